"""Per-launch roofline of one train step from bench.py's per-kernel table (gpurun_out/kernel_table_n1_b64.json):
algorithmic FLOPs and bytes (each activation tensor read once / written once, bf16 = 2 B, fp32 = 4 B) against the
measured peaks in MEASURED_PEAKS.json -> which roof bounds the launch and the fraction of THAT roof it reaches.
usage: python tools/layer_roofline.py gpurun_out/kernel_table_n1_b64.json > profiles/r01_layer_roofline.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM, TF = pk.get("hbm_gbs", 6548.2) * 1e9, pk.get("bf16_tflops_sustained", 1398.6) * 1e12
d = json.load(open(sys.argv[1]))
print(f"# peaks: HBM {HBM/1e9:.0f} GB/s, bf16 {TF/1e12:.0f} TFLOP/s (sustained); times = CUDA events around each C-ABI call of one eager step (B = 64)")
print(f"{'call':18s} {'shape':34s} {'us':>7s} {'GFLOP':>7s} {'MB':>7s} {'TF/s':>7s} {'GB/s':>7s}  bound   frac")
tot_t = tot_ideal = 0.0
for name, i, ms in d["launches"]:
    if not name.startswith("conv2d"):
        continue
    bx = 4 if i["x_dtype"] == 0 else 2
    by = 4 if i["y_dtype"] == 0 else 2
    X = i["N"] * i["H"] * i["W"] * i["C"] * bx
    Y = i["N"] * i["P"] * i["Q"] * i["K"] * by
    byts = X + Y
    fl = i["flops"]
    t = ms * 1e-3
    t_t, t_m = fl / TF, byts / HBM
    bound = "tensor" if t_t > t_m else "hbm"
    ideal = max(t_t, t_m)
    tot_t += t; tot_ideal += ideal
    shape = f"{i['H']}x{i['W']} C{i['C']} K{i['K']} k{i['R']} s{i['stride']}"
    print(f"{name[7:]:18s} {shape:34s} {t*1e6:7.1f} {fl/1e9:7.1f} {byts/1e6:7.1f} {fl/t/1e12:7.1f} {byts/t/1e9:7.0f}  {bound:6s} {ideal/t:5.2f}")
print(f"# conv launches: {tot_t*1e6:.0f} us measured (eager, includes launch gaps), {tot_ideal*1e6:.0f} us at the binding roof of each launch -> {tot_ideal/tot_t:.2f} of roofline overall")
