"""Turns the long-format CSV of `ncu --metrics ... --csv` (tools/ncu_step.py run) into a per-launch table and a
per-kernel-family summary: time, DRAM bytes, achieved GB/s, tensor-pipe utilisation.
usage: python tools/ncu_step_table.py gpurun_out/step_metrics_r1.csv [gpurun_out/step_calls.json profiles/r01_traffic.json]
           > profiles/r01_step_kernels.txt
With the call list written by tools/ncu_step.py it also writes, per bench.py call family (conv2d_fprop.tcgen05, ...),
the launches, time and DRAM bytes per call: the `traffic` figure of bench.py's roofline object."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
i0 = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
launch = collections.OrderedDict()
for r in rows[i0 + 1:]:
    if len(r) < 15:
        continue
    d = launch.setdefault(int(r[0]), {"name": r[4], "grid": r[8], "block": r[7]})
    val = r[14].replace(",", "")
    try:
        val = float(val)
    except ValueError:
        pass
    if r[12] == "gpu__time_duration.sum" and r[13] == "ns":
        val = val / 1e3
    if r[12] == "gpu__time_duration.sum" and r[13] == "ms":
        val = val * 1e3
    d[r[12]] = val


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    n = n.replace("urir::", "")
    return n[:58]


tot = sum(d["gpu__time_duration.sum"] for d in launch.values())
print(f"# one eager train step, B = 64: {len(launch)} launches, {tot:.0f} us under ncu (serialised, cold cache: compare shares)")
print(f"{'id':>4} {'kernel':58s} {'grid':>14} {'us':>8} {'dramR MB':>9} {'dramW MB':>9} {'GB/s':>7} {'tensor%':>7} {'sm%':>6}")
fam = collections.OrderedDict()
for i, d in launch.items():
    t = d["gpu__time_duration.sum"]
    rd, wr = d.get("dram__bytes_read.sum", 0) / 1e6, d.get("dram__bytes_write.sum", 0) / 1e6
    tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0)
    smp = d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", 0)
    print(f"{i:4d} {short(d['name']):58s} {d['grid']:>14} {t:8.1f} {rd:9.1f} {wr:9.1f} {(rd + wr) / t * 1e3:7.0f} {tp:7.1f} {smp:6.1f}")
    f = fam.setdefault(short(d["name"]), [0, 0.0, 0.0, 0.0, 0.0])
    f[0] += 1; f[1] += t; f[2] += rd; f[3] += wr; f[4] += tp * t
print()
print(f"{'kernel family':58s} {'n':>4} {'us':>9} {'share':>6} {'dramR MB':>9} {'dramW MB':>9} {'GB/s':>7} {'tensor%':>7}")
for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:58s} {f[0]:4d} {f[1]:9.1f} {100 * f[1] / tot:5.1f}% {f[2]:9.1f} {f[3]:9.1f} {(f[2] + f[3]) / f[1] * 1e3:7.0f} {f[4] / f[1]:7.1f}")

if len(sys.argv) > 3:
    import json
    calls = json.load(open(sys.argv[2]))
    ours = [d for d in launch.values() if "urir::" in d["name"]]
    need = sum(c["kernels"] for c in calls)
    out = {"note": "per C-ABI call of one eager train step (B = 64) under ncu: kernels, us, DRAM read+write bytes",
           "aligned": need == len(ours), "families": {}}
    if need == len(ours):
        i = 0
        for c in calls:
            ks = ours[i:i + c["kernels"]]; i += c["kernels"]
            # bench.py's family key: the kernel family that served a convolution call, the call name otherwise
            key = ("conv." + c["family"]) if c.get("family") else c["name"] + ((".tcgen05" if c.get("tc") else ".simt") if c["name"].startswith("conv2d") else "")
            f = out["families"].setdefault(key, {"calls": 0, "kernels": 0, "us": 0.0, "dram_bytes": 0.0})
            f["calls"] += 1; f["kernels"] += len(ks)
            f["us"] += sum(k["gpu__time_duration.sum"] for k in ks)
            f["dram_bytes"] += sum(k.get("dram__bytes_read.sum", 0) + k.get("dram__bytes_write.sum", 0) for k in ks)
        for f in out["families"].values():
            f["dram_bytes_per_call"] = f["dram_bytes"] / max(f["calls"], 1)
            f["dram_bytes_per_launch"] = f["dram_bytes"] / max(f["kernels"], 1)
    else:
        out["error"] = f"call list expects {need} liburir kernels, ncu saw {len(ours)}"
    json.dump(out, open(sys.argv[3], "w"), indent=1)
