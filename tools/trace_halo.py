"""Timeline of CTA 0 of conv_halo_kernel (clock64 stamps): python tools/trace_halo.py [case] [stats]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
trace = torch.zeros(128 * 8 + 2 * 148, dtype=torch.int64, device="cuda")
os.environ["URIR_HALO_TRACE"] = hex(trace.data_ptr())
if len(sys.argv) > 2: os.environ["PROF_STATS"] = "1"
os.environ["PROF_REPS"] = "2"
import prof_conv
prof_conv.run(sys.argv[1] if len(sys.argv) > 1 else "fprop_full")
per = trace.cpu()[1024:].view(148, 2)
print("per-CTA cycles min/mean/max:", int(per[:, 0].min()), float(per[:, 0].float().mean()), int(per[:, 0].max()), " ns min/mean/max:", int(per[:, 1].min()), float(per[:, 1].float().mean()), int(per[:, 1].max()), " MHz:", float((per[:, 0].float() / per[:, 1].float()).mean() * 1e3))
t = trace.cpu()[:1024].view(128, 8)
t0 = int(t[0, 0])
print("tile: mma[top, tempty-ok, full-ok, issued] epi[tfull-ok, done] prod[before-empty, issued]   (cycles since first stamp)")
for i in list(range(0, 12)) + list(range(40, 48)):
    print(i, [int(v) - t0 for v in t[i]])
print("CTA0 span cycles first stamp -> last epilogue done:", int(t[:78, 5].max()) - t0, " tiles 60..77 per-tile:", (t[77, 3] - t[60, 3]).item() / 17)
d = (t[60, 3] - t[20, 3]).item() / 40
print("steady-state cycles per tile (MMA warp):", d)
for name, a, b in (("mma wait tempty", 0, 1), ("mma wait full", 1, 2), ("mma issue", 2, 3), ("epi work", 4, 5), ("prod wait+issue", 6, 7)):
    print(name, float((t[20:60, b] - t[20:60, a]).float().mean()))
print("epi: tfull-ok -> first tmem ld done", float((t[20:60, 6] - t[20:60, 4]).float().mean()), " -> all blocks processed/stores issued", float((t[20:60, 7] - t[20:60, 6]).float().mean()), " -> arrive", float((t[20:60, 5] - t[20:60, 7]).float().mean()))
print("epi loop period (same group)", float((t[22:60, 4] - t[20:58, 4]).float().mean()), " done -> next tfull-ok", float((t[22:60, 4] - t[20:58, 5]).float().mean()))
print("epi wait (tfull after previous done)", float((t[21:60, 4] - t[20:59, 5]).float().mean()))
