"""Achieved HBM GB/s of the STFT / iSTFT kernels through the C-ABI (algorithmic bytes: 38.4 kB of waveform + 184.3 kB of
padded spectrogram per sample). Buffers rotate over > 126 MB so that no launch finds its input in L2.
usage: python tools/prof_stft.py [B ...]"""
import ctypes as C
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_rir_b200 import _lib as L

pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = pk.get("hbm_gbs", 6548.2)
Bs = [int(a) for a in sys.argv[1:]] or [4, 64, 256, 1024]
d = L.StftDesc(256, 128, 64, 9600, 129, 151, 144, 160, 0, 1, 1)
L.load()
for B in Bs:
    per = B * (9600 * 4 + 144 * 160 * 2 * 4)
    nset = max(2, -(-400_000_000 // per)) if B >= 64 else 2
    nset = min(nset, 64)
    g = torch.Generator(device="cuda").manual_seed(0)
    wavs = [torch.randn(B, 9600, device="cuda", generator=g) * 0.1 for _ in range(nset)]
    specs = [torch.empty(B, 144, 160, 2, device="cuda") for _ in range(nset)]
    outs = [torch.empty(B, 9600, device="cuda") for _ in range(nset)]
    for name, fn in (("stft_ampphase", lambda i: L.call("stft_ampphase", wavs[i].data_ptr(), B, C.byref(d), specs[i].data_ptr())),
                     ("istft_from_ampphase", lambda i: L.call("istft_from_ampphase", specs[i].data_ptr(), B, C.byref(d), outs[i].data_ptr()))):
        for i in range(nset):
            fn(i)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(nset):
                fn(i)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / nset)
        gbs = per / best / 1e6
        print(f"{name:20s} B={B:5d}: {best*1e3:8.1f} us  {gbs:7.0f} GB/s = {gbs/HBM:5.3f} of the measured HBM peak ({HBM:.0f} GB/s); {nset} rotating buffer sets")
