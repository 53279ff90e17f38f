"""Runs a handful of conv launches at bench shapes (B=64) for ncu / CUDA-event timing.
usage: python tools/prof_conv.py [which ...]   which in {wgrad_full, wgrad_mid, fprop_full, fprop_deep, dgrad_s2, head_wgrad}"""
import ctypes as C
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_rir_b200 import _lib as L

CASES = {
    "wgrad_full": ("wgrad", 64, 144, 160, 32, 32, 3, 1),
    "wgrad_mid": ("wgrad", 64, 36, 40, 128, 128, 3, 1),
    "fprop_full": ("fprop", 64, 144, 160, 32, 32, 3, 1),
    "fprop_full64": ("fprop", 64, 144, 160, 64, 32, 3, 1),
    "fprop_deep": ("fprop", 64, 9, 10, 512, 512, 3, 1),
    "fprop_mid": ("fprop", 64, 36, 40, 256, 128, 3, 1),
    "dgrad_s2": ("dgrad", 64, 144, 160, 32, 64, 3, 2),
    "head_wgrad": ("wgrad", 64, 144, 160, 32, 2, 6, 1),
    "head_fprop": ("fprop", 64, 144, 160, 32, 2, 6, 1),
    "head_dgrad": ("dgrad", 64, 144, 160, 32, 2, 6, 1),
    "stem_fprop": ("fprop", 64, 144, 160, 2, 32, 3, 1),
    "stem_wgrad": ("wgrad", 64, 144, 160, 2, 32, 3, 1),
    "dgrad_full": ("dgrad", 64, 144, 160, 32, 32, 3, 1),
    "wgrad_half": ("wgrad", 64, 72, 80, 64, 64, 3, 1),
    "fprop_half": ("fprop", 64, 72, 80, 64, 64, 3, 1),
    "dgrad_half": ("dgrad", 64, 72, 80, 64, 64, 3, 1),
    "fprop_d4a": ("fprop", 64, 72, 80, 128, 64, 3, 1),
    "dgrad_d4a": ("dgrad", 64, 72, 80, 128, 64, 3, 1),
    "dgrad_full64": ("dgrad", 64, 144, 160, 64, 32, 3, 1),
    "wgrad_full64": ("wgrad", 64, 144, 160, 64, 32, 3, 1),
    "wgrad_d4a": ("wgrad", 64, 72, 80, 128, 64, 3, 1),
    "dgrad_s2_half": ("dgrad", 64, 72, 80, 64, 128, 3, 2),
    "fprop_l3": ("fprop", 64, 36, 40, 128, 128, 3, 1),
    "dgrad_l3": ("dgrad", 64, 36, 40, 128, 128, 3, 1),
    "fprop_l3a": ("fprop", 64, 36, 40, 256, 128, 3, 1),
    "fprop_l4": ("fprop", 64, 18, 20, 256, 256, 3, 1),
    "wgrad_l3a": ("wgrad", 64, 36, 40, 256, 128, 3, 1),
    "wgrad_l3": ("wgrad", 64, 36, 40, 128, 128, 3, 1),
    "wgrad_l4a": ("wgrad", 64, 18, 20, 512, 256, 3, 1),
    "wgrad_l4": ("wgrad", 64, 18, 20, 256, 256, 3, 1),
    "wgrad_l5": ("wgrad", 64, 9, 10, 512, 512, 3, 1),
    "wgrad_s2_l4": ("wgrad", 64, 18, 20, 256, 512, 3, 2),
    "wgrad_s2_l3": ("wgrad", 64, 36, 40, 128, 256, 3, 2),
    "wgrad_s2_full": ("wgrad", 64, 144, 160, 32, 64, 3, 2),
    "dgrad_l4": ("dgrad", 64, 18, 20, 256, 256, 3, 1),
    "fprop_l4a": ("fprop", 64, 18, 20, 512, 256, 3, 1),
    "dgrad_l4a": ("dgrad", 64, 18, 20, 512, 256, 3, 1),
    "dgrad_l3a": ("dgrad", 64, 36, 40, 256, 128, 3, 1),
    "dgrad_deep": ("dgrad", 64, 9, 10, 512, 512, 3, 1),
    "fprop_s2_l4": ("fprop", 64, 36, 40, 128, 256, 3, 2),
    "fprop_s2_l5": ("fprop", 64, 18, 20, 256, 512, 3, 2),
    "dgrad_s2_l4": ("dgrad", 64, 36, 40, 128, 256, 3, 2),
    "dgrad_s2_l5": ("dgrad", 64, 18, 20, 256, 512, 3, 2),
    "fprop_s2": ("fprop", 64, 144, 160, 32, 64, 3, 2),
}


def run(which, reps=int(os.environ.get("PROF_REPS", "3"))):
    op, N, H, W, Cc, K, k, s = CASES[which]
    P, pt = L.same_pad(H, k, s); Q, pl = L.same_pad(W, k, s)
    y_ld = K
    xt = torch.float32 if Cc == 2 else torch.bfloat16
    yt = torch.float32 if K == 2 else torch.bfloat16
    x = torch.randn(N, H, W, Cc, device="cuda").to(xt)
    dy = torch.randn(N, P, Q, y_ld, device="cuda").to(yt)
    w_ck = torch.randn(k * k, Cc, K, device="cuda").to(torch.bfloat16)
    w_kc = torch.randn(k * k, K, Cc, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(k, k, Cc, K, device="cuda")
    acc = 1 if os.environ.get("PROF_ACC") else 0
    d = L.ConvDesc(N, H, W, Cc, K, k, k, s, pt, pl, P, Q, Cc, 0, y_ld, 0, L.dtype_code(x), L.dtype_code(dy), L.IMPL_AUTO, 0, acc)
    stats = torch.zeros(2 * max(Cc, K), device="cuda") if os.environ.get("PROF_STATS") and Cc > 2 and K > 2 else None
    sp = stats.data_ptr() if stats is not None else None
    ts = []
    if reps > 1:      # bring the SM clock up before timing (short runs otherwise time at idle clocks)
        a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
        t_end = __import__("time").time() + 0.3
        while __import__("time").time() < t_end:
            (a @ a); torch.cuda.synchronize()
    w_up2 = None
    if op == "dgrad" and s == 2 and k == 3 and L.load().urir_conv_path(C.byref(d), 3) == 1:
        w_up2 = torch.randn(4, 4 * Cc, K, device="cuda").to(torch.bfloat16)
    def launch():
        if w_up2 is not None and not os.environ.get("PROF_NO_UP2"):
            L.call("conv2d_dgrad_up2", C.byref(d), dy.data_ptr(), w_up2.data_ptr(), None, x.data_ptr())
        elif op == "wgrad":
            L.call("conv2d_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr())
        elif op == "fprop":
            L.call("conv2d_fprop", C.byref(d), x.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), None, dy.data_ptr(), sp)
        else:
            L.call("conv2d_dgrad_sums" if (sp is not None and os.environ.get("PROF_SUMS")) else "conv2d_dgrad", C.byref(d), dy.data_ptr(),
                   w_ck.data_ptr(), w_kc.data_ptr(), None, x.data_ptr(), sp)
    inner = 10 if reps > 1 else 1      # back-to-back launches per timing: host launch latency must not count
    if reps > 1 and not os.environ.get("PROF_NO_GRAPH"):
        # the C-ABI call costs ~20-25 us of host time (ctypes + two tensor-map encodes + launch): kernels shorter than that
        # are host-bound when launched one by one, so the timed launches are replayed from a CUDA graph like the real step
        launch(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(inner):
                launch()
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g.replay(); e0.record()
            g.replay()
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / inner)
        reps = 0
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launch(); e0.record()
        for _ in range(inner):
            launch()
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / inner)
    fl = 2.0 * N * P * Q * K * Cc * k * k
    print(f"{which:14s} {min(ts):8.3f} ms  {fl / (min(ts) * 1e-3) / 1e12:8.1f} TF/s", flush=True)


if __name__ == "__main__":
    for w in (sys.argv[1:] or list(CASES)):
        run(w)
