"""Probe: the vector block's Dense(8192 -> 1440) through the 1x1 conv entry points (tcgen05) vs the SIMT dense kernels."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_rir_b200 import _lib as L

B, Kd, N = 64, 8192, 1440


def timeit(fn, inner=5, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


x = torch.randn(B, Kd, device="cuda").to(torch.bfloat16)
w = (torch.randn(Kd, N, device="cuda") / 90).to(torch.bfloat16)
dy = torch.randn(B, N, device="cuda").to(torch.bfloat16)
w_ck = w.reshape(1, Kd, N).contiguous()
w_kc = w.t().reshape(1, N, Kd).contiguous()
for (n, h, ww) in ((64, 1, 1), (1, 8, 8), (1, 1, 64)):
    d = L.ConvDesc(n, h, ww, Kd, N, 1, 1, 1, 0, 0, h, ww, Kd, 0, N, 0, L.BF16, L.BF16, L.IMPL_TC, 0, 0)
    y = torch.zeros(B, N, device="cuda", dtype=torch.bfloat16)
    dx = torch.zeros(B, Kd, device="cuda", dtype=torch.bfloat16)
    dw = torch.zeros(Kd, N, device="cuda")
    for name, fn, ref, out in (
        ("fprop", lambda: L.call("conv2d_fprop", C.byref(d), x.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), None, y.data_ptr(), None), lambda: x.float() @ w.float(), y),
        ("dgrad", lambda: L.call("conv2d_dgrad", C.byref(d), dy.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), None, dx.data_ptr(), None), lambda: dy.float() @ w.float().t(), dx),
        ("wgrad", lambda: L.call("conv2d_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr()), lambda: x.float().t() @ dy.float(), dw),
    ):
        try:
            t = timeit(fn)
            r = ref()
            err = float((out.float() - r).norm() / r.norm())
            print(f"({n},{h},{ww}) {name}: {t*1e3:7.1f} us  rel err {err:.2e}")
        except Exception as e:
            print(f"({n},{h},{ww}) {name}: {str(e)[:150]}")
