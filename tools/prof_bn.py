"""CUDA-event timing of the BatchNorm kernels at bench shapes (B=64): achieved HBM GB/s per launch.
usage: python tools/prof_bn.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unet_rir_b200 import _lib as L

SHAPES = [(64 * 144 * 160, 32), (64 * 72 * 80, 64), (64 * 36 * 40, 128), (64 * 18 * 20, 256), (64 * 9 * 10, 512)]


def timeit(fn, inner=10, reps=3):
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


def main():
    a = torch.randn(4096, 4096, device="cuda", dtype=torch.bfloat16)
    for _ in range(50):
        a @ a
    for npix, C in SHAPES:
        x = torch.randn(npix, C, device="cuda").to(torch.bfloat16)
        dy = torch.randn(npix, C, device="cuda").to(torch.bfloat16)
        y = torch.empty_like(x)
        ss = torch.randn(2 * C, device="cuda"); mr = torch.rand(2 * C, device="cuda") + 0.5
        gamma = torch.ones(C, device="cuda"); sums = torch.zeros(2 * C, device="cuda")
        dg, db, dbias = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        nb = npix * C * 2
        t = timeit(lambda: L.call("bn_relu_fwd", x.data_ptr(), C, 0, ss.data_ptr(), y.data_ptr(), C, 0, npix, C, 1))
        print(f"C={C:4d} npix={npix:8d} fwd    {t*1e3:7.1f} us {2*nb/t/1e6:7.0f} GB/s")
        t = timeit(lambda: L.call("bn_relu_bwd_reduce", dy.data_ptr(), C, 0, x.data_ptr(), C, 0, ss.data_ptr(), mr.data_ptr(), sums.data_ptr(), npix, C, 0))
        print(f"C={C:4d} npix={npix:8d} reduce {t*1e3:7.1f} us {2*nb/t/1e6:7.0f} GB/s")
        t = timeit(lambda: L.call("bn_relu_bwd_apply", dy.data_ptr(), C, 0, x.data_ptr(), C, 0, ss.data_ptr(), mr.data_ptr(), gamma.data_ptr(), sums.data_ptr(), y.data_ptr(), C, 0, dg.data_ptr(), db.data_ptr(), dbias.data_ptr(), npix, C, 0))
        print(f"C={C:4d} npix={npix:8d} apply  {t*1e3:7.1f} us {3*nb/t/1e6:7.0f} GB/s")


if __name__ == "__main__":
    main()
