#!/usr/bin/env python
"""Round-2 GPU probes (run under gpurun; results land in gpurun_out/r02_probe.json):
  1. run-to-run gradient spread of identical train steps with the fixed-order reductions off and on (B = 16),
     and what the deterministic mode costs in step time;
  2. cudaLimitMaxL2FetchGranularity 32 / 64 / 128 against kernels that read a 64-byte channel slice of a 128-byte-pitch
     buffer (the level-1 skip-concat halves): BatchNorm backward reduce and the stride-2 halo fprop.
"""
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch

from unet_rir_b200 import _lib as L
from unet_rir_b200 import plan as PL
from unet_rir_b200.engine import UNetEngine


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-300))


def one_step(params, x, y, emb, mask, B):
    eng = UNetEngine(kernels=3)
    eng.load_state_dict(params)
    eng.forward(x, emb, training=True, dropout_mask=mask)
    n = B * 144 * 160
    eng.loss_and_grad(y, 1.0 / n, 1.0 / n)
    eng.backward(eng._buffers(B)["g_out"])
    torch.cuda.synchronize()
    return eng


def spread(B=16):
    g = torch.Generator().manual_seed(3)
    params = PL.keras_init(PL.layer_plan(kernels=3), seed=500)
    x = torch.rand(B, 144, 160, 2, generator=g).cuda(); y = torch.rand(B, 144, 160, 2, generator=g).cuda()
    emb = torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32).cuda()
    mask = ((torch.rand(B, 1440, generator=g) > 0.3).float() / 0.7).cuda()
    out = {}
    for det in (False, True):
        L.set_deterministic(det)
        runs = [one_step(params, x, y, emb, mask, B) for _ in range(3)]
        per = {}
        for n in ("enc1.down.w", "enc3.blk.c1.w", "enc5.blk.c1.w", "vec.dense.w", "dec2.fuse.w", "dec5.blk.c1.w", "head.w"):
            per[n] = max(rel_l2(runs[i].grad[n], runs[0].grad[n]) for i in (1, 2))
        flat = max(rel_l2(runs[i].G, runs[0].G) for i in (1, 2))
        outd = max(float((runs[i]._buffers(B)["out"] - runs[0]._buffers(B)["out"]).abs().max()) for i in (1, 2))
        # eager step time in this mode
        eng = runs[0]
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(3):
            eng._forward_body(B, True, injected_mask=True)
            eng.loss_and_grad(y, 1.0 / (B * 23040), 1.0 / (B * 23040))
            eng._backward_body(B)
        torch.cuda.synchronize()
        out["deterministic" if det else "atomics"] = {"rel_l2_flat_gradient_run_to_run": flat, "per_tensor": per,
                                                      "max_abs_output_diff": outd,
                                                      "bit_identical": bool(all(torch.equal(runs[i].G, runs[0].G) for i in (1, 2))),
                                                      "eager_fwd_bwd_ms": (time.perf_counter() - t0) / 3 * 1e3}
        del runs
        torch.cuda.empty_cache()
    L.set_deterministic(False)
    return out


def l2_fetch_granularity():
    import urir_testutil as U
    rt = None
    for name in ("libcudart.so.12", "libcudart.so"):
        try:
            rt = ctypes.CDLL(name); break
        except OSError:
            pass
    if rt is None:
        import glob
        cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "lib", "libcudart*.so*")) + \
            glob.glob("/usr/local/cuda/lib64/libcudart.so*")
        rt = ctypes.CDLL(cands[0])
    LIMIT = 0x05                                    # cudaLimitMaxL2FetchGranularity
    res = {}
    B, Hh, Ww = 64, 144, 160
    npix = B * Hh * Ww
    cat = torch.randn(B, Hh, Ww, 64, device="cuda").to(torch.bfloat16)       # [e1 | up1]: two 64-byte halves per pixel
    xfull = torch.randn(B, Hh, Ww, 32, device="cuda").to(torch.bfloat16)
    ss = torch.rand(64, device="cuda"); mr = torch.rand(64, device="cuda") + 0.5
    sums = torch.zeros(64, device="cuda")
    w = torch.randn(3, 3, 32, 64, device="cuda") * 0.05
    w_ck, w_kc = U.prep_weights(w)
    y = torch.empty(B, Hh // 2, Ww // 2, 64, device="cuda", dtype=torch.bfloat16)
    d_slice = U.conv_desc(B, Hh, Ww, 32, 64, 3, 2, x_ld=64, x_coff=0)
    d_dense = U.conv_desc(B, Hh, Ww, 32, 64, 3, 2)

    def t(fn, it=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(it):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / it * 1e3

    for gran in (128, 64, 32):
        rc = rt.cudaDeviceSetLimit(ctypes.c_int(LIMIT), ctypes.c_size_t(gran))
        val = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(val), ctypes.c_int(LIMIT))
        r = {"set_rc": int(rc), "limit_now": int(val.value)}
        r["bn_bwd_reduce_slice_us"] = t(lambda: L.call("bn_relu_bwd_reduce", cat.data_ptr(), 64, 0, xfull.data_ptr(), 32, 0,
                                                         ss.data_ptr(), mr.data_ptr(), sums.data_ptr(), npix, 32, 0))
        r["bn_bwd_reduce_dense_us"] = t(lambda: L.call("bn_relu_bwd_reduce", xfull.data_ptr(), 32, 0, xfull.data_ptr(), 32, 0,
                                                         ss.data_ptr(), mr.data_ptr(), sums.data_ptr(), npix, 32, 0))
        r["halo_s2_fprop_slice_us"] = t(lambda: U.run_fprop(d_slice, cat, w_ck, w_kc, None, y))
        r["halo_s2_fprop_dense_us"] = t(lambda: U.run_fprop(d_dense, xfull, w_ck, w_kc, None, y))
        res[str(gran)] = r
    rt.cudaDeviceSetLimit(ctypes.c_int(LIMIT), ctypes.c_size_t(128))
    return res


if __name__ == "__main__":
    out = {}
    for name, fn in (("spread", spread), ("l2_fetch_granularity", l2_fetch_granularity)):
        try:
            out[name] = fn()
        except Exception as ex:
            out[name] = {"error": repr(ex)}
        print(name, json.dumps(out[name]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_probe.json"), "w"), indent=1)
