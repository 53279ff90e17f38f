"""Helpers shared by the GPU parity tests: raw C-ABI conv calls and error metrics."""
import ctypes as C

import torch

from unet_rir_b200 import _lib as L


def rel_l2(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def bf16_round(t):
    return t.to(torch.bfloat16).to(torch.float32)


def prep_weights(w_hwio_cuda):
    """fp32 HWIO (cuda) -> (w_ck, w_kc) bf16 via urir_weight_prep."""
    kh, kw, c, k = w_hwio_cuda.shape
    w_ck = torch.empty(kh * kw, c, k, dtype=torch.bfloat16, device="cuda")
    w_kc = torch.empty(kh * kw, k, c, dtype=torch.bfloat16, device="cuda")
    src = w_hwio_cuda.contiguous()
    L.call("weight_prep", src.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), kh * kw, c, k)
    torch.cuda.current_stream().synchronize()      # `src` may be a temporary of the caller
    return w_ck, w_kc


def conv_desc(N, H, W, Cc, K, k, stride, x_ld=None, x_coff=0, y_ld=None, y_coff=0, x_dtype=L.BF16, y_dtype=L.BF16,
              impl=L.IMPL_AUTO, act=0, accumulate=0):
    P, pt = L.same_pad(H, k, stride)
    Q, pl = L.same_pad(W, k, stride)
    return L.ConvDesc(N, H, W, Cc, K, k, k, stride, pt, pl, P, Q, x_ld or Cc, x_coff, y_ld or K, y_coff, x_dtype,
                      y_dtype, impl, act, accumulate)


def run_fprop(d, x, w_ck, w_kc, bias, y, stats=None):
    L.call("conv2d_fprop", C.byref(d), x.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), L.ptr(bias), y.data_ptr(),
           L.ptr(stats))


def run_dgrad(d, dy, w_ck, w_kc, bias, dx, stats=None):
    L.call("conv2d_dgrad", C.byref(d), dy.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), L.ptr(bias), dx.data_ptr(),
           L.ptr(stats))


def run_dgrad_sums(d, dy, w_ck, w_kc, bias, dx, stats):
    L.call("conv2d_dgrad_sums", C.byref(d), dy.data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), L.ptr(bias), dx.data_ptr(),
           L.ptr(stats))


def run_wgrad(d, x, dy, dw):
    L.call("conv2d_wgrad", C.byref(d), x.data_ptr(), dy.data_ptr(), dw.data_ptr())
