"""ctypes binding of liburir.so (the C-ABI declared in include/urir.h).

The product path has no CPU fallback: if the library is missing, or a call fails, this raises.
PyTorch is used only for device memory and streams; every `urir_*` call is enqueued on
`torch.cuda.current_stream()` so the calls are CUDA-graph capturable.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liburir.so")

F32, BF16 = 0, 1
IMPL_AUTO, IMPL_SIMT, IMPL_TC, IMPL_HALO, IMPL_DEEP = 0, 1, 2, 3, 4
ACT_NONE, ACT_SIGMOID, ACT_RELU = 0, 1, 2


class UrirError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "N", "H", "W", "C", "K", "R", "S", "stride", "pad_top", "pad_left", "P", "Q",
        "x_ld", "x_coff", "y_ld", "y_coff", "x_dtype", "y_dtype", "impl", "act", "accumulate")]


class StftDesc(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "n_fft", "win_length", "hop_length", "n_samples", "n_bins", "n_frames", "H_pad", "W_pad",
        "pad_mode", "remove_mean", "normalized")]


_vp, _i, _ll, _f, _d, _u64 = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_double, C.c_uint64
_PROTOS = {
    "urir_version": (C.c_int, []),
    "urir_last_error": (C.c_char_p, []),
    "urir_launch_count": (C.c_longlong, [_i]),
    "urir_family_calls": (C.c_longlong, [_i]),
    "urir_family_name": (C.c_char_p, [_i]),
    "urir_conv2d_fprop": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "urir_conv2d_dgrad": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "urir_conv2d_dgrad_sums": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "urir_conv2d_wgrad": (_i, [C.POINTER(ConvDesc), _vp, _vp, _vp, _vp]),
    "urir_conv_path": (_i, [C.POINTER(ConvDesc), _i]),
    "urir_set_pdl": (_i, [_i]),
    "urir_set_deterministic": (_i, [_i]),
    "urir_l2_reg_batched": (_i, [_vp, _i, _f, _vp, _vp]),
    "urir_weight_prep": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "urir_weight_prep_up2": (_i, [_vp, _vp, _i, _i, _vp]),
    "urir_conv2d_dgrad_up2": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "urir_weight_prep_batched": (_i, [_vp, _i, _vp]),
    "urir_weight_fold_bn_batched": (_i, [_vp, _i, _vp]),
    "urir_channel_sum": (_i, [_vp, _i, _ll, _i, _i, _i, _vp, _vp]),
    "urir_bn_finalize": (_i, [_vp, _d, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp, _i, _vp]),
    "urir_bn_relu_fwd": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _ll, _i, _i, _vp]),
    "urir_bn_relu_fwd_train": (_i, [_vp, _i, _i, _vp, _d, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _vp,
                                    _vp, _i, _i, _ll, _i, _vp]),
    "urir_bn_relu_bwd_reduce": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "urir_bn_relu_bwd_apply": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp,
                                    _vp, _ll, _i, _i, _vp]),
    "urir_embedding_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "urir_embedding_bwd": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "urir_dense_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "urir_dense_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "urir_dropout_mask": (_i, [_vp, _ll, _f, _u64, _vp, _vp]),
    "urir_ampphase_loss": (_i, [_vp, _vp, _ll, _f, _f, _i, _vp, _vp, _vp, _i, _vp]),
    "urir_mse2_loss": (_i, [_vp, _vp, _ll, _f, _i, _vp, _vp, _vp]),
    "urir_adam": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _f, _f, _f, _vp]),
    "urir_sgd": (_i, [_vp, _vp, _ll, _vp, _vp]),
    "urir_nadam": (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _f, _f, _f, _vp]),
    "urir_lamb": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _f, _f, _f, _f, _vp]),
    "urir_step_increment": (_i, [_vp, _vp]),
    "urir_axpy": (_i, [_vp, _vp, _f, _ll, _vp]),
    "urir_sumsq": (_i, [_vp, _ll, _f, _vp, _i, _vp]),
    "urir_add_bf16": (_i, [_vp, _vp, _vp, _ll, _vp]),
    "urir_add_bf16_strided": (_i, [_vp, _i, _i, _vp, _i, _i, _vp, _i, _i, _ll, _i, _vp]),
    "urir_cast_f32_to_bf16": (_i, [_vp, _vp, _ll, _vp]),
    "urir_cast_pad_bf16": (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    "urir_stft_ampphase": (_i, [_vp, _i, C.POINTER(StftDesc), _vp, _vp]),
    "urir_istft_from_ampphase": (_i, [_vp, _i, C.POINTER(StftDesc), _vp, _vp]),
}
EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib = None


def load():
    """Loads liburir.so; raises UrirError (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise UrirError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` "
                "(or unet-rir_b200/csrc/build.sh). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def last_error() -> str:
    return load().urir_last_error().decode("utf-8", "replace")


def check(rc: int, what: str = ""):
    if rc != 0:
        raise UrirError(f"{what or 'urir call'} failed (rc={rc}): {last_error()}")


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int | None:
    if t is None:
        return None
    return t.data_ptr()


def set_deterministic(enabled: bool) -> bool:
    """Fixed-order cross-CTA reductions (bit-identical reruns; slower). Returns the previous setting. CUDA graphs
    captured before the switch keep the mode they were captured in."""
    return bool(load().urir_set_deterministic(1 if enabled else 0))


def launch_count(kind: int = 0) -> int:
    return int(load().urir_launch_count(kind))


FAMILIES = ("simt", "igemm", "halo", "halo_s2_fprop", "halo_up2", "thin_gemm", "head_fprop", "wgrad_tc", "wgrad_halo",
            "wgrad_halo_s2", "thin_wgrad", "deep")


def family_calls() -> dict:
    """name -> number of successful urir_conv2d_* calls served by that kernel family since the library was loaded."""
    lib = load()
    return {lib.urir_family_name(i).decode(): int(lib.urir_family_calls(i)) for i in range(len(FAMILIES))}


_profile = None     # when a list: every call appends (name, conv-desc-or-None, start_event, end_event)


def profile_begin():
    """Starts recording a CUDA-event pair around every urir_* call (bench.py's per-kernel timing)."""
    global _profile
    _profile = []


def profile_end():
    """-> list of (name, info dict, milliseconds); synchronises."""
    global _profile
    rec, _profile = _profile or [], None
    torch.cuda.synchronize()
    return [(n, info, s.elapsed_time(e)) for n, info, s, e in rec]


def _conv_info(name, args):
    d = args[0]._obj if hasattr(args[0], "_obj") else None
    if not isinstance(d, ConvDesc):
        return {}
    op = {"conv2d_fprop": 0, "conv2d_dgrad": 1, "conv2d_dgrad_sums": 1, "conv2d_wgrad": 2, "conv2d_dgrad_up2": 3}[name]
    tc = int(load().urir_conv_path(C.byref(d), op))
    xs, ys = (4 if d.x_dtype == F32 else 2), (4 if d.y_dtype == F32 else 2)
    xb, yb = d.N * d.H * d.W * d.C * xs, d.N * d.P * d.Q * d.K * ys
    wb = d.R * d.S * d.C * d.K * (4 if op == 2 else 2)
    # algorithmic bytes: every operand once (+ the destination once more when the call accumulates into it)
    extra = 0
    if d.accumulate and op != 2:
        extra = yb if op == 0 else xb
    return dict(tc=tc, N=d.N, H=d.H, W=d.W, C=d.C, K=d.K, R=d.R, S=d.S, stride=d.stride, P=d.P, Q=d.Q,
                x_ld=d.x_ld, y_ld=d.y_ld, x_dtype=d.x_dtype, y_dtype=d.y_dtype, impl=d.impl,
                flops=2.0 * d.N * d.P * d.Q * d.K * d.C * d.R * d.S, bytes=float(xb + yb + wb + extra))


def _elementwise_info(name, args):
    """Algorithmic HBM bytes of the bandwidth-bound calls (bf16 activations read / written once per pass)."""
    try:
        if name == "bn_relu_fwd_train":      # x ... npix, C are the last two arguments
            return {"bytes": 2.0 * args[-2] * args[-1] * 2}
        if name == "bn_relu_fwd":
            return {"bytes": 2.0 * args[7] * args[8] * 2}
        if name == "bn_relu_bwd_reduce":
            return {"bytes": 2.0 * args[9] * args[10] * 2}
        if name == "bn_relu_bwd_apply":
            return {"bytes": 3.0 * args[16] * args[17] * 2}
        if name == "ampphase_loss":          # y_true, y_pred fp32 read, grad fp32 written (when asked for)
            return {"bytes": float(args[2]) * 2 * 4 * (3 if args[7] else 2)}
        if name in ("adam", "nadam"):        # p, m, v read + written, g read
            return {"bytes": 7.0 * 4 * args[4]}
        if name == "sgd":
            return {"bytes": 3.0 * 4 * args[2]}
    except Exception:
        pass
    return {}


def call(name: str, *args):
    """Calls `urir_<name>`; the last argument (stream) is appended automatically."""
    fn = getattr(load(), "urir_" + name)
    if _profile is None:
        check(fn(*args, stream()), name)
        return
    conv = name.startswith("conv2d")
    before = family_calls() if conv else None
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    check(fn(*args, stream()), name)
    e.record()
    info = _conv_info(name, args) if conv else _elementwise_info(name, args)
    if conv:
        after = family_calls()
        fams = [k for k in after if after[k] != before[k]]
        info["family"] = fams[0] if fams else "?"
    _profile.append((name, info, s, e))


def same_pad(in_size: int, k: int, s: int):
    """TF SAME geometry: out = ceil(in/s), total = max((out-1)s + k - in, 0), before = total//2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    return out, total // 2


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise UrirError(f"unsupported dtype {t.dtype}")
