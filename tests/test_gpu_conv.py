"""GPU parity: conv fprop / dgrad / wgrad through the C-ABI (both the tcgen05 implicit-GEMM path and
the CUDA-core path) against the CPU oracle's TF-SAME convolutions on the same seeded bf16-rounded
inputs. Tolerances: bf16 outputs rel-L2 <= 4e-3 (one bf16 rounding of an fp32-accumulated result);
fp32 outputs (wgrad, stats) rel-L2 <= 2e-4 (summation order only)."""
import ctypes as C

import pytest
import torch

from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L

pytestmark = pytest.mark.gpu

BF16_TOL = 4e-3
F32_TOL = 2e-4

# (N, H, W, C, K, k, stride): every (k, stride, Cin, Cout) class of the canonical U-Net
# at reduced spatial size, plus the kernels=6 default and ragged tiles.
CONV_CASES = [
    (2, 16, 32, 32, 32, 3, 1),      # E1b / D5b class (BLOCK_K=32 swizzle-64 path)
    (2, 16, 32, 64, 32, 3, 1),      # D5a class
    (3, 12, 20, 64, 64, 3, 1),      # ragged box, batch 3
    (2, 12, 20, 128, 64, 3, 1),     # D4a class
    (2, 18, 20, 128, 128, 3, 1),    # E3b class
    (2, 9, 10, 256, 128, 3, 1),     # D3a class, tile spanning images
    (4, 9, 10, 512, 512, 3, 1),     # E5b class
    (2, 16, 32, 32, 64, 3, 2),      # E2a class (stride 2, Cin 32)
    (2, 24, 40, 64, 128, 3, 2),     # E3a class
    (2, 18, 20, 256, 512, 3, 2),    # E5a class
    (2, 16, 16, 64, 64, 6, 1),      # kernels=6, stride 1 (pad 2,3)
    (2, 16, 16, 64, 128, 6, 2),     # kernels=6, stride 2 (pad 2,2)
]


def _inputs(N, H, W, Cc, K, k, stride, seed=0):
    g = torch.Generator().manual_seed(seed)
    P, _ = L.same_pad(H, k, stride)
    Q, _ = L.same_pad(W, k, stride)
    x = U.bf16_round(torch.randn(N, H, W, Cc, generator=g))
    w = U.bf16_round(torch.randn(k, k, Cc, K, generator=g) / (k * (Cc ** 0.5)))
    bias = torch.randn(K, generator=g)
    dy = U.bf16_round(torch.randn(N, P, Q, K, generator=g))
    return x, w, bias, dy, P, Q


def _oracle_fprop(x, w, bias, stride):
    return O.conv2d_same(x.permute(0, 3, 1, 2), w, bias, stride).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("impl", [L.IMPL_SIMT, L.IMPL_TC], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[str(c) for c in CONV_CASES])
def test_fprop_dgrad_wgrad(case, impl):
    N, H, W, Cc, K, k, stride = case
    x, w, bias, dy, P, Q = _inputs(*case)
    xg, dyg = x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16)
    w_ck, w_kc = U.prep_weights(w.cuda())
    bg = bias.cuda()
    d = U.conv_desc(N, H, W, Cc, K, k, stride, impl=impl)

    # ---- fprop (+ fused per-channel statistics)
    y = torch.empty(N, P, Q, K, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * K, device="cuda")
    U.run_fprop(d, xg, w_ck, w_kc, bg, y, stats)
    ref = _oracle_fprop(x, w, bias, stride)
    assert U.rel_l2(y.float(), ref) < BF16_TOL
    assert U.rel_l2(stats[:K], ref.sum(dim=(0, 1, 2))) < 5e-3 or U.max_abs(stats[:K], ref.sum(dim=(0, 1, 2))) < 1e-2
    assert U.rel_l2(stats[K:], (ref ** 2).sum(dim=(0, 1, 2))) < F32_TOL * 10

    # ---- dgrad (the oracle's autograd of the same conv) with bias (= Conv2DTranspose forward)
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    yr = _oracle_fprop(xr, wr, None, stride)
    gx, gw = torch.autograd.grad(yr, [xr, wr], dy)
    cbias = torch.randn(Cc, generator=torch.Generator().manual_seed(7))
    dx = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    dstats = torch.zeros(2 * Cc, device="cuda")
    cbg = cbias.cuda()
    U.run_dgrad(d, dyg, w_ck, w_kc, cbg, dx, dstats)
    refdx = gx + cbias
    assert U.rel_l2(dx.float(), refdx) < BF16_TOL
    assert U.max_abs(dstats[:Cc], refdx.sum(dim=(0, 1, 2))) < 2e-3 * float(refdx.abs().sum(dim=(0, 1, 2)).max())

    # ---- dgrad with accumulate
    base = U.bf16_round(torch.randn(N, H, W, Cc, generator=torch.Generator().manual_seed(9)))
    dx2 = base.cuda().to(torch.bfloat16)
    d.accumulate = 1
    U.run_dgrad(d, dyg, w_ck, w_kc, None, dx2, None)
    d.accumulate = 0
    assert U.rel_l2(dx2.float(), gx + base) < BF16_TOL

    # ---- wgrad
    dw = torch.full((k, k, Cc, K), 3.0, device="cuda")      # must be overwritten, not accumulated
    U.run_wgrad(d, xg, dyg, dw)
    assert U.rel_l2(dw, gw) < F32_TOL


def test_conv_transpose_matches_oracle():
    """Conv2DTranspose(k, s=2, SAME) forward == urir_conv2d_dgrad of the strided conv, written into
    the right half of a concat buffer (u_net.py:297-308)."""
    for k in (3, 6):
        g = torch.Generator().manual_seed(k)
        N, Hs, Ws, Cin, Cout = 2, 9, 10, 64, 32
        x = U.bf16_round(torch.randn(N, Hs, Ws, Cin, generator=g))
        w = U.bf16_round(torch.randn(k, k, Cout, Cin, generator=g) / (k * 8))     # Keras (kh,kw,out,in)
        bias = torch.randn(Cout, generator=g)
        ref = O.conv2d_transpose_same(x.permute(0, 3, 1, 2), w, bias, 2).permute(0, 2, 3, 1)
        for impl in (L.IMPL_SIMT, L.IMPL_TC):
            cat = torch.zeros(N, 2 * Hs, 2 * Ws, 2 * Cout, dtype=torch.bfloat16, device="cuda")
            w_ck, w_kc = U.prep_weights(w.cuda())
            d = U.conv_desc(N, 2 * Hs, 2 * Ws, Cout, Cin, k, 2, x_ld=2 * Cout, x_coff=Cout, impl=impl)
            xg, bg = x.cuda().to(torch.bfloat16), bias.cuda()
            U.run_dgrad(d, xg, w_ck, w_kc, bg, cat, None)
            assert U.rel_l2(cat[..., Cout:].float(), ref) < BF16_TOL, (k, impl)
            assert float(cat[..., :Cout].float().abs().max()) == 0.0      # left half untouched


def test_stem_and_head_shapes():
    """The two bandwidth-bound layers that run on CUDA cores by design: fp32 2-channel stem and the
    6x6 32->2 sigmoid head (u_net.py:269-276 with Cin=2; u_net.py:248-249)."""
    g = torch.Generator().manual_seed(3)
    N, H, W = 2, 16, 32
    x = torch.rand(N, H, W, 2, generator=g)
    w = U.bf16_round(torch.randn(3, 3, 2, 32, generator=g) * 0.2)
    b = torch.randn(32, generator=g) * 0.1
    w_ck, w_kc = U.prep_weights(w.cuda())
    y = torch.empty(N, H, W, 32, dtype=torch.bfloat16, device="cuda")
    d = U.conv_desc(N, H, W, 2, 32, 3, 1, x_dtype=L.F32)
    xg, bg = x.cuda(), b.cuda()
    U.run_fprop(d, xg, w_ck, w_kc, bg, y)
    assert U.rel_l2(y.float(), _oracle_fprop(x, w, b, 1)) < BF16_TOL
    d.impl = L.IMPL_SIMT                                   # the CUDA-core stem kernel stays available
    U.run_fprop(d, xg, w_ck, w_kc, bg, y)
    assert U.rel_l2(y.float(), _oracle_fprop(x, w, b, 1)) < BF16_TOL

    xh = U.bf16_round(torch.randn(N, H, W, 32, generator=g))
    wh = U.bf16_round(torch.randn(6, 6, 32, 2, generator=g) * 0.05)
    bh = torch.randn(2, generator=g) * 0.1
    w_ck, w_kc = U.prep_weights(wh.cuda())
    out = torch.empty(N, H, W, 2, device="cuda")
    dh = U.conv_desc(N, H, W, 32, 2, 6, 1, y_dtype=L.F32, act=L.ACT_SIGMOID)
    xhg, bhg = xh.cuda().to(torch.bfloat16), bh.cuda()
    U.run_fprop(dh, xhg, w_ck, w_kc, bhg, out)
    ref = torch.sigmoid(_oracle_fprop(xh, wh, bh, 1))
    assert U.max_abs(out, ref) < 1e-5


def test_small_channel_tensor_core_paths():
    """Channel counts below the 32-wide MMA tile ride on TMA zero-fill + masked epilogues: the 6x6 32->2
    sigmoid head (fprop fp32 out, wgrad from the padded bf16 gradient) and the 1x1 16->512 vector
    projection (u_net.py:248-249, 262)."""
    g = torch.Generator().manual_seed(11)
    N, H, W = 2, 16, 32
    # ---- head fprop on tensor cores
    xh = U.bf16_round(torch.randn(N, H, W, 32, generator=g))
    wh = U.bf16_round(torch.randn(6, 6, 32, 2, generator=g) * 0.05)
    bh = torch.randn(2, generator=g) * 0.1
    w_ck, w_kc = U.prep_weights(wh.cuda())
    xhg, bhg = xh.cuda().to(torch.bfloat16), bh.cuda()
    out = torch.empty(N, H, W, 2, device="cuda")
    dh = U.conv_desc(N, H, W, 32, 2, 6, 1, y_dtype=L.F32, act=L.ACT_SIGMOID, impl=L.IMPL_TC)
    U.run_fprop(dh, xhg, w_ck, w_kc, bhg, out)
    ref = torch.sigmoid(_oracle_fprop(xh, wh, bh, 1))
    assert U.max_abs(out, ref) < 1e-5
    # ---- head wgrad: dy = bf16 gradient with a 16-byte pixel pitch (2 valid channels of 8)
    dz = U.bf16_round(torch.randn(N, H, W, 2, generator=g))
    dz8 = torch.zeros(N, H, W, 8, dtype=torch.bfloat16, device="cuda")
    dz8[..., :2] = dz.cuda().to(torch.bfloat16)
    xr, wr = xh.clone(), wh.clone().requires_grad_(True)
    gw, = torch.autograd.grad(_oracle_fprop(xr, wr, None, 1), [wr], dz)
    dw = torch.full((6, 6, 32, 2), 7.0, device="cuda")
    dwd = U.conv_desc(N, H, W, 32, 2, 6, 1, y_ld=8, impl=L.IMPL_TC)
    U.run_wgrad(dwd, xhg, dz8, dw)
    assert U.rel_l2(dw, gw) < F32_TOL
    # ---- vector projection: 1x1, C=16 -> K=64, accumulate into an existing tensor
    xp = U.bf16_round(torch.randn(N, 9, 10, 16, generator=g))
    wp = U.bf16_round(torch.randn(1, 1, 16, 64, generator=g) * 0.2)
    bp = torch.randn(64, generator=g) * 0.1
    base = U.bf16_round(torch.randn(N, 9, 10, 64, generator=g))
    w_ck, w_kc = U.prep_weights(wp.cuda())
    xpg, bpg = xp.cuda().to(torch.bfloat16), bp.cuda()
    y = base.cuda().to(torch.bfloat16)
    dp = U.conv_desc(N, 9, 10, 16, 64, 1, 1, impl=L.IMPL_TC, accumulate=1)
    U.run_fprop(dp, xpg, w_ck, w_kc, bpg, y)
    assert U.rel_l2(y.float(), _oracle_fprop(xp, wp, bp, 1) + base) < BF16_TOL
    dy = U.bf16_round(torch.randn(N, 9, 10, 64, generator=g))
    xr, wr = xp.clone().requires_grad_(True), wp.clone().requires_grad_(True)
    gx, gw = torch.autograd.grad(_oracle_fprop(xr, wr, None, 1), [xr, wr], dy)
    dyg = dy.cuda().to(torch.bfloat16)
    dx = torch.empty(N, 9, 10, 16, dtype=torch.bfloat16, device="cuda")
    dp.accumulate = 0
    U.run_dgrad(dp, dyg, w_ck, w_kc, None, dx)
    assert U.rel_l2(dx.float(), gx) < BF16_TOL
    dw = torch.empty(1, 1, 16, 64, device="cuda")
    U.run_wgrad(dp, xpg, dyg, dw)
    assert U.rel_l2(dw, gw) < F32_TOL


def test_head_dgrad_from_fp32_gradient():
    """Input-gradient of the K=2 head conv from the fp32 loss gradient (specialised CUDA-core kernel)."""
    g = torch.Generator().manual_seed(12)
    for (N, H, W, k) in ((2, 16, 32, 6), (1, 19, 45, 6), (2, 9, 10, 3)):
        x = U.bf16_round(torch.randn(N, H, W, 32, generator=g)).requires_grad_(True)
        w = U.bf16_round(torch.randn(k, k, 32, 2, generator=g) * 0.05)
        dz = torch.randn(N, H, W, 2, generator=g)
        gx, = torch.autograd.grad(_oracle_fprop(x, w, None, 1), [x], dz)
        w_ck, w_kc = U.prep_weights(w.cuda())
        dzg = dz.cuda()
        dx = torch.empty(N, H, W, 32, dtype=torch.bfloat16, device="cuda")
        d = U.conv_desc(N, H, W, 32, 2, k, 1, y_dtype=L.F32)
        U.run_dgrad(d, dzg, w_ck, w_kc, None, dx)
        assert U.rel_l2(dx.float(), gx) < BF16_TOL, (N, H, W, k)


def test_stem_wgrad_on_tensor_cores_and_batched_weight_prep():
    """enc1.down's kernel gradient from the 16-byte-pitch bf16 copy of the 2-channel input (TMA zero-fills
    channels 2..31 of the MMA tile), and the one-launch bf16 operand refresh."""
    g = torch.Generator().manual_seed(13)
    N, H, W = 2, 16, 32
    x = torch.rand(N, H, W, 2, generator=g)
    xq = U.bf16_round(x)
    w = U.bf16_round(torch.randn(3, 3, 2, 32, generator=g) * 0.2).requires_grad_(True)
    dy = U.bf16_round(torch.randn(N, H, W, 32, generator=g))
    gw, = torch.autograd.grad(_oracle_fprop(xq, w, None, 1), [w], dy)
    xg = x.cuda()
    x8 = torch.zeros(N, H, W, 8, dtype=torch.bfloat16, device="cuda")
    L.call("cast_pad_bf16", xg.data_ptr(), x8.data_ptr(), N * H * W, 2, 8)
    assert U.max_abs(x8[..., :2].float(), xq) == 0.0 and float(x8[..., 2:].float().abs().max()) == 0.0
    dyg = dy.cuda().to(torch.bfloat16)
    dw = torch.full((3, 3, 2, 32), 5.0, device="cuda")
    d = U.conv_desc(N, H, W, 2, 32, 3, 1, x_ld=8, impl=L.IMPL_TC)
    U.run_wgrad(d, x8, dyg, dw)
    assert U.rel_l2(dw, gw) < F32_TOL
    # batched weight prep == per-kernel weight prep
    # vector path (K % 4 == 0, even C) with full and ragged 64x64 tiles, and the scalar path (K = 2, K = 6, odd C)
    ws = [torch.randn(*s, generator=g).cuda() for s in
          [(3, 3, 32, 64), (6, 6, 32, 2), (3, 3, 2, 32), (1, 1, 96, 100), (2, 1, 70, 132), (1, 1, 33, 64), (1, 2, 64, 6)]]
    outs, rows = [], []
    for wt in ws:
        kh, kw, c, k = wt.shape
        ck = torch.zeros(kh * kw, c, k, dtype=torch.bfloat16, device="cuda"); kc = torch.zeros(kh * kw, k, c, dtype=torch.bfloat16, device="cuda")
        outs.append((ck, kc)); rows.append([wt.data_ptr(), ck.data_ptr(), kc.data_ptr(), kh * kw, c, k])
    table = torch.tensor(rows, dtype=torch.int64, device="cuda")
    L.call("weight_prep_batched", table.data_ptr(), len(rows))
    for wt, (ck, kc) in zip(ws, outs):
        r_ck, r_kc = U.prep_weights(wt)
        assert torch.equal(ck, r_ck) and torch.equal(kc, r_kc)


THIN_CASES = [(2, 16, 32, 3), (3, 24, 48, 6), (1, 8, 16, 6), (2, 144, 160, 6), (1, 40, 64, 5)]


@pytest.mark.parametrize("case", THIN_CASES, ids=[str(c) for c in THIN_CASES])
def test_thin_channel_tensor_core_kernels(case):
    """conv_thin.cu: the 2-channel stem (fprop, wgrad) and head (dgrad, wgrad). 3x3 and 6x6 kernels run on the warp-MMA
    kernels (A fragments straight from the bf16 halo patch / ldmatrix.trans of the wide tile), other tap counts (the 5x5
    case) on the tcgen05 kernels that build an im2col tile in shared memory. The thin operand is rounded to bf16 on chip,
    so the oracle gets the bf16-rounded tensor."""
    N, H, W, k = case
    g = torch.Generator().manual_seed(21 + k)
    # ---- stem: x fp32 [.,2] -> 32 channels
    x = torch.rand(N, H, W, 2, generator=g)
    xq = U.bf16_round(x)
    w = U.bf16_round(torch.randn(k, k, 2, 32, generator=g) * 0.2)
    b = torch.randn(32, generator=g) * 0.1
    dy = U.bf16_round(torch.randn(N, H, W, 32, generator=g))
    w_ck, w_kc = U.prep_weights(w.cuda())
    xg, bg, dyg = x.cuda(), b.cuda(), dy.cuda().to(torch.bfloat16)
    d = U.conv_desc(N, H, W, 2, 32, k, 1, x_dtype=L.F32)
    assert L.load().urir_conv_path(d, 0) == 1 and L.load().urir_conv_path(d, 2) == 1
    y = torch.empty(N, H, W, 32, dtype=torch.bfloat16, device="cuda")
    U.run_fprop(d, xg, w_ck, w_kc, bg, y)
    assert U.rel_l2(y.float(), _oracle_fprop(xq, w, b, 1)) < BF16_TOL
    # written inside a wider buffer (channel slice)
    wide = torch.zeros(N, H, W, 64, dtype=torch.bfloat16, device="cuda")
    d2 = U.conv_desc(N, H, W, 2, 32, k, 1, x_dtype=L.F32, y_ld=64, y_coff=32)
    U.run_fprop(d2, xg, w_ck, w_kc, bg, wide)
    assert torch.equal(wide[..., 32:], y) and float(wide[..., :32].float().abs().max()) == 0.0
    wr = w.clone().requires_grad_(True)
    gw, = torch.autograd.grad(_oracle_fprop(xq, wr, None, 1), [wr], dy)
    dw = torch.full((k, k, 2, 32), 5.0, device="cuda")
    U.run_wgrad(d, xg, dyg, dw)
    assert U.rel_l2(dw, gw) < F32_TOL
    # ---- head: 32 channels -> fp32 [.,2]
    xh = U.bf16_round(torch.randn(N, H, W, 32, generator=g))
    wh = U.bf16_round(torch.randn(k, k, 32, 2, generator=g) * 0.05)
    dz = torch.randn(N, H, W, 2, generator=g)
    dzq = U.bf16_round(dz)
    xr, whr = xh.clone().requires_grad_(True), wh.clone().requires_grad_(True)
    gx, gwh = torch.autograd.grad(_oracle_fprop(xr, whr, None, 1), [xr, whr], dzq)
    w_ck, w_kc = U.prep_weights(wh.cuda())
    dh = U.conv_desc(N, H, W, 32, 2, k, 1, y_dtype=L.F32)
    assert L.load().urir_conv_path(dh, 1) == 1 and L.load().urir_conv_path(dh, 2) == 1
    dx = torch.empty(N, H, W, 32, dtype=torch.bfloat16, device="cuda")
    U.run_dgrad(dh, dz.cuda(), w_ck, w_kc, None, dx)
    assert U.rel_l2(dx.float(), gx) < BF16_TOL
    dwh = torch.full((k, k, 32, 2), 7.0, device="cuda")
    U.run_wgrad(dh, xh.cuda().to(torch.bfloat16), dz.cuda(), dwh)
    assert U.rel_l2(dwh, gwh) < F32_TOL
    # head forward (conv_head.cu): horizontal taps in GEMM-N, shifted sum in the epilogue, fused sigmoid
    bh = torch.randn(2, generator=g) * 0.1
    out = torch.full((N, H, W, 2), -1.0, device="cuda")
    dhf = U.conv_desc(N, H, W, 32, 2, k, 1, y_dtype=L.F32, act=L.ACT_SIGMOID)
    assert L.load().urir_conv_path(dhf, 0) == 1
    U.run_fprop(dhf, xh.cuda().to(torch.bfloat16), w_ck, w_kc, bh.cuda(), out)
    assert U.max_abs(out, torch.sigmoid(_oracle_fprop(xh, wh, bh, 1))) < 1e-5


HALO_CASES = [
    (2, 16, 32, 32, 32, 3),      # E1b / D5b class: 64-byte rows (swizzle-64), SBO 640
    (2, 16, 32, 64, 32, 3),      # D5a class: 128-byte rows, SBO 1280
    (3, 24, 48, 64, 64, 3),      # E2b / D4b class
    (2, 24, 32, 128, 64, 3),     # D4a class: two channel chunks per tile
    (2, 20, 40, 32, 64, 3),      # ragged tiles (H % 8 != 0, W % 16 != 0)
    (2, 16, 16, 32, 32, 6),      # kernels=6 (pad 2,3): 13 x 21 halo box
    (5, 72, 80, 64, 64, 3),      # more tiles than SMs: several tiles per persistent CTA, both TMEM stages
]


@pytest.mark.parametrize("case", HALO_CASES, ids=[str(c) for c in HALO_CASES])
def test_halo_tile_fprop_dgrad(case):
    """conv_halo.cu: persistent CTAs, one TMA halo box per tile, taps as UMMA descriptor offsets, resident
    weights. Forced through URIR_IMPL_HALO at small shapes (AUTO only picks it for >= 4 tiles per SM)."""
    N, H, W, Cc, K, k = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, k, 1, seed=5)
    xg, dyg = x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16)
    w_ck, w_kc = U.prep_weights(w.cuda())
    d = U.conv_desc(N, H, W, Cc, K, k, 1, impl=L.IMPL_HALO)
    y = torch.empty(N, H, W, K, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * K, device="cuda")
    U.run_fprop(d, xg, w_ck, w_kc, bias.cuda(), y, stats)
    ref = _oracle_fprop(x, w, bias, 1)
    assert U.rel_l2(y.float(), ref) < BF16_TOL
    assert U.max_abs(stats[:K], ref.sum(dim=(0, 1, 2))) < 2e-3 * float(ref.abs().sum(dim=(0, 1, 2)).max())
    assert U.rel_l2(stats[K:], (ref ** 2).sum(dim=(0, 1, 2))) < F32_TOL * 10
    # output into the right half of a concat buffer
    cat = torch.zeros(N, H, W, 2 * K, dtype=torch.bfloat16, device="cuda")
    d2 = U.conv_desc(N, H, W, Cc, K, k, 1, y_ld=2 * K, y_coff=K, impl=L.IMPL_HALO)
    U.run_fprop(d2, xg, w_ck, w_kc, bias.cuda(), cat, None)
    assert torch.equal(cat[..., K:], y) and float(cat[..., :K].float().abs().max()) == 0.0
    # dgrad
    xr = x.clone().requires_grad_(True)
    gx, = torch.autograd.grad(_oracle_fprop(xr, w, None, 1), [xr], dy)
    dx = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    dstats = torch.zeros(2 * Cc, device="cuda")
    U.run_dgrad(d, dyg, w_ck, w_kc, None, dx, dstats)
    assert U.rel_l2(dx.float(), gx) < BF16_TOL
    assert U.max_abs(dstats[:Cc], gx.sum(dim=(0, 1, 2))) < 2e-3 * float(gx.abs().sum(dim=(0, 1, 2)).max())
    # the sums-only entry point (bias gradients): same dx, same channel sums, second half of the buffer untouched
    dx2 = torch.empty_like(dx)
    sums = torch.zeros(2 * Cc, device="cuda")
    U.run_dgrad_sums(d, dyg, w_ck, w_kc, None, dx2, sums)
    assert torch.equal(dx2, dx)
    assert U.max_abs(sums[:Cc], gx.sum(dim=(0, 1, 2))) < 2e-3 * float(gx.abs().sum(dim=(0, 1, 2)).max())
    assert float(sums[Cc:].abs().max()) == 0.0


WGRAD_HALO_CASES = [
    (2, 16, 32, 32, 32),      # E1b / D5b class: one MMA (M = 4 vertical slots x 32 ch, N = 3 x 32) per 16 pixels
    (2, 16, 32, 64, 32),      # D5a class: 128-byte x rows, two MMAs (vertical taps {0,1} and {2})
    (3, 24, 48, 64, 64),      # E2b / D4b class: N = 192
    (2, 24, 32, 128, 64),     # D4a class: two 64-channel x blocks (grid.y = 2)
    (5, 72, 80, 64, 64),      # many tiles per persistent CTA (ring wrap-around)
    (2, 20, 40, 32, 64),      # ragged tiles (H % 8 != 0, W % 16 != 0): TMA zero fill
    (2, 36, 40, 128, 128),    # E3b / D3b class: dy in two 64-channel blocks (grid.z), two x blocks, ragged rows
    (2, 36, 40, 256, 128),    # D3a class
]


@pytest.mark.parametrize("case", WGRAD_HALO_CASES, ids=[str(c) for c in WGRAD_HALO_CASES])
def test_halo_tile_wgrad(case):
    """conv_wgrad_halo.cu: all nine taps from one x halo box (vertical taps stacked along GEMM-M) and one dy halo
    box (horizontal taps stacked along GEMM-N). Also with x and dy living inside wider (concat) buffers."""
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 1, seed=11)
    xr, wr = x.clone(), w.clone().requires_grad_(True)
    gw, = torch.autograd.grad(_oracle_fprop(xr, wr, None, 1), [wr], dy)
    d = U.conv_desc(N, H, W, Cc, K, 3, 1, impl=L.IMPL_HALO)
    assert L.load().urir_conv_path(d, 2) == 1
    dw = torch.full((3, 3, Cc, K), 3.0, device="cuda")
    U.run_wgrad(d, x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16), dw)
    assert U.rel_l2(dw, gw) < F32_TOL
    # operands as channel slices of wider buffers (skip-concat layout)
    xw = torch.randn(N, H, W, Cc + 32).to(torch.bfloat16).cuda()
    xw[..., 32:] = x.cuda().to(torch.bfloat16)
    dyw = torch.randn(N, H, W, 2 * K).to(torch.bfloat16).cuda()
    dyw[..., :K] = dy.cuda().to(torch.bfloat16)
    d2 = U.conv_desc(N, H, W, Cc, K, 3, 1, x_ld=Cc + 32, x_coff=32, y_ld=2 * K, y_coff=0, impl=L.IMPL_HALO)
    dw2 = torch.zeros(3, 3, Cc, K, device="cuda")
    U.run_wgrad(d2, xw, dyw, dw2)
    assert U.rel_l2(dw2, gw) < F32_TOL


UP2_CASES = [
    (2, 16, 32, 32, 64),      # D5t / E2a-dgrad class: N = 4C = 128, K = 64
    (3, 24, 40, 64, 128),     # D4t / E3a-dgrad class: two GEMM-N tiles, two K chunks, ragged half-resolution tiles
    (5, 144, 160, 32, 64),    # many tiles per persistent CTA
]


@pytest.mark.parametrize("case", UP2_CASES, ids=[str(c) for c in UP2_CASES])
def test_stride2_dgrad_as_2x2_on_half_grid(case):
    """urir_conv2d_dgrad_up2: the input gradient of a 3x3 stride-2 SAME conv (= Conv2DTranspose forward) as one
    2x2 stride-1 problem with GEMM-N = (parity, channel); with bias, into a concat slice, and accumulating."""
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 2, seed=13)
    xr = x.clone().requires_grad_(True)
    gx, = torch.autograd.grad(_oracle_fprop(xr, w, None, 2), [xr], dy)
    dyg = dy.cuda().to(torch.bfloat16)
    w_up2 = torch.empty(4, 4 * Cc, K, dtype=torch.bfloat16, device="cuda")
    L.call("weight_prep_up2", w.cuda().contiguous().data_ptr(), w_up2.data_ptr(), Cc, K)
    d = U.conv_desc(N, H, W, Cc, K, 3, 2)
    assert L.load().urir_conv_path(d, 3) == 1
    cbias = torch.randn(Cc, generator=torch.Generator().manual_seed(7))
    dx = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    L.call("conv2d_dgrad_up2", C.byref(d), dyg.data_ptr(), w_up2.data_ptr(), cbias.cuda().data_ptr(), dx.data_ptr())
    assert U.rel_l2(dx.float(), gx + cbias) < BF16_TOL
    # right half of a concat buffer (Conv2DTranspose forward writes next to the skip connection)
    cat = torch.zeros(N, H, W, 2 * Cc, dtype=torch.bfloat16, device="cuda")
    d2 = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=Cc)
    L.call("conv2d_dgrad_up2", C.byref(d2), dyg.data_ptr(), w_up2.data_ptr(), cbias.cuda().data_ptr(), cat.data_ptr())
    assert torch.equal(cat[..., Cc:], dx) and float(cat[..., :Cc].float().abs().max()) == 0.0
    # accumulate into the left half (encoder dgrad adds to the skip gradient)
    base = U.bf16_round(torch.randn(N, H, W, 2 * Cc, generator=torch.Generator().manual_seed(9)))
    acc = base.cuda().to(torch.bfloat16)
    d3 = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=0, accumulate=1)
    L.call("conv2d_dgrad_up2", C.byref(d3), dyg.data_ptr(), w_up2.data_ptr(), None, acc.data_ptr())
    assert U.rel_l2(acc[..., :Cc].float(), gx + base[..., :Cc]) < BF16_TOL
    assert torch.equal(acc[..., Cc:].cpu().float(), base[..., Cc:])


S2_FPROP_CASES = [
    (2, 16, 32, 32, 64),      # E2a class: one channel chunk, N = 64
    (3, 24, 40, 64, 128),     # E3a class: two chunks, ragged half-resolution tiles
    (5, 144, 160, 32, 64),    # many tiles per persistent CTA
]


@pytest.mark.parametrize("case", S2_FPROP_CASES, ids=[str(c) for c in S2_FPROP_CASES])
def test_stride2_fprop_through_parity_planes(case):
    """conv_halo_s2_fprop: the four parity planes of x as halo boxes, taps as descriptor offsets (forced through
    URIR_IMPL_HALO); input read from a concat slice, with the fused channel statistics."""
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 2, seed=17)
    ref = _oracle_fprop(x, w, bias, 2)
    w_ck, w_kc = U.prep_weights(w.cuda())
    xw = torch.randn(N, H, W, 2 * Cc).to(torch.bfloat16).cuda()
    xw[..., Cc:] = x.cuda().to(torch.bfloat16)
    d = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=Cc, impl=L.IMPL_HALO)
    assert L.load().urir_conv_path(d, 0) == 1
    y = torch.empty(N, P, Q, K, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * K, device="cuda")
    U.run_fprop(d, xw, w_ck, w_kc, bias.cuda(), y, stats)
    assert U.rel_l2(y.float(), ref) < BF16_TOL
    assert U.max_abs(stats[:K], ref.sum(dim=(0, 1, 2))) < 2e-3 * float(ref.abs().sum(dim=(0, 1, 2)).max())
    assert U.rel_l2(stats[K:], (ref ** 2).sum(dim=(0, 1, 2))) < F32_TOL * 10


WGRAD_S2_CASES = [
    (2, 16, 32, 32, 64),      # E2a / D5t class: dy has one 64-channel half
    (3, 24, 40, 64, 128),     # E3a / D4t class: two x channel blocks, two dy halves, ragged tiles
    (5, 144, 160, 32, 64),    # many tiles per persistent CTA
]


@pytest.mark.parametrize("case", WGRAD_S2_CASES, ids=[str(c) for c in WGRAD_S2_CASES])
def test_stride2_wgrad_through_parity_planes(case):
    """conv_wgrad_halo_s2: nine taps of a stride-2 kernel gradient from the four parity planes of x (stacked along
    GEMM-M) and a dy box with one halo column (column shift stacked along GEMM-N); operands inside concat buffers."""
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 2, seed=19)
    wr = w.clone().requires_grad_(True)
    gw, = torch.autograd.grad(_oracle_fprop(x, wr, None, 2), [wr], dy)
    xw = torch.randn(N, H, W, 2 * Cc).to(torch.bfloat16).cuda()
    xw[..., :Cc] = x.cuda().to(torch.bfloat16)
    d = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=0, impl=L.IMPL_HALO)
    assert L.load().urir_conv_path(d, 2) == 1
    dw = torch.full((3, 3, Cc, K), 3.0, device="cuda")
    U.run_wgrad(d, xw, dy.cuda().to(torch.bfloat16), dw)
    assert U.rel_l2(dw, gw) < F32_TOL


# ------------------------------------------------------------------------------------------------
# round 2: the persistent padded-sequence kernel of the deep stride-1 layers (conv_deep.cu, URIR_IMPL_DEEP)
# ------------------------------------------------------------------------------------------------
DEEP_CASES = [
    # (N, H, W, C, K): every deep stride-1 3x3 layer of the canonical net (E3b..E5b, D2a/b, D3a/b) at reduced batch,
    # chosen so that CTAs get 0, 1 and several M tiles, tiles span images, and two rounds occur (36x40 with 12 images)
    (3, 9, 10, 512, 512),
    (64, 9, 10, 512, 512),          # the benchmark's bottleneck layer as is: 55 M tiles x 4 N tiles
    (5, 18, 20, 256, 256),
    (16, 18, 20, 512, 256),
    (12, 36, 40, 256, 128),         # 148 CTAs x 5-6 tiles: two rounds of three
    (2, 36, 40, 128, 128),
    (2, 36, 76, 128, 256),          # long-RIR width (config 5): Wp = 77
]


@pytest.mark.parametrize("case", DEEP_CASES, ids=[str(c) for c in DEEP_CASES])
def test_deep_kernel_fprop_dgrad(case):
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 1, seed=N + H)
    xg, dyg = x.cuda().to(torch.bfloat16), dy.cuda().to(torch.bfloat16)
    w_ck, w_kc = U.prep_weights(w.cuda())
    bg = bias.cuda()
    d = U.conv_desc(N, H, W, Cc, K, 3, 1, impl=L.IMPL_DEEP)
    fam0 = L.family_calls()["deep"]
    # AUTO picks this kernel for these shapes
    da = U.conv_desc(N, H, W, Cc, K, 3, 1, impl=L.IMPL_AUTO)
    assert L.load().urir_conv_path(C.byref(da), 0) == 1
    y = torch.full((N, H, W, K), 7.0, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * K, device="cuda")
    U.run_fprop(d, xg, w_ck, w_kc, bg, y, stats)
    ref = _oracle_fprop(x, w, bias, 1)
    assert U.rel_l2(y.float(), ref) < BF16_TOL, U.rel_l2(y.float(), ref)
    assert U.rel_l2(stats[:K], ref.sum(dim=(0, 1, 2))) < 5e-3 or U.max_abs(stats[:K], ref.sum(dim=(0, 1, 2))) < 1e-2 * (N ** 0.5)
    assert U.rel_l2(stats[K:], (ref ** 2).sum(dim=(0, 1, 2))) < F32_TOL * 10
    # bit-comparable with the one-tile-per-CTA tcgen05 kernel (same bf16 operands, fp32 accumulation, one rounding)
    y2 = torch.empty_like(y)
    U.run_fprop(U.conv_desc(N, H, W, Cc, K, 3, 1, impl=L.IMPL_TC), xg, w_ck, w_kc, bg, y2, None)
    assert U.rel_l2(y.float(), y2.float()) < 2e-3
    # dgrad, with the bias-gradient statistics
    xr = x.clone().requires_grad_(True)
    yr = _oracle_fprop(xr, w, None, 1)
    (gx,) = torch.autograd.grad(yr, [xr], dy)
    dx = torch.full((N, H, W, Cc), 7.0, dtype=torch.bfloat16, device="cuda")
    dstats = torch.zeros(2 * Cc, device="cuda")
    U.run_dgrad(d, dyg, w_ck, w_kc, None, dx, dstats)
    assert U.rel_l2(dx.float(), gx) < BF16_TOL, U.rel_l2(dx.float(), gx)
    assert U.max_abs(dstats[:Cc], gx.sum(dim=(0, 1, 2))) < 2e-3 * float(gx.abs().sum(dim=(0, 1, 2)).max())
    assert L.family_calls()["deep"] - fam0 == 2
    # sums-only statistics (bias gradients): same dx, same sums, no sums of squares
    dx2, sums = torch.empty_like(dx), torch.zeros(2 * Cc, device="cuda")
    U.run_dgrad_sums(d, dyg, w_ck, w_kc, None, dx2, sums)
    assert torch.equal(dx2, dx) and float(sums[Cc:].abs().max()) == 0.0
    assert U.max_abs(sums[:Cc], gx.sum(dim=(0, 1, 2))) < 2e-3 * float(gx.abs().sum(dim=(0, 1, 2)).max())


def test_deep_kernel_concat_slices_and_relu():
    """Reads a channel slice of a wider buffer and writes into one (skip-concat halves, u_net.py:308), ReLU epilogue of
    the BatchNorm-folded inference path, AUTO dispatch."""
    N, H, W, Cc, K = 4, 18, 20, 256, 256
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 1, seed=5)
    w_ck, w_kc = U.prep_weights(w.cuda())
    xbuf = torch.zeros(N, H, W, 2 * Cc, dtype=torch.bfloat16, device="cuda")
    xbuf[..., Cc:] = x.cuda().to(torch.bfloat16)
    xbuf[..., :Cc] = 9.0                                         # must not be read
    ybuf = torch.full((N, H, W, 2 * K), 3.0, dtype=torch.bfloat16, device="cuda")
    d = U.conv_desc(N, H, W, Cc, K, 3, 1, x_ld=2 * Cc, x_coff=Cc, y_ld=2 * K, y_coff=0, impl=L.IMPL_AUTO, act=L.ACT_RELU)
    fam0 = L.family_calls()["deep"]
    U.run_fprop(d, xbuf, w_ck, w_kc, bias.cuda(), ybuf, None)
    assert L.family_calls()["deep"] == fam0 + 1
    ref = torch.relu(_oracle_fprop(x, w, bias, 1))
    assert U.rel_l2(ybuf[..., :K].float(), ref) < BF16_TOL
    assert float((ybuf[..., K:].float() - 3.0).abs().max()) == 0.0


DEEP_S2_CASES = [
    # (N, H, W, C, K): the deep strided layers of the canonical net (E4a 36x40 128 -> 256, E5a 18x20 256 -> 512; the
    # decoder's Conv2DTranspose layers are the same shapes run as dgrad), reduced batch and the benchmark's batch
    (3, 18, 20, 256, 512),
    (64, 18, 20, 256, 512),
    (5, 36, 40, 128, 256),
    (64, 36, 40, 128, 256),
    (2, 36, 76, 128, 128),          # long-RIR width (config 5), one channel tile
]


@pytest.mark.parametrize("case", DEEP_S2_CASES, ids=[str(c) for c in DEEP_S2_CASES])
def test_deep_kernel_stride2(case):
    """conv_deep.cu on the half-resolution grid: stride-2 fprop from the four parity planes of x (strided tensor maps,
    4 / 2 / 2 / 1 taps each) and stride-2 dgrad as four output-parity classes over the same dy tiles -- with the
    Conv2DTranspose bias, into a concat half, and accumulating into an existing gradient (encoder skip). Forced through
    URIR_IMPL_DEEP: AUTO keeps conv_igemm for these layers (measured not slower, see deep_supported)."""
    N, H, W, Cc, K = case
    x, w, bias, dy, P, Q = _inputs(N, H, W, Cc, K, 3, 2, seed=N + W)
    w_ck, w_kc = U.prep_weights(w.cuda())
    fam0 = L.family_calls()["deep"]
    d = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=Cc, impl=L.IMPL_DEEP)
    xw = torch.randn(N, H, W, 2 * Cc).to(torch.bfloat16).cuda()
    xw[..., Cc:] = x.cuda().to(torch.bfloat16)
    y = torch.full((N, P, Q, K), 7.0, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2 * K, device="cuda")
    U.run_fprop(d, xw, w_ck, w_kc, bias.cuda(), y, stats)
    ref = _oracle_fprop(x, w, bias, 2)
    assert U.rel_l2(y.float(), ref) < BF16_TOL, U.rel_l2(y.float(), ref)
    assert U.max_abs(stats[:K], ref.sum(dim=(0, 1, 2))) < 2e-3 * float(ref.abs().sum(dim=(0, 1, 2)).max())
    assert U.rel_l2(stats[K:], (ref ** 2).sum(dim=(0, 1, 2))) < F32_TOL * 10
    assert L.family_calls()["deep"] - fam0 == 1
    # dgrad = Conv2DTranspose forward: bias added, written into the right half of a concat buffer
    xr = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad(_oracle_fprop(xr, w, None, 2), [xr], dy)
    cb = torch.randn(Cc)
    dyg = dy.cuda().to(torch.bfloat16)
    cat = torch.full((N, H, W, 2 * Cc), 3.0, dtype=torch.bfloat16, device="cuda")
    d = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=Cc, impl=L.IMPL_DEEP)
    U.run_dgrad(d, dyg, w_ck, w_kc, cb.cuda(), cat)
    assert L.family_calls()["deep"] - fam0 == 2
    assert U.rel_l2(cat[..., Cc:].float(), gx + cb) < BF16_TOL, U.rel_l2(cat[..., Cc:].float(), gx + cb)
    assert float((cat[..., :Cc].float() - 3.0).abs().max()) == 0.0
    # dgrad accumulating into the skip gradient (left half)
    g = torch.Generator().manual_seed(3)
    old = U.bf16_round(torch.randn(N, H, W, Cc, generator=g))
    cat[..., :Cc] = old.cuda().to(torch.bfloat16)
    d = U.conv_desc(N, H, W, Cc, K, 3, 2, x_ld=2 * Cc, x_coff=0, impl=L.IMPL_DEEP, accumulate=1)
    U.run_dgrad(d, dyg, w_ck, w_kc, None, cat)
    assert L.family_calls()["deep"] - fam0 == 3
    assert U.rel_l2(cat[..., :Cc].float(), gx + old) < BF16_TOL, U.rel_l2(cat[..., :Cc].float(), gx + old)
    # same bf16 operands through the one-tile-per-CTA kernel
    dx2 = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    U.run_dgrad(U.conv_desc(N, H, W, Cc, K, 3, 2, impl=L.IMPL_TC), dyg, w_ck, w_kc, cb.cuda(), dx2)
    assert U.rel_l2(cat[..., Cc:].float(), dx2.float()) < 2e-3
