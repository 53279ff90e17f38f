"""GPU parity of the bandwidth-bound kernels through the C-ABI against CPU restatements:
BatchNorm+ReLU fwd/bwd, amp/phase loss, Adam, the vector block, STFT / iSTFT."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import signal_oracle as SO
from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L

pytestmark = pytest.mark.gpu


def test_bn_relu_forward_backward():
    g = torch.Generator().manual_seed(0)
    N, H, W, Cc = 3, 10, 12, 64
    ld, coff = 128, 64                       # live inside a concat buffer
    x = U.bf16_round(torch.randn(N, H, W, Cc, generator=g) * 2 + 0.5)
    gamma = torch.rand(Cc, generator=g) + 0.5
    beta = torch.randn(Cc, generator=g) * 0.3
    dy = U.bf16_round(torch.randn(N, H, W, Cc, generator=g))
    npix = N * H * W
    # oracle (training-mode BN with biased variance, eps 1e-3) + autograd
    xr = x.clone().requires_grad_(True); gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    mean = xr.mean(dim=(0, 1, 2)); var = xr.var(dim=(0, 1, 2), unbiased=False)
    yr = torch.relu((xr - mean) * torch.rsqrt(var + O.BN_EPS) * gr + br)
    gx, gg, gb = torch.autograd.grad(yr, [xr, gr, br], dy)

    gam_g, bet_g = gamma.cuda(), beta.cuda()        # keep device copies alive across the async calls
    xg = x.cuda().to(torch.bfloat16)
    stats = torch.stack([x.sum(dim=(0, 1, 2)), (x * x).sum(dim=(0, 1, 2))]).flatten().cuda()
    mm, mv = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    ss, mr = torch.empty(2 * Cc, device="cuda"), torch.empty(2 * Cc, device="cuda")
    L.call("bn_finalize", stats.data_ptr(), float(npix), gam_g.data_ptr(), bet_g.data_ptr(),
           mm.data_ptr(), mv.data_ptr(), 0.99, 1e-3, 0, ss.data_ptr(), mr.data_ptr(), Cc)
    assert U.rel_l2(mm, 0.01 * mean.detach()) < 1e-4
    assert U.rel_l2(mv, 0.99 + 0.01 * var.detach()) < 1e-5
    ybuf = torch.zeros(N, H, W, ld, dtype=torch.bfloat16, device="cuda")
    L.call("bn_relu_fwd", xg.data_ptr(), Cc, 0, ss.data_ptr(), ybuf.data_ptr(), ld, coff, npix, Cc, 1)
    assert U.rel_l2(ybuf[..., coff:].float(), yr.detach()) < 4e-3
    assert float(ybuf[..., :coff].float().abs().max()) == 0.0
    # the fused training-mode launch (finalize + apply) gives the same tensors and moving statistics
    mm2, mv2 = torch.zeros(Cc, device="cuda"), torch.ones(Cc, device="cuda")
    ss2, mr2 = torch.empty(2 * Cc, device="cuda"), torch.empty(2 * Cc, device="cuda")
    ybuf2 = torch.zeros(N, H, W, ld, dtype=torch.bfloat16, device="cuda")
    L.call("bn_relu_fwd_train", xg.data_ptr(), Cc, 0, stats.data_ptr(), float(npix), gam_g.data_ptr(), bet_g.data_ptr(),
           mm2.data_ptr(), mv2.data_ptr(), 0.99, 1e-3, 0, ss2.data_ptr(), mr2.data_ptr(), ybuf2.data_ptr(), ld, coff,
           npix, Cc)
    assert torch.equal(ybuf2, ybuf) and torch.equal(ss2, ss) and torch.equal(mr2, mr)
    assert torch.equal(mm2, mm) and torch.equal(mv2, mv)

    dybuf = torch.zeros(N, H, W, ld, dtype=torch.bfloat16, device="cuda")
    dybuf[..., coff:] = dy.cuda().to(torch.bfloat16)
    sums = torch.empty(2 * Cc, device="cuda")
    L.call("bn_relu_bwd_reduce", dybuf.data_ptr(), ld, coff, xg.data_ptr(), Cc, 0, ss.data_ptr(), mr.data_ptr(),
           sums.data_ptr(), npix, Cc, 0)
    dx = torch.empty(N, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    dgamma, dbeta, dbias = (torch.empty(Cc, device="cuda") for _ in range(3))
    L.call("bn_relu_bwd_apply", dybuf.data_ptr(), ld, coff, xg.data_ptr(), Cc, 0, ss.data_ptr(), mr.data_ptr(),
           gam_g.data_ptr(), sums.data_ptr(), dx.data_ptr(), Cc, 0, dgamma.data_ptr(), dbeta.data_ptr(),
           dbias.data_ptr(), npix, Cc, 0)
    assert U.rel_l2(dx.float(), gx) < 6e-3
    assert U.rel_l2(dgamma, gg) < 1e-4
    assert U.rel_l2(dbeta, gb) < 1e-4
    assert U.max_abs(dbias, gx.sum(dim=(0, 1, 2))) < 0.05       # analytically zero; bf16 rounding noise

    # inference mode: scale/shift from the moving statistics
    L.call("bn_finalize", None, 0.0, gam_g.data_ptr(), bet_g.data_ptr(), mm.data_ptr(), mv.data_ptr(),
           0.99, 1e-3, 0, ss.data_ptr(), mr.data_ptr(), Cc)
    rs = torch.rsqrt(mv.cpu() + 1e-3) * gamma
    assert U.rel_l2(ss[:Cc], rs) < 1e-5
    assert U.max_abs(ss[Cc:], beta - mm.cpu() * rs) < 1e-5


@pytest.mark.parametrize("kind", ["amp_phase", "dp"])
def test_ampphase_loss_and_grad(kind):
    g = torch.Generator().manual_seed(1)
    B, H, W = 3, 16, 20
    yt = torch.rand(B, H, W, 2, generator=g)
    z = torch.randn(B, H, W, 2, generator=g, requires_grad=True)
    yp = torch.sigmoid(z)
    if kind == "amp_phase":           # amp_phase_trainer.py:143-168
        loss, lp, ls = O.amp_phase_loss(yt, yp)
        w_amp = w_ph = 1.0 / (B * H * W)
    else:                             # main_training.py:203-235, alpha .9, global batch 2B
        alpha, gb = 0.9, 2 * B
        loss = O.dp_loss(yt, yp, alpha, gb)
        _, lp, ls = O.amp_phase_loss(yt, yp)
        w_amp, w_ph = alpha / (H * W * 2 * gb), (1 - alpha) / (H * W * 2 * gb)
    gz, = torch.autograd.grad(loss, z)
    losses = torch.empty(4, device="cuda"); grad = torch.empty(B, H, W, 2, device="cuda")
    yt_g, yp_g = yt.cuda(), yp.detach().cuda()
    g16 = torch.zeros(B, H, W, 8, dtype=torch.bfloat16, device="cuda")
    L.call("ampphase_loss", yt_g.data_ptr(), yp_g.data_ptr(), B * H * W, w_amp, w_ph, 1,
           losses.data_ptr(), grad.data_ptr(), g16.data_ptr(), 8)
    assert U.rel_l2(g16[..., :2].float(), gz) < 4e-3 and float(g16[..., 2:].float().abs().max()) == 0.0
    assert abs(float(losses[0]) - float(loss)) < 1e-5 * max(1.0, abs(float(loss)))
    assert abs(float(losses[1]) - float(lp)) < 1e-5 and abs(float(losses[2]) - float(ls)) < 1e-6
    assert U.rel_l2(grad, gz) < 1e-5


def test_adam_matches_keras_convention():
    g = torch.Generator().manual_seed(2)
    n = 10007
    p = torch.randn(n, generator=g); p0 = p.clone()
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.cuda(), m.cuda(), v.cuda()
    lr = torch.tensor([1e-3], device="cuda"); step = torch.zeros(1, dtype=torch.int32, device="cuda")
    params, ms, vs = {"w": p}, {"w": m}, {"w": v}
    for t in range(1, 4):
        gr = torch.randn(n, generator=g) * (10.0 ** (t - 2))
        O.keras_adam_step(params, {"w": gr}, ms, vs, t, 1e-3)
        gr_g = gr.cuda()
        L.call("adam", pc.data_ptr(), gr_g.data_ptr(), mc.data_ptr(), vc.data_ptr(), n, lr.data_ptr(),
               step.data_ptr(), 0.9, 0.999, 1e-7)
        L.call("step_increment", step.data_ptr())
    assert int(step) == 3
    assert U.rel_l2(pc - p0.cuda(), p - p0) < 1e-4      # powf/sqrtf vs the CPU's double-precision scalars
    assert U.rel_l2(vc, v) < 1e-4


def test_vector_block_dense_embedding():
    g = torch.Generator().manual_seed(4)
    B, T, D, Nn = 5, 32, 256, 1440
    idx = torch.randint(0, 2000, (B, T), generator=g, dtype=torch.int32)
    table = (torch.rand(2000, D, generator=g) - 0.5) * 0.1
    w = U.bf16_round((torch.rand(T * D, Nn, generator=g) - 0.5) * 0.05)
    bias = torch.randn(Nn, generator=g) * 0.1
    mask = (torch.rand(B, Nn, generator=g) > 0.3).float() / 0.7
    x = U.bf16_round(table[idx.long()].reshape(B, -1))
    ref = (x @ w + bias) * mask
    xg = torch.empty(B, T * D, dtype=torch.bfloat16, device="cuda")
    idx_g, table_g, bias_g, mask_g = idx.cuda(), table.cuda(), bias.cuda(), mask.cuda()
    L.call("embedding_fwd", idx_g.data_ptr(), table_g.data_ptr(), xg.data_ptr(), B, T, D, 2000)
    assert U.max_abs(xg.float(), x) == 0.0
    wg = w.cuda().to(torch.bfloat16)
    wg_kn = torch.empty(1, T * D, Nn, dtype=torch.bfloat16, device="cuda")
    wg_nk = torch.empty(1, Nn, T * D, dtype=torch.bfloat16, device="cuda")
    L.call("weight_prep", w.cuda().contiguous().data_ptr(), wg_kn.data_ptr(), wg_nk.data_ptr(), 1, T * D, Nn)
    assert torch.equal(wg_kn[0], wg) and torch.equal(wg_nk[0], wg.t())
    out = torch.empty(B, Nn, dtype=torch.bfloat16, device="cuda")
    L.call("dense_fwd", xg.data_ptr(), wg_kn.data_ptr(), wg_nk.data_ptr(), bias_g.data_ptr(), mask_g.data_ptr(),
           out.data_ptr(), B, T * D, Nn)
    assert U.rel_l2(out.float(), ref) < 6e-3          # two bf16 roundings: Dense output, then the dropout scale
    dy = U.bf16_round(torch.randn(B, Nn, generator=g))
    ge = U.bf16_round(dy * mask)                       # the masked gradient is a bf16 tensor-core operand
    dw = torch.full((T * D, Nn), 7.0, device="cuda"); db = torch.empty(Nn, device="cuda")
    dx = torch.empty(B, T * D, dtype=torch.bfloat16, device="cuda")
    dy_g = dy.cuda().to(torch.bfloat16)
    dy_eff = torch.empty(B, Nn, dtype=torch.bfloat16, device="cuda")
    L.call("dense_bwd", xg.data_ptr(), wg_kn.data_ptr(), wg_nk.data_ptr(), dy_g.data_ptr(), mask_g.data_ptr(),
           dy_eff.data_ptr(), dw.data_ptr(), db.data_ptr(), dx.data_ptr(), B, T * D, Nn)
    assert U.max_abs(dy_eff.float(), ge) == 0.0
    assert U.rel_l2(dw, x.t() @ ge) < 1e-4
    assert U.rel_l2(db, ge.sum(0)) < 1e-4
    assert U.rel_l2(dx.float(), ge @ w.t()) < 4e-3      # bf16 output
    dtab = torch.empty(2000, D, device="cuda")
    L.call("embedding_bwd", idx_g.data_ptr(), dx.data_ptr(), L.BF16, dtab.data_ptr(), B, T, D, 2000)
    ref_t = torch.zeros(2000, D).index_add_(0, idx.long().flatten(), dx.float().cpu().reshape(B * T, D))
    assert U.rel_l2(dtab, ref_t) < 1e-5
    dx32 = dx.float()
    L.call("embedding_bwd", idx_g.data_ptr(), dx32.data_ptr(), L.F32, dtab.data_ptr(), B, T, D, 2000)
    assert U.rel_l2(dtab, ref_t) < 1e-5
    # dropout mask: right keep-rate, right scale, new mask per step
    m1 = torch.empty(B, Nn, device="cuda"); m2 = torch.empty(B, Nn, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.call("dropout_mask", m1.data_ptr(), B * Nn, 0.3, 500, step.data_ptr())
    step += 1
    L.call("dropout_mask", m2.data_ptr(), B * Nn, 0.3, 500, step.data_ptr())
    assert {round(float(v), 4) for v in np.unique(m1.cpu().numpy())} <= {0.0, round(1 / 0.7, 4)}
    assert abs(float((m1 > 0).float().mean()) - 0.7) < 0.03 and not torch.equal(m1, m2)


def _stft_desc(pad_mode=0, normalized=1, remove_mean=1):
    return L.StftDesc(256, 128, 64, 9600, 129, 151, 144, 160, pad_mode, remove_mean, normalized)


@pytest.mark.parametrize("pad_mode", [0, 1], ids=["constant", "reflect"])
def test_stft_ampphase_and_inverse(pad_mode):
    rng = np.random.default_rng(0)
    B = 5
    wav = SO.synthetic_rir(B, rng) + 0.01            # non-zero mean, removed on device
    d = _stft_desc(pad_mode)
    spec = torch.empty(B, 144, 160, 2, device="cuda")
    wav_g = torch.from_numpy(wav).cuda()
    L.call("stft_ampphase", wav_g.data_ptr(), B, C.byref(d), spec.data_ptr())
    spec = spec.cpu().numpy()
    mode = "constant" if pad_mode == 0 else "reflect"
    for i in range(B):
        ref = SO.preprocess(wav[i], pad_mode=mode)
        assert np.abs(spec[i, :, :, 0] - ref[:, :, 0]).max() < 2e-4          # normalised log-amplitude
        # phase is ill-conditioned where |S| ~ 0 and wraps at +-pi: compare as unit phasors where loud
        loud = ref[:129, :151, 0] > 0.35
        dphi = 2 * math.pi * (spec[i, :129, :151, 1] - ref[:129, :151, 1])
        assert np.abs(np.sin(dphi / 2))[loud].max() < 2e-3
        assert spec[i, 129:, :, :].max() == 0.0 and spec[i, :, 151:, :].max() == 0.0   # TensorPadder zeros
    # inverse on the oracle's spectrogram: PostProcess.post_process
    feat = np.stack([SO.preprocess(w, pad_mode="constant") for w in wav])
    out = torch.empty(B, 9600, device="cuda")
    feat_g = torch.from_numpy(feat).cuda()
    L.call("istft_from_ampphase", feat_g.data_ptr(), B, C.byref(_stft_desc()), out.data_ptr())
    out = out.cpu().numpy()
    for i in range(B):
        ref = SO.post_process(feat[i])
        missa = 20 * math.log10(np.linalg.norm(out[i] - ref) / np.linalg.norm(ref))
        assert missa < -60.0, missa                                          # SURVEY 8c tolerance
    # round trip wav -> GPU stft -> GPU istft (preprocess.py:201-205 prints this misalignment)
    spec_g = torch.empty(B, 144, 160, 2, device="cuda")
    L.call("stft_ampphase", wav_g.data_ptr(), B, C.byref(_stft_desc()), spec_g.data_ptr())
    back = torch.empty(B, 9600, device="cuda")
    L.call("istft_from_ampphase", spec_g.data_ptr(), B, C.byref(_stft_desc()), back.data_ptr())
    back = back.cpu().numpy()
    for i in range(B):
        x0 = SO.remove_mean(wav[i])
        inner = slice(256, 9600 - 256)
        assert 20 * math.log10(np.linalg.norm(back[i][inner] - x0[inner]) / np.linalg.norm(x0[inner])) < -60.0


def test_unsupported_shapes_fail_loudly():
    d = L.StftDesc(512, 128, 64, 9600, 257, 151, 272, 160, 0, 1, 1)
    spec = torch.empty(1, 272, 160, 2, device="cuda"); wav = torch.zeros(1, 9600, device="cuda")
    with pytest.raises(L.UrirError):
        L.call("stft_ampphase", wav.data_ptr(), 1, C.byref(d), spec.data_ptr())
    with pytest.raises(L.UrirError):
        L.call("ampphase_loss", wav.data_ptr(), wav.data_ptr(), 3, 1.0, 1.0, 0, wav.data_ptr(), None, None, 0)


def test_stft_long_rir():
    """BASELINE config 5: 0.4 s at 48 kHz = 19200 samples -> (129, 301) -> padded (144, 304); STFT features and the
    inverse against the numpy oracle."""
    from unet_rir_b200.postprocess import post_process_batch
    from unet_rir_b200.preprocess import preprocess_batch
    rng = np.random.default_rng(2)
    wav = SO.synthetic_rir(2, rng, length=19200)
    spec = preprocess_batch(wav, padded=(144, 304)).cpu().numpy()
    assert spec.shape == (2, 144, 304, 2)
    for i in range(2):
        w = wav[i] - wav[i].mean()
        a, p = SO.normalize(*SO.extract(w))
        assert a.shape == (129, 301)
        assert np.abs(spec[i, :129, :301, 0] - a).max() < 2e-4
        assert float(np.abs(spec[i, 129:]).max()) == 0.0 and float(np.abs(spec[i, :, 301:]).max()) == 0.0
    back = post_process_batch(spec, des_shape=(129, 301)).cpu().numpy()
    assert back.shape == (2, 19200)
    for i in range(2):
        w = wav[i] - wav[i].mean()
        assert 20 * np.log10(np.linalg.norm(back[i][128:-128] - w[128:-128]) / np.linalg.norm(w[128:-128])) < -60


# ------------------------------------------------------------------------------------------------
# round 2: the other optimisers of the reference (amp_phase_trainer.py:30-35, trainer.py:37-38) on the device
# ------------------------------------------------------------------------------------------------
def test_sgd_nadam_lamb_match_keras_conventions():
    g = torch.Generator().manual_seed(12)
    sizes = [1000, 37, 4096, 8]                       # four "variables" laid out back to back (offsets 4-aligned)
    offs, o = [], 0
    for s in sizes:
        offs.append(o); o += (s + 3) // 4 * 4
    n = o
    p0 = torch.randn(n, generator=g)
    for o_, s in zip(offs, sizes):
        p0[o_ + s:o_ + (s + 3) // 4 * 4] = 0
    grads = [torch.randn(n, generator=g) * (10.0 ** (t - 2)) for t in range(4)]
    lr = 2e-3
    lr_dev = torch.tensor([lr], device="cuda")

    # ---- SGD (tf.keras.optimizers.SGD defaults: no momentum)
    pc = p0.clone().cuda()
    ref = {"w": p0.clone()}
    for gr in grads:
        O.keras_sgd_step(ref, {"w": gr}, lr)
        L.call("sgd", pc.data_ptr(), gr.cuda().data_ptr(), n, lr_dev.data_ptr())
    assert U.rel_l2(pc, ref["w"]) < 1e-6

    # ---- Nadam (momentum-schedule product carried in device memory across steps)
    pc, mc, vc = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    coef = torch.ones(4, device="cuda")
    ref, m, v, state = {"w": p0.clone().double()}, {"w": torch.zeros(n).double()}, {"w": torch.zeros(n).double()}, {"step": 0, "m_schedule": 1.0}
    for gr in grads:
        O.keras_nadam_step(ref, {"w": gr.double()}, m, v, state, lr)
        L.call("nadam", pc.data_ptr(), gr.cuda().data_ptr(), mc.data_ptr(), vc.data_ptr(), n, lr_dev.data_ptr(),
               step.data_ptr(), coef.data_ptr(), 0.9, 0.999, 1e-7)
        L.call("step_increment", step.data_ptr())
    assert abs(float(coef[0]) - state["m_schedule"]) < 1e-6 * state["m_schedule"]
    assert U.rel_l2(pc - p0.cuda(), (ref["w"] - p0.double()).float()) < 2e-5
    assert U.rel_l2(mc, m["w"].float()) < 1e-6 and U.rel_l2(vc, v["w"].float()) < 5e-5      # fp32 v over gradients spanning 1e-2 .. 1e1

    # ---- LAMB: per-variable trust ratio
    pc, mc, vc = p0.clone().cuda(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    upd, norms = torch.empty(n, device="cuda"), torch.zeros(2 * len(sizes), device="cuda")
    table = torch.tensor([[o_, s] for o_, s in zip(offs, sizes)], dtype=torch.int64, device="cuda")
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    names = [f"v{i}" for i in range(len(sizes))]
    ref = {k: p0[o_:o_ + s].clone().double() for k, o_, s in zip(names, offs, sizes)}
    m = {k: torch.zeros_like(t) for k, t in ref.items()}
    v = {k: torch.zeros_like(t) for k, t in ref.items()}
    for t, gr in enumerate(grads, 1):
        O.tfa_lamb_step(ref, {k: gr[o_:o_ + s].double() for k, o_, s in zip(names, offs, sizes)}, m, v, t, lr)
        L.call("lamb", pc.data_ptr(), gr.cuda().data_ptr(), mc.data_ptr(), vc.data_ptr(), upd.data_ptr(), table.data_ptr(),
               len(sizes), norms.data_ptr(), lr_dev.data_ptr(), step.data_ptr(), 0.9, 0.999, 1e-6, 0.0)
        L.call("step_increment", step.data_ptr())
    for k, o_, s in zip(names, offs, sizes):
        got, want = pc[o_:o_ + s].cpu() - p0[o_:o_ + s], (ref[k] - p0[o_:o_ + s].double()).float()
        assert U.rel_l2(got, want) < 5e-5, (k, U.rel_l2(got, want))


def test_trainer_optimizer_selection_runs_on_the_device():
    """'nadam' / 'sgd' / 'adam' substring selection (amp_phase_trainer.py:30-35): each drives its own device kernel inside
    the captured step; one Nadam step equals the oracle's Keras Nadam applied to the device's gradient."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(13)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    for name, expect in (("Nadam", "nadam"), ("my_sgd", "sgd"), ("adam", "adam")):
        unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
        eng = unet.model.engine
        tr = Trainer(0.9, 1, name.lower(), [ModelCheckpoint("/tmp/urir_opt", False, 0), EarlyStopping(5)], [False, 0], 1e-3, "opt")
        assert tr.optimizer == expect
        tr.dropout = False
        p0 = eng.P.clone()
        tr.step(x, y, emb, unet)
        gdev = eng.G.clone()
        step = eng.P - p0
        if expect == "sgd":      # p - lr * g in fp32: exact up to the rounding of p itself (|p| ~ 5e-2, ulp ~ 4e-9)
            assert float((eng.P - (p0 - 1e-3 * gdev)).abs().max()) < 1e-8
        elif expect == "nadam":
            ref, m, v = {"w": p0.double().cpu()}, {"w": torch.zeros_like(p0).double().cpu()}, {"w": torch.zeros_like(p0).double().cpu()}
            O.keras_nadam_step(ref, {"w": gdev.double().cpu()}, m, v, {"step": 0, "m_schedule": 1.0}, 1e-3)
            assert U.rel_l2(step.cpu(), (ref["w"] - p0.double().cpu()).float()) < 1e-4
        else:
            ref, m, v = {"w": p0.double().cpu()}, {"w": torch.zeros_like(p0).double().cpu()}, {"w": torch.zeros_like(p0).double().cpu()}
            O.keras_adam_step(ref, {"w": gdev.double().cpu()}, m, v, 1, 1e-3)
            assert U.rel_l2(step.cpu(), (ref["w"] - p0.double().cpu()).float()) < 1e-4
        for _ in range(2):                      # capture + replay stay finite
            l = tr.step(x, y, emb, unet)[0]
        assert bool(torch.isfinite(l))


def test_generic_trainer_mse_loss_and_lamb_step():
    """trainer.py: loss = squared error over both channels, differentiated as the SUM of the per-pixel channel means
    (tape.gradient of a non-scalar), optimiser 'lamb' -> tensorflow_addons LAMB. One step against torch autograd of the same
    scalar through the oracle graph (on the device's forward state) and the oracle's LAMB."""
    from unet_rir_b200.dl_models.u_net import UNet
    from unet_rir_b200.trainer import EarlyStopping, ModelCheckpoint, Trainer
    g = torch.Generator().manual_seed(17)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    # the loss kernel alone
    yp = torch.rand(2, 144, 160, 2, generator=g)
    out = torch.empty(4, device="cuda"); grad = torch.empty(2, 144, 160, 2, device="cuda")
    yd, ypd = y.cuda(), yp.cuda()           # named: a temporary's storage would be recycled for the second .cuda()
    L.call("mse2_loss", yd.data_ptr(), ypd.data_ptr(), 2 * 144 * 160, 0.5, 1, out.data_ptr(), grad.data_ptr())
    d = yp - y
    assert abs(float(out[0]) - 0.5 * float((d ** 2).sum())) < 1e-4 * float((d ** 2).sum())
    assert abs(float(out[3]) - float((d ** 2).mean())) < 1e-6 and abs(float(out[2]) - float((d[..., 0] ** 2).mean())) < 1e-6
    assert abs(float(out[1]) - float((1 - torch.cos(2 * math.pi * d[..., 1])).mean())) < 1e-5
    assert U.rel_l2(grad.cpu(), d * yp * (1 - yp)) < 1e-5
    # one LAMB step through the trainer
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    eng = unet.model.engine
    tr = Trainer(0.9, 1, "lamb", [ModelCheckpoint("/tmp/urir_gt", False, 0), EarlyStopping(5)], [False, 0], 1e-3, "gt")
    assert tr.optimizer == "lamb"
    tr.dropout = False
    p0 = {n: eng.param[n].detach().cpu().clone() for n in eng.trainable_names()}
    loss, lp, ls = tr.step(x, y, emb, unet)
    outp = eng._buffers(2)["out"].cpu()
    assert abs(float(loss) - float(((outp - y) ** 2).mean())) < 1e-5
    ref = {n: p0[n].double() for n in p0}
    m = {n: torch.zeros_like(t) for n, t in ref.items()}
    v = {n: torch.zeros_like(t) for n, t in ref.items()}
    O.tfa_lamb_step(ref, {n: eng.grad[n].double().cpu() for n in ref}, m, v, 1, 1e-3)
    for n in ("enc3.blk.c1.w", "dec2.fuse.w", "vec.dense.w", "head.w", "enc1.blk.bn1.gamma"):
        got, want = eng.param[n].detach().cpu() - p0[n], (ref[n] - p0[n].double()).float()
        assert U.rel_l2(got, want) < 2e-3, (n, U.rel_l2(got, want))
    for _ in range(2):
        assert bool(torch.isfinite(tr.step(x, y, emb, unet)[0]))
