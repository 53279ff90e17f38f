"""GPU parity of the whole U-Net path: engine forward / backward / train step (bf16 tcgen05 path and
the CUDA-core cross-check) against the fp32 CPU oracle on identical weights, inputs and dropout mask.

Tolerances (bf16 activations and GEMM operands vs the oracle's fp32; TF itself would run these convs
in TF32): per-layer activations rel-L2 <= 3e-2, sigmoid output abs <= 1e-2 (eval) and
rel-L2 <= 5e-3 (eval) / 1.5e-2 (batch-stat BN on a batch of 2), losses rel 2e-3.
Gradients are compared per tensor, rel-L2 <= 2.5e-2, with autograd through the oracle graph evaluated on
the DEVICE's own forward state (UNetOracle.override = engine.forward_state(): stored activations, BatchNorm
statistics and hence ReLU gates substituted straight-through). Reason, measured on the CPU alone: a network
of 13 BN+ReLU layers is chaotic at this precision -- between the pure-fp32 and a bf16-storage evaluation of
the SAME graph ~0.3 % of the ReLU gates flip, each layer adds ~5 % rel-L2 of unbiased gradient noise and the
deepest tensors differ by ~30 %; even a 1e-4 perturbation of one BN mean moves deep gradients by a few
percent. Pinning the forward state isolates what the backward kernels compute. Against the pure-fp32 oracle
the test additionally requires cosine similarity >= 0.9 and a norm ratio within 15 % for every kernel.
Biases of BN-followed convs (true gradient analytically zero) are checked in absolute terms."""
import math

import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L
from unet_rir_b200.engine import UNetEngine

pytestmark = pytest.mark.gpu


def _setup(B=2, kernels=3, seed=0, shape=(144, 160, 2)):
    g = torch.Generator().manual_seed(seed)
    om = O.UNetOracle(input_shape=shape, kernels=kernels)
    params = O.init_params(om.plan, seed=500)
    # make BN affine / biases non-trivial so their gradients are exercised
    for n, _, kind in om.plan:
        if kind in ("gamma",):
            params[n] = 1 + 0.2 * torch.randn(params[n].shape, generator=g)
        elif kind in ("beta", "bias"):
            params[n] = 0.1 * torch.randn(params[n].shape, generator=g)
    x = torch.rand(B, *shape, generator=g)
    y = torch.rand(B, *shape, generator=g)
    x[:, 129:], x[:, :, 151:], y[:, 129:], y[:, :, 151:] = 0, 0, 0, 0     # TensorPadder zeros
    emb = torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32)
    dim = (shape[0] // 16) * (shape[1] // 16) * 16
    mask = (torch.rand(B, dim, generator=g) > 0.3).float() / 0.7
    return om, params, x, y, emb, mask


@pytest.mark.parametrize("kernels", [3, 6])
def test_forward_eval_matches_oracle(kernels):
    om, params, x, y, emb, mask = _setup(B=2, kernels=kernels)
    eng = UNetEngine(kernels=kernels)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=False).float().cpu()
    ref = om.forward(params, x, emb, training=False)
    assert U.max_abs(out, ref) < 1e-2 and U.rel_l2(out, ref) < 5e-3


def test_train_forward_backward_matches_oracle():
    om, params, x, y, emb, mask = _setup(B=2, kernels=3)
    om.taps = {}
    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), grads32, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    taps = {k: v.detach() for k, v in om.taps.items()}
    om.taps = None
    oq = O.UNetOracle(kernels=3, emulate_bf16=True)

    results = {}
    for impl in (L.IMPL_SIMT, L.IMPL_AUTO):
        eng = UNetEngine(kernels=3, impl=impl)
        eng.load_state_dict(params)
        out = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
        dbg = eng.debug_tensors()
        for name, t in dbg.items():
            if name in taps:
                assert U.rel_l2(t, taps[name]) < 3e-2, (impl, name, U.rel_l2(t, taps[name]))
        # training-mode BN on a batch of 2 amplifies the bf16 rounding of single elements
        assert U.max_abs(out.float(), ref_out) < 3e-2 and U.rel_l2(out.float(), ref_out) < 1.5e-2
        n = 2 * 144 * 160
        losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
        assert abs(float(losses[0]) - float(loss)) < 2e-3 * float(loss)
        assert abs(float(losses[1]) - float(lp)) < 2e-3 * float(lp)
        assert abs(float(losses[2]) - float(ls)) < 2e-3 * float(ls)
        eng.backward(eng._buffers(2)["g_out"])
        torch.cuda.synchronize()
        oq.override = eng.forward_state()
        _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
        oq.override = None
        bad = []
        for name in eng.trainable_names():
            got, ref = eng.grad[name].cpu(), grads[name]
            scale = float(ref.abs().max())
            is_dead_bias = name.endswith(".b") and (".blk." in name or ".fuse" in name)
            if is_dead_bias:      # true gradient is 0 (BatchNorm removes the mean); only rounding noise
                ok = U.max_abs(got, ref) < 2e-3
            else:
                ok = U.rel_l2(got, ref) < 2.5e-2 or U.max_abs(got, ref) < 1e-7 + 1e-3 * scale
                if name.endswith(".w") and ok:      # statistical agreement with the pure-fp32 evaluation
                    r32 = grads32[name].flatten().double()
                    gd = got.flatten().double()
                    cos = float((r32 @ gd) / (r32.norm() * gd.norm() + 1e-300))
                    ok = cos > 0.9 and 0.85 < float(gd.norm() / r32.norm()) < 1.15
            if not ok:
                bad.append((name, U.rel_l2(got, ref), U.max_abs(got, ref), scale))
        assert not bad, (impl, bad)
        results[impl] = {k: v.clone() for k, v in eng.grad.items()}
        # moving statistics followed Keras momentum .99
        new_stats = {}
        om.forward(params, x, emb, training=True, dropout_mask=mask, new_stats=new_stats)
        for k, v in new_stats.items():
            assert U.rel_l2(eng.state[k].cpu(), v) < 2e-2 or U.max_abs(eng.state[k].cpu(), v) < 1e-3, k


def test_adam_step_moves_parameters_like_oracle():
    om, params, x, y, emb, mask = _setup(B=2, kernels=3, seed=1)
    eng = UNetEngine(kernels=3)
    eng.load_state_dict(params)
    st = O.new_opt_state(params, om.plan)
    p0 = {k: v.clone() for k, v in params.items()}
    lr = 1e-3
    _, ref_grads, _ = O.train_step(om, params, st, x, y, emb, lr, dropout_mask=mask, apply=True)
    eng.set_lr(lr)
    eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    n = 2 * 144 * 160
    eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    eng.backward(eng._buffers(2)["g_out"])
    eng.adam_step()
    assert int(eng.step_dev) == 1
    # the first Adam step moves every weight by ~lr*sign(g), so the update is only well conditioned
    # where |g| is not tiny: compare direction there, and the update size everywhere
    for name in ("enc3.blk.c1.w", "dec2.fuse.w", "dec5.up.w", "vec.dense.w", "head.w", "enc1.down.w"):
        du_ref = params[name] - p0[name]
        du = eng.param[name].cpu() - p0[name]
        g = ref_grads[name].abs()
        big = g > g.median()
        agree = float((torch.sign(du[big]) == torch.sign(du_ref[big])).float().mean())
        assert agree > 0.85, (name, agree)          # fp32 vs bf16 evaluation: sign agreement is statistical
        assert abs(float(du.abs().mean()) - float(du_ref.abs().mean())) < 0.05 * float(du_ref.abs().mean()), name


def test_trainer_train_epoch_loop(capsys):
    """Trainer.train (amp_phase_trainer.py:37-127): history rows [loss, phase, stft] per epoch, loss = phase + stft,
    validation rows equal model_loss of the eval forward, checkpoint / early-stopping protocol, the console lines."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(11)

    class Gen:                                   # DataGenerator contract: (spec_in, emb, spec_out) per index
        def __init__(self, n):
            self.items = [(torch.rand(2, 144, 160, 2, generator=g), torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32),
                           torch.rand(2, 144, 160, 2, generator=g)) for _ in range(n)]
        def __len__(self):
            return len(self.items)
        def __getitem__(self, i):
            return self.items[i]

    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    mc, es = ModelCheckpoint("/tmp/urir_train_loop", False, 1), EarlyStopping(2)
    tr = Trainer(0.9, 3, "adam", [mc, es], [True, 1], 1e-3, "loop")
    train_gen, val_gen = Gen(3), Gen(2)
    model, hist = tr.train(unet, train_gen, val_gen)
    out = capsys.readouterr().out
    assert model is unet and 1 <= hist.epochs <= 3
    th, vh = hist.train_loss_history, hist.val_loss_history
    assert th.shape == (hist.epochs, 3) and vh.shape == (hist.epochs, 3)
    assert np.isfinite(th).all() and np.isfinite(vh).all()
    assert np.allclose(th[:, 0], th[:, 1] + th[:, 2], rtol=1e-5) and np.allclose(vh[:, 0], vh[:, 1] + vh[:, 2], rtol=1e-5)
    assert abs(tr.learning_rate - 1e-3 * np.exp(-0.25 * (hist.epochs - 1 - 1))) < 1e-12 or hist.epochs == 1
    # the last epoch's validation row, recomputed from the final weights
    rows = []
    for i in range(len(val_gen)):
        spec_in, emb, spec_out = val_gen[i]
        rows.append([float(v) for v in tr.model_loss(spec_out, unet.model([spec_in, emb], training=False))])
    assert np.allclose(np.mean(rows, axis=0), vh[-1], rtol=1e-4)
    assert abs(mc.val_loss_min - min(10, float(vh[:, 0].min()))) < 1e-6
    assert out.count("[INFO]: Starting epoch") == hist.epochs and "Perdidas training:" in out and " - Perdidas combinadas: " in out


def test_many_graph_replays_stay_healthy():
    """Soak test: 600 replays of the captured train step. Guards the pipelines' mbarrier protocols against
    timing-dependent hangs (a two-issuer ring whose stage ownership alternated between fills once aliased mbarrier
    parities about once per thousand steps: a trapped 'unspecified launch failure') and checks that training moves."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(3)
    B = 16
    x = torch.rand(B, 144, 160, 2, generator=g); y = torch.rand(B, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_soak", False, 0), EarlyStopping(5)], [False, 0], 1e-4, "soak")
    first = last = None
    for i in range(600):
        l = tr.step(x, y, emb, unet)[0]
        if i == 0:
            first = float(l)
    torch.cuda.synchronize()
    last = float(l)
    assert last == last and last < first, (first, last)


def test_data_parallel_step_world1_matches_oracle_dp_loss():
    """DistributedTrainer (main_training.py's compute_loss: alpha-weighted amp/phase terms / global batch + the L2
    kernel regulariser of the nine strided layers) on one replica: loss value against the oracle's dp_loss, and the
    regulariser's gradient 2 * 0.001 * W present in the flat gradient (one batched launch)."""
    from unet_rir_b200.dl_models.u_net import UNet
    from unet_rir_b200.main_training import DistributedTrainer
    om, params, x, y, emb, mask = _setup(B=2, kernels=3, seed=2)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    eng = unet.model.engine
    eng.load_state_dict(params)
    dt = DistributedTrainer(unet, per_replica_batch=2, alpha=0.9, lr=0.0, loss="dp", world=1, use_cuda_graph=False,
                            dropout=False)
    loss = float(dt.train_step(x, emb, y))
    pred = om.forward(params, x, emb, training=True, dropout_mask=None)
    ref = float(O.dp_loss(y, pred, 0.9, 2, l2_losses=om.l2_losses(params), num_replicas=1))
    assert abs(loss - ref) < 3e-3 * abs(ref), (loss, ref)
    # the regulariser alone (one batched launch): reg = 0.001 * sum ||W||^2, G += 2 * 0.001 * W on the nine strided
    # kernels and nothing anywhere else
    eng.G.zero_()
    reg = float(eng.l2_loss_and_grad(1.0)[0])
    ref_reg = float(sum(om.l2_losses(params)))
    assert abs(reg - ref_reg) < 1e-4 * ref_reg, (reg, ref_reg)
    names = set(O.l2_regularised_names(om.plan))
    assert len(names) == 9
    for n in eng.trainable_names():
        g = eng.grad[n].cpu()
        if n in names:
            assert U.rel_l2(g, 2 * O.L2_COEF * params[n]) < 1e-6, n
        else:
            assert float(g.abs().max()) == 0.0, n


def test_train_step_with_default_kernel_size_6():
    """The constructor default kernels=6 (u_net.py:42): 6x6 SAME convs (pad 2,3 / 2,2) take the generic tcgen05 paths
    (no stride-2 halo / up-2 kernels). One training forward + backward against the oracle on the device's own forward
    state, same tolerances as the kernels=3 test."""
    om, params, x, y, emb, mask = _setup(B=2, kernels=6)
    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), grads32, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    eng = UNetEngine(kernels=6)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    assert U.max_abs(out.float(), ref_out) < 3e-2 and U.rel_l2(out.float(), ref_out) < 1.5e-2
    n = 2 * 144 * 160
    losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    assert abs(float(losses[0]) - float(loss)) < 2e-3 * float(loss)
    eng.backward(eng._buffers(2)["g_out"])
    torch.cuda.synchronize()
    oq = O.UNetOracle(kernels=6, emulate_bf16=True)
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    bad = []
    for name in eng.trainable_names():
        got, ref = eng.grad[name].cpu(), grads[name]
        scale = float(ref.abs().max())
        if name.endswith(".b") and (".blk." in name or ".fuse" in name):
            ok = U.max_abs(got, ref) < 2e-3
        else:
            ok = U.rel_l2(got, ref) < 2.5e-2 or U.max_abs(got, ref) < 1e-7 + 1e-3 * scale
        if not ok:
            bad.append((name, U.rel_l2(got, ref), U.max_abs(got, ref), scale))
    assert not bad, bad


def test_prefetched_inputs_give_the_same_step():
    """Trainer.prefetch (copy-stream staging of the next batch) + step == step alone; a step called with other
    objects than the prefetched ones ignores the staged copy."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(5)
    B = 4
    batches = [(torch.rand(B, 144, 160, 2, generator=g).pin_memory(), torch.rand(B, 144, 160, 2, generator=g).pin_memory(),
                torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32).pin_memory()) for _ in range(3)]
    losses = []
    for use_prefetch in (False, True):
        unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
        tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_pf", False, 0), EarlyStopping(5)], [False, 0], 1e-4, "pf")
        tr.dropout = False
        out = []
        if use_prefetch:
            tr.prefetch(*batches[0], unet)
        for i, (x, y, e) in enumerate(batches):
            l = tr.step(x, y, e, unet)[0]
            if use_prefetch and i + 1 < len(batches):
                tr.prefetch(*batches[i + 1], unet)
            out.append(float(l))
        if use_prefetch:                       # stale prefetch: announce batch 0, then step on batch 1
            tr.prefetch(*batches[0], unet)
            l_stale = float(tr.step(*batches[1], unet)[0])
            assert np.isfinite(l_stale)
        losses.append(out)
    for a, b in zip(*losses):
        assert abs(a - b) < 2e-3 * abs(a), losses      # same data, same init; BN statistics / gradients use atomics


@pytest.mark.parametrize("mode", [1, 2, 3])
def test_alternate_block_modes(mode):
    """mode 1 convolutional_block_2, 2 residual_block_1, 3 residual_block_2 (u_net.py:280-287, 324-386): eval forward
    against the fp32 oracle, then a training forward + backward against the oracle evaluated on the device's own
    forward state (same method and tolerances as the mode-0 test)."""
    g = torch.Generator().manual_seed(mode)
    om = O.UNetOracle(kernels=3, mode=mode)
    params = O.init_params(om.plan, seed=500)
    for n, _, kind in om.plan:
        if kind == "gamma":
            params[n] = 1 + 0.2 * torch.randn(params[n].shape, generator=g)
        elif kind in ("beta", "bias"):
            params[n] = 0.1 * torch.randn(params[n].shape, generator=g)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    mask = (torch.rand(2, 1440, generator=g) > 0.3).float() / 0.7
    eng = UNetEngine(kernels=3, mode=mode)
    assert eng.trainable_names() == O.trainable_names(om.plan)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=False).float().cpu()
    ref = om.forward(params, x, emb, training=False)
    assert U.max_abs(out, ref) < 2e-2 and U.rel_l2(out, ref) < 1e-2

    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), _, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    out = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    assert U.rel_l2(out.float(), ref_out) < 2.5e-2
    n = 2 * 144 * 160
    losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    assert abs(float(losses[0]) - float(loss)) < 5e-3 * float(loss)
    eng.backward(eng._buffers(2)["g_out"])
    torch.cuda.synchronize()
    oq = O.UNetOracle(kernels=3, mode=mode, emulate_bf16=True)
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    bad = []
    for name in eng.trainable_names():
        got, refg = eng.grad[name].cpu(), grads[name]
        scale = float(refg.abs().max())
        if name.endswith(".b") and (".blk." in name or ".fuse" in name):
            ok = U.max_abs(got, refg) < 3e-3
        elif name.endswith(".down.b") or name.endswith(".up.b") or name == "vec.proj.b":
            # sums over all pixels of a gradient that just left a BatchNorm backward (zero mean per channel):
            # cancellation-dominated, and these nets double / triple the number of BN layers -- direction and size
            got_d, ref_d = got.double().flatten(), refg.double().flatten()
            cos = float((got_d @ ref_d) / (got_d.norm() * ref_d.norm() + 1e-300))
            ok = U.rel_l2(got, refg) < 0.3 and cos > 0.95
        else:
            ok = U.rel_l2(got, refg) < 3e-2 or U.max_abs(got, refg) < 1e-7 + 1e-3 * scale
        if not ok:
            bad.append((name, U.rel_l2(got, refg), U.max_abs(got, refg), scale))
    assert not bad, bad


def test_without_batchnorm():
    """BatchNorm=False (constructor option, u_net.py:367 `if BatchNorm:`): conv -> ReLU through the same kernels with
    identity coefficients; forward and gradients against the oracle built the same way."""
    g = torch.Generator().manual_seed(9)
    om = O.UNetOracle(kernels=3, BatchNorm=False)
    params = O.init_params(om.plan, seed=500)
    for n, _, kind in om.plan:
        if kind == "bias":
            params[n] = 0.05 * torch.randn(params[n].shape, generator=g)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    mask = (torch.rand(2, 1440, generator=g) > 0.3).float() / 0.7
    eng = UNetEngine(kernels=3, BatchNorm=False)
    assert eng.trainable_names() == O.trainable_names(om.plan)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=False).float().cpu()
    ref = om.forward(params, x, emb, training=False)
    assert U.max_abs(out, ref) < 1e-2 and U.rel_l2(out, ref) < 5e-3
    st = O.new_opt_state(params, om.plan)
    eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    n = 2 * 144 * 160
    eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    eng.backward(eng._buffers(2)["g_out"])
    torch.cuda.synchronize()
    oq = O.UNetOracle(kernels=3, BatchNorm=False, emulate_bf16=True)
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    bad = []
    for name in eng.trainable_names():
        got, refg = eng.grad[name].cpu(), grads[name]
        scale = float(refg.abs().max())
        if not (U.rel_l2(got, refg) < 3e-2 or U.max_abs(got, refg) < 1e-7 + 2e-3 * scale):
            bad.append((name, U.rel_l2(got, refg), U.max_abs(got, refg), scale))
    assert not bad, bad


def test_long_rir_input_shape():
    """BASELINE config 5 (long-RIR sweep): 0.4 s RIRs -> 301 frames -> input_shape (144, 304, 2). Ragged tiles at every
    level (304 = 19 * 16, 19-wide bottleneck), Dense width 9 * 19 * 16 = 2736. Eval forward against the oracle and one
    training step's losses; gradients of a few tensors against the oracle on the device's forward state."""
    shape = (144, 304, 2)
    om, params, x, y, emb, mask = _setup(B=2, kernels=3, seed=4, shape=shape)
    eng = UNetEngine(input_shape=shape, kernels=3)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=False).float().cpu()
    ref = om.forward(params, x, emb, training=False)
    assert U.max_abs(out, ref) < 1e-2 and U.rel_l2(out, ref) < 5e-3
    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), _, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    n = 2 * shape[0] * shape[1]
    losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    assert abs(float(losses[0]) - float(loss)) < 3e-3 * float(loss)
    eng.backward(eng._buffers(2)["g_out"])
    torch.cuda.synchronize()
    oq = O.UNetOracle(input_shape=shape, kernels=3, emulate_bf16=True)
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    for name in ("enc1.down.w", "enc2.down.w", "enc3.blk.c1.w", "enc5.blk.c1.w", "vec.dense.w", "vec.proj.w",
                 "dec2.up.w", "dec3.fuse.w", "dec5.blk.c1.w", "head.w", "dec4.fuse_bn.gamma"):
        assert U.rel_l2(eng.grad[name].cpu(), grads[name]) < 2.5e-2, (name, U.rel_l2(eng.grad[name].cpu(), grads[name]))


# ------------------------------------------------------------------------------------------------
# round 2: parity ON THE BENCHMARKED PATH -- the graph-captured Trainer.step at the batch sizes whose AUTO dispatch
# selects the persistent halo / stride-2 halo / up-2 / halo-wgrad kernels (B = 16: the reference's per-replica batch,
# main_training.py:44; B = 64: what bench.py times), checked against the oracle like the B = 2 eager tests above.
# ------------------------------------------------------------------------------------------------
def _grad_report(eng, grads, grads32, tol_rel=2.5e-2):
    bad = []
    for name in eng.trainable_names():
        got, ref = eng.grad[name].cpu(), grads[name]
        scale = float(ref.abs().max())
        if name.endswith(".b") and (".blk." in name or ".fuse" in name):      # BN-followed conv bias: true gradient 0
            ok = U.max_abs(got, ref) < 2e-3
        else:
            ok = U.rel_l2(got, ref) < tol_rel or U.max_abs(got, ref) < 1e-7 + 1e-3 * scale
            if name.endswith(".w") and ok and grads32 is not None:
                r32, gd = grads32[name].flatten().double(), got.flatten().double()
                cos = float((r32 @ gd) / (r32.norm() * gd.norm() + 1e-300))
                ok = cos > 0.9 and 0.85 < float(gd.norm() / r32.norm()) < 1.15
        if not ok:
            bad.append((name, U.rel_l2(got, ref), U.max_abs(got, ref), scale))
    return bad


@pytest.mark.parametrize("B", [16, 64])
def test_graph_captured_step_matches_oracle_at_benchmark_batch(B):
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    om, params, x, y, emb, _ = _setup(B=B, kernels=3, seed=20 + B)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    eng = unet.model.engine
    eng.load_state_dict(params)
    tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_b16", False, 0), EarlyStopping(5)], [False, 0], 0.0, "parity")
    fam0 = L.family_calls()
    for _ in range(3):                       # eager, capture + replay, replay -- all with lr = 0: weights stay put
        tr.step(x, y, emb, unet)
    fam1 = L.family_calls()
    graph = [g for g in tr._graphs.values() if isinstance(g, torch.cuda.CUDAGraph)]
    assert len(graph) == 1, tr._graphs
    # the dispatch that was captured is the benchmark's: every persistent-halo family ran (plus the deep-layer,
    # thin and head kernels), nothing fell back to the CUDA-core path
    expect = ["halo", "halo_s2_fprop", "halo_up2", "wgrad_halo", "wgrad_halo_s2", "wgrad_tc", "thin_gemm",
              "head_fprop", "thin_wgrad"]
    ran = {k: fam1[k] - fam0[k] for k in fam1}
    assert all(ran[k] > 0 for k in expect), ran
    assert ran["igemm"] + ran["deep"] > 0 and ran["simt"] == 0, ran
    for n in eng.trainable_names():          # lr = 0 really left the masters alone
        assert torch.equal(eng.param[n].cpu(), params[n]), n
    m3, v3 = eng.M.clone(), eng.V.clone()
    p3 = eng.P.clone()
    assert int(eng.step_dev) == 3
    # ---- step 4: a graph REPLAY with a non-zero learning rate (lr lives in device memory)
    lr = 1e-3
    tr.learning_rate = lr
    l4 = [float(v) for v in tr.step(x, y, emb, unet)]
    torch.cuda.synchronize()
    mask = eng._buffers(B)["mask"].cpu()     # the Dropout mask this replay drew (device counter-based generator)
    assert 0.6 < float((mask > 0).float().mean()) < 0.8
    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), grads32, ref_out = O.train_step(om, params, st, x, y, emb, lr, dropout_mask=mask, apply=False)
    out = eng._buffers(B)["out"].float().cpu()
    assert U.max_abs(out, ref_out) < 3e-2 and U.rel_l2(out, ref_out) < 1.5e-2, (U.max_abs(out, ref_out), U.rel_l2(out, ref_out))
    for got, ref in zip(l4, (loss, lp, ls)):
        assert abs(got - float(ref)) < 2e-3 * float(ref), (l4, float(loss), float(lp), float(ls))
    oq = O.UNetOracle(kernels=3, emulate_bf16=True)
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, lr, dropout_mask=mask, apply=False)
    bad = _grad_report(eng, grads, grads32)
    assert not bad, bad
    # ---- Adam inside the replayed graph: Keras update from the device's own gradient, t = 4
    g = eng.G
    m4 = 0.9 * m3 + 0.1 * g
    v4 = 0.999 * v3 + 0.001 * g * g
    lr_t = lr * math.sqrt(1 - 0.999 ** 4) / (1 - 0.9 ** 4)
    want = p3 - lr_t * m4 / (v4.sqrt() + 1e-7)
    assert U.rel_l2(eng.M, m4) < 1e-5 and U.rel_l2(eng.V, v4) < 1e-5        # fp32 evaluation order of b*v + (1-b)*g*g
    assert float((eng.P - want).abs().max()) < 2e-7 + 1e-6 * lr, float((eng.P - want).abs().max())
    assert int(eng.step_dev) == 4
    # and the bf16 operand copies inside the graph follow the new masters
    w_ck, _ = eng.wops["dec3.fuse.w"]
    assert torch.equal(w_ck.float().cpu().reshape(-1), eng.param["dec3.fuse.w"].to(torch.bfloat16).float().cpu().reshape(-1))


def test_deterministic_mode_gives_bit_identical_steps():
    """URIR_DETERMINISTIC / urir_set_deterministic: fixed-order commits of every cross-CTA reduction (BatchNorm statistics
    from the conv epilogues, BatchNorm backward sums, weight-gradient splits, bias gradients, embedding gradient, loss).
    Two independent engines stepping the same batch must agree bit for bit in loss, output and all 77 gradients."""
    om, params, x, y, emb, mask = _setup(B=16, kernels=3, seed=31)
    prev = L.set_deterministic(True)
    try:
        runs = []
        for _ in range(2):
            eng = UNetEngine(kernels=3)
            eng.load_state_dict(params)
            eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
            n = 16 * 144 * 160
            losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n).clone()
            eng.backward(eng._buffers(16)["g_out"])
            torch.cuda.synchronize()
            runs.append((losses.cpu(), eng._buffers(16)["out"].cpu().clone(), eng.G.cpu().clone()))
        assert torch.equal(runs[0][0], runs[1][0])
        assert torch.equal(runs[0][1], runs[1][1])
        assert torch.equal(runs[0][2], runs[1][2]), float((runs[0][2] - runs[1][2]).abs().max())
    finally:
        L.set_deterministic(prev)


def test_dropout_masks_advance_without_the_adam_step_counter():
    """Dropout draws from its own device counter, advanced by every training forward: Nadam / SGD / external optimisers and
    repeated forwards all get fresh masks, also under CUDA-graph replay."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    for opt in ("nadam", "sgd", "adam"):
        unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
        eng = unet.model.engine
        tr = Trainer(0.9, 1, opt, [ModelCheckpoint("/tmp/urir_do", False, 0), EarlyStopping(5)], [False, 0], 1e-5, "do")
        masks = []
        for _ in range(4):                           # eager, capture, replay, replay
            tr.step(x, y, emb, unet)
            masks.append(eng._buffers(2)["mask"].cpu().clone())
        for i in range(3):
            assert not torch.equal(masks[i], masks[i + 1]), (opt, i)
        assert int(eng.drop_ctr_dev) == 4 and int(eng.step_dev) == 4
    # two replicas of a DistributedTrainer draw different masks (rank-mixed seed)
    from unet_rir_b200.main_training import DistributedTrainer
    ms = []
    for rank in (0, 1):
        unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
        dt = DistributedTrainer(unet, per_replica_batch=2, world=1, rank=rank, use_cuda_graph=False)
        dt.train_step(x, emb, y)
        ms.append(unet.model.engine._buffers(2)["mask"].cpu().clone())
    assert not torch.equal(ms[0], ms[1])


def test_autograd_path_with_an_external_optimiser():
    """model.model(..., training=True) under autograd + a torch optimiser on trainable_variables (the reference's
    tape.gradient / apply_gradients pair): the update must reach the bf16 operand copies, and load_weights must work
    after the leaves were created."""
    from unet_rir_b200.dl_models.u_net import UNet
    g = torch.Generator().manual_seed(8)
    x = torch.rand(2, 144, 160, 2, generator=g); y = torch.rand(2, 144, 160, 2, generator=g).cuda()
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g, dtype=torch.int32)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    eng = unet.model.engine
    before = unet.model([x, emb], training=False).clone()
    opt = torch.optim.SGD(unet.model.trainable_variables, lr=0.5)
    out = unet.model([x, emb], training=True)
    loss = ((out - y) ** 2).mean()
    loss.backward()
    assert all(v.grad is not None for v in unet.model.trainable_variables)
    opt.step()
    after = unet.model([x, emb], training=False).clone()
    assert float((after - before).abs().max()) > 1e-4           # the step changed what the kernels compute with
    w_ck, _ = eng.wops["enc2.down.w"]
    assert torch.equal(w_ck.float().cpu().reshape(-1), eng.param["enc2.down.w"].detach().to(torch.bfloat16).float().cpu().reshape(-1))
    sd = eng.state_dict()
    unet.model.save_weights("/tmp/urir_autograd_w.pt")
    unet.load_weights("/tmp/urir_autograd_w.pt")                 # in-place copy into leaves that require grad
    again = unet.model([x, emb], training=False)
    assert torch.equal(again, after)
    assert all(torch.equal(sd[k], v) for k, v in eng.state_dict().items())


def test_dp_trainer_graphs_are_keyed_by_shard_size_and_epoch_loop(tmp_path):
    """DistributedTrainer at world 1: one CUDA graph per shard size (a later step with another B must not replay the
    first graph); train_loop = the reference's epoch loop (main_training.py:332-391): validation pass through
    test_step(training=True), checkpoint at epochs 0, 2, ... with max_to_keep=2, LR decay from the given epoch."""
    from unet_rir_b200.dl_models.u_net import UNet
    from unet_rir_b200.main_training import CheckpointManager, DistributedTrainer, train_loop
    g = torch.Generator().manual_seed(12)

    def batch(B):
        return (torch.rand(B, 144, 160, 2, generator=g), torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32),
                torch.rand(B, 144, 160, 2, generator=g))

    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    dt = DistributedTrainer(unet, per_replica_batch=4, alpha=0.9, lr=1e-4, loss="dp", world=1, dropout=False)
    b4, b2 = batch(4), batch(2)
    l4 = [float(dt.train_step(*b4)) for _ in range(3)]
    l2 = [float(dt.train_step(*b2)) for _ in range(3)]
    assert set(dt._graphs) == {4, 2} and all(isinstance(v, torch.cuda.CUDAGraph) for v in dt._graphs.values())
    assert all(np.isfinite(l4 + l2))
    # the B = 2 steps really ran on the B = 2 batch: compare with an eager trainer from the same state
    unet_b = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    unet_b.model.engine.load_state_dict(unet.model.engine.state_dict())
    eb = DistributedTrainer(unet_b, per_replica_batch=4, alpha=0.9, lr=1e-4, loss="dp", world=1, dropout=False, use_cuda_graph=False)
    a, b = float(dt.train_step(*b2)), float(eb.train_step(*b2))
    assert abs(a - b) < 2e-3 * abs(b), (a, b)

    class Gen:
        def __init__(self, n): self.items = [batch(4) for _ in range(n)]
        def __len__(self): return len(self.items)
        def __getitem__(self, i): return self.items[i]

    mgr = CheckpointManager(dt, str(tmp_path), max_to_keep=2)
    hist = train_loop(dt, Gen(2), Gen(1), n_epochs=3, manager=mgr, lr_exp_decay=(True, 1), verbose=False)
    assert [h["epoch"] for h in hist] == [1, 2, 3]
    assert hist[0]["checkpoint"] and hist[1]["checkpoint"] is None and hist[2]["checkpoint"]
    assert sorted(f for f in __import__("os").listdir(tmp_path)) == ["ckpt-1.pt", "ckpt-2.pt"]
    assert abs(hist[0]["lr"] - 1e-4) < 1e-10 and abs(hist[2]["lr"] - 1e-4 * 0.9 ** 2) < 1e-10      # lr * 0.9^(epoch/start), fp32 on the device
    assert all(np.isfinite([h["loss"], h["train_mse"], h["train_phase"], h["val_mse"], h["val_phase"]]).all() for h in hist)
    moved = unet.model.engine.state["enc1.blk.bn1.moving_mean"].clone()
    dt.test_step(*batch(4))                                      # training=True: the moving statistics move (:300)
    assert not torch.equal(moved, unet.model.engine.state["enc1.blk.bn1.moving_mean"])
    st = dt.checkpoint_state()
    dt.load_checkpoint_state(st)
    assert int(unet.model.engine.step_dev) == st["step"]


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under `gpurun --gpus 2`; log kept in profiles/)")
def test_two_gpu_dp_gradient_equals_sum_of_shard_gradients():
    """tools/dp_parity.py at world 2 over NCCL: the flat gradient after the bucketed all-reduce == sum of the per-shard
    engine gradients (1e-5), bit-identical on both ranks, graph replay == eager, and != the full-batch-BatchNorm gradient."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(root, "tools", "dp_parity.py"), "--batch", "8"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_convergence_curve_tracks_the_fp32_oracle():
    """Loss-trajectory parity: 100 Adam steps on ONE fixed structured batch -- spectrograms of synthetic
    RIRs through the real STFT feature path, not U(0,1) noise -- GPU bf16 graph-captured Trainer.step against the fp32
    CPU oracle's Trainer.step from the same weights, with the SAME Dropout mask each step (read back from the device).
    Band (stated here, measured on B200, curve kept in profiles/r02_convergence.json): both losses fall below a fifth
    of their starting value (measured: 1.123 -> 0.110 and 0.107, a factor of ten), the two curves stay within 8 % of each
    other at every step and within 2.5 % on average (measured: 5.9 % worst, at step 93 where the loss is 0.11 and
    Dropout makes consecutive steps differ by as much; 1.3 % mean; 1 - 1.7 % over the first 40 steps).
    Runs with the fixed-order reductions (urir_set_deterministic): with plain atomics the run-to-run spread of the
    gradients moves the late part of the curve by a few per cent from run to run -- one run in about ten left the
    1.5 % early band this test first had -- and a band test should not depend on the arrival order of atomics."""
    import json, os
    from oracle import signal_oracle as SO
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
    from unet_rir_b200.dl_models.u_net import UNet
    from unet_rir_b200.preprocess import preprocess_batch
    B, steps, lr = 8, 100, 2e-4
    rng = np.random.default_rng(21)
    x = preprocess_batch(SO.synthetic_rir(B, rng)).cpu()
    y = preprocess_batch(SO.synthetic_rir(B, rng)).cpu()
    g = torch.Generator().manual_seed(21)
    emb = torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32)
    om = O.UNetOracle(kernels=3)
    params = O.init_params(om.plan, seed=500)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    eng = unet.model.engine
    eng.load_state_dict(params)
    tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_conv", False, 0), EarlyStopping(5)], [False, 0], lr, "conv")
    st = O.new_opt_state(params, om.plan)
    gpu, cpu = [], []
    prev_det = L.set_deterministic(True)
    try:
        for i in range(steps):
            l = tr.step(x, y, emb, unet)
            gpu.append([float(v) for v in l])
            mask = eng._buffers(B)["mask"].cpu()
            (lo, lp, ls), _, _ = O.train_step(om, params, st, x, y, emb, lr, dropout_mask=mask, apply=True)
            cpu.append([float(lo), float(lp), float(ls)])
    finally:
        L.set_deterministic(prev_det)
    gpu, cpu = np.array(gpu), np.array(cpu)
    rel = np.abs(gpu[:, 0] - cpu[:, 0]) / cpu[:, 0]
    out = {"batch": B, "steps": steps, "lr": lr, "optimizer": "adam", "data": "STFT features of synthetic RIRs (signal path kernels)",
           "gpu_loss": gpu[:, 0].tolist(), "oracle_loss": cpu[:, 0].tolist(), "gpu_phase": gpu[:, 1].tolist(),
           "oracle_phase": cpu[:, 1].tolist(), "gpu_amp": gpu[:, 2].tolist(), "oracle_amp": cpu[:, 2].tolist(),
           "max_rel_diff": float(rel.max()), "mean_rel_diff": float(rel.mean())}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "convergence_curve.json"), "w") as f:
        json.dump(out, f)
    assert gpu[-1, 0] < 0.2 * gpu[0, 0] and cpu[-1, 0] < 0.2 * cpu[0, 0], (gpu[0, 0], gpu[-1, 0], cpu[0, 0], cpu[-1, 0])
    assert rel.max() < 0.08 and rel.mean() < 0.025 and rel[:40].max() < 0.025, (float(rel.max()), float(rel.mean()), float(rel[:40].max()))
