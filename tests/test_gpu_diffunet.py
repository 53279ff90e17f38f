"""GPU parity of the sibling model DiffUNet (dl_models/diff_u_net.py: 2x2 strided / transposed convolutions, 3x3 fuse
convolution, Embedding(1500, 128) -> Dense(46080) -> Dropout(.5) added to the bottleneck, linear 1x1 head) against the
fp32 oracle of the same graph: eval forward, one training forward / loss / backward (gradients with the device's forward
state pinned, as in test_gpu_model.py), and one step of the generic trainer (trainer.py) through the class interface.
Tolerances as for UNet; the head is linear, so outputs are compared by rel-L2 only."""
import numpy as np
import pytest
import torch

from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L
from unet_rir_b200.engine import UNetEngine

pytestmark = pytest.mark.gpu
SHAPE = (144, 160, 2)


def _setup(B=2, seed=3):
    g = torch.Generator().manual_seed(seed)
    om = O.UNetOracle(input_shape=SHAPE, kernels=2, arch="diff")
    params = O.init_params(om.plan, seed=500)
    for n, _, kind in om.plan:
        if kind == "gamma":
            params[n] = 1 + 0.2 * torch.randn(params[n].shape, generator=g)
        elif kind in ("beta", "bias"):
            params[n] = 0.1 * torch.randn(params[n].shape, generator=g)
    x = torch.rand(B, *SHAPE, generator=g)
    y = torch.rand(B, *SHAPE, generator=g)
    emb = torch.randint(0, 1500, (B, 2, 16), generator=g, dtype=torch.int32)
    dim = 9 * 10 * 512
    mask = (torch.rand(B, dim, generator=g) > 0.5).float() / 0.5
    return om, params, x, y, emb, mask


def test_diff_plan_matches_the_reference_graph():
    om = O.UNetOracle(input_shape=SHAPE, kernels=2, arch="diff")
    shapes = {n: s for n, s, _ in om.plan}
    assert shapes["enc1.down.w"] == (2, 2, 2, 32) and shapes["enc5.down.w"] == (2, 2, 256, 512)
    assert shapes["dec2.up.w"] == (2, 2, 256, 512) and shapes["dec2.fuse.w"] == (3, 3, 512, 256)
    assert shapes["vec.emb"] == (1500, 128) and shapes["vec.dense.w"] == (4096, 46080)
    assert shapes["head.w"] == (1, 1, 32, 2) and "vec.proj.w" not in shapes
    eng = UNetEngine(kernels=2, arch="diff")
    assert [(n, tuple(s)) for n, s, _ in eng.plan] == [(n, tuple(s)) for n, s, _ in om.plan]
    trainable = sum(int(np.prod(s)) for n, s, k in om.plan if k in O.TRAINABLE_KINDS)
    # by hand (kernels + biases): strided 2x2 convs 288 + 8,256 + 32,896 + 131,328 + 524,800 = 697,568; 2x2 ConvTs 524,544 +
    # 131,200 + 32,832 + 8,224 = 696,800; nine 3x3 block convs 3,927,488; four 3x3 fuse convs 1,179,904 + 295,040 + 73,792 +
    # 18,464 = 1,567,200; 13 BN (gamma, beta) 2 x 1,952 = 3,904; Embedding 192,000; Dense 4096 x 46080 + 46080 = 188,789,760;
    # head 66
    assert trainable == 697_568 + 696_800 + 3_927_488 + 1_567_200 + 3_904 + 192_000 + 188_789_760 + 66 == 195_874_786, trainable
    assert sum(1 for n, s, k in om.plan if k in O.TRAINABLE_KINDS) == 75


def test_diff_forward_eval_matches_oracle():
    om, params, x, y, emb, mask = _setup()
    eng = UNetEngine(kernels=2, arch="diff")
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=False).float().cpu()
    ref = om.forward(params, x, emb, training=False)
    assert U.rel_l2(out, ref) < 5e-3, U.rel_l2(out, ref)
    assert float(ref.min()) < 0.0 or float(ref.max()) > 1.0           # linear head: not a sigmoid range


def test_diff_train_forward_backward_matches_oracle():
    om, params, x, y, emb, mask = _setup()
    om.taps = {}
    st = O.new_opt_state(params, om.plan)
    (loss, lp, ls), grads32, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    taps = {k: v.detach() for k, v in om.taps.items()}
    om.taps = None
    oq = O.UNetOracle(input_shape=SHAPE, kernels=2, arch="diff", emulate_bf16=True)
    for impl in (L.IMPL_SIMT, L.IMPL_AUTO):
        eng = UNetEngine(kernels=2, arch="diff", impl=impl)
        eng.load_state_dict(params)
        tc0 = L.launch_count(1)
        out = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
        for name, t in eng.debug_tensors().items():
            if name in taps:
                assert U.rel_l2(t, taps[name]) < 3e-2, (impl, name, U.rel_l2(t, taps[name]))
        # linear head: no sigmoid to compress the bf16 rounding carried by d5 (per-layer taps are held to 3e-2 above) and
        # no 0.5 offset inflating the norm the error is divided by -- UNet's 1.5e-2 bound applies to sigmoid outputs
        assert U.rel_l2(out.float().cpu(), ref_out) < 3e-2
        n = 2 * 144 * 160
        losses = eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
        assert abs(float(losses[0]) - float(loss)) < 3e-3 * float(loss)
        eng.backward(eng._buffers(2)["g_out"])
        torch.cuda.synchronize()
        if impl == L.IMPL_AUTO:
            assert L.launch_count(1) - tc0 >= 40          # the 2x2 / 1x1 layers run on the tcgen05 kernels too
        oq.override = eng.forward_state()
        _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
        oq.override = None
        bad = []
        for name in eng.trainable_names():
            got, ref = eng.grad[name].cpu(), grads[name]
            scale = float(ref.abs().max())
            if name.endswith(".b") and (".blk." in name or ".fuse" in name):      # true gradient is 0 behind BatchNorm
                ok = U.max_abs(got, ref) < 2e-3
            elif name.endswith(".down.b") or name.endswith(".up.b"):
                # nearly dead: the bias feeds conv -> BatchNorm, whose backward output sums to zero per channel, so the true
                # gradient is a border effect only (|g| ~ 2e-2 here) while the bf16 rounding of that zero-sum tensor leaves a
                # random pixel-sum of the same kind as for the dead biases above (measured 2-3e-3, varying run to run)
                ok = U.rel_l2(got, ref) < 2.5e-2 or U.max_abs(got, ref) < 5e-3
            else:
                ok = U.rel_l2(got, ref) < 2.5e-2 or U.max_abs(got, ref) < 1e-7 + 1e-3 * scale
            if not ok:
                bad.append((name, U.rel_l2(got, ref), U.max_abs(got, ref), scale))
        assert not bad, (impl, bad)


def test_diff_class_interface_and_generic_trainer_step(tmp_path):
    from unet_rir_b200.dl_models.diff_u_net import DiffUNet
    from unet_rir_b200.trainer import EarlyStopping, ModelCheckpoint, Trainer
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, *SHAPE, generator=g); y = torch.rand(2, *SHAPE, generator=g)
    emb = torch.randint(0, 1500, (2, 2, 16), generator=g, dtype=torch.int32)
    net = DiffUNet(input_shape=SHAPE, inf_vector_shape=(2, 16), mode=0, number_filters_0=32)
    assert len(net.model.trainable_variables) == 75           # UNet's 77 minus the 1x1 projection's kernel and bias
    tr = Trainer(0.9, 1, "adam", [ModelCheckpoint(str(tmp_path / "ck"), False, 0), EarlyStopping(5)], [False, 0], 1e-3, "diff")
    tr.dropout = False
    l0 = float(tr.step(x, y, emb, net)[0])
    out = net.model.engine._buffers(2)["out"].cpu()
    assert abs(l0 - float(((out - y) ** 2).mean())) < 1e-5 * max(1.0, l0)
    for _ in range(30):
        l1 = float(tr.step(x, y, emb, net)[0])
    assert np.isfinite(l1) and l1 < 0.7 * l0, (l0, l1)
    # save / load round trip keeps the six-entry parameter list of the reference and the weights
    net.save(str(tmp_path / "m"))
    net2 = DiffUNet.load(str(tmp_path / "m"))
    a = net.model([x, emb], training=False); b = net2.model([x, emb], training=False)
    assert torch.equal(a, b)
    assert float(DiffUNet.rmse_coef(y, a.cpu())) > 0 and float(DiffUNet.l1_norm(y, y)) == 0.0
