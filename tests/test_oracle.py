"""CPU suite, part 1: the oracle against everything that can pin it here.

  * the reference's own pure-numpy / pure-Python code, executed by tests/golden/make_golden.py
    (Normalizer, TensorPadder, sigmoid, ModelCheckpoint / EarlyStopping decision traces);
  * structural known answers the reference holds: 77 trainable tensors / 20,955,010 parameters of the
    canonical model, 9600 samples -> (129, 151) -> (144, 160) (postprocess.py:54, dataset.py:70);
  * an independent implementation of the published STFT / iSTFT algorithm (torch.stft / torch.istft);
  * self-consistency: SAME-padding arithmetic, Conv2DTranspose == input-gradient of the SAME conv,
    normalise o denormalise = identity, STFT -> iSTFT round trip (preprocess.py:201-205).
The U-Net, STFT and loss restatements remain "parity unpinned" w.r.t. TensorFlow / librosa binaries.
"""
import json
import math
import os

import numpy as np
import torch

from oracle import signal_oracle as SO
from oracle import unet_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_normalizer_padder_sigmoid_match_reference_outputs():
    g = np.load(os.path.join(GOLD, "preprocess_golden.npz"))
    a_n, p_n = SO.normalize(g["amp"], g["phase"])
    assert np.array_equal(a_n, g["amp_norm"]) and np.array_equal(p_n, g["phase_norm"])
    a_d, p_d = SO.denormalize(g["amp_norm"], g["phase_norm"])
    assert np.array_equal(a_d, g["amp_denorm"]) and np.array_equal(p_d, g["phase_denorm"])
    assert np.array_equal(SO.pad(g["amp_norm"]), g["amp_pad"]) and np.array_equal(SO.pad(g["phase_norm"]), g["phase_pad"])
    assert SO.pad(g["amp_norm"]).dtype == g["amp_pad"].dtype == np.float64      # np.r_ with np.zeros promotes
    a_u, p_u = SO.un_pad(g["amp_pad"], g["phase_pad"], (129, 151))
    assert np.array_equal(a_u, g["amp_unpad"]) and np.array_equal(p_u, g["phase_unpad"])
    assert np.array_equal(SO.pad(g["big"]), g["big_out"])                       # larger input returned unchanged
    assert abs(float(g["amp_norm"][0, 0])) < 1e-6                                # |S| = 0 -> the -100 dB floor (fp32 rounding)


def test_product_host_classes_match_reference_outputs():
    """The product's numpy-facing classes (same names as the reference) against the same fixtures."""
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint
    from unet_rir_b200.preprocess import Normalizer, TensorPadder, sigmoid
    g = np.load(os.path.join(GOLD, "preprocess_golden.npz"))
    nz = Normalizer()
    a_n, p_n = nz.normalize(g["amp"], g["phase"])
    assert np.array_equal(a_n, g["amp_norm"]) and np.array_equal(p_n, g["phase_norm"])
    a_d, p_d = nz.denormalize(g["amp_norm"], g["phase_norm"])
    assert np.array_equal(a_d, g["amp_denorm"]) and np.array_equal(p_d, g["phase_denorm"])
    a_p, p_p = TensorPadder((144, 160)).pad_amp_phase(g["amp_norm"], g["phase_norm"])
    assert np.array_equal(a_p, g["amp_pad"]) and np.array_equal(p_p, g["phase_pad"]) and a_p.dtype == np.float64
    a_u, p_u = TensorPadder.un_pad(a_p, p_p, (129, 151))
    assert np.array_equal(a_u, g["amp_unpad"]) and np.array_equal(p_u, g["phase_unpad"])
    assert np.array_equal(TensorPadder((144, 160)).transform(g["big"]), g["big_out"])
    assert np.allclose(sigmoid(0.5, (144, 160)), g["sigmoid"], atol=1e-7)

    tr = json.load(open(os.path.join(GOLD, "callbacks_golden.json")))

    class FakeModel:
        saved = 0

        def save(self, path):
            self.saved += 1

    mc, es, fm = ModelCheckpoint("ckpt", True, 0), EarlyStopping(tr["patience"]), FakeModel()
    for row in tr["trace"]:
        imp = mc.checkpoint(train_loss=row["train"], val_loss=row["val"], model=fm)
        stop = es.stop_count(improve=imp)
        assert (bool(imp), bool(stop), es.count, mc.val_loss_min, mc.train_loss_min, fm.saved) == \
               (row["improve"], row["stop"], row["count"], row["val_min"], row["train_min"], row["saved"])


def test_model_inventory_matches_keras_summary_numbers():
    m = O.UNetOracle(kernels=3)
    names = O.trainable_names(m.plan)
    p = O.init_params(m.plan)
    assert len(names) == 77
    assert sum(p[n].numel() for n in names) == 20_955_010
    assert sum(t.numel() for t in p.values()) == 20_958_914
    assert p["vec.dense.w"].shape == (8192, 1440) and p["vec.emb"].shape == (2000, 256)
    assert len(O.l2_regularised_names(m.plan)) == 9
    from unet_rir_b200 import plan as PL
    mine = PL.layer_plan(kernels=3)
    assert [(n, tuple(s), k) for n, s, k in mine] == [(n, tuple(s), k) for n, s, k in m.plan]
    for mode in (1, 2, 3):
        assert [(n, tuple(s)) for n, s, _ in PL.layer_plan(kernels=3, mode=mode)] == \
               [(n, tuple(s)) for n, s, _ in O.layer_plan(kernels=3, mode=mode)]
    ki = PL.keras_init(mine, seed=500)
    assert all(torch.equal(ki[n], p[n]) for n in p)


def test_same_padding_arithmetic():
    assert O.same_pad(144, 3, 1) == (144, 1, 1)
    assert O.same_pad(144, 3, 2) == (72, 0, 1)      # even input, k3 s2: pad only after
    assert O.same_pad(9, 3, 2) == (5, 1, 1)
    assert O.same_pad(160, 6, 1) == (160, 2, 3)
    assert O.same_pad(160, 6, 2) == (80, 2, 2)
    from unet_rir_b200._lib import same_pad
    for n in (9, 10, 18, 144, 160):
        for k in (1, 2, 3, 6):
            for s in (1, 2):
                out, before, _ = O.same_pad(n, k, s)
                assert same_pad(n, k, s) == (out, before)


def test_conv_transpose_is_input_gradient_of_same_conv():
    g = torch.Generator().manual_seed(0)
    for k in (2, 3, 6):
        x = torch.randn(2, 5, 6, 7, generator=g, dtype=torch.float64)                 # small grid, NCHW
        w = torch.randn(k, k, 4, 5, generator=g, dtype=torch.float64)                 # (kh,kw,out=4,in=5)
        y = O.conv2d_transpose_same(x, w, None, 2)
        assert y.shape == (2, 4, 12, 14)
        big = torch.randn(2, 4, 12, 14, generator=g, dtype=torch.float64, requires_grad=True)
        fwd = O.conv2d_same(big, w, None, 2)                                          # HWIO with I=4, O=5
        gin, = torch.autograd.grad(fwd, big, x)
        assert float((gin - y).abs().max()) < 1e-12


def test_stft_shape_and_independent_implementation():
    rng = np.random.default_rng(1)
    x = SO.remove_mean(SO.synthetic_rir(1, rng)[0])
    S = SO.stft(x)
    assert S.shape == (129, 151) and S.dtype == np.complex64
    w = torch.hann_window(128, periodic=True, dtype=torch.float64)
    for mode in ("constant", "reflect"):
        St = torch.stft(torch.tensor(x, dtype=torch.float64), 256, 64, 128, window=w, center=True, pad_mode=mode,
                        return_complex=True).numpy()
        assert np.abs(SO.stft(x, pad_mode=mode) - St).max() < 1e-5
    St = torch.stft(torch.tensor(x, dtype=torch.float64), 256, 64, 128, window=w, center=True, pad_mode="constant",
                    return_complex=True)
    yt = torch.istft(St, 256, 64, 128, window=w, center=True, length=9600).numpy()
    assert np.abs(SO.istft(S) - yt).max() < 1e-6
    f = SO.preprocess(x)
    assert f.shape == (144, 160, 2) and f.dtype == np.float32
    assert f[129:].max() == 0 and f[:, 151:].max() == 0


def test_round_trip_misalignment():
    """preprocess.py:201-205 prints this number; through normalise/pad/un-pad/denormalise it must stay tiny."""
    rng = np.random.default_rng(2)
    for x in SO.synthetic_rir(3, rng):
        x0 = SO.remove_mean(x)
        y = SO.post_process(SO.preprocess(x))
        assert y.shape == (9600,)
        assert 20 * math.log10(np.linalg.norm(y - x0) / np.linalg.norm(x0)) < -100.0


def test_losses_and_adam_conventions():
    g = torch.Generator().manual_seed(3)
    yt, yp = torch.rand(2, 8, 8, 2, generator=g), torch.rand(2, 8, 8, 2, generator=g)
    loss, lp, ls = O.amp_phase_loss(yt, yp)
    assert abs(float(ls) - float(((yt[..., 0] - yp[..., 0]) ** 2).mean())) < 1e-7
    assert abs(float(loss) - float(lp + ls)) < 1e-7
    assert float(O.amp_phase_loss(yt, yt)[0]) < 1e-7
    # phase loss is periodic: a shift by one full turn changes nothing
    sh = yp.clone(); sh[..., 1] += 1.0
    assert abs(float(O.amp_phase_loss(yt, sh)[1]) - float(lp)) < 1e-5
    # DP loss: alpha-weighted, / (H*W*2), / global batch; wrap does not change the value
    d = O.dp_loss(yt, yp, 0.9, 4)
    amp = ((yt[..., 0] - yp[..., 0]) ** 2).sum()
    ph = (1 - torch.cos(2 * math.pi * (yt[..., 1] - yp[..., 1]))).sum()
    assert abs(float(d) - float((0.9 * amp + 0.1 * ph) / (8 * 8 * 2) / 4)) < 1e-6
    # Keras Adam: first step moves by lr * g / (|g| + eps*sqrt(1-b2)/...) ~ lr * sign(g)
    p = {"w": torch.tensor([1.0, -2.0])}; gr = {"w": torch.tensor([0.5, -4.0])}
    m = {"w": torch.zeros(2)}; v = {"w": torch.zeros(2)}
    O.keras_adam_step(p, gr, m, v, 1, 0.1)
    assert torch.allclose(p["w"], torch.tensor([0.9, -1.9]), atol=1e-5)


def test_oracle_train_step_decreases_loss():
    torch.manual_seed(0)
    om = O.UNetOracle(input_shape=(32, 32, 2), kernels=3, number_filters_0=8)
    params = O.init_params(om.plan, seed=1)
    st = O.new_opt_state(params, om.plan)
    g = torch.Generator().manual_seed(5)
    x, y = torch.rand(2, 32, 32, 2, generator=g), torch.rand(2, 32, 32, 2, generator=g)
    emb = torch.randint(0, 2000, (2, 2, 16), generator=g)
    first = None
    for _ in range(6):
        (loss, _, _), _, _ = O.train_step(om, params, st, x, y, emb, 1e-2)
        first = first if first is not None else float(loss)
    assert float(loss) < first


def test_room_embeddings_match_reference_rooms_py():
    """unet_rir_b200.rooms against the reference's own rooms.py, executed from its source by tests/golden/make_golden.py:
    1440 (room, zone, array type, loudspeaker, microphone) combinations, every one of the 16 entries equal."""
    import json
    from unet_rir_b200 import rooms as R
    d = json.load(open(os.path.join(GOLD, "rooms_golden.json")))
    assert len(d["cases"]) == 1440
    for c in d["cases"]:
        got = R.uts_room(c["characteristics"][0]).return_embedding(c["characteristics"])
        assert [float(v) for v in got] == c["embedding"], c
    assert R.return_room([994]) == "Large" and R.return_room([1]) is None
    assert max(max(c["embedding"]) for c in d["cases"]) < 2000          # fits Embedding(2000, 256), u_net.py:257


def test_generic_trainer_callbacks_match_reference_trace():
    """trainer.py:175-205 executed from the reference source (tests/golden/make_golden.py): `improve` needs min_delta, the
    running minimum / saving do not."""
    import json
    from unet_rir_b200.trainer import EarlyStopping, ModelCheckpoint
    tr = json.load(open(os.path.join(GOLD, "callbacks_golden.json")))["generic_trainer"]

    class FakeModel:
        saved = 0

        def save(self, path):
            self.saved += 1

    mc, es, fm = ModelCheckpoint("ckpt", True, 0, tr["min_delta"]), EarlyStopping(tr["patience"]), FakeModel()
    seen = set()
    for row in tr["trace"]:
        imp = mc.checkpoint(train_loss=row["train"], val_loss=row["val"], model=fm)
        stop = es.stop_count(improve=imp)
        assert (bool(imp), bool(stop), es.count, mc.val_loss_min, mc.train_loss_min, fm.saved) == \
               (row["improve"], row["stop"], row["count"], row["val_min"], row["train_min"], row["saved"])
        seen.add((row["improve"], row["saved"] > 0))
    assert any(not r["improve"] and r["val_min"] == r["val"] for r in tr["trace"])     # a saved-but-not-improved epoch occurs


def test_diffunet_plan_known_answers():
    """DiffUNet (dl_models/diff_u_net.py:205-322): the oracle's and the package's variable inventories agree and add up to
    the hand-computed parameter count (see tests/test_gpu_diffunet.py for the breakdown); the oracle graph runs and its
    linear head leaves the sigmoid range."""
    import numpy as np
    import torch
    from oracle import unet_oracle as O
    from unet_rir_b200 import plan as PL
    om = O.UNetOracle(input_shape=(144, 160, 2), kernels=2, arch="diff")
    mine = PL.layer_plan((144, 160, 2), (2, 16), 0, 32, 2, True, arch="diff")
    assert [(n, tuple(s), k) for n, s, k in mine] == [(n, tuple(s), k) for n, s, k in om.plan]
    assert sum(int(np.prod(s)) for n, s, k in om.plan if k in O.TRAINABLE_KINDS) == 195_874_786
    assert sum(1 for n, s, k in om.plan if k in O.TRAINABLE_KINDS) == 75
    # UNet's own inventory is untouched by the arch switch (20,958,914 parameters including the BN moving statistics)
    assert sum(int(np.prod(s)) for n, s, k in PL.layer_plan(kernels=3)) == 20_958_914
    p = O.init_params(om.plan, seed=1)
    g = torch.Generator().manual_seed(0)
    y = om.forward(p, torch.rand(1, 144, 160, 2, generator=g), torch.randint(0, 1500, (1, 2, 16), generator=g, dtype=torch.int32))
    assert y.shape == (1, 144, 160, 2) and (float(y.min()) < 0 or float(y.max()) > 1)
