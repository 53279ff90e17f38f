"""CPU suite, part 2: host logic and the C-ABI surface (no device work).

  * liburir.so loads without a GPU and exports every function include/urir.h declares;
  * ConvDesc / StftDesc ctypes layouts match the C structs;
  * plan / parameter inventory, optimizer-name matching, LR schedules, batch contract, bucket ranges;
  * world-size-2 gloo run of the data-parallel arithmetic (sharding, bucketed SUM all-reduce, loss
    scaling, per-replica BatchNorm) against a single process, using the CPU oracle as the model.
"""
import ctypes
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "urir.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(urir_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    from unet_rir_b200 import _lib as L
    lib = L.load()
    declared = _header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/urir.h but not exported"
    assert sorted(L.EXPORTED_SYMBOLS) == declared          # the Python binding covers the whole header
    assert lib.urir_version() == 100
    assert L.launch_count(0) >= 0 and L.last_error() == ""


def test_ctypes_struct_layouts_match_header():
    from unet_rir_b200 import _lib as L
    src = open(os.path.join(ROOT, "include", "urir.h")).read()

    def fields(struct):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), src, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in re.findall(r"int32_t\s+([^;]+);", body):
            out += [f.strip() for f in decl.split(",")]
        return out

    assert [f for f, _ in L.ConvDesc._fields_] == fields("urir_conv_desc")
    assert [f for f, _ in L.StftDesc._fields_] == fields("urir_stft_desc")
    assert ctypes.sizeof(L.ConvDesc) == 4 * len(L.ConvDesc._fields_)


def test_bad_arguments_are_rejected_without_a_device():
    from unet_rir_b200 import _lib as L
    lib = L.load()
    d = L.ConvDesc(1, 8, 8, 32, 32, 3, 3, 3, 1, 1, 8, 8, 32, 0, 32, 0, L.BF16, L.BF16, 0, 0, 0)   # stride 3
    rc = lib.urir_conv2d_fprop(ctypes.byref(d), 1, 1, 1, None, 1, None, None)
    assert rc == -1 and "stride" in L.last_error()
    d.stride = 1
    d.x_ld = 16                                                                                     # slice > pitch
    assert lib.urir_conv2d_wgrad(ctypes.byref(d), 1, 1, 1, None) == -1
    assert lib.urir_conv_path(ctypes.byref(L.ConvDesc(1, 8, 8, 32, 32, 3, 3, 1, 1, 1, 8, 8, 32, 0, 32, 0, 1, 1, 0, 0, 0)), 0) == 1
    assert lib.urir_conv_path(ctypes.byref(L.ConvDesc(1, 8, 8, 2, 32, 3, 3, 1, 1, 1, 8, 8, 2, 0, 32, 0, 0, 1, 0, 0, 0)), 0) == 0
    with pytest.raises(L.UrirError):
        L.check(-1, "probe")


def test_trainer_host_logic():
    from unet_rir_b200.amp_phase_trainer import EarlyStopping, History, ModelCheckpoint, Trainer
    cbs = [ModelCheckpoint("x", False, 0), EarlyStopping(2)]
    assert Trainer(0.5, 3, "nadam", cbs, [True, 1], 1e-3, "f").optimizer == "nadam"      # "nadam" contains "adam"
    assert Trainer(0.5, 3, "my_adam_v2", cbs, [True, 1], 1e-3, "f").optimizer == "adam"
    assert Trainer(0.5, 3, "sgd", cbs, [True, 1], 1e-3, "f").optimizer == "sgd"
    with pytest.raises(ValueError):
        Trainer(0.5, 3, "rmsprop", cbs, [True, 1], 1e-3, "f")
    t = Trainer(0.5, 4, "adam", cbs, [True, 2], 1e-3, "f")
    assert t.train_loss_history.shape == (4, 3) and t.train_loss_history.dtype == np.float32
    assert t.alpha == 0.5 and t.lr_exp_decay is True and t.lr_exp_decay_epoch == 2
    h = History(2, np.zeros((2, 3)), np.ones((2, 3)))
    assert h.epochs == 2 and h.train_acc_history is None
    from unet_rir_b200.main_training import gradient_buckets, lr_schedule, shard_batch
    assert lr_schedule(5e-7, 10) == 5e-7 and abs(lr_schedule(5e-7, 80) - 5e-7 * 0.9) < 1e-18
    assert abs(lr_schedule(5e-7, 160) - 5e-7 * 0.81) < 1e-18
    a = np.arange(16).reshape(8, 2)
    assert np.array_equal(shard_batch([a], 1, 4)[0], a[2:4])


def test_flat_parameter_layout_and_buckets():
    from unet_rir_b200 import plan as PL
    from unet_rir_b200.main_training import gradient_buckets
    plan = PL.layer_plan(kernels=3)
    off, offsets = 0, {}
    for name, shape, kind in plan:
        if kind in PL.TRAINABLE_KINDS:
            n = int(np.prod(shape)); offsets[name] = (off, n); off = (off + n + 3) // 4 * 4
    b = gradient_buckets(offsets, off)
    assert b[0][1] == off and b[0][0] == b[1][1] and b[1][0] == b[2][1] and b[2][0] == 0
    names0 = [n for n, (o, _) in offsets.items() if b[0][0] <= o < b[0][1]]
    assert names0[0] == "dec2.up.w" and names0[-1] == "head.b"
    names1 = [n for n, (o, _) in offsets.items() if b[1][0] <= o < b[1][1]]
    assert names1 == ["vec.emb", "vec.dense.w", "vec.dense.b", "vec.proj.w", "vec.proj.b"]
    assert sum(PL.l2_regularised(n) for n in offsets) == 9
    assert all(o % 4 == 0 for o, _ in offsets.values())            # 16-byte aligned for vector reductions


class _FakeDataset:
    seed = 500

    def __init__(self, n=40):
        self.index_in = list(range(n)); self.index_out = list(range(n))[::-1]
        self.amp = np.random.default_rng(0).random((n, 144, 160)).astype(np.float64)
        self.emb = np.arange(n * 16).reshape(n, 16) % 1999

    def __getitem__(self, i):
        return self.amp[i], self.amp[i] * 0.5, self.emb[i]

    def return_characteristics(self):
        return [["Room", "A", "Planar", i, i] for i in range(len(self.index_in))]


def test_data_generator_batch_contract():
    from unet_rir_b200.datageneratorv2 import DataGenerator
    ds = _FakeDataset(40)
    tr = DataGenerator(ds, batch_size=4, partition="train", shuffle=False)
    va = DataGenerator(ds, batch_size=4, partition="val", shuffle=False)
    te = DataGenerator(ds, batch_size=4, partition="test", shuffle=False, characteristics=True)
    assert (len(tr), len(va), len(te)) == (28 // 4, 8 // 4, 4 // 4)            # 70 / 20 / 10, remainder dropped
    spec_in, emb, spec_out = tr[0]
    assert spec_in.shape == (4, 144, 160, 2) and spec_in.dtype == np.float32
    assert emb.shape == (4, 2, 16) and emb.dtype == np.int32 and spec_out.shape == (4, 144, 160, 2)
    i0 = tr.index_in[0]
    assert np.array_equal(spec_in[0, :, :, 0], ds.amp[i0].astype(np.float32))
    assert np.array_equal(emb[0, 0], ds.emb[i0]) and np.array_equal(emb[0, 1], ds.emb[tr.index_out[0]])
    a, b, c = tr.__next__()                                                      # (spec_in, spec_out, emb) order
    assert np.array_equal(a, spec_in) and np.array_equal(b, spec_out) and np.array_equal(c, emb)
    assert len(te[0]) == 4 and te[0][3].shape == (4, 5, 2)


_DP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from oracle import unet_oracle as O
from unet_rir_b200.main_training import gradient_buckets, shard_batch, init_distributed
rank, world, _ = init_distributed(backend="gloo")
torch.set_num_threads(2)
om = O.UNetOracle(input_shape=(32, 32, 2), kernels=3, number_filters_0=8)
params = O.init_params(om.plan, seed=7)
g = torch.Generator().manual_seed(3)
B = 4
x, y = torch.rand(B, 32, 32, 2, generator=g), torch.rand(B, 32, 32, 2, generator=g)
emb = torch.randint(0, 2000, (B, 2, 16), generator=g)
names = O.trainable_names(om.plan)
def flat_grads(xs, ys, es, gb, replicas):
    st = O.new_opt_state(params, om.plan)
    _, grads, _ = O.train_step(om, params, st, xs, ys, es, 1e-3, loss_kind="dp", alpha=0.9, global_batch=gb,
                               num_replicas=replicas, apply=False)
    offs, off = {}, 0
    for n in names:
        offs[n] = (off, grads[n].numel()); off = (off + grads[n].numel() + 3) // 4 * 4
    flat = torch.zeros(off)
    for n in names:
        o, k = offs[n]; flat[o:o + k] = grads[n].flatten()
    return flat, offs, off
xs, ys, es = shard_batch([x, y, emb], rank, world)
flat, offs, n_flat = flat_grads(xs, ys, es, B, world)
for lo, hi in gradient_buckets({k: v for k, v in offs.items()}, n_flat):     # bucketed SUM all-reduce
    dist.all_reduce(flat[lo:hi], op=dist.ReduceOp.SUM)
if rank == 0:
    # single process: each shard through its own BatchNorm statistics, gradients summed
    ref = sum(flat_grads(*shard_batch([x, y, emb], r, world), B, world)[0] for r in range(world))
    err = float((flat - ref).norm() / ref.norm())
    full = flat_grads(x, y, emb, B, 1)[0]                                   # different BN statistics: must differ
    print("DP_ERR", err, float((flat - full).norm() / full.norm()))
# URIR_DP_GATHER_DENSE: a Dense kernel gradient summed over replicas == the product of the all-gathered operands
gr = torch.Generator().manual_seed(100 + rank)
xr, dyr = torch.randn(4, 24, generator=gr), torch.randn(4, 10, generator=gr)
summed = xr.t() @ dyr
dist.all_reduce(summed, op=dist.ReduceOp.SUM)
xa, dya = [torch.empty(4 * world, t.shape[1]) for t in (xr, dyr)]
dist.all_gather_into_tensor(xa, xr); dist.all_gather_into_tensor(dya, dyr)
if rank == 0:
    print("GATHER_ERR", float((xa.t() @ dya - summed).norm() / summed.norm()))
dist.destroy_process_group()
'''


def test_data_parallel_arithmetic_world2_gloo(tmp_path):
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29631", OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29631", str(script), ROOT],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("DP_ERR")][0].split()
    assert float(line[1]) < 1e-5            # bucketed all-reduce == sum of per-replica gradients
    assert float(line[2]) > 1e-3            # and is NOT the full-batch-BN gradient (BN is per replica)
    gather = [l for l in out.stdout.splitlines() if l.startswith("GATHER_ERR")][0].split()
    assert float(gather[1]) < 1e-6          # gathered-operand Dense gradient == all-reduced Dense gradient


def test_rt60_and_edc_of_known_decays():
    """rir_generation.rt60 / edc_db (torch, device-agnostic) against the oracle's numpy version and the analytic value
    for x[t] = N(0,1) * exp(-6.91 t / (rt60 * sr)) (60 dB of energy decay after rt60 seconds)."""
    import numpy as np
    from oracle import signal_oracle as SO
    from unet_rir_b200 import rir_generation as RG
    rng = np.random.default_rng(1)
    want = np.array([0.1, 0.2, 0.35])
    h = SO.synthetic_rir(3, rng, rt60_s=want, length=9600)
    got = RG.rt60(torch.as_tensor(h)).numpy()
    for i in range(3):
        assert abs(got[i] - SO.rt60(h[i])) < 1e-9 * max(1.0, got[i])
        assert abs(got[i] - want[i]) < 0.1 * want[i]
    e = RG.edc_db(torch.as_tensor(h)).numpy()
    assert np.allclose(e[:, 0], 0.0) and (np.diff(e, axis=1) <= 1e-12).all()
    assert np.abs(e[0] - SO.edc_db(h[0])).max() < 1e-9


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys,
    produced without a GPU and without touching /root/reference."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "unet_train_samples_per_sec" and line["unit"] == "samples/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"]


def test_multiply_shift_tile_division_is_exact_below_2_pow_20():
    """conv_thin.cu thin_tile_coords divides tile indices by tiles_w / tiles_h as (tile * ceil(2^40 / d)) >> 40 and
    thin_geometry_ok admits only launches with fewer than 2^20 tiles: the identity must hold for every such index."""
    n = np.arange(1 << 20, dtype=np.uint64)
    for d in (1, 2, 3, 5, 7, 9, 10, 18, 20, 36, 40, 72, 80, 144, 160, 1000, 4097, 1 << 19, (1 << 20) - 1):
        mag = np.uint64(((1 << 40) + d - 1) // d)
        assert np.array_equal((n * mag) >> np.uint64(40), n // np.uint64(d)), d


# ------------------------------------------------------------------------------------------------
# round 2: Keras weight import (f1), per-room reports (f3), checkpoint manager (a18)
# ------------------------------------------------------------------------------------------------
def _keras_topological_shuffle(pairs, rng):
    """A functional Keras model lists layers by depth, not creation: emulate with a random permutation of LAYERS
    (variables of one layer stay together), which the importer must undo through class + auto-name index."""
    groups, cur = [], None
    for plan_name, kname in pairs:
        layer = kname.split("/")[0]
        if layer != cur:
            groups.append([]); cur = layer
        groups[-1].append((plan_name, kname))
    order = rng.permutation(len(groups))
    return [pv for g in order for pv in groups[g]]


@pytest.mark.parametrize("mode,kernels", [(0, 3), (0, 6), (3, 3)])
def test_keras_npz_import_maps_every_variable(tmp_path, mode, kernels):
    """tools/export_tf_weights.py's file format -> plan: 77 trainable + 26 moving statistics (mode 0) land on the right
    plan entries whatever the order of the variables in the file and whatever the auto-name counters started at;
    HWIO / HWOI / (in, out) layouts are taken as they are (u_net.py:192-199, 269-304)."""
    from unet_rir_b200 import keras_weights as KW
    from unet_rir_b200 import plan as PL
    plan = PL.layer_plan(mode=mode, kernels=kernels)
    ref = PL.keras_init(plan, seed=7)
    rng = np.random.default_rng(1)
    for n in ref:                                   # make every tensor distinctive, including BN state
        ref[n] = ref[n] + torch.from_numpy(rng.standard_normal(tuple(ref[n].shape)).astype(np.float32)) * 0.01
    pairs = KW.keras_names_for_plan(plan, offsets={"conv2d": 24, "batch_normalization": 13, "conv2d_transpose": 4})
    if mode == 0:
        assert len(pairs) == 103 and sum(1 for n, _, k in plan if k in PL.TRAINABLE_KINDS) == 77
    assert ("vec.dense.w", "encoder_inf_dense/kernel:0") in pairs
    shuffled = _keras_topological_shuffle(pairs, rng)
    path = tmp_path / "w.npz"
    np.savez(path, names=np.array([k for _, k in shuffled]),
             **{f"arr_{i}": ref[p].numpy() for i, (p, _) in enumerate(shuffled)})
    got = KW.read_keras_npz(str(path), plan)
    assert list(got) == [n for n, _, _ in plan]
    for n in ref:
        assert got[n].dtype == torch.float32 and torch.equal(got[n], ref[n]), n
    # shapes as Keras stores them
    assert tuple(got["dec2.up.w"].shape) == (kernels, kernels, 256, 512)          # Conv2DTranspose: (kh, kw, out, in)
    assert tuple(got["enc2.down.w"].shape) == (kernels, kernels, 32, 64)          # Conv2D: HWIO
    assert tuple(got["vec.dense.w"].shape) == (2 * 16 * 256, 9 * 10 * 16)
    # a file from a differently configured model is refused, not silently mis-mapped
    other = PL.layer_plan(mode=mode, kernels=3 if kernels == 6 else 6)
    with pytest.raises(ValueError):
        KW.read_keras_npz(str(path), other)
    names = [k for _, k in shuffled][:-2]
    np.savez(tmp_path / "short.npz", names=np.array(names), **{f"arr_{i}": ref[p].numpy() for i, (p, _) in enumerate(shuffled[:-2])})
    with pytest.raises(ValueError):
        KW.read_keras_npz(str(tmp_path / "short.npz"), plan)


def test_per_room_reports_follow_the_reference_format(tmp_path):
    """rir_generation.py:303-532: six room groups in the reference's order, np.mean per group, positional(4) numbers for
    the three spectrogram metrics, scientific(4) for the waveform / misalignment ones, the three files' names / headers."""
    from unet_rir_b200 import rir_generation as R
    rooms = ["HemiAnechoicRoom", "LargeMeetingRoom", "LargeMeetingRoom", "ShoeBoxRoom", "SmallMeetingRoom", "LargeMeetingRoom"]
    res = {"total_mse": np.array([.5, .25, .75, .125, 1., .5]), "amp_mse": np.array([.1, .2, .3, .4, .5, .1]),
           "phase_loss": np.array([1., 1., 1., 1., 1., 1.]), "wav_mse": np.array([1e-3, 2e-3, 3e-3, 4e-3, 5e-3, 1e-3]),
           "wav_mse_50ms": np.array([1e-4] * 6), "missa_amp_db": np.array([-10., -20., -30., -5., -1., -10.]),
           "missa_wav_db": np.array([-1., -2., -3., -4., -5., -6.]), "t_model_inference_avg": 0.0123456, "t_postprocess": 2e-4}
    table = R.write_reports(res, rooms, "unet", str(tmp_path), 4, 1.5, t_loss=3e-5)
    assert table["Large"]["n"] == 3 and abs(table["Large"]["total_mse"] - 0.5) < 1e-12 and table["Medium"]["n"] == 0
    assert math.isnan(table["Medium"]["wav_mse"])
    losses = open(tmp_path / "unet_losses.csv").read().splitlines()
    assert losses[0] == ("room,n samples,MSE spectrogram,MSE magnitude,1-cos(y-y_) phase,MSE waveform,MSE waveform 50ms,"
                         "Misalignment magnitude,Misalignment waveform")
    assert [l.split(",")[0] for l in losses[1:]] == ["Global", "HemiAnechoic", "Large", "Medium", "Shoe", "Small"]
    assert losses[3] == "Large,3,0.5,0.2,1.,2.e-03,1.e-04,-2.e+01,-3.6667e+00"
    t = open(tmp_path / "unet_infer_time.csv").read().splitlines()
    assert t[0] == "n_samples,t_model_inference_avg,batch_size,t_postprocess,t_loss_calc,t_global"
    assert t[1] == "6,0.01235,4,0.0002,0.00003,1.5"
    txt = open(tmp_path / "unet_results_inference.txt").read()
    assert txt.startswith("unet results:\n\nTook 0.01235 s on average to infer spectrograms with batch size of 4\n")
    assert "LargeMeetingRoom losses (3 samples):\nTotal loss: 0.5 (MSE whole spectrogram)\t|\tAmplitude loss: 0.2 (MSE amplitude)" in txt
    assert "SmallMeetingRoom losses: (1 samples)\n" in txt and "Misalignment loss (amplitude): -1.e+00 (dB)\t|\t Misalignment loss (wav): -5.e+00 (dB)\n" in txt


def test_checkpoint_manager_keeps_the_last_two(tmp_path):
    """tf.train.CheckpointManager(max_to_keep=2) + save every second epoch (main_training.py:172, 363-364)."""
    from unet_rir_b200.main_training import CheckpointManager

    class FakeTrainer:
        def __init__(self): self.v = 0
        def checkpoint_state(self): return {"v": self.v}
        def load_checkpoint_state(self, st): self.v = st["v"]

    tr = FakeTrainer()
    m = CheckpointManager(tr, str(tmp_path), max_to_keep=2)
    assert m.latest_checkpoint is None
    for epoch in range(6):
        tr.v = epoch
        if epoch % 2 == 0:
            m.save()
    assert sorted(os.listdir(tmp_path)) == ["ckpt-2.pt", "ckpt-3.pt"]
    tr2 = FakeTrainer()
    m2 = CheckpointManager(tr2, str(tmp_path), max_to_keep=2)
    assert m2.restore_latest().endswith("ckpt-3.pt") and tr2.v == 4
    m2.save()
    assert sorted(os.listdir(tmp_path)) == ["ckpt-3.pt", "ckpt-4.pt"]
