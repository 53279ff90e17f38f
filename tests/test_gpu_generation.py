"""GPU parity of the generation path (BASELINE config 4, rir_generation.py:160-225): spectrogram -> U-Net
(training=False) -> un-pad / denormalise / inverse STFT, and the per-sample metrics, against the CPU oracle.

Tolerances: generated spectrogram abs <= 1e-2 (bf16 net vs fp32 oracle); the inverse-STFT stage alone (same
feature into both) waveform misalignment <= -60 dB; end to end (each side's own feature) the waveforms agree to
<= -18 dB -- the normalised log-amplitude spans 100 dB, so a 1e-2 spectrogram error is ~1 dB of amplitude --
while the room-acoustic quantities the north star names agree tightly: RT60 within 3 %, energy-decay curve
within 0.5 dB (mean absolute, down to -40 dB). batch_metrics equals the oracle's per-sample metrics to 1e-6."""
import numpy as np
import pytest
import torch

from oracle import signal_oracle as SO
from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import rir_generation as RG
from unet_rir_b200.dl_models.u_net import UNet
from unet_rir_b200.postprocess import PostProcess, griffinlim_batch, post_process_batch
from unet_rir_b200.preprocess import preprocess_batch

pytestmark = pytest.mark.gpu


def _missa_db(a, b):
    return 20 * np.log10(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_generate_batch_matches_oracle_pipeline():
    rng = np.random.default_rng(7)
    B = 4                                                   # rir_generation.py:45
    wav_src = SO.synthetic_rir(B, rng, rt60_s=[0.25, 0.4, 0.6, 0.9])
    wav_tgt = SO.synthetic_rir(B, rng, rt60_s=[0.3, 0.5, 0.7, 1.0])
    spec_in = preprocess_batch(wav_src)                     # GPU STFT features, (B,144,160,2)
    spec_tgt = preprocess_batch(wav_tgt)
    g = torch.Generator().manual_seed(0)
    emb = torch.randint(0, 2000, (B, 2, 16), generator=g, dtype=torch.int32)
    om = O.UNetOracle(kernels=3)
    params = O.init_params(om.plan, seed=500)
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    unet.model.engine.load_state_dict(params)

    feat, wav_pred = RG.generate_batch(unet, spec_in, emb)
    feat_c, wav_c = feat.float().cpu().numpy(), wav_pred.cpu().numpy()
    assert feat_c.shape == (B, 144, 160, 2) and wav_c.shape == (B, 9600)
    ref_feat = om.forward(params, spec_in.cpu(), emb, training=False).numpy()
    assert np.abs(feat_c - ref_feat).max() < 1e-2

    # inverse-STFT stage alone: the GPU feature through the oracle's post-processing
    for i in range(B):
        assert _missa_db(wav_c[i], SO.post_process(feat_c[i])) < -60.0
    # end to end, each side with its own feature
    ref_wav = np.stack([SO.post_process(ref_feat[i]) for i in range(B)])
    for i in range(B):
        assert _missa_db(wav_c[i], ref_wav[i]) < -18.0, _missa_db(wav_c[i], ref_wav[i])
        r_gpu, r_ref = SO.rt60(wav_c[i]), SO.rt60(ref_wav[i])
        if np.isfinite(r_ref):
            assert abs(r_gpu - r_ref) < 0.03 * r_ref, (r_gpu, r_ref)
        e_gpu, e_ref = SO.edc_db(wav_c[i]), SO.edc_db(ref_wav[i])
        sel = e_ref > -40.0
        assert np.abs(e_gpu - e_ref)[sel].mean() < 0.5

    # the reference's per-sample metrics for the batch, on the GPU vs the oracle's numpy version
    wav_true = torch.as_tensor(wav_tgt - wav_tgt.mean(axis=1, keepdims=True)).cuda()
    m = RG.batch_metrics(spec_tgt, feat, wav_true, wav_pred)
    for i in range(B):
        want = SO.generation_metrics(spec_tgt[i].cpu().numpy(), feat_c[i].astype(np.float64), wav_true[i].cpu().numpy(), wav_c[i])
        for k, v in want.items():
            assert abs(float(m[k][i]) - v) < 1e-6 * max(1.0, abs(v)), (k, float(m[k][i]), v)
        assert abs(float(m["rt60_pred"][i]) - SO.rt60(wav_c[i])) < 1e-6 or not np.isfinite(SO.rt60(wav_c[i]))


def test_round_trip_features_to_waveform():
    """preprocess_batch -> post_process_batch reproduces a mean-removed RIR (the reference's own sanity print,
    preprocess.py:201-205): misalignment <= -60 dB on the interior samples."""
    rng = np.random.default_rng(3)
    wav = SO.synthetic_rir(3, rng)
    back = post_process_batch(preprocess_batch(wav)).cpu().numpy()
    ref = wav - wav.mean(axis=1, keepdims=True)
    for i in range(3):
        assert _missa_db(back[i][128:-128], ref[i][128:-128]) < -60.0


def test_griffin_lim_matches_oracle_and_recovers_the_magnitudes():
    """PostProcess(algorithm='gl') (postprocess.py:130-131, librosa.griffinlim defaults) as a batched GPU loop: with the
    same initial phases the waveform follows the numpy oracle (<= -40 dB after 32 iterations of fp32 vs fp64), and the
    result's STFT magnitude is consistent with the input (spectral convergence below 0.35 for decaying-noise RIRs)."""
    rng = np.random.default_rng(11)
    wav = SO.synthetic_rir(2, rng, rt60_s=[0.3, 0.6])
    S = np.stack([np.abs(SO.stft(w - w.mean())) for w in wav])                     # (2, 129, 151)
    init = np.exp(2j * np.pi * rng.random(S.shape))
    got = griffinlim_batch(S.astype(np.float32), init_angles=init).cpu().numpy()
    for i in range(2):
        ref = SO.griffinlim(S[i], init_angles=init[i])
        assert got[i].shape == ref.shape == (9600,)
        assert _missa_db(got[i], ref) < -40.0, _missa_db(got[i], ref)
        sc = np.linalg.norm(np.abs(SO.stft(got[i])) - S[i]) / np.linalg.norm(S[i])
        assert sc < 0.35, sc
    # the class interface: normalised padded feature in, waveform out (random initial phases)
    feat = preprocess_batch(wav)[0].cpu().numpy()
    w = PostProcess("t", algorithm="gl").post_process(feat, [1, 2])
    assert w.shape == (9600,) and np.isfinite(w).all()
    sc = np.linalg.norm(np.abs(SO.stft(w)) - S[0]) / np.linalg.norm(S[0])
    assert sc < 0.4, sc


def test_dataset_directory_walk(tmp_path):
    """Dataset(dir) (dataset.py:121-182) on a small tree of synthetic wav files: filename parsing, room / array filters,
    room-geometry embeddings, per-room in/out pairing with the seed-500 shuffle, GPU-preprocessed spectrograms equal
    to the oracle's per-file pipeline, and the DataGenerator batch contract on top."""
    from scipy.io import wavfile
    from unet_rir_b200 import rooms as R
    from unet_rir_b200.datageneratorv2 import DataGenerator
    from unet_rir_b200.dataset import Dataset
    rng = np.random.default_rng(5)
    root = tmp_path / "room_impulse"
    files = {}
    for room in ("HemiAnechoicRoom", "SmallMeetingRoom", "AnechoicRoom"):
        for zone in ("ZoneA", "ZoneC"):
            for arr in ("PlanarMicrophoneArray", "CircularMicrophoneArray"):
                d = root / room / zone / arr
                d.mkdir(parents=True)
                for l, m in ((3, 5), (7, 12), (22, 64)):
                    w = SO.synthetic_rir(1, rng, length=12000)[0] * 0.1
                    name = f"{room}_{zone}_{arr}_L{l}_M{m}.wav"
                    wavfile.write(str(d / name), 48000, w)
                    files[(room, zone[-1], arr.replace("MicrophoneArray", ""), str(l), str(m))] = w
    ds = Dataset(str(tmp_path), "room_impulse", room=["All"], array=["PlanarMicrophoneArray"], room_characteristics=True)
    assert len(ds) == 2 * 2 * 3                                   # anechoic room and circular arrays filtered out
    assert ds.amp.shape == (12, 144, 160) and ds.emb.shape == (12, 16) and ds.emb.dtype == np.int32
    chars = ds.return_characteristics()
    for i in range(len(ds)):
        ch = chars[i]
        assert ch[2] == "Planar" and ch[0] in ("HemiAnechoicRoom", "SmallMeetingRoom")
        assert list(ds.emb[i]) == [int(v) for v in R.uts_room(ch[0]).return_embedding(ch)]
        w = files[tuple(ch)][:9600]
        a, p = SO.preprocess(w)[..., 0], SO.preprocess(w)[..., 1]
        assert np.abs(ds.amp[i] - a).max() < 2e-4
        # pairs stay inside one room; index_out is a permutation of index_in
        assert chars[ds.index_in[i]][0] == chars[ds.index_out[i]][0]
    assert sorted(ds.index_in) == sorted(ds.index_out) == list(range(12))
    assert ds.index_in[:6] == [i for i in range(12) if chars[i][0] == "HemiAnechoicRoom"]
    gen = DataGenerator(ds, batch_size=4, partition="all", shuffle=False)
    spec_in, emb, spec_out = gen[0]
    assert spec_in.shape == (4, 144, 160, 2) and emb.shape == (4, 2, 16) and spec_out.dtype == np.float32
    dbg = Dataset(str(tmp_path), "room_impulse", room=["All"], array=["PlanarMicrophoneArray"], debugging=True)
    assert len(dbg) == 3


def test_eval_forward_graph_replay_equals_eager():
    """The inference forward is captured per batch size on its second call and replayed afterwards: the eager call, the
    capture call and replays on NEW inputs and after a weight change all equal an engine with the graph disabled."""
    from unet_rir_b200.engine import UNetEngine
    g = torch.Generator().manual_seed(2)
    a, b = UNetEngine(kernels=3), UNetEngine(kernels=3)
    b.eval_cuda_graph = False
    b.load_state_dict(a.state_dict())
    for step in range(4):
        x = torch.rand(4, 144, 160, 2, generator=g).cuda()
        emb = torch.randint(0, 2000, (4, 2, 16), generator=g, dtype=torch.int32).cuda()
        if step == 3:                                   # weights change between replays
            sd = a.state_dict()
            sd["head.b"] = sd["head.b"] + 0.5
            a.load_state_dict(sd); b.load_state_dict(sd)
        ya = a.forward(x, emb, training=False).clone()
        yb = b.forward(x, emb, training=False).clone()
        assert torch.equal(ya, yb), step
    assert not isinstance(a._eval_graphs[4], str)
