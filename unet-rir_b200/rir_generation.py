"""Batched RIR generation: spectrogram -> U-Net (training=False) -> inverse STFT, with the reference's
per-sample metrics (rir_generation.py:43-536, hot loops at :160-225).

The reference runs the model per batch of 4 and then a per-sample Python/librosa loop (:170-181); here
the whole batch goes through one engine forward and ONE iSTFT kernel launch, and the seven metrics of
:195-225 (plus the Schroeder energy-decay / RT60 agreement the north star asks for) are computed for the
batch on the GPU. Module-level `amplitude_loss` / `phase_loss` keep the reference's names (:31-40).

`python -m unet_rir_b200.rir_generation` runs the reference's `__main__` flow on the synthetic dataset
(the reference's dataset and checkpoints do not exist here; its own `__main__` does not even parse,
rir_generation.py:63 is over-indented).
"""
from __future__ import annotations

import math
import time

import numpy as np
import torch

from .datageneratorv2 import DataGenerator
from .dataset import Dataset
from .dl_models.u_net import UNet
from .postprocess import PostProcess, post_process_batch


def amplitude_loss(y_true, y_pred):
    """tf.keras.losses.mean_squared_error: mean over the LAST axis (:31-34)."""
    yt, yp = torch.as_tensor(y_true, dtype=torch.float32), torch.as_tensor(y_pred, dtype=torch.float32)
    return ((yt - yp.to(yt.device)) ** 2).mean(dim=-1)


def phase_loss(y_true, y_pred):
    """mean over the last axis of 1 - cos of the de-normalised phase difference (:36-40)."""
    yt, yp = torch.as_tensor(y_true, dtype=torch.float32), torch.as_tensor(y_pred, dtype=torch.float32)
    d = (yt * 2 * math.pi - math.pi) - (yp.to(yt.device) * 2 * math.pi - math.pi)
    return (1 - torch.cos(d)).mean(dim=-1)


def edc_db(h):
    """Schroeder backward-integrated energy decay curve, dB, batched over the leading dims."""
    e = torch.flip(torch.cumsum(torch.flip(h.double() ** 2, dims=[-1]), dim=-1), dims=[-1])
    e = e / e[..., :1].clamp_min(1e-300)
    return 10 * torch.log10(e.clamp_min(1e-30))


def rt60(h, sr=48000, lo=-5.0, hi=-25.0):
    """RT60 (T20 fit of the EDC between lo and hi dB, extrapolated to 60 dB), one value per row."""
    d = edc_db(h)
    t = torch.arange(d.shape[-1], device=d.device, dtype=torch.float64) / sr
    m = ((d <= lo) & (d >= hi)).double()
    n = m.sum(-1).clamp_min(2)
    tm, dm = (t * m).sum(-1) / n, (d * m).sum(-1) / n
    cov = ((t - tm[..., None]) * (d - dm[..., None]) * m).sum(-1)
    var = (((t - tm[..., None]) ** 2) * m).sum(-1).clamp_min(1e-30)
    slope = cov / var
    return torch.where(slope < 0, -60.0 / slope, torch.full_like(slope, float("nan")))


def batch_metrics(spec_true, spec_pred, wav_true, wav_pred, sr=48000):
    """The per-sample numbers of rir_generation.py:195-225 for a whole batch (dict of (B,) tensors),
    plus rt60_true / rt60_pred / edc_mae_db."""
    st, sp = spec_true.double(), spec_pred.double()
    a_t, p_t, a_p, p_p = st[..., 0], st[..., 1], sp[..., 0], sp[..., 1]
    B = st.shape[0]
    out = {}
    out["amp_mse"] = ((a_t - a_p) ** 2).reshape(B, -1).mean(1)                       # :195
    out["phase_loss"] = (1 - torch.cos((p_t - p_p) * 2 * math.pi)).reshape(B, -1).mean(1)   # :196
    out["total_mse"] = ((st - sp) ** 2).reshape(B, -1).mean(1)                       # :197
    num = (a_p - a_t).reshape(B, -1).norm(dim=1)
    den = a_t.reshape(B, -1).norm(dim=1)
    out["missa_amp_db"] = 20 * torch.log10(num / den)                                # :203-205
    wt, wp = wav_true.double(), wav_pred.double()
    out["wav_mse"] = ((wt - wp) ** 2).mean(1)                                        # :215
    n50 = int(0.05 * sr)
    out["wav_mse_50ms"] = ((wt[:, :n50] - wp[:, :n50]) ** 2).mean(1)                 # :218
    out["missa_wav_db"] = 20 * torch.log10((wp - wt).norm(dim=1) / wt.norm(dim=1))   # :221-223
    out["rt60_true"], out["rt60_pred"] = rt60(wt, sr), rt60(wp, sr)
    et, ep = edc_db(wt), edc_db(wp)
    valid = (et > -40.0).double()
    out["edc_mae_db"] = ((et - ep).abs() * valid).sum(1) / valid.sum(1).clamp_min(1)
    return out


def generate_batch(model: UNet, spec_in, emb, diff_gen=False):
    """One hot-loop iteration of :160-181 for a whole batch: returns (spec_generated, wav_pred) on the GPU."""
    with torch.no_grad():
        spec_generated = model.model([spec_in, emb], training=False)
        feat = spec_generated
        if diff_gen:                                                                 # :173-176
            spec_in_t = torch.as_tensor(spec_in, dtype=torch.float32).to(feat.device)
            feat = torch.stack([spec_generated[..., 0], spec_generated[..., 1] + spec_in_t[..., 1]], dim=-1)
        wav_pred = post_process_batch(feat)
    return feat, wav_pred


def generate(model: UNet, generator: DataGenerator, dataset: Dataset, diff_gen=False, max_batches=None,
             verbose=True):
    """The reference's generation + loss loop over a test generator; returns dict of per-sample arrays and timings."""
    results, t_inf, t_post = {}, [], []
    n = len(generator) if max_batches is None else min(len(generator), max_batches)
    for i in range(n):
        spec_in, emb, spec_out = generator[i][:3]
        lo = i * generator.batch_size
        idx_out = generator.index_out[lo:lo + generator.batch_size]
        wav_true = torch.as_tensor(np.stack([dataset.waveform(j) for j in idx_out])).cuda()
        torch.cuda.synchronize(); t0 = time.time()
        with torch.no_grad():
            spec_generated = model.model([spec_in, emb], training=False)
        torch.cuda.synchronize(); t1 = time.time()
        feat = spec_generated
        if diff_gen:
            feat = torch.stack([spec_generated[..., 0],
                                spec_generated[..., 1] + torch.as_tensor(spec_in[..., 1]).cuda()], dim=-1)
        wav_pred = post_process_batch(feat)
        torch.cuda.synchronize(); t2 = time.time()
        t_inf.append(t1 - t0); t_post.append((t2 - t1) / len(idx_out))
        m = batch_metrics(torch.as_tensor(spec_out).cuda(), feat, wav_true, wav_pred)
        for k, v in m.items():
            results.setdefault(k, []).append(v.cpu().numpy())
    results = {k: np.concatenate(v) for k, v in results.items()}
    results["t_model_inference_avg"] = float(np.mean(t_inf)) if t_inf else float("nan")
    results["t_postprocess"] = float(np.mean(t_post)) if t_post else float("nan")
    if verbose:
        for k in ("total_mse", "amp_mse", "phase_loss", "wav_mse", "wav_mse_50ms", "missa_amp_db", "missa_wav_db"):
            print(f"{k:>14s}: {np.nanmean(results[k]):.6g}")
        print(f"inference s/batch: {results['t_model_inference_avg']:.6f}   postprocess s/sample: {results['t_postprocess']:.6f}")
    return results


if __name__ == '__main__':

    batch_size = 4
    debug = False
    target_size = (144, 160, 2)
    models = ['unet_diff_full']
    algorithm = 'ph'
    diff_gen = True

    dataset = Dataset(None, 'room_impulse', normalization=True, debugging=debug, extract=False,
                      room_characteristics=True, room=['All'], array=None, n_synthetic=160)
    test_generator = DataGenerator(dataset, batch_size=batch_size, partition='test', shuffle=False,
                                   characteristics=True)
    for name in models:
        print("Generating with UNET")
        trained_model = UNet(input_shape=target_size,
                             inf_vector_shape=(2, 16),
                             mode=0,
                             number_filters_0=32,
                             kernels=3,
                             name=name
                             )
        print("Initializing from scratch.")
        postprocessor = PostProcess(name, algorithm=algorithm)
        print('Generating wavs and obtaining loss')
        generate(trained_model, test_generator, dataset, diff_gen=diff_gen)
