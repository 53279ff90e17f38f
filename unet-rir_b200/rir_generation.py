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


def batch_metrics(spec_true, spec_pred, wav_true, wav_pred, sr=48000, spec_generated=None):
    """The per-sample numbers of rir_generation.py:195-225 for a whole batch (dict of (B,) tensors),
    plus rt60_true / rt60_pred / edc_mae_db. spec_pred is the feature that was post-processed (with diff_gen its
    phase channel is output + input phase, :173-176); spec_generated is the RAW model output, which is what the
    reference's total loss compares with the target (:197) -- it defaults to spec_pred."""
    st, sp = spec_true.double(), spec_pred.double()
    sg = sp if spec_generated is None else spec_generated.double()
    a_t, p_t, a_p, p_p = st[..., 0], st[..., 1], sp[..., 0], sp[..., 1]
    B = st.shape[0]
    out = {}
    out["amp_mse"] = ((a_t - a_p) ** 2).reshape(B, -1).mean(1)                       # :195
    out["phase_loss"] = (1 - torch.cos((p_t - p_p) * 2 * math.pi)).reshape(B, -1).mean(1)   # :196
    out["total_mse"] = ((st - sg) ** 2).reshape(B, -1).mean(1)                       # :197
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


# ---- per-room reports (rir_generation.py:227-532) ------------------------------------------------------------
ROOM_GROUPS = (("Global", None), ("HemiAnechoic", "HemiAnechoicRoom"), ("Large", "LargeMeetingRoom"),
               ("Medium", "MediumMeetingRoom"), ("Shoe", "ShoeBoxRoom"), ("Small", "SmallMeetingRoom"))
# (results key, CSV column, positional?) in the reference's column order (:369-418)
REPORT_METRICS = (("total_mse", "MSE spectrogram", True), ("amp_mse", "MSE magnitude", True),
                  ("phase_loss", "1-cos(y-y_) phase", True), ("wav_mse", "MSE waveform", False),
                  ("wav_mse_50ms", "MSE waveform 50ms", False), ("missa_amp_db", "Misalignment magnitude", False),
                  ("missa_wav_db", "Misalignment waveform", False))


def _fmt(v, positional, precision=4):
    v = float(v)
    return np.format_float_positional(v, precision=precision) if positional else np.format_float_scientific(v, precision=precision)


def room_table(results, rooms):
    """Means of the seven metrics over all samples and per room group, as the reference accumulates them in its
    per-room lists (:227-290, 303-356): -> {group: {"n": count, metric: mean}}; an empty group's means are nan
    (np.mean of an empty list, as there)."""
    rooms = np.asarray(list(rooms))
    table = {}
    for label, room in ROOM_GROUPS:
        sel = np.ones(len(rooms), dtype=bool) if room is None else rooms == room
        row = {"n": int(sel.sum())}
        for key, _, _ in REPORT_METRICS:
            vals = np.asarray(results[key])[sel]
            row[key] = float(np.mean(vals)) if len(vals) else float("nan")
        table[label] = row
    return table


def write_reports(results, rooms, name, out_dir, batch_size, t_global, t_loss=float("nan")):
    """<name>_infer_time.csv, <name>_losses.csv and <name>_results_inference.txt with the reference's columns, row
    order, labels and number formats (:358-532). `rooms` = room name of every sample's TARGET (characteristic[j,:,1][0])."""
    import csv
    import os
    os.makedirs(out_dir, exist_ok=True)
    n = len(rooms)
    table = room_table(results, rooms)
    t_inf, t_post = results.get("t_model_inference_avg", float("nan")), results.get("t_postprocess", float("nan"))
    pos5 = lambda v: np.format_float_positional(float(v), precision=5)
    with open(os.path.join(out_dir, f"{name}_infer_time.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["n_samples", "t_model_inference_avg", "batch_size", "t_postprocess", "t_loss_calc", "t_global"])
        w.writerow([n, pos5(t_inf), batch_size, pos5(t_post), pos5(t_loss), pos5(t_global)])
    with open(os.path.join(out_dir, f"{name}_losses.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["room", "n samples"] + [col for _, col, _ in REPORT_METRICS])
        for label, _ in ROOM_GROUPS:
            row = table[label]
            w.writerow([label, row["n"]] + [_fmt(row[key], positional) for key, _, positional in REPORT_METRICS])
    with open(os.path.join(out_dir, f"{name}_results_inference.txt"), "w") as f:
        f.write(f"{name} results:\n\n")
        f.write(f"Took {pos5(t_inf)} s on average to infer spectrograms with batch size of {batch_size}\n")
        f.write(f"Took {pos5(t_post)} s on average to postprocess and generate each spectrogram and waveform\n")
        f.write(f"Took {pos5(t_loss)} s on average to obtain the losses for each waveform\n")
        f.write(f"Took {pos5(t_global)} s to generate, postprocess and obtain loss for {n} samples\n\n")
        for label, room in ROOM_GROUPS:
            r = table[label]
            if room is None:
                f.write("Total losses:\n")
            elif label == "Small":
                f.write(f"{room} losses: ({r['n']} samples)\n")       # the reference's punctuation differs for this one (:521)
            else:
                f.write(f"{room} losses ({r['n']} samples):\n")
            f.write(f"Total loss: {_fmt(r['total_mse'], True)} (MSE whole spectrogram)\t|\tAmplitude loss: "
                    f"{_fmt(r['amp_mse'], True)} (MSE amplitude)\t|\tPhase loss: {_fmt(r['phase_loss'], True)} "
                    f"(1-cos(y_true - y_pred))\n")
            f.write(f"Waveform loss: {_fmt(r['wav_mse'], False)} (MSE)\t|\t 50 ms waveform loss: "
                    f"{_fmt(r['wav_mse_50ms'], False)} (MSE)" + (" \n" if label == "Small" else "\n"))
            f.write(f"Misalignment loss (amplitude): {_fmt(r['missa_amp_db'], False)} (dB)\t|\t Misalignment loss (wav): "
                    f"{_fmt(r['missa_wav_db'], False)} (dB)\n")
            if label != "Small":
                f.write("\n")
    return table


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
             verbose=True, report_dir=None, report_name="unet"):
    """The reference's generation + loss loop over a test generator; returns dict of per-sample arrays and timings."""
    results, t_inf, t_post, rooms = {}, [], [], []
    n = len(generator) if max_batches is None else min(len(generator), max_batches)
    chars = dataset.return_characteristics() if hasattr(dataset, "return_characteristics") else None
    t_begin = time.time()
    for i in range(n):
        spec_in, emb, spec_out = generator[i][:3]
        lo = i * generator.batch_size
        idx_out = generator.index_out[lo:lo + generator.batch_size]
        wav_true = torch.as_tensor(np.stack([dataset.waveform(j) for j in idx_out])).cuda()
        rooms.extend((chars[j][0] if chars is not None else "") for j in idx_out)     # characteristic[j, :, 1][0] (:209)
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
        m = batch_metrics(torch.as_tensor(spec_out).cuda(), feat, wav_true, wav_pred, spec_generated=spec_generated)
        for k, v in m.items():
            results.setdefault(k, []).append(v.cpu().numpy())
    results = {k: np.concatenate(v) for k, v in results.items()}
    # the reference drops the first (warm-up) timing of each list (:358-360)
    results["t_model_inference_avg"] = float(np.mean(t_inf[1:] or t_inf)) if t_inf else float("nan")
    results["t_postprocess"] = float(np.mean(t_post[1:] or t_post)) if t_post else float("nan")
    results["t_global"] = time.time() - t_begin
    results["rooms"] = rooms
    if report_dir is not None:
        write_reports(results, rooms, report_name, report_dir, generator.batch_size, results["t_global"])
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
        generate(trained_model, test_generator, dataset, diff_gen=diff_gen,
                 report_dir=f'../generated_rir_distributed/{name}_{algorithm}', report_name=name)
