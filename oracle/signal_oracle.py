"""CPU oracle for the amp/phase signal path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: librosa is not installed in the build container and the reference pins no
version, so `stft`/`istft` restate librosa's published algorithm (librosa.core.spectrum.stft /
istft, 0.9-0.10 era) in numpy; the reference's own call sites anchor the parameters:
  * preprocess.py:13-18   FeatureExtractor.extract  -> extract
  * preprocess.py:21-41   Normalizer                -> normalize / denormalize
  * preprocess.py:60-113  TensorPadder              -> pad / un_pad
  * preprocess.py:51-57   Loader.load (mean removal)-> remove_mean
  * postprocess.py:54-133 PostProcess.post_process  -> post_process (algorithm 'ph')
  * rir_generation.py:195-225 per-sample metrics    -> generation_metrics
  * (new, BASELINE north_star) Schroeder EDC / RT60 -> edc_db / rt60

Structural known answers held by the reference: 9600 samples -> (129, 151) bins x frames
(postprocess.py:54, dataset.py:62-70), padded to (144, 160); iSTFT returns 9600 samples.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.
"""
from __future__ import annotations

import math

import numpy as np

N_FFT, WIN_LENGTH, HOP_LENGTH = 256, 128, 64          # dataset.py:62-64
SR, DURATION = 48000, 0.2                              # dataset.py:66-67
INPUT_SHAPE = (144, 160)                               # dataset.py:70
MD = 100.0                                             # preprocess.py:23
EP = 10 ** (-1 * MD / 20)                              # preprocess.py:24


def hann_periodic(n):
    """scipy.signal.get_window('hann', n, fftbins=True), which librosa uses."""
    return 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)


def padded_window(n_fft=N_FFT, win_length=WIN_LENGTH):
    """librosa.util.pad_center(window, size=n_fft): the window is centred inside the FFT frame."""
    w = hann_periodic(win_length)
    lp = (n_fft - win_length) // 2
    out = np.zeros(n_fft)
    out[lp:lp + win_length] = w
    return out


def stft(y, n_fft=N_FFT, win_length=WIN_LENGTH, hop_length=HOP_LENGTH, pad_mode="constant"):
    """librosa.stft(y, n_fft, hop_length, win_length) with center=True, window='hann'.

    pad_mode: 'constant' (librosa >= 0.10 default) or 'reflect' (< 0.10) -- unpinned by the
    reference, taken as a parameter. Returns complex64 (1+n_fft/2, n_frames).
    """
    y = np.asarray(y)
    w = padded_window(n_fft, win_length)
    yp = np.pad(y.astype(np.float64), n_fft // 2, mode=pad_mode)
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    frames = yp[idx] * w[None, :]
    return np.fft.rfft(frames, axis=1).T.astype(np.complex64)


def istft(S, n_fft=N_FFT, win_length=WIN_LENGTH, hop_length=HOP_LENGTH, length=None):
    """librosa.istft(S, n_fft, hop_length, win_length), center=True, window='hann':
    irfft -> * window -> overlap-add -> / window-sum-square where > tiny -> trim n_fft//2."""
    S = np.asarray(S)
    n_frames = S.shape[1]
    w = padded_window(n_fft, win_length)
    expected = n_fft + hop_length * (n_frames - 1)
    y = np.zeros(expected, dtype=np.float64)
    ytmp = np.fft.irfft(S.T.astype(np.complex128), n=n_fft, axis=1) * w[None, :]
    wss = np.zeros(expected, dtype=np.float64)
    w2 = w * w
    for t in range(n_frames):
        y[t * hop_length:t * hop_length + n_fft] += ytmp[t]
        wss[t * hop_length:t * hop_length + n_fft] += w2
    tiny = np.finfo(np.float32).tiny
    nz = wss > tiny
    y[nz] /= wss[nz]
    y = y[n_fft // 2:]
    if length is None:
        y = y[:expected - n_fft]           # librosa: y[n_fft//2 : -n_fft//2]
    else:
        y = y[:length]
    return y.astype(np.float32)


def remove_mean(signal):
    """Loader.load's `signal -= np.mean(signal)` (preprocess.py:56)."""
    s = np.asarray(signal, dtype=np.float32)
    return s - np.mean(s)


def extract(waveform, pad_mode="constant"):
    """FeatureExtractor.extract (preprocess.py:13-18)."""
    S = stft(waveform, pad_mode=pad_mode)
    return np.abs(S), np.angle(S)


def normalize(amp, phase):
    """Normalizer.normalize (preprocess.py:26-32)."""
    amp_norm = 20 * np.log10(amp / 128 + EP)
    amp_norm = (amp_norm + MD) / MD
    phase_norm = (phase + math.pi) / (2 * math.pi)
    return amp_norm, phase_norm


def denormalize(amp_norm, phase_norm):
    """Normalizer.denormalize (preprocess.py:34-41)."""
    amp = (amp_norm * MD) - MD
    amp = (10 ** (amp / 20) - EP) * 128
    phase = (phase_norm * 2 * math.pi) - math.pi
    phase = (phase + math.pi) % (2 * math.pi) - math.pi
    return amp, phase


def pad(t, desired=INPUT_SHAPE):
    """TensorPadder.transform (preprocess.py:75-105): zero rows below, zero columns right;
    returns the input unchanged if it is larger than `desired` in either dim."""
    t = np.asarray(t)
    if t.shape[0] > desired[0] or t.shape[1] > desired[1]:
        return t
    out = np.zeros(desired, dtype=np.float64)
    out[:t.shape[0], :t.shape[1]] = t
    return out


def un_pad(amp, phase, desired_shape):
    """TensorPadder.un_pad (preprocess.py:107-113)."""
    return (np.asarray(amp)[:desired_shape[0], :desired_shape[1]],
            np.asarray(phase)[:desired_shape[0], :desired_shape[1]])


def preprocess(wav, pad_mode="constant"):
    """Dataset.preprocess (dataset.py:214-223) minus file loading: (T,) -> (144,160,2) f32."""
    amp, ph = extract(remove_mean(wav), pad_mode)
    an, pn = normalize(amp, ph)
    return np.stack([pad(an), pad(pn)], axis=-1).astype(np.float32)


def post_process(feature, des_shape=(129, 151), n_fft=N_FFT, win_length=WIN_LENGTH,
                 hop_length=HOP_LENGTH):
    """PostProcess.post_process with algorithm 'ph' (postprocess.py:54-133), no file output."""
    stft_n, phase_n = feature[:, :, 0], feature[:, :, 1]
    a, p = un_pad(stft_n, phase_n, des_shape)
    a, p = denormalize(a.astype(np.float64), p.astype(np.float64))
    S = a * (np.cos(p) + 1j * np.sin(p))
    return istft(S, n_fft, win_length, hop_length)


def griffinlim(S, n_iter=32, momentum=0.99, init_angles=None, rng=None, n_fft=N_FFT, win_length=WIN_LENGTH,
               hop_length=HOP_LENGTH, pad_mode="constant"):
    """librosa.griffinlim with its defaults (postprocess.py:130-131 passes only n_fft / win_length / hop_length):
    fast Griffin-Lim, 32 iterations, momentum 0.99, random initial phases. S = magnitudes (n_bins, n_frames).
    Third-party algorithm (librosa, unpinned version): restated from its published form --
        angles_0 = exp(2 pi i U);  rebuilt_k = STFT(ISTFT(S * angles_k));
        angles_{k+1} = normalise(rebuilt_k - momentum / (1 + momentum) * rebuilt_{k-1});  y = ISTFT(S * angles_n)."""
    S = np.asarray(S, dtype=np.float64)
    if init_angles is None:
        rng = rng or np.random.default_rng()
        init_angles = np.exp(2j * np.pi * rng.random(S.shape))
    angles = np.asarray(init_angles, dtype=np.complex128)
    rebuilt = np.zeros_like(angles)
    length = hop_length * (S.shape[1] - 1)
    for _ in range(n_iter):
        tprev = rebuilt
        inverse = istft(S * angles, n_fft, win_length, hop_length, length=length)
        rebuilt = stft(inverse, n_fft, win_length, hop_length, pad_mode=pad_mode)
        angles = rebuilt - (momentum / (1 + momentum)) * tprev
        angles = angles / (np.abs(angles) + 1e-16)
    return istft(S * angles, n_fft, win_length, hop_length, length=length)


# -- metrics ---------------------------------------------------------------------------
def generation_metrics(spec_true, spec_pred, wav_true, wav_pred):
    """The seven per-sample numbers of rir_generation.py:195-225."""
    a_t, p_t = spec_true[..., 0].astype(np.float64), spec_true[..., 1].astype(np.float64)
    a_p, p_p = spec_pred[..., 0].astype(np.float64), spec_pred[..., 1].astype(np.float64)
    out = {}
    out["amp_mse"] = float(np.mean((a_t - a_p) ** 2))
    out["phase_loss"] = float(np.mean(1 - np.cos((p_t - p_p) * 2 * math.pi)))
    out["total_mse"] = float(np.mean((spec_true.astype(np.float64) - spec_pred) ** 2))
    out["missa_amp_db"] = 20 * math.log10(np.linalg.norm((a_p - a_t).ravel()) /
                                          np.linalg.norm(a_t.ravel()))
    wt, wp = wav_true.astype(np.float64), wav_pred.astype(np.float64)
    out["wav_mse"] = float(np.mean((wt - wp) ** 2))
    out["wav_mse_50ms"] = float(np.mean((wt[:2400] - wp[:2400]) ** 2))
    out["missa_wav_db"] = 20 * math.log10(np.linalg.norm(wp - wt) / np.linalg.norm(wt))
    return out


def edc_db(h):
    """Schroeder backward-integrated energy decay curve in dB (new metric, north_star)."""
    e = np.cumsum((np.asarray(h, dtype=np.float64) ** 2)[::-1])[::-1]
    e = e / max(e[0], 1e-300)
    return 10 * np.log10(np.maximum(e, 1e-30))


def rt60(h, sr=SR, lo=-5.0, hi=-25.0):
    """RT60 from a T20 (default) line fit of the EDC between `lo` and `hi` dB."""
    d = edc_db(h)
    idx = np.where((d <= lo) & (d >= hi))[0]
    if len(idx) < 2:
        return float("nan")
    t = idx / sr
    slope, _ = np.polyfit(t, d[idx], 1)
    return float(-60.0 / slope) if slope < 0 else float("nan")


def synthetic_rir(n, rng, rt60_s=None, length=9600, sr=SR):
    """Synthetic exponentially decaying noise RIRs: N(0,1)*exp(-6.91 t/(rt60 sr))."""
    if rt60_s is None:
        rt60_s = rng.uniform(0.05, 1.3, size=n)
    t = np.arange(length)[None, :]
    x = rng.standard_normal((n, length)) * np.exp(-6.91 * t / (np.asarray(rt60_s)[:, None] * sr))
    return x.astype(np.float32)
