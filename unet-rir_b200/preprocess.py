"""Signal pre-processing with the reference's class names (preprocess.py:7-121), computed by liburir's
STFT kernel on the GPU instead of librosa on the CPU.

  FeatureExtractor(n_fft, win_length, hop_length).extract(waveform) -> (amp, phase)   (:13-18)
  Normalizer().normalize / denormalize                                               (:26-41)
  Loader(sample_rate, duration, mono).load(path)                                     (:51-57)
  TensorPadder(desired_shape).pad_amp_phase / transform / un_pad                     (:70-113)
  sigmoid(beta, dimensions)                                                          (:116-121)
plus the batched fused entry point the hot path uses:
  preprocess_batch(wavs) == Dataset.preprocess (dataset.py:214-223) for a whole batch: mean removal,
  STFT, |.|/angle, normalisation and zero-padding to (144, 160) in ONE kernel launch.
Per-sample class methods accept and return numpy arrays like the reference; preprocess_batch takes
numpy or torch and returns a CUDA tensor (B, 144, 160, 2).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

PAD_MODE = {"constant": 0, "reflect": 1}


def stft_desc(n_samples, n_fft=256, win_length=128, hop_length=64, padded=(144, 160), pad_mode="constant",
              remove_mean=True, normalized=True):
    n_frames = 1 + n_samples // hop_length
    n_bins = n_fft // 2 + 1
    H_pad = padded[0] if padded else n_bins
    W_pad = padded[1] if padded else n_frames
    return L.StftDesc(n_fft, win_length, hop_length, n_samples, n_bins, n_frames, H_pad, W_pad,
                      PAD_MODE[pad_mode], int(remove_mean), int(normalized))


def preprocess_batch(wavs, padded=(144, 160), pad_mode="constant", remove_mean=True, normalized=True,
                     n_fft=256, win_length=128, hop_length=64, out=None):
    """(B, T) waveforms -> (B, H_pad, W_pad, 2) float32 CUDA tensor, channel 0 = (normalised log-)amplitude,
    channel 1 = (normalised) phase: Loader's mean removal + FeatureExtractor + Normalizer + TensorPadder."""
    w = wavs if isinstance(wavs, torch.Tensor) else torch.as_tensor(np.asarray(wavs), dtype=torch.float32)
    w = w.to("cuda", torch.float32, non_blocking=True).contiguous()
    if w.dim() == 1:
        w = w[None]
    B, T = w.shape
    d = stft_desc(T, n_fft, win_length, hop_length, padded, pad_mode, remove_mean, normalized)
    if out is None:
        out = torch.empty(B, d.H_pad, d.W_pad, 2, dtype=torch.float32, device="cuda")
    L.call("stft_ampphase", w.data_ptr(), B, C.byref(d), out.data_ptr())
    return out


class FeatureExtractor:
    def __init__(self, n_fft, win_length, hop_length):
        self.n_fft = n_fft
        self.win_length = win_length
        self.hop_length = hop_length

    def extract(self, waveform):
        spec = preprocess_batch(np.asarray(waveform, dtype=np.float32)[None], padded=None, remove_mean=False,
                                normalized=False, n_fft=self.n_fft, win_length=self.win_length,
                                hop_length=self.hop_length)[0].cpu().numpy()
        return spec[:, :, 0], spec[:, :, 1]


class Normalizer:
    def __init__(self):
        self.md = 100
        self.ep = 10 ** (-1 * self.md / 20)

    def normalize(self, amp, phase):
        amp_norm = 20 * np.log10(amp / (128) + self.ep)
        amp_norm = (amp_norm + self.md) / self.md
        phase_norm = (phase + math.pi) / (2 * math.pi)
        return amp_norm, phase_norm

    def denormalize(self, amp_norm, phase_norm):
        amp = (amp_norm * self.md) - self.md
        amp = (10 ** (amp / 20) - self.ep) * (128)
        phase = (phase_norm * 2 * math.pi) - math.pi
        phase = (phase + math.pi) % (2 * math.pi) - math.pi
        return amp, phase


class Loader:
    """librosa.load(path, sr, duration, mono) is replaced by scipy.io.wavfile + polyphase resampling
    (librosa / soundfile are not installed); then `signal -= mean` as in the reference (:56)."""

    def __init__(self, sample_rate, duration, mono):
        self.sample_rate = sample_rate
        self.duration = duration
        self.mono = mono

    def load(self, file_path):
        from scipy.io import wavfile
        from scipy.signal import resample_poly
        sr, data = wavfile.read(file_path)
        if data.dtype.kind == "i":
            data = data.astype(np.float32) / float(np.iinfo(data.dtype).max + 1)
        elif data.dtype.kind == "u":
            data = (data.astype(np.float32) - 128.0) / 128.0
        data = data.astype(np.float32)
        if data.ndim == 2:
            data = data.mean(axis=1) if self.mono else data[:, 0]
        if sr != self.sample_rate:
            g = math.gcd(int(sr), int(self.sample_rate))
            data = resample_poly(data, self.sample_rate // g, sr // g).astype(np.float32)
        n = int(round(self.duration * self.sample_rate))
        signal = data[:n]
        signal = signal - np.mean(signal)
        return signal


class TensorPadder:

    def __init__(self, desired_shape):
        self.current_shape = None
        self.desired_shape = desired_shape
        self.c_rows = None
        self.c_columns = None
        self.n_rows = None
        self.n_columns = None

    def pad_amp_phase(self, amp, phase):
        return self.transform(amp), self.transform(phase)

    def transform(self, tensor):
        if self.get_needed_transform(tensor):
            return self.col_transform(self.row_transform(tensor))
        return tensor

    def get_needed_transform(self, tensor):
        self.current_shape = tensor.shape
        self.c_rows, self.c_columns = tensor.shape[0], tensor.shape[1]
        if tensor.shape[0] > self.desired_shape[0] or tensor.shape[1] > self.desired_shape[1]:
            return False
        self.n_rows = self.desired_shape[0] - tensor.shape[0]
        self.n_columns = self.desired_shape[1] - tensor.shape[1]
        return True

    def row_transform(self, tensor):
        return np.r_[tensor, np.zeros((self.n_rows, self.c_columns))]

    def col_transform(self, tensor):
        # one allocation instead of the reference's per-column np.c_ loop (:100-105); same result
        return np.c_[tensor, np.zeros((self.desired_shape[0], self.n_columns))]

    @staticmethod
    def un_pad(amp, phase, desired_shape):
        amp_d = np.asarray(amp)[:desired_shape[0], :desired_shape[1]]
        phase_d = np.asarray(phase)[:desired_shape[0], :desired_shape[1]]
        return amp_d, phase_d


def sigmoid(beta, dimensions):
    x = np.linspace(-10, 10, dimensions[1])
    z = 1 / (1 + np.exp(-(x + 5) * beta))
    z = np.flip(z)
    return np.tile(z, (dimensions[0], 1))
