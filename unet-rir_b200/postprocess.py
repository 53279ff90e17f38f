"""`PostProcess` with the reference's interface (postprocess.py:25-171); the un-pad -> denormalise ->
polar-to-complex -> inverse STFT chain (:78-129) is ONE liburir kernel on the GPU.

  PostProcess(folder, algorithm=None).post_process(feature, vector, des_shape=(129,151), n_fft=256,
                                                   win_length=128, hop_length=64, sr=48000) -> waveform
  post_process_batch(features) -> (B, n_samples) CUDA tensor     (what rir_generation's loop batches into)
`algorithm='gl'` (librosa.griffinlim, :130-131) is a batched GPU loop over the same STFT / iSTFT kernels
(`griffinlim_batch`).
Files are only written when `write_files=True` (the reference always writes, :73-74).
"""
from __future__ import annotations

import ctypes as C
import os
import pathlib

import numpy as np
import torch

from . import _lib as L
from .preprocess import Normalizer, TensorPadder, stft_desc


def post_process_batch(features, des_shape=(129, 151), n_fft=256, win_length=128, hop_length=64, normalized=True,
                       out=None):
    """(B, H_pad, W_pad, 2) normalised amp/phase (numpy or torch) -> (B, hop*(frames-1)) float32 CUDA tensor."""
    f = features if isinstance(features, torch.Tensor) else torch.as_tensor(np.asarray(features), dtype=torch.float32)
    f = f.to("cuda", torch.float32, non_blocking=True).contiguous()
    if f.dim() == 3:
        f = f[None]
    B, Hp, Wp, _ = f.shape
    n_samples = hop_length * (des_shape[1] - 1)
    d = stft_desc(n_samples, n_fft, win_length, hop_length, (Hp, Wp), "constant", False, normalized)
    if d.n_bins != des_shape[0] or d.n_frames != des_shape[1]:
        raise L.UrirError(f"des_shape {des_shape} is not the STFT shape ({d.n_bins}, {d.n_frames}) of n_fft={n_fft}")
    if out is None:
        out = torch.empty(B, n_samples, dtype=torch.float32, device="cuda")
    L.call("istft_from_ampphase", f.data_ptr(), B, C.byref(d), out.data_ptr())
    return out


def griffinlim_batch(amp, n_iter=32, momentum=0.99, init_angles=None, seed=None, n_fft=256, win_length=128,
                     hop_length=64, padded=(144, 160)):
    """librosa.griffinlim (fast Griffin-Lim: 32 iterations, momentum 0.99, random initial phases -- the defaults the
    reference relies on, postprocess.py:130-131) for a batch of magnitude spectrograms on the GPU.

    amp: (B, n_bins, n_frames) linear magnitudes (numpy or torch). Every iteration is one iSTFT launch, one STFT launch
    (liburir kernels, un-normalised mode) and a few elementwise torch ops on (B, n_bins, n_frames) tensors.
    init_angles: optional complex (B, n_bins, n_frames) unit phasors (parity tests inject them). -> (B, n_samples)."""
    S = amp if isinstance(amp, torch.Tensor) else torch.as_tensor(np.asarray(amp), dtype=torch.float32)
    S = S.to("cuda", torch.float32)
    if S.dim() == 2:
        S = S[None]
    B, nb, nf = S.shape
    n_samples = hop_length * (nf - 1)
    Hp, Wp = max(padded[0], nb), max(padded[1], nf)
    if init_angles is None:
        gen = torch.Generator(device="cuda")
        gen.manual_seed(int(seed) if seed is not None else int(np.random.randint(0, 2 ** 31 - 1)))
        phase = 2 * np.pi * torch.rand(B, nb, nf, generator=gen, device="cuda")
        angles = torch.polar(torch.ones_like(phase), phase)
    else:
        angles = torch.as_tensor(init_angles).to("cuda", torch.complex64)
    d_inv = stft_desc(n_samples, n_fft, win_length, hop_length, (Hp, Wp), "constant", False, False)
    spec = torch.zeros(B, Hp, Wp, 2, dtype=torch.float32, device="cuda")
    spec[:, :nb, :nf, 0] = S
    reb = torch.empty(B, Hp, Wp, 2, dtype=torch.float32, device="cuda")
    wav = torch.empty(B, n_samples, dtype=torch.float32, device="cuda")
    rebuilt = torch.zeros(B, nb, nf, dtype=torch.complex64, device="cuda")
    c = momentum / (1.0 + momentum)

    def inverse(ang):
        spec[:, :nb, :nf, 1] = torch.angle(ang)
        L.call("istft_from_ampphase", spec.data_ptr(), B, C.byref(d_inv), wav.data_ptr())

    for _ in range(n_iter):
        tprev = rebuilt
        inverse(angles)
        L.call("stft_ampphase", wav.data_ptr(), B, C.byref(d_inv), reb.data_ptr())
        rebuilt = torch.polar(reb[:, :nb, :nf, 0], reb[:, :nb, :nf, 1])
        angles = rebuilt - c * tprev
        angles = angles / (angles.abs() + 1e-16)
    inverse(angles)
    return wav


class PostProcess:

    def __init__(self, folder, algorithm=None, write_files=False):
        self.stft = None
        self.phase = None
        self.waveform = None
        self.normalizer = Normalizer()
        self.padder = TensorPadder((144, 160))
        self.algorithm = 'gl' if algorithm == 'gl' else 'ph'
        self.write_files = write_files
        self.wav_path = "../generated_rir_distributed/" + folder + f'_{self.algorithm}'

    def post_process(self, feature, vector, des_shape=(129, 151),
                     n_fft=256, win_length=128, hop_length=64, sr=48000):
        if self.algorithm == 'gl':          # magnitudes only: un-pad, denormalise, Griffin-Lim phase retrieval
            f = feature.detach().float().cpu().numpy() if isinstance(feature, torch.Tensor) else np.asarray(feature, dtype=np.float32)
            a, p = self.padder.un_pad(f[:, :, 0], f[:, :, 1], des_shape)
            a, _ = self.normalizer.denormalize(np.asarray(a, dtype=np.float32), np.asarray(p, dtype=np.float32))
            self.waveform = griffinlim_batch(a[None], n_fft=n_fft, win_length=win_length, hop_length=hop_length)[0].cpu().numpy()
            if self.write_files:
                self.save_wav(sr, vector)
                self.save_stft(f)
            return self.waveform
        if isinstance(feature, torch.Tensor):
            feature_t = feature.detach()
        else:
            feature_t = np.asarray(feature, dtype=np.float32)
        self.waveform = post_process_batch(feature_t, des_shape, n_fft, win_length, hop_length)[0].cpu().numpy()
        if self.write_files:
            self.save_wav(sr, vector)
            self.save_stft(np.asarray(feature_t.cpu() if isinstance(feature_t, torch.Tensor) else feature_t))
        return self.waveform

    @staticmethod
    def get_stft_phase(feature):
        return feature[:, :, 0], feature[:, :, 1]

    def de_shape(self, stft, phase, des_shape):
        return self.padder.un_pad(stft, phase, des_shape)

    def denormalize(self, stft, phase):
        return self.normalizer.denormalize(stft, phase)

    def istft(self, denorm_f, denorm_p, n_fft, win_length, hop_length):
        """amp, phase (n_bins, n_frames), already un-padded and denormalised -> self.waveform; Griffin-Lim phase
        retrieval from the magnitudes alone when algorithm == 'gl' (postprocess.py:128-131)."""
        if self.algorithm == 'gl':
            a = np.asarray(denorm_f, dtype=np.float32)
            self.waveform = griffinlim_batch(a[None], n_fft=n_fft, win_length=win_length, hop_length=hop_length,
                                             padded=a.shape)[0].cpu().numpy()
            return
        feat = np.stack([np.asarray(denorm_f, dtype=np.float32), np.asarray(denorm_p, dtype=np.float32)], axis=-1)
        self.waveform = post_process_batch(feat, feat.shape[:2], n_fft, win_length, hop_length,
                                           normalized=False)[0].cpu().numpy()

    def save_wav(self, sr, vector):
        from scipy.io.wavfile import write
        vector_name = ""
        for value in vector:
            vector_name += "-" + str(int(value))
        self.wav_name = "RIR" + vector_name
        self._create_directory_if_none(self.wav_path + "/rir/")
        write(os.path.join(self.wav_path + "/rir/", self.wav_name + ".wav"), sr, self.waveform)

    def save_stft(self, feature):
        self._create_directory_if_none(self.wav_path + "/stft/")
        np.save(os.path.join(self.wav_path + "/stft/", self.wav_name) + ".npy", feature)

    @staticmethod
    def _create_directory_if_none(dir_path):
        directory = pathlib.Path(dir_path)
        if not directory.exists():
            os.makedirs(dir_path)
