"""`PostProcess` with the reference's interface (postprocess.py:25-171); the un-pad -> denormalise ->
polar-to-complex -> inverse STFT chain (:78-129) is ONE liburir kernel on the GPU.

  PostProcess(folder, algorithm=None).post_process(feature, vector, des_shape=(129,151), n_fft=256,
                                                   win_length=128, hop_length=64, sr=48000) -> waveform
  post_process_batch(features) -> (B, n_samples) CUDA tensor     (what rir_generation's loop batches into)
`algorithm='gl'` (Griffin-Lim, :130-131) is a SURVEY 8(f) "next" item and raises NotImplementedError.
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
        if self.algorithm == 'gl':
            raise NotImplementedError("Griffin-Lim synthesis (postprocess.py:130-131) is not built yet")
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
        """amp, phase (n_bins, n_frames), already un-padded and denormalised -> self.waveform."""
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
