"""`Dataset` with the reference's constants and item contract (dataset.py:11-244), fed by synthetic RIRs.

The reference walks `room/zone/array/*.wav` files and builds room-geometry embeddings (dataset.py:123-223,
rooms.py); that file walking is out of scope and the dataset directory does
not exist here. What the hot path needs is kept: the STFT constants (:62-70), `__getitem__(i) ->
(amp, phase, emb)` with amp/phase (144,160) normalised + padded spectrograms and emb a 16-int vector,
`index_in/index_out` pairs made by a seeded shuffle (:173-182), `seed = 500` (:76) and
`return_characteristics()`. Spectrograms are produced by the GPU pre-processing kernel for the whole
dataset in one batched launch (the reference runs librosa per file on the CPU, :214-223).
"""
from __future__ import annotations

import random

import numpy as np

from .preprocess import preprocess_batch


def synthetic_rirs(n, seed=500, length=9600, sr=48000):
    """Exponentially decaying noise, rt60 ~ U(0.05, 1.3) s (the range of dataset.py:86-91)."""
    rng = np.random.default_rng(seed)
    rt60 = rng.uniform(0.05, 1.3, size=n)
    t = np.arange(length)[None, :]
    x = rng.standard_normal((n, length)) * np.exp(-6.91 * t / (rt60[:, None] * sr))
    return x.astype(np.float32), rt60


class Dataset:
    def __init__(self, dir_dataset, dataset_name, normalization=True, debugging=False, extract=False,
                 room_characteristics=False, room=None, array=None, n_synthetic=256, seed=500):
        self.dir_dataset = dir_dataset
        self.dataset_name = dataset_name

        'Constants'
        self.n_fft = 256
        self.win_length = 128
        self.hop_length = 64
        self.duration = 0.2  # in seconds
        self.sr = 48000
        self.mono = True
        self.input_shape = (144, 160)

        self.normalization = normalization
        self.debugging = debugging
        self.room_characteristics = room_characteristics
        self.seed = seed  # Seed for consistency at selecting training / validation and test datasets

        if dir_dataset is not None:
            raise NotImplementedError("reading the room_impulse directory tree (dataset.py:123-223) is not built; "
                                      "pass dir_dataset=None for synthetic RIRs")
        n = int(n_synthetic)
        self.wavs, self.rt60 = synthetic_rirs(n, seed, int(self.duration * self.sr), self.sr)
        spec = preprocess_batch(self.wavs, padded=self.input_shape, normalized=normalization).cpu().numpy()
        self.amp, self.phase = spec[..., 0], spec[..., 1]
        rng = np.random.default_rng(seed + 1)
        # 16-int vectors with the reference's value range (< 2000: Embedding vocab, u_net.py:257); the last
        # entry is rt60 in ms like rooms.py:94
        self.emb = rng.integers(0, 1282, size=(n, 16)).astype(np.int32)
        self.emb[:, 15] = np.clip((self.rt60 * 1000).astype(np.int32), 0, 1999)
        self.characteristics = [["SyntheticRoom", "A", "Planar", int(i), int(i)] for i in range(n)]
        # in -> out pairs: a seeded shuffle of the same room's items (dataset.py:173-182)
        self.index_in = list(range(n))
        self.index_out = list(range(n))
        random.Random(self.seed).shuffle(self.index_out)

    def __len__(self):
        return len(self.index_in)

    def __getitem__(self, i):
        return self.amp[i], self.phase[i], self.emb[i]

    def return_characteristics(self):
        return self.characteristics

    def waveform(self, i):
        """mean-removed waveform of item i (what Loader.load returns for the item's file)."""
        w = self.wavs[i]
        return w - np.mean(w)
