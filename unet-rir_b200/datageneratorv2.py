"""`DataGenerator` with the reference's batch contract (datageneratorv2.py:8-102).

__getitem__(idx) -> (spec_in f32 (B,144,160,2), emb i32 (B,2,16), spec_out f32 (B,144,160,2)[, characteristic]);
emb[:,0] = source vector, emb[:,1] = target vector; 70/20/10 train/val/test split of a seeded shuffle
(:25-43); __len__ = N // B. The reference defines __iter__ without __next__ although Trainer.train calls
__next__ and unpacks (spec_in, spec_out, emb) (amp_phase_trainer.py:65): __next__ is provided here in
that order, and __getitem__ keeps the (spec_in, emb, spec_out) order main_training.py:85 uses.
"""
from __future__ import annotations

import random

import numpy as np


class DataGenerator:

    def __init__(self, dataset, batch_size=32, partition='train', shuffle=True, characteristics=False):
        self.dataset = dataset
        self.batch_size = batch_size
        self.partition = partition
        self.shuffle = shuffle
        self.characteristics = characteristics

        self._idx = 0
        temp = list(zip(dataset.index_in, dataset.index_out))
        random.Random(dataset.seed).shuffle(temp)
        index_in, index_out = zip(*temp)
        self.index_in, self.index_out = list(index_in), list(index_out)
        self.characteristics_list = self.dataset.return_characteristics()

        n = len(self.index_in)
        if partition == 'train':
            sl = slice(0, int(0.7 * n))
        elif partition == 'val':
            sl = slice(int(0.7 * n), int(0.9 * n))
        elif partition == 'test':
            sl = slice(int(0.9 * n), n)
        else:
            sl = slice(0, n)
        self.index_in, self.index_out = self.index_in[sl], self.index_out[sl]

    def __len__(self):
        return int(len(self.index_in) // self.batch_size)

    def __iter__(self):
        return self

    def __next__(self):
        if self._idx >= len(self):
            self._idx = 0
            self.on_epoch_end()
        spec_in, emb, spec_out = self.__getitem__(self._idx)[:3]
        self._idx += 1
        return spec_in, spec_out, emb

    def on_epoch_end(self):
        if self.shuffle:
            temp = list(zip(self.index_in, self.index_out))
            random.shuffle(temp)
            index_in, index_out = zip(*temp)
            self.index_in, self.index_out = list(index_in), list(index_out)

    def __getitem__(self, idx):
        lo, hi = idx * self.batch_size, (idx + 1) * self.batch_size
        stft_in, phase_in, emb_in, char_in = [], [], [], []
        stft_out, phase_out, emb_out, char_out = [], [], [], []
        for i in range(lo, hi):
            a, p, e = self.dataset.__getitem__(self.index_in[i])
            stft_in.append(a); phase_in.append(p); emb_in.append(e)
            a, p, e = self.dataset.__getitem__(self.index_out[i])
            stft_out.append(a); phase_out.append(p); emb_out.append(e)
            if self.characteristics:
                char_in.append(self.characteristics_list[self.index_in[i]])
                char_out.append(self.characteristics_list[self.index_out[i]])
        spectrogram_in = np.stack((stft_in, phase_in), axis=-1).astype('float32')
        spectrogram_out = np.stack((stft_out, phase_out), axis=-1).astype('float32')
        embedding = np.stack((emb_in, emb_out), axis=1).astype('int32')
        if self.characteristics:
            characteristic = np.stack((char_in, char_out), axis=2)
            return spectrogram_in, embedding, spectrogram_out, characteristic
        return spectrogram_in, embedding, spectrogram_out
