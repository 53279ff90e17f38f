"""`Dataset` with the reference's constants and item contract (dataset.py:11-244).

Two sources:
  * `dir_dataset=None`: synthetic RIRs (exponentially decaying noise) -- what the benchmarks and most tests use, the
    real UTS dataset does not exist here;
  * `dir_dataset=<path>`: the reference's directory walk (dataset.py:123-182): `<dir>/<name>/<room>/<zone>/<array>/
    <Room>_Zone<Z>_<Type>MicrophoneArray_L<l>_M<m>.wav`, filtered by `room` / `array`, room-geometry embeddings from
    rooms.py, per-room index lists, `index_out` = the per-room lists shuffled with seed 500.
Kept from the reference in both cases: the STFT constants (:62-70), `__getitem__(i) -> (amp, phase, emb)` with
amp / phase (144,160) normalised + padded spectrograms and emb a 16-int vector, `seed = 500` (:76) and
`return_characteristics()`. Spectrograms are produced by the GPU pre-processing kernel for the whole dataset in batched
launches (the reference runs librosa per file on the CPU, :214-223).
"""
from __future__ import annotations

import os
import random

import numpy as np

from .preprocess import Loader, preprocess_batch
from .rooms import UTS_ROOMS, uts_room


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

        self.rooms = (['HemiAnechoicRoom', 'LargeMeetingRoom', 'MediumMeetingRoom', 'ShoeBoxRoom', 'SmallMeetingRoom']
                      if room in (None, ['All']) else list(room))                                   # dataset.py:35-38
        self.array = ['PlanarMicrophoneArray', 'CircularMicrophoneArray'] if array is None else list(array)   # :19-22
        if dir_dataset is not None:
            if extract:
                self.extract_files()
            self._load_directory()
            return
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

    # ------------------------------------------------------------------ the reference's directory walk
    def extract_files(self):
        """Unzip `<array>.zip` archives in place (dataset.py:92-113)."""
        import zipfile
        root = os.path.join(self.dir_dataset, self.dataset_name)
        for room_folder in sorted(os.listdir(root)):
            for zone_folder in sorted(os.listdir(os.path.join(root, room_folder))):
                zone_path = os.path.join(root, room_folder, zone_folder)
                for entry in sorted(os.listdir(zone_path)):
                    if entry.endswith(".zip"):
                        with zipfile.ZipFile(os.path.join(zone_path, entry), 'r') as z:
                            z.extractall(zone_path)
                        os.remove(os.path.join(zone_path, entry))

    def _load_directory(self):
        """dataset.py:121-182. Folder entries are visited in sorted order (the reference takes os.listdir order, which is
        filesystem-dependent). With `debugging`, loading stops after the first array folder that contributed files."""
        loader = Loader(sample_rate=self.sr, duration=self.duration, mono=self.mono)
        n_samples = int(round(self.duration * self.sr))
        root = os.path.join(self.dir_dataset, self.dataset_name)
        by_room = {name: [] for name in UTS_ROOMS}
        wavs, embs, chars = [], [], []
        done = False
        for room_folder in sorted(os.listdir(root)):
            if done:
                break
            for zone_folder in sorted(os.listdir(os.path.join(root, room_folder))):
                if done:
                    break
                zone_path = os.path.join(root, room_folder, zone_folder)
                for array_folder in sorted(os.listdir(zone_path)):
                    array_path = os.path.join(zone_path, array_folder)
                    if not os.path.isdir(array_path):
                        continue
                    matched = False
                    for rir_file in sorted(os.listdir(array_path)):
                        ch = rir_file.split('_')
                        if len(ch) < 5 or ch[0] not in self.rooms or ch[2] not in self.array or ch[0] not in UTS_ROOMS:
                            continue
                        ch = [ch[0], ch[1].replace('Zone', ''), ch[2].replace('MicrophoneArray', ''),
                              ch[3].replace('L', ''), ch[4].replace('M', '').replace('.wav', '')]
                        w = loader.load(os.path.join(array_path, rir_file))
                        if len(w) < n_samples:                       # short file: zero tail (amplitude floor)
                            w = np.concatenate([w, np.zeros(n_samples - len(w), dtype=np.float32)])
                        by_room[ch[0]].append(len(wavs))
                        wavs.append(w.astype(np.float32))
                        embs.append(uts_room(ch[0]).return_embedding(ch))
                        chars.append(ch)
                        matched = True
                    if self.debugging and matched:
                        done = True
                        break
        if not wavs:
            raise FileNotFoundError(f"no RIR files of rooms {self.rooms} / arrays {self.array} under {root}")
        self.wavs = np.stack(wavs)
        self.rt60 = np.array([e[15] / 1000.0 for e in embs])
        amp, phase = [], []
        for lo in range(0, len(wavs), 1024):                     # whole-dataset preprocessing in batched GPU launches
            spec = preprocess_batch(self.wavs[lo:lo + 1024], padded=self.input_shape, normalized=self.normalization,
                                    remove_mean=False).cpu().numpy()           # Loader.load already removed the mean
            amp.append(spec[..., 0]); phase.append(spec[..., 1])
        self.amp, self.phase = np.concatenate(amp), np.concatenate(phase)
        self.emb = np.asarray(embs, dtype=np.int32)
        self.characteristics = chars if self.room_characteristics else None
        # :173-182 -- inputs in room order, outputs = the same rooms' items shuffled with the fixed seed (the anechoic
        # room gets an embedding but, as in the reference, no index list)
        order = ["HemiAnechoicRoom", "LargeMeetingRoom", "MediumMeetingRoom", "SmallMeetingRoom", "ShoeBoxRoom"]
        self.index_in = [i for name in order for i in by_room[name]]
        self.index_out = []
        for name in order:
            lst = list(by_room[name])
            random.Random(self.seed).shuffle(lst)
            self.index_out += lst

    def __len__(self):
        return len(self.amp)

    def __getitem__(self, i):
        return self.amp[i], self.phase[i], self.emb[i]

    def return_characteristics(self):
        return self.characteristics

    def waveform(self, i):
        """mean-removed waveform of item i (what Loader.load returns for the item's file)."""
        w = self.wavs[i]
        return w - np.mean(w)
