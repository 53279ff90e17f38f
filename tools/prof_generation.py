"""Throughput of the generation path (BASELINE config 4) and achieved HBM GB/s of the signal kernels.
usage: python tools/prof_generation.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import signal_oracle as SO           # synthetic RIRs only (test infrastructure, not the timed path)
from unet_rir_b200 import rir_generation as RG
from unet_rir_b200.dl_models.u_net import UNet
from unet_rir_b200.postprocess import post_process_batch
from unet_rir_b200.preprocess import preprocess_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256


def timeit(fn, inner=10, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    return best


rng = np.random.default_rng(0)
wav = torch.as_tensor(SO.synthetic_rir(B, rng)).cuda()
spec = preprocess_batch(wav)
emb = torch.randint(0, 2000, (B, 2, 16), dtype=torch.int32).cuda()
unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
t = timeit(lambda: preprocess_batch(wav))
print(f"stft_ampphase   B={B}: {t*1e3:8.1f} us  {B/t*1e3:10.0f} samples/s  {(B*(9600*4 + 144*160*2*4))/t/1e6:7.0f} GB/s (38.4 kB in + 184.3 kB out per sample)")
t = timeit(lambda: post_process_batch(spec))
print(f"istft           B={B}: {t*1e3:8.1f} us  {B/t*1e3:10.0f} samples/s  {(B*(9600*4 + 144*160*2*4))/t/1e6:7.0f} GB/s")
t = timeit(lambda: unet.model([spec, emb], training=False), inner=5)
print(f"U-Net inference B={B}: {t*1e3:8.1f} us  {B/t*1e3:10.0f} samples/s  {B*9.076e9/t/1e9:7.1f} TFLOP/s (9.076 GF/sample)")
t = timeit(lambda: RG.generate_batch(unet, spec, emb), inner=5)
print(f"generate_batch  B={B}: {t*1e3:8.1f} us  {B/t*1e3:10.0f} samples/s  (spectrogram -> U-Net -> iSTFT waveform)")
