"""Graph-replayed Trainer.step time at B = 64 (device-resident inputs), best and median of several 50-step loops: the quick
A/B tool for kernel-selection switches (URIR_* environment variables). usage: python tools/step_time.py [batch]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_batch
from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
from unet_rir_b200.dl_models.u_net import UNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x, y, e = [t.cuda() for t in synthetic_batch(B, 1)]
unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_t", False, 0), EarlyStopping(5)], [False, 0], 1e-5, "s")
for _ in range(10):
    tr.step(x, y, e, unet)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        tr.step(x, y, e, unet)
    t.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(t) / 50)
ts.sort()
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("URIR_"))
print(f"step B={B} [{tag}] best {ts[0]:.3f} ms  median {ts[2]:.3f} ms", flush=True)
