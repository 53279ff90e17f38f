"""Step time of the data-parallel trainer's 4-graph step at world = 1 (no NCCL) next to the single-graph Trainer, same
device-resident inputs: isolates what the DP step costs beyond the collectives. usage: python tools/time_dp_world1.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_batch
from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
from unet_rir_b200.dl_models.u_net import UNet
from unet_rir_b200.main_training import DistributedTrainer

B = 64
x, y, e = [t.cuda() for t in synthetic_batch(B, 1)]
for mode in ("single", "dp", "dp_amp"):
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    if mode == "single":
        tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_t", False, 0), EarlyStopping(5)], [False, 0], 1e-5, "s")
        step = lambda: tr.step(x, y, e, unet)[0]
    else:
        dt = DistributedTrainer(unet, per_replica_batch=B, alpha=0.9, lr=5e-7, loss="dp" if mode == "dp" else "amp", world=1)
        step = lambda: dt.train_step(x, e, y)
    for _ in range(10):
        step()
    torch.cuda.synchronize()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        step()
    t.record()
    torch.cuda.synchronize()
    print(mode, "ms/step", s.elapsed_time(t) / 50, flush=True)
