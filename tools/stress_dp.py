"""Stress: many steps of the data-parallel trainer's 4-graph step (world = 1, no NCCL) and of the single-graph Trainer,
to shake out timing-dependent protocol bugs. usage: python tools/stress_dp.py [steps] [mode: dp|single]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_batch
from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
from unet_rir_b200.dl_models.u_net import UNet
from unet_rir_b200.main_training import DistributedTrainer

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
mode = sys.argv[2] if len(sys.argv) > 2 else "dp"
B = 64
unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
x, y, e = synthetic_batch(B, 1)
if mode == "dp":
    dt = DistributedTrainer(unet, per_replica_batch=B, alpha=0.9, lr=5e-7, loss="dp", world=1)
    step = lambda: dt.train_step(x, e, y)
else:
    tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_stress", False, 0), EarlyStopping(5)], [False, 0], 1e-5, "s")
    step = lambda: tr.step(x, y, e, unet)[0]
for i in range(steps):
    l = step()
    if i % 50 == 0:
        print(i, float(l), flush=True)
torch.cuda.synchronize()
print("done", float(l))
