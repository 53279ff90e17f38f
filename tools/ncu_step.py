"""One eager train step of the bench workload (B = 64) bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off ... python tools/ncu_step.py [batch]
so that ncu sees exactly the kernels of one step (forward, loss, backward, Adam, weight refresh)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synthetic_batch
from unet_rir_b200.amp_phase_trainer import EarlyStopping, ModelCheckpoint, Trainer
from unet_rir_b200.dl_models.u_net import UNet

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
eng = unet.model.engine
tr = Trainer(0.9, 1, "adam", [ModelCheckpoint("/tmp/urir_ncu", False, 0), EarlyStopping(5)], [False, 0], 1e-5, "ncu")
tr.use_cuda_graph = False
x, y, e = synthetic_batch(B, 1)
tr.step(x, y, e, unet)
torch.cuda.synchronize()
# record, per C-ABI call of the profiled step, how many liburir kernels it launched (urir_launch_count deltas):
# tools/ncu_step_table.py uses it to attribute ncu's per-kernel DRAM bytes to bench.py's call families
import json
from unet_rir_b200 import _lib as L
calls = []
_orig_call = L.call
def _counting_call(name, *args):
    n0 = L.launch_count(0)
    f0 = L.family_calls() if name.startswith("conv2d") else None
    _orig_call(name, *args)
    info = L._conv_info(name, args) if name.startswith("conv2d") else {}
    fam = None
    if f0 is not None:
        f1 = L.family_calls()
        fam = next((k for k in f1 if f1[k] != f0[k]), None)
    calls.append({"name": name, "tc": info.get("tc"), "family": fam, "kernels": L.launch_count(0) - n0})
L.call = _counting_call
import unet_rir_b200.engine as _E, unet_rir_b200.amp_phase_trainer as _T
torch.cuda.profiler.start()
tr._device_step(eng, B)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
L.call = _orig_call
os.makedirs("gpurun_out", exist_ok=True)
json.dump(calls, open("gpurun_out/step_calls.json", "w"))
print("ok", [float(v) for v in eng.losses_dev[:3]])
