"""Clock stamps of the first tiles of a weight-gradient kernel (URIR_WGRAD_TRACE): producer / issuer barrier waits per iteration.
usage: python tools/trace_wgrad.py [prof_conv case, default wgrad_full]"""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
tr = torch.zeros(1024, dtype=torch.int64, device="cuda")
os.environ["URIR_WGRAD_TRACE"] = hex(tr.data_ptr())
from unet_rir_b200 import _lib as L
import tools.prof_conv as P
P.run(sys.argv[1] if len(sys.argv) > 1 else "wgrad_full", reps=1)
t = tr.cpu().view(256, 4)
t0 = int(t[0, 0])
print("it   prod_before_empty  prod_after_empty  mma_before_full  mma_after_full   (cycles since start)")
for it in list(range(0, 24)) + list(range(100, 108)):
    print(it, [int(v) - t0 for v in t[it]])
