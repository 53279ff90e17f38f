"""Layer-by-layer comparison of backward activations (gradients w.r.t. raw conv outputs) device vs oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L
from unet_rir_b200.engine import UNetEngine
from test_gpu_model import _setup

B = 2
om, params, x, y, emb, mask = _setup(B=B, kernels=3)
oq = O.UNetOracle(kernels=3, emulate_bf16=True)
oq.taps = {}
names = O.trainable_names(oq.plan)
leaf = {n: t.detach().clone().requires_grad_(n in names) for n, t in params.items()}
out = oq.forward(leaf, x, emb, training=True, dropout_mask=mask, new_stats={})
for t in oq.taps.values():
    t.retain_grad()
loss, lp, ls = O.amp_phase_loss(y, out)
loss.backward()
eng = UNetEngine(kernels=3, impl=L.IMPL_SIMT if len(sys.argv) > 1 else L.IMPL_AUTO)
eng.load_state_dict(params)
o = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
n = B * 144 * 160
eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
eng.backward(eng._buffers(B)["g_out"])
torch.cuda.synchronize()
b = eng._buffers(B)
dbg = eng.debug_tensors()
print("forward taps (device vs bf16-contract oracle):")
for k in ["enc1.down", "enc1.blk.c1", "enc3.blk.c1", "enc5.blk.c1", "bottleneck", "dec2.up", "dec2.fuse", "dec3.blk.c1", "dec5.fuse", "dec5.blk.c1"]:
    print(f"  {k:14s} rel {U.rel_l2(dbg[k], oq.taps[k].detach()):.5f}")
print("out rel", U.rel_l2(o.float(), out.detach()))
print("backward taps:")
pairs = [("head", None)]
order = []
for j in (5, 4, 3, 2):
    i = 6 - j
    nn = 32 * 2 ** (i - 1)
    order += [(f"dec{j}.blk.c1", b[f"g_rb{i}"]), (f"dec{j}.fuse", b[f"g_rf{i}"]), (f"dec{j}.up", b[f"g_cat{i}"][..., nn:])]
order += [("bottleneck", b["g_z"])]
for i in (5, 4, 3, 2, 1):
    order += [(f"enc{i}.blk.c1", b[f"g_r{i}"]), (f"enc{i}.down", b[f"g_t{i}"])]
for k, dev in order:
    ref = oq.taps[k].grad.permute(0, 2, 3, 1)
    print(f"  {k:14s} rel {U.rel_l2(dev.float(), ref):.5f}   |ref| {float(ref.norm()):.3e}")
