"""Per-tensor PARAMETER gradients, engine vs oracle with the device forward state substituted (debug aid;
tools/grad_check_activations.py compares the gradients w.r.t. the raw conv outputs layer by layer)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import unet_oracle as O
import urir_testutil as U
from unet_rir_b200 import _lib as L
from unet_rir_b200.engine import UNetEngine
from test_gpu_model import _setup

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
om, params, x, y, emb, mask = _setup(B=B, kernels=3)
st = O.new_opt_state(params, om.plan)
(loss, lp, ls), _, ref_out = O.train_step(om, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
oq = O.UNetOracle(kernels=3, emulate_bf16=True)
_, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
for impl in (L.IMPL_SIMT, L.IMPL_AUTO):
    eng = UNetEngine(kernels=3, impl=impl)
    eng.load_state_dict(params)
    out = eng.forward(x.cuda(), emb.cuda(), training=True, dropout_mask=mask.cuda())
    n = B * 144 * 160
    eng.loss_and_grad(y.cuda(), 1.0 / n, 1.0 / n)
    eng.backward(eng._buffers(B)["g_out"])
    torch.cuda.synchronize()
    # reference gradients: autograd through the oracle graph with the DEVICE's forward state substituted
    oq.override = eng.forward_state()
    _, grads, _ = O.train_step(oq, params, st, x, y, emb, 1e-3, dropout_mask=mask, apply=False)
    oq.override = None
    print("impl", impl, "out rel", U.rel_l2(out.float(), ref_out))
    for name in eng.trainable_names():
        got, ref = eng.grad[name].cpu(), grads[name]
        print(f"  {name:24s} rel {U.rel_l2(got, ref):9.4f}  maxabs {U.max_abs(got, ref):10.3e}  scale {float(ref.abs().max()):10.3e}")
