"""Data-parallel check at world >= 2 (torchrun): the Dense gradient formed from all-gathered operands against the plain
all-reduced gradient, on identical replicas / shards / dropout masks. Prints the relative difference of the Dense
kernel and bias gradients after the first step and of all parameters after three steps, then step times of both modes.
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp_gather.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from bench import synthetic_batch
from unet_rir_b200.dl_models.u_net import UNet
from unet_rir_b200.main_training import DistributedTrainer, init_distributed

rank, world, local = init_distributed()
B = 64
x, y, e = [t.cuda() for t in synthetic_batch(B, 1 + rank)]
res = {}
for tag, mode in (("1", "1"), ("0", "0"), ("0b", "0")):
    os.environ["URIR_DP_GATHER_DENSE"] = mode
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    dt = DistributedTrainer(unet, per_replica_batch=B, alpha=0.9, lr=1e-4, loss="dp", use_cuda_graph=True)
    assert dt.gather_dense == (mode == "1")
    eng = unet.model.engine
    losses = [float(dt.train_step(x, e, y))]                            # eager step: gradients at identical weights
    torch.cuda.synchronize()
    g1 = (eng.grad["vec.dense.w"].clone(), eng.grad["vec.dense.b"].clone())
    if mode == "1":      # direct check, independent of run-to-run noise: gathered operands -> torch matmul
        xa, dya = [t.float() for t in dt._gathered]
        ref_w, ref_b = xa.t() @ dya, dya.sum(0)
        lx, ldy = [t.float().reshape(B, -1) for t in eng.dense_operands(B)]
        loc = lx.t() @ ldy; dist.all_reduce(loc)
        print(f"rank {rank}: kernel vs matmul(gathered) rel {float((g1[0] - ref_w).norm() / ref_w.norm()):.2e} "
              f"bias {float((g1[1] - ref_b).norm() / ref_b.norm()):.2e}  matmul(gathered) vs allreduce(local matmul) "
              f"{float((ref_w - loc).norm() / loc.norm()):.2e}", flush=True)
    losses += [float(dt.train_step(x, e, y)) for _ in range(2)]          # capture, replay
    torch.cuda.synchronize()
    res[tag] = (g1[0], g1[1], eng.P.clone(), losses)
    for _ in range(10):
        dt.train_step(x, e, y)
    torch.cuda.synchronize(); dist.barrier()
    s, t = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(50):
        dt.train_step(x, e, y)
    t.record(); torch.cuda.synchronize()
    if rank == 0:
        print("gather" if mode == "1" else "allreduce", "ms/step", s.elapsed_time(t) / 50, flush=True)
rel = lambda a, b: float((a - b).norm() / (b.norm() + 1e-30))
a, b, b2 = res["1"], res["0"], res["0b"]
print(f"rank {rank}: run-to-run noise of the all-reduce mode itself: dense.w rel {rel(b2[0], b[0]):.2e} dense.b rel {rel(b2[1], b[1]):.2e}", flush=True)
# replicas must agree with each other as well
chk = a[2].clone(); dist.all_reduce(chk, op=dist.ReduceOp.MAX); same = float((chk - a[2]).abs().max())
print(f"rank {rank}: dense.w rel {rel(a[0], b[0]):.2e}  dense.b rel {rel(a[1], b[1]):.2e}  params rel {rel(a[2], b[2]):.2e} "
      f"replica divergence {same:.1e}  losses {a[3]} vs {b[3]}", flush=True)
dist.destroy_process_group()
