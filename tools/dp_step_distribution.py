#!/usr/bin/env python
"""Step-time DISTRIBUTION of the data-parallel train step: per-step CUDA-event times of 200+ steps,
device-resident inputs, max over ranks per step, percentiles printed as one JSON line by rank 0.

    [URIR_DP_GRAPH=one|segments|off] [URIR_DP_GATHER_DENSE=1] [NCCL_MAX_CTAS=n] \
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29513 \
        tools/dp_step_distribution.py [--batch 64] [--steps 200] [--repeats 3] [--tag name]
World 1 works too (python tools/dp_step_distribution.py): the same step body without collectives."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--repeats", type=int, default=3)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from bench import synthetic_batch
    from unet_rir_b200.dl_models.u_net import UNet
    from unet_rir_b200.main_training import DistributedTrainer, init_distributed
    rank, world, local = init_distributed("nccl")
    dev = torch.device("cuda", local)
    B = args.batch
    unet = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3)
    dt = DistributedTrainer(unet, per_replica_batch=B, alpha=0.9, lr=5e-7, loss="dp", world=world, rank=rank)
    x, y, e = synthetic_batch(B, 100 + rank)
    for _ in range(4):
        dt.train_step(x, e, y)
    torch.cuda.synchronize()
    runs = []
    for rep in range(args.repeats):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        evs[0].record()
        for i in range(args.steps):
            dt._run(B)
            evs[i + 1].record()
        torch.cuda.synchronize()
        ms = torch.tensor([evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        runs.append(ms.cpu().numpy())
    if rank == 0:
        allms = np.concatenate(runs)
        out = {"tag": args.tag, "world": world, "per_gpu_batch": B, "graph": os.environ.get("URIR_DP_GRAPH", "one"),
               "gather_dense": os.environ.get("URIR_DP_GATHER_DENSE", "0"), "nccl_max_ctas": os.environ.get("NCCL_MAX_CTAS"),
               "steps": args.steps, "repeats": args.repeats,
               "mean_ms_per_run": [float(r.mean()) for r in runs],
               "p05": float(np.percentile(allms, 5)), "p50": float(np.percentile(allms, 50)),
               "p95": float(np.percentile(allms, 95)), "p99": float(np.percentile(allms, 99)), "max": float(allms.max()),
               "frac_above_1.1x_median": float((allms > 1.1 * np.median(allms)).mean()),
               "samples_per_s_at_mean": float(B * world / (allms.mean() * 1e-3))}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    dt._graphs.clear()
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
