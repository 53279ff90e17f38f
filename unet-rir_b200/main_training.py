"""Data-parallel training with the semantics of the reference's main_training.py, one process per
GPU (torchrun) instead of tf.distribute.MirroredStrategy (main_training.py:56-60, 203-235, 253-327).

What is kept exactly:
  * per-replica batch (16 in the reference, :44) and global_batch = per_replica * replicas (:60);
  * compute_loss (:203-235): alpha*(a_t-a_p)^2 + (1-alpha)*(1-cos(wrap(2*pi*(p_t-p_p)))) per element,
    divided by H*W*2 and by the GLOBAL batch, plus sum(l2 losses)/replicas, so that the SUM of the
    per-replica gradients is the gradient of the global mean loss;
  * gradients are sum-all-reduced across replicas before Adam (:268, inside apply_gradients);
  * BatchNormalization statistics stay per replica (plain BN under MirroredStrategy, u_net.py:368);
  * Adam lr 5e-7 with lr*0.9^(epoch/80) from epoch 80 (:43, :342-344).
What changes: the all-reduce is NCCL over NVLink through torch.distributed, issued per gradient
bucket from a side stream as soon as the backward segment that produces the bucket has been
enqueued (decoder+head, then bottleneck, then encoder -- reverse layer order), so the gradient traffic
overlaps the remaining backward kernels. Each segment is a CUDA graph. The Dense layer's kernel gradient (47 of
the 84 MB) can instead be formed from all-gathered operands (1.2 MB per replica; URIR_DP_GATHER_DENSE=1): same
result, 56 % less NVLink traffic, but no faster on NVSwitch, so the plain all-reduce stays the default.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import plan as PL
from .dl_models.u_net import UNet


def init_distributed(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_batch(global_batch_arrays, rank, world):
    """experimental_distribute_dataset (:114): contiguous B/world slices of the global batch."""
    out = []
    for a in global_batch_arrays:
        n = len(a)
        per = n // world
        out.append(a[rank * per:(rank + 1) * per])
    return out


def gradient_buckets(offsets, n_flat):
    """Flat-gradient ranges in the order backward finishes them: (decoder+head), (vector block),
    (encoder). `offsets` is the engine's name -> (offset, size) map in Keras variable order."""
    o_vec = offsets["vec.emb"][0]
    o_dec = offsets["dec2.up.w"][0]
    return [(o_dec, n_flat), (o_vec, o_dec), (0, o_vec)]


def lr_schedule(lr0, epoch, start=80):
    """main_training.py:342-344."""
    return lr0 * 0.9 ** (epoch / start) if epoch >= start else lr0


class DistributedTrainer:
    """One replica of the synchronous data-parallel train step."""

    def __init__(self, model: UNet, per_replica_batch=16, alpha=0.9, lr=5e-7, loss="dp", world=None,
                 use_cuda_graph=True, dropout=True):
        self.model = model
        self.eng = model.model.engine
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.per_replica_batch = per_replica_batch
        self.global_batch = per_replica_batch * self.world
        self.alpha, self.lr, self.loss = alpha, lr, loss
        self.use_cuda_graph, self.dropout = use_cuda_graph, dropout
        self.buckets = gradient_buckets(self.eng.offsets, self.eng.n_flat)
        self._graphs = None
        self._calls = 0
        self.overlap = os.environ.get("URIR_DP_OVERLAP", "1") != "0"
        # Dense layer (56 % of all parameters): all-gather its two small operands and form the global-batch kernel
        # gradient locally instead of all-reducing 47 MB of fp32 gradient (see engine.dense_grad_from_gathered)
        # (opt-in: measured equal to the plain all-reduce on NVSwitch at 2 and 8 GPUs, profiles/r01_layer_roofline.txt)
        self.gather_dense = self.world > 1 and self.overlap and os.environ.get("URIR_DP_GATHER_DENSE", "0") == "1"
        o = self.eng.offsets
        self._o_dense = o["vec.dense.w"][0]
        self._o_after_dense = o["vec.dense.b"][0] + o["vec.dense.b"][1]
        self._gathered = None
        self.eng.set_lr(lr)

    # loss weights: w_amp * sum sq err + w_ph * sum (1 - cos)
    def _weights(self, B):
        H, W, _ = self.eng.input_shape
        if self.loss == "dp":                       # compute_loss, main_training.py:226-231
            den = H * W * 2 * self.global_batch
            return self.alpha / den, (1.0 - self.alpha) / den
        n = self.global_batch * H * W               # amp_phase_trainer.py:155 averaged over the global batch
        return 1.0 / n, 1.0 / n

    # ---- the three backward segments and the optimiser, as separately capturable bodies
    def _seg_forward_loss_decoder(self, B):
        e = self.eng
        b = e._buffers(B)
        e._forward_body(B, training=True, dropout=self.dropout)
        wa, wp = self._weights(B)
        L.call("ampphase_loss", b["y_true"].data_ptr(), b["out"].data_ptr(), B * e.input_shape[0] * e.input_shape[1],
               wa, wp, 1, e.losses_dev.data_ptr(), b["g_out"].data_ptr(), None, 0)
        e._backward_body(B, segment=0)

    def _seg_bottleneck(self, B):
        self.eng._backward_body(B, segment=1, dense_dw=not self.gather_dense)

    def _seg_encoder(self, B):
        self.eng._backward_body(B, segment=2)

    def _seg_optimizer(self, B):
        # scale_regularization_loss(sum l2) (:232-233) contributes (2*0.001*W)/replicas per replica, i.e.
        # 2*0.001*W after the SUM all-reduce: added once here, after the reduce, on every replica alike.
        if self.loss == "dp":
            self.eng.l2_loss_and_grad(1.0)
        self.eng.adam_step()

    def _run(self, B):
        segs = (self._seg_forward_loss_decoder, self._seg_bottleneck, self._seg_encoder, self._seg_optimizer)
        if self.use_cuda_graph and self._graphs is None and self._calls == 1:
            torch.cuda.synchronize()
            graphs = []
            for s in segs:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    s(B)
                graphs.append(g)
            self._graphs = graphs
        works = []
        for i, s in enumerate(segs):
            if i == 3:
                if self.world > 1 and not self.overlap:
                    # one all-reduce of the whole flat gradient after backward: nothing runs beside the persistent
                    # one-CTA-per-SM conv kernels (an NCCL kernel resident on a few SMs forces their static tile
                    # schedule into a second wave)
                    works.append(dist.all_reduce(self.eng.G, op=dist.ReduceOp.SUM, async_op=True))
                for w in works:
                    w.wait()
            if self._graphs is not None:
                self._graphs[i].replay()
            else:
                s(B)
            if i < 3 and self.world > 1 and self.overlap:
                lo, hi = self.buckets[i]
                if not self.gather_dense:
                    works.append(dist.all_reduce(self.eng.G[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                elif i == 0:
                    works.append(dist.all_reduce(self.eng.G[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
                elif i == 1:
                    # bottleneck bucket = [embedding | Dense kernel, bias | projection]: gather the Dense operands,
                    # reduce the projection now; the embedding table rides with the (adjacent) encoder bucket
                    gathers = self._gather_dense_operands(B)
                    works.append(dist.all_reduce(self.eng.G[self._o_after_dense:hi], op=dist.ReduceOp.SUM, async_op=True))
                else:
                    # Dense kernel gradient of the global batch on the side stream, beside the encoder's backward
                    e = self.eng
                    with torch.cuda.stream(e.side):
                        for w in gathers:
                            w.wait()
                        e.dense_grad_from_gathered(*self._gathered)
                    works.append(dist.all_reduce(e.G[0:self._o_dense], op=dist.ReduceOp.SUM, async_op=True))
            if i == 2 and self.gather_dense:
                torch.cuda.current_stream().wait_stream(self.eng.side)
        self._calls += 1

    def _gather_dense_operands(self, B):
        x, dy = self.eng.dense_operands(B)
        if self._gathered is None or self._gathered[0].shape[0] != self.world * B:
            self._gathered = (torch.empty(self.world * B, x.shape[1], dtype=x.dtype, device=x.device),
                              torch.empty(self.world * B, dy.shape[1], dtype=dy.dtype, device=dy.device))
        return [dist.all_gather_into_tensor(self._gathered[0], x.view(B, -1), async_op=True),
                dist.all_gather_into_tensor(self._gathered[1], dy.view(B, -1), async_op=True)]

    def prefetch(self, spec_in, emb, spec_out):
        """Start the host -> device copy of the next shard on a copy stream (same argument order as train_step)."""
        self.eng.prefetcher.prefetch(spec_in, emb, spec_out)

    def train_step(self, spec_in, emb, spec_out):
        """inputs = this replica's shard (spec_in, emb, spec_out), the tuple order of the reference's
        train_step (main_training.py:254). Returns the replica's loss contribution as a 0-d tensor;
        `strategy.reduce(SUM)` of those (:326) is the global loss."""
        e = self.eng
        dev = e.device
        B = int(spec_in.shape[0])
        if not e.prefetcher.take(spec_in, emb, spec_out):
            spec_in = torch.as_tensor(spec_in).to(dev, torch.float32, non_blocking=True)
            spec_out = torch.as_tensor(spec_out).to(dev, torch.float32, non_blocking=True)
            emb = torch.as_tensor(emb).to(dev, torch.int32, non_blocking=True)
            e.stage(spec_in, emb, spec_out)
        self._run(B)
        loss = e.losses_dev[0].clone()
        if self.loss == "dp":
            loss = loss + e.reg_dev[0] / self.world      # reg_dev = sum(l2); each replica's share is 1/replicas
        return loss

    def set_epoch(self, epoch):
        self.eng.set_lr(lr_schedule(self.lr, epoch))


def main(n_epochs=2, steps_per_epoch=8, per_replica_batch=16, lr=5e-7, alpha=0.9, seed=500):
    """Synthetic-data stand-in for the reference's `__main__` (the dataset directory of
    main_training.py:75 does not exist here): same model, loss, optimiser and schedule."""
    from .dataset import Dataset
    from .datageneratorv2 import DataGenerator
    rank, world, local = init_distributed()
    model = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3,
                 name="U-Net")
    trainer = DistributedTrainer(model, per_replica_batch, alpha, lr, world=world)
    dataset = Dataset(None, "synthetic", n_synthetic=per_replica_batch * world * steps_per_epoch * 2, seed=seed)
    gen = DataGenerator(dataset, batch_size=per_replica_batch * world, partition="train", shuffle=True)
    for epoch in range(n_epochs):
        t0 = time.time()
        trainer.set_epoch(epoch)
        total, nb = 0.0, 0
        for i in range(min(len(gen), steps_per_epoch)):
            spec_in, emb, spec_out = shard_batch(gen[i], rank, world)
            loss = trainer.train_step(spec_in, emb, spec_out)
            if world > 1:
                dist.all_reduce(loss)
            total += float(loss); nb += 1
        gen.on_epoch_end()
        if rank == 0:
            print(f"Epoch {epoch + 1}, Loss: {total / max(nb, 1):.6f}, Epoch time: {time.time() - t0:.2f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
