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
import random
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
    """One replica of the synchronous data-parallel train step.

    The whole step -- forward, loss, the three backward segments, the three NCCL all-reduces of the gradient buckets
    (each issued on NCCL's stream as soon as its segment is enqueued, joined before the optimiser), the L2 regulariser
    and Adam -- is ONE CUDA graph per shard size (URIR_DP_GRAPH=segments restores round 1's four graphs with eagerly
    launched collectives, =off runs everything eagerly). World 1 runs the same body without the collectives."""

    def __init__(self, model: UNet, per_replica_batch=16, alpha=0.9, lr=5e-7, loss="dp", world=None,
                 use_cuda_graph=True, dropout=True, rank=None):
        self.model = model
        self.eng = model.model.engine
        self.world = world if world is not None else (dist.get_world_size() if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank() if dist.is_initialized() else 0)
        self.per_replica_batch = per_replica_batch
        self.global_batch = per_replica_batch * self.world
        self.alpha, self.lr, self.loss = alpha, lr, loss
        self.use_cuda_graph, self.dropout = use_cuda_graph, dropout
        self.buckets = gradient_buckets(self.eng.offsets, self.eng.n_flat)
        self.graph_mode = os.environ.get("URIR_DP_GRAPH", "one") if use_cuda_graph else "off"
        self._graphs = {}            # shard size B -> "warm" | CUDAGraph | [four segment graphs]
        self.overlap = os.environ.get("URIR_DP_OVERLAP", "1") != "0"
        # Dense layer (56 % of all parameters): all-gather its two small operands and form the global-batch kernel
        # gradient locally instead of all-reducing 47 MB of fp32 gradient (see engine.dense_grad_from_gathered)
        # (opt-in: measured equal to the plain all-reduce on NVSwitch at 2 and 8 GPUs, profiles/r01_layer_roofline.txt)
        self.gather_dense = self.world > 1 and self.overlap and os.environ.get("URIR_DP_GATHER_DENSE", "0") == "1"
        o = self.eng.offsets
        self._o_dense = o["vec.dense.w"][0]
        self._o_after_dense = o["vec.dense.b"][0] + o["vec.dense.b"][1]
        self._gathered = None
        # A bucket's all-reduce must see the weight / Dense gradients the side stream produces, but the MAIN stream need not:
        # the collective is issued from an auxiliary stream that waits for both, and the next segment's dgrad chain keeps
        # running beside those gradient kernels as in the single-GPU step (URIR_DP_JOIN=1 restores round 1's joins, which
        # serialised e.g. the Dense gradient in front of the whole encoder backward: +0.25 ms per step at world 1).
        self._segment_join = (os.environ.get("URIR_DP_JOIN", "0") == "1" or self.graph_mode == "segments" or self.gather_dense
                              or not self.overlap)
        self._aux = torch.cuda.Stream(device=self.eng.device)
        # MirroredStrategy gives every replica its own Dropout mask: mix the rank into the mask stream's seed
        self.eng.dropout_seed = (self.eng.dropout_seed + 0x9E3779B1 * self.rank) & 0x7FFFFFFFFFFFFFFF
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
               wa, wp, int(e.head_sigmoid), e.losses_dev.data_ptr(), b["g_out"].data_ptr(), None, 0)
        e._backward_body(B, segment=0, join=self._segment_join)

    def _seg_bottleneck(self, B):
        self.eng._backward_body(B, segment=1, dense_dw=not self.gather_dense, join=self._segment_join)

    def _seg_encoder(self, B):
        self.eng._backward_body(B, segment=2)

    def _seg_optimizer(self, B):
        # scale_regularization_loss(sum l2) (:232-233) contributes (2*0.001*W)/replicas per replica, i.e.
        # 2*0.001*W after the SUM all-reduce: added once here, after the reduce, on every replica alike.
        if self.loss == "dp":
            self.eng.l2_loss_and_grad(1.0)
        self.eng.adam_step()

    def _segments(self):
        return (self._seg_forward_loss_decoder, self._seg_bottleneck, self._seg_encoder, self._seg_optimizer)

    def _reduce_bucket(self, i, B, works, state):
        """All-reduce of gradient bucket i (async, on NCCL's stream) right after backward segment i was enqueued."""
        e = self.eng
        lo, hi = self.buckets[i]
        if not self._segment_join and not self.gather_dense:
            main = torch.cuda.current_stream()
            self._aux.wait_stream(main)
            self._aux.wait_stream(e.side)
            with torch.cuda.stream(self._aux):
                works.append(dist.all_reduce(e.G[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            state["aux"] = True
            return
        if not self.gather_dense or i == 0:
            works.append(dist.all_reduce(e.G[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        elif i == 1:
            # bottleneck bucket = [embedding | Dense kernel, bias | projection]: gather the Dense operands,
            # reduce the projection now; the embedding table rides with the (adjacent) encoder bucket
            state["gathers"] = self._gather_dense_operands(B)
            works.append(dist.all_reduce(e.G[self._o_after_dense:hi], op=dist.ReduceOp.SUM, async_op=True))
        else:
            # Dense kernel gradient of the global batch on the side stream, beside the encoder's backward
            e.side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(e.side):
                for w in state["gathers"]:
                    w.wait()
                e.dense_grad_from_gathered(*self._gathered)
            works.append(dist.all_reduce(e.G[0:self._o_dense], op=dist.ReduceOp.SUM, async_op=True))
            torch.cuda.current_stream().wait_stream(e.side)

    def _step_body(self, B, graphs=None):
        """The whole step in stream order; `graphs` = four captured segment graphs to replay instead of issuing."""
        works, state = [], {}
        for i, seg in enumerate(self._segments()):
            if i == 3 and graphs is None and self._bucketwise_optimizer(works, state):
                break
            if i == 3:
                if self.world > 1 and not self.overlap:
                    # one all-reduce of the whole flat gradient after backward (nothing runs beside the conv kernels)
                    works.append(dist.all_reduce(self.eng.G, op=dist.ReduceOp.SUM, async_op=True))
                for w in works:
                    w.wait()
                if state.get("aux"):
                    torch.cuda.current_stream().wait_stream(self._aux)
                self.eng._join_side()
            if graphs is not None:
                graphs[i].replay()
            else:
                seg(B)
            if i < 3 and self.world > 1 and self.overlap:
                self._reduce_bucket(i, B, works, state)

    def _bucketwise_optimizer(self, works, state):
        """Adam bucket by bucket, each as soon as ITS all-reduce has finished: the decoder's and the vector block's updates
        (95 % of the parameters, ~75 of Adam's ~90 us) run while the encoder bucket -- issued last, with nothing left to hide
        behind -- is still on the wire. The regulariser's gradient is added per part right before the bucket that holds those
        kernels. Same arithmetic as the one-launch optimiser (URIR_DP_ADAM_BUCKETS=0 restores it)."""
        if not (self.world > 1 and self.overlap and len(works) == 3 and state.get("aux") and not self.gather_dense
                and os.environ.get("URIR_DP_ADAM_BUCKETS", "1") != "0"):
            return False
        e = self.eng
        up4 = lambda v: (v + 3) // 4 * 4
        (d_lo, d_hi), (v_lo, _), _ = self.buckets            # decoder + head | vector block | encoder, in reduce order
        ranges = [(up4(d_lo), d_hi), (up4(v_lo), up4(d_lo)), (0, up4(v_lo))]
        for k, (lo, hi) in enumerate(ranges):
            works[k].wait()
            if k == 2:
                e._join_side()
            if self.loss == "dp" and k != 1:
                e.l2_loss_and_grad(1.0, part="dec" if k == 0 else "enc")
            e.adam_range(lo, hi)
        torch.cuda.current_stream().wait_stream(self._aux)
        e.adam_finish()
        return True

    def _run(self, B):
        st = self._graphs.get(B)
        if self.graph_mode == "off":
            self._step_body(B)
        elif st is None:                         # first call for this shard size: eager (kernel attributes, NCCL warm-up)
            self._step_body(B)
            self._graphs[B] = "warm"
        elif st == "warm":                       # second call: capture, then run
            torch.cuda.synchronize()
            if self.graph_mode == "segments":
                graphs = []
                for seg in self._segments():
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        seg(B)
                    graphs.append(g)
                self._graphs[B] = graphs
                self._step_body(B, graphs)
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._step_body(B)
                self._graphs[B] = g
                g.replay()
        elif isinstance(st, list):
            self._step_body(B, st)
        else:
            st.replay()

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
            loss = loss + (e.reg_dev[0] + e.reg_dev[1]) / self.world      # reg_dev = sum(l2) (in two parts when the optimiser ran bucket by bucket); each replica's share is 1/replicas
        return loss

    def test_step(self, spec_in, emb, spec_out):
        """The reference's test_step (main_training.py:293-316): a forward with training=True (its literal argument:
        BatchNormalization uses -- and updates -- batch statistics, Dropout is active), then the two validation metrics
        of the shard: mean squared amplitude error and mean 1 - cos of the wrapped phase difference.
        Returns (loss_amplitude, loss_phase) as 0-d device tensors."""
        e = self.eng
        dev = e.device
        B = int(spec_in.shape[0])
        spec_in = torch.as_tensor(spec_in).to(dev, torch.float32, non_blocking=True)
        emb = torch.as_tensor(emb).to(dev, torch.int32, non_blocking=True)
        y = torch.as_tensor(spec_out).to(dev, torch.float32, non_blocking=True).contiguous()
        e.sync_operands()
        out = e.forward(spec_in, emb, training=True, dropout=self.dropout)
        n = B * e.input_shape[0] * e.input_shape[1]
        res = torch.empty(4, dtype=torch.float32, device=dev)
        L.call("ampphase_loss", y.data_ptr(), out.data_ptr(), n, 1.0 / n, 1.0 / n, 0, res.data_ptr(), None, None, 0)
        return res[2], res[1]

    def set_epoch(self, epoch, start=80):
        self.eng.set_lr(lr_schedule(self.lr, epoch, start))

    # ---- tf.train.Checkpoint(optimizer=..., model=...) contents (main_training.py:171)
    def checkpoint_state(self):
        e = self.eng
        return {"model": e.state_dict(), "adam_m": e.M.detach().cpu().clone(), "adam_v": e.V.detach().cpu().clone(),
                "step": int(e.step_dev.item()), "lr": float(e.lr_dev.item())}

    def load_checkpoint_state(self, st):
        e = self.eng
        e.load_state_dict(st["model"])
        with torch.no_grad():
            e.M.copy_(st["adam_m"]); e.V.copy_(st["adam_v"])
        e.step_dev.fill_(int(st["step"]))
        e.set_lr(st["lr"])


class CheckpointManager:
    """tf.train.CheckpointManager(checkpoint, directory, max_to_keep=2) (main_training.py:171-172): numbered saves
    `ckpt-<n>.pt`, the oldest deleted beyond max_to_keep; `latest_checkpoint` / `restore_latest` for a restart."""

    def __init__(self, trainer: DistributedTrainer, directory, max_to_keep=2):
        self.trainer, self.directory, self.max_to_keep = trainer, directory, max_to_keep
        os.makedirs(directory, exist_ok=True)
        self._kept = sorted((f for f in os.listdir(directory) if f.startswith("ckpt-") and f.endswith(".pt")),
                            key=lambda f: int(f[5:-3]))
        self._n = int(self._kept[-1][5:-3]) if self._kept else 0

    @property
    def latest_checkpoint(self):
        return os.path.join(self.directory, self._kept[-1]) if self._kept else None

    def save(self):
        self._n += 1
        name = f"ckpt-{self._n}.pt"
        torch.save(self.trainer.checkpoint_state(), os.path.join(self.directory, name))
        self._kept.append(name)
        while len(self._kept) > self.max_to_keep:
            os.remove(os.path.join(self.directory, self._kept.pop(0)))
        return os.path.join(self.directory, name)

    def restore_latest(self):
        path = self.latest_checkpoint
        if path is not None:
            self.trainer.load_checkpoint_state(torch.load(path, map_location="cpu"))
        return path


def _global_mean(values, world):
    """tf.keras.metrics.Mean over every replica's updates (ON_READ sum aggregation, main_training.py:237-242)."""
    if not values:
        return float("nan")
    t = torch.stack([v.reshape(()).float() for v in values]).mean()
    if world > 1:
        dist.all_reduce(t)
        t = t / world
    return float(t)


def train_loop(trainer: DistributedTrainer, train_generator, val_generator, n_epochs, manager=None,
               lr_exp_decay=(True, 80), rank=0, world=1, max_steps=None, verbose=True):
    """The epoch loop of main_training.py:332-391: LR decay from epoch lr_exp_decay[1], one pass over the training
    generator (each rank takes its contiguous slice of every global batch, :114), one pass over the validation
    generator through test_step (training=True, as written there), a checkpoint every second epoch (`epoch % 2 == 0`),
    the reference's four console lines. Returns one dict per epoch."""
    history = []
    t_start = time.time()
    for epoch in range(n_epochs):
        t0 = time.time()
        if lr_exp_decay[0]:
            trainer.set_epoch(epoch, lr_exp_decay[1])
        n_train = len(train_generator) if max_steps is None else min(len(train_generator), max_steps)
        n_val = len(val_generator) if max_steps is None else min(len(val_generator), max_steps)
        losses, amp, ph = [], [], []
        for i in range(n_train):
            spec_in, emb, spec_out = shard_batch(train_generator[i][:3], rank, world)
            losses.append(trainer.train_step(spec_in, emb, spec_out))
            ld = trainer.eng.losses_dev.clone()
            amp.append(ld[2]); ph.append(ld[1])
        total = torch.stack(losses).sum() if losses else torch.zeros((), device=trainer.eng.device)
        if world > 1:
            dist.all_reduce(total)                      # strategy.reduce(SUM, per_replica_losses) (:326-327)
        train_loss = float(total) / max(n_train, 1)
        v_amp, v_ph = [], []
        for i in range(n_val):
            spec_in, emb, spec_out = shard_batch(val_generator[i][:3], rank, world)
            a, p = trainer.test_step(spec_in, emb, spec_out)
            v_amp.append(a); v_ph.append(p)
        saved = None
        if manager is not None and epoch % 2 == 0 and rank == 0:
            saved = manager.save()
        row = {"epoch": epoch + 1, "loss": train_loss, "train_mse": _global_mean(amp, world),
               "train_phase": _global_mean(ph, world), "val_mse": _global_mean(v_amp, world),
               "val_phase": _global_mean(v_ph, world), "lr": float(trainer.eng.lr_dev.item()), "checkpoint": saved,
               "seconds": time.time() - t0}
        history.append(row)
        random.seed(0x5EED + epoch)      # every rank must reshuffle the global batches identically (on_epoch_end uses `random`)
        for g in (train_generator, val_generator):
            if hasattr(g, "on_epoch_end"):
                g.on_epoch_end()
        if verbose and rank == 0:
            print("Epoch {}, Loss: {}, Epoch time: {}\nTrain | MSE Loss: {}, Phase Loss: {}\n"
                  "Val   | MSE Loss: {}, Phase Loss: {}\nlr    | {}".format(
                      row["epoch"], row["loss"], row["seconds"], row["train_mse"], row["train_phase"], row["val_mse"],
                      row["val_phase"], row["lr"]))
    if verbose and rank == 0:
        print("Training complete, took " + str(time.time() - t_start))
    return history


def main(n_epochs=2, steps_per_epoch=8, per_replica_batch=16, lr=5e-7, alpha=0.9, seed=500, file_name=None,
         lr_exp_decay=(True, 80)):
    """Synthetic-data stand-in for the reference's `__main__` (the dataset directory of main_training.py:75 does not
    exist here): same model, loss, optimiser, schedule, validation pass and checkpoint cadence."""
    from .dataset import Dataset
    from .datageneratorv2 import DataGenerator
    rank, world, local = init_distributed()
    model = UNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, kernels=3,
                 name="U-Net")
    trainer = DistributedTrainer(model, per_replica_batch, alpha, lr, world=world, rank=rank)
    gb = per_replica_batch * world
    dataset = Dataset(None, "synthetic", n_synthetic=int(gb * steps_per_epoch / 0.7) + 4 * gb, seed=seed)
    train_generator = DataGenerator(dataset, batch_size=gb, partition="train", shuffle=True)
    val_generator = DataGenerator(dataset, batch_size=gb, partition="val", shuffle=True)
    manager = CheckpointManager(trainer, file_name, max_to_keep=2) if file_name else None
    history = train_loop(trainer, train_generator, val_generator, n_epochs, manager, lr_exp_decay, rank, world,
                         max_steps=steps_per_epoch)
    if world > 1:
        # graphs that captured NCCL kernels must go before the communicator does (teardown behind them can hang)
        trainer._graphs.clear()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
    return history


if __name__ == "__main__":
    main()
