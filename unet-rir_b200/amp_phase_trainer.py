"""`Trainer` and friends with the reference's names and semantics (amp_phase_trainer.py:12-306),
running the step on the sm_100a engine.

Trainer.step(spec_in, spec_out, emb, model) does what the reference does under its GradientTape
(amp_phase_trainer.py:130-141): forward with training=True, amp-MSE + phase-(1-cos) loss, gradients
for all 77 trainable variables, optimiser update -- as one fused device sequence
(forward -> fused loss fwd+bwd -> backward -> Adam -> bf16 operand refresh), captured in a CUDA graph
after the first call for a given batch size. It returns (loss, loss_phase, loss_stft) as 0-d device
tensors so the loop does not synchronise per step (the reference syncs once per epoch, :94-96).

Documented differences from the reference, which is not runnable as shipped:
  * generators may yield either (spec_in, spec_out, emb) via __next__ (what Trainer.train unpacks,
    :65) or be indexable DataGenerators returning (spec_in, emb, spec_out); both are accepted.
  * matplotlib is imported lazily inside plot_graphs (it is not installed here).
"""
from __future__ import annotations

import json
import math
import sys
import time

import numpy as np
import torch

from . import _lib as L


def _dev_tensor(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(device, non_blocking=True)


def _mean(values):
    """np.mean over a list of python floats or 0-d device tensors (one sync)."""
    if len(values) == 0:
        return float("nan")
    if isinstance(values[0], torch.Tensor):
        return float(torch.stack([v.reshape(()) for v in values]).float().mean().item())
    return float(np.mean(values))


class Trainer:

    def __init__(self, alpha, n_epochs, optimizer, callbacks, lr_exp_decay, lr0, file_name):

        'Initialization'
        self.alpha = alpha            # stored and unused, like the reference (:17, loss at :155)
        self.n_epochs = n_epochs
        self.lr0 = lr0
        self.model_checkpoint = callbacks[0]
        self.early_stop = callbacks[1]
        self.lr_exp_decay = lr_exp_decay[0]
        self.lr_exp_decay_epoch = lr_exp_decay[1]
        self.file_name = file_name

        'Secondary initialization'
        self.train_loss_history = np.empty((n_epochs, 3), dtype=np.float32)
        self.val_loss_history = np.empty((n_epochs, 3), dtype=np.float32)

        # substring match in the reference's order: "nadam" contains "adam" (:30-35)
        if 'nadam' in optimizer:
            self.optimizer = 'nadam'
        elif 'sgd' in optimizer:
            self.optimizer = 'sgd'
        elif 'adam' in optimizer:
            self.optimizer = 'adam'
        else:
            raise ValueError(f"optimizer must contain 'nadam', 'sgd' or 'adam', got {optimizer!r}")
        self.learning_rate = lr0
        self._graphs = {}
        self.use_cuda_graph = True
        self.dropout = True

    # ------------------------------------------------------------------ epoch loop (:37-127)
    @staticmethod
    def _next_batch(gen, i):
        if hasattr(gen, "__next__"):
            return gen.__next__()                       # (spec_in, spec_out, emb), as :65 unpacks
        spec_in, emb, spec_out = gen.__getitem__(i)[:3]  # DataGenerator order (datageneratorv2.py:101)
        return spec_in, spec_out, emb

    def train(self, model, train_generator, val_generator):
        """Epoch loop of the reference (:37-127): same console lines, same history rows [loss, phase, stft], same
        checkpoint / early-stopping protocol; returns (model, History)."""
        print("[INFO]: Training model...")
        n_train, n_val = len(train_generator), len(val_generator)
        last = -1
        for epoch in range(self.n_epochs):
            last = epoch
            print("\n[INFO]: Starting epoch {}/{}...".format(epoch + 1, self.n_epochs), end="\n")
            sys.stdout.flush()
            t_start = time.time()
            self._apply_lr_schedule(epoch)
            train_sums = self._train_epoch(model, train_generator, n_train)
            val_sums = self._validate(model, val_generator, n_val)
            print("took {:.4} seconds".format(time.time() - t_start))

            train_means = [_mean(col) for col in train_sums]
            self._report("Perdidas training:", " - Perdidas: ", train_means)
            val_means = [_mean(col) for col in val_sums]
            self._report("Perdidas validation:", " - Perdidas combinadas: ", val_means)
            self.train_loss_history[epoch, :] = train_means
            self.val_loss_history[epoch, :] = val_means

            improved = self.model_checkpoint.checkpoint(train_loss=train_means[0], val_loss=val_means[0], model=model)
            if self.early_stop.stop_count(improve=improved):
                break
        done = last + 1
        return model, History(done, self.train_loss_history[:done, :], self.val_loss_history[:done, :])

    def _apply_lr_schedule(self, epoch):
        # exponential decay over the last epochs (:56-59)
        if self.lr_exp_decay and epoch >= self.lr_exp_decay_epoch:
            self.learning_rate = self.lr0 * np.exp(-0.25 * (epoch - self.lr_exp_decay_epoch))

    def _train_epoch(self, model, gen, n_batches):
        """One pass over the training generator; returns three lists (loss, phase, stft) of per-step device scalars."""
        cols = ([], [], [])
        batch = self._next_batch(gen, 0) if n_batches > 0 else None
        for i in range(n_batches):
            spec_in, spec_out, emb = batch
            step_losses = self.step(spec_in, spec_out, emb, model)
            if i + 1 < n_batches:             # next batch's H2D copy overlaps this step (enqueued, not waited for)
                batch = self._next_batch(gen, i + 1)
                self.prefetch(batch[0], batch[1], batch[2], model)
            for col, v in zip(cols, step_losses):
                col.append(v)
        return cols

    def _validate(self, model, gen, n_batches):
        cols = ([], [], [])
        for i in range(n_batches):
            spec_in, spec_out, emb = self._next_batch(gen, i)
            with torch.no_grad():
                generated = model.model([spec_in, emb], training=False)
            for col, v in zip(cols, self.model_loss(spec_out, generated)):
                col.append(v)
        return cols

    @staticmethod
    def _report(title, first_label, means):
        print(title)
        for label, v in zip((first_label, " - Perdidas fase: ", " - Perdidas módulo: "), means):
            print(label + str(v))

    # ------------------------------------------------------------------ one optimiser step (:130-141)
    def _device_step(self, eng, B):
        """forward -> loss -> backward -> optimiser on the engine's static buffers."""
        H, W, _ = eng.input_shape
        n = B * H * W
        b = eng._buffers(B)
        eng._forward_body(B, training=True, dropout=self.dropout)
        L.call("ampphase_loss", b["y_true"].data_ptr(), b["out"].data_ptr(), n, 1.0 / n, 1.0 / n, int(eng.head_sigmoid),
               eng.losses_dev.data_ptr(), b["g_out"].data_ptr(), None, 0)
        eng._backward_body(B)
        if self.optimizer == 'adam':
            eng.adam_step()
        elif self.optimizer == 'sgd':
            eng.sgd_step()
        else:
            eng.nadam_step()

    def prefetch(self, spec_in, spec_out, emb, model):
        """Optional: start the host -> device copy of the NEXT batch (same argument order as step) on a copy
        stream while the current step computes; the following step(...) called with the same objects uses it.
        train() does this for every batch."""
        model.model.engine.prefetcher.prefetch(spec_in, emb, spec_out)

    def step(self, spec_in, spec_out, emb, model):
        eng = model.model.engine
        dev = eng.device
        B = int(spec_in.shape[0])
        eng.set_lr(self.learning_rate)
        if not eng.prefetcher.take(spec_in, emb, spec_out):      # not announced with prefetch(): copy now
            spec_in = _dev_tensor(spec_in, torch.float32, dev)
            spec_out = _dev_tensor(spec_out, torch.float32, dev)
            emb = _dev_tensor(emb, torch.int32, dev)
            eng.stage(spec_in, emb, spec_out)
        eng.sync_operands()              # the masters may have been changed through torch since the last refresh
        if not self.use_cuda_graph:
            self._device_step(eng, B)
        else:
            key = (id(eng), B, self.optimizer, self.dropout)
            state = self._graphs.get(key)
            if state is None:                      # first call: eager (also sets kernel attributes)
                self._device_step(eng, B)
                self._graphs[key] = "warm"
            elif state == "warm":                  # second call: capture, then replay
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._device_step(eng, B)
                self._graphs[key] = g
                g.replay()
            else:
                state.replay()
        losses = eng.losses_dev.clone()
        return losses[0], losses[1], losses[2]

    # ------------------------------------------------------------------ losses (:143-168)
    def model_loss(self, y_true, y_pred):
        """(loss, loss_phase, loss_stft): loss_stft = mean((a_t-a_p)^2), loss_phase =
        mean(1-cos(2*pi*(p_t-p_p))), loss = loss_phase + loss_stft."""
        dev = y_pred.device if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda else torch.device("cuda")
        yt = _dev_tensor(y_true, torch.float32, dev).contiguous()
        yp = _dev_tensor(y_pred, torch.float32, dev).contiguous()
        n = yt.numel() // 2
        out = torch.empty(4, dtype=torch.float32, device=dev)
        L.call("ampphase_loss", yt.data_ptr(), yp.data_ptr(), n, 1.0 / n, 1.0 / n, 0, out.data_ptr(), None, None, 0)
        return out[0], out[1], out[2]

    def amplitude_loss(self, y_true, y_pred):
        yt = _dev_tensor(y_true, torch.float32, "cuda")
        yp = _dev_tensor(y_pred, torch.float32, "cuda")
        pair_t = torch.stack([yt, torch.zeros_like(yt)], dim=-1).contiguous()
        pair_p = torch.stack([yp, torch.zeros_like(yp)], dim=-1).contiguous()
        return self.model_loss(pair_t, pair_p)[2]

    def phase_loss(self, y_true, y_pred):
        yt = _dev_tensor(y_true, torch.float32, "cuda")
        yp = _dev_tensor(y_pred, torch.float32, "cuda")
        pair_t = torch.stack([torch.zeros_like(yt), yt], dim=-1).contiguous()
        pair_p = torch.stack([torch.zeros_like(yp), yp], dim=-1).contiguous()
        return self.model_loss(pair_t, pair_p)[1]


########################################################
# Callbacks
########################################################

class ModelCheckpoint(object):
    """Best-validation bookkeeping of the reference (:175-203): both minima start at 10, a strictly smaller validation
    loss counts as an improvement (and saves the model when save_best_only is set)."""

    def __init__(self, filepath, save_best_only, verbose):
        self.filepath, self.save_best_only, self.verbose = filepath, save_best_only, verbose
        self.train_loss_min = self.val_loss_min = 10

    def checkpoint(self, train_loss, val_loss, model):
        if not val_loss < self.val_loss_min:
            if self.verbose:
                print('Validation loss did not improve')
            return False
        if self.verbose:
            print('Validation loss improved from ' + str(self.val_loss_min) + ' to ' + str(val_loss))
        if self.save_best_only:
            model.save(self.filepath)
        self.val_loss_min, self.train_loss_min = val_loss, train_loss
        return True


class EarlyStopping(object):
    """Stops when `patience` consecutive epochs brought no improvement (:206-223); the counter resets on improvement and
    the comparison is `==`, so a patience of 0 never fires."""

    def __init__(self, patience):
        self.patience = patience
        self.count = 0

    def stop_count(self, improve):
        self.count = 0 if improve else self.count + 1
        return self.count == self.patience


class History(object):

    def __init__(self, epochs, train_loss_history, val_loss_history, train_acc_history=None, val_acc_history=None):
        self.epochs = epochs
        self.train_loss_history = train_loss_history
        self.train_acc_history = train_acc_history
        self.val_loss_history = val_loss_history
        self.val_acc_history = val_acc_history


def plot_graphs(x1=None, y1=None, x2=None, y2=None, x3=None, y3=None, x4=None, y4=None, label1='', label2='',
                label3='', label4='', filename='./Graphic.png'):
    """Up to four labelled loss curves over epochs into one PNG (the helper of amp_phase_trainer.py:246-275; same
    signature). A missing x axis becomes 0..len(y)-1. matplotlib is imported on use: it is optional here."""
    import matplotlib.pyplot as plt
    curves = [(x, y, lab) for x, y, lab in ((x1, y1, label1), (x2, y2, label2), (x3, y3, label3), (x4, y4, label4))
              if y is not None]
    with plt.style.context("ggplot"):
        fig, ax = plt.subplots()
        for x, y, lab in curves:
            ax.plot(np.arange(len(y)) if x is None else x, y, label=lab)
        ax.set(title="Graphic", xlabel="Epoch ", ylabel="Loss")
        ax.legend()
        fig.savefig(filename)
        plt.close(fig)


def params_saver(file_name, batch_size, optimizer, criterion, lr, BatchNorm, normalization, epochs,
                 callbacks, alpha, beta, number_filters_0):
    """Writes <file_name>/hiperparametros.json with the run's hyper-parameters and the best losses the checkpoint
    callback saw (amp_phase_trainer.py:278-297: same file name, same keys)."""
    checkpoint, early_stop = callbacks[0], callbacks[1]
    record = dict(batch_size=batch_size, optimizer=str(optimizer), criterion=str(criterion), epochs=epochs, lr=lr,
                  alpha=alpha, beta=beta, batch_norm=BatchNorm, normalization=normalization,
                  number_filters_0=number_filters_0, val_loss=float(checkpoint.val_loss_min),
                  train_loss=float(checkpoint.train_loss_min), patience=early_stop.patience)
    with open(file_name + '/hiperparametros.json', 'w') as fp:
        json.dump(record, fp)


def rmse_coef(y_true, y_pred):
    yt = torch.as_tensor(y_true, dtype=torch.float32).flatten()
    yp = torch.as_tensor(y_pred, dtype=torch.float32).flatten().to(yt.device)
    return torch.sqrt(torch.mean((yt - yp) ** 2) + 1.0e-12)


def fit_mse(unet, x_train1, x_train2, y_train, x_val1, x_val2, y_val, batch_size, num_epochs, steps_per_epoch,
            learning_rate, callbacks):
    """UNet.compile_and_fit (u_net.py:83-118): Adam with InverseTimeDecay(lr, steps_per_epoch*100, rate 1),
    MeanSquaredError over both channels, shuffle=False, EarlyStopping(val_loss, patience 20)."""
    eng = unet.model.engine
    hist = {"loss": [], "val_loss": []}
    early = callbacks[1] if len(callbacks) > 1 else EarlyStopping(20)
    it, best = 0, float("inf")
    n_train = len(x_train1)
    for epoch in range(num_epochs):
        tl = []
        for i in range(0, n_train, batch_size):
            xb = _dev_tensor(x_train1[i:i + batch_size], torch.float32, eng.device)
            eb = _dev_tensor(x_train2[i:i + batch_size], torch.int32, eng.device)
            yb = _dev_tensor(y_train[i:i + batch_size], torch.float32, eng.device)
            eng.set_lr(learning_rate / (1.0 + it / (steps_per_epoch * 100.0)))
            out = eng.forward(xb, eb, training=True)
            diff = out - yb
            n = diff.numel()
            eng.backward((2.0 / n) * diff * out * (1.0 - out) if eng.head_sigmoid else (2.0 / n) * diff)
            eng.adam_step()
            tl.append((diff * diff).mean())
            it += 1
        vl = []
        for i in range(0, len(x_val1), batch_size):
            xb = _dev_tensor(x_val1[i:i + batch_size], torch.float32, eng.device)
            eb = _dev_tensor(x_val2[i:i + batch_size], torch.int32, eng.device)
            yb = _dev_tensor(y_val[i:i + batch_size], torch.float32, eng.device)
            out = eng.forward(xb, eb, training=False)
            vl.append(((out - yb) ** 2).mean())
        hist["loss"].append(_mean(tl)); hist["val_loss"].append(_mean(vl))
        improve = hist["val_loss"][-1] < best
        best = min(best, hist["val_loss"][-1])
        if early.stop_count(improve):
            break
    return hist
