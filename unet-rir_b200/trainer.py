"""The reference's GENERIC trainer (trainer.py:13-205) on the sm_100a engine: same constructor, epoch loop, callbacks and
step as amp_phase_trainer.Trainer, with the three things that file does differently:

  * loss (:146-156): `loss = amplitude_loss(y_true, y_pred)` = tf.keras.losses.mean_squared_error over the last axis, i.e.
    squared error over BOTH channels -- and NOT reduced further, so `tape.gradient` differentiates the SUM of the
    (B, H, W) per-pixel values: d/dy_pred = (y_pred - y_true) per element. step() returns the MEAN of that tensor as
    `loss` (what the epoch loop's np.mean makes of it, :93-96), `loss_phase` = mean(1 - cos) and `loss_stft` = mean squared
    amplitude error, like the reference's model_loss;
  * optimiser (:30-38): the same substring matching plus `'lamb'` -> tensorflow_addons LAMB (engine.lamb_step);
  * ModelCheckpoint(min_delta) (:175-205): "improved" (what EarlyStopping sees) requires val_loss + min_delta <
    val_loss_min, while saving and the running minimum follow the plain comparison.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import amp_phase_trainer as _A
from .amp_phase_trainer import EarlyStopping, History, params_saver, plot_graphs, rmse_coef  # noqa: F401  (same names)


class Trainer(_A.Trainer):

    def __init__(self, alpha, n_epochs, optimizer, callbacks, lr_exp_decay, lr0, file_name):
        lamb = ('nadam' not in optimizer) and ('sgd' not in optimizer) and ('adam' not in optimizer) and ('lamb' in optimizer)
        super().__init__(alpha, n_epochs, 'adam' if lamb else optimizer, callbacks, lr_exp_decay, lr0, file_name)
        if lamb:
            self.optimizer = 'lamb'

    def _device_step(self, eng, B):
        H, W, _ = eng.input_shape
        n = B * H * W
        b = eng._buffers(B)
        eng._forward_body(B, training=True, dropout=self.dropout)
        # gradient of sum over pixels of mean over the two channels of the squared error: 0.5 * SSE
        L.call("mse2_loss", b["y_true"].data_ptr(), b["out"].data_ptr(), n, 0.5, int(eng.head_sigmoid), eng.losses_dev.data_ptr(), b["g_out"].data_ptr())
        eng._backward_body(B)
        if self.optimizer == 'adam':
            eng.adam_step()
        elif self.optimizer == 'sgd':
            eng.sgd_step()
        elif self.optimizer == 'lamb':
            eng.lamb_step()
        else:
            eng.nadam_step()

    def step(self, spec_in, spec_out, emb, model):
        super().step(spec_in, spec_out, emb, model)
        losses = model.model.engine.losses_dev.clone()
        return losses[3], losses[1], losses[2]

    def model_loss(self, y_true, y_pred):
        """(loss, loss_phase, loss_stft) = (MSE over both channels, mean(1 - cos), mean squared amplitude error)."""
        dev = y_pred.device if isinstance(y_pred, torch.Tensor) and y_pred.is_cuda else torch.device("cuda")
        yt = _A._dev_tensor(y_true, torch.float32, dev).contiguous()
        yp = _A._dev_tensor(y_pred, torch.float32, dev).contiguous()
        n = yt.numel() // 2
        out = torch.empty(4, dtype=torch.float32, device=dev)
        L.call("mse2_loss", yt.data_ptr(), yp.data_ptr(), n, 0.5, 0, out.data_ptr(), None)
        return out[3], out[1], out[2]


def amplitude_loss(y_true, y_pred):
    """tf.keras.losses.mean_squared_error: mean over the LAST axis only (trainer.py:159-162)."""
    yt, yp = torch.as_tensor(y_true, dtype=torch.float32), torch.as_tensor(y_pred, dtype=torch.float32)
    return ((yt - yp.to(yt.device)) ** 2).mean(dim=-1)


def phase_loss(y_true, y_pred):
    """K.mean(1 - cos(2 pi (y_true - y_pred))) (trainer.py:164-168): a scalar."""
    yt, yp = torch.as_tensor(y_true, dtype=torch.float32), torch.as_tensor(y_pred, dtype=torch.float32)
    return (1 - torch.cos((yt - yp.to(yt.device)) * 2 * torch.pi)).mean()


class ModelCheckpoint(object):
    def __init__(self, filepath, save_best_only, verbose, min_delta=0.0001):
        self.filepath, self.save_best_only, self.verbose, self.min_delta = filepath, save_best_only, verbose, min_delta
        self.train_loss_min = self.val_loss_min = 10

    def checkpoint(self, train_loss, val_loss, model):
        improve = val_loss + self.min_delta < self.val_loss_min
        if val_loss < self.val_loss_min:
            if self.verbose:
                print('Validation loss improved from ' + str(self.val_loss_min) + ' to ' + str(val_loss))
            if self.save_best_only:
                model.save(self.filepath)
            self.val_loss_min, self.train_loss_min = val_loss, train_loss
        elif self.verbose:
            print('Validation loss did not improve')
        return improve
