"""`UNet` with the reference's constructor, attributes and methods (dl_models/u_net.py:34-199),
backed by the sm_100a engine instead of a Keras graph.

`UNet(...).model` is the callable the reference's trainers use:
    model.model([spec_in, emb], training=bool) -> (B, H, W, 2) float32 in (0, 1)
    model.model.trainable_variables            -> 77 tensors in Keras order
    model.model.losses                         -> the l2(0.001) kernel-regulariser terms
Inputs may be numpy arrays or torch tensors, NHWC float32 spectrograms and (B, 2, 16) int32
embedding ids, exactly the DataGenerator batch contract (datageneratorv2.py:88-102). Outputs are torch
tensors on the GPU (use `.cpu().numpy()` where the reference used `.numpy()`).

Deviations, all deliberate and visible:
  * weights are saved as a torch state dict (`weights.pt`); h5py / Keras `weights.h5` do not exist here. Weights
    trained with the reference come in through `load_weights("x.npz")` (tools/export_tf_weights.py, keras_weights.py).
  * `_save_parameters` also stores `kernels`. The reference omits it (u_net.py:180-187), so its own
    `UNet.load` rebuilds with `BatchNorm` shifted into the `kernels` slot -- a latent bug we do not copy.
  * all four block modes (0 convolutional_block_1 ... 3 residual_block_2) are wired on the device with
    `BatchNorm=True` or `False`; every reference call site uses mode 0 with BatchNorm.
"""
from __future__ import annotations

import os
import pickle

import numpy as np
import torch

from .. import _lib as L
from .. import plan as PL
from ..engine import UNetEngine


def _to_dev(x, dtype, device):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(x), dtype=dtype).to(device, non_blocking=True)


class _UNetFn(torch.autograd.Function):
    """Lets `loss.backward()` drive the engine's backward pass (the GradientTape of
    amp_phase_trainer.py:133-138) when a caller builds its own loss in torch."""

    @staticmethod
    def forward(ctx, model, spec, emb, dropout_mask, *params):
        out = model.engine.forward(spec, emb, training=True, dropout_mask=dropout_mask).clone()
        ctx.model = model
        ctx.save_for_backward(out)
        return out

    @staticmethod
    def backward(ctx, gout):
        (out,) = ctx.saved_tensors
        eng = ctx.model.engine
        # through the fused sigmoid head (the sibling DiffUNet's head is linear)
        eng.backward((gout * out * (1.0 - out)).contiguous() if eng.head_sigmoid else gout.contiguous())
        return (None, None, None, None) + tuple(eng.grad[n] for n in eng.trainable_names())


class UNetModel:
    """Stand-in for the `tf.keras.Model` held in `UNet.model`."""

    def __init__(self, engine: UNetEngine, name="U-Net"):
        self.engine = engine
        self.name = name
        self._leaves = None

    # -- variables --------------------------------------------------------------------------
    @property
    def trainable_variables(self):
        if self._leaves is None:
            self._leaves = [self.engine.param[n].requires_grad_(True) for n in self.engine.trainable_names()]
        return self._leaves

    @property
    def variables(self):
        return list(self.engine.param.values()) + list(self.engine.state.values())

    @property
    def losses(self):
        """kernel_regularizer=l2(0.001) terms (u_net.py:274,302), one scalar per regularised kernel."""
        out = []
        for n in self.engine.trainable_names():
            if PL.l2_regularised(n):
                t = torch.zeros(1, device=self.engine.device)
                p = self.engine.param[n]
                L.call("sumsq", p.data_ptr(), p.numel(), PL.L2_COEF, t.data_ptr(), 0)
                out.append(t[0])
        return out

    # -- call -------------------------------------------------------------------------------
    def __call__(self, inputs, training=False, dropout_mask=None):
        spec, emb = inputs
        dev = self.engine.device
        spec = _to_dev(spec, torch.float32, dev)
        emb = _to_dev(emb, torch.int32, dev)
        if training and torch.is_grad_enabled():
            return _UNetFn.apply(self, spec, emb, dropout_mask, *self.trainable_variables)
        return self.engine.forward(spec, emb, training=training, dropout_mask=dropout_mask).clone()

    def predict(self, inputs, batch_size=32, verbose=0):
        spec, emb = inputs
        outs = []
        with torch.no_grad():
            for i in range(0, len(spec), batch_size):
                outs.append(self([spec[i:i + batch_size], emb[i:i + batch_size]], training=False).cpu())
        return torch.cat(outs).numpy()

    # -- weights ----------------------------------------------------------------------------
    def get_weights(self):
        return [v.detach().cpu().numpy() for v in self.variables]

    def save_weights(self, path):
        torch.save(self.engine.state_dict(), path)

    def load_weights(self, path):
        """`.pt`: a state dict saved by save_weights; `.npz`: weights trained with the reference under TensorFlow,
        exported by tools/export_tf_weights.py (keras_weights.read_keras_npz maps them onto the plan)."""
        if str(path).endswith(".npz"):
            from ..keras_weights import read_keras_npz
            self.engine.load_state_dict(read_keras_npz(path, self.engine.plan))
        else:
            self.engine.load_state_dict(torch.load(path, map_location="cpu"))

    def count_params(self):
        return sum(v.numel() for v in self.variables)

    def summary(self, print_fn=print):
        eng = self.engine
        print_fn(f'Model: "{self.name}"')
        print_fn(f"{'variable':34s}{'shape':>24s}{'params':>12s}")
        for name, shape, kind in eng.plan:
            n = int(np.prod(shape))
            print_fn(f"{name:34s}{str(tuple(shape)):>24s}{n:12d}")
        tr = sum(int(np.prod(s)) for n, s, k in eng.plan if k in PL.TRAINABLE_KINDS)
        tot = sum(int(np.prod(s)) for n, s, k in eng.plan)
        print_fn(f"Total params: {tot:,}\nTrainable params: {tr:,}\nNon-trainable params: {tot - tr:,}")


class UNet:
    """Same constructor as the reference (dl_models/u_net.py:40-45)."""

    def __init__(self, input_shape, inf_vector_shape,
                 learning_rate=1e-5,
                 mode=0, number_filters_0=32, kernels=6, BatchNorm=True,
                 resize_factor_0=None, res_factor=None,
                 name="U-Net"
                 ):
        # the reference only sets these attributes when the arguments are None (u_net.py:46-49)
        self.res_factor = [2, 2] if res_factor is None else res_factor
        self.resize_factor_0 = [1, 1] if resize_factor_0 is None else resize_factor_0
        if list(self.res_factor) != [2, 2] or list(self.resize_factor_0) != [1, 1]:
            raise NotImplementedError("only res_factor=[2,2], resize_factor_0=[1,1] (the reference defaults) are built")

        self.input_shape = input_shape
        self.inf_vector_shape = inf_vector_shape
        self.learning_rate = learning_rate
        self.mode = mode
        self.number_filters_0 = number_filters_0
        self.kernels = kernels
        self.BatchNorm = BatchNorm
        self.name = name

        self.model = None
        self._model_input = None
        self._build()

    def _build(self):
        engine = UNetEngine(self.input_shape, self.inf_vector_shape, self.mode, self.number_filters_0,
                            self.kernels, self.BatchNorm)
        self.model = UNetModel(engine, name="U-Net")

    def summary(self):
        self.model.summary()

    def get_callbacks(self):
        """The reference returns Keras CSVLogger + EarlyStopping(patience=20) (u_net.py:72-81); here the
        equivalents from amp_phase_trainer are returned for use with compile_and_fit."""
        from ..amp_phase_trainer import EarlyStopping, ModelCheckpoint
        return [ModelCheckpoint(f"{self.name}_ckpt", save_best_only=False, verbose=0), EarlyStopping(patience=20)]

    def compile_and_fit(self, x_train1, x_train2, y_train, x_val1, x_val2,
                        y_val, batch_size, num_epochs, steps_per_epoch):
        """model.compile(Adam(InverseTimeDecay), MSE) + fit(shuffle=False) (u_net.py:83-118).
        Returns a dict with 'loss' and 'val_loss' per epoch like `History.history`."""
        from ..amp_phase_trainer import fit_mse
        return fit_mse(self, x_train1, x_train2, y_train, x_val1, x_val2, y_val, batch_size, num_epochs,
                       steps_per_epoch, self.learning_rate, self.get_callbacks())

    def save(self, save_folder="."):
        self._create_folder_if_it_doesnt_exist(save_folder)
        self._save_parameters(save_folder)
        self._save_weights(save_folder)

    def load_weights(self, weights_path):
        self.model.load_weights(weights_path)

    def predict_stft(self, inputs):
        return self.model.predict(inputs)

    @classmethod
    def load(cls, save_folder="."):
        with open(os.path.join(save_folder, "parameters.pkl"), "rb") as f:
            parameters = pickle.load(f)
        ue = UNet(**parameters) if isinstance(parameters, dict) else UNet(*parameters)
        ue.load_weights(os.path.join(save_folder, "weights.pt"))
        return ue

    @staticmethod
    def _create_folder_if_it_doesnt_exist(folder):
        if not os.path.exists(folder):
            os.makedirs(folder)

    def _save_parameters(self, save_folder):
        parameters = dict(input_shape=self.input_shape, inf_vector_shape=self.inf_vector_shape,
                          learning_rate=self.learning_rate, mode=self.mode,
                          number_filters_0=self.number_filters_0, kernels=self.kernels,
                          BatchNorm=self.BatchNorm)
        with open(os.path.join(save_folder, "parameters.pkl"), "wb") as f:
            pickle.dump(parameters, f)

    def _save_weights(self, save_folder):
        self.model.save_weights(os.path.join(save_folder, "weights.pt"))


if __name__ == "__main__":
    unet = UNet(input_shape=(144, 160, 2),
                inf_vector_shape=(2, 16),
                mode=0,
                number_filters_0=32,
                kernels=3,
                name='Unet'
                )
    unet.summary()
