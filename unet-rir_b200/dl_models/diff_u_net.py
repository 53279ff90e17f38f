"""`DiffUNet` with the reference's constructor, attributes and methods (dl_models/diff_u_net.py:34-199), on the same
sm_100a engine as `UNet`.

What differs from `UNet` in the reference's graph (diff_u_net.py:205-322), and therefore here (plan.ARCHS["diff"]):
  * strided convolutions and Conv2DTranspose use 2x2 kernels (:272-279, :300-307; stride 1 in block 1, SAME padding
    puts the single pad row / column after the data), the decoder's fuse convolution is 3x3 (:312);
  * vector block: Embedding(1500, 128) -> Flatten -> Dense(H5 * W5 * 16 F0) -> Dropout(.5) -> Reshape, added to the
    bottleneck with no 1x1 projection (:261-270, :230-231);
  * head: UpSampling2D((1, 1)) (identity) -> Conv2D(2, 1x1, linear) (:256-257), so outputs are not confined to (0, 1);
  * no `kernels` constructor argument; `parameters.pkl` holds six entries (:180-187).

`DiffUNet(...).model([spec_in, emb], training=bool)`, `.trainable_variables`, `.losses`, `save` / `load` /
`load_weights` / `predict_stft` / `compile_and_fit` behave as documented in dl_models/u_net.py of this package; the
generic trainer (trainer.py) and the amp/phase trainer both differentiate through the linear head.
"""
from __future__ import annotations

import os
import pickle

import torch

from ..engine import UNetEngine
from .u_net import UNet, UNetModel


class DiffUNet(UNet):
    """Same constructor as the reference (dl_models/diff_u_net.py:40-45)."""

    def __init__(self, input_shape, inf_vector_shape,
                 learning_rate=1e-5,
                 mode=0, number_filters_0=32, BatchNorm=True,
                 resize_factor_0=None, res_factor=None,
                 name="U-Net"
                 ):
        # the reference only sets these attributes when the arguments are None (diff_u_net.py:46-49)
        self.res_factor = [2, 2] if res_factor is None else res_factor
        self.resize_factor_0 = [1, 1] if resize_factor_0 is None else resize_factor_0
        if list(self.res_factor) != [2, 2] or list(self.resize_factor_0) != [1, 1]:
            raise NotImplementedError("only res_factor=[2,2], resize_factor_0=[1,1] (the reference defaults) are built")
        self.input_shape = input_shape
        self.inf_vector_shape = inf_vector_shape
        self.learning_rate = learning_rate
        self.mode = mode
        self.number_filters_0 = number_filters_0
        self.kernels = 2                     # not a constructor argument: the strided kernels are 2x2 (:275, :303)
        self.BatchNorm = BatchNorm
        self.name = name
        self.model = None
        self._model_input = None
        self._build()

    def _build(self):
        engine = UNetEngine(self.input_shape, self.inf_vector_shape, self.mode, self.number_filters_0,
                            self.kernels, self.BatchNorm, arch="diff")
        self.model = UNetModel(engine, name="U-Net")

    @classmethod
    def load(cls, save_folder="."):
        with open(os.path.join(save_folder, "parameters.pkl"), "rb") as f:
            parameters = pickle.load(f)
        ue = DiffUNet(*parameters)
        ue.load_weights(os.path.join(save_folder, "weights.pt"))
        return ue

    def _save_parameters(self, save_folder):
        parameters = [self.input_shape, self.inf_vector_shape, self.learning_rate, self.mode, self.number_filters_0,
                      self.BatchNorm]
        with open(os.path.join(save_folder, "parameters.pkl"), "wb") as f:
            pickle.dump(parameters, f)

    @staticmethod
    def mse_coef(y_true, y_pred):
        """mean squared error over all elements (diff_u_net.py:386-394)."""
        d = torch.as_tensor(y_true, dtype=torch.float32).reshape(-1) - torch.as_tensor(y_pred, dtype=torch.float32).reshape(-1)
        return (d * d).mean()

    @staticmethod
    def rmse_coef(y_true, y_pred):
        """sqrt(mean squared error + 1e-12) over all elements (diff_u_net.py:396-403)."""
        return torch.sqrt(DiffUNet.mse_coef(y_true, y_pred) + 1.0e-12)

    @staticmethod
    def rmse_coef_slicing(y_true, y_pred):
        """RMSE over the slice [0:32, 0:160, 20:32, 0:1] of both tensors (diff_u_net.py:405-416)."""
        a = torch.as_tensor(y_true, dtype=torch.float32)[0:32, 0:160, 20:32, 0:1]
        b = torch.as_tensor(y_pred, dtype=torch.float32)[0:32, 0:160, 20:32, 0:1]
        return torch.sqrt(((a - b) ** 2).mean())

    @staticmethod
    def l1_norm(y_true, y_pred):
        """sum of absolute differences (diff_u_net.py:418-426)."""
        return (torch.as_tensor(y_true, dtype=torch.float32) - torch.as_tensor(y_pred, dtype=torch.float32)).abs().sum()


if __name__ == "__main__":
    unet = DiffUNet(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32, name='DiffUnet')
    unet.summary()
