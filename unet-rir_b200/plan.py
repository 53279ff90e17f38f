"""Variable inventory of the U-Net, in Keras creation order.

Mirrors UNet._build of the reference (dl_models/u_net.py:201-251): the order here is the order
of `model.trainable_variables` that Trainer.step walks (amp_phase_trainer.py:137-139), so flat
parameter / gradient / Adam-moment buffers laid out in this order are interchangeable with a list
of Keras variables. Shapes use the Keras layouts: Conv2D HWIO, Conv2DTranspose (kh, kw, out, in),
Dense (in, out), Embedding (vocab, dim).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

TRAINABLE_KINDS = ("conv_w", "convT_w", "bias", "gamma", "beta", "emb", "dense_w")
STATE_KINDS = ("moving_mean", "moving_var")
EMB_VOCAB, EMB_DIM = 2000, 256          # Embedding(2000, 256)   u_net.py:257
VEC_CH = 16                             # shape[2] = 16           u_net.py:255
DROPOUT_RATE = 0.3                      # Dropout(.3)             u_net.py:260
BN_EPS, BN_MOMENTUM = 1e-3, 0.99        # Keras BatchNormalization defaults (u_net.py:368)
L2_COEF = 1e-3                          # kernel_regularizer=l2(0.001) (u_net.py:274,302)

# The two sibling graphs built on this plan. None = "the constructor's `kernels` argument".
#   unet: dl_models/u_net.py:201-263   -- k x k strided / transposed / fuse convolutions, Embedding(2000, 256),
#         Dense -> Dropout(.3) -> Reshape(H5, W5, 16) -> 1x1 projection, 6x6 sigmoid head
#   diff: dl_models/diff_u_net.py:205-276 -- 2x2 strided / transposed convolutions (:272-279, :300-307), 3x3 fuse
#         convolution (:312), Embedding(1500, 128), Dense straight to (H5, W5, 16 F0) -> Dropout(.5) (:261-270, no
#         projection), linear 1x1 head (:257)
ARCHS = {
    "unet": dict(down_k=None, up_k=None, fuse_k=None, head_k=6, head_sigmoid=True, emb_vocab=EMB_VOCAB, emb_dim=EMB_DIM,
                 vec_ch=VEC_CH, proj=True, dropout=DROPOUT_RATE),
    "diff": dict(down_k=2, up_k=2, fuse_k=3, head_k=1, head_sigmoid=False, emb_vocab=1500, emb_dim=128,
                 vec_ch=None, proj=False, dropout=0.5),
}


def arch_params(arch, number_filters_0, kernels):
    a = dict(ARCHS[arch])
    for key in ("down_k", "up_k", "fuse_k"):
        if a[key] is None:
            a[key] = kernels
    if a["vec_ch"] is None:
        a["vec_ch"] = number_filters_0 * 16
    return a


def layer_plan(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32,
               kernels=6, BatchNorm=True, arch="unet"):
    """-> list of (name, shape, kind)."""
    F0 = number_filters_0
    A = arch_params(arch, number_filters_0, kernels)
    plan = []

    def conv(name, kh, cin, cout):
        plan.append((name + ".w", (kh, kh, cin, cout), "conv_w"))
        plan.append((name + ".b", (cout,), "bias"))

    def convT(name, kh, cin, cout):
        plan.append((name + ".w", (kh, kh, cout, cin), "convT_w"))
        plan.append((name + ".b", (cout,), "bias"))

    def bn(name, c):
        if BatchNorm:
            for suffix, kind in ((".gamma", "gamma"), (".beta", "beta"),
                                 (".moving_mean", "moving_mean"), (".moving_var", "moving_var")):
                plan.append((name + suffix, (c,), kind))

    def block(name, cin, n):
        # u_net.py:324-386: mode 0 one conv+BN+ReLU; modes 1,2 two; mode 3 two + a conv shortcut
        conv(name + ".c1", 3, cin, n); bn(name + ".bn1", n)
        if mode in (1, 2, 3):
            conv(name + ".c2", 3, n, n); bn(name + ".bn2", n)
        if mode == 3:
            conv(name + ".c3", 3, cin, n); bn(name + ".bn3", n)

    cin = input_shape[2]
    for i, m in enumerate([1, 2, 4, 8, 16]):
        n = F0 * m
        conv(f"enc{i + 1}.down", A["down_k"], cin, n)
        block(f"enc{i + 1}.blk", n, n)
        cin = n
    H5, W5 = input_shape[0] // 16, input_shape[1] // 16
    dim = H5 * W5 * A["vec_ch"]
    plan.append(("vec.emb", (A["emb_vocab"], A["emb_dim"]), "emb"))
    plan.append(("vec.dense.w", (inf_vector_shape[0] * inf_vector_shape[1] * A["emb_dim"], dim), "dense_w"))
    plan.append(("vec.dense.b", (dim,), "bias"))
    if A["proj"]:
        conv("vec.proj", 1, A["vec_ch"], F0 * 16)
    for j, m in zip([2, 3, 4, 5], [8, 4, 2, 1]):
        n = F0 * m
        convT(f"dec{j}.up", A["up_k"], cin, n)
        conv(f"dec{j}.fuse", A["fuse_k"], 2 * n, n); bn(f"dec{j}.fuse_bn", n)
        block(f"dec{j}.blk", n, n)
        cin = n
    conv("head", A["head_k"], cin, 2)
    return plan


def keras_init(plan, seed=500):
    """Keras-default initialisers: glorot_uniform kernels, zero bias, Embedding U(-.05,.05),
    BN gamma = 1, beta = 0, moving mean 0 / var 1. CPU float32 tensors."""
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, shape, kind in plan:
        if kind in ("conv_w", "convT_w"):
            kh, kw, a, b = shape
            lim = math.sqrt(6.0 / (kh * kw * a + kh * kw * b))
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif kind == "dense_w":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g) * 2 - 1) * lim
        elif kind == "emb":
            t = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind in ("gamma", "moving_var"):
            t = torch.ones(shape)
        else:
            t = torch.zeros(shape)
        out[name] = t
    return out


def l2_regularised(name: str) -> bool:
    """l2(0.001) sits on the strided encoder convs and the ConvTs only (u_net.py:269-275, 297-303)."""
    return name.endswith(".w") and (".down" in name or ".up" in name)
