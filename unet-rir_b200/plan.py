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


def layer_plan(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0, number_filters_0=32,
               kernels=6, BatchNorm=True):
    """-> list of (name, shape, kind)."""
    F0, k = number_filters_0, kernels
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
        conv(f"enc{i + 1}.down", k, cin, n)
        block(f"enc{i + 1}.blk", n, n)
        cin = n
    H5, W5 = input_shape[0] // 16, input_shape[1] // 16
    dim = H5 * W5 * VEC_CH
    plan.append(("vec.emb", (EMB_VOCAB, EMB_DIM), "emb"))
    plan.append(("vec.dense.w", (inf_vector_shape[0] * inf_vector_shape[1] * EMB_DIM, dim), "dense_w"))
    plan.append(("vec.dense.b", (dim,), "bias"))
    conv("vec.proj", 1, VEC_CH, F0 * 16)
    for j, m in zip([2, 3, 4, 5], [8, 4, 2, 1]):
        n = F0 * m
        convT(f"dec{j}.up", k, cin, n)
        conv(f"dec{j}.fuse", k, 2 * n, n); bn(f"dec{j}.fuse_bn", n)
        block(f"dec{j}.blk", n, n)
        cin = n
    conv("head", 6, cin, 2)
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
