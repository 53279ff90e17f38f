"""CPU oracle for the U-Net hot path -- TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (igmsalinas/unet-rir) ships no tests, fixtures or golden
vectors, and TensorFlow/Keras cannot be imported in the build container, so this file is a
restatement of the reference arithmetic that is justified line by line against its source and
against the documented TF/Keras defaults, not against reference outputs.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product path (unet_rir_b200) never does.

What is restated (all citations into /root/reference):
  * dl_models/u_net.py:201-251   UNet._build          -> UNetOracle.forward
  * dl_models/u_net.py:253-263   vector_block         -> UNetOracle._vector_block
  * dl_models/u_net.py:265-289   encoding_block       -> UNetOracle._encoding_block
  * dl_models/u_net.py:291-321   decoding_block       -> UNetOracle._decoding_block
  * dl_models/u_net.py:324-386   the four block modes -> UNetOracle._block
  * amp_phase_trainer.py:143-168 model_loss           -> amp_phase_loss
  * main_training.py:203-235     compute_loss (DP)    -> dp_loss
  * amp_phase_trainer.py:30-35   Keras Adam/SGD/Nadam -> keras_adam_step / keras_sgd_step / keras_nadam_step
  * trainer.py:37-38             tfa LAMB             -> tfa_lamb_step

Tensors are NHWC float32 at the interface, exactly like the reference's
(datageneratorv2.py:88-102). Kernels are stored in Keras layouts: Conv2D HWIO,
Conv2DTranspose HWOI (kh, kw, out, in), Dense (in, out), Embedding (vocab, dim).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3        # Keras BatchNormalization default epsilon (u_net.py:368 passes no args)
BN_MOMENTUM = 0.99   # Keras default momentum
L2_COEF = 1e-3       # kernel_regularizer=l2(0.001) (u_net.py:274,302)


# ----------------------------------------------------------------------------------------
# TF "SAME" padding arithmetic
# ----------------------------------------------------------------------------------------
def same_pad(in_size: int, k: int, s: int):
    """TF SAME: out = ceil(in/s); total = max((out-1)*s + k - in, 0); before = total // 2."""
    out = -(-in_size // s)
    total = max((out - 1) * s + k - in_size, 0)
    before = total // 2
    return out, before, total - before


def conv2d_same(x_nchw, w_hwio, b, stride):
    """tf.keras.layers.Conv2D(padding='same') on an NCHW tensor with a HWIO kernel."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    _, pt, pb = same_pad(x_nchw.shape[2], kh, stride)
    _, pl, pr = same_pad(x_nchw.shape[3], kw, stride)
    x = F.pad(x_nchw, (pl, pr, pt, pb))
    w = w_hwio.permute(3, 2, 0, 1)  # OIHW
    return F.conv2d(x, w, b, stride=stride)


def conv2d_transpose_same(x_nchw, w_hwoi, b, stride):
    """tf.keras.layers.Conv2DTranspose(padding='same').

    TF defines it as the input-gradient of the SAME forward conv whose input has size in*s:
    y[o] = sum_{i,r : i*s + r - before = o} x[i] w[r], o in [0, in*s), where `before` is the
    forward conv's leading pad. Built here from the full transposed conv, then cropped.
    """
    kh, kw = w_hwoi.shape[0], w_hwoi.shape[1]
    H, W = x_nchw.shape[2], x_nchw.shape[3]
    _, pt, _ = same_pad(H * stride, kh, stride)
    _, pl, _ = same_pad(W * stride, kw, stride)
    w = w_hwoi.permute(3, 2, 0, 1)  # (in, out, kh, kw) as torch wants
    full = F.conv_transpose2d(x_nchw, w, None, stride=stride)
    need_h, need_w = pt + H * stride, pl + W * stride
    if full.shape[2] < need_h or full.shape[3] < need_w:   # k < s corner: zero extend
        full = F.pad(full, (0, max(need_w - full.shape[3], 0), 0, max(need_h - full.shape[2], 0)))
    y = full[:, :, pt:pt + H * stride, pl:pl + W * stride]
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    return y


# ----------------------------------------------------------------------------------------
# Parameter inventory (Keras creation order == model.trainable_variables order)
# ----------------------------------------------------------------------------------------
# What the sibling DiffUNet (dl_models/diff_u_net.py:205-322) changes in the graph: 2x2 strided / transposed kernels
# (:272-279, :300-307), 3x3 fuse convolution (:312), Embedding(1500, 128) -> Dense(H5 W5 16 F0) -> Dropout(.5) added to the
# bottleneck without a 1x1 projection (:261-270, :230-231), linear 1x1 head (:256-257).
def arch_table(arch, number_filters_0, kernels):
    if arch == "unet":
        return dict(down_k=kernels, up_k=kernels, fuse_k=kernels, head_k=6, head_sigmoid=True, emb_vocab=2000, emb_dim=256,
                    vec_ch=16, proj=True)
    if arch == "diff":
        return dict(down_k=2, up_k=2, fuse_k=3, head_k=1, head_sigmoid=False, emb_vocab=1500, emb_dim=128,
                    vec_ch=16 * number_filters_0, proj=False)
    raise ValueError(arch)


def layer_plan(input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0,
               number_filters_0=32, kernels=6, BatchNorm=True, arch="unet"):
    """Returns the ordered list of (name, shape, kind) for every variable of the model.

    kind in {conv_w, convT_w, bias, gamma, beta, moving_mean, moving_var, emb, dense_w}.
    Order follows layer creation order in UNet._build (u_net.py:201-251), which is the order of
    `model.trainable_variables` that Trainer.step walks (amp_phase_trainer.py:137-139).
    """
    F0 = number_filters_0
    A = arch_table(arch, number_filters_0, kernels)
    plan = []

    def conv(name, kh, cin, cout):
        plan.append((name + ".w", (kh, kh, cin, cout), "conv_w"))
        plan.append((name + ".b", (cout,), "bias"))

    def convT(name, kh, cin, cout):
        plan.append((name + ".w", (kh, kh, cout, cin), "convT_w"))
        plan.append((name + ".b", (cout,), "bias"))

    def bn(name, c):
        if BatchNorm:
            plan.append((name + ".gamma", (c,), "gamma"))
            plan.append((name + ".beta", (c,), "beta"))
            plan.append((name + ".moving_mean", (c,), "moving_mean"))
            plan.append((name + ".moving_var", (c,), "moving_var"))

    def block(name, cin, n, first_k=3):
        # u_net.py:324-386 ; every mode starts with conv(first_k)+BN+ReLU
        if mode == 0:
            conv(name + ".c1", first_k, cin, n); bn(name + ".bn1", n)
        elif mode == 1:
            conv(name + ".c1", 3, cin, n); bn(name + ".bn1", n)
            conv(name + ".c2", 3, n, n); bn(name + ".bn2", n)
        elif mode == 2:
            conv(name + ".c1", 3, cin, n); bn(name + ".bn1", n)
            conv(name + ".c2", 3, n, n); bn(name + ".bn2", n)
        elif mode == 3:
            conv(name + ".c1", 3, cin, n); bn(name + ".bn1", n)
            conv(name + ".c2", 3, n, n); bn(name + ".bn2", n)
            conv(name + ".c3", 3, cin, n); bn(name + ".bn3", n)

    cin = input_shape[2]
    mults = [1, 2, 4, 8, 16]
    for i, m in enumerate(mults):
        n = F0 * m
        conv(f"enc{i+1}.down", A["down_k"], cin, n)
        block(f"enc{i+1}.blk", n, n)
        cin = n
    H5 = input_shape[0] // 16
    W5 = input_shape[1] // 16
    dim = H5 * W5 * A["vec_ch"]
    plan.append(("vec.emb", (A["emb_vocab"], A["emb_dim"]), "emb"))
    plan.append(("vec.dense.w", (inf_vector_shape[0] * inf_vector_shape[1] * A["emb_dim"], dim), "dense_w"))
    plan.append(("vec.dense.b", (dim,), "bias"))
    if A["proj"]:
        conv("vec.proj", 1, A["vec_ch"], F0 * 16)
    for j, m in zip([2, 3, 4, 5], [8, 4, 2, 1]):
        n = F0 * m
        convT(f"dec{j}.up", A["up_k"], cin, n)
        # concat([skip(n), up(n)]) -> conv(kernels)+BN+ReLU (u_net.py:308-310) -> mode block
        conv(f"dec{j}.fuse", A["fuse_k"], 2 * n, n); bn(f"dec{j}.fuse_bn", n)
        block(f"dec{j}.blk", n, n)
        cin = n
    conv("head", A["head_k"], cin, 2)
    return plan


def init_params(plan, seed=500, dtype=torch.float32):
    """Keras-default initialisers: glorot_uniform kernels, zero bias,
    Embedding U(-0.05, 0.05), BN gamma=1 beta=0 mean=0 var=1."""
    g = torch.Generator().manual_seed(seed)
    p = OrderedDict()
    for name, shape, kind in plan:
        if kind in ("conv_w", "convT_w"):
            kh, kw, a, b = shape
            # keras glorot: fan_in = kh*kw*shape[-2], fan_out = kh*kw*shape[-1]
            fan_in, fan_out = kh * kw * a, kh * kw * b
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * lim
        elif kind == "dense_w":
            lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            t = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * lim
        elif kind == "emb":
            t = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * 0.05
        elif kind in ("bias", "beta", "moving_mean"):
            t = torch.zeros(shape, dtype=dtype)
        elif kind in ("gamma", "moving_var"):
            t = torch.ones(shape, dtype=dtype)
        else:
            raise ValueError(kind)
        p[name] = t
    return p


TRAINABLE_KINDS = ("conv_w", "convT_w", "bias", "gamma", "beta", "emb", "dense_w")


def trainable_names(plan):
    return [n for n, _, kind in plan if kind in TRAINABLE_KINDS]


def l2_regularised_names(plan):
    """kernel_regularizer=l2 sits only on the strided encoder convs and the ConvTs
    (u_net.py:269-275, 297-303)."""
    return [n for n, _, kind in plan if n.endswith(".w") and (".down" in n or ".up" in n)]


# ----------------------------------------------------------------------------------------
# The model
# ----------------------------------------------------------------------------------------
class UNetOracle:
    """Functional restatement of UNet._build. `params` is an OrderedDict name->tensor."""

    def __init__(self, input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0,
                 number_filters_0=32, kernels=6, BatchNorm=True, bn_moving_var_unbiased=False,
                 emulate_bf16=False, arch="unet"):
        # emulate_bf16: evaluate the SAME graph under the device path's storage contract -- conv / Dense
        # operands (activations, kernels) and every stored activation rounded to bfloat16, fp32
        # accumulation, fp32 BatchNorm statistics taken before the rounding (straight-through in
        # backward). ReLU gates then open and close on the same values as on the device, which is what
        # makes per-tensor GRADIENT comparisons meaningful: against the pure-fp32 evaluation ~0.3 % of the
        # gates differ and every ReLU layer adds ~5 % rel-L2 of (unbiased) gradient noise.
        self.q = emulate_bf16
        self.input_shape = tuple(input_shape)
        self.inf_vector_shape = tuple(inf_vector_shape)
        self.mode = mode
        self.F0 = number_filters_0
        self.kernels = kernels
        self.BatchNorm = BatchNorm
        self.bn_unbiased = bn_moving_var_unbiased
        self.arch = arch
        self.A = arch_table(arch, number_filters_0, kernels)
        self.plan = layer_plan(input_shape, inf_vector_shape, mode, number_filters_0, kernels, BatchNorm, arch)
        self.taps = None   # optional dict collecting intermediates for per-layer parity
        # optional dict name -> tensor: forward VALUES to substitute (straight-through) at every stored
        # tensor and BatchNorm statistic. With the device path's own forward state plugged in, autograd
        # back-propagates through exactly the ReLU gates / saved activations the device used, which turns
        # the (chaotic) end-to-end gradient comparison into a test of the backward kernels alone.
        self.override = None

    # -- layers ---------------------------------------------------------------------
    def _r(self, t):
        """round-to-bf16 with a straight-through gradient (identity when not emulating)."""
        if not self.q:
            return t
        return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()

    def _sub(self, name, t):
        """value substitution with a straight-through gradient"""
        if self.override is not None and name in self.override:
            return t + (self.override[name].to(t.dtype) - t).detach()
        return t

    def _tap(self, name, t):
        if self.taps is not None:
            self.taps[name] = t

    def _bn_relu(self, x, p, name, training, new_stats):
        if self.BatchNorm:
            g, b = p[name + ".gamma"], p[name + ".beta"]
            if training:
                mean = x.mean(dim=(0, 2, 3))
                var = x.var(dim=(0, 2, 3), unbiased=False)
                if self.q:      # device: statistics from the fp32 accumulators, gradient flows via the stored x
                    mean, var = mean.detach(), var.detach()
                    xq = self._r(x)
                    mean = mean + (xq.mean(dim=(0, 2, 3)) - xq.mean(dim=(0, 2, 3)).detach())
                    var = var + (xq.var(dim=(0, 2, 3), unbiased=False) - xq.var(dim=(0, 2, 3), unbiased=False).detach())
                mean, var = self._sub(name + ".mean", mean), self._sub(name + ".var", var)
                if new_stats is not None:
                    n = x.shape[0] * x.shape[2] * x.shape[3]
                    mv = var * (n / max(n - 1, 1)) if self.bn_unbiased else var
                    new_stats[name + ".moving_mean"] = (
                        p[name + ".moving_mean"] * BN_MOMENTUM + mean.detach() * (1 - BN_MOMENTUM))
                    new_stats[name + ".moving_var"] = (
                        p[name + ".moving_var"] * BN_MOMENTUM + mv.detach() * (1 - BN_MOMENTUM))
            else:
                mean, var = p[name + ".moving_mean"], p[name + ".moving_var"]
            inv = torch.rsqrt(var + BN_EPS)
            x = (self._r(x) - mean.view(1, -1, 1, 1)) * (inv * g).view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        if self.override is not None and (name + ".out") in self.override:
            # pin the ReLU gates to the substituted output: gradient passes exactly where that output is > 0
            y = x * (self.override[name + ".out"] > 0).to(x.dtype)
            return self._sub(name + ".out", y)
        return self._r(F.relu(x))

    def _cbr(self, x, p, cname, bname, training, new_stats):
        x = self._sub(cname, conv2d_same(x, self._r(p[cname + ".w"]), p[cname + ".b"], 1))
        self._tap(cname, x)
        return self._bn_relu(x, p, bname, training, new_stats)

    def _block(self, x, p, name, training, new_stats):
        m = self.mode
        if m == 0:      # convolutional_block_1 (u_net.py:363-371)
            return self._cbr(x, p, name + ".c1", name + ".bn1", training, new_stats)
        if m == 1:      # convolutional_block_2 (u_net.py:373-386)
            y = self._cbr(x, p, name + ".c1", name + ".bn1", training, new_stats)
            return self._cbr(y, p, name + ".c2", name + ".bn2", training, new_stats)
        if m == 2:      # residual_block_1 (u_net.py:324-339)
            y = self._cbr(x, p, name + ".c1", name + ".bn1", training, new_stats)
            y = self._cbr(y, p, name + ".c2", name + ".bn2", training, new_stats)
            return y + x
        if m == 3:      # residual_block_2 (u_net.py:341-361)
            y = self._cbr(x, p, name + ".c1", name + ".bn1", training, new_stats)
            y = self._cbr(y, p, name + ".c2", name + ".bn2", training, new_stats)
            z = self._cbr(x, p, name + ".c3", name + ".bn3", training, new_stats)
            return y + z
        raise ValueError(m)

    def _encoding_block(self, x, p, i, stride, training, new_stats):
        x = self._sub(f"enc{i}.down", self._r(conv2d_same(x, self._r(p[f"enc{i}.down.w"]), p[f"enc{i}.down.b"], stride)))   # no BN / act
        self._tap(f"enc{i}.down", x)
        return self._block(x, p, f"enc{i}.blk", training, new_stats)

    def _decoding_block(self, x, skip, p, j, training, new_stats):
        x = self._sub(f"dec{j}.up", self._r(conv2d_transpose_same(x, self._r(p[f"dec{j}.up.w"]), p[f"dec{j}.up.b"], 2)))
        self._tap(f"dec{j}.up", x)
        x = torch.cat([skip, x], dim=1)            # skip FIRST (u_net.py:308)
        x = self._cbr(x, p, f"dec{j}.fuse", f"dec{j}.fuse_bn", training, new_stats)
        return self._block(x, p, f"dec{j}.blk", training, new_stats)

    def _vector_block(self, emb_idx, p, training, dropout_mask):
        B = emb_idx.shape[0]
        H5, W5 = self.input_shape[0] // 16, self.input_shape[1] // 16
        f = p["vec.emb"][emb_idx.long()]                        # (B, 2, 16, 256)
        x = self._r(f.reshape(B, -1))                          # Flatten, row-major
        x = x @ self._r(p["vec.dense.w"]) + p["vec.dense.b"]   # Dense(dim)
        self._tap("vec.dense", x)
        if training and dropout_mask is not None:              # Dropout(.3), inverted
            x = x * dropout_mask
        x = self._sub("vec.dense.out", self._r(x))
        x = x.reshape(B, H5, W5, self.A["vec_ch"]).permute(0, 3, 1, 2)       # Reshape((H5, W5, 16)) NHWC
        if self.A["proj"]:
            x = conv2d_same(x, self._r(p["vec.proj.w"]), p["vec.proj.b"], 1)
        return x

    # -- forward ----------------------------------------------------------------------
    def forward(self, params, spec_nhwc, emb_idx, training=False, dropout_mask=None,
                new_stats=None):
        """model([spec, emb], training=...) -> (B, H, W, 2) in (0, 1).

        dropout_mask: (B, dim) tensor of {0, 1/(1-rate)} injected for determinism;
        None means rate 0. new_stats: dict that receives the updated BN moving statistics.
        """
        p = params
        x = spec_nhwc.permute(0, 3, 1, 2)
        e1 = self._encoding_block(x, p, 1, 1, training, new_stats)
        e2 = self._encoding_block(e1, p, 2, 2, training, new_stats)
        e3 = self._encoding_block(e2, p, 3, 2, training, new_stats)
        e4 = self._encoding_block(e3, p, 4, 2, training, new_stats)
        e5 = self._encoding_block(e4, p, 5, 2, training, new_stats)
        v = self._vector_block(emb_idx, p, training, dropout_mask)
        z = self._sub("bottleneck", self._r(e5 + v))            # Add() (u_net.py:229)
        self._tap("bottleneck", z)
        d2 = self._decoding_block(z, e4, p, 2, training, new_stats)
        d3 = self._decoding_block(d2, e3, p, 3, training, new_stats)
        d4 = self._decoding_block(d3, e2, p, 4, training, new_stats)
        d5 = self._decoding_block(d4, e1, p, 5, training, new_stats)
        out = conv2d_same(d5, self._r(p["head.w"]), p["head.b"], 1)   # UpSampling2D((1,1)) = identity
        self._tap("head", out)
        if not self.A["head_sigmoid"]:                                # DiffUNet: activation='linear'
            return out.permute(0, 2, 3, 1)
        return torch.sigmoid(out).permute(0, 2, 3, 1)

    def l2_losses(self, params):
        """model.model.losses : one 0.001*sum(w^2) per regularised kernel."""
        return [L2_COEF * (params[n] ** 2).sum() for n in l2_regularised_names(self.plan)]


# ----------------------------------------------------------------------------------------
# Losses
# ----------------------------------------------------------------------------------------
def amp_phase_loss(y_true, y_pred):
    """Trainer.model_loss (amp_phase_trainer.py:143-168): returns (loss, loss_phase, loss_stft)."""
    a_t, p_t = y_true[..., 0], y_true[..., 1]
    a_p, p_p = y_pred[..., 0], y_pred[..., 1]
    loss_stft = ((a_t - a_p) ** 2).mean()
    pt = p_t * 2 * math.pi - math.pi
    pp = p_p * 2 * math.pi - math.pi
    loss_phase = (1 - torch.cos(pt - pp)).mean()
    return loss_phase + loss_stft, loss_phase, loss_stft


def dp_loss(y_true, y_pred, alpha, global_batch_size, l2_losses=None, num_replicas=1):
    """compute_loss of main_training.py:203-235 for one replica's shard.

    loss_object_amplitude on expand_dims(...,-1) is the per-element squared error; phase_loss
    wraps the difference into [-pi, pi) before 1-cos (main_training.py:184-190);
    `per_example_loss /= prod(shape(y_true)[1:])` then compute_average_loss = sum / global batch;
    scale_regularization_loss divides the L2 sum by the replica count.
    """
    a_t, p_t = y_true[..., 0], y_true[..., 1]
    a_p, p_p = y_pred[..., 0], y_pred[..., 1]
    amp = (a_t - a_p) ** 2
    d = (p_t * 2 * math.pi - math.pi) - (p_p * 2 * math.pi - math.pi)
    d = torch.remainder(d + math.pi, 2 * math.pi) - math.pi
    ph = 1 - torch.cos(d)
    per = alpha * amp + (1 - alpha) * ph
    per = per / float(np.prod(y_true.shape[1:]))
    loss = per.sum() / global_batch_size
    if l2_losses:
        loss = loss + sum(l2_losses) / num_replicas
    return loss


# ----------------------------------------------------------------------------------------
# Optimisers (Keras conventions)
# ----------------------------------------------------------------------------------------
def keras_adam_step(params, grads, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Adam: w -= lr*sqrt(1-b2^t)/(1-b1^t) * m/(sqrt(v)+eps), t = step (1-based)."""
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    for n in grads:
        g = grads[n]
        m[n].mul_(beta1).add_(g, alpha=1 - beta1)
        v[n].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        params[n].sub_(lr_t * m[n] / (v[n].sqrt() + eps))


def keras_nadam_step(params, grads, m, v, state, lr, beta1=0.9, beta2=0.999, eps=1e-7):
    """tf.keras.optimizers.Nadam (optimizer_v2/nadam.py, the "nadam" branch of amp_phase_trainer.py:30-31).
    state = {"step": completed steps, "m_schedule": running product of the momentum schedule (1.0 at the start)}."""
    t = state["step"] + 1
    u_t = beta1 * (1.0 - 0.5 * 0.96 ** (0.004 * t))
    u_t1 = beta1 * (1.0 - 0.5 * 0.96 ** (0.004 * (t + 1)))
    ms = state["m_schedule"] * u_t
    ms_next = ms * u_t1
    state["m_schedule"], state["step"] = ms, t
    for n in grads:
        g = grads[n]
        g_prime = g / (1.0 - ms)
        m[n].mul_(beta1).add_(g, alpha=1 - beta1)
        v[n].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        m_prime = m[n] / (1.0 - ms_next)
        v_prime = v[n] / (1.0 - beta2 ** t)
        m_bar = (1.0 - u_t) * g_prime + u_t1 * m_prime
        params[n].sub_(lr * m_bar / (v_prime.sqrt() + eps))


def tfa_lamb_step(params, grads, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
    """tensorflow_addons.optimizers.LAMB (trainer.py:37-38), step is 1-based: Adam direction with bias correction,
    per-variable trust ratio ||w|| / ||update|| (1 when either norm is zero)."""
    for n in grads:
        g = grads[n]
        m[n].mul_(beta1).add_(g, alpha=1 - beta1)
        v[n].mul_(beta2).addcmul_(g, g, value=1 - beta2)
        upd = (m[n] / (1 - beta1 ** step)) / ((v[n] / (1 - beta2 ** step)).sqrt() + eps) + weight_decay * params[n]
        wn, un = float(params[n].norm()), float(upd.norm())
        ratio = wn / un if (wn > 0 and un > 0) else 1.0
        params[n].sub_(lr * ratio * upd)


def keras_sgd_step(params, grads, lr):
    for n in grads:
        params[n].sub_(lr * grads[n])


# ----------------------------------------------------------------------------------------
# One training step, for the CPU baseline and whole-step parity
# ----------------------------------------------------------------------------------------
def train_step(model: UNetOracle, params, opt_state, spec_in, spec_out, emb, lr,
               dropout_mask=None, loss_kind="amp_phase", alpha=0.9, global_batch=None,
               num_replicas=1, apply=True):
    """Trainer.step (amp_phase_trainer.py:130-141) or the DP train_step
    (main_training.py:253-268). Returns (losses tuple, grads dict)."""
    names = trainable_names(model.plan)
    leaf = OrderedDict()
    for n, t in params.items():
        leaf[n] = t.detach().clone().requires_grad_(n in names)
    new_stats = {}
    y = model.forward(leaf, spec_in, emb, training=True, dropout_mask=dropout_mask,
                      new_stats=new_stats)
    if loss_kind == "amp_phase":
        loss, lp, ls = amp_phase_loss(spec_out, y)
    else:
        gb = global_batch or spec_in.shape[0]
        loss = dp_loss(spec_out, y, alpha, gb, model.l2_losses(leaf), num_replicas)
        with torch.no_grad():
            _, lp, ls = amp_phase_loss(spec_out, y)
    gl = torch.autograd.grad(loss, [leaf[n] for n in names], allow_unused=True)
    grads = OrderedDict()
    for n, g in zip(names, gl):
        grads[n] = g if g is not None else torch.zeros_like(params[n])
    if apply:
        opt_state["step"] += 1
        with torch.no_grad():
            keras_adam_step(params, grads, opt_state["m"], opt_state["v"], opt_state["step"], lr)
            for n, t in new_stats.items():
                params[n].copy_(t)
    return (loss.detach(), lp.detach(), ls.detach()), grads, y.detach()


def new_opt_state(params, plan):
    names = trainable_names(plan)
    return {"step": 0,
            "m": OrderedDict((n, torch.zeros_like(params[n])) for n in names),
            "v": OrderedDict((n, torch.zeros_like(params[n])) for n in names)}
