"""Import of weights trained with the reference (TensorFlow / Keras) into the sm_100a engine.

The reference saves `weights.h5` (UNet.save, dl_models/u_net.py:192-199) or a tf.train.Checkpoint
(main_training.py:171-172, restored by rir_generation.py:125-133). Neither h5py nor TensorFlow exists on the B200
image, so the import is two steps:

  1. where TensorFlow is installed, `tools/export_tf_weights.py` rebuilds the reference model, restores the
     checkpoint / h5 file and writes every variable of `model.model.weights` with its Keras name into one `.npz`
     (`names` + `arr_<i>`; nothing but numpy is needed to read it);
  2. here, `read_keras_npz` maps those variables onto the engine's plan (plan.py) and `UNet.load_weights("x.npz")`
     loads them.

The mapping is BY LAYER CLASS AND CREATION INDEX, not by position: a functional Keras model lists its layers in
topological order (the embedding branch is interleaved with the deepest encoder layers), and Keras' auto-name counters
(`conv2d_7`, `batch_normalization_3`, ...) are offset when other models were built earlier in the process. Within one
class the auto-name index IS creation order, which is the order UNet._build creates layers in (u_net.py:201-251) and
therefore the order of plan.layer_plan. Layouts need no transposition: Conv2D kernels are HWIO, Conv2DTranspose
kernels (kh, kw, out, in), Dense (in, out), Embedding (vocab, dim) on both sides.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np
import torch

# plan kind -> (Keras layer class auto-name, Keras variable name)
_KERAS_VAR = {"conv_w": "kernel", "convT_w": "kernel", "dense_w": "kernel", "bias": "bias", "gamma": "gamma",
              "beta": "beta", "moving_mean": "moving_mean", "moving_var": "moving_variance", "emb": "embeddings"}
_CLASS_ALIAS = {"encoder_inf_dense": "dense"}       # the Dense layer carries an explicit name (u_net.py:259)
_NAME_RE = re.compile(r"^(?:.*/)?(?P<layer>[A-Za-z_0-9]+?)(?:_(?P<idx>\d+))?/(?P<var>[a-z_]+)(?::\d+)?$")


def plan_layers(plan):
    """Groups the plan's variables into layers in creation order: [(layer prefix, keras class, [(name, shape, kind)])]."""
    layers, seen = [], {}
    for name, shape, kind in plan:
        prefix = name.rsplit(".", 1)[0]
        if name == "vec.emb":
            prefix = "vec.emb"
        if prefix not in seen:
            if kind == "conv_w":
                cls = "conv2d"
            elif kind == "convT_w":
                cls = "conv2d_transpose"
            elif kind == "dense_w":
                cls = "dense"
            elif kind == "emb":
                cls = "embedding"
            elif kind in ("gamma", "beta", "moving_mean", "moving_var"):
                cls = "batch_normalization"
            else:
                raise ValueError(f"{name}: a layer cannot start with a {kind}")
            seen[prefix] = len(layers)
            layers.append((prefix, cls, []))
        layers[seen[prefix]][2].append((name, tuple(shape), kind))
    return layers


def keras_names_for_plan(plan, offsets=None):
    """The Keras variable names a fresh TF process gives the reference model, in plan order: [(plan name, keras name)].
    offsets: optional {class: first index} to emulate counters that did not start at zero."""
    counters = dict(offsets or {})
    out = []
    for prefix, cls, variables in plan_layers(plan):
        i = counters.get(cls, 0)
        counters[cls] = i + 1
        lname = "encoder_inf_dense" if cls == "dense" else (cls if i == 0 else f"{cls}_{i}")
        for name, _, kind in variables:
            out.append((name, f"{lname}/{_KERAS_VAR[kind]}:0"))
    return out


def map_keras_variables(names, arrays, plan):
    """names[i] = Keras variable name of arrays[i] (any order) -> OrderedDict plan name -> float32 tensor.
    Raises ValueError on a missing / surplus layer or a shape mismatch (which is what a wrong `kernels`,
    `number_filters_0` or `mode` looks like)."""
    by_class = {}
    for n, a in zip(names, arrays):
        m = _NAME_RE.match(str(n))
        if m is None:
            raise ValueError(f"cannot parse the Keras variable name {n!r}")
        layer, idx, var = m.group("layer"), int(m.group("idx") or 0), m.group("var")
        layer = _CLASS_ALIAS.get(layer, layer)
        by_class.setdefault(layer, {}).setdefault(idx, {})[var] = np.asarray(a)
    queues = {cls: [layers[i] for i in sorted(layers)] for cls, layers in by_class.items()}
    state = OrderedDict()
    for prefix, cls, variables in plan_layers(plan):
        if not queues.get(cls):
            raise ValueError(f"the file has no more '{cls}' layers but the model still needs one for {prefix}")
        kv = queues[cls].pop(0)
        for name, shape, kind in variables:
            var = _KERAS_VAR[kind]
            if var not in kv:
                raise ValueError(f"{prefix}: Keras layer of class {cls} has no variable '{var}' (has {sorted(kv)})")
            a = kv[var]
            if tuple(a.shape) != shape:
                raise ValueError(f"{name}: file has shape {tuple(a.shape)}, the model needs {shape}")
            state[name] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))
    left = {cls: len(q) for cls, q in queues.items() if q}
    if left:
        raise ValueError(f"the file holds layers the model does not have: {left}")
    return state


def read_keras_npz(path, plan):
    """`.npz` written by tools/export_tf_weights.py -> state dict for UNetEngine.load_state_dict."""
    with np.load(path, allow_pickle=False) as z:
        names = [str(n) for n in z["names"]]
        arrays = [z[f"arr_{i}"] for i in range(len(names))]
    return map_keras_variables(names, arrays, plan)
