#!/usr/bin/env python
"""Run this WHERE TENSORFLOW AND THE REFERENCE ARE INSTALLED (not on the B200 image, which has neither TF nor h5py).

    python tools/export_tf_weights.py --reference /path/to/unet-rir --out unet_weights.npz \
        [--h5 path/to/weights.h5 | --ckpt-dir ../results/unet] [--kernels 3] [--filters 32] [--mode 0]

It rebuilds the reference's UNet exactly as its call sites do (main_training.py:155-161, rir_generation.py:117-123),
restores either a Keras `weights.h5` (UNet.load_weights, dl_models/u_net.py:192-199) or the latest tf.train.Checkpoint
of a directory (main_training.py:171-172 / rir_generation.py:125-133), and writes every variable of
`model.model.weights` -- the 77 trainable tensors and the 26 BatchNormalization moving statistics -- with its Keras
name into ONE .npz:   names = array of variable names,  arr_<i> = value of names[i].
`unet_rir_b200.UNet.load_weights("unet_weights.npz")` reads that file (unet_rir_b200/keras_weights.py maps the
variables onto the engine's plan by layer class and creation index; no layout change is needed).
"""
import argparse
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True, help="checkout of igmsalinas/unet-rir")
    ap.add_argument("--out", required=True)
    ap.add_argument("--h5", default=None)
    ap.add_argument("--ckpt-dir", default=None)
    ap.add_argument("--kernels", type=int, default=3)
    ap.add_argument("--filters", type=int, default=32)
    ap.add_argument("--mode", type=int, default=0)
    ap.add_argument("--height", type=int, default=144)
    ap.add_argument("--width", type=int, default=160)
    args = ap.parse_args()

    sys.path.insert(0, args.reference)
    import tensorflow as tf
    from dl_models.u_net import UNet

    unet = UNet(input_shape=(args.height, args.width, 2), inf_vector_shape=(2, 16), mode=args.mode,
                number_filters_0=args.filters, kernels=args.kernels, name="U-Net")
    if args.h5:
        unet.load_weights(args.h5)
    elif args.ckpt_dir:
        ckpt = tf.train.Checkpoint(model=unet.model)
        latest = tf.train.latest_checkpoint(args.ckpt_dir)
        if latest is None:
            raise SystemExit(f"no checkpoint under {args.ckpt_dir}")
        ckpt.restore(latest).expect_partial()
    else:
        print("no --h5 / --ckpt-dir: exporting the freshly initialised model (useful for parity runs)")
    weights = unet.model.weights
    names = np.array([w.name for w in weights])
    arrays = {f"arr_{i}": w.numpy() for i, w in enumerate(weights)}
    np.savez(args.out, names=names, **arrays)
    n = sum(int(np.prod(a.shape)) for a in arrays.values())
    print(f"wrote {len(names)} variables, {n:,} values to {args.out}")


if __name__ == "__main__":
    main()
