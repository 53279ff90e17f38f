"""Generates tests/golden/*.npz / *.json from the REFERENCE's own code, run in the build container.

TensorFlow / librosa are not installed, so the reference's modules cannot be imported; but a few of its
classes are pure numpy / pure Python. This script lifts exactly those definitions out of the reference
sources with `ast` (no reference source is copied into the repo), executes them, and records their
outputs on seeded inputs:
  * preprocess.py : Normalizer.normalize / denormalize, TensorPadder.pad_amp_phase / un_pad, sigmoid
  * amp_phase_trainer.py : ModelCheckpoint.checkpoint / EarlyStopping.stop_count decision traces
  * trainer.py : the generic trainer's ModelCheckpoint(min_delta) / EarlyStopping decision trace
The U-Net graph (Keras), the STFT (librosa) and the losses (tf ops) cannot be executed here; for those the
oracle is a restatement (parity unpinned, see oracle/*.py headers).

Run:  python tests/golden/make_golden.py   (needs /root/reference; the outputs are committed)
"""
import ast
import json
import math
import os

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def lift(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in names]
    ns = {"np": np, "math": math}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    rng = np.random.default_rng(500)
    pre = lift(os.path.join(REF, "preprocess.py"), {"Normalizer", "TensorPadder", "sigmoid"})
    amp = np.abs(rng.standard_normal((129, 151))).astype(np.float32) * np.exp(rng.uniform(-12, 3, (129, 151))).astype(np.float32)
    amp[0, 0] = 0.0                                      # exact zero -> the -100 dB floor
    phase = rng.uniform(-math.pi, math.pi, (129, 151)).astype(np.float32)
    nz = pre["Normalizer"]()
    a_n, p_n = nz.normalize(amp, phase)
    a_d, p_d = nz.denormalize(a_n, p_n)
    pad = pre["TensorPadder"]((144, 160))
    a_p, p_p = pad.pad_amp_phase(a_n, p_n)
    a_u, p_u = pre["TensorPadder"].un_pad(a_p, p_p, (129, 151))
    big = rng.standard_normal((150, 10))
    big_out = pre["TensorPadder"]((144, 160)).transform(big)     # larger than desired in dim 0 -> unchanged
    sig = pre["sigmoid"](0.5, (144, 160))
    np.savez_compressed(os.path.join(OUT, "preprocess_golden.npz"), amp=amp, phase=phase, amp_norm=a_n, phase_norm=p_n,
                        amp_denorm=a_d, phase_denorm=p_d, amp_pad=a_p, phase_pad=p_p, amp_unpad=a_u, phase_unpad=p_u,
                        big=big, big_out=big_out, sigmoid=sig.astype(np.float32))

    cb = lift(os.path.join(REF, "amp_phase_trainer.py"), {"ModelCheckpoint", "EarlyStopping"})

    class FakeModel:
        def __init__(self): self.saved = 0
        def save(self, path): self.saved += 1

    val = [12.0, 9.5, 9.7, 9.1, 9.1, 9.3, 9.2, 9.4, 8.0, 8.1, 8.2, 8.3]
    trn = [11.0, 9.0, 8.0, 7.0, 6.5, 6.0, 5.5, 5.0, 4.5, 4.0, 3.5, 3.0]
    mc, es, fm = cb["ModelCheckpoint"]("ckpt", True, 0), cb["EarlyStopping"](3), FakeModel()
    trace = []
    for t, v in zip(trn, val):
        imp = mc.checkpoint(train_loss=t, val_loss=v, model=fm)
        stop = es.stop_count(improve=imp)
        trace.append({"train": t, "val": v, "improve": bool(imp), "stop": bool(stop), "count": es.count,
                      "val_min": mc.val_loss_min, "train_min": mc.train_loss_min, "saved": fm.saved})
    # trainer.py: the generic trainer's ModelCheckpoint(min_delta) (:175-205) on a trace with sub-min_delta improvements
    cb2 = lift(os.path.join(REF, "trainer.py"), {"ModelCheckpoint", "EarlyStopping"})
    val2 = [12.0, 9.5, 9.49995, 9.4999, 9.3, 9.29995, 9.4, 9.29991, 9.2998, 9.0]
    mc2, es2, fm2 = cb2["ModelCheckpoint"]("ckpt", True, 0), cb2["EarlyStopping"](3), FakeModel()
    trace2 = []
    for i, v in enumerate(val2):
        imp = mc2.checkpoint(train_loss=float(10 - i), val_loss=v, model=fm2)
        stop = es2.stop_count(improve=imp)
        trace2.append({"train": float(10 - i), "val": v, "improve": bool(imp), "stop": bool(stop), "count": es2.count,
                       "val_min": mc2.val_loss_min, "train_min": mc2.train_loss_min, "saved": fm2.saved})
    json.dump({"patience": 3, "trace": trace, "generic_trainer": {"patience": 3, "min_delta": 0.0001, "trace": trace2}},
              open(os.path.join(OUT, "callbacks_golden.json"), "w"), indent=1)
    # rooms.py: the 16-int embedding of every (room, zone, array type, loudspeaker, microphone) combination sampled
    rm = lift(os.path.join(REF, "rooms.py"), {"Quadrilateral", "Room", "UTSRoom", "return_room"})
    rooms = {"AnechoicRoom": (490, 722, 490, 722, 90, 90, 90, 90, 529, [245, 361], 45),
             "HemiAnechoicRoom": (490, 722, 490, 722, 90, 90, 90, 90, 529, [245, 361], 52),
             "SmallMeetingRoom": (355, 410, 401, 378, 96, 90, 85, 88, 300, [175.5, 205], 497),
             "MediumMeetingRoom": (736, 520, 650, 434.5, 81, 92, 98, 89, 300, [368, 217.5], 659),
             "LargeMeetingRoom": (994, 923, 1087, 1022, 81.4, 105, 81.3, 92.3, 300, [497, 486.25], 1281),
             "ShoeBoxRoom": (600, 1175, 600, 1175, 90, 90, 90, 90, 300, [300, 881.25], 667)}      # dataset.py:84-89
    cases = []
    for name, args in rooms.items():
        room = rm["UTSRoom"](*args)
        for zone in "ABCDE":
            for array in ("Planar", "Circular"):
                for l in (1, 7, 22, 60):
                    for m in (1, 8, 9, 30, 31, 64):
                        ch = [name, zone, array, str(l), str(m)]
                        cases.append({"characteristics": ch, "embedding": [float(v) for v in room.return_embedding(ch)]})
    json.dump({"rooms": {k: list(v[:9]) + [v[9], v[10]] for k, v in rooms.items()}, "cases": cases},
              open(os.path.join(OUT, "rooms_golden.json"), "w"))
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
