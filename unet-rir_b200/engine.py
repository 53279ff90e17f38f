"""Device-side execution of the U-Net hot path through liburir's C-ABI.

This is the B200 replacement for what TensorFlow does under `model.model([spec, emb],
training=...)`, `tape.gradient` and `optimizer.apply_gradients` in the reference
(amp_phase_trainer.py:130-141, main_training.py:253-268). Python only sequences the calls and
owns the buffers (torch tensors as device memory); every arithmetic op is a liburir kernel.

Data layout in HBM (all NHWC):
  * activations bf16. Each encoder output e_i is written by its BN+ReLU pass straight into the
    left half of the decoder's concat buffer cat_i = [e_i | up_i] and the Conv2DTranspose writes
    the right half, so `concatenate` (u_net.py:308) never exists as a copy.
  * raw (pre-BN) conv outputs bf16 are kept for the backward pass; BN batch statistics come out
    of the conv epilogue (fp32 sum / sum-of-squares per channel).
  * parameters: ONE flat fp32 buffer in Keras variable order (plan.py), with flat gradient and
    Adam-moment twins, so the optimiser is a single kernel and the DP all-reduce works on
    contiguous buckets. bf16 operand copies of every conv kernel ([tap][C][K] and [tap][K][C])
    and of the Dense kernel are refreshed after each optimiser step.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
from collections import OrderedDict

import numpy as np
import torch

from . import _lib as L
from . import plan as PL


class View:
    """A C-channel slice, starting at channel `coff`, of an NHWC buffer with `ld` channels."""
    __slots__ = ("buf", "coff", "C")

    def __init__(self, buf, coff=0, C=None):
        self.buf, self.coff = buf, coff
        self.C = buf.shape[3] - coff if C is None else C

    N = property(lambda s: s.buf.shape[0])
    H = property(lambda s: s.buf.shape[1])
    W = property(lambda s: s.buf.shape[2])
    ld = property(lambda s: s.buf.shape[3])
    npix = property(lambda s: s.buf.shape[0] * s.buf.shape[1] * s.buf.shape[2])

    def ptr(self):
        return self.buf.data_ptr()

    def tensor(self):
        return self.buf[..., self.coff:self.coff + self.C]


def _align(n, a=4):
    return (n + a - 1) // a * a


class InputPrefetcher:
    """Double-buffered host -> device staging of a batch on a copy stream, so the H2D copy of batch i+1 overlaps the
    train step of batch i (the reference's tf.data pipeline does the same for MirroredStrategy,
    main_training.py:114). `prefetch` enqueues the copies; `take` hands the staged device tensors to the step that
    is called with the SAME host objects (identity match), after making the compute stream wait for the copy."""

    def __init__(self, eng):
        self.eng, self.stream, self.sets, self.pending, self.k = eng, None, {}, None, 0

    def _set(self, B, k):
        key = (B, k)
        if key not in self.sets:
            H, W, Cin = self.eng.input_shape
            dev = self.eng.device
            self.sets[key] = {"x": torch.empty(B, H, W, Cin, dtype=torch.float32, device=dev),
                              "y": torch.empty(B, H, W, 2, dtype=torch.float32, device=dev),
                              "emb": torch.empty(B, self.eng.T, dtype=torch.int32, device=dev),
                              "ready": None, "consumed": None}
        return self.sets[key]

    @staticmethod
    def _host(t, dtype):
        t = t if isinstance(t, torch.Tensor) else torch.as_tensor(np.asarray(t))
        return t if t.dtype == dtype else t.to(dtype)

    def prefetch(self, spec_in, emb, spec_out):
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=self.eng.device)
        B = int(spec_in.shape[0])
        s = self._set(B, self.k)
        self.k ^= 1
        if s["consumed"] is not None:
            self.stream.wait_event(s["consumed"])          # the step that last read this set has copied it out
        with torch.cuda.stream(self.stream):
            s["x"].copy_(self._host(spec_in, torch.float32).reshape(s["x"].shape), non_blocking=True)
            s["y"].copy_(self._host(spec_out, torch.float32).reshape(s["y"].shape), non_blocking=True)
            s["emb"].copy_(self._host(emb, torch.int32).reshape(s["emb"].shape), non_blocking=True)
            s["ready"] = torch.cuda.Event()
            s["ready"].record(self.stream)
        self.pending = (id(spec_in), id(emb), id(spec_out), s)

    def take(self, spec_in, emb, spec_out):
        """Stages a prefetched batch into the engine's static input buffers; False when nothing matches."""
        pd = self.pending
        if pd is None or pd[:3] != (id(spec_in), id(emb), id(spec_out)):
            return False
        self.pending = None
        s = pd[3]
        cur = torch.cuda.current_stream()
        cur.wait_event(s["ready"])
        self.eng.stage(s["x"], s["emb"], s["y"])           # device -> device into the buffers the CUDA graph reads
        s["consumed"] = torch.cuda.Event()
        s["consumed"].record(cur)
        return True


class UNetEngine:
    def __init__(self, input_shape=(144, 160, 2), inf_vector_shape=(2, 16), mode=0,
                 number_filters_0=32, kernels=6, BatchNorm=True, device="cuda", seed=500,
                 impl=L.IMPL_AUTO, bn_unbiased_moving_var=False, arch="unet"):
        if mode not in (0, 1, 2, 3):
            raise ValueError("mode must be 0 (convolutional_block_1), 1 (convolutional_block_2), 2 (residual_block_1) "
                             "or 3 (residual_block_2)  (u_net.py:280-287)")
        H, W, Cin = input_shape
        if H % 16 or W % 16:
            raise ValueError("input H and W must be multiples of 16 (four stride-2 stages)")
        L.load()
        self.input_shape, self.inf_vector_shape = tuple(input_shape), tuple(inf_vector_shape)
        self.mode, self.F0, self.kernels, self.BatchNorm = mode, number_filters_0, kernels, BatchNorm
        self.device = torch.device(device)
        self.impl = impl
        self.bn_unbiased = int(bn_unbiased_moving_var)
        # arch "unet" = dl_models/u_net.py, "diff" = dl_models/diff_u_net.py (plan.ARCHS lists what differs)
        self.arch = arch
        self.A = PL.arch_params(arch, number_filters_0, kernels)
        self.head_sigmoid = bool(self.A["head_sigmoid"])
        self.plan = PL.layer_plan(input_shape, inf_vector_shape, mode, number_filters_0, kernels, BatchNorm, arch)
        self.T = inf_vector_shape[0] * inf_vector_shape[1]
        self.H5, self.W5 = H // 16, W // 16
        self.dense_n = self.H5 * self.W5 * self.A["vec_ch"]
        self._build_params(seed)
        self._bufs = {}
        self.dropout_seed = seed
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        # Dropout's own stream position: advanced by every training forward that draws a mask (a device-side
        # increment right after the mask launch, so CUDA-graph replays, optimisers that never touch step_dev
        # (Nadam, an external torch optimiser) and repeated forwards all get fresh masks)
        self.drop_ctr_dev = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.lr_dev = torch.zeros(1, dtype=torch.float32, device=self.device)
        self.losses_dev = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.reg_dev = torch.zeros(2, dtype=torch.float32, device=self.device)    # [whole / decoder part, encoder part]
        self.side = torch.cuda.Stream(device=self.device)
        self.prefetcher = InputPrefetcher(self)
        self.eval_cuda_graph = os.environ.get("URIR_NO_EVAL_GRAPH", "0") != "1"
        self._eval_graphs = {}
        self.fold_bn_eval = self.BatchNorm and os.environ.get("URIR_NO_BN_FOLD", "0") != "1"
        self.overlap_wgrad = os.environ.get("URIR_NO_OVERLAP", "0") != "1"
        self._side_dirty = False

    # ------------------------------------------------------------------ parameters
    def _build_params(self, seed):
        dev = self.device
        off, soff = 0, 0
        self.offsets, self.state_offsets, self.shapes, self.kinds = OrderedDict(), OrderedDict(), {}, {}
        for name, shape, kind in self.plan:
            n = 1
            for s in shape:
                n *= s
            self.shapes[name], self.kinds[name] = tuple(shape), kind
            if kind in PL.TRAINABLE_KINDS:
                self.offsets[name] = (off, n); off = _align(off + n)
            else:
                self.state_offsets[name] = (soff, n); soff = _align(soff + n)
        self.n_flat = off
        self.P = torch.zeros(off, dtype=torch.float32, device=dev)
        self.G = torch.zeros(off, dtype=torch.float32, device=dev)
        self.M = torch.zeros(off, dtype=torch.float32, device=dev)
        self.V = torch.zeros(off, dtype=torch.float32, device=dev)
        self.S = torch.zeros(max(soff, 4), dtype=torch.float32, device=dev)
        self.param, self.grad, self.state = OrderedDict(), OrderedDict(), OrderedDict()
        for name, (o, n) in self.offsets.items():
            self.param[name] = self.P[o:o + n].view(self.shapes[name])
            self.grad[name] = self.G[o:o + n].view(self.shapes[name])
        for name, (o, n) in self.state_offsets.items():
            self.state[name] = self.S[o:o + n].view(self.shapes[name])
        # bf16 operand copies
        self.wops = {}
        for name, shape, kind in self.plan:
            if kind in ("conv_w", "convT_w"):
                kh, kw, a, b = shape
                self.wops[name] = (torch.empty(kh * kw, a, b, dtype=torch.bfloat16, device=dev),
                                   torch.empty(kh * kw, b, a, dtype=torch.bfloat16, device=dev))
        kd, dn = self.shapes["vec.dense.w"]
        self.dense_w16 = torch.empty(kd, dn, dtype=torch.bfloat16, device=dev)       # [Kd][N]
        self.dense_w16_t = torch.empty(dn, kd, dtype=torch.bfloat16, device=dev)     # [N][Kd]
        rows = [[self.param["vec.dense.w"].data_ptr(), self.dense_w16.data_ptr(), self.dense_w16_t.data_ptr(), 1, kd, dn]]
        for name, (w_ck, w_kc) in self.wops.items():
            taps, a, b = w_ck.shape
            rows.append([self.param[name].data_ptr(), w_ck.data_ptr(), w_kc.data_ptr(), taps, a, b])
        self.wprep_table = torch.tensor(rows, dtype=torch.int64, device=dev)
        # w_up2 operand of the stride-2 layers (3x3 only): Conv2DTranspose forward and strided-conv dgrad as one
        # 2x2 problem on the half-resolution grid (urir_conv2d_dgrad_up2); allocated where the kernel applies
        self.wup2 = {}
        if self.A["down_k"] == 3 and self.A["up_k"] == 3:
            H, W, _ = self.input_shape
            for name, shape, kind in self.plan:
                if kind == "convT_w" or (kind == "conv_w" and name.endswith(".down.w") and not name.startswith("enc1.")):
                    lvl = int(name[3]) if name.startswith("enc") else 7 - int(name[3])     # level of the LARGE tensor + 1
                    h, w = H >> (lvl - 2), W >> (lvl - 2)
                    kh, kw, c, k = shape
                    d = L.ConvDesc(1, h, w, c, k, 3, 3, 2, 0, 0, h // 2, w // 2, c, 0, k, 0, L.BF16, L.BF16, self.impl, 0, 0)
                    if h % 2 == 0 and w % 2 == 0 and L.load().urir_conv_path(C.byref(d), 3) == 1:
                        self.wup2[name[:-2]] = torch.empty(4, 4 * c, k, dtype=torch.bfloat16, device=dev)
        # per-BN scratch: stats (zeroed each forward), scale/shift, mean/rstd, backward sums
        bn_names = [n[:-len(".gamma")] for n in self.offsets if n.endswith(".gamma")]
        tot = sum(2 * self.shapes[b + ".gamma"][0] for b in bn_names)
        self.stats_arena = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.ss_arena = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.mr_arena = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.sums_arena = torch.zeros(tot, dtype=torch.float32, device=dev)
        self.bn_slot = {}
        o = 0
        for b in bn_names:
            c = self.shapes[b + ".gamma"][0]
            self.bn_slot[b] = (o, 2 * c); o += 2 * c
        # BatchNorm=False (u_net.py:367 `if BatchNorm:`): conv -> ReLU runs through the same kernels with identity
        # coefficients (scale 1, shift 0, mean 0, rstd 1, gamma 1, zero reduction sums), one set per channel count
        self._ident = {}
        # bias-gradient statistics emitted by dgrad epilogues (zeroed each backward)
        self.bstat_arena = torch.zeros(2 * 4 * (self.F0 * 16) * 12, dtype=torch.float32, device=dev)
        self.load_state_dict(PL.keras_init(self.plan, seed))

    def trainable_names(self):
        return list(self.offsets)

    def state_dict(self):
        d = OrderedDict((k, v.detach().cpu().clone()) for k, v in self.param.items())
        d.update((k, v.detach().cpu().clone()) for k, v in self.state.items())
        return d

    def load_state_dict(self, sd):
        for k, v in sd.items():
            dst = self.param.get(k, None)
            if dst is None:
                dst = self.state[k]
            if tuple(v.shape) != tuple(dst.shape):
                raise ValueError(f"{k}: shape {tuple(v.shape)} != {tuple(dst.shape)}")
            with torch.no_grad():           # the views may be autograd leaves (UNetModel.trainable_variables)
                dst.copy_(torch.as_tensor(v, dtype=torch.float32))
        self.refresh_operands()

    def refresh_operands(self):
        """fp32 masters -> bf16 operand layouts (after load / after every optimiser step)."""
        self._p_version = self.P._version
        L.call("weight_prep_batched", self.wprep_table.data_ptr(), self.wprep_table.shape[0])
        for name, buf in self.wup2.items():
            L.call("weight_prep_up2", self.param[name + ".w"].data_ptr(), buf.data_ptr(), buf.shape[1] // 4, buf.shape[2])

    def sync_operands(self):
        """Refreshes the bf16 operand copies when the fp32 masters were modified through torch (an external
        optimiser stepping `trainable_variables`, a manual `param[...].copy_`): views share P's version counter.
        Kernel-side updates (adam_step etc.) refresh the operands themselves."""
        if self.P._version != getattr(self, "_p_version", -1):
            self.refresh_operands()

    # ------------------------------------------------------------------ buffers
    def _buffers(self, B):
        if B in self._bufs:
            return self._bufs[B]
        dev, bf = self.device, torch.bfloat16
        H, W, Cin = self.input_shape
        b = {}

        def act(h, w, c, dtype=bf):
            return torch.zeros(B, h, w, c, dtype=dtype, device=dev)

        b["x_in"] = act(H, W, Cin, torch.float32)
        b["emb"] = torch.zeros(B, self.T, dtype=torch.int32, device=dev)
        b["out"] = act(H, W, 2, torch.float32)
        b["g_out"] = act(H, W, 2, torch.float32)
        b["y_true"] = act(H, W, 2, torch.float32)
        for i in range(1, 6):
            n = self.F0 * 2 ** (i - 1)
            h, w = (H, W) if i == 1 else (H >> (i - 1), W >> (i - 1))
            for nm in ("t", "r"):
                b[f"{nm}{i}"] = act(h, w, n); b[f"g_{nm}{i}"] = act(h, w, n)
            if i < 5:
                b[f"cat{i}"] = act(h, w, 2 * n); b[f"g_cat{i}"] = act(h, w, 2 * n)
                for nm in ("rf", "f", "rb", "d"):
                    b[f"{nm}{i}"] = act(h, w, n); b[f"g_{nm}{i}"] = act(h, w, n)
            else:
                b["z"] = act(h, w, n); b["g_z"] = act(h, w, n)
        if self.mode != 0:
            # block modes 1-3 (u_net.py:324-386): activation after c1, raw c2 output, (2/3) activation after c2,
            # (3) raw / activated shortcut conv -- and the matching gradients, per block
            for i in range(1, 6):
                n = self.F0 * 2 ** (i - 1)
                h, w = (H, W) if i == 1 else (H >> (i - 1), W >> (i - 1))
                for key in ([f"enc{i}"] + ([f"dec{6 - i}"] if i < 5 else [])):
                    names = ["a1", "raw2"] + (["a2"] if self.mode >= 2 else []) + (["raw3", "a3"] if self.mode == 3 else [])
                    for nm in names:
                        b[f"{key}.{nm}"] = act(h, w, n)
                        if nm != "a2" and nm != "a3":
                            b[f"g_{key}.{nm}"] = act(h, w, n)
        b["embflat"] = torch.zeros(B, self.T * self.A["emb_dim"], dtype=bf, device=dev)
        b["g_embflat"] = torch.zeros(B, self.T * self.A["emb_dim"], dtype=bf, device=dev)
        b["v16"] = act(self.H5, self.W5, self.A["vec_ch"])
        # without the 1x1 projection (diff_u_net.py:261-270) the Dense output is added to e5 directly, so its gradient
        # IS the bottleneck gradient
        b["g_v16"] = act(self.H5, self.W5, self.A["vec_ch"]) if self.A["proj"] else b["g_z"]
        b["g_v16_eff"] = torch.zeros(B, self.dense_n, dtype=bf, device=dev)
        b["mask"] = torch.ones(B, self.dense_n, dtype=torch.float32, device=dev)
        self._bufs[B] = b
        return b

    # ------------------------------------------------------------------ op helpers
    def _desc(self, x: View, y: View, k, stride, act=L.ACT_NONE, accumulate=0):
        P, pt = L.same_pad(x.H, k, stride)
        Q, pl = L.same_pad(x.W, k, stride)
        assert (P, Q) == (y.H, y.W), ((x.H, x.W), (y.H, y.W), k, stride)
        return L.ConvDesc(x.N, x.H, x.W, x.C, y.C, k, k, stride, pt, pl, P, Q, x.ld, x.coff, y.ld, y.coff,
                          L.dtype_code(x.buf), L.dtype_code(y.buf), self.impl, act, accumulate)

    def _w(self, name):
        w_ck, w_kc = self.wops[name + ".w"]
        return w_ck.data_ptr(), w_kc.data_ptr()

    def _conv_fprop(self, name, x, y, k, stride, stats=None, act=L.ACT_NONE, accumulate=0, bias=True, w_kc=None,
                    bias_ptr=None):
        d = self._desc(x, y, k, stride, act, accumulate)
        w_ck, w_kc_default = self._w(name)
        if bias_ptr is None:
            bias_ptr = self.param[name + ".b"].data_ptr() if bias else None
        L.call("conv2d_fprop", C.byref(d), x.ptr(), w_ck, w_kc_default if w_kc is None else w_kc, bias_ptr, y.ptr(),
               None if stats is None else stats.data_ptr())

    def _conv_dgrad(self, name, dy, dx, k, stride, stats=None, accumulate=0, bias=False):
        """dx (the conv's input side) from dy (its output side); desc describes the forward conv."""
        d = self._desc(dx, dy, k, stride, L.ACT_NONE, accumulate)
        if stride == 2 and stats is None and name in self.wup2:
            L.call("conv2d_dgrad_up2", C.byref(d), dy.ptr(), self.wup2[name].data_ptr(),
                   self.param[name + ".b"].data_ptr() if bias else None, dx.ptr())
            return
        w_ck, w_kc = self._w(name)
        # every caller reads only the channel SUMS of the statistics (bias gradients): the sums-only entry point
        L.call("conv2d_dgrad" if stats is None else "conv2d_dgrad_sums", C.byref(d), dy.ptr(), w_ck, w_kc,
               self.param[name + ".b"].data_ptr() if bias else None, dx.ptr(),
               None if stats is None else stats.data_ptr())

    def _conv_wgrad(self, name, x, dy, k, stride):
        """Weight gradient on the side stream: it only feeds the optimiser, so it runs concurrently with the
        dgrad / BatchNorm chain of the main stream (fills the SMs the small deep-layer grids leave idle and
        overlaps tensor-bound with HBM-bound kernels). Joined by _join_side() at the end of a backward segment;
        captured into the CUDA graph as a parallel branch."""
        # accumulate = 1: the flat gradient buffer was zeroed once at the start of backward (no per-layer memset)
        d = self._desc(x, dy, k, stride, accumulate=1)
        with self._side_branch():
            L.call("conv2d_wgrad", C.byref(d), x.ptr(), dy.ptr(), self.grad[name + ".w"].data_ptr())

    @contextlib.contextmanager
    def _side_branch(self):
        """Work that only feeds the optimiser: ordered after everything issued so far on the current stream,
        run on the side stream (or inline when overlap is off), joined later by _join_side()."""
        if not self.overlap_wgrad:
            yield
            return
        self.side.wait_stream(torch.cuda.current_stream())
        self._side_dirty = True
        with torch.cuda.stream(self.side):
            yield

    def _join_side(self):
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self.side)
            self._side_dirty = False

    def _slot(self, arena, bn):
        o, n = self.bn_slot[bn]
        return arena[o:o + n]

    # ------------------------------------------------------------------ inference: BatchNorm folded into the convs
    @staticmethod
    def _conv_of_bn(bname):
        return bname.replace("fuse_bn", "fuse") if bname.endswith("fuse_bn") else bname.replace(".bn", ".c")

    def _fold_setup(self):
        """bf16 [tap][K][C] copies of every BN-followed kernel with the inference-mode BN scale folded in, folded fp32
        biases, and the device table urir_weight_fold_bn_batched reads."""
        dev = self.device
        self.wfold, self.bfold, rows = {}, {}, []
        for bname in self.bn_slot:
            cname = self._conv_of_bn(bname)
            kh, kw, cin, cout = self.shapes[cname + ".w"]
            self.wfold[cname] = torch.empty(kh * kw, cout, cin, dtype=torch.bfloat16, device=dev)
            self.bfold[cname] = torch.empty(cout, dtype=torch.float32, device=dev)
            rows.append([self.param[cname + ".w"].data_ptr(), self._slot(self.ss_arena, bname).data_ptr(),
                         self.param[cname + ".b"].data_ptr(), self.wfold[cname].data_ptr(), self.bfold[cname].data_ptr(),
                         kh * kw, cin, cout])
        self._fold_table = torch.tensor(rows, dtype=torch.int64, device=dev)

    def _fold_refresh(self):
        """moving statistics -> scale / shift of every BN layer, then ONE launch folds them into kernels and biases
        (re-done on every inference forward: the weights or statistics may have changed since the last one)."""
        if getattr(self, "_fold_table", None) is None:
            self._fold_setup()
        for bname in self.bn_slot:
            c = self.shapes[bname + ".gamma"][0]
            ss, mr = self._slot(self.ss_arena, bname), self._slot(self.mr_arena, bname)
            L.call("bn_finalize", None, 1.0, self.param[bname + ".gamma"].data_ptr(), self.param[bname + ".beta"].data_ptr(),
                   self.state[bname + ".moving_mean"].data_ptr(), self.state[bname + ".moving_var"].data_ptr(),
                   PL.BN_MOMENTUM, PL.BN_EPS, self.bn_unbiased, ss.data_ptr(), mr.data_ptr(), c)
        L.call("weight_fold_bn_batched", self._fold_table.data_ptr(), self._fold_table.shape[0])

    def _identity(self, c):
        if c not in self._ident:
            dev = self.device
            one, zero = torch.ones(c, device=dev), torch.zeros(c, device=dev)
            self._ident[c] = {"ss": torch.cat([one, zero]), "mr": torch.cat([zero, one]), "gamma": one.clone(),
                              "sums": torch.zeros(2 * c, device=dev)}
        return self._ident[c]

    def _cbr_fwd(self, cname, bname, x, raw, out, k, training):
        """Conv2D(k, SAME) -> BatchNormalization -> ReLU  (convolutional_block_1, u_net.py:363-371)."""
        if not self.BatchNorm:
            self._conv_fprop(cname, x, raw, k, 1)
            L.call("bn_relu_fwd", raw.ptr(), raw.ld, raw.coff, self._identity(raw.C)["ss"].data_ptr(), out.ptr(), out.ld,
                   out.coff, raw.npix, raw.C, 1)
            return
        if not training and self.fold_bn_eval:
            d = self._desc(x, out, k, 1, L.ACT_RELU)
            if L.load().urir_conv_path(C.byref(d), 0) == 1:          # tensor-core kernels carry the ReLU epilogue
                # conv -> BN(moving statistics) -> ReLU as ONE kernel: scale folded into the bf16 kernel, shift into the
                # bias (_fold_refresh), ReLU in the epilogue, output straight into `out` -- no raw tensor, no BN pass
                self._conv_fprop(cname, x, out, k, 1, act=L.ACT_RELU, w_kc=self.wfold[cname].data_ptr(),
                                 bias_ptr=self.bfold[cname].data_ptr())
                return
        stats = self._slot(self.stats_arena, bname) if training else None
        self._conv_fprop(cname, x, raw, k, 1, stats=stats)
        c = raw.C
        ss, mr = self._slot(self.ss_arena, bname), self._slot(self.mr_arena, bname)
        if training:
            L.call("bn_relu_fwd_train", raw.ptr(), raw.ld, raw.coff, stats.data_ptr(), float(raw.npix),
                   self.param[bname + ".gamma"].data_ptr(), self.param[bname + ".beta"].data_ptr(),
                   self.state[bname + ".moving_mean"].data_ptr(), self.state[bname + ".moving_var"].data_ptr(),
                   PL.BN_MOMENTUM, PL.BN_EPS, self.bn_unbiased, ss.data_ptr(), mr.data_ptr(),
                   out.ptr(), out.ld, out.coff, raw.npix, c)
            return
        L.call("bn_finalize", None if stats is None else stats.data_ptr(), float(raw.npix),
               self.param[bname + ".gamma"].data_ptr(), self.param[bname + ".beta"].data_ptr(),
               self.state[bname + ".moving_mean"].data_ptr(), self.state[bname + ".moving_var"].data_ptr(),
               PL.BN_MOMENTUM, PL.BN_EPS, self.bn_unbiased, ss.data_ptr(), mr.data_ptr(), c)
        L.call("bn_relu_fwd", raw.ptr(), raw.ld, raw.coff, ss.data_ptr(), out.ptr(), out.ld, out.coff,
               raw.npix, c, 1)

    def _cbr_bwd(self, cname, bname, x, raw, g_out, g_raw, k, g_x=None, g_x_stats=None, accumulate=0):
        if not self.BatchNorm:          # ReLU backward only: g_raw = g_out * (raw > 0); conv bias gradient = its pixel sum
            idn = self._identity(raw.C)
            L.call("bn_relu_bwd_apply", g_out.ptr(), g_out.ld, g_out.coff, raw.ptr(), raw.ld, raw.coff,
                   idn["ss"].data_ptr(), idn["mr"].data_ptr(), idn["gamma"].data_ptr(), idn["sums"].data_ptr(),
                   g_raw.ptr(), g_raw.ld, g_raw.coff, None, None, self.grad[cname + ".b"].data_ptr(), raw.npix, raw.C, 1)
            self._conv_wgrad(cname, x, g_raw, k, 1)
            if g_x is not None:
                self._conv_dgrad(cname, g_raw, g_x, k, 1, stats=g_x_stats, accumulate=accumulate)
            return
        ss, mr = self._slot(self.ss_arena, bname), self._slot(self.mr_arena, bname)
        sums = self._slot(self.sums_arena, bname)
        c = raw.C
        L.call("bn_relu_bwd_reduce", g_out.ptr(), g_out.ld, g_out.coff, raw.ptr(), raw.ld, raw.coff,
               ss.data_ptr(), mr.data_ptr(), sums.data_ptr(), raw.npix, c, 1)
        L.call("bn_relu_bwd_apply", g_out.ptr(), g_out.ld, g_out.coff, raw.ptr(), raw.ld, raw.coff,
               ss.data_ptr(), mr.data_ptr(), self.param[bname + ".gamma"].data_ptr(), sums.data_ptr(),
               g_raw.ptr(), g_raw.ld, g_raw.coff, self.grad[bname + ".gamma"].data_ptr(),
               self.grad[bname + ".beta"].data_ptr(), self.grad[cname + ".b"].data_ptr(), raw.npix, c, 1)
        self._conv_wgrad(cname, x, g_raw, k, 1)
        if g_x is not None:
            self._conv_dgrad(cname, g_raw, g_x, k, 1, stats=g_x_stats, accumulate=accumulate)

    def _add(self, a, bv, out):
        L.call("add_bf16_strided", a.ptr(), a.ld, a.coff, bv.ptr(), bv.ld, bv.coff, out.ptr(), out.ld, out.coff, a.npix, a.C)

    def _blk_fwd(self, key, blk, x, raw1, out, training):
        """The mode-selected feature block (u_net.py:280-287, 314-319) on block input x -> out.
        mode 0 convolutional_block_1, 1 convolutional_block_2, 2 residual_block_1 (+ x), 3 residual_block_2 (+ conv shortcut)."""
        m = self.mode
        if m == 0:
            self._cbr_fwd(blk + ".c1", blk + ".bn1", x, raw1, out, 3, training)
            return
        b = self._buffers(x.N)
        a1, raw2 = View(b[key + ".a1"]), View(b[key + ".raw2"])
        self._cbr_fwd(blk + ".c1", blk + ".bn1", x, raw1, a1, 3, training)
        if m == 1:
            self._cbr_fwd(blk + ".c2", blk + ".bn2", a1, raw2, out, 3, training)
            return
        a2 = View(b[key + ".a2"])
        self._cbr_fwd(blk + ".c2", blk + ".bn2", a1, raw2, a2, 3, training)
        if m == 2:
            self._add(a2, x, out)
        else:
            raw3, a3 = View(b[key + ".raw3"]), View(b[key + ".a3"])
            self._cbr_fwd(blk + ".c3", blk + ".bn3", x, raw3, a3, 3, training)
            self._add(a2, a3, out)

    def _blk_bwd(self, key, blk, x, raw1, g_out, g_raw1, g_x, g_x_stats=None):
        """Backward of _blk_fwd: g_out = dL/d(block output) -> parameter gradients of the block and g_x = dL/dx
        (overwritten). Returns True when g_x_stats received the channel sums of the COMPLETE g_x."""
        m = self.mode
        if m == 0:
            self._cbr_bwd(blk + ".c1", blk + ".bn1", x, raw1, g_out, g_raw1, 3, g_x=g_x, g_x_stats=g_x_stats)
            return True
        b = self._buffers(x.N)
        a1, raw2 = View(b[key + ".a1"]), View(b[key + ".raw2"])
        g_a1, g_raw2 = View(b[f"g_{key}.a1"]), View(b[f"g_{key}.raw2"])
        # the Add of modes 2 / 3 passes g_out unchanged to both summands
        self._cbr_bwd(blk + ".c2", blk + ".bn2", a1, raw2, g_out, g_raw2, 3, g_x=g_a1)
        if m == 1:
            self._cbr_bwd(blk + ".c1", blk + ".bn1", x, raw1, g_a1, g_raw1, 3, g_x=g_x, g_x_stats=g_x_stats)
            return True
        self._cbr_bwd(blk + ".c1", blk + ".bn1", x, raw1, g_a1, g_raw1, 3, g_x=g_x)
        if m == 2:
            self._add(g_x, g_out, g_x)                      # identity shortcut
        else:
            raw3, g_raw3 = View(b[key + ".raw3"]), View(b[f"g_{key}.raw3"])
            self._cbr_bwd(blk + ".c3", blk + ".bn3", x, raw3, g_out, g_raw3, 3, g_x=g_x, accumulate=1)
        return False

    # ------------------------------------------------------------------ forward
    def stage(self, spec_in, emb, spec_out=None):
        """Copies one batch into the static input buffers (outside any captured graph)."""
        B = spec_in.shape[0]
        b = self._buffers(B)
        b["x_in"].copy_(spec_in.reshape(b["x_in"].shape), non_blocking=True)
        b["emb"].copy_(emb.reshape(B, self.T), non_blocking=True)
        if spec_out is not None:
            b["y_true"].copy_(spec_out.reshape(b["y_true"].shape), non_blocking=True)
        self._last_B = B
        return b

    def forward(self, spec_in, emb, training=False, dropout_mask=None, dropout=True):
        """model([spec_in, emb], training) -> fp32 (B, H, W, 2) in (0, 1) (a static buffer).

        dropout_mask: optional (B, dim) fp32 tensor of {0, 1/(1-rate)} to inject (parity tests);
        otherwise a fresh counter-based mask is drawn when training and dropout is True.
        """
        b = self.stage(spec_in, emb)
        B = spec_in.shape[0]
        self.sync_operands()
        if not training and self.eval_cuda_graph:
            # inference (rir_generation.py:160-170 runs batches of 4): ~60 small launches, host-bound when issued one
            # by one, so the eval forward of each batch size is captured once (after a warm-up call) and replayed.
            # Weights, BN moving statistics and I/O all live in fixed buffers, so the graph stays valid across updates.
            st = self._eval_graphs.get(B)
            if st is None:
                self._forward_body(B, False)
                self._eval_graphs[B] = "warm"
            elif st == "warm":
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize()
                with torch.cuda.graph(g):
                    self._forward_body(B, False)
                self._eval_graphs[B] = g
                g.replay()
            else:
                st.replay()
            self._last_B = B
            return b["out"]
        if training and dropout_mask is not None:
            b["mask"].copy_(dropout_mask)
        return self._forward_body(B, training, dropout, injected_mask=dropout_mask is not None)

    def _forward_body(self, B, training, dropout=True, injected_mask=False):
        b = self._buffers(B)
        kd, ku, kf, A = self.A["down_k"], self.A["up_k"], self.A["fuse_k"], self.A
        if training:
            self.stats_arena.zero_()
        elif self.fold_bn_eval:
            self._fold_refresh()
        # ---- vector block (u_net.py:253-263): independent of the encoder, so it runs on the side stream
        main = torch.cuda.current_stream()
        fork = self.overlap_wgrad
        if fork:
            self.side.wait_stream(main)
        with torch.cuda.stream(self.side if fork else main):
            L.call("embedding_fwd", b["emb"].data_ptr(), self.param["vec.emb"].data_ptr(), b["embflat"].data_ptr(),
                   B, self.T, A["emb_dim"], A["emb_vocab"])
            mask = None
            if training:
                if injected_mask:
                    mask = b["mask"]
                elif dropout:
                    L.call("dropout_mask", b["mask"].data_ptr(), b["mask"].numel(), A["dropout"],
                           self.dropout_seed, self.drop_ctr_dev.data_ptr())
                    L.call("step_increment", self.drop_ctr_dev.data_ptr())
                    mask = b["mask"]
            self._fwd_mask = mask
            L.call("dense_fwd", b["embflat"].data_ptr(), self.dense_w16.data_ptr(), self.dense_w16_t.data_ptr(),
                   self.param["vec.dense.b"].data_ptr(), L.ptr(mask), b["v16"].data_ptr(), B, self.T * A["emb_dim"],
                   self.dense_n)
        # ---- encoder (encoding_block, u_net.py:265-289)
        x = View(b["x_in"])
        for i in range(1, 6):
            n = self.F0 * 2 ** (i - 1)
            t, r = View(b[f"t{i}"]), View(b[f"r{i}"])
            e = View(b[f"cat{i}"], 0, n) if i < 5 else View(b["z"])
            self._conv_fprop(f"enc{i}.down", x, t, kd, 1 if i == 1 else 2)
            self._blk_fwd(f"enc{i}", f"enc{i}.blk", t, r, e, training)
            x = e
        # ---- Add (u_net.py:229): z = e5 + proj(v16)
        if fork:
            main.wait_stream(self.side)
        z = View(b["z"])
        if A["proj"]:
            self._conv_fprop("vec.proj", View(b["v16"]), z, 1, 1, accumulate=1)
        else:
            self._add(z, View(b["v16"]), z)
        # ---- decoder (decoding_block, u_net.py:291-321)
        x = z
        for j in (2, 3, 4, 5):
            i = 6 - j
            n = self.F0 * 2 ** (i - 1)
            cat = View(b[f"cat{i}"])
            up = View(b[f"cat{i}"], n, n)
            self._conv_dgrad(f"dec{j}.up", x, up, ku, 2, bias=True)     # Conv2DTranspose forward
            self._cbr_fwd(f"dec{j}.fuse", f"dec{j}.fuse_bn", cat, View(b[f"rf{i}"]), View(b[f"f{i}"]), kf, training)
            self._blk_fwd(f"dec{j}", f"dec{j}.blk", View(b[f"f{i}"]), View(b[f"rb{i}"]), View(b[f"d{i}"]), training)
            x = View(b[f"d{i}"])
        # ---- head: Conv2D(2, 6x6, same) + sigmoid (u_net.py:247-249) / Conv2D(2, 1x1, linear) (diff_u_net.py:257)
        self._conv_fprop("head", x, View(b["out"]), A["head_k"], 1, act=L.ACT_SIGMOID if self.head_sigmoid else L.ACT_NONE)
        self._last_B = B
        return b["out"]

    # ------------------------------------------------------------------ backward
    def backward(self, g_head):
        """g_head: fp32 (B,H,W,2) gradient w.r.t. the head's PRE-sigmoid output. Fills self.grad."""
        B = self._last_B
        b = self._buffers(B)
        if g_head.data_ptr() != b["g_out"].data_ptr():
            b["g_out"].copy_(g_head)
        self._backward_body(B)

    def _bstat(self, key, c):
        """Persistent slot of the bias-statistics arena ([sum | sumsq] of 2c floats) for `key`."""
        slots = self.__dict__.setdefault("_bstat_slots", {})
        if key not in slots:
            o = sum(n for _, n in slots.values())
            assert o + 2 * c <= self.bstat_arena.numel()
            slots[key] = (o, 2 * c)
        o, n = slots[key]
        return self.bstat_arena[o:o + n]

    def _backward_body(self, B, segment=None, dense_dw=True, join=True):
        """segment: None = everything; 0 = head + decoder, 1 = bottleneck / vector block, 2 = encoder
        (the order gradients become final, used for bucketed all-reduce overlap). dense_dw=False leaves the Dense
        kernel / bias gradients to dense_grad_from_gathered (data-parallel: gather operands, not gradients).
        join=False: do not make the main stream wait for the side-stream gradient kernels at the end of segments 0 / 1
        (the caller orders its consumer -- the bucket's all-reduce -- after BOTH streams itself, so the next segment's
        dgrad chain keeps overlapping the weight / Dense gradients exactly as in the unsegmented step)."""
        b = self._buffers(B)
        kd, ku, kf, A = self.A["down_k"], self.A["up_k"], self.A["fuse_k"], self.A
        hk = A["head_k"]
        if segment in (None, 0):
            self.bstat_arena.zero_()
            self.sums_arena.zero_()
            # one zero-fill of the flat gradient buffer replaces a memset per wgrad / bias-gradient launch; the
            # Dense kernel's gradient (plain stores, 47 MB) is skipped
            lo, n = self.offsets["vec.dense.w"]
            self.G[:lo].zero_()
            self.G[lo + n:].zero_()
            g_out = View(b["g_out"])
            d1 = View(b["d1"])
            # head
            self._conv_wgrad("head", d1, g_out, hk, 1)
            with self._side_branch():
                L.call("channel_sum", g_out.ptr(), L.F32, g_out.npix, 2, 2, 0, self.grad["head.b"].data_ptr())
            self._conv_dgrad("head", g_out, View(b["g_d1"]), hk, 1)
            # decoder, top (level 1) down to level 4
            for j in (5, 4, 3, 2):
                i = 6 - j
                n = self.F0 * 2 ** (i - 1)
                cat, g_cat = View(b[f"cat{i}"]), View(b[f"g_cat{i}"])
                self._blk_bwd(f"dec{j}", f"dec{j}.blk", View(b[f"f{i}"]), View(b[f"rb{i}"]),
                              View(b[f"g_d{i}"]), View(b[f"g_rb{i}"]), View(b[f"g_f{i}"]))
                st = self._bstat(f"dec{j}.cat", 2 * n)
                self._cbr_bwd(f"dec{j}.fuse", f"dec{j}.fuse_bn", cat, View(b[f"rf{i}"]), View(b[f"g_f{i}"]),
                              View(b[f"g_rf{i}"]), kf, g_x=g_cat, g_x_stats=st)
                # Conv2DTranspose: bias grad = channel sums of its output gradient (right half of g_cat)
                self.grad[f"dec{j}.up.b"].copy_(st[n:2 * n])
                g_up = View(b[f"g_cat{i}"], n, n)
                x_in = View(b[f"d{i + 1}"]) if j > 2 else View(b["z"])
                g_x_in = View(b[f"g_d{i + 1}"]) if j > 2 else View(b["g_z"])
                self._conv_wgrad(f"dec{j}.up", g_up, x_in, ku, 2)
                st2 = self._bstat("z", x_in.C) if (j == 2 and A["proj"]) else None
                self._conv_fprop(f"dec{j}.up", g_up, g_x_in, ku, 2, stats=st2, bias=False)   # ConvT dgrad
                if st2 is not None:
                    self.grad["vec.proj.b"].copy_(st2[:x_in.C])
            if segment == 0 and join:
                self._join_side()
        if segment in (None, 1):
            # bottleneck: z = e5 + proj(v16)
            g_z = View(b["g_z"])
            if A["proj"]:
                self._conv_wgrad("vec.proj", View(b["v16"]), g_z, 1, 1)
                self._conv_dgrad("vec.proj", g_z, View(b["g_v16"]), 1, 1)
            # the Dense / Embedding gradients only feed the optimiser: side stream, next to the encoder's dgrad chain
            with self._side_branch():
                L.call("dense_bwd", b["embflat"].data_ptr(), self.dense_w16.data_ptr(), self.dense_w16_t.data_ptr(),
                       b["g_v16"].data_ptr(), L.ptr(self._fwd_mask), b["g_v16_eff"].data_ptr(),
                       self.grad["vec.dense.w"].data_ptr() if dense_dw else None,
                       self.grad["vec.dense.b"].data_ptr() if dense_dw else None,
                       b["g_embflat"].data_ptr(), B, self.T * A["emb_dim"], self.dense_n)
                L.call("embedding_bwd", b["emb"].data_ptr(), b["g_embflat"].data_ptr(), L.BF16,
                       self.grad["vec.emb"].data_ptr(), B, self.T, A["emb_dim"], A["emb_vocab"])
            if segment == 1 and join:
                self._join_side()
        if segment in (None, 2):
            # encoder, level 5 up to level 1
            g_e = View(b["g_z"])
            for i in (5, 4, 3, 2, 1):
                n = self.F0 * 2 ** (i - 1)
                t, r = View(b[f"t{i}"]), View(b[f"r{i}"])
                st = self._bstat(f"enc{i}.t", n)
                g_t = View(b[f"g_t{i}"])
                if self._blk_bwd(f"enc{i}", f"enc{i}.blk", t, r, g_e, View(b[f"g_r{i}"]), g_t, g_x_stats=st):
                    self.grad[f"enc{i}.down.b"].copy_(st[:n])
                else:       # g_t was completed after the dgrad epilogue that produced the statistics: sum it directly
                    L.call("channel_sum", g_t.ptr(), L.BF16, g_t.npix, n, g_t.ld, g_t.coff,
                           self.grad[f"enc{i}.down.b"].data_ptr())
                if i > 1:
                    e_prev = View(b[f"cat{i - 1}"], 0, n // 2)
                    g_e_prev = View(b[f"g_cat{i - 1}"], 0, n // 2)
                    self._conv_wgrad(f"enc{i}.down", e_prev, g_t, kd, 2)
                    self._conv_dgrad(f"enc{i}.down", g_t, g_e_prev, kd, 2, accumulate=1)
                    g_e = g_e_prev
                else:
                    self._conv_wgrad("enc1.down", View(b["x_in"]), g_t, kd, 1)
            self._join_side()

    def dense_operands(self, B):
        """(x, dy) of the Dense layer's kernel gradient dw = x^T dy after backward segment 1: the flattened
        embeddings [B, T*EMB_DIM] and the (dropout-masked) output gradient [B, dense_n], both bf16."""
        b = self._buffers(B)
        return b["embflat"], (b["g_v16_eff"] if self._fwd_mask is not None else b["g_v16"])

    def dense_grad_from_gathered(self, x_all, dy_all):
        """Dense kernel and bias gradients of the GLOBAL batch from the all-gathered operands
        (rows = world * B): sum_r x_r^T dy_r == [x_1; ..; x_R]^T [dy_1; ..; dy_R]. 1.2 MB of operands per replica
        cross the NVLinks instead of a 47 MB fp32 gradient."""
        rows = x_all.shape[0]
        L.call("dense_bwd", x_all.data_ptr(), self.dense_w16.data_ptr(), self.dense_w16_t.data_ptr(),
               dy_all.data_ptr(), None, None, self.grad["vec.dense.w"].data_ptr(),
               self.grad["vec.dense.b"].data_ptr(), None, rows, self.T * self.A["emb_dim"], self.dense_n)

    # ------------------------------------------------------------------ loss + optimiser
    def loss_and_grad(self, y_true, w_amp, w_ph, need_grad=True):
        """Fused amp/phase loss on the last forward output; returns the device tensor
        [loss, mean(1-cos), mean sq err, 0] and leaves dL/d(pre-sigmoid) in the g_out buffer."""
        b = self._buffers(self._last_B)
        b["y_true"].copy_(y_true.reshape(b["y_true"].shape))
        npix = b["out"].numel() // 2
        L.call("ampphase_loss", b["y_true"].data_ptr(), b["out"].data_ptr(), npix, float(w_amp), float(w_ph),
               int(self.head_sigmoid),
               self.losses_dev.data_ptr(), b["g_out"].data_ptr() if need_grad else None, None, 0)
        return self.losses_dev

    def l2_loss_and_grad(self, scale, part=None):
        """DP loss regulariser (main_training.py:232-233): reg = scale * 0.001 * sum ||W||^2 over the
        strided convs and ConvTs; its gradient 2*scale*0.001*W is added to the flat gradient. One launch.
        part = "dec" / "enc": only the Conv2DTranspose kernels / only the encoder's strided kernels (the bucket-wise
        optimiser of the data-parallel step); their sums land in reg_dev[0] / reg_dev[1], the whole sum in reg_dev[0]."""
        if getattr(self, "_l2_tables", None) is None:
            names = [n for n in self.offsets if PL.l2_regularised(n)]
            def table(sel):
                rows = [[self.param[n].data_ptr(), self.grad[n].data_ptr(), self.param[n].numel()] for n in sel]
                return torch.tensor(rows, dtype=torch.int64, device=self.device)
            self._l2_tables = {None: table(names), "dec": table([n for n in names if n.startswith("dec")]),
                               "enc": table([n for n in names if not n.startswith("dec")])}
        t = self._l2_tables[part]
        slot = 1 if part == "enc" else 0
        if part is None:
            self.reg_dev[1:].zero_()
        L.call("l2_reg_batched", t.data_ptr(), t.shape[0], float(scale * PL.L2_COEF), self.reg_dev[slot:].data_ptr())
        return self.reg_dev

    def adam_step(self, beta1=0.9, beta2=0.999, eps=1e-7):
        L.call("adam", self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self.n_flat,
               self.lr_dev.data_ptr(), self.step_dev.data_ptr(), beta1, beta2, eps)
        L.call("step_increment", self.step_dev.data_ptr())
        self.refresh_operands()

    def adam_range(self, lo, hi, beta1=0.9, beta2=0.999, eps=1e-7):
        """Adam on flat elements [lo, hi) only (lo a multiple of 4: 16-byte accesses); the step counter is NOT advanced --
        adam_finish() does that once every range of the step has been updated."""
        if hi > lo:
            L.call("adam", self.P.data_ptr() + 4 * lo, self.G.data_ptr() + 4 * lo, self.M.data_ptr() + 4 * lo,
                   self.V.data_ptr() + 4 * lo, hi - lo, self.lr_dev.data_ptr(), self.step_dev.data_ptr(), beta1, beta2, eps)

    def adam_finish(self):
        L.call("step_increment", self.step_dev.data_ptr())
        self.refresh_operands()

    def nadam_step(self, beta1=0.9, beta2=0.999, eps=1e-7):
        """tf.keras.optimizers.Nadam on the flat buffers (amp_phase_trainer.py:30-31); the momentum-schedule product
        lives in device memory, so the step is graph capturable like adam_step."""
        if getattr(self, "nadam_coef", None) is None:
            self.nadam_coef = torch.ones(4, dtype=torch.float32, device=self.device)
        L.call("nadam", self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self.n_flat,
               self.lr_dev.data_ptr(), self.step_dev.data_ptr(), self.nadam_coef.data_ptr(), beta1, beta2, eps)
        L.call("step_increment", self.step_dev.data_ptr())
        self.refresh_operands()

    def lamb_step(self, beta1=0.9, beta2=0.999, eps=1e-6, weight_decay=0.0):
        """tensorflow_addons LAMB (trainer.py:37-38): per-variable trust ratio over the 77 trainable tensors."""
        if getattr(self, "_lamb_table", None) is None:
            rows = [[o, n] for o, n in self.offsets.values()]
            self._lamb_table = torch.tensor(rows, dtype=torch.int64, device=self.device)
            self._lamb_upd = torch.empty_like(self.P)
            self._lamb_norms = torch.zeros(2 * len(rows), dtype=torch.float32, device=self.device)
        L.call("lamb", self.P.data_ptr(), self.G.data_ptr(), self.M.data_ptr(), self.V.data_ptr(), self._lamb_upd.data_ptr(),
               self._lamb_table.data_ptr(), self._lamb_table.shape[0], self._lamb_norms.data_ptr(), self.lr_dev.data_ptr(),
               self.step_dev.data_ptr(), beta1, beta2, eps, weight_decay)
        L.call("step_increment", self.step_dev.data_ptr())
        self.refresh_operands()

    def sgd_step(self):
        L.call("sgd", self.P.data_ptr(), self.G.data_ptr(), self.n_flat, self.lr_dev.data_ptr())
        L.call("step_increment", self.step_dev.data_ptr())
        self.refresh_operands()

    def set_lr(self, lr):
        # the captured step reads the rate from device memory; rewritten only when it changes (Trainer.step calls this
        # every step: one tiny launch less on the host path between two graph replays)
        lr = float(lr)
        if lr != getattr(self, "_lr_host", None):
            self.lr_dev.fill_(lr)
            self._lr_host = lr

    # ------------------------------------------------------------------ debugging / parity
    def forward_state(self):
        """Every tensor the last TRAINING forward stored, keyed for UNetOracle.override: raw conv outputs,
        BN+ReLU outputs, BatchNorm batch statistics (fp32 CPU; activations NCHW)."""
        b = self._buffers(self._last_B)
        st = {}

        def nchw(t):
            return t.float().permute(0, 3, 1, 2).contiguous().cpu()

        def bn(bname):
            if not self.BatchNorm:
                return
            mr = self._slot(self.mr_arena, bname).float().cpu()
            c = mr.numel() // 2
            st[bname + ".mean"] = mr[:c].clone()
            st[bname + ".var"] = (1.0 / (mr[c:] ** 2) - PL.BN_EPS).clone()

        for i in range(1, 6):
            n = self.F0 * 2 ** (i - 1)
            st[f"enc{i}.down"] = nchw(b[f"t{i}"])
            st[f"enc{i}.blk.c1"] = nchw(b[f"r{i}"])
            bn(f"enc{i}.blk.bn1")
            # the last BN+ReLU output of the block IS the block output for modes 0 / 1 (e5 is skipped: the vector
            # projection is accumulated into the same buffer afterwards)
            if i < 5 and self.mode <= 1:
                st[f"enc{i}.blk.bn{self.mode + 1}.out"] = nchw(b[f"cat{i}"][..., :n])

        def extra(key):
            blk = key + ".blk"
            st[blk + ".bn1.out"] = nchw(b[key + ".a1"])
            st[blk + ".c2"] = nchw(b[key + ".raw2"]); bn(blk + ".bn2")
            if self.mode >= 2:
                st[blk + ".bn2.out"] = nchw(b[key + ".a2"])
            if self.mode == 3:
                st[blk + ".c3"] = nchw(b[key + ".raw3"]); bn(blk + ".bn3")
                st[blk + ".bn3.out"] = nchw(b[key + ".a3"])

        if self.mode != 0:
            for i in range(1, 6):
                extra(f"enc{i}")
        st["bottleneck"] = nchw(b["z"])
        st["vec.dense.out"] = b["v16"].float().reshape(self._last_B, -1).cpu()
        for j in (2, 3, 4, 5):
            i = 6 - j
            n = self.F0 * 2 ** (i - 1)
            st[f"dec{j}.up"] = nchw(b[f"cat{i}"][..., n:])
            st[f"dec{j}.fuse"] = nchw(b[f"rf{i}"]); bn(f"dec{j}.fuse_bn"); st[f"dec{j}.fuse_bn.out"] = nchw(b[f"f{i}"])
            st[f"dec{j}.blk.c1"] = nchw(b[f"rb{i}"]); bn(f"dec{j}.blk.bn1")
            if self.mode != 0:
                extra(f"dec{j}")
            if self.mode <= 1:
                st[f"dec{j}.blk.bn{self.mode + 1}.out"] = nchw(b[f"d{i}"])
        return st

    def debug_tensors(self):
        """Intermediate tensors of the last forward, keyed like the oracle's taps (fp32, NCHW)."""
        b = self._buffers(self._last_B)
        out = {}
        for i in range(1, 6):
            out[f"enc{i}.down"] = b[f"t{i}"]
            out[f"enc{i}.blk.c1"] = b[f"r{i}"]
        for j in (2, 3, 4, 5):
            i = 6 - j
            n = self.F0 * 2 ** (i - 1)
            out[f"dec{j}.up"] = b[f"cat{i}"][..., n:]
            out[f"dec{j}.fuse"] = b[f"rf{i}"]
            out[f"dec{j}.blk.c1"] = b[f"rb{i}"]
        out["bottleneck"] = b["z"]
        return {k: v.float().permute(0, 3, 1, 2).contiguous() for k, v in out.items()}
