"""Device engine: stacked BiLSTM -> Dense -> CTC on libasr_b200 kernels.

Host code is PyTorch only for device memory, streams and torch.distributed; all
arithmetic is the hand-written CUDA behind include/asr_b200.h (no torch ops on
the compute path, no autograd, no CPU fallback).

Mirrors the graph the reference builds in core/models.py:217-281 (brsmv1) /
:55-73 (graves2006) + ctc_model (:31-52) and the optimiser set-up of
train.py:133-143.  Internal layout is time-major [T, N, *].
"""
from __future__ import annotations

import ctypes as C
import math
import dataclasses
from dataclasses import dataclass

import numpy as np
import torch

from ._lib import (GEMM_BACKGROUND, GEMM_TILE128, LSTM_SHARED_SM, CastJob, ConvGeom, ConvPlan, LstmBwdArgs, LstmFwdArgs, LstmVariant,
                   LstmVariantGrads, cur_stream, lib, ptr)

F16, BF16 = 0, 1
DP_SLICE_BYTES = 128 << 20     # gradient buckets above this are all-reduced in overlapped slices (engine.backward)
F16_LO = 16          # fp16(v - fp16(v)): the low half of a split-precision operand (csrc/utils.cu)
OUT_F32, OUT_F16, OUT_BF16 = 0, 1, 2


def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


TC_WIDTHS = (128, 256, 384, 512, 640, 768, 832, 896)     # widths the tensor-core recurrences are instantiated for


def tc_width(H: int) -> int:
    """The width the tensor-core recurrences run a layer of H units at: H itself when instantiated, else the next
    instantiated width (the extra units are zero-padded: see ParamBucket).  Narrow layers (H <= 128, e.g. the
    BiLSTM-100 of graves2006) and layers wider than 896 are left alone (fp32 persistent / general-cell engines)."""
    if H <= 128 or H > TC_WIDTHS[-1] or H in TC_WIDTHS:
        return H
    return next(w for w in TC_WIDTHS if w >= H)


def _pad_blocks(a: np.ndarray, axis: int, blocks: int, H: int, Hp: int) -> np.ndarray:
    """axis = `blocks` consecutive blocks of H -> blocks of Hp, zero-filled behind each block."""
    shp = list(a.shape)
    assert shp[axis] == blocks * H
    out = np.zeros(shp[:axis] + [blocks, Hp] + shp[axis + 1:], a.dtype)
    idx = [slice(None)] * out.ndim
    idx[axis + 1] = slice(0, H)
    out[tuple(idx)] = a.reshape(shp[:axis] + [blocks, H] + shp[axis + 1:])
    return out.reshape(shp[:axis] + [blocks * Hp] + shp[axis + 1:])


def _unpad_blocks(a: np.ndarray, axis: int, blocks: int, H: int, Hp: int) -> np.ndarray:
    shp = list(a.shape)
    assert shp[axis] == blocks * Hp
    idx = [slice(None)] * (a.ndim + 1)
    idx[axis + 1] = slice(0, H)
    v = a.reshape(shp[:axis] + [blocks, Hp] + shp[axis + 1:])[tuple(idx)]
    return np.ascontiguousarray(v).reshape(shp[:axis] + [blocks * H] + shp[axis + 1:])


@dataclass
class ModelSpec:
    num_features: int = 26
    num_hiddens: int = 512
    num_layers: int = 3
    num_classes: int = 28
    weight_decay: float = 0.0        # l2 on W, U of every LSTM and the Dense kernel (models.py:263-264,279)
    name: str = "brsmv1"
    dropout: float = 0.0             # variational dropout_W = dropout_U (core/models.py:265-266), train phase only
    # brsmv1 switches (core/models.py:217-281); any of them routes the recurrence to the general-cell engine
    zoneout: float = 0.0             # zoneout_c = zoneout_h (core/models.py:267-268)
    layer_norm: tuple | None = None  # (gain_init, bias_init)            (core/layers.py:407-422)
    mi: tuple | None = None          # (alpha_init, beta1_init, beta2_init) (core/layers.py:391-405)
    residual: str | None = None      # merge mode; 'sum' is built (core/models.py:253-255, 273-274)
    input_dropout: bool = False      # element-wise Dropout(dropout) on the (projected) input (core/models.py:257-258)

    # heterogeneous stacks (eyben, core/models.py:76-103): per-layer widths and a linear input projection
    layer_hiddens: tuple | None = None   # overrides num_hiddens / num_layers when set
    input_dense: int | None = None       # TimeDistributed(Dense(n)) in front of the first BiLSTM (no residual merges)

    # convolutional front end of BASELINE configs[3] (DeepSpeech2-style; NOT in the reference): layers of
    # (C_out, kt, kf, stride_t, stride_f), "same"-style zero padding (k - 1) // 2, bias, clipped ReLU at conv_clip
    conv_front: tuple | None = None
    conv_clip: float = 20.0

    def conv_shapes(self, T=None):
        """per conv layer: (C_in, F_in, C_out, F_out, K) and, when T is given, the output frame counts."""
        out, C, F = [], 1, self.num_features
        for (co, kt, kf, st, sf) in self.conv_front or ():
            Fo = (F + 2 * ((kf - 1) // 2) - kf) // sf + 1
            To = None if T is None else (T + 2 * ((kt - 1) // 2) - kt) // st + 1
            out.append(dict(C_in=C, F_in=F, C_out=co, F_out=Fo, K=kt * kf * C, T_in=T, T_out=To))
            C, F, T = co, Fo, To
        return out

    @property
    def lstm_in(self) -> int:
        """feature width the first BiLSTM (or its input projection) sees"""
        if self.conv_front:
            last = self.conv_shapes()[-1]
            return last["F_out"] * last["C_out"]
        return self.num_features

    @property
    def hs(self):
        return tuple(self.layer_hiddens) if self.layer_hiddens else (self.num_hiddens,) * self.num_layers

    @property
    def proj_width(self):
        """width of the input projection, or None: 2H for the residual stack (core/models.py:253-255), input_dense else."""
        if self.residual is not None:
            return 2 * self.hs[0]
        return self.input_dense

    @property
    def general(self) -> bool:
        """switches only the general-cell path implements (layer norm needs whole-row statistics every step; the
        residual / projection / heterogeneous stacks use its per-layer operand plumbing)."""
        if self.conv_front and (self.layer_norm is not None or self.residual is not None or self.input_dropout or
                                self.input_dense or len(set(self.hs)) > 1 or self.zoneout or self.mi is not None):
            raise NotImplementedError("the convolutional front end is built in front of the default BiLSTM stack only")
        return bool(self.layer_norm is not None or self.residual is not None or self.input_dropout or self.input_dense
                    or len(set(self.hs)) > 1)

    @property
    def elementwise(self) -> bool:
        """multiplicative integration / zoneout: element-wise in the step, a template switch of the tensor-core kernels."""
        return bool(self.zoneout or self.mi is not None)


class ParamBucket:
    """One flat fp32 parameter vector (+ grad, Adam m/v, l2 mask) with named views.

    Per layer the order is Wf, Wb, Uf, Ub, bf, bb so that [Uf|Ub] is a contiguous
    [2,H,4H] block and [bf|bb] a contiguous [2,4H] block (what the kernels take).

    `logical_h` (optional): the model's own width when `spec` is the zero-padded device spec (tc_width): load() pads
    every tensor per gate / per direction block, export() cuts the padding off again.  A padded unit has W = U = b = 0,
    so z = 0, i = f = o = 0.5, g = 0, c = h = 0 at every step, its dz is 0 in BPTT and every gradient, Adam moment and
    l2 term that touches it stays exactly 0: the padded model IS the logical model.
    """

    def __init__(self, spec: ModelSpec, device, logical_h: int | None = None):
        self.spec = spec
        self.logical_h = logical_h if (logical_h and logical_h != spec.num_hiddens) else None
        C = spec.num_classes
        shapes = []
        for i, cs in enumerate(spec.conv_shapes()):          # conv{i}.W [C_out, (kt, kf, c_in)], conv{i}.b [C_out]
            shapes += [(f"conv{i}.W", (cs["C_out"], cs["K"])), (f"conv{i}.b", (cs["C_out"],))]
        D = spec.lstm_in
        if spec.proj_width:
            shapes += [("proj.W", (D, spec.proj_width)), ("proj.b", (spec.proj_width,))]
            D = spec.proj_width
        for l, H in enumerate(spec.hs):
            shapes += [(f"l{l}.Wf", (D, 4 * H)), (f"l{l}.Wb", (D, 4 * H)),
                       (f"l{l}.Uf", (H, 4 * H)), (f"l{l}.Ub", (H, 4 * H)),
                       (f"l{l}.bf", (4 * H,)), (f"l{l}.bb", (4 * H,))]
            if spec.mi is not None:          # [2, 4H]: forward | backward direction
                shapes += [(f"l{l}.{n}", (2, 4 * H)) for n in ("mi_alpha", "mi_beta1", "mi_beta2")]
            if spec.layer_norm is not None:
                for n, wdt in (("uh", 4 * H), ("wx", 4 * H), ("c", H)):
                    shapes += [(f"l{l}.ln_gain_{n}", (2, wdt)), (f"l{l}.ln_bias_{n}", (2, wdt))]
            D = 2 * H
        shapes += [("dense.W", (D, C)), ("dense.b", (C,))]
        self.shapes = dict(shapes)
        self.offsets, off = {}, 0
        for k, s in shapes:
            self.offsets[k] = off
            off += int(np.prod(s))
            off = (off + 3) // 4 * 4                       # keep every tensor 16-byte aligned
        self.numel = off
        self.loads = 0
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.flat)
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.decay = torch.zeros(off, dtype=torch.uint8, device=device)
        for k in self.shapes:
            if k.endswith((".Wf", ".Wb", ".Uf", ".Ub")) or k in ("dense.W", "proj.W") or (k.startswith("conv") and k.endswith(".W")):
                self._view(self.decay, k).fill_(1)

    def _view(self, flat, k):
        o, s = self.offsets[k], self.shapes[k]
        return flat[o:o + int(np.prod(s))].view(*s)

    def p(self, k):
        return self._view(self.flat, k)

    def g(self, k):
        return self._view(self.grad, k)

    def _blocks(self, k):
        """(axis, number of H-wide blocks) pairs of tensor k that carry the hidden width."""
        if k == "dense.W":
            return [(0, 2)]
        if not k.startswith("l") or "." not in k:
            return []
        l, n = k.split(".", 1)
        if n in ("Wf", "Wb"):
            return [(1, 4)] + ([(0, 2)] if int(l[1:]) > 0 else [])
        if n in ("Uf", "Ub"):
            return [(0, 1), (1, 4)]
        if n in ("bf", "bb"):
            return [(0, 4)]
        if n.startswith("mi_"):
            return [(1, 4)]
        raise KeyError(k)

    def load(self, params: dict, which="flat"):
        self.loads += 1
        for k, v in params.items():
            v = np.asarray(v, dtype=np.float32)
            if self.logical_h:
                for axis, blocks in self._blocks(k):
                    v = _pad_blocks(v, axis, blocks, self.logical_h, self.spec.num_hiddens)
            self._view(getattr(self, which), k).copy_(torch.as_tensor(v))

    def export(self, which="flat") -> dict:
        src = getattr(self, which)
        out = {k: self._view(src, k).detach().cpu().numpy().copy() for k in self.shapes}
        if self.logical_h:
            for k in out:
                for axis, blocks in self._blocks(k):
                    out[k] = _unpad_blocks(out[k], axis, blocks, self.logical_h, self.spec.num_hiddens)
        return out


class AcousticEngine:
    """Forward / backward / optimiser step for one rank."""

    def __init__(self, spec: ModelSpec, device="cuda:0", seed=4321, init_params: dict | None = None, pad_width=True,
                 overlap=True, fp16_storage=True):
        """pad_width: widths without a tensor-core instantiation run zero-padded at the next one (else on the general
        cell); overlap: gradient GEMMs / weight preparation on a side stream beside the recurrences; fp16_storage: zx and
        the saved gates / cell state are fp16 in HBM and move through TMA (csrc/lstm_tc4.cu) where that engine takes the
        shape."""
        # widths without a tensor-core instantiation (e.g. the BiLSTM-800 of BASELINE config 4) run zero-padded at the
        # next instantiated width; self.spec is the device spec, self.user_spec the model's own
        self.user_spec = spec
        Hl = spec.num_hiddens
        if not spec.general and not spec.layer_hiddens and tc_width(Hl) != Hl and pad_width:
            spec = dataclasses.replace(spec, num_hiddens=tc_width(Hl))
        self.spec = spec
        self.logical_h = Hl if spec.num_hiddens != Hl else None
        self.device = torch.device(device)
        torch.cuda.set_device(self.device)
        lib.load()
        self.params = ParamBucket(spec, self.device, logical_h=self.logical_h)
        if init_params is None:
            init_params = self.keras_init(self.user_spec, seed)
        self.params.load(init_params)
        self.step_count = 0
        self._ws = {}
        self._views = {}
        self._shape = None
        self._sqnorm = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device=self.device)
        self._weights_version = -1
        # dW/dU GEMMs of layer l run on a low-priority side stream while the BPTT recurrence of layer l-1 (128 of the
        # 148 SMs, which it owns exclusively — see exclusive_smem() in csrc/lstm_tc2.cu) runs on the high-priority
        # main stream; the GEMM CTAs fill the 20 idle SMs and never delay the critical path.
        self._main = torch.cuda.Stream(device=self.device, priority=-1)
        self._side = torch.cuda.Stream(device=self.device, priority=0)
        self.overlap = bool(overlap)
        self.dp_slices = None         # data parallel: per-slice all-reduce hook (True) or one collective (False); None = by bucket size
        self.fp16_storage = bool(fp16_storage)
        self.shared_sm = False          # let other kernels' CTAs share SMs with the recurrences / pin the small GEMM tiling
        self._seed = int(seed)
        self._mask_seed, self._mask_offset = (seed + 17) * 0x9E3779B1 & 0xFFFFFFFFFFFFFFFF, 0
        self._l2 = torch.zeros(1, dtype=torch.float64, device=self.device)
        self._prepared_for = None

    @property
    def params_version(self):
        """changes whenever the fp32 masters may have changed (optimiser steps, loads): keys the cached 16-bit operands"""
        return (self.step_count, self.params.loads)

    def set_rank(self, rank: int):
        """data parallel: same parameters on every rank (same init seed), different dropout / zoneout / noise streams."""
        self._mask_seed = ((self._seed + 17) * 0x9E3779B1 + 0x632BE59BD9B4E019 * int(rank)) & 0xFFFFFFFFFFFFFFFF

    @property
    def lstm_opts(self) -> int:
        return LSTM_SHARED_SM if self.shared_sm else 0

    @property
    def gemm_flags(self) -> int:
        return GEMM_TILE128 if self.shared_sm else 0

    def add_gaussian_noise(self, xt, n_real, std, offset):
        """GaussianNoise(std) on the first n_real utterances of time-major features [T, Np, F] (train phase,
        core/models.py:67,251); returns the advanced generator offset."""
        T, Np, F = xt.shape
        lib.asr_add_gaussian_noise(ptr(xt), T, n_real * F, Np * F, float(std), self._mask_seed ^ 0xA5A5A5A5, int(offset), cur_stream())
        return int(offset) + T * n_real * F

    def l2_penalty(self):
        """sum of the l2(weight_decay) regularisers (core/models.py:263-264, 279) as a device scalar: the masked squared
        norm kernel with a zero gradient scale gives (2 c p)^2 summed over the decayed tensors; c = sqrt(wd) / 2."""
        wd = float(self.spec.weight_decay)
        P = self.params
        if not wd:
            return torch.zeros((), dtype=torch.float32, device=self.device)
        lib.asr_grad_sqnorm(ptr(P.flat), ptr(P.flat), ptr(P.decay), P.numel, 0.0, 0.5 * math.sqrt(wd), ptr(self._l2), cur_stream())
        return self._l2[0].float()

    # ------------------------------------------------------------ dropout masks
    def sample_masks(self, N):
        """Keras-1 LSTM.get_constants: one B_W [N, D] and one B_U [N, H] mask per direction and layer, sampled
        once per batch, constant over time, scaled by 1/(1-p) (K.dropout).  Returns {layer: {Wf,Wb,Uf,Ub}} plus the
        packed views W2 [2,N,D] / U2 [2,N,H] the kernels take.  One asr_dropout_mask launch fills all of them."""
        sp, p = self.spec, float(self.spec.dropout)
        hs = sp.hs
        widths = [((sp.proj_width or sp.lstm_in) if l == 0 else 2 * hs[l - 1]) for l in range(len(hs))]
        total = sum(2 * N * (D + H) for D, H in zip(widths, hs))
        flat = self._buf("dropout_masks", (total,), torch.float32)
        lib.asr_dropout_mask(ptr(flat), total, p, self._mask_seed, self._mask_offset, cur_stream())
        self._mask_offset += total
        out, o = {}, 0
        for l, (D, H) in enumerate(zip(widths, hs)):
            W2 = flat[o:o + 2 * N * D].view(2, N, D); o += 2 * N * D
            U2 = flat[o:o + 2 * N * H].view(2, N, H); o += 2 * N * H
            out[l] = {"Wf": W2[0], "Wb": W2[1], "Uf": U2[0], "Ub": U2[1], "W2": W2, "U2": U2}
        return out

    @staticmethod
    def _packed(mk, key):
        """[2, N, *] contiguous (fwd | bwd) view of a layer's masks: the packed view when sampled here, a stack when
        the caller supplied separate tensors."""
        if key + "2" in mk:
            return mk[key + "2"]
        return torch.stack([mk[key + "f"], mk[key + "b"]]).contiguous()

    # ------------------------------------------------------------------ init
    @staticmethod
    def keras_init(spec: ModelSpec, seed: int) -> dict:
        """Keras-1.2.2 initialisers for LSTM(consume_less='gpu') and Dense:
        glorot_uniform W, orthogonal(1.1) U, zero b with forget slice = 1."""
        rng = np.random.RandomState(seed)
        out = {}

        def glorot(shape):
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            return rng.uniform(-lim, lim, size=shape).astype(np.float32)

        def orth(shape):
            a = rng.normal(0.0, 1.0, shape)
            u, _, v = np.linalg.svd(a, full_matrices=False)
            q = u if u.shape == tuple(shape) else v
            return (1.1 * q.reshape(shape)).astype(np.float32)

        for i, cs in enumerate(spec.conv_shapes()):
            fan_in, fan_out = cs["K"], cs["K"] // cs["C_in"] * cs["C_out"]
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            out[f"conv{i}.W"] = rng.uniform(-lim, lim, size=(cs["C_out"], cs["K"])).astype(np.float32)
            out[f"conv{i}.b"] = np.zeros(cs["C_out"], np.float32)
        D = spec.lstm_in
        if spec.proj_width:
            out["proj.W"] = glorot((D, spec.proj_width))
            out["proj.b"] = np.zeros(spec.proj_width, np.float32)
            D = spec.proj_width
        for l, H in enumerate(spec.hs):
            for d in ("f", "b"):
                out[f"l{l}.W{d}"] = glorot((D, 4 * H))
                out[f"l{l}.U{d}"] = orth((H, 4 * H))
                b = np.zeros(4 * H, np.float32)
                b[H:2 * H] = 1.0
                out[f"l{l}.b{d}"] = b
            if spec.mi is not None:          # k_init(k) = k * ones (core/initializers.py:6-10)
                for n, k in zip(("mi_alpha", "mi_beta1", "mi_beta2"), spec.mi):
                    out[f"l{l}.{n}"] = np.full((2, 4 * H), float(k), np.float32)
            if spec.layer_norm is not None:
                g0, b0 = spec.layer_norm
                for n, wdt in (("uh", 4 * H), ("wx", 4 * H), ("c", H)):
                    out[f"l{l}.ln_gain_{n}"] = np.full((2, wdt), float(g0), np.float32)
                    out[f"l{l}.ln_bias_{n}"] = np.full((2, wdt), float(b0), np.float32)
            D = 2 * H
        out["dense.W"] = glorot((D, spec.num_classes))
        out["dense.b"] = np.zeros(spec.num_classes, np.float32)
        return out

    # ------------------------------------------------------------- workspace
    def _buf(self, name, shape, dtype, zero=False):
        t = self._ws.get(name)
        n = int(np.prod(shape))
        if t is None or t.numel() < n or t.dtype != dtype:
            t = (torch.zeros if zero else torch.empty)(n, dtype=dtype, device=self.device)
            self._ws[name] = t
        v = t[:n].view(*shape)
        self._views[name] = v          # last shaped view (self._ws holds the flat storage)
        return v

    def _alloc(self, T, N, training):
        sp = self.spec
        H, L, Cc = sp.hs[0], len(sp.hs), sp.num_classes
        R = T * N
        w = {}
        D0 = _pad8(sp.lstm_in)
        sdt = torch.float16 if self._fp16 else torch.float32       # zx / gates / cell storage (csrc/lstm_tc4.cu)
        w["x16"] = self._buf("x16", (R, D0), torch.float16, zero=True)
        w["zx"] = self._buf("zx", (R, 8 * H), sdt)
        for l in range(L):
            w[f"h16.{l}"] = self._buf(f"h16.{l}", (R, 2 * H), torch.float16)
        w["logits"] = self._buf("logits", (T, N, Cc), torch.float32)
        if training:
            w["xT16"] = self._buf("xT16", (sp.lstm_in, R), torch.bfloat16)
            for l in range(L):
                w[f"hT16.{l}"] = self._buf(f"hT16.{l}", (2 * H, R), torch.bfloat16)
                w[f"gates.{l}"] = self._buf(f"gates.{l}", (R, 8 * H), sdt)
                w[f"cell.{l}"] = self._buf(f"cell.{l}", (R, 2 * H), sdt)
            w["dlogits"] = self._buf("dlogits", (T, N, Cc), torch.float32)
            w["dl16"] = self._buf("dl16", (R, _pad8(Cc)), torch.bfloat16, zero=True)
            w["dlT16"] = self._buf("dlT16", (Cc, R), torch.bfloat16)
            w["dhA"] = self._buf("dhA", (R, 2 * H), torch.float32)
            w["dhB"] = self._buf("dhB", (R, 2 * H), torch.float32)
            for l in range(L):      # per layer: the dW/dU GEMMs of layer l overlap the recurrence of layer l-1
                w[f"dz16.{l}"] = self._buf(f"dz16.{l}", (R, 8 * H), torch.bfloat16)
                w[f"dzT16.{l}"] = self._buf(f"dzT16.{l}", (8 * H, R), torch.bfloat16)
            w["loss"] = self._buf("loss", (N,), torch.float32)
        return w

    # ---------------------------------------------------- weight operand prep
    def _prep_weights(self, training, first_only=False, skip_first=False):
        """16-bit tensor-core operands derived from the fp32 masters (once per step).
        first_only / skip_first split the work: what the first recurrent layer's forward pass needs (its W^T and U^T)
        is made on the main stream, everything else (the other layers, every BPTT operand, the Dense pair) on the side
        stream beside the first recurrence; each of the two is ONE asr_cast_batch launch over its list of tensors."""
        sp, P = self.spec, self.params
        Cc = sp.num_classes
        jobs = []

        def rows(src, ld_src, dst, ld_dst, n_rows, n_cols, dtype, exact=False):   # asr_cast_rows semantics; exact: no K-padding
            jobs.append(CastJob(src.data_ptr(), ld_src, dst.data_ptr(), ld_dst, n_rows, n_cols, dtype, 2 if exact else 0))

        def transpose(src, ld_src, dst, ld_dst, n_rows, n_cols, dtype):  # asr_cast_transpose semantics
            jobs.append(CastJob(src.data_ptr(), ld_src, dst.data_ptr(), ld_dst, n_rows, n_cols, dtype, 1))

        D = sp.proj_width or sp.lstm_in                             # width the first BiLSTM sees
        for l, H in enumerate(sp.hs):
            Dp = _pad8(D)
            fwd_ops = not (skip_first and l == 0)
            rest = not first_only
            if first_only and l > 0:
                break
            if fwd_ops:
                wt = self._buf(f"WcatT16.{l}", (8 * H, Dp), torch.float16, zero=True)     # [8H, D]  fwd B operand
                for i, d in enumerate("fb"):
                    transpose(P.p(f"l{l}.W{d}"), 4 * H, wt[i * 4 * H:], Dp, D, 4 * H, F16)
                if sp.layer_norm is not None:     # split-precision projection (fp16 rounding residuals), see _forward_general
                    wl = self._buf(f"WcatT16lo.{l}", (8 * H, Dp), torch.float16, zero=True)
                    for i, d in enumerate("fb"):
                        transpose(P.p(f"l{l}.W{d}"), 4 * H, wl[i * 4 * H:], Dp, D, 4 * H, F16_LO)
                ut = self._buf(f"UT16.{l}", (2, 4 * H, H), torch.float16)                 # [2, 4H, H] U^T, fwd recurrence
                for i, d in enumerate("fb"):
                    transpose(P.p(f"l{l}.U{d}"), 4 * H, ut[i], H, H, 4 * H, F16)
            if rest and training:
                ub = self._buf(f"Ub16.{l}", (2, H, 4 * H), torch.bfloat16)             # [2, H, 4H] U, BPTT recurrence
                rows(P.p(f"l{l}.Uf"), 4 * H, ub, 4 * H, 2 * H, 4 * H, BF16)
            if rest and training and (l > 0 or sp.proj_width or sp.conv_front):
                wc = self._buf(f"Wcat16.{l}", (D, 8 * H), torch.bfloat16)              # [D, 8H]  dX B operand
                for i, d in enumerate("fb"):
                    # the two directions are column blocks of one matrix and the jobs of a launch run concurrently: no fill
                    # past 4H (at 4H % 8 != 0 the forward block's padding would land on the backward block's first columns)
                    rows(P.p(f"l{l}.W{d}"), 4 * H, wc[:, i * 4 * H:], 8 * H, D, 4 * H, BF16, exact=True)
            D = 2 * H
        if not first_only:
            dp, H2 = _pad8(Cc), 2 * sp.hs[-1]
            wd = self._buf("WdT16", (Cc, _pad8(H2)), torch.float16, zero=True)             # [C, 2H (padded to 8)] logits B operand
            transpose(P.p("dense.W"), Cc, wd, _pad8(H2), H2, Cc, F16)
            if training:
                wdb = self._buf("Wd16", (H2, dp), torch.bfloat16, zero=True)               # [2H, Cpad] dTop B operand
                rows(P.p("dense.W"), Cc, wdb, dp, H2, Cc, BF16)
        if jobs:                                                    # one launch for the whole list (csrc/utils.cu)
            lib.asr_cast_batch((CastJob * len(jobs))(*jobs), len(jobs), cur_stream())

    def _gemm(self, din, dout, M, N, K, A, lda, B, ldb, Cm, ldc, bias=None, alpha=1.0, acc=0):
        lib.asr_gemm_tn_ex(din, dout, M, N, K, ptr(A), lda, ptr(B), ldb, ptr(Cm), ldc, ptr(bias), float(alpha), acc,
                           self.gemm_flags, cur_stream())

    # ------------------------------------------------- convolutional front end (BASELINE configs[3])
    def _conv_forward(self, x, training):
        """x f32 [T, N, F] -> f32 [T', N, F' * C].  Per layer ONE tcgen05 GEMM whose A operand is the overlapping-row view
        of the zero-padded batch-major fp16 activations (csrc/conv.cu: no patch matrix), B the banded frequency-Toeplitz
        image of the kernel, bias in the epilogue; then clipped ReLU, written straight into the next layer's padded
        input (or, for the last layer, as the time-major input of the first BiLSTM)."""
        sp, P, st = self.spec, self.params, cur_stream()
        T, N, F = x.shape
        Cin = 1
        self._conv = []
        layers = []
        fresh = getattr(self, "_conv_key", None) != (T, N, F)   # a new shape re-uses the flat buffers with another layout:
        self._conv_key = (T, N, F)                              # what is padding now may hold old activations
        if fresh:
            self._conv_zero_g = True
        for i, (co, kt, kf, s_t, s_f) in enumerate(sp.conv_front):
            g = ConvGeom(T=T, N=N, F=F, C=Cin, kt=kt, kf=kf, st=s_t, sf=s_f, pt=(kt - 1) // 2, pf=(kf - 1) // 2)
            pl = ConvPlan()
            lib.asr_conv_plan_for(C.byref(g), C.byref(pl))
            layers.append((g, pl, co))
            T, F, Cin = pl.t_out, pl.f_out, co
        for i, (g, pl, co) in enumerate(layers):
            W_in, W_out, M = g.F * g.C, pl.f_out * co, g.N * pl.rows
            # padded input [N, t_padded, W_in] + slack for the windows of the scratch rows; zeroed once, padding never written
            xp = self._buf(f"conv{i}.xp", (g.N * pl.t_padded * W_in + pl.k_padded + 64,), torch.float16, zero=True)
            if i == 0:
                if fresh:
                    xp.zero_()
                lib.asr_conv_pack(ptr(x), C.byref(g), ptr(xp), st)
            wt = self._buf(f"conv{i}.wt16", (W_out, pl.k_padded), torch.float16)
            w2 = self._buf(f"conv{i}.w2_16", (W_in, g.kt * W_out), torch.bfloat16) if (training and i > 0) else None
            bt = self._buf(f"conv{i}.bias_t", (W_out,), torch.float32)
            lib.asr_conv_toeplitz(ptr(P.p(f"conv{i}.W")), ptr(P.p(f"conv{i}.b")), C.byref(g), co, ptr(wt), pl.k_padded, ptr(w2),
                                  ptr(bt), st)
            z = self._buf(f"conv{i}.z", (M, W_out), torch.float32)
            self._gemm(F16, OUT_F32, M, W_out, pl.k_padded, xp, g.st * W_in, wt, pl.k_padded, z, W_out, bias=bt)
            last = i == len(layers) - 1
            if last:                                   # fp16 copy in GEMM-row space (the backward mask) + the BiLSTM's input
                y16, y_rows, y_row0 = self._buf(f"conv{i}.y16", (M, W_out), torch.float16), pl.rows, 0
                y32 = self._buf("conv.out32", (pl.t_out, g.N, W_out), torch.float32)
            else:                                      # the next layer's padded input
                gn, pn = layers[i + 1][0], layers[i + 1][1]
                y16 = self._buf(f"conv{i + 1}.xp", (g.N * pn.t_padded * W_out + pn.k_padded + 64,), torch.float16, zero=True)
                y_rows, y_row0, y32 = pn.t_padded, gn.pt, None
                if fresh:
                    y16.zero_()
            lib.asr_conv_act(ptr(z), C.byref(g), co, float(sp.conv_clip), ptr(y16), y_rows, y_row0, ptr(y32), st)
            self._conv.append(dict(geom=g, plan=pl, co=co, xp=xp, w2=w2, y16=y16, y_rows=y_rows, y_row0=y_row0))
        return y32

    def _conv_backward(self, dx):
        """dx f32 [T' * N, F' * C] (dL/d input of the first BiLSTM, time-major) -> conv{i}.W / conv{i}.b gradients."""
        sp, P, st = self.spec, self.params, cur_stream()
        gout, g_ts, g_ns, g_row0 = dx, self._conv[-1]["geom"].N, 1, 0
        for i in range(len(sp.conv_front) - 1, -1, -1):
            c = self._conv[i]
            g, pl, co = c["geom"], c["plan"], c["co"]
            W_in, W_out, M = g.F * g.C, pl.f_out * co, g.N * pl.rows
            # dL/dz in GEMM-row space behind kt - 1 zero rows (the mirrored input-gradient GEMM reads them), its transpose, f32
            gbuf = self._buf(f"conv{i}.g16", ((g.kt - 1 + M) * W_out,), torch.bfloat16, zero=True)
            if getattr(self, "_conv_zero_g", False):
                gbuf.zero_()
            g16 = gbuf[(g.kt - 1) * W_out:]
            gT16 = self._buf(f"conv{i}.gT16", (W_out, M), torch.bfloat16)
            g32 = self._buf(f"conv{i}.g32", (M, W_out), torch.float32)
            lib.asr_conv_act_backward(ptr(gout), g_ts, g_ns, g_row0, ptr(c["y16"]), c["y_rows"], c["y_row0"], C.byref(g), co,
                                      float(sp.conv_clip), ptr(g16), ptr(gT16), ptr(g32), st)
            cs = self._buf(f"conv{i}.colsum", (W_out,), torch.float32)
            lib.asr_colsum(ptr(g32), W_out, M, W_out, ptr(cs), st)
            # d(banded matrix) [W_out, K] = g^T [W_out, M] . view(xp)^T [K, M]^T, then fold the bins back into the kernel
            xuT = self._buf(f"conv{i}.xuT16", (pl.k, M), torch.bfloat16)
            lib.asr_conv_unfold_t(ptr(c["xp"]), C.byref(g), ptr(xuT), M, st)
            dwt = self._buf(f"conv{i}.dwt", (W_out, pl.k), torch.float32)
            self._gemm(BF16, OUT_F32, W_out, pl.k, M, gT16, M, xuT, M, dwt, pl.k)
            lib.asr_conv_toeplitz_grad(ptr(dwt), pl.k, ptr(cs), C.byref(g), co, ptr(P.g(f"conv{i}.W")), ptr(P.g(f"conv{i}.b")), st)
            if i > 0:
                # dL/d(padded input) [N * t_padded, W_in] = view(zero-padded g)[., (dkt', f', co)] . w2^T  (stride 1 in time)
                assert g.st == 1, "the input gradient of a time-strided inner conv layer is not implemented"
                dxp = self._buf(f"conv{i}.dxp", (M, W_in), torch.float32)
                self._gemm(BF16, OUT_F32, M, W_in, g.kt * W_out, gbuf, W_out, c["w2"], g.kt * W_out, dxp, W_in)
                gout, g_ts, g_ns, g_row0 = dxp, 1, pl.t_padded, g.pt
        self._conv_zero_g = False

    # ---------------------------------------------------------------- forward
    def forward(self, feats_tm: torch.Tensor, training=False, masks=None, zmasks=None, input_mask=None) -> torch.Tensor:
        """feats_tm: f32 [T, N, F] time-major on device -> logits f32 [T, N, C].
        masks: {layer: {Wf,Wb [N,D], Uf,Ub [N,H]}} variational-dropout masks (training only); sampled when
        spec.dropout > 0 and none are given."""
        sp, P = self.spec, self.params
        T, N, Fd = feats_tm.shape
        assert Fd == sp.num_features and feats_tm.is_cuda and feats_tm.dtype == torch.float32
        if sp.conv_front:                               # [T, N, F] -> [T', N, F' * C] (utterances stay independent)
            Np = self._padded_batch(self.out_frames(T), N)
            if Np != N:
                fp = self._buf("feats_pad_conv", (T, Np, Fd), torch.float32)
                fp[:, :N].copy_(feats_tm)
                fp[:, N:].zero_()
                feats_tm = fp
            feats_tm = self._conv_forward(feats_tm.contiguous(), training)
            out = self._forward_padded(feats_tm, training, masks, zmasks, input_mask, N)
            self._conv_pad = (N, Np) if Np != N else None
            return out
        return self._forward_padded(feats_tm, training, masks, zmasks, input_mask, None)

    def out_frames(self, T):
        """frames the BiLSTM stack sees for T input frames (the conv front end strides over time)"""
        for (_, kt, _, st, _) in self.spec.conv_front or ():
            T = (T + 2 * ((kt - 1) // 2) - kt) // st + 1
        return T

    def out_lengths(self, in_len):
        """per-utterance frame counts behind the conv front end (index glue on a handful of integers)"""
        for (_, kt, _, st, _) in self.spec.conv_front or ():
            in_len = torch.div(in_len + 2 * ((kt - 1) // 2) - kt, st, rounding_mode="floor") + 1
        return in_len.clamp_min(0).to(torch.int32) if self.spec.conv_front else in_len

    def _forward_padded(self, feats_tm, training, masks, zmasks, input_mask, n_real):
        sp, P = self.spec, self.params
        T, N, Fd = feats_tm.shape
        if n_real is not None:                          # conv path: the batch is already padded; callers see n_real samples
            if self.logical_h:
                masks, zmasks = self._widen_masks(masks, zmasks)
            self._pad = (n_real, N) if n_real != N else None
            if masks is not None and n_real != N:
                masks = {l: {k: torch.cat([v, torch.ones(N - n_real, v.shape[1], dtype=v.dtype, device=v.device)])
                             for k, v in m.items() if k in ("Wf", "Wb", "Uf", "Ub")} for l, m in masks.items()}
            logits = self._forward(feats_tm, training, masks, zmasks, input_mask)
            self.last_logits = logits[:, :n_real].contiguous() if n_real != N else logits
            return self.last_logits
        if self.logical_h:                              # zero-padded width: widen caller-supplied masks (values irrelevant)
            masks, zmasks = self._widen_masks(masks, zmasks)
        # ragged batches (the last batch of an epoch, predict.py's batch of 1): the tensor-core recurrences work on
        # groups of 8 / 16 samples, so the batch is padded with zero utterances up to the next group boundary.  Utterances
        # are independent (no batch statistics anywhere on the path) and backward() pads dlogits with zero rows, so
        # the padding contributes nothing to any gradient; callers only ever see the first N samples.
        Np = self._padded_batch(T, N)
        self._pad = (N, Np) if Np != N else None
        if self._pad:
            fp = self._buf("feats_pad", (T, Np, Fd), torch.float32)
            fp[:, :N].copy_(feats_tm)
            fp[:, N:].zero_()
            if masks is not None:
                masks = {l: {k: torch.cat([v, torch.ones(Np - N, v.shape[1], dtype=v.dtype, device=v.device)])
                             for k, v in m.items() if k in ("Wf", "Wb", "Uf", "Ub")} for l, m in masks.items()}
            if input_mask is not None:                  # [T * N, D] time-major rows -> [T * Np, D]
                im = torch.ones(T, Np, input_mask.shape[1], dtype=input_mask.dtype, device=input_mask.device)
                im[:, :N].copy_(input_mask.view(T, N, -1))
                input_mask = im.view(T * Np, -1)
            self.last_logits = self._forward(fp, training, masks, zmasks, input_mask)[:, :N].contiguous()
        else:
            self.last_logits = self._forward(feats_tm, training, masks, zmasks, input_mask)
        return self.last_logits

    def _widen_masks(self, masks, zmasks):
        H, Hp = self.logical_h, self.spec.num_hiddens

        def widen(v, axis, blocks):
            if v.shape[axis] != blocks * H:
                return v                                # layer-0 input masks [N, F], or already device-width
            shp = list(v.shape)
            out = torch.ones(shp[:axis] + [blocks, Hp] + shp[axis + 1:], dtype=v.dtype, device=v.device)
            out.narrow(axis + 1, 0, H).copy_(v.reshape(shp[:axis] + [blocks, H] + shp[axis + 1:]))
            return out.reshape(shp[:axis] + [blocks * Hp] + shp[axis + 1:])

        if masks is not None:
            masks = {l: {k: (widen(v, 1, 1) if k in ("Uf", "Ub") else widen(v, 1, 2) if (k in ("Wf", "Wb") and l > 0) else v)
                         for k, v in m.items() if k in ("Wf", "Wb", "Uf", "Ub")} for l, m in masks.items()}
        if zmasks is not None:
            zmasks = {l: {k: widen(v, v.dim() - 1, 1) for k, v in m.items()} for l, m in zmasks.items()}
        return masks, zmasks

    def _padded_batch(self, T, N):
        """N itself when the tensor-core engine takes it (or cannot take the model at all); else the next group boundary."""
        sp = self.spec
        H = sp.hs[0]
        if not sp.general:
            if lib.asr_lstm_fuses_masks(T, N, H, self.lstm_opts):
                return N
            for Np in (_pad8(N), (N + 15) // 16 * 16):
                if Np != N and lib.asr_lstm_fuses_masks(T, Np, H, self.lstm_opts):
                    return Np
        # fp32 / general-cell paths: the dU GEMM reads h and dz shifted by one time step = N columns of the transposed
        # 16-bit copies, and a GEMM operand starts on a 16-byte boundary: N must be a multiple of 8 there as well
        return _pad8(N)

    def _forward(self, feats_tm, training, masks, zmasks, input_mask):
        sp, P = self.spec, self.params
        T, N, Fd = feats_tm.shape
        assert Fd == sp.lstm_in
        # the persistent engines cover the reference's shapes; anything else (e.g. H = 800) runs on the general cell
        # (the persistent engines hand 2H-wide fp16 rows straight to the next GEMM: 2H must be a multiple of 8 — 16-byte
        # rows — e.g. graves2006(num_hiddens=50) is not; the general path pads its operand copies instead)
        self._use_general = (sp.general or not lib.asr_lstm_persistent_supported(T, N, sp.hs[0], int(training), self.lstm_opts)
                             or (sp.elementwise and not lib.asr_lstm_fuses_variants(T, N, sp.hs[0], self.lstm_opts))
                             or (2 * sp.hs[0]) % 8 != 0)
        if self._use_general:
            return self._forward_general(feats_tm, training, masks, zmasks, input_mask)
        H, L, Cc = sp.hs[0], len(sp.hs), sp.num_classes
        R = T * N
        # 16-bit storage of zx / gates / cell + TMA staging: the default tensor-core recurrence where it takes the shape
        # (not when the recurrences share their SMs with another kernel — the pipelined evaluator's beam search: the
        # nine-warp TMA kernels lose more to the shared issue slots than the four-warp ones, 5 250 vs 6 395 clips/s at C5)
        self._fp16 = bool(self.fp16_storage and not sp.elementwise and not self.shared_sm and
                          lib.asr_lstm_fp16_storage(T, N, H, self.lstm_opts))
        zdt = OUT_F16 if self._fp16 else OUT_F32
        w = self._alloc(T, N, training)
        self._w, self._T, self._N = w, T, N
        if training and masks is None and sp.dropout > 0:
            masks = self.sample_masks(N)
        self._masks = masks if training else None
        masks = self._masks
        if training and zmasks is None and sp.zoneout > 0:
            zmasks = self.sample_zoneout_masks(T)
        self._zmasks = zmasks if training else None
        self._zpacked = {}
        st = cur_stream()
        feats_tm = feats_tm.contiguous()
        D0 = _pad8(Fd)
        main = torch.cuda.current_stream()
        prep_ev = None
        if not training and self._prepared_for == (self.params_version, False):
            pass                                         # inference on unchanged parameters: the 16-bit operands are in place
        elif self.overlap and L > 1 and training:
            # operands of the first layer's forward pass here; everything else on the side stream under its recurrence
            self._prep_weights(training, first_only=True)
            self._side.wait_stream(main)                 # the previous step's optimiser update / readers are behind us
            with torch.cuda.stream(self._side):
                self._prep_weights(training, skip_first=True)
                if training:
                    lib.asr_cast_transpose(ptr(feats_tm), Fd, ptr(w["xT16"]), R, R, Fd, BF16, cur_stream())
                prep_ev = torch.cuda.Event()
                prep_ev.record(self._side)
            feats_tm.record_stream(self._side)
        else:
            self._prep_weights(training)
            if training:
                lib.asr_cast_transpose(ptr(feats_tm), Fd, ptr(w["xT16"]), R, R, Fd, BF16, st)
        self._prepared_for = (self.params_version, bool(training))
        lib.asr_cast_rows(ptr(feats_tm), Fd, ptr(w["x16"]), D0, R, Fd, F16, st)
        x16, D = w["x16"], D0
        src, src_dt, src_ld, Dl = feats_tm, 2, Fd, Fd              # layer input before masking
        # with dropout the recurrence of layer l-1 writes the masked operand copies of layer l itself (fused side
        # stores) when the selected engine supports it; otherwise asr_mask_cast makes them
        fuse = masks is not None and bool(lib.asr_lstm_fuses_masks(T, N, H, self.lstm_opts))
        self._fused = fuse
        prev = None                                                 # fused outputs of the previous layer
        for l in range(L):
            if l == 1 and prep_ev is not None:
                main.wait_event(prep_ev)                   # the side-stream operand preparation (long finished by now)
            mask_u = None
            zx = w["zx"]
            if training and sp.mi is not None:             # the backward pass of MI re-reads every layer's Wx
                zx = w[f"zx.{l}"] = self._buf(f"zx.{l}", (R, 8 * H), torch.float32)
                w[f"uh.{l}"] = self._buf(f"uh.{l}", (R, 8 * H), torch.float32)
            if masks is None:
                self._gemm(F16, zdt, R, 8 * H, D, x16, D, self._ws[f"WcatT16.{l}"], D, zx, 8 * H)
            else:
                mk = masks[l]
                mask_u = self._packed(mk, "U")
                self._views[f"maskU.{l}"] = mask_u
                for i, d in enumerate("fb"):
                    if prev is not None:
                        xm = prev["hm16"][i]
                        if training:
                            self._views[f"xmT16.{l}.{i}"] = prev["hmT16"][i]
                    else:
                        mw = mk["W" + d].contiguous()
                        xm = self._buf(f"xm16.{i}", (R, D), torch.float16)
                        lib.asr_mask_cast(ptr(src), src_dt, src_ld, ptr(mw), N, ptr(xm), F16, D, R, Dl, 0, st)
                        if training:
                            xmT = self._buf(f"xmT16.{l}.{i}", (Dl, R), torch.bfloat16)
                            lib.asr_mask_cast(ptr(src), src_dt, src_ld, ptr(mw), N, ptr(xmT), BF16, R, R, Dl, 1, st)
                    lib.asr_gemm_tn(F16, zdt, R, 4 * H, D, ptr(xm), D, ptr(self._views[f"WcatT16.{l}"][i * 4 * H:]), D,
                                    ptr(zx[:, i * 4 * H:]), 8 * H, None, 1.0, 0, st)
            top = l == L - 1
            fz = dict(mask_next=None, hm16=None, hmT16=None, hT16u=None)
            cur = None
            if fuse and not top:
                cur = dict(hm16=self._buf(f"hm16.{l}", (2, R, 2 * H), torch.float16),
                           hmT16=self._buf(f"hmT16.{l}", (2, 2 * H, R), torch.bfloat16) if training else None)
                mnext = self._packed(masks[l + 1], "W")
                self._views[f"maskW.{l + 1}"] = mnext
                fz.update(mask_next=ptr(mnext).value, hm16=ptr(cur["hm16"]).value,
                          hmT16=ptr(cur["hmT16"]).value if training else None)
            if fuse and top and training:
                fz["hT16u"] = ptr(self._buf("topT16", (2 * H, R), torch.bfloat16)).value
            if sp.elementwise:                             # MI / zoneout switches of the tensor-core kernels
                fz["zoneout"] = float(sp.zoneout)
                if sp.mi is not None:
                    fz["mi"] = P.p(f"l{l}.mi_alpha").data_ptr()       # alpha | beta1 | beta2 are contiguous in the bucket
                    fz["uh"] = ptr(w[f"uh.{l}"]).value if training else None
                if self._zmasks is not None:
                    zp = self._zpacked[l] = torch.stack([self._zmasks[l]["h"], self._zmasks[l]["c"]]).contiguous()
                    fz["zmask"] = zp.data_ptr()
            gbuf = ptr(w[f"gates.{l}"]).value if training else None
            cbuf = ptr(w[f"cell.{l}"]).value if training else None
            st16 = dict(zx=None, zx16=ptr(zx).value, gates=None, cell=None, gates16=gbuf, cell16=cbuf) if self._fp16 else \
                dict(zx=ptr(zx).value, gates=gbuf, cell=cbuf)
            a = LstmFwdArgs(T=T, N=N, H=H, training=int(training),
                            bias=ptr(P.p(f"l{l}.bf")).value, U=ptr(P.p(f"l{l}.Uf")).value,
                            U16=ptr(self._ws[f"UT16.{l}"]).value,
                            h16=ptr(w[f"h16.{l}"]).value if (top or not fuse) else None,
                            hT16=ptr(w[f"hT16.{l}"]).value if training else None, h32=None,
                            flags=ptr(self._flags).value, mask_u=ptr(mask_u).value if mask_u is not None else None,
                            opts=self.lstm_opts, **st16, **fz)
            lib.asr_lstm_forward(C.byref(a), st)
            prev = cur
            x16, D = w[f"h16.{l}"], 2 * H
            src, src_dt, src_ld, Dl = w[f"h16.{l}"], 0, 2 * H, 2 * H
        self._gemm(F16, OUT_F32, R, Cc, 2 * H, x16, 2 * H, self._ws["WdT16"], 2 * H, w["logits"], Cc,
                   bias=P.p("dense.b"))
        return w["logits"]

    # ------------------------------------------------- general path (brsmv1 switches on)
    def sample_zoneout_masks(self, T):
        """One keep mask per (layer, direction, time step, unit), shared by the batch: K.dropout(h_diff, level,
        noise_shape=(output_dim,)) inside the step (core/layers_utils.py:34-42).  {layer: {h, c: f32 [2, T, H]}}."""
        sp = self.spec
        total = sum(2 * 2 * T * H for H in sp.hs)
        flat = self._buf("zoneout_masks", (total,), torch.float32)
        lib.asr_bernoulli_mask(ptr(flat), total, float(sp.zoneout), 1.0, self._mask_seed, self._mask_offset, cur_stream())
        self._mask_offset += total
        out, o = {}, 0
        for l, H in enumerate(sp.hs):
            out[l] = {}
            for k in ("h", "c"):
                out[l][k] = flat[o:o + 2 * T * H].view(2, T, H)
                o += 2 * T * H
        return out

    def _variant(self, l, zm):
        sp, P = self.spec, self.params

        def pp(name):
            return P.p(f"l{l}.{name}").data_ptr()

        kw = dict(ln_eps=1e-5, zoneout_h=float(sp.zoneout), zoneout_c=float(sp.zoneout))
        if sp.mi is not None:
            kw.update(mi_alpha=pp("mi_alpha"), mi_beta1=pp("mi_beta1"), mi_beta2=pp("mi_beta2"))
        if sp.layer_norm is not None:
            for n in ("uh", "wx", "c"):
                kw[f"ln_gain_{n}"], kw[f"ln_bias_{n}"] = pp(f"ln_gain_{n}"), pp(f"ln_bias_{n}")
        if zm is not None:
            kw.update(zmask_h=zm["h"].data_ptr(), zmask_c=zm["c"].data_ptr())
        return LstmVariant(**kw)

    def _operands(self, name, src32, Dl, mk, training, split=False):
        """16-bit GEMM operands of one layer input (fp32 [R, Dl]): fp16 [R, pad8(Dl)] for the projection and (training)
        its bf16 transpose [Dl, R] for dW; one pair per direction when the variational masks B_W differ.
        split=True adds the fp16 rounding residual as a third operand (split-precision projection)."""
        R, Dp, st = src32.shape[0], _pad8(Dl), cur_stream()
        out = []
        for i in range(2 if mk is not None else 1):
            a16 = self._buf(f"{name}.in16.{i}", (R, Dp), torch.float16, zero=True)
            aT = self._buf(f"{name}.inT16.{i}", (Dl, R), torch.bfloat16) if training else None
            if split:
                m32 = src32
                if mk is not None:
                    mw = mk["W" + "fb"[i]].contiguous()
                    m32 = self._buf(f"{name}.in32m.{i}", (R, Dl), torch.float32)
                    lib.asr_add_mask(ptr(src32), None, ptr(mw), mw.shape[0], ptr(m32), R, Dl, st)
                lo = self._buf(f"{name}.in16lo.{i}", (R, Dp), torch.float16, zero=True)
                lib.asr_cast_rows(ptr(m32), Dl, ptr(a16), Dp, R, Dl, F16, st)
                lib.asr_cast_rows(ptr(m32), Dl, ptr(lo), Dp, R, Dl, F16_LO, st)
                if training:
                    lib.asr_cast_transpose(ptr(m32), Dl, ptr(aT), R, R, Dl, BF16, st)
                out.append((a16, aT, lo))
                continue
            if mk is None:
                lib.asr_cast_rows(ptr(src32), Dl, ptr(a16), Dp, R, Dl, F16, st)
                if training:
                    lib.asr_cast_transpose(ptr(src32), Dl, ptr(aT), R, R, Dl, BF16, st)
            else:
                mw = mk["W" + "fb"[i]].contiguous()
                lib.asr_mask_cast(ptr(src32), 2, Dl, ptr(mw), mw.shape[0], ptr(a16), F16, Dp, R, Dl, 0, st)
                if training:
                    lib.asr_mask_cast(ptr(src32), 2, Dl, ptr(mw), mw.shape[0], ptr(aT), BF16, R, R, Dl, 1, st)
            out.append((a16, aT))
        return out

    def _forward_general(self, feats_tm, training, masks, zmasks, input_mask):
        sp, P = self.spec, self.params
        T, N, Fd = feats_tm.shape
        hs, Cc = sp.hs, sp.num_classes
        L, Hmax, PW = len(hs), max(hs), sp.proj_width
        R, st = T * N, cur_stream()
        if sp.residual not in (None, "sum"):
            raise NotImplementedError("residual merge mode %r: only 'sum' keeps the layer width (core/models.py:273-274)" % sp.residual)
        if sp.residual is not None and len(set(hs)) > 1:
            raise NotImplementedError("the residual merge needs equal layer widths")
        self._T, self._N = T, N
        if training and masks is None and sp.dropout > 0:
            masks = self.sample_masks(N)
        if training and zmasks is None and sp.zoneout > 0:
            zmasks = self.sample_zoneout_masks(T)
        if training and input_mask is None and sp.input_dropout and sp.dropout > 0:
            Din = PW or Fd
            input_mask = self._buf("input_mask", (R, Din), torch.float32)
            lib.asr_bernoulli_mask(ptr(input_mask), R * Din, float(sp.dropout), 1.0 / (1.0 - float(sp.dropout)), self._mask_seed,
                                   self._mask_offset, cur_stream())
            self._mask_offset += R * Din
        if not training:
            masks = zmasks = input_mask = None
        self._masks, self._zmasks, self._input_mask = masks, zmasks, input_mask
        self._prep_weights(training)
        w = self._w = {}
        x32 = feats_tm.contiguous().view(R, Fd)
        cur32, Dl = x32, Fd
        self._gen = dict(ops=[], x_ops=None)
        if PW:                                     # TimeDistributed(Dense(PW)): residual stack (2H) or eyben's input layer
            split = sp.layer_norm is not None      # everything that feeds a layer-normalised product runs split-precision
            self._gen["x_ops"] = self._operands("x", x32, Fd, None, training, split)[0]
            Fp = _pad8(Fd)
            wp = self._buf("WpT16", (PW, Fp), torch.float16, zero=True)
            lib.asr_cast_transpose(ptr(P.p("proj.W")), PW, ptr(wp), Fp, Fd, PW, F16, st)
            cur32 = self._buf("res32.in", (R, PW), torch.float32)
            self._gemm(F16, OUT_F32, R, PW, Fp, self._gen["x_ops"][0], Fp, wp, Fp, cur32, PW, bias=P.p("proj.b"))
            if split:
                wpl = self._buf("WpT16lo", (PW, Fp), torch.float16, zero=True)
                lib.asr_cast_transpose(ptr(P.p("proj.W")), PW, ptr(wpl), Fp, Fd, PW, F16_LO, st)
                self._gemm(F16, OUT_F32, R, PW, Fp, self._gen["x_ops"][0], Fp, wpl, Fp, cur32, PW, acc=1)
                self._gemm(F16, OUT_F32, R, PW, Fp, self._gen["x_ops"][2], Fp, wp, Fp, cur32, PW, acc=1)
            Dl = PW
        if input_mask is not None:
            dst = cur32 if PW else self._buf("x32.masked", (R, Dl), torch.float32)
            lib.asr_add_mask(ptr(cur32), None, ptr(input_mask.contiguous()), R, ptr(dst), R, Dl, st)
            cur32 = dst
        for l, H in enumerate(hs):
            w["zx"] = self._buf("zx", (R, 8 * H), torch.float32)
            mk = masks[l] if masks is not None else None
            # layer normalisation divides Wx by its row deviation, which amplifies the fp16 operand rounding of the
            # projection ~16x (measured 4.6e-3 on the logits, CPU emulation 4.56e-3): with LN on, the projection runs
            # split-precision (hi*hi + hi*lo + lo*hi, three tensor-core GEMMs accumulating in fp32)
            split = sp.layer_norm is not None
            ops = self._operands(f"l{l}", cur32, Dl, mk, training, split)
            self._gen["ops"].append(ops)
            Dp = _pad8(Dl)
            wt = self._views[f"WcatT16.{l}"]
            wl = self._views[f"WcatT16lo.{l}"] if split else None
            for i in range(2 if mk is not None else 1):
                ncol = 4 * H if mk is not None else 8 * H
                cdst = w["zx"][:, i * 4 * H:] if mk is not None else w["zx"]
                lib.asr_gemm_tn(F16, OUT_F32, R, ncol, Dp, ptr(ops[i][0]), Dp, ptr(wt[i * 4 * H:]), Dp, ptr(cdst), 8 * H,
                                None, 1.0, 0, st)
                if split:
                    lib.asr_gemm_tn(F16, OUT_F32, R, ncol, Dp, ptr(ops[i][0]), Dp, ptr(wl[i * 4 * H:]), Dp, ptr(cdst), 8 * H,
                                    None, 1.0, 1, st)
                    lib.asr_gemm_tn(F16, OUT_F32, R, ncol, Dp, ptr(ops[i][2]), Dp, ptr(wt[i * 4 * H:]), Dp, ptr(cdst), 8 * H,
                                    None, 1.0, 1, st)
            mask_u = None
            if mk is not None:
                mask_u = self._buf(f"maskU.{l}", (2, N, H), torch.float32)
                mask_u[0].copy_(mk["Uf"]); mask_u[1].copy_(mk["Ub"])
            h32 = self._buf(f"h32.{l}", (R, 2 * H), torch.float32)
            w[f"h32.{l}"] = h32
            if training:
                # zx is overwritten by the next layer, the backward pass needs every layer's: keep per-layer copies
                w[f"zx.{l}"] = self._buf(f"zx.{l}", (R, 8 * H), torch.float32)
                w[f"zx.{l}"].copy_(w["zx"])
                w[f"hT16.{l}"] = self._buf(f"hT16.{l}", (2 * H, R), torch.bfloat16)
                w[f"gates.{l}"] = self._buf(f"gates.{l}", (R, 8 * H), torch.float32)
                w[f"cell.{l}"] = self._buf(f"cell.{l}", (R, 2 * H), torch.float32)
                w[f"uh.{l}"] = self._buf(f"uh.{l}", (R, 8 * H), torch.float32)
            a = LstmFwdArgs(T=T, N=N, H=H, training=int(training), zx=ptr(w["zx"]).value,
                            bias=ptr(P.p(f"l{l}.bf")).value, U=ptr(P.p(f"l{l}.Uf")).value, U16=None, h16=None,
                            hT16=ptr(w[f"hT16.{l}"]).value if training else None, h32=ptr(h32).value,
                            gates=ptr(w[f"gates.{l}"]).value if training else None,
                            cell=ptr(w[f"cell.{l}"]).value if training else None,
                            flags=ptr(self._flags).value, mask_u=ptr(mask_u).value if mask_u is not None else None)
            v = self._variant(l, zmasks[l] if zmasks is not None else None)
            lib.asr_lstm_cell_forward(C.byref(a), C.byref(v), ptr(w[f"uh.{l}"]) if training else None, st)
            if sp.residual is not None:
                nxt = self._buf(f"res32.{l}", (R, 2 * H), torch.float32)
                lib.asr_add_mask(ptr(h32), ptr(cur32), None, 1, ptr(nxt), R, 2 * H, st)
                cur32 = nxt
            else:
                cur32 = h32
            Dl = 2 * H
        H2 = 2 * hs[-1]
        top = self._operands("top", cur32, H2, None, training)[0]
        self._gen["top"] = top
        w["logits"] = self._buf("logits", (T, N, Cc), torch.float32)
        self._gemm(F16, OUT_F32, R, Cc, _pad8(H2), top[0], _pad8(H2), self._views["WdT16"], _pad8(H2), w["logits"], Cc,
                   bias=P.p("dense.b"))
        if training:
            w["dlogits"] = self._buf("dlogits", (T, N, Cc), torch.float32)
            w["loss"] = self._buf("loss", (N,), torch.float32)
        return w["logits"]

    def _backward_general(self, dlogits):
        sp, P, w = self.spec, self.params, self._w
        T, N = self._T, self._N
        hs, Cc = sp.hs, sp.num_classes
        L, H2 = len(hs), 2 * hs[-1]
        R, st, cp = T * N, cur_stream(), _pad8(Cc)
        masks, zmasks = self._masks, self._zmasks
        dl16 = self._buf("dl16", (R, cp), torch.bfloat16, zero=True)
        dlT16 = self._buf("dlT16", (Cc, R), torch.bfloat16)
        lib.asr_cast_rows(ptr(dlogits), Cc, ptr(dl16), cp, R, Cc, BF16, st)
        lib.asr_cast_transpose(ptr(dlogits), Cc, ptr(dlT16), R, R, Cc, BF16, st)
        lib.asr_colsum(ptr(dlogits), Cc, R, Cc, ptr(P.g("dense.b")), st)
        self._gemm(BF16, OUT_F32, H2, Cc, R, self._gen["top"][1], R, dlT16, R, P.g("dense.W"), Cc)
        dcur, dname = self._buf("dhA", (R, H2), torch.float32), "dhA"
        self._gemm(BF16, OUT_F32, R, H2, cp, dl16, cp, self._views["Wd16"], cp, dcur, H2)
        for l in range(L - 1, -1, -1):
            H = hs[l]
            dwx = self._buf("dwx32", (R, 8 * H), torch.float32)
            duh = self._buf("duh32", (R, 8 * H), torch.float32)
            dwx16 = self._buf("dwx16", (R, 8 * H), torch.bfloat16)
            dwxT16 = self._buf("dwxT16", (8 * H, R), torch.bfloat16)
            duhT16 = self._buf("duhT16", (8 * H, R), torch.bfloat16)
            mk = masks[l] if masks is not None else None
            mask_u = self._views[f"maskU.{l}"] if mk is not None else None
            b = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dcur).value, gates=ptr(w[f"gates.{l}"]).value,
                            cell=ptr(w[f"cell.{l}"]).value, U=ptr(P.p(f"l{l}.Uf")).value, U16=None, dz16=None, dzT16=None,
                            dz32=ptr(dwx).value, dbias=ptr(P.g(f"l{l}.bf")).value, flags=ptr(self._flags).value,
                            mask_u=ptr(mask_u).value if mask_u is not None else None)
            v = self._variant(l, zmasks[l] if zmasks is not None else None)
            gk = {}
            if sp.mi is not None:
                gk.update({n: P.g(f"l{l}.{n}").data_ptr() for n in ("mi_alpha", "mi_beta1", "mi_beta2")})
            if sp.layer_norm is not None:
                for n in ("uh", "wx", "c"):
                    gk[f"ln_gain_{n}"], gk[f"ln_bias_{n}"] = P.g(f"l{l}.ln_gain_{n}").data_ptr(), P.g(f"l{l}.ln_bias_{n}").data_ptr()
            g = LstmVariantGrads(**gk)
            lib.asr_lstm_cell_backward(C.byref(b), C.byref(v), ptr(w[f"zx.{l}"]), ptr(w[f"uh.{l}"]), ptr(duh), C.byref(g), st)
            lib.asr_cast_rows(ptr(dwx), 8 * H, ptr(dwx16), 8 * H, R, 8 * H, BF16, st)
            lib.asr_cast_transpose(ptr(dwx), 8 * H, ptr(dwxT16), R, R, 8 * H, BF16, st)
            lib.asr_cast_transpose(ptr(duh), 8 * H, ptr(duhT16), R, R, 8 * H, BF16, st)
            ops = self._gen["ops"][l]
            Dl = ops[0][1].shape[0]
            hT = w[f"hT16.{l}"]
            for i, d in enumerate("fb"):
                xT = ops[i if mk is not None else 0][1]
                lib.asr_gemm_tn(BF16, OUT_F32, Dl, 4 * H, R, ptr(xT), R, ptr(dwxT16[i * 4 * H:]), R,
                                ptr(P.g(f"l{l}.W{d}")), 4 * H, None, 1.0, 0, st)
                if T > 1:
                    Kk = (T - 1) * N
                    hA, dzB = hT[i * H:(i + 1) * H], duhT16[i * 4 * H:(i + 1) * 4 * H]
                    Ap, Bp = (hA, dzB[:, N:]) if i == 0 else (hA[:, N:], dzB)
                    lib.asr_gemm_tn(BF16, OUT_F32, H, 4 * H, Kk, C.c_void_p(Ap.data_ptr()), R, C.c_void_p(Bp.data_ptr()), R,
                                    ptr(P.g(f"l{l}.U{d}")), 4 * H, None, 1.0, 0, st)
                else:
                    P.g(f"l{l}.U{d}").zero_()
            need_dx = l > 0 or bool(sp.proj_width)
            if need_dx:
                wc = self._views[f"Wcat16.{l}"]
                res = sp.residual is not None
                if mk is None:
                    # dX = dWx . Wcat^T ; with the residual merge the identity branch adds d(o_{l+1}) (GEMM accumulate)
                    if res:
                        lib.asr_gemm_tn(BF16, OUT_F32, R, Dl, 8 * H, ptr(dwx16), 8 * H, ptr(wc), 8 * H, ptr(dcur), Dl, None, 1.0, 1, st)
                    else:
                        dname = "dhB" if dname == "dhA" else "dhA"
                        nxt = self._buf(dname, (R, Dl), torch.float32)
                        lib.asr_gemm_tn(BF16, OUT_F32, R, Dl, 8 * H, ptr(dwx16), 8 * H, ptr(wc), 8 * H, ptr(nxt), Dl, None, 1.0, 0, st)
                        dcur = nxt
                else:
                    part = [self._buf(f"dxpart.{i}", (R, Dl), torch.float32) for i in range(2)]
                    for i in range(2):
                        lib.asr_gemm_tn(BF16, OUT_F32, R, Dl, 4 * H, ptr(dwx16[:, i * 4 * H:]), 8 * H, ptr(wc[:, i * 4 * H:]), 8 * H,
                                        ptr(part[i]), Dl, None, 1.0, 0, st)
                    comb = self._buf("dxcomb", (R, Dl), torch.float32)
                    lib.asr_mask_combine(ptr(part[0]), ptr(part[1]), ptr(mk["Wf"].contiguous()), ptr(mk["Wb"].contiguous()), N,
                                         ptr(comb), R, Dl, st)
                    if res:
                        lib.asr_add_mask(ptr(dcur), ptr(comb), None, 1, ptr(dcur), R, Dl, st)
                    else:
                        dcur = comb
        PW = sp.proj_width
        if PW:
            if self._input_mask is not None:
                lib.asr_add_mask(ptr(dcur), None, ptr(self._input_mask.contiguous()), R, ptr(dcur), R, PW, st)
            dT = self._buf("dprojT16", (PW, R), torch.bfloat16)
            lib.asr_cast_transpose(ptr(dcur), PW, ptr(dT), R, R, PW, BF16, st)
            Fd = sp.num_features
            lib.asr_gemm_tn(BF16, OUT_F32, Fd, PW, R, ptr(self._gen["x_ops"][1]), R, ptr(dT), R, ptr(P.g("proj.W")), PW,
                            None, 1.0, 0, st)
            lib.asr_colsum(ptr(dcur), PW, R, PW, ptr(P.g("proj.b")), st)

    # ------------------------------------------------------------------- CTC
    def ctc(self, logits, in_len, labels_flat, label_off, max_label_len, grad_scale=1.0, want_grad=True):
        T, N, Cc = logits.shape
        w = self._w
        wsb = lib.asr_ctc_workspace_bytes(T, N, max_label_len)
        ws = self._buf("ctc_ws", (wsb // 4 + 1,), torch.float32)
        loss = self._buf("loss", (N,), torch.float32)      # shaped by the logits handed in (forward() may have padded)
        grad = self._buf("dlogits", (T, N, Cc), torch.float32)
        lib.asr_ctc_loss_grad(ptr(logits), T, N, Cc, ptr(in_len), ptr(labels_flat), ptr(label_off), max_label_len,
                              Cc - 1, float(grad_scale), ptr(loss), ptr(grad), ptr(ws), cur_stream())
        return loss, grad

    def greedy(self, logits, in_len, merge_repeated=True):
        T, N, Cc = logits.shape
        out = self._buf("greedy_out", (N, T), torch.int32)
        out_len = self._buf("greedy_len", (N,), torch.int32)
        lib.asr_ctc_greedy(ptr(logits), T, N, Cc, ptr(in_len), Cc - 1, int(merge_repeated), ptr(out), ptr(out_len),
                           cur_stream())
        return out, out_len

    def ler(self, hyp, hyp_len, labels_flat, label_off, max_label_len, normalize=True):
        """core/metrics.py:4-8 on the device: per-utterance tf.edit_distance(hyp, truth, normalize=True) of a decode
        kernel's output [N, T] / [N] against the sparse labels -> f32 [N] (no host round trip)."""
        N = hyp.shape[0]
        out = self._buf("ler_out", (N,), torch.float32)
        lib.asr_edit_distance(ptr(hyp), N, hyp.shape[1], ptr(hyp_len), ptr(labels_flat), ptr(label_off),
                              int(max_label_len), int(normalize), ptr(out), cur_stream())
        return out

    def beam(self, logits, in_len, beam_width=100, merge_repeated=True, tag=""):
        """tag: suffix of the workspace / output buffer names, so that two searches can be in flight on two streams
        (a pipelined evaluator decodes group k while group k+1 runs forward)."""
        T, N, Cc = logits.shape
        wsb = lib.asr_ctc_beam_workspace_bytes(T, N, Cc, beam_width)
        ws = self._buf("beam_ws" + tag, (wsb // 4 + 1,), torch.int32)
        out = self._buf("beam_out" + tag, (N, T), torch.int32)
        out_len = self._buf("beam_len" + tag, (N,), torch.int32)
        lib.asr_ctc_beam(ptr(logits), T, N, Cc, ptr(in_len), Cc - 1, beam_width, int(merge_repeated), ptr(out),
                         ptr(out_len), ptr(ws), cur_stream())
        return out, out_len

    # --------------------------------------------------------------- backward
    def _layer_slice(self, l):
        """flat-bucket range of layer l's parameters (Wf .. bb are contiguous per layer)."""
        P, sp = self.params, self.spec
        lo = P.offsets[f"l{l}.Wf"]
        hi = P.offsets[f"l{l + 1}.Wf"] if l + 1 < sp.num_layers else P.offsets["dense.W"]
        return lo, hi

    def backward(self, dlogits: torch.Tensor, allreduce=None):
        """dlogits f32 [T,N,C] (already scaled by 1/global_batch) -> fills params.grad.
        allreduce (data parallel): called ONCE, on the whole flat gradient bucket, behind the last GEMM of the backward
        pass when the bucket is small (< DP_SLICE_BYTES; self.dp_slices = False forces it); for a large bucket (or with
        self.dp_slices = True) it is called on slices of the bucket as soon as they are complete — the
        Dense slice behind its own GEMM, layer l's [Wf|Wb] behind its two dW GEMMs and the rest of the layer behind its dU
        GEMMs (while the BPTT of layer l-1 runs), layer 0 as one slice; the slices tile the bucket exactly once.  Measured on
        2 and on 8 B200s the single collective wins for C2's 59 MB: the step is short enough that
        NCCL's CTAs, the background GEMMs and the next BPTT's cooperative launch fight over the same 20 idle SMs.
        Returns the handles (objects with .wait()) the callable returned, if any."""
        sp, P, w = self.spec, self.params, self._w
        handles = []
        # one collective for a bucket the links move in a fraction of a millisecond (C2: 59 MB), slices overlapped with the
        # BPTT for a large one (configs[3]: 277 MB, 0.9 ms at 2 GPUs if left to the end) — both measured, DESIGN.md section 6
        slices = self.dp_slices if self.dp_slices is not None else (P.numel * 4 > DP_SLICE_BYTES)
        whole, allreduce = allreduce, (allreduce if slices else None)
        if getattr(self, "_pad", None):                 # forward() padded the batch: zero gradient rows for the padding
            n, npad = self._pad
            dl = self._buf("dlogits_pad", (dlogits.shape[0], npad, dlogits.shape[2]), torch.float32)
            dl[:, :n].copy_(dlogits)
            dl[:, n:].zero_()
            dlogits = dl
        if self._use_general:
            self._backward_general(dlogits)
            if whole is not None:
                handles.append(whole(P.grad))
            return handles
        T, N = self._T, self._N
        H, L, Cc = sp.hs[0], len(sp.hs), sp.num_classes
        R = T * N
        st = cur_stream()
        cp = _pad8(Cc)
        lib.asr_cast_rows(ptr(dlogits), Cc, ptr(w["dl16"]), cp, R, Cc, BF16, st)
        lib.asr_cast_transpose(ptr(dlogits), Cc, ptr(w["dlT16"]), R, R, Cc, BF16, st)
        lib.asr_colsum(ptr(dlogits), Cc, R, Cc, ptr(P.g("dense.b")), st)
        top = L - 1
        topT = w[f"hT16.{top}"]
        fuse = self._masks is not None and getattr(self, "_fused", False)
        if self._masks is not None and fuse:
            topT = self._views["topT16"]          # unmasked transposed copy written by the top layer's recurrence
        elif self._masks is not None:
            # with dropout hT16 holds h * B_U (the dU operand); the Dense kernel saw the unmasked h
            ones = self._buf("ones_mask", (N, 2 * H), torch.float32)
            ones.fill_(1.0)
            topT = self._buf("topT16", (2 * H, R), torch.bfloat16)
            lib.asr_mask_cast(ptr(w[f"h16.{top}"]), 0, 2 * H, ptr(ones), N, ptr(topT), BF16, R, R, 2 * H, 1, st)
        # dWd [2H, C] = topT [2H, R] . dlT [C, R]^T
        self._gemm(BF16, OUT_F32, 2 * H, Cc, R, topT, R, w["dlT16"], R, P.g("dense.W"), Cc)
        if allreduce is not None:                       # the Dense slice is complete: reduce it under the whole BPTT
            handles.append(allreduce(P.grad[P.offsets["dense.W"]:]))
        # dTop [R, 2H] = dl16 [R, Cpad] . Wd16 [2H, Cpad]^T
        dh, other = w["dhA"], w["dhB"]
        self._gemm(BF16, OUT_F32, R, 2 * H, cp, w["dl16"], cp, self._ws["Wd16"], cp, dh, 2 * H)
        masks = self._masks
        dh2, mask_dh = None, None                  # fused path: the layer above left its dX as two masked partials
        for l in range(L - 1, -1, -1):
            mask_u = self._views[f"maskU.{l}"] if masks is not None else None
            vz = {}
            if sp.elementwise:
                vz["zoneout"] = float(sp.zoneout)
                if sp.mi is not None:
                    duhT = self._buf(f"duhT16.{l}", (8 * H, R), torch.bfloat16)
                    vz.update(mi=P.p(f"l{l}.mi_alpha").data_ptr(), zx=ptr(w[f"zx.{l}"]).value, uh=ptr(w[f"uh.{l}"]).value,
                              dmi=P.g(f"l{l}.mi_alpha").data_ptr(), duhT16=ptr(duhT).value)
                if self._zmasks is not None:
                    vz["zmask"] = self._zpacked[l].data_ptr()
            st16 = dict(gates16=ptr(w[f"gates.{l}"]).value, cell16=ptr(w[f"cell.{l}"]).value) if self._fp16 else \
                dict(gates=ptr(w[f"gates.{l}"]).value, cell=ptr(w[f"cell.{l}"]).value)
            a = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh).value, opts=self.lstm_opts, **st16, U=ptr(P.p(f"l{l}.Uf")).value,
                            U16=ptr(self._ws[f"Ub16.{l}"]).value,
                            dz16=ptr(w[f"dz16.{l}"]).value, dzT16=ptr(w[f"dzT16.{l}"]).value, dz32=None,
                            dbias=ptr(P.g(f"l{l}.bf")).value, flags=ptr(self._flags).value,
                            mask_u=ptr(mask_u).value if mask_u is not None else None,
                            dh2=ptr(dh2).value if dh2 is not None else None,
                            mask_dh=ptr(mask_dh).value if mask_dh is not None else None, **vz)
            lib.asr_lstm_backward(C.byref(a), st)
            dh2, mask_dh = None, None
            D = sp.lstm_in if l == 0 else 2 * H
            xT = w["xT16"] if l == 0 else w[f"hT16.{l - 1}"]
            hT = w[f"hT16.{l}"]
            dzT = w[f"dzT16.{l}"]
            def input_gradient():
                nonlocal dh, other, dh2, mask_dh
                if l == 0 and sp.conv_front:
                    # dL/d(conv output) [R, D] = dz . Wcat^T (per direction with the fused dropout masks), then the conv backward
                    dx0 = self._buf("dx0", (R, D), torch.float32)
                    wc = self._views[f"Wcat16.{l}"]
                    if masks is None:
                        self._gemm(BF16, OUT_F32, R, D, 8 * H, w[f"dz16.{l}"], 8 * H, wc, 8 * H, dx0, D)
                    else:
                        part = [self._buf(f"dx0part.{i}", (R, D), torch.float32) for i in range(2)]
                        for i in range(2):
                            lib.asr_gemm_tn(BF16, OUT_F32, R, D, 4 * H, ptr(w[f"dz16.{l}"][:, i * 4 * H:]), 8 * H,
                                            ptr(wc[:, i * 4 * H:]), 8 * H, ptr(part[i]), D, None, 1.0, 0, st)
                        mk = masks[l]
                        lib.asr_mask_combine(ptr(part[0]), ptr(part[1]), ptr(mk["Wf"].contiguous()), ptr(mk["Wb"].contiguous()), N,
                                             ptr(dx0), R, D, st)
                    self._conv_backward(dx0)
                elif l > 0 and masks is None:
                    # dX [R, 2H] = dz16 [R, 8H] . Wcat16 [2H, 8H]^T
                    self._gemm(BF16, OUT_F32, R, 2 * H, 8 * H, w[f"dz16.{l}"], 8 * H, self._ws[f"Wcat16.{l}"], 8 * H,
                               other, 2 * H)
                    dh, other = other, dh
                elif l > 0:
                    # dX = (dz_f . Wf^T) * B_Wf + (dz_b . Wb^T) * B_Wb   (each direction's LSTM masked its own input)
                    part = [self._buf(f"dxpart.{i}", (R, 2 * H), torch.float32) for i in range(2)]
                    wc = self._views[f"Wcat16.{l}"]
                    for i in range(2):
                        lib.asr_gemm_tn(BF16, OUT_F32, R, 2 * H, 4 * H, ptr(w[f"dz16.{l}"][:, i * 4 * H:]), 8 * H,
                                        ptr(wc[:, i * 4 * H:]), 8 * H, ptr(part[i]), 2 * H, None, 1.0, 0, st)
                    mk = masks[l]
                    if fuse:                            # the BPTT kernel of layer l-1 combines the partials with B_Wf / B_Wb
                        dh, dh2, mask_dh = part[0], part[1], self._views[f"maskW.{l}"]
                    else:
                        lib.asr_mask_combine(ptr(part[0]), ptr(part[1]), ptr(mk["Wf"].contiguous()), ptr(mk["Wb"].contiguous()), N,
                                             ptr(other), R, 2 * H, st)
                        dh, other = other, dh

            def weight_gradients():
                main = torch.cuda.current_stream()
                side = self._side if self.overlap else main
                if side is not main:
                    # behind the dX GEMMs above: they feed the next BPTT, and a background CTA that got its SM first holds it for
                    # a whole K = T*N tile (the second dX GEMM ran 0.20 instead of 0.11 ms when both started together)
                    side.wait_stream(main)
                with torch.cuda.stream(side):
                    sst = cur_stream()
                    bg = GEMM_BACKGROUND if (side is not main and l > 0) else 0      # runs beside the next BPTT
                    lo, hi = self._layer_slice(l)
                    if l == 0 and not sp.conv_front:
                        lo = 0
                    mid = P.offsets[f"l{l}.Uf"]
                    for i, d in enumerate("fb"):
                        # dW_dir [D, 4H] = (x * B_W)^T [D, R] . dzT_dir [4H, R]^T
                        xTd = xT if masks is None else self._views[f"xmT16.{l}.{i}"]
                        lib.asr_gemm_tn_ex(BF16, OUT_F32, D, 4 * H, R, ptr(xTd), R, ptr(dzT[i * 4 * H:]), R,
                                           ptr(P.g(f"l{l}.W{d}")), 4 * H, None, 1.0, 0, bg, sst)
                    if allreduce is not None and l > 0:     # [Wf | Wb] is complete on this stream: reduce it under the dU GEMMs
                        handles.append(allreduce(P.grad[lo:mid]))  # (layer 0 is the exposed tail: one collective, one latency)
                    for i, d in enumerate("fb"):
                        # dU_dir [H, 4H] = h_prev^T . dz  with the one-step time shift of the recurrence
                        if T > 1:
                            Kk = (T - 1) * N
                            hA = hT[i * H:(i + 1) * H]
                            dzB = (self._views[f"duhT16.{l}"] if sp.mi is not None else dzT)[i * 4 * H:(i + 1) * 4 * H]
                            if i == 0:   # forward direction: h_{t-1} with dz_t
                                Ap, Bp = hA, dzB[:, N:]
                            else:        # reverse direction: h_{t+1} with dz_t
                                Ap, Bp = hA[:, N:], dzB
                            lib.asr_gemm_tn_ex(BF16, OUT_F32, H, 4 * H, Kk, C.c_void_p(Ap.data_ptr()), R,
                                               C.c_void_p(Bp.data_ptr()), R, ptr(P.g(f"l{l}.U{d}")), 4 * H, None, 1.0, 0, bg, sst)
                        else:
                            P.g(f"l{l}.U{d}").zero_()
                    if allreduce is not None:               # [Uf | Ub | biases (| switches)]: the rest of layer l's slice
                        handles.append(allreduce(P.grad[(mid if l > 0 else lo):hi]))

            if l == 0:              # nothing follows but the conv front end's backward pass: let it run beside these GEMMs
                weight_gradients()
                input_gradient()
            else:                   # the dX GEMMs feed the next BPTT: first, alone on the machine
                input_gradient()
                weight_gradients()
        if self.overlap:
            torch.cuda.current_stream().wait_stream(self._side)
        if allreduce is not None and sp.conv_front:     # the conv front end's gradients: the head of the bucket
            handles.append(allreduce(P.grad[:self._layer_slice(0)[0]]))
        if whole is not None and allreduce is None:     # the default: one collective of the whole bucket
            handles.append(whole(P.grad))
        return handles

    # -------------------------------------------------------------- optimiser
    def optimizer_step(self, lr=1e-3, clipnorm=400.0, beta1=0.9, beta2=0.999, eps=1e-8, opt="adam",
                       momentum=0.9):
        P, st = self.params, cur_stream()
        wd = float(self.spec.weight_decay)
        mask = ptr(P.decay) if wd else None
        lib.asr_grad_sqnorm(ptr(P.grad), ptr(P.flat), mask, P.numel, 1.0, wd, ptr(self._sqnorm), st)
        self.step_count += 1
        if opt == "adam":
            lib.asr_adam_step(ptr(P.flat), ptr(P.grad), ptr(P.m), ptr(P.v), mask, P.numel, 1.0, wd,
                              ptr(self._sqnorm), float(clipnorm or 0.0), float(lr), beta1, beta2, eps,
                              self.step_count, st)
        else:
            lib.asr_sgd_step(ptr(P.flat), ptr(P.grad), ptr(P.m), mask, P.numel, 1.0, wd, ptr(self._sqnorm),
                             float(clipnorm or 0.0), float(lr), float(momentum), st)

    def grad_norm(self) -> float:
        return math.sqrt(float(self._sqnorm.item()))

    def lstm_status(self) -> int:
        """0 = ok; non-zero = a persistent kernel's watchdog fired (see ASR_ERR_TIMEOUT)."""
        return int(self._flags[64].item())

    # ------------------------------------------------------------ whole step
    def train_step(self, feats_tm, in_len, labels_flat, label_off, max_label_len, global_batch=None,
                   allreduce=None, masks=None, zmasks=None, input_mask=None, **opt):
        """One optimisation step on time-major features; returns the per-utterance CTC loss tensor [N]."""
        N = feats_tm.shape[1]
        caller = torch.cuda.current_stream()
        main = self._main if self.overlap else caller
        if main is not caller:
            main.wait_stream(caller)
        with torch.cuda.stream(main):
            logits = self.forward(feats_tm, training=True, masks=masks, zmasks=zmasks, input_mask=input_mask)
            in_len = self.out_lengths(in_len)
            loss, dlogits = self.ctc(logits, in_len, labels_flat, label_off, max_label_len,
                                     grad_scale=1.0 / float(global_batch or N))
            for h in self.backward(dlogits, allreduce=allreduce):
                if h is not None and hasattr(h, "wait"):
                    h.wait()                            # async collectives: the current stream waits for them
            self.optimizer_step(**opt)
        if main is not caller:
            caller.wait_stream(main)
        return loss


def pack_labels(labels, device):
    """list of int sequences -> (flat i32, offsets i32 [N+1], max_len) on device
    (the sparse-label contract of datasets/dataset_generator.py:237-251)."""
    lens = [len(l) for l in labels]
    off = np.zeros(len(labels) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    flat = np.concatenate([np.asarray(l, dtype=np.int32) for l in labels]) if sum(lens) else np.zeros(1, np.int32)
    return (torch.as_tensor(flat, device=device), torch.as_tensor(off, device=device), int(max(lens) if lens else 0))
