"""ctypes binding of libasr_b200.so (the C ABI in include/asr_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an
exception is raised.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
(or ``make -C asr-study_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ASR_B200_LIB selects another build of the SAME library (e.g. the -DASR_LSTM_PROFILE build); never a fallback
LIB_PATH = os.environ.get("ASR_B200_LIB") or os.path.join(_HERE, "libasr_b200.so")


class AsrError(RuntimeError):
    pass


class MfccConfig(C.Structure):
    _fields_ = [("fs", C.c_float), ("win_len", C.c_float), ("win_step", C.c_float),
                ("num_filt", C.c_int32), ("nfft", C.c_int32), ("low_freq", C.c_float),
                ("high_freq", C.c_float), ("pre_emph", C.c_float), ("kind", C.c_int32),
                ("num_cep", C.c_int32), ("cep_lifter", C.c_int32), ("append_energy", C.c_int32),
                ("d", C.c_int32), ("dd", C.c_int32), ("mean_norm", C.c_int32), ("var_norm", C.c_int32),
                ("eps", C.c_float), ("stride", C.c_int32), ("num_context", C.c_int32)]


class LstmFwdArgs(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("H", C.c_int32), ("training", C.c_int32),
                ("zx", C.c_void_p), ("bias", C.c_void_p), ("U", C.c_void_p), ("U16", C.c_void_p),
                ("h16", C.c_void_p), ("hT16", C.c_void_p), ("h32", C.c_void_p),
                ("gates", C.c_void_p), ("cell", C.c_void_p), ("flags", C.c_void_p), ("mask_u", C.c_void_p),
                ("mask_next", C.c_void_p), ("hm16", C.c_void_p), ("hmT16", C.c_void_p), ("hT16u", C.c_void_p),
                ("mi", C.c_void_p), ("uh", C.c_void_p), ("zoneout", C.c_float), ("zmask", C.c_void_p),
                ("zx16", C.c_void_p), ("gates16", C.c_void_p), ("cell16", C.c_void_p), ("opts", C.c_int32)]


class LstmBwdArgs(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("H", C.c_int32),
                ("dh", C.c_void_p), ("gates", C.c_void_p), ("cell", C.c_void_p),
                ("U", C.c_void_p), ("U16", C.c_void_p), ("dz16", C.c_void_p), ("dzT16", C.c_void_p),
                ("dz32", C.c_void_p), ("dbias", C.c_void_p), ("flags", C.c_void_p), ("mask_u", C.c_void_p),
                ("dh2", C.c_void_p), ("mask_dh", C.c_void_p),
                ("mi", C.c_void_p), ("zx", C.c_void_p), ("uh", C.c_void_p), ("dmi", C.c_void_p), ("duhT16", C.c_void_p),
                ("zoneout", C.c_float), ("zmask", C.c_void_p),
                ("gates16", C.c_void_p), ("cell16", C.c_void_p), ("opts", C.c_int32)]


LSTM_SHARED_SM, LSTM_PIN_FP32, LSTM_GROUP16 = 1, 2, 4
GEMM_BACKGROUND, GEMM_TILE128 = 1, 2


class ConvGeom(C.Structure):
    _fields_ = [("T", C.c_int32), ("N", C.c_int32), ("F", C.c_int32), ("C", C.c_int32), ("kt", C.c_int32), ("kf", C.c_int32),
                ("st", C.c_int32), ("sf", C.c_int32), ("pt", C.c_int32), ("pf", C.c_int32)]


class CastJob(C.Structure):
    _fields_ = [("src", C.c_void_p), ("ld_src", C.c_int64), ("dst16", C.c_void_p), ("ld_dst", C.c_int64), ("rows", C.c_int64),
                ("cols", C.c_int32), ("dtype", C.c_int32), ("transpose", C.c_int32)]


class ConvPlan(C.Structure):
    _fields_ = [("t_out", C.c_int32), ("f_out", C.c_int32), ("rows", C.c_int32), ("t_padded", C.c_int32), ("k", C.c_int32),
                ("k_padded", C.c_int32)]


class LstmVariant(C.Structure):
    _fields_ = [("mi_alpha", C.c_void_p), ("mi_beta1", C.c_void_p), ("mi_beta2", C.c_void_p),
                ("ln_gain_uh", C.c_void_p), ("ln_bias_uh", C.c_void_p), ("ln_gain_wx", C.c_void_p),
                ("ln_bias_wx", C.c_void_p), ("ln_gain_c", C.c_void_p), ("ln_bias_c", C.c_void_p),
                ("ln_eps", C.c_float), ("zoneout_h", C.c_float), ("zoneout_c", C.c_float),
                ("zmask_h", C.c_void_p), ("zmask_c", C.c_void_p)]


class LstmVariantGrads(C.Structure):
    _fields_ = [("mi_alpha", C.c_void_p), ("mi_beta1", C.c_void_p), ("mi_beta2", C.c_void_p),
                ("ln_gain_uh", C.c_void_p), ("ln_bias_uh", C.c_void_p), ("ln_gain_wx", C.c_void_p),
                ("ln_bias_wx", C.c_void_p), ("ln_gain_c", C.c_void_p), ("ln_bias_c", C.c_void_p)]


_P, _I32, _I64, _F, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/asr_b200.h declares
SIGNATURES = {
    "asr_last_error": (C.c_char_p, []),
    "asr_version": (_I32, []),
    "asr_launch_count": (_I64, []),
    "asr_mfcc_plan_create": (_I32, [C.POINTER(MfccConfig), C.POINTER(_P)]),
    "asr_mfcc_plan_destroy": (None, [_P]),
    "asr_mfcc_num_feats": (_I32, [_P]),
    "asr_mfcc_num_frames": (_I32, [_P, _I64]),
    "asr_mfcc_workspace_bytes": (_SZ, [_P, _I32]),
    "asr_mfcc_workspace_bytes_ex": (_SZ, [_P, _I32, _I32]),
    "asr_mfcc_forward": (_I32, [_P, _P, _P, _I32, _I32, _P, _P, _I32, _P, _P]),
    "asr_mfcc_forward_host": (_I32, [_P, _P, _I64, _P]),
    "asr_gemm_tn": (_I32, [_I32, _I32, _I32, _I32, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _F, _I32, _P]),
    "asr_gemm_tn_ex": (_I32, [_I32, _I32, _I32, _I32, _I32, _P, _I64, _P, _I64, _P, _I64, _P, _F, _I32, _I32, _P]),
    "asr_lstm_flags_bytes": (_SZ, []),
    "asr_lstm_fuses_masks": (_I32, [_I32, _I32, _I32, _I32]),
    "asr_lstm_persistent_supported": (_I32, [_I32, _I32, _I32, _I32, _I32]),
    "asr_lstm_fuses_variants": (_I32, [_I32, _I32, _I32, _I32]),
    "asr_lstm_fp16_storage": (_I32, [_I32, _I32, _I32, _I32]),
    "asr_lstm_forward": (_I32, [C.POINTER(LstmFwdArgs), _P]),
    "asr_lstm_backward": (_I32, [C.POINTER(LstmBwdArgs), _P]),
    "asr_lstm_cell_forward": (_I32, [C.POINTER(LstmFwdArgs), C.POINTER(LstmVariant), _P, _P]),
    "asr_lstm_cell_backward": (_I32, [C.POINTER(LstmBwdArgs), C.POINTER(LstmVariant), _P, _P, _P,
                                      C.POINTER(LstmVariantGrads), _P]),
    "asr_ctc_workspace_bytes": (_SZ, [_I32, _I32, _I32]),
    "asr_ctc_loss_grad": (_I32, [_P, _I32, _I32, _I32, _P, _P, _P, _I32, _I32, _F, _P, _P, _P, _P]),
    "asr_ctc_greedy": (_I32, [_P, _I32, _I32, _I32, _P, _I32, _I32, _P, _P, _P]),
    "asr_edit_distance": (_I32, [_P, _I32, _I32, _P, _P, _P, _I32, _I32, _P, _P]),
    "asr_ctc_beam_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32]),
    "asr_ctc_beam": (_I32, [_P, _I32, _I32, _I32, _P, _I32, _I32, _I32, _P, _P, _P, _P]),
    "asr_grad_sqnorm": (_I32, [_P, _P, _P, _I64, _F, _F, _P, _P]),
    "asr_adam_step": (_I32, [_P, _P, _P, _P, _P, _I64, _F, _F, _P, _F, _F, _F, _F, _F, _I32, _P]),
    "asr_sgd_step": (_I32, [_P, _P, _P, _P, _I64, _F, _F, _P, _F, _F, _F, _P]),
    "asr_cast_rows": (_I32, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P]),
    "asr_cast_transpose": (_I32, [_P, _I64, _P, _I64, _I64, _I32, _I32, _P]),
    "asr_cast_batch": (_I32, [C.POINTER(CastJob), _I32, _P]),
    "asr_colsum": (_I32, [_P, _I64, _I64, _I32, _P, _P]),
    "asr_mask_cast": (_I32, [_P, _I32, _I64, _P, _I32, _P, _I32, _I64, _I64, _I32, _I32, _P]),
    "asr_dropout_mask": (_I32, [_P, _I64, _F, C.c_uint64, C.c_uint64, _P]),
    "asr_bernoulli_mask": (_I32, [_P, _I64, _F, _F, C.c_uint64, C.c_uint64, _P]),
    "asr_add_gaussian_noise": (_I32, [_P, _I64, _I32, _I64, _F, C.c_uint64, C.c_uint64, _P]),
    "asr_add_mask": (_I32, [_P, _P, _P, _I64, _P, _I64, _I32, _P]),
    "asr_mask_combine": (_I32, [_P, _P, _P, _P, _I32, _P, _I64, _I32, _P]),
    "asr_conv_plan_for": (_I32, [C.POINTER(ConvGeom), C.POINTER(ConvPlan)]),
    "asr_conv_pack": (_I32, [_P, C.POINTER(ConvGeom), _P, _P]),
    "asr_conv_toeplitz": (_I32, [_P, _P, C.POINTER(ConvGeom), _I32, _P, _I64, _P, _P, _P]),
    "asr_conv_act": (_I32, [_P, C.POINTER(ConvGeom), _I32, _F, _P, _I32, _I32, _P, _P]),
    "asr_conv_act_backward": (_I32, [_P, _I64, _I64, _I64, _P, _I32, _I32, C.POINTER(ConvGeom), _I32, _F, _P, _P, _P, _P]),
    "asr_conv_unfold_t": (_I32, [_P, C.POINTER(ConvGeom), _P, _I64, _P]),
    "asr_conv_toeplitz_grad": (_I32, [_P, _I64, _P, C.POINTER(ConvGeom), _I32, _P, _P, _P]),
}


class _Lib:
    """Lazy handle; attribute access returns a checked callable."""

    def __init__(self):
        self._dll = None

    def load(self):
        if self._dll is None:
            if not os.path.exists(LIB_PATH):
                raise AsrError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                               "there is no CPU / PyTorch fallback for this path")
            dll = C.CDLL(LIB_PATH)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(dll, name)          # AttributeError if the .so lacks a declared symbol
                fn.restype, fn.argtypes = res, args
            self._dll = dll
        return self._dll

    def raw(self, name):
        return getattr(self.load(), name)

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        fn = self.raw(name)
        res = SIGNATURES[name][0]
        if res is not _I32 or name in ("asr_version", "asr_mfcc_num_feats", "asr_mfcc_num_frames", "asr_lstm_fuses_masks",
                                     "asr_lstm_persistent_supported", "asr_lstm_fuses_variants", "asr_lstm_fp16_storage"):
            return fn

        def checked(*a):
            rc = fn(*a)
            if rc != 0:
                msg = self.raw("asr_last_error")()
                raise AsrError(f"{name} failed ({rc}): {msg.decode() if msg else ''}")
            return rc
        return checked


lib = _Lib()


def ptr(t):
    """device/host pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
