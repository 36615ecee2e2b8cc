"""ctypes wrapper of oracle/ctc_beam.c (the C restatement of beam_decode_single; TEST INFRASTRUCTURE ONLY).
build() compiles it with gcc into oracle/_build/ (done by __graft_entry__.build(); rebuilt on demand when missing)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "ctc_beam.c")
LIB = os.path.join(_HERE, "_build", "libctc_beam_oracle.so")
_dll = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"], check=True)
    return LIB


def _load():
    global _dll
    if _dll is None:
        _dll = C.CDLL(build())
        _dll.ctc_beam_oracle_batch.restype = None
        _dll.ctc_beam_oracle_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_void_p, C.c_void_p]
    return _dll


def beam_decode(logits, seq_lens, blank=None, beam_width=100, merge_repeated=True):
    """logits [N, T, C] (batch-major, like oracle.ctc.beam_decode) -> list of label lists."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    N, T, Cc = lg.shape
    blank = Cc - 1 if blank is None else blank
    sl = np.ascontiguousarray(seq_lens, dtype=np.int32)
    out = np.empty((N, T), np.int32)
    out_len = np.empty(N, np.int32)
    _load().ctc_beam_oracle_batch(lg.ctypes.data, N, T, Cc, sl.ctypes.data, int(blank), int(beam_width), int(bool(merge_repeated)),
                                  out.ctypes.data, out_len.ctypes.data)
    return [out[n, :out_len[n]].tolist() for n in range(N)]
