"""CPU: the C-ABI library loads and exports every symbol include/asr_b200.h declares; host-side
argument validation fails loudly (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(asr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from asr_study_b200._lib import LIB_PATH, SIGNATURES, lib
    assert os.path.exists(LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    dll = lib.load()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(dll, n), f"{n} declared in include/asr_b200.h but not exported"
        assert n in SIGNATURES, f"{n} has no ctypes signature in _lib.py"
    assert lib.asr_version() >= 100


def test_bad_arguments_fail_loudly_without_a_gpu():
    from asr_study_b200 import AsrError
    from asr_study_b200._lib import MfccConfig, lib
    cfg = MfccConfig(fs=16000, win_len=0.025, win_step=0.01, num_filt=40, nfft=256, low_freq=20, high_freq=7800,
                     pre_emph=0.97, kind=0, num_cep=13, cep_lifter=22, append_energy=1, d=1, dd=0, mean_norm=1,
                     var_norm=1, eps=1e-8, stride=1, num_context=0)
    h = C.c_void_p()
    with pytest.raises(AsrError, match="nfft"):
        lib.asr_mfcc_plan_create(C.byref(cfg), C.byref(h))
    with pytest.raises(AsrError, match="null"):
        lib.asr_gemm_tn(0, 0, 8, 8, 8, None, 8, None, 8, None, 8, None, 1.0, 0, None)
    with pytest.raises(AsrError):
        lib.asr_ctc_greedy(None, 1, 1, 1, None, 0, 1, None, None, None)
    assert lib.asr_ctc_workspace_bytes(999, 32, 49) == 32 * (((2 * 999 * 99 + 999 + 1) & ~1) + 4 * 999 + 2) * 4
    assert lib.asr_lstm_flags_bytes() > 0


def test_missing_library_is_an_error_not_a_fallback(monkeypatch):
    import asr_study_b200._lib as L
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libasr_b200.so")
    fresh = L._Lib()
    with pytest.raises(L.AsrError, match="no CPU"):
        fresh.load()


def test_shape_queries_of_the_tensor_core_recurrences():
    """Host-only queries (no launch): which (N, H) the tensor-core recurrences take, what the engine pads to, and that
    the scratch every launch clears fits the buffer asr_lstm_flags_bytes() sizes."""
    from asr_study_b200._lib import lib
    from asr_study_b200.engine import TC_WIDTHS, tc_width
    T = 999
    for H in TC_WIDTHS:
        for N in (8, 16, 32, 40, 64, 128, 256):          # any multiple of 8: more groups than one wave -> several launches
            assert lib.asr_lstm_fuses_masks(T, N, H, 0) == 1 and lib.asr_lstm_fuses_variants(T, N, H, 0) == 1, (N, H)
            assert lib.asr_lstm_persistent_supported(T, N, H, 1, 0) == 1
        for N in (1, 5, 13, 43):                          # ragged: the engine pads these (engine._padded_batch)
            assert lib.asr_lstm_fuses_masks(T, N, H, 0) == 0, (N, H)
    for H in (100, 200, 320, 800, 960, 1024):             # no instantiation: zero-padded by the engine, or general cell
        assert lib.asr_lstm_fuses_masks(T, 32, H, 0) == 0
        assert (tc_width(H) in TC_WIDTHS) == (128 < H <= 896)
    assert lib.asr_lstm_persistent_supported(T, 2, 100, 1, 0) == 1       # graves2006 / C1: the fp32 persistent engine
    assert lib.asr_lstm_persistent_supported(T, 16, 800, 1, 0) == 0      # un-padded BiLSTM-800: general cell only
    # widest exchange ring of any launch (BPTT reduce-scatter, H = 768: 24 CTAs, three 16-sample groups per wave)
    ring = 2 * 3 * 2 * 24 * 8 * 24 * 32 * 8
    assert lib.asr_lstm_flags_bytes() >= 8192 + ring
    # error path of the label-error-rate entry
    from asr_study_b200 import AsrError
    with pytest.raises(AsrError, match="null"):
        lib.asr_edit_distance(None, 1, 1, None, None, None, 1, 1, None, None)
