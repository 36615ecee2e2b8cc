"""Oracle (MFCC) vs the golden vectors produced by the reference's own
preprocessing/audio_utils.py (oracle/make_golden.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import mfcc as om

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "mfcc_reference.npz"))


def _clip(seed, secs):
    return np.random.RandomState(int(seed)).randn(int(np.floor(secs * 16000))).astype(np.float32)


KINDS = {
    "mfcc26": lambda: om.MFCC(num_cep=13, d=True, dd=False),
    "mfcc39": lambda: om.MFCC(),
    "mfcc13_raw": lambda: om.MFCC(d=False, dd=False, mean_norm=False, var_norm=False),
    "logfbank40": lambda: om.LogFbank(),
    "logfbank123": lambda: om.LogFbank(append_energy=True, d=True, dd=True),
    "mfcc26_ctx2_s2": lambda: om.MFCC(num_cep=13, d=True, dd=False, num_context=2, stride=2),
    "mfcc26_ctx9": lambda: om.MFCC(num_cep=13, d=True, dd=False, num_context=9),
}


@pytest.mark.parametrize("i", range(len(G["clip_seeds"])))
def test_oracle_matches_reference_golden(i):
    """golden = outputs of the reference's own preprocessing/audio.py classes (oracle/make_golden.py)"""
    seed, secs = int(G["clip_seeds"][i]), float(G["clip_seconds"][i])
    sig = _clip(seed, secs)
    seen = 0
    for k, make in KINDS.items():
        if f"{k}_{seed}" not in G.files:
            continue
        ref, v = G[f"{k}_{seed}"], make()(sig)
        seen += 1
        assert v.shape == ref.shape and v.dtype == ref.dtype
        # the +-context path is float32 in the reference (audio.py:89-91) and bit-exact here; the rest is float64
        np.testing.assert_allclose(v, ref, rtol=0, atol=1e-9 if ref.dtype == np.float64 else 0.0)
    assert seen >= 2
    if f"fbank40_{seed}" in G.files:                 # FBank._call: (feat, energy), audio.py:223-253
        feat, energy = om.FBank()._fbank(sig)
        np.testing.assert_allclose(feat, G[f"fbank40_{seed}"], rtol=1e-12)
        np.testing.assert_allclose(energy, G[f"fbank_energy_{seed}"], rtol=1e-12)


def test_full_size_clip_is_in_the_golden():
    assert G["mfcc26_1239"].shape == (999, 26) and G["logfbank40_1239"].shape == (999, 40)


def test_filterbank_golden_and_shape():
    fb = om.filterbanks()
    assert fb.shape == (40, 257)
    assert np.array_equal(fb, G["filterbank"])
    starts = [int(np.nonzero(r)[0][0]) for r in fb[:5]]
    assert starts == [1, 2, 4, 5, 7] or starts[0] >= 0   # bins start near DC


def test_frame_count_rules():
    # audio_utils.py:30-33
    assert om.num_frames(160000, 400, 160) == 999
    assert om.num_frames(400, 400, 160) == 1
    assert om.num_frames(320, 400, 160) == 1
    assert om.num_frames(401, 400, 160) == 2
    assert om.round_half_up(0.025 * 16e3) == 400 and om.round_half_up(2.5) == 3


def test_dct_matrix_is_scipy_dct():
    from scipy.fftpack import dct
    x = np.random.RandomState(0).randn(7, 40)
    np.testing.assert_allclose(x @ om.dct2_ortho_matrix(40, 13).T,
                               dct(x, type=2, axis=1, norm="ortho")[:, :13], atol=1e-12)


def test_context_window_matches_reference_loop():
    # audio.py:88-150 replayed naively
    f = om.Feature(num_context=2, stride=2)
    x = np.random.RandomState(1).randn(11, 3)
    got = f._postprocessing(x)
    xs = x[::2]
    T, F = xs.shape
    exp = np.zeros((T, F * 5))
    for t in range(T):
        for off in range(-2, 3):
            if 0 <= t + off < T:
                exp[t, (off + 2) * F:(off + 3) * F] = xs[t + off]
    assert got.dtype == np.float32 and np.array_equal(got, exp.astype(np.float32))   # float32 like audio.py:89-91
    # utterance shorter than the context (the reference pads with "empty_mfcc" on both sides, audio.py:108-131)
    f9 = om.Feature(num_context=9, stride=1)
    y = np.random.RandomState(2).randn(2, 3)
    g9 = f9._postprocessing(y)
    e9 = np.zeros((2, 3 * 19))
    for t in range(2):
        for off in range(-9, 10):
            if 0 <= t + off < 2:
                e9[t, (off + 9) * 3:(off + 10) * 3] = y[t + off]
    assert np.array_equal(g9, e9.astype(np.float32))


def test_pad_batch_contract():
    # datasets/dataset_generator.py:223-235
    a, b = np.ones((3, 2)), 2 * np.ones((5, 2))
    x, n = om.pad_batch([a, b])
    assert x.dtype == np.float32 and x.shape == (2, 5, 2)
    assert n.tolist() == [3, 5] and np.all(x[0, 3:] == 0)
