"""Generate tests/golden/mfcc_reference.npz from the REFERENCE's own code (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden

preprocessing/audio.py and preprocessing/audio_utils.py are loaded verbatim by path from /root/reference
(oracle/ref_shim.py:load_reference_audio supplies what they need at import time under Python 3.12: a `librosa`
stand-in that the ndarray branch never calls, the py2 names `unicode` / `xrange`, and scipy 0.19's
`scipy.signal.hamming` = today's scipy.signal.windows.hamming).  The golden vectors are the outputs of the reference's
own classes — Feature.__call__ of MFCC / LogFbank (audio.py:41-65: _call -> _postprocessing -> _standarize) and
FBank._call — NOT a replay.  While writing them this script asserts that the restatement oracle/mfcc.py agrees
(float64: 1e-9; the +-context path, which the reference runs in float32: bit-exact), and the committed file lets the
GPU box (no /root/reference) check both the oracle and the CUDA kernel against the reference.

Reference behaviours this pinned down (kept in oracle/mfcc.py):
  * FBank()(sig) raises inside _standarize: FBank._call returns the tuple (feat, energy) (audio.py:253), so FBank is only
    usable through its subclasses; the golden holds the raw (feat, energy) pair.
  * with num_context > 0 the widened matrix is np.float32 (audio.py:89-91) and CMVN runs in float32.
"""
from __future__ import annotations

import os

import numpy as np

from . import mfcc as om
from . import ref_shim as rs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (seed, seconds): sub-frame clip, non-multiples of the hop, and the BASELINE size (10 s -> 999 frames)
CLIPS = [(1234, 0.5), (1235, 1.0), (1236, 1.3712), (1237, 0.02), (1238, 2.0), (1239, 10.0)]


def main():
    audio = rs.load_reference_audio()
    os.makedirs(OUT, exist_ok=True)
    fb = audio.FBank()._filterbanks if hasattr(audio.FBank(), "_filterbanks") else audio.FBank()._get_filterbanks()
    assert np.array_equal(fb, om.filterbanks()), "filterbank restatement differs"
    out = {"filterbank": np.asarray(fb), "clip_seeds": np.array([c[0] for c in CLIPS]),
           "clip_seconds": np.array([c[1] for c in CLIPS])}
    kinds = {
        "mfcc26": (dict(num_cep=13, d=True, dd=False), "MFCC"),
        "mfcc39": (dict(), "MFCC"),
        "mfcc13_raw": (dict(d=False, dd=False, mean_norm=False, var_norm=False), "MFCC"),
        "logfbank40": (dict(), "LogFbank"),
        "logfbank123": (dict(append_energy=True, d=True, dd=True), "LogFbank"),
        "mfcc26_ctx2_s2": (dict(num_cep=13, d=True, dd=False, num_context=2, stride=2), "MFCC"),
        "mfcc26_ctx9": (dict(num_cep=13, d=True, dd=False, num_context=9), "MFCC"),
    }
    worst = 0.0
    for seed, secs in CLIPS:
        sig = np.random.RandomState(seed).randn(int(np.floor(secs * 16000))).astype(np.float32)
        big = secs >= 10.0
        for k, (kw, cls) in kinds.items():
            if big and k not in ("mfcc26", "logfbank40"):         # keep the fixture small: the C2 / C4 features at 10 s
                continue
            if (k in ("mfcc26_ctx9", "logfbank123") and secs > 0.5) or (k == "mfcc26_ctx2_s2" and secs > 1.0):
                continue
            ref = np.asarray(getattr(audio, cls)(**kw)(sig.copy()))
            got = getattr(om, cls)(**kw)(sig.copy())
            assert ref.shape == got.shape and ref.dtype == got.dtype, (k, seed, ref.shape, got.shape, ref.dtype, got.dtype)
            err = float(np.max(np.abs(ref - got)))
            worst = max(worst, err)
            assert err < (1e-9 if ref.dtype == np.float64 else 1e-12), (seed, k, err)
            out[f"{k}_{seed}"] = ref
        if not big:
            feat, energy = audio.FBank()._call(sig.copy())          # FBank alone: (feat, energy), audio.py:223-253
            ofeat, oen = om.FBank()._fbank(sig.copy())
            assert np.allclose(feat, ofeat, rtol=1e-12, atol=0) and np.allclose(energy, oen, rtol=1e-12, atol=0)
            out[f"fbank40_{seed}"], out[f"fbank_energy_{seed}"] = np.asarray(feat), np.asarray(energy)
    try:
        audio.FBank()(np.zeros(800, np.float32) + 1.0)
        raise AssertionError("the reference's FBank()(sig) was expected to fail (tuple through _standarize)")
    except (ValueError, TypeError):
        pass
    np.savez_compressed(os.path.join(OUT, "mfcc_reference.npz"), **out)
    print(f"wrote {OUT}/mfcc_reference.npz ; worst |oracle-reference| = {worst:.3e} ; "
          f"{os.path.getsize(os.path.join(OUT, 'mfcc_reference.npz'))} bytes")


if __name__ == "__main__":
    main()
