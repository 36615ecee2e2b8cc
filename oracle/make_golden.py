"""Generate tests/golden/*.npz from the REFERENCE's own code (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python -m oracle.make_golden

The reference's preprocessing/audio_utils.py imports under py3 and is loaded
verbatim by path from /root/reference (read-only).  preprocessing/audio.py does
not import (librosa, py2 `unicode`/`xrange`, scipy.signal.hamming), so its class
bodies are replayed here line by line *on top of the reference's own sigproc
functions* (preemphasis, framesig, powspec, delta) plus scipy.fftpack.dct and
scipy.signal.windows.hamming — exactly the calls audio.py:235-253, 350-365,
428-440, 70-75 make.  The result pins oracle/mfcc.py, and is committed so the
GPU box (no /root/reference) can check against it.
"""
from __future__ import annotations

import importlib.util
import os

import numpy as np
from scipy.fftpack import dct
from scipy.signal.windows import hamming

from . import mfcc as om

REF = "/root/reference/preprocessing/audio_utils.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "tests", "golden")


def load_ref_sigproc():
    spec = importlib.util.spec_from_file_location("ref_audio_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ref_fbank(sp, sig, fs=16e3, nfft=512, fb=None):
    sig = sp.preemphasis(sig, 0.97)                              # audio.py:235
    frames = sp.framesig(sig, 0.025 * fs, 0.01 * fs, hamming)    # audio.py:237-240
    pspec = sp.powspec(frames, nfft)                             # audio.py:242
    energy = np.sum(pspec, 1)                                    # audio.py:244
    energy = np.where(energy == 0, np.finfo(float).eps, energy)  # audio.py:246
    feat = np.dot(pspec, fb.T)                                   # audio.py:249
    feat = np.where(feat == 0, np.finfo(float).eps, feat)        # audio.py:251
    return feat, energy, pspec


def ref_mfcc(sp, sig, fb, num_cep=13, d=True, dd=True, L=22, eps=1e-8):
    feat, energy, _ = ref_fbank(sp, sig, fb=fb)
    feat = np.log(feat)                                          # audio.py:352
    feat = dct(feat, type=2, axis=1, norm="ortho")[:, :num_cep]  # audio.py:353
    n = np.arange(num_cep)                                       # audio.py:383-385
    feat = (1 + (L / 2.0) * np.sin(np.pi * n / L)) * feat
    feat[:, 0] = np.log(energy + eps)                            # audio.py:358
    if d:                                                        # audio.py:360-365
        dl = sp.delta(feat, 2)
        feat = np.hstack([feat, dl])
        if dd:
            feat = np.hstack([feat, sp.delta(dl, 2)])
    return feat


def ref_logfbank(sp, sig, fb):
    feat, _, _ = ref_fbank(sp, sig, fb=fb)
    return np.log(feat)                                          # audio.py:430


def ref_cmvn(feats, eps=1e-8):
    feats = np.array(feats, dtype=np.float64)                    # audio.py:70-75
    feats -= np.mean(feats, axis=0, keepdims=True)
    feats /= (np.std(feats, axis=0, keepdims=True) + eps)
    return feats


def ref_filterbanks(num_filt=40, nfft=512, fs=16e3, low=20, high=7800):
    """audio.py:255-277 replayed with range for xrange."""
    hz2mel = lambda hz: 2595 * np.log10(1 + hz / 700.0)
    mel2hz = lambda mel: 700 * (10 ** (mel / 2595.0) - 1)
    pts = np.linspace(hz2mel(low), hz2mel(high), num_filt + 2)
    bin = np.floor((nfft + 1) * mel2hz(pts) / fs)
    fbank = np.zeros([num_filt, int(nfft / 2 + 1)])
    for j in range(0, num_filt):
        for i in range(int(bin[j]), int(bin[j + 1])):
            fbank[j, i] = (i - bin[j]) / (bin[j + 1] - bin[j])
        for i in range(int(bin[j + 1]), int(bin[j + 2])):
            fbank[j, i] = (bin[j + 2] - i) / (bin[j + 2] - bin[j + 1])
    return fbank


CLIPS = [(1234, 0.5), (1235, 1.0), (1236, 1.3712), (1237, 0.02), (1238, 2.0)]


def main():
    sp = load_ref_sigproc()
    os.makedirs(OUT, exist_ok=True)
    fb = ref_filterbanks()
    assert np.array_equal(fb, om.filterbanks()), "filterbank restatement differs"
    out = {"filterbank": fb, "clip_seeds": np.array([c[0] for c in CLIPS]),
           "clip_seconds": np.array([c[1] for c in CLIPS])}
    worst = 0.0
    for seed, secs in CLIPS:
        sig = np.random.RandomState(seed).randn(int(np.floor(secs * 16000))).astype(np.float32)
        # stage-by-stage check of the restated sigproc against the reference's
        assert np.array_equal(sp.preemphasis(sig, 0.97), om.preemphasis(sig, 0.97))
        fr_ref = sp.framesig(sp.preemphasis(sig, 0.97), 400.0, 160.0, hamming)
        fr_or = om.framesig(om.preemphasis(sig, 0.97), 400.0, 160.0)
        assert fr_ref.shape == fr_or.shape and np.array_equal(fr_ref, fr_or)
        assert np.allclose(sp.powspec(fr_ref, 512), om.powspec(fr_or, 512), rtol=1e-13, atol=0)
        g = {
            "mfcc26": ref_cmvn(ref_mfcc(sp, sig, fb, d=True, dd=False)),
            "mfcc39": ref_cmvn(ref_mfcc(sp, sig, fb, d=True, dd=True)),
            "mfcc13_raw": ref_mfcc(sp, sig, fb, d=False, dd=False),
            "logfbank40": ref_cmvn(ref_logfbank(sp, sig, fb)),
        }
        o = {
            "mfcc26": om.MFCC(num_cep=13, d=True, dd=False)(sig),
            "mfcc39": om.MFCC()(sig),
            "mfcc13_raw": om.MFCC(d=False, dd=False).cepstra(sig),
            "logfbank40": om.LogFbank()(sig),
        }
        for k in g:
            assert g[k].shape == o[k].shape, (k, g[k].shape, o[k].shape)
            err = np.max(np.abs(g[k] - o[k]))
            worst = max(worst, err)
            assert err < 1e-9, (seed, k, err)
            out[f"{k}_{seed}"] = g[k]
    np.savez_compressed(os.path.join(OUT, "mfcc_reference.npz"), **out)
    print(f"wrote {OUT}/mfcc_reference.npz ; worst |oracle-reference| = {worst:.3e}")


if __name__ == "__main__":
    main()
