"""Oracle: MFCC / log-mel front end (numpy, fp64 like the reference).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.

Restates /root/reference/preprocessing/audio.py and audio_utils.py.  The
signal-processing helpers are re-derived (not copied) and validated against the
reference's own audio_utils.py by oracle/make_golden.py.
"""
from __future__ import annotations

import decimal
import math

import numpy as np


# --------------------------------------------------------------------------- #
# sigproc  (reference: preprocessing/audio_utils.py)
# --------------------------------------------------------------------------- #
def round_half_up(x) -> int:
    """audio_utils.py:11-14 — decimal ROUND_HALF_UP to an int."""
    return int(decimal.Decimal(x).quantize(decimal.Decimal("1"),
                                           rounding=decimal.ROUND_HALF_UP))


def preemphasis(sig, coeff=0.97):
    """audio_utils.py:143-150 — y[0]=x[0]; y[n]=x[n]-coeff*x[n-1].

    Keeps the input dtype (float32 pcm stays float32, as numpy does there).
    """
    sig = np.asarray(sig)
    out = np.empty_like(sig)
    out[0] = sig[0]
    out[1:] = sig[1:] - coeff * sig[:-1]
    return out


def num_frames(slen: int, frame_len: int, frame_step: int) -> int:
    """audio_utils.py:30-33."""
    if slen <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * slen - frame_len) / frame_step))


def hamming(n: int) -> np.ndarray:
    """scipy.signal.hamming(n) (symmetric) — audio.py:182 default win_fun."""
    from scipy.signal.windows import hamming as _h   # same routine the reference calls
    return _h(n)


def framesig(sig, frame_len, frame_step, winfunc=hamming):
    """audio_utils.py:17-50 — zero-pad to whole frames, gather, window (fp64)."""
    sig = np.asarray(sig)
    slen = len(sig)
    frame_len = int(round_half_up(frame_len))
    frame_step = int(round_half_up(frame_step))
    nf = num_frames(slen, frame_len, frame_step)
    padlen = (nf - 1) * frame_step + frame_len
    pad = np.zeros(padlen, dtype=np.float64)
    pad[:slen] = sig
    idx = np.arange(frame_len)[None, :] + (np.arange(nf) * frame_step)[:, None]
    return pad[idx] * winfunc(frame_len)[None, :]


def powspec(frames, nfft):
    """audio_utils.py:98-120 — (1/NFFT) * |rfft(frame, NFFT)|^2."""
    spec = np.fft.rfft(frames, nfft)
    return (1.0 / nfft) * np.square(np.abs(spec))


def delta(feat, N=2):
    """audio_utils.py:153-173 — regression deltas with edge replication.

    d[t] = sum_{n=-N..N} n * f[t+n] / (2 * sum_{i=1..N} i^2); returns ndarray.
    """
    feat = np.asarray(feat, dtype=np.float64)
    T = feat.shape[0]
    padded = np.concatenate([np.repeat(feat[:1], N, axis=0), feat,
                             np.repeat(feat[-1:], N, axis=0)], axis=0)
    denom = sum(2 * i * i for i in range(1, N + 1))
    out = np.zeros_like(feat)
    for n in range(-N, N + 1):
        out += n * padded[N + n:N + n + T]
    return out / denom


# --------------------------------------------------------------------------- #
# Feature classes  (reference: preprocessing/audio.py)
# --------------------------------------------------------------------------- #
def hz2mel(hz):
    """audio.py:279-290."""
    return 2595 * np.log10(1 + hz / 700.0)


def mel2hz(mel):
    """audio.py:292-303."""
    return 700 * (10 ** (mel / 2595.0) - 1)


def filterbanks(num_filt=40, nfft=512, fs=16e3, low_freq=20, high_freq=7800):
    """audio.py:201-203, 255-277 — unnormalised triangles on floor()ed bins."""
    mel_points = np.linspace(hz2mel(low_freq), hz2mel(high_freq), num_filt + 2)
    bins = np.floor((nfft + 1) * mel2hz(mel_points) / fs)
    fb = np.zeros([num_filt, int(nfft / 2 + 1)])
    for j in range(num_filt):
        for i in range(int(bins[j]), int(bins[j + 1])):
            fb[j, i] = (i - bins[j]) / (bins[j + 1] - bins[j])
        for i in range(int(bins[j + 1]), int(bins[j + 2])):
            fb[j, i] = (bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])
    return fb


def dct2_ortho_matrix(n_in: int, n_out: int) -> np.ndarray:
    """scipy.fftpack.dct(type=2, norm='ortho') as an [n_out, n_in] matrix
    (audio.py:353)."""
    k = np.arange(n_out)[:, None]
    n = np.arange(n_in)[None, :]
    m = np.sqrt(2.0 / n_in) * np.cos(np.pi * k * (2 * n + 1) / (2.0 * n_in))
    m[0] *= 1.0 / np.sqrt(2.0)
    return m


def lifter_coeffs(num_cep: int, L=22) -> np.ndarray:
    """audio.py:369-388."""
    if L > 0:
        n = np.arange(num_cep)
        return 1 + (L / 2.0) * np.sin(np.pi * n / L)
    return np.ones(num_cep)


class Feature:
    """audio.py:18-157 (array input only; file loading is librosa, out of scope)."""

    def __init__(self, fs=16e3, eps=1e-8, stride=1, num_context=0,
                 mean_norm=True, var_norm=True):
        self.fs, self.eps = fs, eps
        self.stride, self.num_context = stride, num_context
        self.mean_norm, self.var_norm = mean_norm, var_norm

    def __call__(self, audio):
        feats = self._call(np.asarray(audio))
        return self._standarize(self._postprocessing(feats))

    def _standarize(self, feats):
        """audio.py:70-75 — per-utterance CMVN, population std, +eps."""
        feats = np.array(feats)              # in place in the reference: float64, or float32 behind a context window
        if self.mean_norm:
            feats -= np.mean(feats, axis=0, keepdims=True)
        if self.var_norm:
            feats /= (np.std(feats, axis=0, keepdims=True) + self.eps)
        return feats

    def _postprocessing(self, feats):
        """audio.py:77-150 — stride, then +-num_context frames (zeros outside)."""
        feats = feats[::self.stride]
        c = self.num_context
        if c == 0:
            return feats
        T, F = feats.shape
        # audio.py:89-91 builds the widened matrix as np.array([], np.float32): with a context window the features are
        # rounded to float32 BEFORE the CMVN of audio.py:65, which then runs in float32 (found by running the reference's
        # own class, oracle/make_golden.py; without a context the whole chain stays float64)
        out = np.zeros((T, F * (2 * c + 1)), dtype=np.float32)
        for off in range(-c, c + 1):
            lo, hi = max(0, -off), min(T, T - off)
            if hi <= lo:                       # utterance shorter than the offset: the whole block stays "empty_mfcc"
                continue
            out[lo:hi, (off + c) * F:(off + c + 1) * F] = feats[lo + off:hi + off]
        return out


class FBank(Feature):
    """audio.py:160-306."""

    def __init__(self, win_len=0.025, win_step=0.01, num_filt=40, nfft=512,
                 low_freq=20, high_freq=7800, pre_emph=0.97, **kw):
        super().__init__(**kw)
        if high_freq > self.fs / 2:
            raise ValueError("high_freq must be less or equal than fs/2")
        self.win_len, self.win_step = win_len, win_step
        self.num_filt, self.nfft = num_filt, nfft
        self.low_freq, self.high_freq = low_freq, high_freq or self.fs / 2
        self.pre_emph = pre_emph
        self._fb = filterbanks(num_filt, nfft, self.fs, low_freq, self.high_freq)
        self.num_feats = num_filt

    def _fbank(self, sig):
        """audio.py:223-253."""
        sig = preemphasis(sig, self.pre_emph)
        frames = framesig(sig, self.win_len * self.fs, self.win_step * self.fs)
        pspec = powspec(frames, self.nfft)
        energy = np.sum(pspec, 1)
        energy = np.where(energy == 0, np.finfo(float).eps, energy)
        feat = np.dot(pspec, self._fb.T)
        feat = np.where(feat == 0, np.finfo(float).eps, feat)
        return feat, energy

    def _call(self, sig):
        return self._fbank(sig)[0]


class MFCC(FBank):
    """audio.py:309-391."""

    def __init__(self, num_cep=13, cep_lifter=22, append_energy=True,
                 d=True, dd=True, **kw):
        super().__init__(**kw)
        self.num_cep, self.cep_lifter = num_cep, cep_lifter
        self.append_energy, self.d, self.dd = append_energy, d, dd
        self.num_feats = (1 + int(d) + int(dd)) * num_cep
        self._dct = dct2_ortho_matrix(self.num_filt, num_cep)
        self._lift = lifter_coeffs(num_cep, cep_lifter)

    def cepstra(self, sig):
        """audio.py:350-358 (before deltas)."""
        feat, energy = self._fbank(sig)
        feat = np.log(feat) @ self._dct.T
        feat = feat * self._lift[None, :]
        if self.append_energy:
            feat[:, 0] = np.log(energy + self.eps)
        return feat

    def _call(self, sig):
        feat = self.cepstra(sig)
        if self.d:                       # audio.py:360-365 (dd only if d)
            d = delta(feat, 2)
            feat = np.hstack([feat, d])
            if self.dd:
                feat = np.hstack([feat, delta(d, 2)])
        return feat


class LogFbank(FBank):
    """audio.py:394-445."""

    def __init__(self, d=False, dd=False, append_energy=False, **kw):
        super().__init__(**kw)
        self.d, self.dd, self.append_energy = d, dd, append_energy
        self.num_feats = (1 + int(d) + int(dd)) * (self.num_filt + int(append_energy))

    def _call(self, sig):
        feat, energy = self._fbank(sig)
        feat = np.log(feat)
        if self.append_energy:
            feat = np.hstack([feat, np.log(energy + self.eps)[:, None]])
        if self.d:
            d = delta(feat, 2)
            feat = np.hstack([feat, d])
            if self.dd:
                feat = np.hstack([feat, delta(d, 2)])
        return feat


def pad_batch(feats_list):
    """datasets/dataset_generator.py:223-235 — float32, zero-pad 'post'.

    Returns (x [N,Tmax,F] float32, lengths [N] int32).
    """
    lens = np.asarray([f.shape[0] for f in feats_list], dtype=np.int32)
    F = feats_list[0].shape[1]
    x = np.zeros((len(feats_list), int(lens.max()), F), dtype=np.float32)
    for i, f in enumerate(feats_list):
        x[i, :f.shape[0]] = f.astype(np.float32)
    return x, lens
