"""Audio front end — the reference's feature-plugin surface on the fused CUDA kernel.

Same names, constructor keywords, ``num_feats`` and ``str()`` as
preprocessing/audio.py:18-465 of the reference, so
``get_from_module('preprocessing.audio', 'mfcc', params=[...])`` (train.py:176-178)
resolves here unchanged.  ``obj(audio)`` takes a 1-D ndarray/list and returns a
fresh host ndarray [T, num_feats]; ``obj.batch(...)`` is the batched device entry
the training loop uses (one kernel launch for the whole batch).

File paths (audio.py:55-59: librosa.load at its default 22 050 Hz, then librosa.resample
to ``fs``) are served by ``load_audio`` below with scipy (RIFF/WAV only, polyphase resampling
straight to ``fs``): librosa is not installable here, so samples differ from the reference's
two-stage resampling at the filter-ripple level; ndarray / list inputs (the hot path) are exact.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from .._lib import AsrError, MfccConfig, cur_stream, lib, ptr

_KIND = {"mfcc": 0, "logfbank": 1, "fbank": 2}


def load_audio(path, fs):
    """WAV file -> mono float32 in [-1, 1] at ``fs`` Hz (stands in for librosa.load + librosa.resample, audio.py:57-58)."""
    from fractions import Fraction

    import scipy.io.wavfile
    import scipy.signal
    sr, x = scipy.io.wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    x = x.astype(np.float32)
    if x.ndim > 1:
        x = x.mean(axis=1)                      # librosa.load(mono=True)
    if int(sr) != int(fs):
        fr = Fraction(int(fs), int(sr)).limit_denominator(1000)
        x = scipy.signal.resample_poly(x, fr.numerator, fr.denominator).astype(np.float32)
    return x


class Feature(object):
    """Base class (audio.py:18-157)."""

    def __init__(self, fs=16e3, eps=1e-8, stride=1, num_context=0, mean_norm=True, var_norm=True):
        self.fs = fs
        self.eps = eps
        self.mean_norm = mean_norm
        self.var_norm = var_norm
        self.stride = stride
        self.num_context = num_context
        self._plan = None
        self._ws = {}
        self._ws_lock = threading.Lock()

    # -- plan -----------------------------------------------------------------
    def _config(self) -> MfccConfig:
        raise NotImplementedError("__call__ must be overrided")

    def _get_plan(self):
        if self._plan is None:
            cfg = self._config()
            h = C.c_void_p()
            lib.asr_mfcc_plan_create(C.byref(cfg), C.byref(h))
            self._plan = h
        return self._plan

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None:
                lib.asr_mfcc_plan_destroy(self._plan)
        except Exception:
            pass

    # -- reference surface ----------------------------------------------------
    def __call__(self, audio):
        if isinstance(audio, str):
            audio = load_audio(audio, self.fs)
        if type(audio) not in (np.ndarray, list) or len(audio) <= 1:
            raise TypeError("audio type is not support")
        pcm = np.ascontiguousarray(np.asarray(audio, dtype=np.float32).reshape(-1))
        plan = self._get_plan()
        T = lib.asr_mfcc_num_frames(plan, pcm.shape[0])
        out = np.empty((T, self.num_feats), dtype=np.float32)
        lib.asr_mfcc_forward_host(plan, pcm.ctypes.data_as(C.c_void_p), pcm.shape[0],
                                  out.ctypes.data_as(C.c_void_p))
        return out

    def num_frames(self, num_samples: int) -> int:
        return lib.asr_mfcc_num_frames(self._get_plan(), int(num_samples))

    def batch(self, pcm, offsets, t_max=None, time_major=True, out=None):
        """Device entry.  pcm: f32 CUDA tensor [sum samples]; offsets: i64 CUDA tensor [n+1].
        Returns (feats f32 [t_max, n, F] (or [n, t_max, F]), lengths i32 [n]) on the device."""
        import torch
        plan = self._get_plan()
        n = offsets.numel() - 1
        F = self.num_feats
        if t_max is None:
            lens = (offsets[1:] - offsets[:-1]).cpu().numpy()
            t_max = max(self.num_frames(int(s)) for s in lens)
        shape = (t_max, n, F) if time_major else (n, t_max, F)
        if out is None:
            out = torch.empty(shape, dtype=torch.float32, device=pcm.device)
        out_len = torch.empty(n, dtype=torch.int32, device=pcm.device)
        # one grow-only zeroed workspace per (device, stream): the kernel leaves the bytes it used zeroed, so a larger
        # buffer serves every smaller batch (real corpora have a different t_max for almost every batch)
        need = lib.asr_mfcc_workspace_bytes_ex(plan, n, int(t_max)) // 8 + 1
        key = (pcm.device, torch.cuda.current_stream(pcm.device).cuda_stream)
        with self._ws_lock:
            ws = self._ws.get(key)
            if ws is None or ws.numel() < need:
                ws = torch.zeros(need, dtype=torch.float64, device=pcm.device)
                self._ws[key] = ws
        lib.asr_mfcc_forward(plan, ptr(pcm), ptr(offsets), n, t_max, ptr(out), ptr(out_len), int(time_major),
                             ptr(ws), cur_stream())
        return out, out_len

    def __str__(self):
        raise NotImplementedError("__str__ must be overrided")

    @property
    def num_feats(self):
        # width of what __call__ returns: the base features widened by +-num_context frames (audio.py:146 — the
        # reference only updates _num_feats inside the first _postprocessing call; the final value is returned here)
        base = self._num_feats
        return base if base is None else base * (1 + 2 * int(self.num_context))


class FBank(Feature):
    """Mel filterbank energies (audio.py:160-306)."""

    def __init__(self, win_len=0.025, win_step=0.01, num_filt=40, nfft=512, low_freq=20, high_freq=7800,
                 pre_emph=0.97, win_fun=None, **kwargs):
        super(FBank, self).__init__(**kwargs)
        if high_freq > self.fs / 2:
            raise ValueError("high_freq must be less or equal than fs/2")
        if win_fun is not None:
            raise ValueError("only the default Hamming analysis window is built into the kernel")
        self.win_len = win_len
        self.win_step = win_step
        self.num_filt = num_filt
        self.nfft = nfft
        self.low_freq = low_freq
        self.high_freq = high_freq or self.fs / 2
        self.pre_emph = pre_emph
        self._num_feats = self.num_filt

    def _base(self, kind, num_cep=13, cep_lifter=22, append_energy=0, d=0, dd=0):
        return MfccConfig(fs=float(self.fs), win_len=float(self.win_len), win_step=float(self.win_step),
                          num_filt=int(self.num_filt), nfft=int(self.nfft), low_freq=float(self.low_freq),
                          high_freq=float(self.high_freq), pre_emph=float(self.pre_emph), kind=_KIND[kind],
                          num_cep=int(num_cep), cep_lifter=int(cep_lifter), append_energy=int(bool(append_energy)),
                          d=int(bool(d)), dd=int(bool(dd)), mean_norm=int(bool(self.mean_norm)),
                          var_norm=int(bool(self.var_norm)), eps=float(self.eps), stride=int(self.stride),
                          num_context=int(self.num_context))

    def _config(self):
        return self._base("fbank")

    def __str__(self):
        return "fbank"


class MFCC(FBank):
    """MFCC (+energy, deltas) (audio.py:309-391)."""

    def __init__(self, num_cep=13, cep_lifter=22, append_energy=True, d=True, dd=True, **kwargs):
        super(MFCC, self).__init__(**kwargs)
        self.num_cep = num_cep
        self.cep_lifter = cep_lifter
        self.append_energy = append_energy
        self.d = d
        self.dd = dd
        self._num_feats = (1 + self.d + self.dd) * self.num_cep

    def _config(self):
        return self._base("mfcc", self.num_cep, self.cep_lifter, self.append_energy, self.d, self.dd)

    def __str__(self):
        return "mfcc"


class LogFbank(FBank):
    """log mel filterbank (audio.py:394-445)."""

    def __init__(self, d=False, dd=False, append_energy=False, **kwargs):
        super(LogFbank, self).__init__(**kwargs)
        self.d = d
        self.dd = dd
        self.append_energy = append_energy
        self._num_feats = (1 + self.d + self.dd) * (self.num_filt + self.append_energy)

    def _config(self):
        return self._base("logfbank", append_energy=self.append_energy, d=self.d, dd=self.dd)

    def __str__(self):
        return "logfbank"


class Raw(Feature):
    """Pass-through (audio.py:448-462): pre-computed features are returned as they are."""

    def __init__(self, **kwargs):
        super(Raw, self).__init__(**kwargs)
        self._num_feats = None

    def __call__(self, x):
        # audio.py:455-459 + :65,70-75: _call and _postprocessing are identity, CMVN still applies.
        # Host-side glue for already-extracted features (not a compute kernel of the path).
        feats = np.array(x, dtype=np.float64)
        if self.mean_norm:
            feats -= np.mean(feats, axis=0, keepdims=True)
        if self.var_norm:
            feats /= (np.std(feats, axis=0, keepdims=True) + self.eps)
        return feats

    def __str__(self):
        return "raw"


raw = Raw()
