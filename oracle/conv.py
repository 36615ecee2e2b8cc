"""Oracle: the convolutional front end of BASELINE configs[3] ("DeepSpeech2-style 2 x Conv + 5 x BiLSTM-800"), numpy fp64.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  PARITY UNPINNED BY THE REFERENCE: the reference has no convolutional
model (README.md:118 lists DeepSpeech 2 under future work), so there is no reference code, test or vector to pin this
against; the semantics are the ones include/asr_b200.h states (cross-correlation, zero padding, bias, clipped ReLU
min(max(z, 0), clip)) with Deep Speech 2's published geometry (Amodei et al. 2015, table 2 / section 3.5: 2-D
convolutions over time and frequency, 32 channels, kernels 41 x 11 and 21 x 11 (frequency x time), strides (2, 2) and
(2, 1), hard-tanh clipped at 20; batch normalisation is left out).  Independent pins (tests/test_oracle_conv.py):
torch.nn.functional.conv2d forward and autograd.

Layout: activations [N, T, F, C]; kernels [C_out, kt, kf, C_in]; the last layer's output is flattened to
[N, T', F' * C] with the channel index fastest — the order the CUDA path emits without a copy.
"""
from __future__ import annotations

import numpy as np

DS2_FRONT = ((32, 11, 41, 2, 2), (32, 11, 21, 1, 2))      # (C_out, kt, kf, stride_t, stride_f) per layer
DS2_CLIP = 20.0


def out_len(n, k, s):
    p = (k - 1) // 2
    return (n + 2 * p - k) // s + 1


def front_out_lengths(lens, layers=DS2_FRONT):
    lens = np.asarray(lens)
    for (_, kt, _, st, _) in layers:
        lens = (lens + 2 * ((kt - 1) // 2) - kt) // st + 1
    return lens


def init_front(rng, num_features, layers=DS2_FRONT):
    p, C = {}, 1
    for i, (co, kt, kf, _, _) in enumerate(layers):
        fan_in, fan_out = kt * kf * C, kt * kf * co
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        p[f"conv{i}.W"] = rng.uniform(-lim, lim, size=(co, kt * kf * C)).astype(np.float32)      # [C_out, (kt, kf, c)]
        p[f"conv{i}.b"] = np.zeros(co, np.float32)
        C = co
    return p


def _patches(x, kt, kf, st, sf):
    """x [N, T, F, C] -> patch matrix [N, T', F', kt, kf, C] (zero padded, cross-correlation)."""
    N, T, F, C = x.shape
    pt, pf = (kt - 1) // 2, (kf - 1) // 2
    To, Fo = (T + 2 * pt - kt) // st + 1, (F + 2 * pf - kf) // sf + 1
    xp = np.zeros((N, T + 2 * pt, F + 2 * pf, C), x.dtype)
    xp[:, pt:pt + T, pf:pf + F] = x
    out = np.zeros((N, To, Fo, kt, kf, C), x.dtype)
    for a in range(kt):
        for b in range(kf):
            out[:, :, :, a, b] = xp[:, a:a + st * To:st, b:b + sf * Fo:sf]
    return out


def front_forward(params, x, layers=DS2_FRONT, clip=DS2_CLIP, dtype=np.float64):
    """x [N, T, F] -> ([N, T', F' * C_last], cache)."""
    a = np.asarray(x, dtype)[..., None]
    cache = []
    for i, (co, kt, kf, st, sf) in enumerate(layers):
        P = _patches(a, kt, kf, st, sf)
        N, To, Fo = P.shape[:3]
        Pm = P.reshape(N * To * Fo, -1)
        z = Pm @ params[f"conv{i}.W"].astype(dtype).T + params[f"conv{i}.b"].astype(dtype)
        y = np.clip(z, 0.0, clip).reshape(N, To, Fo, co)
        cache.append((a.shape, Pm, y, (kt, kf, st, sf)))
        a = y
    N, To, Fo, C = a.shape
    return a.reshape(N, To, Fo * C), cache


def front_backward(params, dout, cache, layers=DS2_FRONT, clip=DS2_CLIP):
    """dout [N, T', F' * C] -> (grads of conv{i}.W / conv{i}.b, d/dx [N, T, F])."""
    grads = {}
    g = None
    for i in range(len(layers) - 1, -1, -1):
        in_shape, Pm, y, (kt, kf, st, sf) = cache[i]
        co = y.shape[-1]
        gy = (dout.reshape(y.shape) if g is None else g) * ((y > 0.0) & (y < clip))
        gm = gy.reshape(-1, co)
        grads[f"conv{i}.W"] = gm.T @ Pm
        grads[f"conv{i}.b"] = gm.sum(axis=0)
        dP = (gm @ params[f"conv{i}.W"].astype(gm.dtype)).reshape(y.shape[:3] + (kt, kf, in_shape[3]))
        N, T, F, C = in_shape
        pt, pf = (kt - 1) // 2, (kf - 1) // 2
        To, Fo = y.shape[1], y.shape[2]
        dxp = np.zeros((N, T + 2 * pt, F + 2 * pf, C), gm.dtype)
        for a in range(kt):
            for b in range(kf):
                dxp[:, a:a + st * To:st, b:b + sf * Fo:sf] += dP[:, :, :, a, b]
        g = dxp[:, pt:pt + T, pf:pf + F]
    return grads, g[..., 0]
