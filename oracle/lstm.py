"""Oracle: Keras-1 (Bi)LSTM forward / backward-through-time, numpy.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  PARITY UNPINNED: the time
loop, Bidirectional wrapper and autodiff live in un-vendored Keras 1.2.2 /
TF 1.3.0; only the cell step is reference code.

Follows /root/reference/core/layers.py:432-469 (LSTM.step, the non-LN / non-MI
/ non-zoneout branch that brsmv1 and graves2006 use by default) under
Keras-1.2.2 semantics (SURVEY.md section 8c hypotheses 1-3):
  * W [D,4H], U [H,4H], b [4H]; gate order i,f,c,o (layers.py:447-450)
  * inner_activation = hard_sigmoid = clip(0.2x+0.5, 0, 1); activation = tanh
  * h0 = c0 = 0; variational dropout masks B_W [N,D], B_U [N,H] constant over
    time, already scaled by 1/(1-p) (layers.py:438-439 use B_U[0], B_W[0])
  * Bidirectional(merge_mode='concat'): the backward copy consumes the sequence
    reversed (go_backwards) and its outputs are reversed back; NO masking, so
    on a zero-padded batch the reverse direction runs over the padding first
    (core/models.py:68-70, 261-271; datasets/dataset_generator.py:227).
All tensors are batch-major [N, T, *] like the reference.
"""
from __future__ import annotations

import numpy as np


def hard_sigmoid(x):
    return np.clip(0.2 * x + 0.5, 0.0, 1.0)


def lstm_forward(x, W, U, b, reverse=False, mask_W=None, mask_U=None,
                 dtype=np.float32, matmul_cast=None):
    """One direction.  x [N,T,D] -> h [N,T,H]; also returns the cache for BPTT.

    matmul_cast: optional callable applied to both matmul operands (used by the
    precision study to emulate fp16/bf16 tensor-core inputs).
    """
    x = np.asarray(x, dtype=dtype)
    N, T, D = x.shape
    H = U.shape[0]
    cast = matmul_cast or (lambda a: a)
    Wc, Uc = cast(W.astype(dtype)), cast(U.astype(dtype))
    xm = x if mask_W is None else x * mask_W[:, None, :].astype(dtype)
    zx = (cast(xm).reshape(N * T, D) @ Wc).reshape(N, T, 4 * H) + b.astype(dtype)
    h = np.zeros((N, H), dtype=dtype)
    c = np.zeros((N, H), dtype=dtype)
    out = np.zeros((N, T, H), dtype=dtype)
    gates = np.zeros((N, T, 4 * H), dtype=dtype)   # activated i,f,g,o
    cs = np.zeros((N, T, H), dtype=dtype)
    order = range(T - 1, -1, -1) if reverse else range(T)
    for t in order:
        hm = h if mask_U is None else h * mask_U.astype(dtype)
        z = zx[:, t] + cast(hm) @ Uc
        i = hard_sigmoid(z[:, :H])
        f = hard_sigmoid(z[:, H:2 * H])
        g = np.tanh(z[:, 2 * H:3 * H])
        o = hard_sigmoid(z[:, 3 * H:])
        c = f * c + i * g
        h = o * np.tanh(c)
        out[:, t] = h
        cs[:, t] = c
        gates[:, t, :H], gates[:, t, H:2 * H] = i, f
        gates[:, t, 2 * H:3 * H], gates[:, t, 3 * H:] = g, o
    cache = dict(x=x, xm=xm, W=W, U=U, gates=gates, cs=cs, out=out,
                 reverse=reverse, mask_W=mask_W, mask_U=mask_U)
    return out, cache


def _dhs(a):
    """d hard_sigmoid / dz expressed on the activated value (0 on the clips).

    TF's clip_by_value gradient passes on the closed interval, so z exactly on
    a kink (a == 0 or 1 reached exactly) would get 0.2; measure-zero for float
    inputs, ignored here and in the CUDA path alike.
    """
    return np.where((a > 0.0) & (a < 1.0), 0.2, 0.0).astype(a.dtype)


def lstm_backward(dout, cache):
    """BPTT for one direction. dout [N,T,H] -> dx [N,T,D], dW, dU, db."""
    x, xm, W, U = cache["x"], cache["xm"], cache["W"], cache["U"]
    gates, cs, out = cache["gates"], cache["cs"], cache["out"]
    reverse, mask_W, mask_U = cache["reverse"], cache["mask_W"], cache["mask_U"]
    N, T, D = x.shape
    H = U.shape[0]
    dt = x.dtype
    dz_all = np.zeros((N, T, 4 * H), dtype=dt)
    dh_next = np.zeros((N, H), dtype=dt)
    dc_next = np.zeros((N, H), dtype=dt)
    fwd_order = list(range(T - 1, -1, -1) if reverse else range(T))
    dU = np.zeros_like(U, dtype=dt)
    for k in range(T - 1, -1, -1):
        t = fwd_order[k]
        i, f = gates[:, t, :H], gates[:, t, H:2 * H]
        g, o = gates[:, t, 2 * H:3 * H], gates[:, t, 3 * H:]
        c = cs[:, t]
        if k > 0:
            tp = fwd_order[k - 1]
            c_prev, h_prev = cs[:, tp], out[:, tp]
        else:
            c_prev = np.zeros((N, H), dtype=dt)
            h_prev = np.zeros((N, H), dtype=dt)
        tc = np.tanh(c)
        dh = dout[:, t] + dh_next
        do = dh * tc * _dhs(o)
        dc = dc_next + dh * o * (1.0 - tc * tc)
        di = dc * g * _dhs(i)
        dg = dc * i * (1.0 - g * g)
        df = dc * c_prev * _dhs(f)
        dz = np.concatenate([di, df, dg, do], axis=1)
        dz_all[:, t] = dz
        hm_prev = h_prev if mask_U is None else h_prev * mask_U.astype(dt)
        dU += hm_prev.T @ dz
        dh_next = dz @ U.T.astype(dt)
        if mask_U is not None:
            dh_next = dh_next * mask_U.astype(dt)
        dc_next = dc * f
    dW = xm.reshape(N * T, D).T @ dz_all.reshape(N * T, 4 * H)
    db = dz_all.sum(axis=(0, 1))
    dx = (dz_all.reshape(N * T, 4 * H) @ W.T.astype(dt)).reshape(N, T, D)
    if mask_W is not None:
        dx = dx * mask_W[:, None, :].astype(dt)
    return dx, dW, dU, db, dz_all


def bilstm_forward(x, params, masks=None, dtype=np.float32, matmul_cast=None):
    """Bidirectional(LSTM) with concat merge.  params = dict(Wf,Uf,bf,Wb,Ub,bb).

    masks = optional dict(Wf,Uf,Wb,Ub) of dropout masks.  Returns [N,T,2H].
    """
    m = masks or {}
    hf, cf = lstm_forward(x, params["Wf"], params["Uf"], params["bf"], False,
                          m.get("Wf"), m.get("Uf"), dtype, matmul_cast)
    hb, cb = lstm_forward(x, params["Wb"], params["Ub"], params["bb"], True,
                          m.get("Wb"), m.get("Ub"), dtype, matmul_cast)
    return np.concatenate([hf, hb], axis=2), (cf, cb)


def bilstm_backward(dout, caches):
    cf, cb = caches
    H = cf["U"].shape[0]
    dxf, dWf, dUf, dbf, _ = lstm_backward(dout[:, :, :H], cf)
    dxb, dWb, dUb, dbb, _ = lstm_backward(dout[:, :, H:], cb)
    grads = dict(Wf=dWf, Uf=dUf, bf=dbf, Wb=dWb, Ub=dUb, bb=dbb)
    return dxf + dxb, grads


# --------------------------------------------------------------------------- #
# Keras-1 initialisers (used for synthetic weights; SURVEY 8c hypothesis 1)
# --------------------------------------------------------------------------- #
def glorot_uniform(rng, shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def orthogonal(rng, shape, scale=1.1):
    """Keras-1 'orthogonal' (scale 1.1) applied to the full [H,4H] matrix."""
    a = rng.normal(0.0, 1.0, shape)
    u, _, v = np.linalg.svd(a, full_matrices=False)
    q = u if u.shape == tuple(shape) else v
    return (scale * q.reshape(shape)).astype(np.float32)


def init_lstm(rng, D, H):
    """Keras-1 LSTM(consume_less='gpu'): one glorot_uniform W [D,4H], one
    orthogonal U [H,4H], b = 0 with forget slice 1 (forget_bias_init='one')."""
    W = glorot_uniform(rng, (D, 4 * H))
    U = orthogonal(rng, (H, 4 * H))
    b = np.zeros(4 * H, dtype=np.float32)
    b[H:2 * H] = 1.0
    return W, U, b
