"""Oracle: Keras-1 (Bi)LSTM forward / backward-through-time, numpy.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  PARITY UNPINNED: the time
loop, Bidirectional wrapper and autodiff live in un-vendored Keras 1.2.2 /
TF 1.3.0; only the cell step is reference code.  Independent pins (tests/test_oracle_lstm.py): torch.nn.LSTM on
the whole bidirectional restatement with the inner activation swapped on both sides (structure: 1e-12), finite
differences for BPTT.

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


# --------------------------------------------------------------------------- #
# LSTM.step with the brsmv1 switches: layer normalisation, multiplicative
# integration, zoneout (core/layers.py:432-469, core/layers_utils.py:16-51)
# --------------------------------------------------------------------------- #
def layer_norm(x, gain, bias, eps=1e-5):
    """core/layers_utils.py:16-19 — moments over the feature axis; note the reference names the *variance* `std`
    and normalises by sqrt(var + eps).  Returns (y, xhat, rstd)."""
    mu = x.mean(axis=1, keepdims=True)
    var = ((x - mu) ** 2).mean(axis=1, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xhat = (x - mu) * rstd
    return xhat * gain + bias, xhat, rstd


def layer_norm_backward(dy, xhat, rstd, gain):
    dxhat = dy * gain
    dx = rstd * (dxhat - dxhat.mean(axis=1, keepdims=True) - xhat * (dxhat * xhat).mean(axis=1, keepdims=True))
    return dx, (dy * xhat).sum(axis=0), dy.sum(axis=0)


def make_variant(H, mi=None, layer_norm=None, zoneout_h=0.0, zoneout_c=0.0, zmask_h=None, zmask_c=None, eps=1e-5):
    """Parameter record for one direction.  mi = (alpha, beta1, beta2) scalars or [4H] arrays
    (core/layers.py:391-405 initialises each as k*ones); layer_norm = (gain, bias) scalars or a dict
    {uh,wx,c: (gain[·], bias[·])} (core/layers.py:407-422); zmask_* = [T,H] keep masks for the train phase
    (one mask per time step shared by the batch: K.dropout(noise_shape=(output_dim,)), layers_utils.py:34-42),
    None = inference blend with (1 - level)."""
    v = dict(mi=None, ln=None, zoneout_h=float(zoneout_h), zoneout_c=float(zoneout_c), zmask_h=zmask_h,
             zmask_c=zmask_c, eps=float(eps))
    if mi is not None:
        v["mi"] = tuple(np.full(4 * H, m, np.float64) if np.isscalar(m) else np.asarray(m, np.float64) for m in mi)
    if layer_norm is not None:
        if isinstance(layer_norm, dict):
            v["ln"] = {k: (np.asarray(g, np.float64), np.asarray(bb, np.float64)) for k, (g, bb) in layer_norm.items()}
        else:
            g0, b0 = layer_norm
            v["ln"] = {"uh": (np.full(4 * H, g0, np.float64), np.full(4 * H, b0, np.float64)),
                       "wx": (np.full(4 * H, g0, np.float64), np.full(4 * H, b0, np.float64)),
                       "c": (np.full(H, g0, np.float64), np.full(H, b0, np.float64))}
    return v


def _zone_coeff(level, mask, t, H, dt):
    """keep coefficient k so that new = prev + k * (candidate - prev) (layers_utils.py:34-42)."""
    if not (0.0 < level < 1.0):
        return None
    if mask is None:
        return np.full(H, 1.0 - level, dtype=dt)
    return np.asarray(mask[t], dtype=dt)


def lstm_cell_forward(x, W, U, b, variant, reverse=False, mask_W=None, mask_U=None, dtype=np.float64):
    """One direction of LSTM.step with the variant switches.  The bias is NOT folded into Wx: LN / MI act on the
    raw products K.dot(x*B_W, W) and K.dot(h*B_U, U) (core/layers.py:438-443)."""
    x = np.asarray(x, dtype=dtype)
    N, T, D = x.shape
    H = U.shape[0]
    W, U, b = W.astype(dtype), U.astype(dtype), b.astype(dtype)
    xm = x if mask_W is None else x * mask_W[:, None, :].astype(dtype)
    wx_raw = (xm.reshape(N * T, D) @ W).reshape(N, T, 4 * H)
    v = variant
    ln, mi = v["ln"], v["mi"]
    h = np.zeros((N, H), dtype)
    c = np.zeros((N, H), dtype)
    out = np.zeros((N, T, H), dtype)
    gates = np.zeros((N, T, 4 * H), dtype)
    cs = np.zeros((N, T, H), dtype)
    uh_raw_all = np.zeros((N, T, 4 * H), dtype)
    order = list(range(T - 1, -1, -1) if reverse else range(T))
    for t in order:
        hm = h if mask_U is None else h * mask_U.astype(dtype)
        uh_raw = hm @ U
        uh_raw_all[:, t] = uh_raw
        if ln is not None:
            uh = layer_norm(uh_raw, ln["uh"][0], ln["uh"][1], v["eps"])[0]
            wx = layer_norm(wx_raw[:, t], ln["wx"][0], ln["wx"][1], v["eps"])[0]
        else:
            uh, wx = uh_raw, wx_raw[:, t]
        if mi is not None:
            z = mi[0] * wx * uh + mi[1] * uh + mi[2] * wx + b
        else:
            z = wx + uh + b
        i = hard_sigmoid(z[:, :H])
        f = hard_sigmoid(z[:, H:2 * H])
        g = np.tanh(z[:, 2 * H:3 * H])
        o = hard_sigmoid(z[:, 3 * H:])
        c_new = f * c + i * g
        kc = _zone_coeff(v["zoneout_c"], v["zmask_c"], t, H, dtype)
        c = c_new if kc is None else c + kc * (c_new - c)
        nc = layer_norm(c, ln["c"][0], ln["c"][1], v["eps"])[0] if ln is not None else c
        h_new = o * np.tanh(nc)
        kh = _zone_coeff(v["zoneout_h"], v["zmask_h"], t, H, dtype)
        h = h_new if kh is None else h + kh * (h_new - h)
        out[:, t], cs[:, t] = h, c
        gates[:, t] = np.concatenate([i, f, g, o], axis=1)
    cache = dict(x=x, xm=xm, W=W, U=U, b=b, gates=gates, cs=cs, out=out, wx_raw=wx_raw, uh_raw=uh_raw_all,
                 reverse=reverse, mask_W=mask_W, mask_U=mask_U, variant=v)
    return out, cache


def lstm_cell_backward(dout, cache):
    """BPTT of lstm_cell_forward.  Returns dx, grads dict (W, U, b + the variant parameters), (dwx_raw, duh_raw)."""
    x, xm, W, U = cache["x"], cache["xm"], cache["W"], cache["U"]
    gates, cs, out, v = cache["gates"], cache["cs"], cache["out"], cache["variant"]
    wx_raw_all, uh_raw_all = cache["wx_raw"], cache["uh_raw"]
    mask_W, mask_U = cache["mask_W"], cache["mask_U"]
    N, T, D = x.shape
    H = U.shape[0]
    dt = x.dtype
    ln, mi = v["ln"], v["mi"]
    fwd_order = list(range(T - 1, -1, -1) if cache["reverse"] else range(T))
    dwx_all = np.zeros((N, T, 4 * H), dt)
    duh_all = np.zeros((N, T, 4 * H), dt)
    dh_carry = np.zeros((N, H), dt)
    dc_carry = np.zeros((N, H), dt)
    g_par = {"b": np.zeros(4 * H, dt)}
    if mi is not None:
        g_par.update(mi_alpha=np.zeros(4 * H, dt), mi_beta1=np.zeros(4 * H, dt), mi_beta2=np.zeros(4 * H, dt))
    if ln is not None:
        for k, w in (("uh", 4 * H), ("wx", 4 * H), ("c", H)):
            g_par["ln_gain_" + k] = np.zeros(w, dt)
            g_par["ln_bias_" + k] = np.zeros(w, dt)
    dU = np.zeros_like(U)
    for k in range(T - 1, -1, -1):
        t = fwd_order[k]
        i, f = gates[:, t, :H], gates[:, t, H:2 * H]
        g, o = gates[:, t, 2 * H:3 * H], gates[:, t, 3 * H:]
        c = cs[:, t]
        if k > 0:
            c_prev, h_prev = cs[:, fwd_order[k - 1]], out[:, fwd_order[k - 1]]
        else:
            c_prev, h_prev = np.zeros((N, H), dt), np.zeros((N, H), dt)
        dh = dout[:, t] + dh_carry
        kh = _zone_coeff(v["zoneout_h"], v["zmask_h"], t, H, dt)
        dh_new, dh_prev = (dh, 0.0) if kh is None else (kh * dh, (1.0 - kh) * dh)
        if ln is not None:
            nc, xhat_c, rstd_c = layer_norm(c, ln["c"][0], ln["c"][1], v["eps"])
        else:
            nc = c
        tnc = np.tanh(nc)
        do = dh_new * tnc
        dnc = dh_new * o * (1.0 - tnc * tnc)
        if ln is not None:
            dc_ln, gg, gb = layer_norm_backward(dnc, xhat_c, rstd_c, ln["c"][0])
            g_par["ln_gain_c"] += gg
            g_par["ln_bias_c"] += gb
        else:
            dc_ln = dnc
        dc = dc_carry + dc_ln
        kc = _zone_coeff(v["zoneout_c"], v["zmask_c"], t, H, dt)
        dc_new, dc_prev = (dc, 0.0) if kc is None else (kc * dc, (1.0 - kc) * dc)
        dz = np.concatenate([dc_new * g * _dhs(i), dc_new * c_prev * _dhs(f), dc_new * i * (1.0 - g * g),
                             do * _dhs(o)], axis=1)
        dc_carry = dc_prev + dc_new * f
        if ln is not None:
            uh, xhat_u, rstd_u = layer_norm(uh_raw_all[:, t], ln["uh"][0], ln["uh"][1], v["eps"])
            wx, xhat_w, rstd_w = layer_norm(wx_raw_all[:, t], ln["wx"][0], ln["wx"][1], v["eps"])
        else:
            uh, wx = uh_raw_all[:, t], wx_raw_all[:, t]
        g_par["b"] += dz.sum(axis=0)
        if mi is not None:
            g_par["mi_alpha"] += (dz * wx * uh).sum(axis=0)
            g_par["mi_beta1"] += (dz * uh).sum(axis=0)
            g_par["mi_beta2"] += (dz * wx).sum(axis=0)
            duh, dwx = dz * (mi[0] * wx + mi[1]), dz * (mi[0] * uh + mi[2])
        else:
            duh, dwx = dz, dz
        if ln is not None:
            duh, gg, gb = layer_norm_backward(duh, xhat_u, rstd_u, ln["uh"][0])
            g_par["ln_gain_uh"] += gg
            g_par["ln_bias_uh"] += gb
            dwx, gg, gb = layer_norm_backward(dwx, xhat_w, rstd_w, ln["wx"][0])
            g_par["ln_gain_wx"] += gg
            g_par["ln_bias_wx"] += gb
        dwx_all[:, t], duh_all[:, t] = dwx, duh
        hm_prev = h_prev if mask_U is None else h_prev * mask_U.astype(dt)
        dU += hm_prev.T @ duh
        dh_rec = duh @ U.T
        if mask_U is not None:
            dh_rec = dh_rec * mask_U.astype(dt)
        dh_carry = dh_rec + dh_prev
    dW = xm.reshape(N * T, D).T @ dwx_all.reshape(N * T, 4 * H)
    dx = (dwx_all.reshape(N * T, 4 * H) @ W.T).reshape(N, T, D)
    if mask_W is not None:
        dx = dx * mask_W[:, None, :].astype(dt)
    g_par.update(W=dW, U=dU)
    return dx, g_par, (dwx_all, duh_all)
