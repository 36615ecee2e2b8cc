"""Oracle: whole acoustic-model training step (numpy): stacked BiLSTM -> Dense
-> CTC loss -> BPTT -> global-norm clip -> Adam.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  PARITY UNPINNED (Keras/TF
arithmetic, see oracle/lstm.py and oracle/ctc.py).

Topology follows /root/reference/core/models.py:
  graves2006 (:55-73)  = 1 x Bidirectional(LSTM(H)) + TimeDistributed(Dense(C))
  brsmv1     (:217-281)= num_layers x Bidirectional(LSTM(H, l2, dropout_W/U))
                         + TimeDistributed(Dense(C, l2))        (no residual/LN/MI)
  ctc_model  (:31-52)  = loss = tf.nn.ctc_loss per utterance; greedy decoder.
Loss reduction / optimiser follow /root/reference/train.py:133-143 under
Keras-1.2.2 semantics: total = mean_N(ctc) + sum l2(weight_decay)*||W||^2 over
W,U of every LSTM and the Dense kernel; grads clipped by *global* norm
(clipnorm=400) and fed to Adam(lr=1e-3, b1=.9, b2=.999, eps=1e-8).
"""
from __future__ import annotations

import numpy as np

from . import ctc as octc
from . import lstm as olstm


def init_params(num_features, num_hiddens, num_layers, num_classes, seed=4321, input_dense=None):
    """Flat dict of fp32 parameters with Keras-1 initialisers.  num_hiddens may be a per-layer sequence and
    input_dense the width of a leading TimeDistributed(Dense) (eyben, core/models.py:76-103)."""
    rng = np.random.RandomState(seed)
    p = {}
    D = num_features
    hs = list(num_hiddens) if isinstance(num_hiddens, (list, tuple)) else [num_hiddens] * num_layers
    if input_dense:
        p["proj.W"] = olstm.glorot_uniform(rng, (D, input_dense))
        p["proj.b"] = np.zeros(input_dense, dtype=np.float32)
        D = input_dense
    for l, num_hiddens in enumerate(hs):
        for d in ("f", "b"):
            W, U, b = olstm.init_lstm(rng, D, num_hiddens)
            p[f"l{l}.W{d}"], p[f"l{l}.U{d}"], p[f"l{l}.b{d}"] = W, U, b
        D = 2 * num_hiddens
    p["dense.W"] = olstm.glorot_uniform(rng, (D, num_classes))
    p["dense.b"] = np.zeros(num_classes, dtype=np.float32)
    return p


def num_layers_of(params):
    return 1 + max(int(k[1:k.index(".")]) for k in params if k.startswith("l"))


def forward(params, x, masks=None, dtype=np.float32, matmul_cast=None):
    """x [N,T,F] -> logits [N,T,C] (linear; softmax lives inside the CTC op)."""
    L = num_layers_of(params)
    h, caches = np.asarray(x, dtype=dtype), []
    for l in range(L):
        lp = {k: params[f"l{l}.{k}"] for k in ("Wf", "Uf", "bf", "Wb", "Ub", "bb")}
        lm = None if masks is None else masks.get(l)
        h, c = olstm.bilstm_forward(h, lp, lm, dtype, matmul_cast)
        caches.append(c)
    N, T, D = h.shape
    cast = matmul_cast or (lambda a: a)
    logits = (cast(h).reshape(N * T, D) @ cast(params["dense.W"].astype(dtype))
              ).reshape(N, T, -1) + params["dense.b"].astype(dtype)
    return logits, (caches, h)


def loss_and_grads(params, x, x_len, labels, weight_decay=0.0, masks=None,
                   dtype=np.float32, global_batch=None):
    """Returns (total_loss, ctc_loss[N], grads dict, logits).

    grads are d(total)/dparam with total = sum_n(ctc_n)/global_batch + l2 terms
    (global_batch defaults to N; a data-parallel rank passes the global size).
    """
    logits, (caches, top) = forward(params, x, masks, dtype)
    N, T, C = logits.shape
    gb = float(global_batch or N)
    ctc, dlogits = octc.ctc_loss_grad(logits, x_len, labels, dtype=dtype)
    dlogits = (dlogits / gb).astype(dtype)
    grads = {}
    D = top.shape[2]
    grads["dense.W"] = top.reshape(N * T, D).T @ dlogits.reshape(N * T, C)
    grads["dense.b"] = dlogits.sum(axis=(0, 1))
    dh = (dlogits.reshape(N * T, C) @ params["dense.W"].T.astype(dtype)).reshape(N, T, D)
    for l in range(len(caches) - 1, -1, -1):
        dh, g = olstm.bilstm_backward(dh, caches[l])
        for k, v in g.items():
            grads[f"l{l}.{k}"] = v
    reg = 0.0
    if weight_decay:
        for k in params:
            if k.endswith((".Wf", ".Uf", ".Wb", ".Ub")) or k == "dense.W":
                reg += weight_decay * float(np.sum(np.square(params[k], dtype=np.float64)))
                grads[k] = grads[k] + 2.0 * weight_decay * params[k]
    total = float(ctc.astype(np.float64).sum()) / gb + reg
    return total, ctc, {k: np.asarray(v, dtype=dtype) for k, v in grads.items()}, logits


def loss_and_grads_conv(params, x, x_len, labels, layers=None, clip=None, weight_decay=0.0, masks=None, dtype=np.float64,
                        global_batch=None):
    """BASELINE configs[3]: convolutional front end (oracle/conv.py; NOT in the reference) in front of the BiLSTM stack.
    x [N, T, F] -> (total, ctc[N], grads incl. conv{i}.W / conv{i}.b, logits [N, T', C], x_len')."""
    from . import conv as ocv
    layers = layers or ocv.DS2_FRONT
    clip = ocv.DS2_CLIP if clip is None else clip
    y, cache = ocv.front_forward(params, x, layers, clip, dtype)
    len2 = ocv.front_out_lengths(x_len, layers)
    lstm_params = {k: v for k, v in params.items() if not k.startswith("conv")}
    logits, (caches, top) = forward(lstm_params, y, masks, dtype)
    N, T, C = logits.shape
    gb = float(global_batch or N)
    ctc, dlogits = octc.ctc_loss_grad(logits, len2, labels, dtype=dtype)
    dlogits = (dlogits / gb).astype(dtype)
    grads = {}
    D = top.shape[2]
    grads["dense.W"] = top.reshape(N * T, D).T @ dlogits.reshape(N * T, C)
    grads["dense.b"] = dlogits.sum(axis=(0, 1))
    dh = (dlogits.reshape(N * T, C) @ params["dense.W"].T.astype(dtype)).reshape(N, T, D)
    for l in range(len(caches) - 1, -1, -1):
        dh, g = olstm.bilstm_backward(dh, caches[l])
        for k, v in g.items():
            grads[f"l{l}.{k}"] = v
    cg, _ = ocv.front_backward(params, dh, cache, layers, clip)
    grads.update(cg)
    reg = 0.0
    if weight_decay:
        for k in params:
            if k.endswith((".Wf", ".Uf", ".Wb", ".Ub")) or k == "dense.W" or (k.startswith("conv") and k.endswith(".W")):
                reg += weight_decay * float(np.sum(np.square(params[k], dtype=np.float64)))
                grads[k] = grads[k] + 2.0 * weight_decay * params[k]
    total = float(ctc.astype(np.float64).sum()) / gb + reg
    return total, ctc, {k: np.asarray(v, dtype=dtype) for k, v in grads.items()}, logits, len2


def global_norm(grads):
    return float(np.sqrt(sum(np.sum(np.square(g, dtype=np.float64)) for g in grads.values())))


def clip_adam_step(params, grads, state, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8,
                   clipnorm=400.0):
    """Keras-1.2.2 optimizers.py: clip_norm (g*c/n when n >= c, global n) then
    Adam with lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t*m/(sqrt(v)+eps).
    state = dict(t=int, m={}, v={}) updated in place; returns the pre-clip norm."""
    n = global_norm(grads)
    scale = clipnorm / n if (clipnorm and n >= clipnorm) else 1.0
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    for k, g in grads.items():
        g = (g * scale).astype(np.float32)
        m = state.setdefault("m", {}).get(k, np.zeros_like(g))
        v = state.setdefault("v", {}).get(k, np.zeros_like(g))
        m = (b1 * m + (1.0 - b1) * g).astype(np.float32)
        v = (b2 * v + (1.0 - b2) * np.square(g)).astype(np.float32)
        params[k] = (params[k] - lr_t * m / (np.sqrt(v) + eps)).astype(np.float32)
        state["m"][k], state["v"][k] = m, v
    return n


# --------------------------------------------------------------------------- #
# synthetic workload (spec: datasets/dummy.py:60-84, seeded)
# --------------------------------------------------------------------------- #
def synth_clip(seed, i, seconds=10.0, fs=16000):
    """dummy.py:71-72 — Gaussian noise clip, seeded, float32."""
    return np.random.RandomState(seed + i).randn(int(np.floor(seconds * fs))).astype(np.float32)


def synth_labels(seed, n, max_label_length=50):
    """dummy.py:80-84 — length randint(2,max), chars a..y -> ids 0..24."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        L = rng.randint(2, max_label_length)
        out.append(rng.randint(0, 25, size=L).astype(np.int32))
    return out


# --------------------------------------------------------------------------- #
# brsmv1 with its switches on (core/models.py:217-281): zoneout, layer norm,
# multiplicative integration, residual merge, input dropout
# --------------------------------------------------------------------------- #
def init_variant_params(params, num_features, num_hiddens, num_layers, layer_norm=None, mi=None, residual=None,
                        seed=99):
    """Adds the extra parameters of the switches to a dict made by init_params():
    l{l}.mi_alpha/beta1/beta2 [2,4H] = k*ones (core/layers.py:391-405), l{l}.ln_gain_*/ln_bias_* [2,4H] / [2,H]
    (core/layers.py:407-422), proj.W [F,2H] glorot + proj.b (the TimeDistributed(Dense(2H)) of
    core/models.py:253-255; with it every layer sees 2H inputs, so l0.W* must be re-drawn by the caller)."""
    H = num_hiddens
    p = dict(params)
    for l in range(num_layers):
        if mi is not None:
            for name, k in zip(("mi_alpha", "mi_beta1", "mi_beta2"), mi):
                p[f"l{l}.{name}"] = np.full((2, 4 * H), k, np.float32)
        if layer_norm is not None:
            g0, b0 = layer_norm
            for name, w in (("uh", 4 * H), ("wx", 4 * H), ("c", H)):
                p[f"l{l}.ln_gain_{name}"] = np.full((2, w), g0, np.float32)
                p[f"l{l}.ln_bias_{name}"] = np.full((2, w), b0, np.float32)
    if residual is not None:
        rng = np.random.RandomState(seed)
        p["proj.W"] = olstm.glorot_uniform(rng, (num_features, 2 * H))
        p["proj.b"] = np.zeros(2 * H, np.float32)
        for d in ("f", "b"):
            W, _, _ = olstm.init_lstm(rng, 2 * H, H)
            p[f"l0.W{d}"] = W
    return p


def _layer_variant(params, l, i, H, zoneout, zmasks):
    kw = {}
    if f"l{l}.mi_alpha" in params:
        kw["mi"] = tuple(params[f"l{l}.{n}"][i] for n in ("mi_alpha", "mi_beta1", "mi_beta2"))
    if f"l{l}.ln_gain_uh" in params:
        kw["layer_norm"] = {k: (params[f"l{l}.ln_gain_{k}"][i], params[f"l{l}.ln_bias_{k}"][i]) for k in ("uh", "wx", "c")}
    if zoneout:
        zm = (zmasks or {}).get(l, {})
        kw.update(zoneout_h=zoneout, zoneout_c=zoneout, zmask_h=zm.get("h" + "fb"[i]), zmask_c=zm.get("c" + "fb"[i]))
    return olstm.make_variant(H, **kw)


def forward_general(params, x, masks=None, zoneout=0.0, zmasks=None, residual=None, input_mask=None, dtype=np.float64):
    """x [N,T,F] -> logits.  masks: {layer: {Wf,Wb,Uf,Ub}} dropout masks; zmasks: {layer: {hf,hb,cf,cb: [T,H]}} zoneout
    keep masks (None = inference blend); residual: None | 'sum' (core/models.py:273-274); input_mask: [N,T,D]
    element-wise Dropout mask applied after the optional projection (core/models.py:257-258)."""
    L = num_layers_of(params)
    o = np.asarray(x, dtype=dtype)
    N, T, _ = o.shape
    ctx = dict(x=o, caches=[], ins=[])
    assert residual in (None, "sum")
    if "proj.W" in params:      # TimeDistributed(Dense): the residual stack's 2H projection (core/models.py:253-255) or
        o = (o.reshape(N * T, -1) @ params["proj.W"].astype(dtype)).reshape(N, T, -1) + params["proj.b"].astype(dtype)   # eyben's input layer (:90-91)
    if input_mask is not None:
        o = o * input_mask
    for l in range(L):
        H = params[f"l{l}.Uf"].shape[0]
        lm = (masks or {}).get(l, {})
        outs, cs = [], []
        for i, d in enumerate("fb"):
            v = _layer_variant(params, l, i, H, zoneout, zmasks)
            out, c = olstm.lstm_cell_forward(o, params[f"l{l}.W{d}"], params[f"l{l}.U{d}"], params[f"l{l}.b{d}"], v,
                                             reverse=(d == "b"), mask_W=lm.get("W" + d), mask_U=lm.get("U" + d), dtype=dtype)
            outs.append(out)
            cs.append(c)
        new_o = np.concatenate(outs, axis=2)
        ctx["caches"].append(cs)
        o = new_o + o if residual is not None else new_o
    D = o.shape[2]
    logits = (o.reshape(N * T, D) @ params["dense.W"].astype(dtype)).reshape(N, T, -1) + params["dense.b"].astype(dtype)
    ctx.update(top=o, residual=residual, input_mask=input_mask)
    return logits, ctx


def loss_and_grads_general(params, x, x_len, labels, weight_decay=0.0, global_batch=None, dtype=np.float64, **kw):
    """Returns (total, ctc[N], grads, logits) like loss_and_grads(), for forward_general()."""
    logits, ctx = forward_general(params, x, dtype=dtype, **kw)
    N, T, C = logits.shape
    gb = float(global_batch or N)
    ctc, dlogits = octc.ctc_loss_grad(logits, x_len, labels, dtype=dtype)
    dlogits = (dlogits / gb).astype(dtype)
    top = ctx["top"]
    D = top.shape[2]
    grads = {"dense.W": top.reshape(N * T, D).T @ dlogits.reshape(N * T, C), "dense.b": dlogits.sum(axis=(0, 1))}
    do = (dlogits.reshape(N * T, C) @ params["dense.W"].T.astype(dtype)).reshape(N, T, D)
    for l in range(len(ctx["caches"]) - 1, -1, -1):
        H = params[f"l{l}.Uf"].shape[0]
        dx = 0.0
        for i, d in enumerate("fb"):
            dxi, gp, _ = olstm.lstm_cell_backward(do[:, :, i * H:(i + 1) * H], ctx["caches"][l][i])
            dx = dx + dxi
            grads[f"l{l}.W{d}"], grads[f"l{l}.U{d}"], grads[f"l{l}.b{d}"] = gp["W"], gp["U"], gp["b"]
            for k, g in gp.items():
                if k.startswith(("mi_", "ln_")):
                    grads.setdefault(f"l{l}.{k}", np.zeros((2,) + g.shape, dtype))[i] = g
        do = do + dx if ctx["residual"] is not None else dx
    if ctx["input_mask"] is not None:
        do = do * ctx["input_mask"]
    if "proj.W" in params:
        F = ctx["x"].shape[2]
        grads["proj.W"] = ctx["x"].reshape(N * T, F).T @ do.reshape(N * T, -1)
        grads["proj.b"] = do.sum(axis=(0, 1))
    reg = 0.0
    if weight_decay:
        for k in params:
            if k.endswith((".Wf", ".Uf", ".Wb", ".Ub")) or k in ("dense.W", "proj.W"):
                reg += weight_decay * float(np.sum(np.square(params[k], dtype=np.float64)))
                grads[k] = grads[k] + 2.0 * weight_decay * params[k]
    total = float(ctc.astype(np.float64).sum()) / gb + reg
    return total, ctc, {k: np.asarray(v, dtype=dtype) for k, v in grads.items()}, logits
