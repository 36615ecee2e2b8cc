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


def init_params(num_features, num_hiddens, num_layers, num_classes, seed=4321):
    """Flat dict of fp32 parameters with Keras-1 initialisers."""
    rng = np.random.RandomState(seed)
    p = {}
    D = num_features
    for l in range(num_layers):
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
