"""Golden vectors for the BiLSTM / model / CTC-glue part of the path, produced by RUNNING THE REFERENCE'S OWN CODE
(core/layers.py LSTM.step, core/layers_utils.py, core/models.py topologies + ctc_model, core/ctc_utils.py) under the
Keras-1 / TF-1.3 look-alike of oracle/ref_shim.py (torch.float64, autograd for the gradients).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (needs /root/reference); writes tests/golden/lstm_reference.npz,
which tests/test_oracle_ref_pin.py (CPU) checks oracle/lstm.py + oracle/model.py + oracle/ctc.py against and
tests/test_gpu_ref_pin.py feeds to the CUDA path.  Usage:  python -m oracle.make_golden_lstm

Cases (all float64):
  seq.*    one reference LSTM layer (core/layers.py:366-479) driven over T steps through its own step(), forward and
           go_backwards, for the switch combinations: default, variational dropout, MI, LN, zoneout (train and test
           phase), everything on
  model.*  whole reference topologies (core/models.py) + ctc_model: brsmv1 in the train phase (dropout 0.2, l2 1e-4) and
           in the test phase, graves2006, eyben, brsmv1 with residual + LN + MI + zoneout + input dropout; per case the
           logits, per-utterance CTC loss, best-path decode, total loss (mean CTC + l2 terms, Keras' compile of
           train.py:140-143) and d(total)/d(every parameter) by autograd through the reference's forward code
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_shim as rs

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "lstm_reference.npz")


def npy(t):
    return t.detach().numpy().copy() if isinstance(t, torch.Tensor) else np.asarray(t)


def lstm_params(layer):
    """our names for one direction's tensors of a reference LSTM layer object."""
    p = {"W": npy(layer.W), "U": npy(layer.U), "b": npy(layer.b)}
    if layer.mi is not None:
        p.update(mi_alpha=npy(layer.mi_alpha), mi_beta1=npy(layer.mi_beta1), mi_beta2=npy(layer.mi_beta2))
    if layer.layer_norm is not None:
        for ours, theirs in (("uh", "Uh"), ("wx", "Wx"), ("c", "new_c")):
            g, b = layer.layer_norm_params[theirs]
            p["ln_gain_" + ours], p["ln_bias_" + ours] = npy(g), npy(b)
    return p


def randomise(layer, rng):
    """replace the constant initial values of the MI / LN parameters (k * ones) by distinct numbers, so that a swapped
    gain / bias or gate block cannot go unnoticed."""
    with torch.no_grad():
        names = []
        if layer.mi is not None:
            names += [layer.mi_alpha, layer.mi_beta1, layer.mi_beta2]
        if layer.layer_norm is not None:
            for g, b in layer.layer_norm_params.values():
                names += [g, b]
        for w in names + [layer.b]:
            w.add_(torch.as_tensor(rng.uniform(-0.3, 0.3, size=tuple(w.shape))))


def split_lstm_log(log, layer, N):
    """K.dropout draws of one LSTM.call in order: get_constants (4 x B_U, 4 x B_W when the level is in (0,1)), then per
    step zoneout_c, zoneout_h (core/layers.py:457-467).  Returns (mask_U, mask_W, zmask_c [T,H], zmask_h [T,H], rest);
    the masks carry the 1/(1-p) scale the way oracle/lstm.py expects them."""
    i, mu, mw = 0, None, None
    if 0 < layer.dropout_U < 1:
        mu = log[0][1] / (1.0 - layer.dropout_U)
        i += 4
    if 0 < layer.dropout_W < 1:
        mw = log[i][1] / (1.0 - layer.dropout_W)
        i += 4
    return mu, mw, i


def seq_cases(LSTM, out):
    N, T, D, H = 3, 6, 5, 4
    combos = {
        "default": {},
        "dropout": dict(dropout_W=0.3, dropout_U=0.3),
        "mi": dict(mi=[0.7, 1.1, 0.9]),
        "ln": dict(layer_norm=[1.2, 0.1]),
        "zoneout": dict(zoneout_h=0.25, zoneout_c=0.35),
        "all": dict(dropout_W=0.3, dropout_U=0.3, mi=[0.7, 1.1, 0.9], layer_norm=[1.2, 0.1], zoneout_h=0.25, zoneout_c=0.35),
    }
    rng = np.random.RandomState(7)
    for name, kw in combos.items():
        for backwards in (False, True):
            for training in (True, False):
                if not training and name not in ("zoneout", "all"):
                    continue
                layer = LSTM(H, return_sequences=True, consume_less="gpu", go_backwards=backwards, **kw)
                layer.build((None, None, D))
                randomise(layer, rng)
                x = rng.randn(N, T, D)
                rs.CTX.reset(training, seed=11)
                y = layer.call(x)                     # processing order (K.rnn); Bidirectional reverses it back
                y = npy(y)[:, ::-1] if backwards else npy(y)
                tag = f"seq.{name}.{'bwd' if backwards else 'fwd'}.{'train' if training else 'test'}"
                out[tag + ".x"], out[tag + ".y"] = x, y
                for k, v in lstm_params(layer).items():
                    out[tag + ".p." + k] = v
                log = rs.CTX.dropout_log
                if training:
                    mu, mw, i = split_lstm_log(log, layer, N)
                    if mu is not None:
                        out[tag + ".mask_U"], out[tag + ".mask_W"] = mu, mw
                    if kw.get("zoneout_h"):
                        zs = log[i:]
                        assert len(zs) == 2 * T
                        order = list(range(T - 1, -1, -1) if backwards else range(T))
                        zc, zh = np.zeros((T, H)), np.zeros((T, H))
                        for s, t in enumerate(order):        # masks indexed by the TIME step they were applied at
                            zc[t], zh[t] = zs[2 * s][1], zs[2 * s + 1][1]
                        out[tag + ".zmask_c"], out[tag + ".zmask_h"] = zc, zh
                out[tag + ".kw"] = np.array(repr(kw))


def model_params(model):
    """flat parameter dict in this repo's naming from the layer objects of a reference model; also the torch leaves."""
    p, leaves, l = {}, {}, 0
    for sym in model.layers():
        layer = getattr(sym, "layer", None)
        if isinstance(layer, rs.Bidirectional):
            for d, sub in (("f", layer.forward_layer), ("b", layer.backward_layer)):
                for k, v in lstm_params(sub).items():
                    if k in ("W", "U", "b"):
                        p[f"l{l}.{k}{d}"] = v
                        leaves[f"l{l}.{k}{d}"] = getattr(sub, k)
                    else:
                        p.setdefault(f"l{l}.{k}", [None, None])["fb".index(d)] = v
                if sub.mi is not None:
                    for k in ("mi_alpha", "mi_beta1", "mi_beta2"):
                        leaves.setdefault(f"l{l}.{k}", [None, None])["fb".index(d)] = getattr(sub, k)
                if sub.layer_norm is not None:
                    for ours, theirs in (("uh", "Uh"), ("wx", "Wx"), ("c", "new_c")):
                        g, b = sub.layer_norm_params[theirs]
                        leaves.setdefault(f"l{l}.ln_gain_{ours}", [None, None])["fb".index(d)] = g
                        leaves.setdefault(f"l{l}.ln_bias_{ours}", [None, None])["fb".index(d)] = b
            l += 1
        elif isinstance(layer, rs.TimeDistributed):
            name = "dense" if l > 0 else "proj"
            p[name + ".W"], p[name + ".b"] = npy(layer.layer.W), npy(layer.layer.b)
            leaves[name + ".W"], leaves[name + ".b"] = layer.layer.W, layer.layer.b
    for k in list(p):
        if isinstance(p[k], list):
            p[k] = np.stack(p[k])
    return p, leaves


def regularisation(model):
    total = 0.0
    for sym in model.layers():
        layer = getattr(sym, "layer", None)
        subs = []
        if isinstance(layer, rs.Bidirectional):
            subs = [layer.forward_layer, layer.backward_layer]
        elif isinstance(layer, rs.TimeDistributed):
            subs = [layer.layer]
        for s in subs:
            for reg, w in s.regularizers:
                total = total + reg(w)
    return total


def model_case(tag, model, out, F, N, T, C, training, seed, rng):
    for sym in model.layers():                       # distinct MI / LN / bias values (see randomise)
        layer = getattr(sym, "layer", None)
        if isinstance(layer, rs.Bidirectional):
            randomise(layer.forward_layer, rng)
            randomise(layer.backward_layer, rng)
    x = rng.randn(N, T, F)
    lens = np.array([T, T - 2, T - 1][:N])
    for n in range(N):
        x[n, lens[n]:] = 0.0                         # zero padding after the utterance (pad_sequences 'post')
    labels = [list(rng.randint(0, C - 1, size=rng.randint(1, 4))) for _ in range(N)]
    logits_sym = [s for s in model.layers() if isinstance(getattr(s, "layer", None), rs.TimeDistributed)][-1]
    feeds = {"inputs": x, "labels": labels, "inputs_length": lens.reshape(-1, 1)}
    (loss, dec), (logits,) = model.run(feeds, training=training, seed=seed, want=[logits_sym])
    p, leaves = model_params(model)
    total = loss.mean() + regularisation(model)      # Keras: loss_weights [1, 0], mean over the batch, + regularisers
    flat = []
    for k, v in leaves.items():
        flat += v if isinstance(v, list) else [v]
    grads = torch.autograd.grad(total, flat, allow_unused=True)
    gi = iter(grads)
    for k, v in leaves.items():
        if isinstance(v, list):
            out[f"{tag}.g.{k}"] = np.stack([npy(next(gi)) for _ in v])
        else:
            out[f"{tag}.g.{k}"] = npy(next(gi))
    for k, v in p.items():
        out[f"{tag}.p.{k}"] = v
    out[tag + ".x"], out[tag + ".lens"] = x, lens
    out[tag + ".labels"] = np.array([np.array(l + [-1] * (4 - len(l))) for l in labels])
    out[tag + ".logits"], out[tag + ".ctc"], out[tag + ".total"] = npy(logits), npy(loss), npy(total)
    out[tag + ".decoded"] = np.array([np.array(d + [-1] * (T - len(d))) for d in dec])
    # every K.dropout draw of the pass, in call order: per Bidirectional forward layer then backward layer
    log, i = list(rs.CTX.dropout_log), 0
    if rs.CTX.noise_log:
        out[tag + ".noise"] = rs.CTX.noise_log[0]
    l = 0
    for sym in model.layers():
        layer = getattr(sym, "layer", None)
        if isinstance(layer, rs.Dropout) and training and 0 < layer.p < 1:
            out[tag + ".input_mask"] = log[i][1] / (1.0 - layer.p)
            i += 1
        if isinstance(layer, rs.Bidirectional) and training:
            for d, sub in (("f", layer.forward_layer), ("b", layer.backward_layer)):
                mu, mw, used = split_lstm_log(log[i:], sub, N)
                i += used
                if mu is not None:
                    out[f"{tag}.mask.{l}.U{d}"], out[f"{tag}.mask.{l}.W{d}"] = mu, mw
                if 0 < sub.zoneout_h < 1:
                    H = sub.output_dim
                    order = list(range(T - 1, -1, -1) if d == "b" else range(T))
                    zc, zh = np.zeros((T, H)), np.zeros((T, H))
                    for s, t in enumerate(order):
                        zc[t], zh[t] = log[i + 2 * s][1], log[i + 2 * s + 1][1]
                    i += 2 * T
                    out[f"{tag}.zmask.{l}.c{d}"], out[f"{tag}.zmask.{l}.h{d}"] = zc, zh
        if isinstance(layer, rs.Bidirectional):
            l += 1
    assert i == len(log) or not training, (tag, i, len(log))


def main():
    mods = rs.install()
    import logging
    logging.disable(logging.WARNING)
    LSTM, M = mods["layers"].LSTM, mods["models"]
    out = {}
    rs._Init.rng = np.random.RandomState(2024)
    seq_cases(LSTM, out)
    rng = np.random.RandomState(5)
    F, N, T, C = 5, 3, 8, 6
    model_case("model.brsmv1_train", M.brsmv1(num_features=F, num_classes=C, num_hiddens=4, num_layers=2, dropout=0.2,
                                              weight_decay=1e-4), out, F, N, T, C, True, 21, rng)
    model_case("model.brsmv1_test", M.brsmv1(num_features=F, num_classes=C, num_hiddens=4, num_layers=3, dropout=0.2,
                                             weight_decay=1e-4), out, F, N, T, C, False, 22, rng)
    model_case("model.graves2006", M.graves2006(num_features=F, num_hiddens=6, num_classes=C, std=0.6), out, F, N, T, C,
               True, 23, rng)
    model_case("model.eyben", M.eyben(num_features=F, num_hiddens=[7, 5, 3], num_classes=C), out, F, N, T, C, False, 24, rng)
    model_case("model.brsmv1_all", M.brsmv1(num_features=F, num_classes=C, num_hiddens=4, num_layers=2, dropout=0.2,
                                            zoneout=0.15, input_dropout=True, weight_decay=1e-4, residual="sum",
                                            layer_norm=[1.0, 0.0], mi=[1.0, 1.0, 1.0]), out, F, N, T, C, True, 25, rng)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, len(out), "arrays", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
