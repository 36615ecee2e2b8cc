"""The oracle (oracle/lstm.py, oracle/model.py, oracle/ctc.py) against golden vectors produced by executing the
reference's own core/layers.py (LSTM.step), core/layers_utils.py, core/models.py (graves2006 / eyben / brsmv1 /
ctc_model) and core/ctc_utils.py under the Keras-1 / TF-1.3 look-alike of oracle/ref_shim.py
(tests/golden/lstm_reference.npz, made by oracle/make_golden_lstm.py in the build container).  float64 on both sides:
the bars below are rounding noise, not tolerances."""
import ast
import os

import numpy as np
import pytest

from oracle import ctc as oc
from oracle import lstm as ol
from oracle import model as om

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lstm_reference.npz"), allow_pickle=False)
SEQ = sorted({k[:-2] for k in G.files if k.startswith("seq.") and k.endswith(".x")})
MODELS = sorted({k[:-2] for k in G.files if k.startswith("model.") and k.endswith(".x")})


def sub(tag, prefix):
    n = len(tag) + 1 + len(prefix)
    return {k[n:]: G[k] for k in G.files if k.startswith(tag + "." + prefix)}


@pytest.mark.parametrize("tag", SEQ)
def test_layer_sequence_matches_reference_step(tag):
    """one direction of the reference's LSTM layer driven over T steps through its own step()"""
    kw = ast.literal_eval(str(G[tag + ".kw"]))
    p = sub(tag, "p.")
    H = p["U"].shape[0]
    training = tag.endswith(".train")
    vk = {}
    if "mi" in kw:
        vk["mi"] = (p["mi_alpha"], p["mi_beta1"], p["mi_beta2"])
    if "layer_norm" in kw:
        vk["layer_norm"] = {k: (p["ln_gain_" + k], p["ln_bias_" + k]) for k in ("uh", "wx", "c")}
    if "zoneout_h" in kw:
        vk.update(zoneout_h=kw["zoneout_h"], zoneout_c=kw["zoneout_c"],
                  zmask_h=G[tag + ".zmask_h"] if training else None, zmask_c=G[tag + ".zmask_c"] if training else None)
    v = ol.make_variant(H, **vk)
    mW = G[tag + ".mask_W"] if tag + ".mask_W" in G.files else None
    mU = G[tag + ".mask_U"] if tag + ".mask_U" in G.files else None
    y, _ = ol.lstm_cell_forward(G[tag + ".x"], p["W"], p["U"], p["b"], v, reverse=".bwd." in tag, mask_W=mW, mask_U=mU,
                                dtype=np.float64)
    np.testing.assert_allclose(y, G[tag + ".y"], rtol=0, atol=1e-12)
    if not kw or set(kw) == {"dropout_W", "dropout_U"}:          # the default-branch restatement (what C2 runs)
        y2, _ = ol.lstm_forward(G[tag + ".x"], p["W"], p["U"], p["b"], reverse=".bwd." in tag, mask_W=mW, mask_U=mU,
                                dtype=np.float64)
        np.testing.assert_allclose(y2, G[tag + ".y"], rtol=0, atol=1e-12)


def _model_inputs(tag):
    params = sub(tag, "p.")
    x, lens = G[tag + ".x"].copy(), G[tag + ".lens"]
    labels = [row[row >= 0].astype(np.int32) for row in G[tag + ".labels"]]
    L = om.num_layers_of(params)
    masks = {l: sub(tag, f"mask.{l}.") for l in range(L)}
    masks = masks if any(masks.values()) else None
    zm = {l: sub(tag, f"zmask.{l}.") for l in range(L)}
    zm = zm if any(zm.values()) else None
    return params, x, lens, labels, masks, zm


@pytest.mark.parametrize("tag", MODELS)
def test_model_matches_reference_topology(tag):
    """logits, per-utterance CTC loss, best-path decode, total loss and every parameter gradient of the reference's
    own model functions (gradients: autograd through the reference's forward code)."""
    params, x, lens, labels, masks, zm = _model_inputs(tag)
    if tag + ".noise" in G.files:                    # GaussianNoise (train phase) is additive on the input
        x = x + G[tag + ".noise"]
    wd = 1e-4 if "brsmv1" in tag else 0.0
    general = tag.endswith(("brsmv1_all", "eyben"))
    if general:
        kw = dict(masks=masks, zmasks=zm)
        if tag.endswith("brsmv1_all"):
            kw.update(zoneout=0.15, residual="sum", input_mask=G[tag + ".input_mask"])
        total, ctc, grads, logits = om.loss_and_grads_general(params, x, lens, labels, weight_decay=wd, **kw)
    else:
        total, ctc, grads, logits = om.loss_and_grads(params, x, lens, labels, weight_decay=wd, masks=masks, dtype=np.float64)
    np.testing.assert_allclose(logits, G[tag + ".logits"], rtol=0, atol=1e-11)
    np.testing.assert_allclose(ctc, G[tag + ".ctc"], rtol=1e-11)
    np.testing.assert_allclose(total, G[tag + ".total"], rtol=1e-11)
    dec = oc.greedy_decode(logits, lens)
    assert dec == [row[row >= 0].tolist() for row in G[tag + ".decoded"]]
    ref_g = sub(tag, "g.")
    assert set(ref_g) == set(grads)
    for k, g in ref_g.items():
        np.testing.assert_allclose(grads[k], g, rtol=0, atol=1e-10 * max(1.0, np.abs(g).max()), err_msg=k)
