"""The CUDA path against golden vectors produced by EXECUTING THE REFERENCE'S OWN core/models.py / core/layers.py /
core/ctc_utils.py (tests/golden/lstm_reference.npz, see oracle/make_golden_lstm.py): the reference's parameters, inputs,
dropout / zoneout masks go through the engine; logits, per-utterance CTC loss, best-path labels and every parameter
gradient come back and are compared with what the reference's code computed (gradients: autograd through the
reference's forward code).  Bars: logits 1e-3 norm-wise (max|d| / max|ref|), CTC loss 1e-3 relative, labels bit-exact,
gradients 1e-2 norm-wise per tensor (16-bit GEMM operands)."""
import os

import numpy as np
import pytest
import torch

from tests.util_gpu import dev, norm_err

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "lstm_reference.npz"), allow_pickle=False)
MODELS = sorted({k[:-2] for k in G.files if k.startswith("model.") and k.endswith(".x")})


def sub(tag, prefix):
    n = len(tag) + 1 + len(prefix)
    return {k[n:]: G[k] for k in G.files if k.startswith(tag + "." + prefix)}


@pytest.mark.parametrize("tag", MODELS)
def test_engine_matches_reference_executed_model(tag):
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    params = {k: v.astype(np.float32) for k, v in sub(tag, "p.").items()}
    x, lens = G[tag + ".x"].copy(), G[tag + ".lens"].astype(np.int32)
    if tag + ".noise" in G.files:                    # GaussianNoise (train phase) is additive on the input
        x = x + G[tag + ".noise"]
    labels = [row[row >= 0].astype(np.int32) for row in G[tag + ".labels"]]
    N, T, F = x.shape
    C = params["dense.b"].shape[0]
    L = 1 + max(int(k[1:k.index(".")]) for k in params if k.startswith("l"))
    hs = tuple(params[f"l{l}.Uf"].shape[0] for l in range(L))
    training = tag + ".mask.0.Uf" in G.files or tag.endswith("graves2006")
    kw = dict(weight_decay=1e-4 if "brsmv1" in tag else 0.0)
    if tag.endswith("brsmv1_all"):
        kw.update(dropout=0.2, zoneout=0.15, layer_norm=(1.0, 0.0), mi=(1.0, 1.0, 1.0), residual="sum", input_dropout=True)
    elif tag.endswith("brsmv1_train"):
        kw.update(dropout=0.2)
    if tag.endswith("eyben"):
        kw.update(layer_hiddens=hs, input_dense=params["proj.W"].shape[1])
    eng = AcousticEngine(ModelSpec(F, hs[0], L, C, **kw), init_params=params)
    masks = {l: {k: dev(v.astype(np.float32)) for k, v in sub(tag, f"mask.{l}.").items()} for l in range(L)}
    masks = masks if any(masks.values()) else None
    zm = None
    if tag + ".zmask.0.hf" in G.files:
        zm = {l: {k: dev(np.stack([G[f"{tag}.zmask.{l}.{k}f"], G[f"{tag}.zmask.{l}.{k}b"]]).astype(np.float32))
                  for k in ("h", "c")} for l in range(L)}
    im = None
    if tag + ".input_mask" in G.files:               # [N, T, D] -> the engine's time-major rows [T * N, D]
        m = G[tag + ".input_mask"]
        im = dev(np.ascontiguousarray(m.transpose(1, 0, 2)).reshape(T * N, -1).astype(np.float32))
    feats = dev(np.ascontiguousarray(x.transpose(1, 0, 2)).astype(np.float32))
    flat, off, mx = pack_labels(labels, "cuda")
    if training:
        loss = eng.train_step(feats, dev(lens), flat, off, mx, masks=masks, zmasks=zm, input_mask=im, lr=1e-3, clipnorm=400.0)
        logits = eng.last_logits
    else:                                            # test-phase goldens: forward + loss; gradients through a
        logits = eng.forward(feats, training=False)  # training-mode pass with every stochastic switch off
        loss, _ = eng.ctc(logits, dev(lens), flat, off, mx, want_grad=False)
        loss = loss.clone()
    torch.cuda.synchronize()
    assert eng.lstm_status() == 0
    got = logits.cpu().numpy().transpose(1, 0, 2)
    assert norm_err(got, G[tag + ".logits"]) < 1e-3, norm_err(got, G[tag + ".logits"])
    np.testing.assert_allclose(loss.cpu().numpy(), G[tag + ".ctc"], rtol=1e-3)
    out, out_len = eng.greedy(logits, dev(lens))
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    ref_dec = [row[row >= 0].tolist() for row in G[tag + ".decoded"]]
    margins = np.sort(G[tag + ".logits"], axis=2)
    margins = (margins[:, :, -1] - margins[:, :, -2])
    if margins.min() > 1e-3 * np.abs(G[tag + ".logits"]).max():      # labels are only defined when no frame is a near-tie
        assert [out[n, :out_len[n]].tolist() for n in range(N)] == ref_dec
    if training:
        dg = eng.params.export("grad")
        ref_g = sub(tag, "g.")
        for k, g in ref_g.items():
            # the engine's bucket holds d(mean CTC)/dp; the reference's total adds the l2 term (folded into the clip /
            # Adam kernels here): add it on this side
            mine = dg[k] + (2e-4 * params[k] if (kw["weight_decay"] and (k.endswith(("Wf", "Wb", "Uf", "Ub")) or k in ("dense.W", "proj.W"))) else 0)
            assert norm_err(mine, g) < 1e-2, (k, norm_err(mine, g))
