"""K6/K7 parity: CTC loss + gradient and best-path decode (C ABI) vs the oracle.
Loss: 1e-3 relative (north_star); gradient: 1e-4 absolute on softmax-scale values;
best-path label indices: bit-exact.  Gradient bar: 5e-4 absolute at T=999 (fp32 lattice, as in TF)."""
import numpy as np
import pytest
import torch

from oracle import ctc as oc
from tests.util_gpu import dev

pytestmark = pytest.mark.gpu


def _run(logits_ntc, lens, labels, grad_scale=1.0):
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    N, T, C = logits_ntc.shape
    eng = AcousticEngine(ModelSpec(4, 8, 1, C), init_params=None)
    eng._w = {}
    lg = dev(np.ascontiguousarray(logits_ntc.transpose(1, 0, 2)))
    flat, off, mx = pack_labels(labels, "cuda")
    loss, grad = eng.ctc(lg, dev(np.asarray(lens, np.int32)), flat, off, mx, grad_scale, want_grad=False)
    return loss.cpu().numpy(), grad.cpu().numpy().transpose(1, 0, 2), eng, lg


@pytest.mark.parametrize("seed,T,N", [(0, 40, 5), (1, 150, 3), (2, 999, 4)])
def test_loss_and_grad_vs_oracle(seed, T, N):
    rng = np.random.RandomState(seed)
    C = 28
    logits = (rng.randn(N, T, C) * 3).astype(np.float32)
    lens = [T] + [int(rng.randint(T // 2, T)) for _ in range(N - 1)]
    labels = [rng.randint(0, 25, size=rng.randint(2, min(50, T // 3))) for _ in range(N)]
    labels[0][1] = labels[0][0]                                   # repeated label
    loss, grad, _, _ = _run(logits, lens, labels, grad_scale=0.5)
    rl, rg = oc.ctc_loss_grad(logits, lens, labels)
    np.testing.assert_allclose(loss, rl, rtol=1e-3)
    # fp32 log-domain lattice (like TF's) vs the fp64 oracle: rounding random-walks over T frames
    assert np.abs(grad - 0.5 * rg).max() < (1e-4 if T < 100 else 5e-4)
    for n in range(N):
        assert np.all(grad[n, lens[n]:] == 0)


def test_tensorflow_known_answers_on_the_device():
    """K6 on the known-answer pair of TensorFlow's own ctc_loss_op_test.py (tests/golden/ctc_tf_known_answer.json):
    -ln p and d loss / d logits as TF prints them (loss 1e-4 relative, gradient 1e-4 absolute: the bars of this file;
    the fp64 oracle meets the same constants to 5e-7 in tests/test_oracle_ctc.py)."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_tf_known_answer.json")))
    logits = np.log(np.stack([np.asarray(c["probs"], np.float64) for c in d["cases"]])).astype(np.float32)
    assert d["blank"] == logits.shape[2] - 1
    loss, grad, _, _ = _run(logits, [5, 5], [np.asarray(c["targets"], np.int32) for c in d["cases"]])
    np.testing.assert_allclose(loss, [c["loss"] for c in d["cases"]], rtol=1e-4)
    assert np.abs(grad - np.stack([np.asarray(c["grad_wrt_logits"]) for c in d["cases"]])).max() < 1e-4


def test_edge_cases_empty_label_long_label_infeasible():
    rng = np.random.RandomState(3)
    T, C = 12, 6
    logits = rng.randn(4, T, C).astype(np.float32)
    labels = [np.array([], np.int32), np.array([1, 1, 1, 1, 1, 1, 1]), np.array([0, 1, 2]), np.array([2])]
    lens = [12, 12, 3, 1]
    loss, grad, _, _ = _run(logits, lens, labels)
    rl, rg = oc.ctc_loss_grad(logits, lens, labels)
    assert np.isinf(rl[1]) and np.isinf(loss[1])                  # 7 repeats need 13 frames
    assert np.all(grad[1] == 0)
    for n in (0, 2, 3):
        assert abs(loss[n] - rl[n]) <= 1e-3 * abs(rl[n]) + 1e-5
        assert np.abs(grad[n] - rg[n]).max() < 1e-4


def test_long_labels_use_wide_lattice():
    rng = np.random.RandomState(4)
    T, C, N = 400, 30, 2
    logits = rng.randn(N, T, C).astype(np.float32)
    labels = [rng.randint(0, 29, size=100), rng.randint(0, 29, size=70)]
    loss, grad, _, _ = _run(logits, [T, T - 7], labels)
    rl, rg = oc.ctc_loss_grad(logits, [T, T - 7], labels)
    np.testing.assert_allclose(loss, rl, rtol=1e-3)
    assert np.abs(grad - rg).max() < 1e-4


def test_greedy_bit_exact_including_ties_and_chunk_boundaries():
    rng = np.random.RandomState(5)
    N, T, C = 6, 700, 28
    logits = np.round(rng.randn(N, T, C) * 2).astype(np.float32)   # many exact ties -> first max must win
    logits[0, :, 27] += 3                                           # mostly blank
    logits[1, 200:520, 5] += 50                                     # long repeat crossing 256-frame chunks
    lens = [700, 700, 513, 256, 255, 1]
    _, _, eng, lg = _run(logits, lens, [np.array([1])] * N)
    out, out_len = eng.greedy(lg, dev(np.asarray(lens, np.int32)))
    out, out_len = out.cpu().numpy(), out_len.cpu().numpy()
    ref = oc.greedy_decode(logits, lens)
    for n in range(N):
        assert out_len[n] == len(ref[n])
        assert out[n, :out_len[n]].tolist() == ref[n]
        assert np.all(out[n, out_len[n]:] == -1)
