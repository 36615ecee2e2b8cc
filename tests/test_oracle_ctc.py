"""Oracle (CTC) cross-checks: brute-force enumeration, torch CTC, finite
differences, decoder semantics.  CPU only.  (Parity unpinned by the reference:
TF is not installed; these are the independent pins SURVEY 8c lists.)"""
import numpy as np
import pytest
import torch

from oracle import ctc as oc


def test_loss_matches_bruteforce_enumeration():
    rng = np.random.RandomState(0)
    T, C, blank = 5, 4, 3
    logits = rng.randn(T, C)
    probs = oc.brute_force_label_probs(logits, blank)
    assert abs(sum(probs.values()) - 1.0) < 1e-12
    for lab in [(0, 1), (1, 1), (2,), (), (0, 1, 2), (0, 0, 0)]:
        loss, _ = oc.ctc_loss_grad_single(logits, T, list(lab), blank)
        p = probs.get(lab, 0.0)
        if p == 0.0:
            assert np.isinf(loss)
        else:
            assert abs(loss - (-np.log(p))) < 1e-10, lab


def test_tf_style_known_answer_is_rederived():
    # 5 frames x 6 classes (blank = 5), the shape of TF's ctc_loss_op_test: we do
    # not trust remembered constants, we re-derive them by enumeration.
    rng = np.random.RandomState(3)
    logits = rng.randn(5, 6)
    probs = oc.brute_force_label_probs(logits, 5)
    for lab in [(0, 1, 2, 1, 0), (0, 1, 1, 0)]:
        loss, _ = oc.ctc_loss_grad_single(logits, 5, list(lab), 5)
        ref = -np.log(probs[lab]) if probs.get(lab, 0) > 0 else np.inf
        assert np.isclose(loss, ref, rtol=1e-10) or (np.isinf(loss) and np.isinf(ref))


def test_tensorflow_known_answers_loss_and_gradient():
    """The known-answer pair of TensorFlow's own ctc_loss_op_test.py (tests/golden/ctc_tf_known_answer.json: inputs,
    -ln p and d loss / d logits 'from Alex Graves' implementation'): the oracle reproduces TF's printed constants to
    their last digit, loss and gradient, for the (merge-repeated, blank = C-1, internal softmax) convention
    tf.nn.ctc_loss is called with at core/ctc_utils.py:68."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_tf_known_answer.json")))
    for case in d["cases"]:
        p = np.asarray(case["probs"], np.float64)
        assert np.abs(p.sum(axis=1) - 1.0).max() < 2e-6                    # typed-in rows are distributions
        loss, grad = oc.ctc_loss_grad_single(np.log(p), p.shape[0], case["targets"], d["blank"])
        assert abs(loss - case["loss"]) < 5e-6                              # TF prints 6 significant digits
        assert np.abs(grad - np.asarray(case["grad_wrt_logits"])).max() < 2e-6
    # batched entry point, float32 like the kernels' inputs, both cases at once
    logits = np.log(np.stack([np.asarray(c["probs"], np.float64) for c in d["cases"]])).astype(np.float32)
    loss, grad = oc.ctc_loss_grad(logits, [5, 5], [np.asarray(c["targets"]) for c in d["cases"]])
    np.testing.assert_allclose(loss, [c["loss"] for c in d["cases"]], atol=5e-6)
    assert np.abs(grad - np.stack([np.asarray(c["grad_wrt_logits"]) for c in d["cases"]])).max() < 2e-6


def _tf_beam_case():
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_tf_beam_known_answer.json")))
    p = np.asarray(d["probs"], np.float64)
    assert np.abs(p.sum(axis=1) - 1.0).max() < 2e-6
    logits = np.zeros((1, d["max_time"], p.shape[1]), np.float32)      # frames past seq_len are zero-padded, as in TF's test
    logits[0, :d["seq_len"]] = (np.log(p) + d["offset"])[:d["seq_len"]]
    return d, p, logits


def test_tensorflow_known_answer_beam_search():
    """TensorFlow's own 'hibernating beam search' test (tests/golden/ctc_tf_beam_known_answer.json): at beam_width 2 the
    top path TF expects is [1, 0] although [0, 1, 0] is the more probable labelling — the oracle's restatement of
    CTCBeamSearchDecoder reproduces that pruning artefact, finds [0, 1, 0] at wider beams like exhaustive enumeration,
    and ignores both the +2.0 offset on the inputs and the frames past the sequence length."""
    d, p, logits = _tf_beam_case()
    assert oc.beam_decode(logits, [d["seq_len"]], beam_width=d["beam_width"], merge_repeated=False) == [d["top_path_width_2"]]
    assert oc.beam_decode(logits, [d["seq_len"]], beam_width=d["beam_width"]) == [d["top_path_width_2"]]
    probs = oc.brute_force_label_probs(np.log(p[:d["seq_len"]]), d["blank"])
    best = max(probs, key=probs.get)
    assert list(best) == d["second_path_width_2"]
    for W in (3, 16, 100):
        assert oc.beam_decode(logits, [d["seq_len"]], beam_width=W) == [list(best)]
    assert oc.beam_decode(logits - d["offset"], [d["seq_len"]], beam_width=2) == [d["top_path_width_2"]]
    assert oc.greedy_decode(logits, [d["seq_len"]]) == [[0, 1, 0]]


def test_tensorflow_known_answers_greedy_decoder_and_edit_distance():
    """tests/golden/ctc_tf_greedy_known_answer.json: TF's own greedy-decoder test (repeats separated by a blank survive
    the merge; frames past the sequence length and -inf inputs are harmless) and the worked example of the
    tf.edit_distance docstring (empty truth -> inf, empty hypothesis -> 1.0, one missing label of two -> 0.5): the
    conventions core/ctc_utils.py:42 and core/metrics.py:8 inherit."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_tf_greedy_known_answer.json")))
    with np.errstate(divide="ignore"):
        logits = np.log(np.asarray(d["probs"], np.float64)).astype(np.float32)
    assert oc.greedy_decode(logits, d["seq_lens"], blank=d["blank"]) == d["decoded"]
    assert oc.greedy_decode(logits, d["seq_lens"]) == d["decoded"]                    # blank defaults to C - 1
    e = d["edit_distance_doc_example"]
    from asr_study_b200.core import metrics
    for h, t, want in zip(e["hyp"], e["truth"], e["expected"]):
        want = float(want)
        for fn in (oc.ler, metrics.ler):
            got = fn([t], [h])
            assert (np.isinf(got) and np.isinf(want)) or got == want, (h, t, got, want)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_loss_and_grad_match_torch(seed):
    rng = np.random.RandomState(seed)
    N, T, C = 4, 37, 28
    logits = rng.randn(N, T, C).astype(np.float32) * 2
    lens = np.array([37, 30, 21, 9])
    labels = [rng.randint(0, 25, size=L) for L in (7, 12, 3, 4)]
    labels[1][3] = labels[1][4]                        # a repeat
    loss, grad = oc.ctc_loss_grad(logits, lens, labels)
    tl = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(tl, dim=2).transpose(0, 1)
    tgt = torch.tensor(np.concatenate(labels), dtype=torch.long)
    out = torch.nn.functional.ctc_loss(lp, tgt, torch.tensor(lens), torch.tensor([len(l) for l in labels]),
                                       blank=C - 1, reduction="none", zero_infinity=False)
    out.sum().backward()
    np.testing.assert_allclose(loss, out.detach().numpy(), rtol=1e-5)
    np.testing.assert_allclose(grad, tl.grad.numpy(), atol=2e-6)
    for n in range(N):                                 # zero grad past seq_len
        assert np.all(grad[n, lens[n]:] == 0)


def test_grad_finite_difference():
    rng = np.random.RandomState(5)
    T, C = 6, 5
    logits = rng.randn(T, C)
    lab = [1, 1, 3]
    _, g = oc.ctc_loss_grad_single(logits, T, lab, C - 1)
    eps = 1e-6
    for (t, k) in [(0, 0), (2, 1), (5, 4), (3, 3)]:
        lp, lm = logits.copy(), logits.copy()
        lp[t, k] += eps
        lm[t, k] -= eps
        fd = (oc.ctc_loss_grad_single(lp, T, lab, C - 1)[0] -
              oc.ctc_loss_grad_single(lm, T, lab, C - 1)[0]) / (2 * eps)
        assert abs(fd - g[t, k]) < 1e-6


def test_greedy_semantics():
    blank = 3
    def onehot(seq):
        x = np.full((len(seq), 4), -5.0)
        for t, k in enumerate(seq):
            x[t, k] = 5.0
        return x
    assert oc.greedy_decode_single(onehot([0, 0, 3, 0, 1, 1, 3, 3, 2]), 9, blank) == [0, 0, 1, 2]
    assert oc.greedy_decode_single(onehot([0, 0, 3, 0, 1, 1, 3, 3, 2]), 4, blank) == [0, 0]
    assert oc.greedy_decode_single(onehot([3, 3]), 2, blank) == []
    tie = np.zeros((1, 4))
    assert oc.greedy_decode_single(tie, 1, blank) == [0]            # first max wins


def test_beam_finds_most_probable_labelling_on_tiny_problems():
    rng = np.random.RandomState(7)
    for trial in range(20):
        T, C = 4, 3
        logits = rng.randn(T, C) * 2
        probs = oc.brute_force_label_probs(logits, C - 1)
        best = max(probs.items(), key=lambda kv: kv[1])[0]
        got = oc.beam_decode_single(logits, T, C - 1, beam_width=100, merge_repeated=False)
        assert tuple(got) == best, (trial, got, best)


def test_beam_merge_repeated_quirk_and_width1():
    rng = np.random.RandomState(11)
    logits = rng.randn(30, 6) * 3
    a = oc.beam_decode_single(logits, 30, 5, 50, merge_repeated=False)
    b = oc.beam_decode_single(logits, 30, 5, 50, merge_repeated=True)
    collapsed = [k for i, k in enumerate(a) if i == 0 or k != a[i - 1]]
    assert b == collapsed
    assert oc.beam_decode_single(logits, 0, 5, 10) == []
    # a peaky posterior: beam == greedy
    peaky = np.full((12, 6), -20.0)
    seq = [0, 0, 5, 1, 5, 5, 2, 2, 5, 0, 5, 5]
    for t, k in enumerate(seq):
        peaky[t, k] = 20.0
    assert oc.beam_decode_single(peaky, 12, 5, 8, merge_repeated=False) == \
        oc.greedy_decode_single(peaky, 12, 5) == [0, 1, 2, 0]


def test_ler():
    assert oc.edit_distance([1, 2, 3], [1, 3]) == 1
    assert oc.edit_distance([], [1, 2]) == 2
    assert abs(oc.ler([[1, 2, 3, 4]], [[1, 2, 4]]) - 0.25) < 1e-12
    assert abs(oc.ler([[1, 2], [3]], [[1, 2], [4]]) - 0.5) < 1e-12


def test_c_beam_oracle_matches_the_python_oracle():
    """oracle/ctc_beam.c (used for the C5 label-error-rate parity on hundreds of full-length clips) against the Python
    restatement it transcribes — itself pinned on TensorFlow's known answer above: identical label sequences on
    near-uniform, peaky and tie-heavy posteriors, several widths, ragged lengths, merge_repeated on and off."""
    from oracle import ctc_beam_c as occ
    rng = np.random.RandomState(11)
    for (N, T, C, W, sharp) in [(4, 30, 6, 1, 1.0), (4, 40, 28, 5, 3.0), (3, 50, 28, 25, 8.0), (2, 60, 28, 100, 0.3),
                                (3, 25, 5, 400, 2.0)]:
        lg = (rng.randn(N, T, C) * sharp).astype(np.float32)
        lg[0] = np.round(lg[0])                                   # exact ties between classes / beams
        lens = [T, max(1, T // 2), T - 3, T][:N]
        for merge in (True, False):
            ref = oc.beam_decode(lg, lens, beam_width=W, merge_repeated=merge)
            got = occ.beam_decode(lg, lens, beam_width=W, merge_repeated=merge)
            assert got == ref, (N, T, C, W, sharp, merge)
    d, p, logits = _tf_beam_case()                                # and TensorFlow's own known answer, directly
    assert occ.beam_decode(logits, [d["seq_len"]], beam_width=d["beam_width"]) == [d["top_path_width_2"]]
    for W in (3, 16, 100):
        assert occ.beam_decode(logits, [d["seq_len"]], beam_width=W) == [d["second_path_width_2"]]
