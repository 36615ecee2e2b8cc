"""K8 parity: device prefix beam search (C ABI) vs the oracle's restatement of TF's CTCBeamSearchDecoder
(top path).  Label sequences must be identical; the LER bar of config 5 follows from that."""
import numpy as np
import pytest
import torch

from oracle import ctc as oc
from tests.util_gpu import dev

pytestmark = pytest.mark.gpu


def _beam(logits_ntc, lens, W, merge=True):
    from asr_study_b200.core import ctc_utils
    y = dev(logits_ntc)
    out = ctc_utils.decode([y, dev(np.asarray(lens, np.int32))], is_greedy=False, beam_width=W, merge_repeated=merge)
    out = out.cpu().numpy()
    return [[int(v) for v in r if v >= 0] for r in out]


@pytest.mark.parametrize("T,C,W,scale", [(20, 5, 4, 2.0), (40, 6, 8, 1.0), (60, 28, 16, 3.0), (50, 28, 100, 1.5),
                                         (120, 28, 100, 4.0), (30, 12, 400, 1.0)])
def test_beam_matches_oracle(T, C, W, scale):
    rng = np.random.RandomState(T * 31 + W)
    N = 4
    logits = (rng.randn(N, T, C) * scale).astype(np.float32)
    lens = [T, T - 3, max(1, T // 2), 1]
    got = _beam(logits, lens, W)
    ref = oc.beam_decode(logits, lens, beam_width=W)
    assert got == ref
    got2 = _beam(logits, lens, W, merge=False)
    ref2 = oc.beam_decode(logits, lens, beam_width=W, merge_repeated=False)
    assert got2 == ref2


def test_tensorflow_known_answer_beam_search_on_the_device():
    """K8 on TensorFlow's own 'hibernating beam search' vectors (tests/golden/ctc_tf_beam_known_answer.json): the top
    path TF expects at beam_width 2 is [1, 0]; wider beams find the more probable [0, 1, 0]."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ctc_tf_beam_known_answer.json")))
    p = np.asarray(d["probs"], np.float64)
    logits = np.zeros((1, d["max_time"], p.shape[1]), np.float32)
    logits[0, :d["seq_len"]] = (np.log(p) + d["offset"])[:d["seq_len"]]
    assert _beam(logits, [d["seq_len"]], d["beam_width"], merge=False) == [d["top_path_width_2"]]
    assert _beam(logits, [d["seq_len"]], d["beam_width"]) == [d["top_path_width_2"]]
    for W in (3, 16, 100):
        assert _beam(logits, [d["seq_len"]], W) == [d["second_path_width_2"]]


def test_beam_peaky_equals_greedy_and_empty():
    T, C = 64, 28
    seq = np.random.RandomState(0).randint(0, C, size=T)
    logits = np.full((2, T, C), -30.0, np.float32)
    for t, k in enumerate(seq):
        logits[:, t, k] = 30.0
    got = _beam(logits, [T, 0], 50, merge=False)
    assert got[0] == oc.greedy_decode_single(logits[0], T, C - 1)
    assert got[1] == []


def test_beam_width_1_and_ler_on_model_like_posteriors():
    """width 100 on smoother, trained-like posteriors (config 5 shape in miniature): LER(beam) vs truth equals
    the oracle's to the digit because the label sequences match."""
    rng = np.random.RandomState(5)
    N, T, C = 6, 150, 28
    truth = [rng.randint(0, 26, size=rng.randint(5, 20)) for _ in range(N)]
    logits = rng.randn(N, T, C).astype(np.float32)
    for n, lab in enumerate(truth):                     # plant the truth with blanks in between
        pos = np.linspace(3, T - 4, len(lab)).astype(int)
        logits[n, :, C - 1] += 2.5
        for p_, k in zip(pos, lab):
            logits[n, p_, k] += 6.0
    got = _beam(logits, [T] * N, 100)
    ref = oc.beam_decode(logits, [T] * N, beam_width=100)
    assert got == ref
    assert abs(oc.ler(truth, got) - oc.ler(truth, ref)) < 1e-12
    assert _beam(logits, [T] * N, 1) == oc.beam_decode(logits, [T] * N, beam_width=1)


def test_beam_full_length_c5_shape():
    """BASELINE configs[4] shape: 999 frames, 28 classes, width 100 — flat (random-init-like) posteriors, where near
    ties are the rule and only a bit-identical log-sum-exp keeps the device on the oracle's beam for 999 frames."""
    rng = np.random.RandomState(99)
    T, C = 999, 28
    logits = np.stack([rng.randn(T, C) * 0.7, rng.randn(T, C) * 3.0]).astype(np.float32)
    got = _beam(logits, [T, T - 11], 100)
    ref = oc.beam_decode(logits, [T, T - 11], beam_width=100)
    assert got == ref
