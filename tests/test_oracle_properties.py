"""Size-independent properties of the oracle (CPU): invariances the domain offers, checked on seeded random inputs.
The GPU suite checks the kernels against the oracle; these check the oracle against the mathematics."""
import numpy as np
import pytest

from oracle import ctc as oc
from oracle import mfcc as omf


@pytest.mark.parametrize("seed", [0, 1])
def test_ctc_invariances(seed):
    rng = np.random.RandomState(seed)
    N, T, C = 3, 50, 28
    logits = (rng.randn(N, T, C) * 2).astype(np.float64)
    lens = [50, 41, 17]
    labels = [rng.randint(0, 26, size=L) for L in (9, 12, 4)]
    loss, grad = oc.ctc_loss_grad(logits, lens, labels, dtype=np.float64)
    # a per-frame additive constant on the logits changes nothing (softmax inside)
    shift = rng.randn(N, T, 1) * 5
    loss2, grad2 = oc.ctc_loss_grad(logits + shift, lens, labels, dtype=np.float64)
    np.testing.assert_allclose(loss2, loss, rtol=1e-12)
    np.testing.assert_allclose(grad2, grad, atol=1e-12)
    # softmax minus occupancy: every gradient row sums to 0; frames past the sequence length get exactly 0
    assert np.abs(grad.sum(axis=2)).max() < 1e-12
    for n in range(N):
        assert not grad[n, lens[n]:].any()
    # the gradient is that of -log p: first-order agreement along a random direction
    d = rng.randn(N, T, C)
    eps = 1e-6
    lp, _ = oc.ctc_loss_grad(logits + eps * d, lens, labels, dtype=np.float64)
    lm, _ = oc.ctc_loss_grad(logits - eps * d, lens, labels, dtype=np.float64)
    np.testing.assert_allclose((lp - lm) / (2 * eps), (grad * d).sum(axis=(1, 2)), rtol=1e-5, atol=1e-7)
    # decoding ignores the same constant; greedy labels are a fixed point of merge-repeats + drop-blank
    assert oc.greedy_decode(logits + shift, lens) == oc.greedy_decode(logits, lens)
    assert oc.beam_decode(logits + shift, lens, beam_width=8) == oc.beam_decode(logits, lens, beam_width=8)


def test_edit_distance_is_a_metric():
    rng = np.random.RandomState(3)
    seqs = [rng.randint(0, 5, size=rng.randint(0, 12)).tolist() for _ in range(12)]
    d = [[oc.edit_distance(a, b) for b in seqs] for a in seqs]
    for i, a in enumerate(seqs):
        assert d[i][i] == 0
        for j, b in enumerate(seqs):
            assert d[i][j] == d[j][i] and abs(len(a) - len(b)) <= d[i][j] <= max(len(a), len(b))
            for k in range(len(seqs)):
                assert d[i][j] <= d[i][k] + d[k][j]


def test_mfcc_after_cmvn_is_invariant_to_the_gain_of_the_recording():
    """Scaling the waveform shifts every log mel energy (and the log frame energy) by the same constant: only c0 moves,
    by a constant over time, and per-utterance mean normalisation removes it — the features of a louder copy of a clip
    are the same features."""
    rng = np.random.RandomState(4)
    clip = rng.randn(16000).astype(np.float32)
    for feat in (omf.MFCC(num_cep=13, d=True, dd=False), omf.MFCC(num_cep=13, d=True, dd=True), omf.LogFbank(num_filt=40)):
        a, b = feat(clip), feat(clip * 7.5)
        assert a.shape == b.shape
        assert np.abs(a - b).max() < 2e-5                  # float32 rounding of the scaled samples, amplified by 1/std
