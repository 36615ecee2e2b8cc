"""K10 parity: asr_edit_distance (tf.edit_distance(hyp, truth, normalize=True), core/metrics.py:4-8) vs the host
dynamic program of asr_study_b200.core.metrics / oracle.ctc.ler — integer work, bit-exact distances."""
import numpy as np
import pytest
import torch

from oracle import ctc as oc

pytestmark = pytest.mark.gpu


def _lev(a, b):
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i] + [0] * len(b)
        for j, cb in enumerate(b, 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb))
        prev = cur
    return prev[len(b)]


def _run(hyps, truths, stride, with_len=True, normalize=True):
    from asr_study_b200._lib import lib, ptr, cur_stream
    from asr_study_b200.engine import pack_labels
    N = len(hyps)
    mat = -np.ones((N, stride), np.int32)
    for i, h in enumerate(hyps):
        mat[i, :len(h)] = h
    hyp = torch.as_tensor(mat, device="cuda")
    hl = torch.as_tensor(np.array([len(h) for h in hyps], np.int32), device="cuda")
    flat, off, mx = pack_labels([np.asarray(t, np.int32) for t in truths], "cuda")
    out = torch.empty(N, dtype=torch.float32, device="cuda")
    lib.asr_edit_distance(ptr(hyp), N, stride, ptr(hl) if with_len else None, ptr(flat), ptr(off), mx, int(normalize),
                          ptr(out), cur_stream())
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("with_len", [True, False])
def test_edit_distance_matches_host_dp(with_len):
    rng = np.random.RandomState(7)
    hyps, truths = [], []
    for _ in range(64):                                   # the regime of the path: a few dozen labels of 28 classes
        truths.append(rng.randint(0, 27, size=rng.randint(1, 60)).tolist())
        h = list(truths[-1])
        for _ in range(rng.randint(0, 12)):               # a noisy copy: substitutions, insertions, deletions
            op, pos = rng.randint(3), rng.randint(0, max(1, len(h)))
            if op == 0 and h:
                h[pos % len(h)] = int(rng.randint(0, 27))
            elif op == 1:
                h.insert(pos, int(rng.randint(0, 27)))
            elif h:
                del h[pos % len(h)]
        hyps.append(h)
    hyps += [[], [], [3, 4], list(rng.randint(0, 27, size=999)), [5] * 40, list(range(20))]
    truths += [[1, 2, 3], [], [], list(rng.randint(0, 27, size=7)), list(rng.randint(0, 3, size=300)), list(range(20))]
    got = _run(hyps, truths, 999, with_len=with_len)
    for h, t, g in zip(hyps, truths, got):
        d = _lev(h, t)
        want = (d / len(t)) if t else (0.0 if not h else float("inf"))
        assert g == np.float32(want), (len(h), len(t), g, want)
    raw = _run(hyps, truths, 999, with_len=with_len, normalize=False)
    for h, t, g in zip(hyps, truths, raw):
        assert g == float(_lev(h, t))
    # batch mean = core.metrics.ler = the oracle's ler on the utterances with a non-empty truth
    keep = [i for i, t in enumerate(truths) if t]
    from asr_study_b200.core import metrics
    assert abs(float(got[keep].mean()) - metrics.ler([truths[i] for i in keep], [hyps[i] for i in keep])) < 1e-6
    assert abs(float(got[keep].mean()) - oc.ler([truths[i] for i in keep], [hyps[i] for i in keep])) < 1e-6
