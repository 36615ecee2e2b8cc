"""Oracle: CTC loss/gradient, best-path and prefix-beam decode, LER (numpy).

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  PARITY PINNED ON THE DEPENDENCY'S OWN KNOWN ANSWERS (the
reference itself holds no test or golden vector for this path): loss, gradient, greedy and beam decode and the
edit distance reproduce the constants of TensorFlow's ctc_loss_op_test.py / ctc_decoder_ops_test.py and the
tf.edit_distance docstring (tests/golden/ctc_tf_*.json, tests/test_oracle_ctc.py).  The
reference only *calls* TensorFlow 1.3.0 here (un-vendored, not installed):
  core/ctc_utils.py:68-70  tf.nn.ctc_loss(labels, time-major logits, seq_len)
  core/ctc_utils.py:42     tf.nn.ctc_greedy_decoder(y_pred, seq_len)
  core/ctc_utils.py:48-50  tf.nn.ctc_beam_search_decoder(..., beam_width,
                           top_paths, merge_repeated)[0][0]
  core/metrics.py:8        tf.reduce_mean(tf.edit_distance(hyp, truth))
This file restates the published algorithms with TF-1.3 conventions (SURVEY.md
8c hypotheses 4-7): softmax applied internally, blank = num_classes-1, standard
merge-repeated topology, zero gradient past seq_len, argmax ties -> lowest
index.  Further independent pins live in tests/ (brute-force path enumeration,
torch.nn.functional.ctc_loss, finite differences).
"""
from __future__ import annotations

import itertools

import numpy as np

NEG_INF = -np.inf


def log_softmax(x, axis=-1):
    x = np.asarray(x, dtype=np.float64)
    m = x.max(axis=axis, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(axis=axis, keepdims=True))


def _shift(a, k):
    """a shifted right by k (k>0) or left (k<0), filled with -inf, same length."""
    out = np.full_like(a, NEG_INF)
    n = len(a)
    if k > 0 and k < n:
        out[k:] = a[:n - k]
    elif k < 0 and -k < n:
        out[:n + k] = a[-k:]
    elif k == 0:
        out[:] = a
    return out


def _lse(*xs):
    m = np.maximum.reduce(xs)
    m_safe = np.where(np.isfinite(m), m, 0.0)
    s = sum(np.exp(x - m_safe) for x in xs)
    with np.errstate(divide="ignore"):
        return np.where(np.isfinite(m), m_safe + np.log(s), NEG_INF)


def ctc_loss_grad_single(logits, seq_len, labels, blank):
    """One utterance.  logits [T,C] (unnormalised), labels: list of ints.

    Returns (loss = -ln p(labels|x), dloss/dlogits [T,C]); rows >= seq_len are 0.
    """
    logits = np.asarray(logits, dtype=np.float64)
    T, C = logits.shape
    seq_len = int(seq_len)
    L = len(labels)
    S = 2 * L + 1
    ext = np.full(S, blank, dtype=np.int64)
    ext[1::2] = labels
    lp = log_softmax(logits[:seq_len], axis=1)            # [t, C]
    # skip[s] : transition s-2 -> s allowed
    skip = np.zeros(S, dtype=bool)
    skip[2:] = (ext[2:] != blank) & (ext[2:] != ext[:-2])
    alpha = np.full((seq_len, S), NEG_INF)
    alpha[0, 0] = lp[0, blank]
    if S > 1:
        alpha[0, 1] = lp[0, ext[1]]
    for t in range(1, seq_len):
        a = alpha[t - 1]
        a1 = _shift(a, 1)
        a2 = np.where(skip, _shift(a, 2), NEG_INF)
        alpha[t] = _lse(a, a1, a2) + lp[t, ext]
    beta = np.full((seq_len, S), NEG_INF)
    beta[seq_len - 1, S - 1] = lp[seq_len - 1, blank]
    if S > 1:
        beta[seq_len - 1, S - 2] = lp[seq_len - 1, ext[S - 2]]
    skip_f = np.zeros(S, dtype=bool)                     # s -> s+2 allowed
    if S > 2:
        skip_f[:-2] = skip[2:]
    for t in range(seq_len - 2, -1, -1):
        b = beta[t + 1]
        b1 = _shift(b, -1)
        b2 = np.where(skip_f, _shift(b, -2), NEG_INF)
        beta[t] = _lse(b, b1, b2) + lp[t, ext]
    ends = [alpha[seq_len - 1, S - 1]] + ([alpha[seq_len - 1, S - 2]] if S > 1 else [])
    log_p = float(_lse(*[np.asarray(e) for e in ends]))
    loss = -log_p
    grad = np.zeros((T, C))
    if np.isfinite(log_p):
        ab = alpha + beta                                # [t, S]
        acc = np.full((seq_len, C), NEG_INF)
        for s in range(S):
            acc[:, ext[s]] = _lse(acc[:, ext[s]], ab[:, s])
        with np.errstate(invalid="ignore"):
            occ = np.exp(acc - lp - log_p)
        occ = np.where(np.isfinite(acc), occ, 0.0)
        grad[:seq_len] = np.exp(lp) - occ
    return loss, grad


def ctc_loss_grad(logits, seq_lens, labels_list, blank=None, dtype=np.float32):
    """Batch form of tf.nn.ctc_loss + its gradient.

    logits [N,T,C] batch-major (the reference transposes to time-major itself,
    core/ctc_utils.py:69).  Returns (loss [N] f32, grad [N,T,C] f32).
    """
    logits = np.asarray(logits)
    N, T, C = logits.shape
    blank = C - 1 if blank is None else blank
    loss = np.zeros(N, dtype=np.float64)
    grad = np.zeros((N, T, C), dtype=np.float64)
    for n in range(N):
        loss[n], grad[n] = ctc_loss_grad_single(logits[n], seq_lens[n],
                                                list(labels_list[n]), blank)
    return loss.astype(dtype), grad.astype(dtype)


def brute_force_label_probs(logits, blank):
    """Enumerate all C^T alignments of a tiny [T,C] problem.

    Returns dict {label tuple: probability} under the CTC collapse map.
    """
    logits = np.asarray(logits, dtype=np.float64)
    T, C = logits.shape
    p = np.exp(log_softmax(logits, axis=1))
    out = {}
    for path in itertools.product(range(C), repeat=T):
        pr = 1.0
        for t, k in enumerate(path):
            pr *= p[t, k]
        lab, prev = [], None
        for k in path:
            if k != prev and k != blank:
                lab.append(k)
            prev = k
        out[tuple(lab)] = out.get(tuple(lab), 0.0) + pr
    return out


# --------------------------------------------------------------------------- #
# decoders
# --------------------------------------------------------------------------- #
def greedy_decode_single(logits, seq_len, blank, merge_repeated=True):
    """tf.nn.ctc_greedy_decoder for one utterance: per-frame argmax (first max
    wins), drop repeats, drop blanks."""
    out, prev = [], -1
    for t in range(int(seq_len)):
        k = int(np.argmax(logits[t]))
        if k != blank and not (merge_repeated and k == prev):
            out.append(k)
        prev = k
    return out


def greedy_decode(logits, seq_lens, blank=None):
    logits = np.asarray(logits)
    blank = logits.shape[2] - 1 if blank is None else blank
    return [greedy_decode_single(logits[n], seq_lens[n], blank)
            for n in range(logits.shape[0])]


class _Entry:
    __slots__ = ("parent", "label", "children", "ob", "ol", "ot", "nb", "nl", "nt")

    def __init__(self, parent, label):
        self.parent, self.label, self.children = parent, label, None
        self.ob = self.ol = self.ot = NEG_INF      # oldp blank/label/total
        self.nb = self.nl = self.nt = NEG_INF      # newp

    def active(self):
        return self.nt != NEG_INF


def _lse2(a, b):
    if a == NEG_INF:
        return b
    if b == NEG_INF:
        return a
    # TF ctc_loss_util.h LogSumExp: max + log1pf(expf(-|a-b|)); the two float functions are emulated as
    # "fp64 then round to fp32" (what a correctly rounded libm returns) so any platform reproduces it.
    f32 = np.float32
    a, b = f32(a), f32(b)
    m, mn = (a, b) if a >= b else (b, a)
    e = f32(np.exp(np.float64(f32(mn - m))))
    return f32(m + f32(np.log1p(np.float64(e))))


def beam_decode_single(logits, seq_len, blank, beam_width=100,
                       merge_repeated=True):
    """TF CTCBeamSearchDecoder (tensorflow/core/util/ctc/ctc_beam_search.h,
    v1.3), top path only, restated from the published algorithm:

    per frame, scores = logits - max(logits); every current leaf updates
    (blank, label) mass from itself and its still-active parent; then leaves, in
    descending old-score order, spawn inactive children that enter the beam iff
    their score beats the current worst leaf (strictly) or the beam is not full.
    Scores are kept in float32 like TF.  Returns the label list of the best
    leaf, with adjacent repeats collapsed when merge_repeated (the TF op quirk).
    """
    logits = np.asarray(logits, dtype=np.float32)
    C = logits.shape[1]
    f32 = np.float32
    root = _Entry(None, -1)
    root.nt, root.nb = f32(0.0), f32(0.0)
    leaves = [root]
    for t in range(int(seq_len)):
        inp = (logits[t] - logits[t].max()).astype(np.float32)
        branches = sorted(leaves, key=lambda e: -e.nt)        # descending newp
        for b in branches:
            b.ob, b.ol, b.ot = b.nb, b.nl, b.nt
        leaves = []
        for b in branches:
            if b.parent is not None:
                if b.parent.active():
                    prev = b.parent.ob if b.label == b.parent.label else b.parent.ot
                    b.nl = f32(_lse2(b.nl, prev))
                b.nl = f32(b.nl + inp[b.label])
            b.nb = f32(b.ot + inp[blank])
            b.nt = f32(_lse2(b.nb, b.nl))
            leaves.append(b)

        def bottom():
            return min(leaves, key=lambda e: e.nt)

        def is_candidate(total):
            return total > NEG_INF and (len(leaves) < beam_width or
                                        total > bottom().nt)

        for b in branches:
            if not is_candidate(b.ot):
                continue
            if b.children is None:
                b.children = [_Entry(b, k) for k in range(C) if k != blank]
            for c in b.children:
                if c.active():
                    continue
                prev = b.ob if c.label == b.label else b.ot
                c.nb = NEG_INF
                c.nl = f32(inp[c.label] + prev) if prev != NEG_INF else NEG_INF
                c.nt = c.nl
                if is_candidate(c.nt):
                    if len(leaves) == beam_width:
                        worst = bottom()
                        leaves.remove(worst)
                        worst.nb = worst.nl = worst.nt = NEG_INF
                    leaves.append(c)
                else:
                    c.ob = c.ol = c.ot = NEG_INF
                    c.nb = c.nl = c.nt = NEG_INF
    best = max(leaves, key=lambda e: e.nt)
    labels, prev, e = [], -1, best
    while e.parent is not None:
        if not merge_repeated or e.label != prev:
            labels.append(e.label)
        prev = e.label
        e = e.parent
    return labels[::-1]


def beam_decode(logits, seq_lens, blank=None, beam_width=100, merge_repeated=True):
    logits = np.asarray(logits)
    blank = logits.shape[2] - 1 if blank is None else blank
    return [beam_decode_single(logits[n], seq_lens[n], blank, beam_width,
                               merge_repeated) for n in range(logits.shape[0])]


# --------------------------------------------------------------------------- #
# label error rate
# --------------------------------------------------------------------------- #
def edit_distance(a, b) -> int:
    """Levenshtein distance between two int sequences."""
    a, b = list(a), list(b)
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1,
                         prev[j - 1] + (a[i - 1] != b[j - 1]))
        prev = cur
    return prev[len(b)]


def ler(truths, hyps) -> float:
    """core/metrics.py:4-8 — mean over the batch of
    tf.edit_distance(hyp, truth, normalize=True) = lev(hyp,truth)/len(truth)."""
    vals = []
    for t, h in zip(truths, hyps):
        d = edit_distance(h, t)
        vals.append(d / len(t) if len(t) else (float("inf") if d else 0.0))
    return float(np.mean(vals))
