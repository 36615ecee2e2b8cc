#!/usr/bin/env python
"""CPU emulation (numpy, fp32 lattice + fp64 per-frame offsets, the arithmetic of csrc/ctc.cu) of two renormalisation
schedules of the CTC alpha / beta recursions at T = 999 against the fp64 oracle:

  every frame, maximum of the PREVIOUS row        (what the kernel does: the warp / block maximum sits on the chain)
  every frame, maximum of the row BEFORE that     (one frame late: the maximum leaves the dependent chain)
  every k-th frame

Reports max |d loss / d logits - oracle| (bar at T = 999: 5e-4 absolute) and the loss error.  Runs anywhere."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ctc as oc          # noqa: E402

F32 = np.float32
NEG = F32(-np.inf)


def lse3(a, b, c):
    m = np.maximum(np.maximum(a, b), c)
    with np.errstate(invalid="ignore"):
        s = np.exp(a - m, dtype=F32) + np.exp(b - m, dtype=F32) + np.exp(c - m, dtype=F32)
        r = m + np.log(s, dtype=F32)
    return np.where(np.isfinite(m), r, NEG).astype(F32)


def lattice(lp, ext, blank, schedule, reverse):
    """lp [T, C] fp32 log-softmax; returns rows [T, S] fp32 and offsets [T] fp64 (row t is relative to offsets[t])."""
    T, S = lp.shape[0], len(ext)
    em = lp[:, ext]                                          # [T, S]
    if reverse:
        em = em[::-1]
        skip = np.array([s + 2 < S and ext[s + 2] != blank and ext[s + 2] != ext[s] for s in range(S)])
    else:
        skip = np.array([s >= 2 and ext[s] != blank and ext[s] != ext[s - 2] for s in range(S)])
    rows = np.full((T, S), NEG, F32)
    offs = np.zeros(T, np.float64)
    init = np.full(S, NEG, F32)
    if reverse:
        init[S - 1] = em[0, S - 1]
        if S > 1:
            init[S - 2] = em[0, S - 2]
    else:
        init[0] = em[0, 0]
        if S > 1:
            init[1] = em[0, 1]
    rows[0] = init
    off = 0.0
    for t in range(1, T):
        prev = rows[t - 1]
        if reverse:
            p1 = np.concatenate([prev[1:], [NEG]])
            p2 = np.where(skip, np.concatenate([prev[2:], [NEG, NEG]]), NEG)
        else:
            p1 = np.concatenate([[NEG], prev[:-1]])
            p2 = np.where(skip, np.concatenate([[NEG, NEG], prev[:-2]]), NEG)
        v = lse3(prev, p1.astype(F32), p2.astype(F32))
        M = F32(0.0)
        if schedule == "prev":
            M = prev.max()
        elif schedule == "lag":
            M = rows[t - 2].max() if t >= 2 else F32(0.0)
        elif schedule.startswith("every"):
            k = int(schedule[5:])
            M = prev.max() if t % k == 0 else F32(0.0)
        if not np.isfinite(M):
            M = F32(0.0)
        rows[t] = np.where(np.isfinite(v), v - M + em[t], NEG).astype(F32)
        off += float(M)
        offs[t] = off
    if reverse:
        rows, offs = rows[::-1], offs[::-1]
    return rows, offs


def loss_grad(logits, labels, blank, schedule):
    T, C = logits.shape
    x = logits.astype(F32)
    m = x.max(axis=1, keepdims=True)
    lse = (m + np.log(np.exp(x - m, dtype=F32).sum(axis=1, keepdims=True), dtype=F32)).astype(F32)
    lp = (x - lse).astype(F32)
    ext = [blank]
    for l in labels:
        ext += [int(l), blank]
    ext = np.array(ext)
    S = len(ext)
    A, offA = lattice(lp, ext, blank, schedule, False)
    B, offB = lattice(lp, ext, blank, schedule, True)
    last = A[T - 1]
    tail = np.logaddexp(np.float64(last[S - 1]), np.float64(last[S - 2])) if S > 1 else np.float64(last[S - 1])
    logp = offA[T - 1] + tail
    grad = np.zeros((T, C), F32)
    for t in range(T):
        v = A[t] + B[t] - lp[t, ext]                          # beta here includes the emission of frame t, like alpha
        kf = F32(offA[t] + offB[t] - logp)
        mm = v[np.isfinite(v)].max()
        e = np.where(np.isfinite(v), np.exp(v - mm, dtype=F32), F32(0))
        occ = np.zeros(C, F32)
        np.add.at(occ, ext, e)
        grad[t] = np.exp(lp[t], dtype=F32) - occ * np.exp(mm + kf, dtype=F32)
    return -logp, grad


def main():
    rng = np.random.RandomState(2)
    T, C, blank = 999, 28, 27
    out = {}
    for scale in (1.0, 3.0):
        logits = (rng.randn(T, C) * scale).astype(np.float32)
        labels = rng.randint(0, 25, size=40)
        labels[1] = labels[0]
        rl, rg = oc.ctc_loss_grad_single(logits.astype(np.float64), T, list(labels), blank)
        for sched in ("prev", "lag", "every4", "every8", "every16"):
            l, g = loss_grad(logits, labels, blank, sched)
            out[f"logit scale {scale}: {sched}"] = {"loss_rel_err": float(abs(l - rl) / rl), "grad_max_abs_err": float(np.abs(g - rg).max())}
    print(json.dumps({"T": T, "bar_grad_abs": 5e-4, "results": out}, indent=1))


if __name__ == "__main__":
    main()
