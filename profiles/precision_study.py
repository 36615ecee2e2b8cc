#!/usr/bin/env python
"""CPU emulation of operand / storage precisions on the forward pass of the C2 stack (3 x BiLSTM-512, T = 999), against
the fp64 oracle: which 16-bit choices keep the logits inside the 1e-3 norm-wise bar.  Runs anywhere (numpy only).

  python profiles/precision_study.py [--frames 999] [--batch 4]

Variants: matmul operands rounded to fp16 / bf16 (what the tensor-core kernels do: fp32 accumulate, fp32 state), and
additionally the hoisted projection zx = x.W + b STORED in fp16 / bf16 (it is fp32 in HBM today: 524 MB per layer at
C2, the output the K = 1024 projection GEMM is co-limited by)."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import lstm as ol          # noqa: E402
from oracle import model as om         # noqa: E402


def bf16(a):
    u = np.asarray(a, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def fp16(a):
    return np.asarray(a, np.float32).astype(np.float16).astype(np.float32)


def lstm_dir(x, W, U, b, reverse, cast, zx_store):
    N, T, D = x.shape
    H = U.shape[0]
    Wc, Uc = cast(W), cast(U)
    zx = zx_store((cast(x).reshape(N * T, D) @ Wc).reshape(N, T, 4 * H) + b)
    h = np.zeros((N, H), np.float32)
    c = np.zeros((N, H), np.float32)
    out = np.zeros((N, T, H), np.float32)
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        z = zx[:, t] + cast(h) @ Uc
        i, f = ol.hard_sigmoid(z[:, :H]), ol.hard_sigmoid(z[:, H:2 * H])
        g, o = np.tanh(z[:, 2 * H:3 * H]), ol.hard_sigmoid(z[:, 3 * H:])
        c = f * c + i * g
        h = (o * np.tanh(c)).astype(np.float32)
        out[:, t] = h
    return out


def forward(params, x, cast, zx_store):
    h = x.astype(np.float32)
    L = om.num_layers_of(params)
    for l in range(L):
        p = {k.split(".", 1)[1]: v for k, v in params.items() if k.startswith(f"l{l}.")}
        h = np.concatenate([lstm_dir(h, p["Wf"], p["Uf"], p["bf"], False, cast, zx_store),
                            lstm_dir(h, p["Wb"], p["Ub"], p["bb"], True, cast, zx_store)], axis=2)
    N, T, D = h.shape
    return (cast(h).reshape(N * T, D) @ cast(params["dense.W"])).reshape(N, T, -1) + params["dense.b"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=999)
    ap.add_argument("--batch", type=int, default=4)
    args = ap.parse_args()
    F, H, L, C = 26, 512, 3, 28
    params = om.init_params(F, H, L, C, seed=4321)
    rng = np.random.RandomState(0)
    x = rng.randn(args.batch, args.frames, F).astype(np.float32)        # CMVN-normalised features are ~N(0, 1)
    ref, _ = om.forward(params, x, dtype=np.float64)
    ident = lambda a: np.asarray(a, np.float32)                          # noqa: E731
    res = {}
    for name, cast, store in (("fp32 operands, fp32 zx", ident, ident),
                              ("fp16 operands, fp32 zx (the kernels)", fp16, ident),
                              ("fp16 operands, fp16 zx", fp16, fp16),
                              ("fp16 operands, bf16 zx", fp16, bf16),
                              ("bf16 operands, fp32 zx", bf16, ident)):
        got = forward(params, x, cast, store)
        res[name] = float(np.abs(got - ref).max() / np.abs(ref).max())
    print(json.dumps({"T": args.frames, "N": args.batch, "bar": 1e-3, "logits_norm_err": res}, indent=1))


if __name__ == "__main__":
    main()
