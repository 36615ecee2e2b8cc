#!/usr/bin/env python
"""Cross-check of the two routes a BiLSTM-800 stack can take (zero-padded tensor-core recurrences vs the fp32 general
cell) at full length: same parameters, same batch, no dropout, lr = 0 -> per-utterance loss, logits and every
parameter gradient of one train step must agree to the tensor-core bars (logits 1e-3, gradients 3e-2 norm-wise).

  python profiles/check_c4_routes.py [--frames 999] [--layers 5] [--batch 16]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=999)
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--hidden", type=int, default=800)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(3)
    N, T, L, H, F, C = args.batch, args.frames, args.layers, args.hidden, 40, 28
    x = torch.from_numpy(rng.randn(T, N, F).astype(np.float32)).to(dev)
    lens = torch.full((N,), T, dtype=torch.int32, device=dev)
    labels = [rng.randint(0, 25, size=rng.randint(2, 50)).astype(np.int32) for _ in range(N)]
    flat, loff, mx = pack_labels(labels, dev)
    params, D = {}, F
    for l in range(L):
        for d in "fb":
            lim = np.sqrt(6.0 / (D + 4 * H))
            params[f"l{l}.W{d}"] = rng.uniform(-lim, lim, size=(D, 4 * H)).astype(np.float32)
            params[f"l{l}.U{d}"] = (1.1 * np.linalg.qr(rng.randn(4 * H, H))[0].T).astype(np.float32)   # orthogonal(1.1), by QR
            params[f"l{l}.b{d}"] = np.concatenate([np.zeros(H), np.ones(H), np.zeros(2 * H)]).astype(np.float32)
        D = 2 * H
    lim = np.sqrt(6.0 / (D + C))
    params["dense.W"] = rng.uniform(-lim, lim, size=(D, C)).astype(np.float32)
    params["dense.b"] = np.zeros(C, np.float32)
    res = {}
    for pad in (True, False):
        os.environ["ASR_B200_PAD_WIDTH"] = "1" if pad else "0"
        eng = AcousticEngine(ModelSpec(F, H, L, C), device=dev, init_params=params)
        loss = eng.train_step(x, lens, flat, loff, mx, lr=0.0, clipnorm=400.0)
        torch.cuda.synchronize()
        assert eng.lstm_status() == 0
        res[pad] = dict(loss=loss.cpu().numpy().copy(), logits=eng.last_logits.cpu().numpy().copy(),
                        grads=eng.params.export("grad"), general=bool(eng._use_general), norm=eng.grad_norm())
        del eng
    a, b = res[True], res[False]

    def nerr(u, v):
        return float(np.abs(u - v).max() / max(np.abs(v).max(), 1e-30))

    out = {"T": T, "N": N, "L": L, "H": H, "routes": [("general" if r["general"] else "tensor-core") for r in (a, b)],
           "loss_tc": a["loss"][:4].tolist(), "loss_general": b["loss"][:4].tolist(),
           "loss_rel_max": float(np.abs(a["loss"] / b["loss"] - 1).max()), "logits_err": nerr(a["logits"], b["logits"]),
           "grad_norm": [a["norm"], b["norm"]],
           "grad_err": {k: nerr(a["grads"][k], b["grads"][k]) for k in a["grads"]}}
    out["grad_err_max"] = max(out["grad_err"].values())
    print(json.dumps(out))


if __name__ == "__main__":
    main()
