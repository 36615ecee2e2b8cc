#!/usr/bin/env python
"""BASELINE config 4's recurrent stack on one B200: 40 log-mel -> 5 x BiLSTM-800 -> Dense-28 -> CTC, train step,
16 utterances per GPU (128 over 8 GPUs), T = 999 (the DS2-style conv front end is not in the reference and is not
built: the stack sees the raw log-fbank frames).  Times the zero-padded tensor-core route (800 -> 832 units,
engine.tc_width) against the general-cell route (ASR_B200_PAD_WIDTH=0) on the same synthetic batch.

  python profiles/bench_c4_stack.py [--steps 5] [--batch 16] [--frames 999]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(pad, args):
    import torch
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    from asr_study_b200.preprocessing import audio
    os.environ["ASR_B200_PAD_WIDTH"] = "1" if pad else "0"
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(3)
    N, T = args.batch, args.frames
    feat = audio.LogFbank(num_filt=40)
    n_samples = 400 + 160 * (T - 1)
    pcm = torch.from_numpy(rng.randn(N * n_samples).astype(np.float32)).to(dev)
    off = (torch.arange(N + 1, dtype=torch.int64) * n_samples).to(dev)
    x, lens = feat.batch(pcm, off, t_max=T, time_major=True)
    labels = [rng.randint(0, 25, size=rng.randint(2, 50)).astype(np.int32) for _ in range(N)]
    flat, loff, mx = pack_labels(labels, dev)
    spec = ModelSpec(40, 800, 5, 28, weight_decay=1e-4, dropout=0.2)
    params = {}
    D = 40
    for l in range(5):
        for d in "fb":
            lim = np.sqrt(6.0 / (D + 3200))
            params[f"l{l}.W{d}"] = rng.uniform(-lim, lim, size=(D, 3200)).astype(np.float32)
            params[f"l{l}.U{d}"] = (1.1 * np.linalg.qr(rng.randn(3200, 800))[0].T).astype(np.float32)   # orthogonal(1.1), by QR
            params[f"l{l}.b{d}"] = np.concatenate([np.zeros(800), np.ones(800), np.zeros(1600)]).astype(np.float32)
        D = 1600
    lim = np.sqrt(6.0 / (1600 + 28))
    params["dense.W"] = rng.uniform(-lim, lim, size=(1600, 28)).astype(np.float32)
    params["dense.b"] = np.zeros(28, np.float32)
    eng = AcousticEngine(spec, device=dev, init_params=params)

    def step():
        return eng.train_step(x, lens, flat, loff, mx, lr=1e-3, clipnorm=400.0)

    for _ in range(2):
        loss = step()
    torch.cuda.synchronize()
    assert eng.lstm_status() == 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    gflop = 3 * (2 * T * 2 * ((40 + 800) + 4 * (1600 + 800)) * 3200 + 2 * T * 1600 * 28) / 1e9     # per utterance, fwd+dX+dW
    return {"route": "tensor-core, 800 -> 832 zero-padded" if pad else "general cell (fp32 CUDA cores)",
            "device_width": eng.spec.num_hiddens, "general": bool(eng._use_general), "ms_per_step": ms,
            "utt_per_s": N / (ms / 1e3), "tflops": N * gflop / ms, "loss_mean": float(loss.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--frames", type=int, default=999)
    ap.add_argument("--skip-general", action="store_true")
    args = ap.parse_args()
    out = {"workload": f"C4 stack: 40 log-mel, 5 x BiLSTM-800, Dense-28, CTC, Adam, dropout 0.2, N={args.batch}, T={args.frames}",
           "runs": [run(True, args)]}
    if not args.skip_general:
        out["runs"].append(run(False, args))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
