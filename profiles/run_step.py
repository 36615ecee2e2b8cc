"""A few C2 training steps (MFCC -> 3xBiLSTM-512 -> CTC -> BPTT -> Adam, dropout 0.2) for ncu captures; no timing."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels  # noqa: E402
from asr_study_b200.preprocessing import audio  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
pcm_np, labels = bench.synth_batch(32, 1234)
pcm = torch.from_numpy(pcm_np.reshape(-1)).to(dev)
off = (torch.arange(33, dtype=torch.int64) * pcm_np.shape[1]).to(dev)
flat, loff, mx = pack_labels(labels, dev)
feat = audio.MFCC(num_cep=13, d=True, dd=False)
eng = AcousticEngine(ModelSpec(26, 512, 3, 28, weight_decay=1e-4, dropout=0.2), device=dev)
for _ in range(steps):
    x, lens = feat.batch(pcm, off, t_max=999, time_major=True)
    loss = eng.train_step(x, lens, flat, loff, mx, lr=1e-3, clipnorm=400.0)
torch.cuda.synchronize()
print("loss", float(loss.mean()), "status", eng.lstm_status())
