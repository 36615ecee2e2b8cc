"""Where the plugin-surface e2e step spends its time beyond the device-resident step (C2 shape, one B200).

Variants, each timed with CUDA events over `steps` steps after a warm-up, plus the host time of one call (how far ahead of
the GPU the host runs):
  engine        engine.train_step on device-resident features (bench.py's `value` without the MFCC launch)
  tob_fixed     CTCModel.train_on_batch on ONE fixed host batch (no generator thread): H2D + step + metrics
  tob_nostats   the same without the metric kernels (decode, label error rate, l2 penalty, read-back)
  plugin        DatasetIterator on a generator thread -> train_on_batch (bench.py's `e2e`)

usage: python profiles/e2e_breakdown.py [steps]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    from asr_study_b200.core import models
    from asr_study_b200.core.models import _GeneratorFeed
    from asr_study_b200.datasets.dataset_generator import DatasetIterator
    from asr_study_b200.engine import pack_labels
    from asr_study_b200.preprocessing import audio

    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    dev = torch.device("cuda:0")
    nb = 32
    pcm_np, labels = bench.synth_batch(nb, 1234)
    feat = audio.MFCC(num_cep=13, d=True, dd=False)
    model = models.brsmv1(num_features=26, num_hiddens=512, num_layers=3, num_classes=28, dropout=0.2, weight_decay=1e-4,
                          device=str(dev), seed=4321)
    model.compile(optimizer=models.Adam(lr=1e-3, clipnorm=400.0))
    eng = model.engine
    flow = DatasetIterator([pcm_np[i] for i in range(nb)], [np.asarray(l, np.int32) for l in labels], batch_size=nb,
                           shuffle=False, input_parser=feat, label_parser=None, rank=0, world_size=1)
    x_fixed, _ = flow.next()
    flat, loff, mx = pack_labels(labels, dev)
    pcm_dev = torch.from_numpy(pcm_np.reshape(-1)).to(dev)
    off_dev = (torch.arange(nb + 1, dtype=torch.int64) * pcm_np.shape[1]).to(dev)
    xt, lens = feat.batch(pcm_dev, off_dev, t_max=bench.T_FRAMES, time_major=True)
    xt = xt.clone()

    def engine_step():
        eng.train_step(xt, lens, flat, loff, mx, global_batch=nb, lr=1e-3, clipnorm=400.0)

    last = {"m": None}

    def tob_fixed():
        m = model.train_on_batch(x_fixed)
        if last["m"] is not None:
            last["m"].result()
        last["m"] = m

    def tob_nostats():
        xd, ln, (fl, of, mxl), N = model._device_batch(x_fixed[0], x_fixed[2], x_fixed[1], True)
        eng.train_step(xd, ln, fl, of, mxl, global_batch=N, lr=1e-3, clipnorm=400.0)

    feed = {"f": None}

    def plugin():
        x, _y = feed["f"].get()
        m = model.train_on_batch(x)
        if last["m"] is not None:
            last["m"].result()
        last["m"] = m

    def timed(fn, n):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        host = []
        e0.record()
        for _ in range(n):
            t = time.perf_counter()
            fn()
            host.append(time.perf_counter() - t)
        e1.record()
        torch.cuda.synchronize()
        return {"ms_per_step": e0.elapsed_time(e1) / n, "host_ms_median": float(np.median(host)) * 1e3,
                "host_ms_min": float(np.min(host)) * 1e3}

    out = {}
    out["engine"] = timed(engine_step, steps)
    out["tob_fixed"] = timed(tob_fixed, steps)
    last["m"] = None
    out["tob_nostats"] = timed(tob_nostats, steps)
    feed["f"] = _GeneratorFeed(flow, (steps + 3) * nb, 10, 1, dev)
    out["plugin"] = timed(plugin, steps)
    feed["f"].close()
    out["engine_again"] = timed(engine_step, steps)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
