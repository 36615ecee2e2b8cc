"""Timeline of ONE C2 training step: a CUDA event before and after every C-ABI call, on the stream the call is enqueued on,
without any synchronisation in between (the step runs as in bench.py: side-stream GEMMs and operand preparation included).
Prints, per call: stream, start offset from the step's first event, duration between its two events (kernel time plus any
wait for a cross-stream dependency in front of it), and the idle gap to the previous call on the same stream.

usage: python profiles/step_timeline.py [c4]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    import torch

    import asr_study_b200.engine as E
    from asr_study_b200.core import models
    from asr_study_b200.engine import pack_labels
    from asr_study_b200.preprocessing import audio

    c4 = len(sys.argv) > 1 and sys.argv[1] == "c4"
    dev = torch.device("cuda:0")
    nb = 16 if c4 else 32
    pcm_np, labels = bench.synth_batch(nb, 1234)
    if c4:
        feat = audio.LogFbank()
        model = models.deep_speech2(num_features=40, num_hiddens=800, num_layers=5, num_classes=28, dropout=0.2,
                                    weight_decay=1e-4, device=str(dev), seed=4321)
    else:
        feat = audio.MFCC(num_cep=13, d=True, dd=False)
        model = models.brsmv1(num_features=26, num_hiddens=512, num_layers=3, num_classes=28, dropout=0.2,
                              weight_decay=1e-4, device=str(dev), seed=4321)
    model.compile(optimizer=models.Adam(lr=1e-3, clipnorm=400.0))
    eng = model.engine
    flat, loff, mx = pack_labels(labels, dev)
    pcm_dev = torch.from_numpy(pcm_np.reshape(-1)).to(dev)
    off_dev = (torch.arange(nb + 1, dtype=torch.int64) * pcm_np.shape[1]).to(dev)
    xt, lens = feat.batch(pcm_dev, off_dev, t_max=bench.T_FRAMES, time_major=True)
    xt = xt.clone()

    def step():
        eng.train_step(xt, lens, flat, loff, mx, global_batch=nb, lr=1e-3, clipnorm=400.0)

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    rec = []

    class Tap:
        def __init__(self, real):
            self.real = real

        def __getattr__(self, n):
            f = getattr(self.real, n)
            if not n.startswith("asr_") or n.endswith("_bytes") or n in ("asr_launch_count", "asr_last_error") \
                    or "supported" in n or "fuses" in n or "storage" in n or "out_shape" in n:
                return f

            def call(*a):
                st = torch.cuda.current_stream()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                r = f(*a)
                e1.record(st)
                shape = ""
                if n.startswith("asr_gemm"):
                    shape = "M%d N%d K%d" % (a[2], a[3], a[4])
                rec.append((n, st.cuda_stream, e0, e1, shape))
                return r
            return call

    real = E.lib
    e_begin, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    E.lib = Tap(real)
    try:
        step()                              # a first traced step keeps the GPU busy while the host enqueues the second one:
        del rec[:]                          # the printed step is in steady state (host ahead of the device), not host-paced
        e_begin.record()
        step()
        e_end.record()
    finally:
        E.lib = real
    torch.cuda.synchronize()
    streams = {}
    last_end = {}
    print("step: %.3f ms (traced, second of two back-to-back steps; the events add ~1 us per call)" % e_begin.elapsed_time(e_end))
    print("%-28s %3s %9s %8s %8s  %s" % ("call", "str", "start", "dur", "gap", "shape"))
    tot = {}
    for n, sid, e0, e1, shape in rec:
        k = streams.setdefault(sid, len(streams))
        start, end = e_begin.elapsed_time(e0), e_begin.elapsed_time(e1)
        gap = start - last_end.get(k, 0.0)
        last_end[k] = end
        print("%-28s %3d %9.3f %8.3f %8.3f  %s" % (n, k, start, end - start, gap, shape))
        if k == 0:
            tot[n] = tot.get(n, 0.0) + (end - start)
            tot["(gap)"] = tot.get("(gap)", 0.0) + max(gap, 0.0)
    print(json.dumps({k: round(v, 3) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])}))


if __name__ == "__main__":
    main()
