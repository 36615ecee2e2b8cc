#!/usr/bin/env python
"""bench.py — utterances/sec of the acoustic training hot path on B200 (BASELINE.json metric).

Workload (configs[1] / C2): synthetic 16 kHz 10 s clips -> 26-dim MFCC (fused kernel) ->
3 x BiLSTM-512 -> Dense-28 -> CTC loss+grad -> BPTT -> global-norm clip + Adam, batch 32 per GPU.
N > 1: data parallel (configs[2] shape), one NCCL all-reduce of the flat gradient bucket per step.

  python bench.py --gpus 1 --steps 5 --warmup 3
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference        # CPU arm: oracle port of the reference path on the host cores

Prints ONE JSON line (see the task contract): value = device-resident throughput; e2e = through the reference-facing
plugin surface: a DatasetIterator (datasets/dataset_generator.py contract: host pcm in, host feature batches out, on
Keras' generator thread) feeding CTCModel.train_on_batch (host features in, host metrics out), every copy inside the timed
region; e2e_engine = the engine-level variant (pinned host pcm -> H2D -> step -> D2H loss).  At N = 1 the line also
carries `infer` (BASELINE configs[4]: 10 k clips, beam 100, LER parity on 256 clips) and `cpu_baseline`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SECONDS, FS = 10.0, 16000
F, H, L, C = 26, 512, 3, 28
T_FRAMES = 999
GFLOP_TRAIN_PER_UTT = 88.80     # fwd+dX+dW LSTM GEMMs + Dense, T=999
METRIC = "utterances/sec (10 s, 26-MFCC, 3xBiLSTM-512, CTC)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled DURING the timed region: NVML in-process (a sample every 20 ms
    from a thread; `nvidia-smi` needs longer than a short timed region just to start on an 8-GPU box), with the
    `nvidia-smi -lms` query of the profiling recipe as the fallback.  prepare() before the warm-up, start() / stop()
    around the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.nvml = self.handle = self.thread = None
        self.sm, self.reasons, self.max_sm, self.on = [], set(), None, False

    def prepare(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        nv = self.nvml
        names = (("hw_slowdown", nv.nvmlClocksEventReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown),
                 ("sw_thermal_slowdown", nv.nvmlClocksEventReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksEventReasonSwPowerCap))
        while self.on:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for name, bit in names:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nvml is not None:
            self.on = True
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.on = False
            self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_sm,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml"}
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def synth_batch(n, seed):
    from oracle.model import synth_clip, synth_labels   # the *spec* of the synthetic workload (datasets/dummy.py)
    pcm = np.stack([synth_clip(seed, i, SECONDS, FS) for i in range(n)])
    labels = synth_labels(seed + 7, n, 50)
    return pcm, labels


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path on the host cores
# --------------------------------------------------------------------------------------
def cpu_step(pcm, labels, params, state, dropout=0.2):
    from oracle import mfcc as omf
    from oracle import model as om
    feat = omf.MFCC(num_cep=13, d=True, dd=(F == 39))
    x, lens = omf.pad_batch([feat(c) for c in pcm])
    masks = None
    if dropout > 0:                                        # the same work as the GPU arm: fresh variational masks per step
        rng = state.setdefault("mask_rng", np.random.RandomState(17))
        masks, D, n = {}, F, len(pcm)
        for l in range(L):
            masks[l] = {k: ((rng.rand(n, w) >= dropout) / (1.0 - dropout)).astype(np.float32)
                        for k, w in (("Wf", D), ("Wb", D), ("Uf", H), ("Ub", H))}
            D = 2 * H
    _, ctc, grads, _ = om.loss_and_grads(params, x, lens, labels, weight_decay=1e-4, masks=masks, dtype=np.float32)
    om.clip_adam_step(params, grads, state.setdefault("adam", {}), lr=1e-3, clipnorm=400.0)
    return float(ctc.mean())


CPU_SAMPLE = 32          # both CPU legs time the same thing: one full step on all 32 clips of the C2 batch


def cpu_arm(n_sample, steps, warmup):
    """The oracle port on ALL host cores whatever the launcher exported: torch.distributed.run sets OMP_NUM_THREADS=1
    for its children, which would pin numpy's BLAS to one thread; the pool size is set here, at run time."""
    from threadpoolctl import threadpool_limits

    from oracle import model as om
    cores = os.cpu_count() or 1
    params = om.init_params(F, H, L, C, seed=4321)
    pcm, labels = synth_batch(n_sample, 1234)
    state = {}
    with threadpool_limits(limits=cores):
        for _ in range(warmup):
            cpu_step(pcm, labels, params, state)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_step(pcm, labels, params, state)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    return n_sample / dt, dt


def cpu_sample_text(n_sample, dt):
    return (f"all {n_sample} clips of one C2 step (full 10 s, T=999): oracle port incl. MFCC + BPTT + clip + Adam, numpy BLAS "
            f"pool set to all {os.cpu_count()} cores by threadpoolctl (launcher-independent), {dt:.1f} s per step")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n_sample = CPU_SAMPLE
    steps, warmup = max(1, min(args.steps, 2)), 0
    val, dt = cpu_arm(n_sample, steps, warmup)
    cores = os.cpu_count()
    sample = cpu_sample_text(n_sample, dt)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "utt/s", "n_gpus": args.gpus, "steps": steps,
           "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": ("C2" if (F, H, L) == (26, 512, 3) else "custom (not the BASELINE config)") +
                                  ": synthetic 16 kHz 10 s clips, %d-MFCC, %dxBiLSTM-%d, Dense-28, CTC, " % (F, L, H) +
                                  "Adam(1e-3, clipnorm 400), l2 1e-4, variational dropout 0.2",
                      "per_gpu_batch": n_sample, "global_batch": n_sample, "frames": T_FRAMES, "parallelism": "cpu",
                      "input_pipeline": "in line"},
           "cpu_baseline": {"value": val, "unit": "utt/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "utt/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from asr_study_b200._lib import lib
    from asr_study_b200.core import models
    from asr_study_b200.datasets.dataset_generator import DatasetIterator
    from asr_study_b200.engine import pack_labels
    from asr_study_b200.preprocessing import audio

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    nb = args.batch
    gb = nb * world
    pcm_np, labels = synth_batch(nb, 1234 + 1000 * rank)
    pcm_host = torch.from_numpy(pcm_np.reshape(-1)).pin_memory()
    off_host = torch.arange(nb + 1, dtype=torch.int64) * pcm_np.shape[1]
    pcm_dev = pcm_host.to(dev)
    off_dev = off_host.to(dev)
    flat, loff, mx = pack_labels(labels, dev)
    if args.config == "c4":
        # BASELINE configs[3]: 40 log-mel -> DeepSpeech2-style 2 x Conv -> 5 x BiLSTM-800 -> Dense-28, 16 utterances per GPU
        feat = audio.LogFbank()
        model = models.deep_speech2(num_features=F, num_hiddens=H, num_layers=L, num_classes=C, dropout=args.dropout,
                                    weight_decay=1e-4, device=str(dev), seed=4321)
    else:
        feat = audio.MFCC(num_cep=13, d=True, dd=(F == 39))
        # the reference's own model factory and optimiser set-up (core/models.py:217-281, train.py:133-143)
        model = models.brsmv1(num_features=F, num_hiddens=H, num_layers=L, num_classes=C, dropout=args.dropout,
                              zoneout=args.zoneout, mi=[1.0, 0.5, 0.5] if args.mi else None, weight_decay=1e-4,
                              device=str(dev), seed=4321)
    model.compile(optimizer=models.Adam(lr=1e-3, clipnorm=400.0))
    eng = model.engine
    loss_host = torch.empty(nb, dtype=torch.float32).pin_memory()
    loss_bufs = [loss_host, torch.empty(nb, dtype=torch.float32).pin_memory()]
    loss_evs = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"k": 0, "last": None}

    def allreduce(g):                  # called by engine.backward: once on the whole gradient bucket (or per slice, --dp-mode slices)
        if world > 1:
            if args.dp_mode == "none":             # experiment: no collective at all (what the ranks do when left alone)
                return None
            return dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)

    eng.dp_slices = {"slices": True, "bucket": False}.get(args.dp_mode)      # auto / none: the engine decides by bucket size
    if world > 1:
        model.set_data_parallel(allreduce, world, rank=rank)

    # input pipeline: batch k+1's H2D copy + fused MFCC launch run on a side stream while batch k trains
    # (asr_study_b200/datasets/prefetch.py); every step still featurises its own batch inside the timed region
    from asr_study_b200.datasets.prefetch import DeviceFeaturePrefetcher
    pre = DeviceFeaturePrefetcher(feat, dev, nb, pcm_np.shape[1], T_FRAMES)

    def train(x, lens):
        return eng.train_step(x, lens, flat, loff, mx, global_batch=gb, allreduce=allreduce, lr=1e-3, clipnorm=400.0)

    def step(pcm):
        if not args.prefetch:
            x, lens = feat.batch(pcm, off_dev, t_max=T_FRAMES, time_major=True)
            return train(x, lens)
        x, lens = pre.get()                                        # features of THIS step (submitted one step ago)
        pre.submit(pcm, off_dev)                                   # the next step's features, concurrent with this step
        loss = train(x, lens)
        pre.release()
        return loss

    def step_e2e():
        if args.prefetch:
            loss = step(pcm_host)                                  # H2D of the next batch's audio inside the timed region
        else:
            loss = step(pcm_host.to(dev, non_blocking=True))       # H2D of this step's audio, inside the timed region
        # D2H of the step's result, every step; the host reads it one step late (while the next step runs) so that
        # enqueueing step k+1 does not wait for step k to drain — the usual logging lag of a training loop
        k = e2e_state["k"]
        e2e_state["k"] = k + 1
        loss_bufs[k & 1].copy_(loss, non_blocking=True)
        loss_evs[k & 1].record()
        if k > 0:
            loss_evs[(k - 1) & 1].synchronize()
            e2e_state["last"] = float(loss_bufs[(k - 1) & 1].mean())
        return loss_bufs[k & 1]

    # ---- e2e through the plugin surface: DatasetIterator (host pcm -> host feature batch, one fused K1 launch per batch,
    # on a generator worker thread like Keras' fit_generator, max_q_size 10) -> CTCModel.train_on_batch (host batch in,
    # [loss, ctc_loss, decoder_loss, decoder_ler] out as Python floats: one device->host read per step)
    from asr_study_b200.core.models import _GeneratorFeed
    flow = DatasetIterator([pcm_np[i] for i in range(nb)], [np.asarray(l, np.int32) for l in labels], batch_size=nb,
                           shuffle=False, input_parser=feat, label_parser=None, rank=0, world_size=1)
    plugin = {"feed": None, "last": None, "read": None}

    def plugin_start(total_steps):
        plugin["feed"] = _GeneratorFeed(flow, total_steps * nb, 10, 1, dev)

    def step_plugin():
        x, _y = plugin["feed"].get()
        m = model.train_on_batch(x)                     # host batch in; metrics come back as a lazily read float sequence
        if plugin["last"] is not None:
            plugin["read"] = plugin["last"].result()    # the host reads every step's metrics, one step late (a logging loop)
        plugin["last"] = m

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    rank_ms = {}

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            every = [torch.zeros_like(ms) for _ in range(world)]
            dist.all_gather(every, ms)
            rank_ms["last"] = [round(float(v.item()) / steps, 4) for v in every]
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    sampler = ClockSampler(local)
    sampler.prepare()
    if args.prefetch:
        pre.submit(pcm_dev, off_dev)                               # prime the pipeline (untimed)
    for _ in range(max(args.warmup, 3)):
        step(pcm_dev)
    torch.cuda.synchronize()
    if eng.lstm_status() != 0:
        raise RuntimeError("persistent LSTM kernel watchdog fired during warm-up")
    sampler.start()
    l0 = lib.asr_launch_count()
    ms = timed(lambda: step(pcm_dev), args.steps)
    rank_ms_value = rank_ms.get("last")
    launches = (lib.asr_launch_count() - l0) // max(args.steps, 1)
    clocks = sampler.stop()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    plugin_start(args.steps + 2)
    for _ in range(2):
        step_plugin()
    ms_plugin = timed(step_plugin, args.steps)
    plugin["feed"].close()
    model.check_status()
    value = gb * args.steps / (ms / 1e3)
    e2e_engine = gb * args.steps / (ms_e2e / 1e3)
    e2e = gb * args.steps / (ms_plugin / 1e3)
    feat_bytes = int(nb * T_FRAMES * F * 4)
    dp_check = dp_equivalence(model, feat, nb, world, rank, dev, torch, dist) if (world > 1 and args.dp_mode != "none") else None

    # per-kernel-class device time (instrumented pass, outside the timed region)
    kern = kernel_breakdown(eng, feat, pcm_dev, off_dev, flat, loff, mx, gb, torch) if rank == 0 else {}
    if rank == 0 and args.config == "c4":
        kern["note"] = "other_ms holds the conv front end's layout kernels (pack, weight expansion, activation, unfold-transpose) beside casts and masks; its GEMMs are in gemm_ms"
    pk = peaks()
    out = None
    if rank == 0:
        step_tf = value * GFLOP_TRAIN_PER_UTT / 1e3 / world                       # TFLOP/s per GPU
        t_rec = eng.out_frames(T_FRAMES)                                          # frames the recurrences walk (conv front: 500)
        gf = 2 * 2.0 * t_rec * nb * H * 4 * H / 1e9                              # recurrent matmul of ONE launch (2 dirs)
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic_r2.json")
        if not os.path.exists(tpath):
            tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath))
        # dominant kernel = the backward recurrence (largest share of the step, profiles/ncu_summary_*.md)
        ms_bwd = kern.get("lstm_bwd_ms", 0.0) / L
        ms_fwd = kern.get("lstm_fwd_ms", 0.0) / L
        c4 = args.config == "c4"
        if c4:
            traffic = None              # the captures under profiles/ are of the C2 kernels
        roof = {"bound": "tensor", "kernel": ("lstmtc2::bwd3_kernel<832> (H = 800 zero-padded; persistent BiLSTM BPTT, one launch per layer)" if c4 else
                                              "lstmtc4::bwd_kernel (persistent BiLSTM BPTT, one launch per layer)"),
                "achieved": gf / ms_bwd if ms_bwd else None, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                "frac": (gf / ms_bwd / pk["tf_burst"]) if ms_bwd else None,
                "peak_source": pk["src"] + " bf16_tflops (burst; kernel timed alone with CUDA events)",
                "flop_per_launch": gf * 1e9, "ms_per_launch": ms_bwd,
                "traffic": (traffic or {}).get("lstm_bwd_bytes_per_launch"),
                "mma_precision": "bf16 operands, fp32 accumulate (TS-mode tcgen05.mma, U tile resident in TMEM)",
                "note": "latency-bound by design at N=32: 999 dependent steps per launch; see profiles/lstm_phases_r2.md",
                "others": {
                    "lstm_fwd": {"achieved": gf / ms_fwd if ms_fwd else None, "ms_per_launch": ms_fwd,
                                 "frac": (gf / ms_fwd / pk["tf_burst"]) if ms_fwd else None,
                                 "traffic": (traffic or {}).get("lstm_fwd_bytes_per_launch")},
                    "gemm_all": {"achieved": (2.0 / 3.0) * gb / world * GFLOP_TRAIN_PER_UTT / kern["gemm_ms"] if kern.get("gemm_ms") else None,
                                 "ms_per_step": kern.get("gemm_ms"), "unit": "TFLOP/s"},
                    "whole_step_vs_lstm_gemm_roofline": {"achieved": step_tf, "peak": pk["tf_sust"],
                                                         "frac": step_tf / pk["tf_sust"],
                                                         "peak_source": pk["src"] + " bf16_tflops_sustained"}}}
        out = {"metric": METRIC if args.config != "c4" else "utterances/sec (10 s, 40 log-mel, 2xConv + 5xBiLSTM-800, CTC)", "value": value, "unit": "utt/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": ("fp16/bf16 tensor-core operands, fp32 saved activations" if args.fp32_storage else "fp16/bf16 tensor-core operands and saved activations") + ", fp32 accumulate / cell state / parameters",
               "data": "synthetic",
               "config": {"workload": ("C4 (BASELINE configs[3]): synthetic 16 kHz 10 s clips, 40 log-mel, DS2-style 2xConv (32 ch, 41x11 s(2,2) + "
                                       "21x11 s(2,1), clipped ReLU 20, no batch norm) -> 320 features x 500 frames, " if args.config == "c4" else "") +
                                      ("C2" if (F, H, L) == (26, 512, 3) else "custom (not the BASELINE headline config)") +
                                      ": synthetic 16 kHz 10 s clips, %d %s, %dxBiLSTM-%d, Dense-28, CTC, " % (F, "log-mel" if args.config == "c4" else "MFCC", L, H) +
                                      "Adam(1e-3, clipnorm 400), l2 1e-4, variational dropout %g" % args.dropout
                                      + (", zoneout %g" % args.zoneout if args.zoneout else "") + (", MI" if args.mi else ""), "per_gpu_batch": nb,
                          "global_batch": gb, "frames": T_FRAMES, "parallelism": f"dp{world}" + ("" if args.dp_mode == "auto" else f" (--dp-mode {args.dp_mode})"),
                          "l2_flush": "per-step working set ~4 GB >> 126 MB L2 (inputs larger than L2)",
                          "input_pipeline": "prefetch: H2D + MFCC of batch k+1 on a side stream during step k" if args.prefetch
                          else "in line"},
               "e2e": {"value": e2e, "unit": "utt/s", "h2d_bytes_per_step": int(pcm_host.numel() * 4) + feat_bytes,
                       "d2h_bytes_per_step": feat_bytes + 16, "ms_per_step": ms_plugin / args.steps,
                       "path": "DatasetIterator (host pcm -> K1 -> host [N,T,F] batch, generator thread, queue 10) -> "
                               "CTCModel.train_on_batch (host batch -> device -> step -> 4 metrics copied back every step, read by the host "
                               "one step late)",
                       "last_metrics": plugin["last"].result() if plugin["last"] is not None else None},
               "e2e_engine": {"value": e2e_engine, "unit": "utt/s", "h2d_bytes_per_step": int(pcm_host.numel() * 4),
                              "d2h_bytes_per_step": int(loss_host.numel() * 4), "ms_per_step": ms_e2e / args.steps,
                              "path": "pinned host pcm -> H2D + K1 one step ahead on a side stream -> engine.train_step -> "
                                      "per-utterance loss D2H, read one step late"},
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "kernel_ms": kern,
               **({"data_parallel_check": dp_check} if dp_check is not None else {}),
               **({"rank_ms_per_step": rank_ms_value} if rank_ms_value else {}),
               "final_loss_mean": float(loss_bufs[(e2e_state["k"] - 1) & 1].mean())}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if args.cpu_baseline and world == 1:            # the CPU port beside the GPU number: rank 0 at N = 1 only
            if args.infer:                               # BASELINE configs[4] inside the default line (driver-visible)
                del model, eng
                torch.cuda.empty_cache()
                out["infer"] = run_infer(args, quiet=True)
            v, dt = cpu_arm(CPU_SAMPLE, 1, 0)
            out["cpu_baseline"] = {"value": v, "unit": "utt/s", "cores": os.cpu_count(), "kind": "port",
                                   "sample": cpu_sample_text(CPU_SAMPLE, dt)}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def dp_equivalence(model, feat, nb, world, rank, dev, torch, dist):
    """Two checks of the data-parallel path on the hardware (outside the timed region, reported in the N > 1 line):
    replica_checksum_spread  max over ranks of |sum(params_rank) - sum(params_0)| after the timed steps — the replicas
                             apply the same reduced gradient with the same kernels and must stay bit-identical: 0.0;
    dp_vs_big_batch_grad     the gradient of ONE global batch computed data-parallel (each rank its 32 utterances, the
                             bucket all-reduced in per-layer slices) against the same global batch computed by rank 0
                             alone as one big batch (dropout off on both sides): max |d| / max |ref| over the bucket."""
    import dataclasses

    from asr_study_b200.engine import AcousticEngine, pack_labels
    eng = model.engine
    sums = torch.stack([eng.params.flat.double().sum(), eng.params.flat.double().abs().sum()])
    gathered = [torch.zeros_like(sums) for _ in range(world)]
    dist.all_gather(gathered, sums)
    spread = max(float((g - gathered[0]).abs().max()) for g in gathered)
    spec0 = dataclasses.replace(eng.user_spec, dropout=0.0)
    e2 = AcousticEngine(spec0, device=dev, init_params=eng.params.export("flat"))
    gb = nb * world

    def grads(pcm_np, labels, allreduce):
        pcm = torch.from_numpy(pcm_np.reshape(-1)).to(dev)
        off = (torch.arange(pcm_np.shape[0] + 1, dtype=torch.int64) * pcm_np.shape[1]).to(dev)
        x, lens = feat.batch(pcm, off, t_max=T_FRAMES, time_major=True)
        flat, loff, mx = pack_labels(labels, dev)
        logits = e2.forward(x, training=True)
        _, dl = e2.ctc(logits, lens, flat, loff, mx, grad_scale=1.0 / gb)
        for h in e2.backward(dl, allreduce=allreduce):
            if h is not None and hasattr(h, "wait"):
                h.wait()
        torch.cuda.synchronize()
        return e2.params.grad.clone()

    pcm_r, lab_r = synth_batch(nb, 1234 + 1000 * rank)
    g_dp = grads(pcm_r, lab_r, lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True))
    err = None
    if rank == 0:
        parts = [synth_batch(nb, 1234 + 1000 * r) for r in range(world)]
        g_big = grads(np.concatenate([p[0] for p in parts]), sum((p[1] for p in parts), []), None)
        err = float((g_dp - g_big).abs().max() / g_big.abs().max())
    dist.barrier()
    return {"replica_checksum_spread": spread, "dp_vs_big_batch_grad": err,
            "note": "spread must be 0.0; the gradient difference is summation order + 16-bit operand rounding of different tilings"}


def kernel_breakdown(eng, feat, pcm_dev, off_dev, flat, loff, mx, gb, torch):
    """CUDA-event time per kernel class over one step (separate pass; explains `value`)."""
    res = {}

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def span(name, fn):
        a, b = ev(), ev()
        torch.cuda.synchronize()
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = res.get(name, 0.0) + a.elapsed_time(b)
        return r

    feat.batch(pcm_dev, off_dev, t_max=T_FRAMES, time_major=True)          # untimed: this thread's workspace / stream set-up
    x, lens = span("mfcc_ms", lambda: feat.batch(pcm_dev, off_dev, t_max=T_FRAMES, time_major=True))
    import asr_study_b200.engine as E
    keys = {"asr_lstm_forward": "lstm_fwd_ms", "asr_lstm_backward": "lstm_bwd_ms", "asr_gemm_tn": "gemm_ms",
            "asr_gemm_tn_ex": "gemm_ms",
            "asr_ctc_loss_grad": "ctc_ms", "asr_adam_step": "adam_ms", "asr_grad_sqnorm": "adam_ms"}

    class Tap:
        def __init__(self, real):
            self.real = real

        def __getattr__(self, n):
            f = getattr(self.real, n)
            if not n.startswith("asr_") or n.endswith("_bytes") or n == "asr_launch_count":
                return f
            return lambda *a: span(keys.get(n, "other_ms"), lambda: f(*a))

    real = E.lib
    E.lib = Tap(real)
    try:
        eng.train_step(x, lens, flat, loff, mx, global_batch=gb, lr=1e-3, clipnorm=400.0)
    finally:
        E.lib = real
    return {k: round(v, 4) for k, v in res.items()}


def run_infer(args, quiet=False):
    """configs[4] (C5): MFCC -> 3xBiLSTM-512 forward -> CTC beam search (width 100) over synthetic 10 s clips on one
    B200; prints clips/s and checks LER parity of the device decode against the oracle's TF-order beam search on a
    sample of the same logits (identical label sequences => identical LER)."""
    import torch

    from asr_study_b200._lib import lib
    from asr_study_b200.engine import AcousticEngine, ModelSpec
    from asr_study_b200.preprocessing import audio
    from oracle import ctc as oc
    from oracle import ctc_beam_c as occ                 # the C restatement of oc.beam_decode (tests hold it to the Python one)
    from oracle.model import synth_clip, synth_labels

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    nb, total, W = 64, args.clips, args.beam_width
    feat = audio.MFCC(num_cep=13, d=True, dd=False)
    eng = AcousticEngine(ModelSpec(F, H, L, C), device=dev, seed=4321, fp16_storage=not args.fp32_storage,
                         overlap=not args.no_engine_overlap)
    # a random-init network emits near-uniform posteriors (logits within +-0.1): every beam is a near tie and fp32
    # totals collide exactly, so the result hangs on tie-breaking order.  Scale the Dense kernel so the posteriors are
    # as peaky as a trained CTC model's (the regime config 5 is about); --sharpen sets the factor.
    eng.params.p("dense.W").mul_(args.sharpen)
    eng._prepared_for = None                              # the masters were edited in place: re-derive the 16-bit operands
    # distinct synthetic clips: enough forward batches that the LER-parity sample holds no clip twice
    n_distinct = max(1, (min(args.ler_sample, args.clips) + nb - 1) // nb)
    pcm_nps = [np.stack([synth_clip(777, j * nb + i, SECONDS, FS) for i in range(nb)]) for j in range(n_distinct)]
    pcm_hosts = [torch.from_numpy(p.reshape(-1)).pin_memory() for p in pcm_nps]
    pcm_np, pcm_host = pcm_nps[0], pcm_hosts[0]
    off = (torch.arange(nb + 1, dtype=torch.int64) * pcm_np.shape[1]).to(dev)
    truth = synth_labels(778, nb * n_distinct, 50)
    fwd_count = {"i": 0}

    def next_pcm():
        fwd_count["i"] += 1
        return pcm_hosts[(fwd_count["i"] - 1) % n_distinct]

    G = max(1, args.decode_group)
    nb_fwd = nb
    # --overlap_decode: the beam search of group k (one warp per utterance, ~17 KB of shared memory each, latency-bound
    # for ~80 ms per launch) runs on a second stream under the forward passes of group k+1.  Its CTAs spread over all
    # SMs, so the forward kernels must be able to share an SM with them: the recurrences give up their exclusive
    # shared-memory reservation and the GEMMs use the non-persistent 96 KB tiling (the persistent one needs 192 KB).
    overlap = bool(args.overlap_decode)
    eng.shared_sm = overlap          # ASR_LSTM_SHARED_SM / ASR_GEMM_TILE128 flags of the C ABI, per launch
    bigs = [torch.empty(T_FRAMES, G * nb_fwd, C, dtype=torch.float32, device=dev) for _ in range(2)]
    big_lens = [torch.empty(G * nb_fwd, dtype=torch.int32, device=dev) for _ in range(2)]
    outs_h = [torch.empty(G * nb_fwd, T_FRAMES, dtype=torch.int32).pin_memory() for _ in range(2)]
    lens_h = [torch.empty(G * nb_fwd, dtype=torch.int32).pin_memory() for _ in range(2)]
    main = torch.cuda.current_stream()
    dec = torch.cuda.Stream(device=dev) if overlap else main
    ev_fwd = [torch.cuda.Event() for _ in range(2)]
    ev_dec = [torch.cuda.Event() for _ in range(2)]
    state = {"k": 0}
    pre = None
    if args.prefetch:                                     # same device input pipeline as the training bench
        from asr_study_b200.datasets.prefetch import DeviceFeaturePrefetcher
        pre = DeviceFeaturePrefetcher(feat, dev, nb_fwd, pcm_np.shape[1], T_FRAMES)
        pre.submit(next_pcm(), off)                       # primed (untimed); every timed batch still copies + featurises one

    def batch():
        """G forward batches (host pcm -> H2D -> MFCC -> BiLSTM) then one beam-search launch over all G * nb utterances
        and the D2H copy of its labels; with overlap the search + copy of this group run on `dec` while the caller goes
        on to the next group's forward passes (double-buffered logits / outputs)."""
        k = state["k"]
        state["k"] = k + 1
        p = k & 1
        big, big_len = bigs[p], big_lens[p]
        if k >= 2:
            ev_dec[p].synchronize()                       # host: the pinned outputs of group k-2 are complete (and consumed)
            main.wait_event(ev_dec[p])                    # device: its search no longer reads big[p]
        for g in range(G):
            if pre is not None:                           # H2D + MFCC of the NEXT forward batch on the prefetcher's stream
                x, lens = pre.get()
                pre.submit(next_pcm(), off)
            else:
                pcm = next_pcm().to(dev, non_blocking=True)
                x, lens = feat.batch(pcm, off, t_max=T_FRAMES, time_major=True)
            logits = eng.forward(x, training=False)
            big[:, g * nb_fwd:(g + 1) * nb_fwd].copy_(logits)
            big_len[g * nb_fwd:(g + 1) * nb_fwd].copy_(lens)
            if pre is not None:
                pre.release()
        ev_fwd[p].record(main)
        dec.wait_event(ev_fwd[p])
        with torch.cuda.stream(dec):
            out, out_len = eng.beam(big, big_len, W, True, tag=str(p))
            outs_h[p].copy_(out, non_blocking=True)
            lens_h[p].copy_(out_len, non_blocking=True)
            ev_dec[p].record(dec)
        if not overlap:
            ev_dec[p].synchronize()
        return p

    def drain():
        for e in ev_dec:
            e.synchronize()
        torch.cuda.synchronize()

    for _ in range(2):
        batch()
    drain()
    nb = G * nb_fwd
    nbatches = max(1, total // nb)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.asr_launch_count()
    e0.record(main)
    for _ in range(nbatches):
        p = batch()
    main.wait_stream(dec)                                 # the last search and its D2H copy are inside the timed region
    e1.record(main)
    drain()
    ms = e0.elapsed_time(e1)
    logits, lens, out, out_len = bigs[p], big_lens[p], outs_h[p], lens_h[p]
    launches = (lib.asr_launch_count() - l0) // nbatches
    # LER parity on a sample: the oracle's TF-order beam search on the SAME logits (the last decoded group)
    k = min(args.ler_sample, nb)
    lg = logits[:, :k].transpose(0, 1).contiguous().cpu().numpy()
    t_or = time.perf_counter()
    ref = occ.beam_decode(lg, [T_FRAMES] * k, beam_width=W)
    t_or = time.perf_counter() - t_or
    got = [[int(v) for v in out[i, :int(out_len[i])]] for i in range(k)]
    same = sum(int(a == b) for a, b in zip(got, ref))
    # utterance u of a decoded group is clip (first forward batch of the group + u // 64) % n_distinct, u % 64: G and the
    # number of distinct batches are both powers of two here, so groups start at distinct-batch 0
    assert (G * nb_fwd) % (n_distinct * nb_fwd) == 0 or n_distinct % G == 0
    truth = [truth[((u // nb_fwd) % n_distinct) * nb_fwd + u % nb_fwd] for u in range(k)]
    res = {"metric": "clips/sec (inference: MFCC -> 3xBiLSTM-512 fwd -> CTC beam search width %d)" % W,
           "value": nb * nbatches / (ms / 1e3), "unit": "clips/s", "n_gpus": 1, "clips": nb * nbatches,
           "ms_per_batch": ms / nbatches, "batch": nb, "forward_batch": nb_fwd, "higher_is_better": True, "data": "synthetic",
           "config": {"workload": "C5: synthetic 16 kHz 10 s clips, 26-MFCC, 3xBiLSTM-512, beam %d, host pcm -> labels" % W,
                      "input_pipeline": "prefetch: H2D + MFCC of forward batch k+1 on a side stream during batch k" if pre is not None else "in line",
                      "decode": ("beam search of group k on a second stream under the forward passes of group k+1" if overlap
                                 else "in line: forward passes, then the beam search, then the next group")},
           "gpu_launches": int(launches),
           "ler_parity": {"sample": k, "identical_label_sequences": same, "oracle": "oracle/ctc_beam.c (%.1f s)" % t_or,
                          "ler_device_vs_truth": oc.ler(truth[:k], got), "ler_oracle_vs_truth": oc.ler(truth[:k], ref),
                          "ler_rel_diff": abs(oc.ler(truth[:k], got) - oc.ler(truth[:k], ref)) / max(oc.ler(truth[:k], ref), 1e-12),
                          "note": "sequences differ only where fp32 beam totals tie exactly (random-init posteriors): TF breaks "
                                  "ties by heap order, which neither restatement can pin without TF; tests/test_gpu_beam.py "
                                  "has the identical-sequence cases"}}
    if not quiet:
        print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="utterances per GPU")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dropout", type=float, default=0.2, help="brsmv1 dropout_W = dropout_U (reference default 0.2)")
    ap.add_argument("--zoneout", type=float, default=0.0, help="brsmv1 zoneout switch (off in the headline config)")
    ap.add_argument("--mi", action="store_true", help="brsmv1 multiplicative-integration switch (off in the headline config)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-prefetch", dest="prefetch", action="store_false",
                    help="featurise each batch in line instead of one step ahead on the side stream")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="train = C2/C3 (default), infer = C5")
    ap.add_argument("--clips", type=int, default=10240, help="infer mode / sub-record: BASELINE configs[4] asks for 10 k clips")
    ap.add_argument("--dp-mode", dest="dp_mode", default="auto", choices=["auto", "bucket", "slices", "none"],
                    help="N > 1: 'auto' = the engine's rule (one all-reduce of the whole gradient bucket behind the backward pass "
                         "below 128 MB, per-layer slices overlapped with the BPTT of the layer below above), 'bucket' / 'slices' "
                         "force one of them, 'none' = no collective (experiment: the ranks diverge; timing only)")
    ap.add_argument("--no-infer", dest="infer", action="store_false", help="skip the configs[4] sub-record of the default line")
    ap.add_argument("--beam_width", type=int, default=100)
    ap.add_argument("--ler_sample", type=int, default=256)
    ap.add_argument("--no-engine-overlap", dest="no_engine_overlap", action="store_true",
                    help="A/B: keep the engine's operand preparation / gradient GEMMs on the main stream")
    ap.add_argument("--fp32-storage", dest="fp32_storage", action="store_true",
                    help="A/B: the fp32-storage recurrences (csrc/lstm_tc2.cu) instead of fp16 storage + TMA (csrc/lstm_tc4.cu)")
    ap.add_argument("--sharpen", type=float, default=60.0, help="infer mode: factor on the random-init Dense kernel")
    ap.add_argument("--decode_group", type=int, default=16,
                    help="infer mode: forward batches decoded by ONE beam-search launch (one warp per utterance: the "
                         "search is latency-bound, so more utterances per launch is nearly free)")
    ap.add_argument("--overlap_decode", type=int, default=1, help="infer mode: decode group k under the forward passes of group k+1")
    ap.add_argument("--hidden", type=int, default=512, help="BiLSTM width (BASELINE config: 512; brsmv1's own default: 256)")
    ap.add_argument("--layers", type=int, default=3, help="BiLSTM layers (BASELINE config: 3; brsmv1's own default: 5)")
    ap.add_argument("--dd", action="store_true", help="39-dim MFCC (13 + delta + delta-delta: brsmv1's own default) instead of 26")
    ap.add_argument("--config", default="c2", choices=["c2", "c4"],
                    help="c2 (default): BASELINE configs[1]/[2]; c4: configs[3], DS2-style 2xConv + 5xBiLSTM-800 on 40 log-mel, 16 utt/GPU")
    args = ap.parse_args()
    global F, H, L, GFLOP_TRAIN_PER_UTT
    F, H, L = (39 if args.dd else 26), args.hidden, args.layers
    # fwd + dX + dW of the LSTM GEMMs and the Dense layer (88.80 GFLOP/utt for the BASELINE config)
    GFLOP_TRAIN_PER_UTT = 3 * (sum(2 * T_FRAMES * 2 * ((F if l == 0 else 2 * H) + H) * 4 * H for l in range(L))
                               + 2 * T_FRAMES * 2 * H * C) / 1e9
    if args.config == "c4":
        F, H, L = 40, 800, 5
        args.batch = 16 if args.batch == 32 else args.batch
        args.cpu_baseline = args.infer = False
        t2, d0 = 500, 320                                  # frames / features behind the conv front end
        conv = 2 * (500 * 20 * 32 * 451) + 3 * (500 * 10 * 32 * 7392) * 1   # conv1 fwd + dW; conv2 fwd + dW + dX (MACs)
        GFLOP_TRAIN_PER_UTT = (3 * (sum(2 * t2 * 2 * ((d0 if l == 0 else 2 * H) + H) * 4 * H for l in range(L))
                                    + 2 * t2 * 2 * H * C) + 2 * conv) / 1e9
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "infer":
        run_infer(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
