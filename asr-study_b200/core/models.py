"""Model factories with the reference's surface (core/models.py:31-281) on the CUDA engine.

``graves2006``, ``eyben`` (BiLSTM part), ``brsmv1`` and ``ctc_model`` keep their names, keyword
arguments and defaults; they return a ``CTCModel`` that answers the Keras calls train.py /
eval.py / predict.py make on it (compile, fit_generator, evaluate_generator, predict,
train_on_batch, optimizer.lr, metrics_names, get_layer, save) — see SURVEY.md 8(b).
``maas`` / ``deep_speech`` are SimpleRNN + clipped-ReLU stacks that cannot even be constructed in
the reference (un-imported names, core/models.py:122,129): out of scope, they raise.
"""
from __future__ import annotations

import io
import json
import os
import time

import numpy as np
import torch

from .._lib import AsrError
from ..engine import AcousticEngine, ModelSpec, pack_labels
from .layers import (LSTM, Bidirectional, Dense, Dropout, GaussianNoise, Input, Tensor, TimeDistributed, _Merge,   # noqa: F401
                     l2, merge, recurrent)

_PAD = 16          # batch rows are padded to a multiple of 16 (tensor-core tile / 16-byte operand rows)


class Adam(object):
    """keras.optimizers.Adam as configured at train.py:137."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-8, clipnorm=0.):
        self.lr, self.beta_1, self.beta_2, self.epsilon, self.clipnorm = lr, beta_1, beta_2, epsilon, clipnorm
        self.kind = "adam"


class SGD(object):
    """keras.optimizers.SGD as configured at train.py:134-135."""

    def __init__(self, lr=0.01, momentum=0., clipnorm=0.):
        self.lr, self.momentum, self.clipnorm = lr, momentum, clipnorm
        self.kind = "sgd"


class _GeneratorFeed(object):
    """Keras-1 GeneratorEnqueuer for the one-worker case: a daemon thread pulls next(generator) into a bounded queue
    until `total_samples` utterances have been fetched (never more: the generator's position after the call is what a
    synchronous loop would leave).  nb_worker=0: no thread."""

    def __init__(self, generator, total_samples, max_q_size, nb_worker, device):
        import queue
        import threading
        self.gen, self.total, self.thread = generator, int(total_samples), None
        if not nb_worker or self.total <= 0:
            return
        self.q = queue.Queue(maxsize=max(1, int(max_q_size)))
        self.stop = threading.Event()

        def work():
            if torch.cuda.is_available() and torch.device(device).type == "cuda":
                torch.cuda.set_device(device)           # new threads start on device 0
            fetched = 0
            try:
                while fetched < self.total and not self.stop.is_set():
                    item = next(self.gen)
                    fetched += np.asarray(item[0][0]).shape[0]
                    while not self.stop.is_set():
                        try:
                            self.q.put(item, timeout=0.1)
                            break
                        except queue.Full:
                            pass
            except BaseException as e:                  # surfaces in get()
                self.q.put(e)

        self.thread = threading.Thread(target=work, daemon=True)
        self.thread.start()

    def get(self):
        if self.thread is None:
            return next(self.gen)
        item = self.q.get()
        if isinstance(item, BaseException):
            raise item
        return item

    def close(self):
        if self.thread is not None:
            self.stop.set()
            self.thread.join(timeout=5.0)


class LazyMetrics(object):
    """[loss, ctc_loss, decoder_loss, decoder_ler] of one batch, as Keras' train_on_batch returns them — a sequence of
    Python floats — whose device->host read-back is asynchronous: the four values are copied into pinned memory behind the
    step on the step's own stream, and the host only waits for them when an element is actually read.  A loop that logs
    the previous batch while the next one trains (what Keras' progress bar amounts to) therefore never drains the GPU."""

    def __init__(self, dev_tensor):
        self._host = torch.empty(dev_tensor.numel(), dtype=torch.float32).pin_memory()
        self._host.copy_(dev_tensor, non_blocking=True)
        self._ev = torch.cuda.Event()
        self._ev.record()
        self._vals = None

    def result(self):
        if self._vals is None:
            self._ev.synchronize()
            self._vals = [float(v) for v in self._host.tolist()]
        return self._vals

    def __len__(self):
        return int(self._host.numel())

    def __getitem__(self, i):
        return self.result()[i]

    def __iter__(self):
        return iter(self.result())

    def __repr__(self):
        return repr(self.result())


def _label_rows(labels):
    if labels is None:
        return None
    if hasattr(labels, "tocsr"):
        m = labels.tocsr()
        return [m.data[m.indptr[i]:m.indptr[i + 1]].astype(np.int32) for i in range(m.shape[0])]
    return [np.asarray(r, dtype=np.int32) for r in labels]


class CTCModel(object):
    """[inputs, labels, inputs_length] -> [ctc loss per utterance, greedy decode] (core/models.py:31-52)."""

    metrics_names = ["loss", "ctc_loss", "decoder_loss", "decoder_ler"]

    def __init__(self, spec: ModelSpec, device=None, seed=4321, is_greedy=True, beam_width=100,
                 merge_repeated=True, input_std_noise=0.0):
        if device is None:
            device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        self.spec, self.seed = spec, int(seed)
        self.engine = AcousticEngine(spec, device=device, seed=seed)
        self.device = self.engine.device
        self.optimizer = Adam()
        self.decoder = dict(is_greedy=is_greedy, beam_width=beam_width, merge_repeated=merge_repeated)
        self.input_std_noise = float(input_std_noise or 0.0)
        self.allreduce = None
        self.world_size = 1
        self.history = {}
        self._noise_offset = 0
        self._stage = {}

    # ---- Keras-facing plumbing --------------------------------------------------
    def compile(self, loss=None, optimizer=None, metrics=None, loss_weights=None, **kw):
        if optimizer is not None:
            self.optimizer = optimizer
        return self

    def get_layer(self, name=None, index=None):
        if name in ("inputs", "labels", "inputs_length", "decoder", "ctc", "beam_search"):
            return name
        raise ValueError("No such layer: %s" % name)

    def set_data_parallel(self, allreduce, world_size, rank=0):
        """allreduce(grad_slice) must SUM in place across ranks; it is called on slices that tile the flat gradient
        bucket exactly once per step, each as soon as it is complete, and may return a handle with .wait().  The
        parameters are initialised from the same seed on every rank; the dropout / zoneout / noise streams are offset
        by the rank so that the replicas draw different masks for their different utterances."""
        self.allreduce, self.world_size = allreduce, int(world_size)
        self.engine.set_rank(int(rank))

    def check_status(self):
        """Raises if a persistent recurrence kernel's watchdog fired (a peer CTA never delivered its slice): its outputs
        are then partial and every later step would train on garbage.  One 4-byte read-back; the training loop calls it
        once per epoch and the evaluator once per call, never per step."""
        st = self.engine.lstm_status()
        if st != 0:
            raise AsrError("persistent BiLSTM kernel aborted (status %d: 1 = watchdog timeout, 2 = TMEM allocation); "
                           "the step's gradients are not valid" % st)

    # ---- batches ------------------------------------------------------------------
    def _device_batch(self, x, x_len, labels, training):
        """host batch (the DatasetIterator contract: x f32 [N, Tmax, F] zero-padded, lengths, sparse labels) -> device
        operands of the engine: time-major features [T, Np, F] (Np = N padded to the 16-sample tile), lengths, packed
        labels.  One pinned staging buffer per shape, asynchronous copies, the batch-major -> time-major permutation done
        by the copy engine on the device (no host transpose): ~0.3 ms of host time per C2 batch instead of ~2.5 ms."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        N, T, F = x.shape
        Np = (N + _PAD - 1) // _PAD * _PAD
        key = (N, T, F)
        st = self._stage.get(key)
        if st is None:
            if len(self._stage) > 8:
                self._stage.clear()                         # real corpora: a new Tmax per batch; keep the cache small
            st = self._stage[key] = dict(host=torch.empty(N, T, F, dtype=torch.float32).pin_memory(),
                                         dev=torch.empty(N, T, F, dtype=torch.float32, device=self.device),
                                         xt=torch.zeros(T, Np, F, dtype=torch.float32, device=self.device),
                                         lens_h=torch.zeros(Np, dtype=torch.int32).pin_memory(),
                                         lens=torch.zeros(Np, dtype=torch.int32, device=self.device))
        if "ev" in st:
            st["ev"].synchronize()                           # the previous asynchronous copies out of the pinned buffers
        st["host"].numpy()[...] = x                          # pageable -> pinned (one memcpy), then async DMA
        st["lens_h"].zero_()
        st["lens_h"].numpy()[:N] = np.asarray(x_len).reshape(-1)[:N]
        st["dev"].copy_(st["host"], non_blocking=True)
        st["lens"].copy_(st["lens_h"], non_blocking=True)
        st.setdefault("ev", torch.cuda.Event())
        xt = st["xt"]
        xt[:, :N].copy_(st["dev"].permute(1, 0, 2))          # layout copy on the device; the padding columns stay zero
        if training and self.input_std_noise > 0:           # GaussianNoise(std), core/models.py:67,251 (train phase only)
            self._noise_offset = self.engine.add_gaussian_noise(xt, N, self.input_std_noise, self._noise_offset)
        rows = _label_rows(labels)
        packed = None
        if rows is not None:
            # packed sparse labels through the same pinned staging: a copy out of pageable memory would first wait for
            # everything queued on the stream (the previous step), i.e. drain the pipeline once per batch
            sizes = [len(r) for r in rows]
            total = int(sum(sizes))
            if st.get("lab_cap", -1) < max(total, 1):
                cap = max(1024, 2 * total)
                st.update(lab_cap=cap, lab_h=torch.zeros(cap, dtype=torch.int32).pin_memory(),
                          lab=torch.zeros(cap, dtype=torch.int32, device=self.device),
                          off_h=torch.zeros(Np + 1, dtype=torch.int32).pin_memory(),
                          off=torch.zeros(Np + 1, dtype=torch.int32, device=self.device))
            if total:
                st["lab_h"].numpy()[:total] = np.concatenate(rows)
            oh = st["off_h"].numpy()
            oh[0] = 0
            oh[1:N + 1] = np.cumsum(sizes)
            oh[N + 1:] = total                                 # the padding utterances carry no labels
            st["lab"].copy_(st["lab_h"], non_blocking=True)
            st["off"].copy_(st["off_h"], non_blocking=True)
            packed = (st["lab"][:max(total, 1)], st["off"], int(max(sizes) if sizes else 0))
        st["ev"].record()                                    # covers every copy out of the pinned buffers above
        return xt, st["lens"], packed, N

    def _decode(self, logits, lens, with_len=False):
        if self.decoder["is_greedy"]:
            out, out_len = self.engine.greedy(logits, lens, True)
        else:
            out, out_len = self.engine.beam(logits, lens, self.decoder["beam_width"], self.decoder["merge_repeated"])
        return (out, out_len) if with_len else out

    def _stats(self, loss, logits, lens, flat, off, mx, N):
        """[loss, ctc_loss, decoder_loss, decoder_ler] of one batch as a DEVICE tensor (no read-back): the decode and the
        label error rate (core/metrics.py:4-8, asr_edit_distance) run on the device, so a training loop only
        synchronises when it reads its logs."""
        out, out_len = self._decode(logits, lens, with_len=True)
        ler = self.engine.ler(out, out_len, flat, off, mx)[:N].mean()
        ctc = loss[:N].mean()
        return torch.stack([ctc + self._reg_dev(), ctc, torch.zeros_like(ctc), ler])

    # ---- train / eval / predict ------------------------------------------------------
    def train_on_batch(self, x, y=None):
        """host batch in, the four Keras metrics out (a float sequence, read back lazily: see LazyMetrics)"""
        return LazyMetrics(self._train_stats(x))

    def _train_stats(self, x):
        feats, labels, x_len = x[0], x[1], x[2]
        xt, lens, (flat, off, mx), N = self._device_batch(feats, x_len, labels, True)
        gb = N * self.world_size
        o = self.optimizer
        kw = dict(lr=o.lr, clipnorm=o.clipnorm, opt=o.kind)
        if o.kind == "adam":
            kw.update(beta1=o.beta_1, beta2=o.beta_2, eps=o.epsilon)
        else:
            kw.update(momentum=o.momentum)
        loss = self.engine.train_step(xt, lens, flat, off, mx, global_batch=gb, allreduce=self.allreduce, **kw)
        return self._stats(loss, self.engine.last_logits, self.engine.out_lengths(lens), flat, off, mx, N)

    def test_on_batch(self, x, y=None):
        feats, labels, x_len = x[0], x[1], x[2]
        xt, lens, (flat, off, mx), N = self._device_batch(feats, x_len, labels, False)
        logits = self.engine.forward(xt, training=False)
        lens = self.engine.out_lengths(lens)
        loss, _ = self.engine.ctc(logits, lens, flat, off, mx, want_grad=False)
        return [float(v) for v in self._stats(loss, logits, lens, flat, off, mx, N).tolist()]

    def predict(self, x, batch_size=None, verbose=0):
        """x = [inputs, inputs_length] (predict mode, utils/core_utils.py:76-93) -> -1-padded label matrix."""
        feats, x_len = x[0], x[1]
        xt, lens, _, N = self._device_batch(feats, x_len, None, False)
        logits = self.engine.forward(xt, training=False)
        return self._decode(logits, self.engine.out_lengths(lens))[:N].cpu().numpy()

    predict_on_batch = predict

    def logits(self, feats, x_len):
        """batch-major logits [N,T,C] (what the reference's TimeDistributed(Dense) emits)."""
        xt, lens, _, N = self._device_batch(feats, x_len, None, False)
        return self.engine.forward(xt, training=False)[:, :N].transpose(0, 1).contiguous()

    def _reg_dev(self):
        """sum of the l2(weight_decay) regularisers (core/models.py:263-264,279) as a device scalar — a reported metric,
        not part of the step (the optimiser folds the l2 gradient in itself)."""
        return self.engine.l2_penalty()

    def _reg(self):
        return float(self._reg_dev().item())

    def fit_generator(self, generator, samples_per_epoch, nb_epoch, validation_data=None, nb_val_samples=None,
                      max_q_size=10, nb_worker=1, callbacks=None, verbose=1, initial_epoch=0, **kw):
        callbacks = callbacks or []
        for cb in callbacks:
            if hasattr(cb, "set_model"):
                cb.set_model(self)
        # Keras runs the generator on a worker thread behind a queue of max_q_size batches (train.py:213-217:
        # max_q_size=10, nb_worker=1), so featurisation overlaps the training step; nb_worker=0 pulls in line.  The
        # worker fetches exactly the batches this call consumes.  Per batch nothing is read back: the metrics are
        # accumulated on the device and read once per epoch.
        feed = _GeneratorFeed(generator, max(0, nb_epoch - initial_epoch) * samples_per_epoch, max_q_size,
                              nb_worker, self.device)
        try:
            for epoch in range(initial_epoch, nb_epoch):
                seen, t0 = 0, time.time()
                agg_dev = torch.zeros(4, dtype=torch.float32, device=self.device)
                while seen < samples_per_epoch:
                    x, y = feed.get()
                    n = np.asarray(x[0]).shape[0]
                    agg_dev += self._train_stats(x) * n
                    seen += n
                agg = agg_dev.cpu().numpy().astype(np.float64)
                self.check_status()
                t_train = time.time() - t0
                logs = dict(zip(["loss", "ctc_loss", "decoder_loss", "decoder_ler"], agg / max(seen, 1)))
                if validation_data is not None and nb_val_samples:
                    v = self.evaluate_generator(validation_data, nb_val_samples)
                    logs.update({"val_" + k: val for k, val in zip(self.metrics_names, v)})
                for k, v in logs.items():
                    self.history.setdefault(k, []).append(float(v))
                if verbose:
                    show = {k: round(float(logs[k]), 4) for k in ("loss", "decoder_ler", "val_loss", "val_decoder_ler")
                            if k in logs}
                    print("Epoch %d/%d - %.1fs - %s - %.1f utt/s" % (epoch + 1, nb_epoch, time.time() - t0, show,
                                                                      seen / max(t_train, 1e-9)))
                for cb in callbacks:
                    if hasattr(cb, "on_epoch_end"):
                        cb.on_epoch_end(epoch, logs)
        finally:
            feed.close()                                    # also on an exception: the worker must not outlive the call
        return self.history

    def evaluate_generator(self, generator, val_samples, max_q_size=10, nb_worker=1, decode_group=None, **kw):
        """Same result as test_on_batch per batch, weighted by batch size (what Keras' evaluate_generator returns), as a
        device pipeline: the forward pass and the CTC loss of every batch run on the current stream and nothing is read
        back per batch; the logits of `decode_group` batches are decoded by ONE launch on a second stream, under the
        next group's forward passes (the beam search is one warp per utterance and latency-bound: a launch over 1 024
        utterances costs what a launch over 64 does), followed by the label-error-rate kernel; one read-back at the
        end.  The generator runs on a worker thread like fit_generator's.  decode_group=1 is the batch-by-batch order."""
        eng, dev = self.engine, self.device
        beam = not self.decoder["is_greedy"]
        G = int(decode_group or (16 if beam else 4))
        main = torch.cuda.current_stream(dev)
        dec = torch.cuda.Stream(device=dev) if G > 1 else main
        # the search CTAs spread over all SMs: let the forward kernels share an SM with them for the duration of this
        # call (ASR_LSTM_SHARED_SM / ASR_GEMM_TILE128 flags of the C ABI, passed per launch by the engine)
        shared_before = eng.shared_sm
        eng.shared_sm = bool(beam and G > 1)
        C = self.spec.num_classes
        losses, seen = [], 0
        ler_sum = torch.zeros((), dtype=torch.float32, device=dev)
        done_ev = [None, None]
        feed = _GeneratorFeed(generator, val_samples, max_q_size, nb_worker, dev)
        try:
            k = 0
            while seen < val_samples:
                batches = []
                while seen < val_samples and len(batches) < G:
                    x, y = feed.get()
                    batches.append(x)
                    seen += np.asarray(x[0]).shape[0]
                p = k & 1
                k += 1
                if done_ev[p] is not None:                 # the search two groups back no longer reads this buffer pair
                    main.wait_event(done_ev[p])
                shapes = [np.asarray(x[0]).shape for x in batches]
                Tmax = max(eng.out_frames(int(sh[1])) for sh in shapes)
                Ntot = sum((int(sh[0]) + _PAD - 1) // _PAD * _PAD for sh in shapes)
                big = eng._buf("eval_logits%d" % p, (Tmax, Ntot, C), torch.float32)
                big_len = eng._buf("eval_len%d" % p, (Ntot,), torch.int32)
                rows, real, c0 = [], [], 0
                for x in batches:
                    # staged one at a time: batches of one shape share their pinned / device staging buffers
                    xt, lens, (flat, off, mx), N = self._device_batch(x[0], x[2], x[1], False)
                    logits = eng.forward(xt, training=False)
                    lens = eng.out_lengths(lens)
                    loss, _ = eng.ctc(logits, lens, flat, off, mx, want_grad=False)
                    losses.append(loss[:N].clone())
                    T, Np = int(xt.shape[0]), int(xt.shape[1])
                    big[:T, c0:c0 + Np].copy_(logits)
                    big_len[c0:c0 + Np].copy_(lens)
                    rows += _label_rows(x[1]) + [np.zeros(0, np.int32)] * (Np - N)
                    real += list(range(c0, c0 + N))
                    c0 += Np
                gflat, goff, gmx = pack_labels(rows, dev)  # the group's labels, padding columns empty
                real_idx = torch.as_tensor(real, dtype=torch.int64, device=dev)
                ev_f = torch.cuda.Event()
                ev_f.record(main)
                dec.wait_event(ev_f)
                with torch.cuda.stream(dec):
                    if beam:
                        out, out_len = eng.beam(big, big_len, self.decoder["beam_width"], self.decoder["merge_repeated"], tag=str(p))
                    else:
                        out, out_len = eng.greedy(big, big_len, True)
                    ler_sum += eng.ler(out, out_len, gflat, goff, gmx)[real_idx].sum()
                    for t in (gflat, goff, real_idx):      # allocated on the caller's stream, consumed on `dec`
                        t.record_stream(dec)
                    done_ev[p] = torch.cuda.Event()
                    done_ev[p].record(dec)
            main.wait_stream(dec)
        finally:
            feed.close()
            eng.shared_sm = shared_before
        self.check_status()
        n_real = sum(int(l.numel()) for l in losses)
        ctc = float(torch.cat(losses).sum().item()) / max(n_real, 1)
        return [ctc + self._reg(), ctc, 0.0, float(ler_sum.item()) / max(n_real, 1)]

    # ---- checkpoint: weights + optimiser state + meta (core/callbacks.py:36-56, utils/core_utils.py:49-131) ----------
    # The reference writes Keras' HDF5 layout plus a `meta` group; h5py is not installable here, so the same content goes
    # into one .npz archive: arrays `p/<name>`, `m/<name>`, `v/<name>` and a JSON document `config` (spec, optimiser,
    # decoder, step, seeds / mask offsets, input_std_noise, meta).  Nothing in it is executable (no pickle).
    def save(self, path, meta=None):
        P = self.engine.params
        arrays = {}
        for which, tag in (("flat", "p"), ("m", "m"), ("v", "v")):
            for k, v in P.export(which).items():
                arrays["%s/%s" % (tag, k)] = v
        spec = dict(self.spec.__dict__)
        for k in ("layer_norm", "mi", "layer_hiddens"):
            spec[k] = list(spec[k]) if spec[k] is not None else None
        cfg = dict(format="asr_b200_ckpt/2", spec=spec, optimizer=dict(self.optimizer.__dict__), decoder=self.decoder,
                   step=int(self.engine.step_count), seed=self.seed, mask_offset=int(self.engine._mask_offset),
                   noise_offset=int(self._noise_offset), input_std_noise=self.input_std_noise, meta=meta or {})
        arrays["config"] = np.frombuffer(json.dumps(cfg, default=_jsonable).encode("utf-8"), dtype=np.uint8)
        buf = io.BytesIO()
        np.savez(buf, **arrays)
        tmp = path + ".tmp"
        with open(tmp, "wb") as f:
            f.write(buf.getvalue())
        os.replace(tmp, path)

    @classmethod
    def load(cls, path, device=None, mode="train", **kw):
        """mode (utils/core_utils.py:49-59): 'train' keeps the saved decoder; 'eval' / 'predict' switch to the beam search
        (width 400 unless overridden: utils/core_utils.py:67-72)."""
        if mode not in ("train", "predict", "eval"):
            raise ValueError("mode must be one of (train, predict, eval)")
        with np.load(path, allow_pickle=False) as z:
            cfg = json.loads(bytes(z["config"]).decode("utf-8"))
            groups = {t: {k[2:]: z[k] for k in z.files if k.startswith(t + "/")} for t in ("p", "m", "v")}
        spec = dict(cfg["spec"])
        for k in ("layer_norm", "mi", "layer_hiddens"):
            spec[k] = tuple(spec[k]) if spec.get(k) is not None else None
        m = cls(ModelSpec(**spec), device=device, seed=cfg.get("seed", 4321), input_std_noise=cfg.get("input_std_noise", 0.0), **kw)
        P = m.engine.params
        P.load(groups["p"])
        P.load(groups["m"], which="m")
        P.load(groups["v"], which="v")
        m.engine.step_count = int(cfg["step"])
        m.engine._mask_offset = int(cfg.get("mask_offset", 0))
        m._noise_offset = int(cfg.get("noise_offset", 0))
        o = cfg["optimizer"]
        m.optimizer = Adam(o["lr"], o["beta_1"], o["beta_2"], o["epsilon"], o["clipnorm"]) if o["kind"] == "adam" \
            else SGD(o["lr"], o["momentum"], o["clipnorm"])
        m.decoder = dict(cfg["decoder"])
        if mode in ("eval", "predict"):
            m.decoder.update(is_greedy=kw.get("is_greedy", False), beam_width=kw.get("beam_width", 400))
        return m, cfg.get("meta", {})


def _jsonable(o):
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, np.ndarray):
        return o.tolist()
    return str(o)


# -------------------------------------------------------------------------------------------
# factories
# -------------------------------------------------------------------------------------------
def _lower(inputs, output):
    """Walk the record graph from `output` back to `inputs` and return the ModelSpec fields of the chain
        Input -> [GaussianNoise] -> [TimeDistributed(Dense(P))] -> [Dropout] ->
        L x ( Bidirectional(LSTM) [-> merge([., previous], 'sum')] ) -> TimeDistributed(Dense(C))
    which is every topology of core/models.py on the BiLSTM hot path (graves2006 :55-73, eyben :76-103, brsmv1 :217-281)
    and the README recipe."""
    if not isinstance(inputs, Tensor) or not isinstance(output, Tensor):
        raise TypeError("ctc_model(inputs, output) takes the symbolic input and logits tensors (core/models.py:31-52)")
    t = output
    if not isinstance(t.producer, TimeDistributed):
        raise NotImplementedError("the logits must come from TimeDistributed(Dense(num_classes)) (core/models.py:71, 278)")
    num_classes, decays = t.producer.layer.output_dim, [t.producer.layer.weight_decay]
    t = t.parents[0]
    lstms, residual = [], None
    while isinstance(t.producer, (Bidirectional, _Merge)):
        if isinstance(t.producer, _Merge):
            new_o, skip = t.parents
            if not isinstance(new_o.producer, Bidirectional) or new_o.parents[0] is not skip:
                raise NotImplementedError("merge([new_o, o]) must add a Bidirectional layer's output to its own input "
                                          "(core/models.py:273-274)")
            residual, t = t.producer.mode, new_o
            res_here = True
        else:
            res_here = False
        lstms.append((t.producer.layer, res_here))
        t = t.parents[0]
    lstms.reverse()
    if not lstms:
        raise NotImplementedError("the engine stacks Bidirectional(LSTM) layers; none found between inputs and logits")
    if residual is not None and not all(r for _, r in lstms):
        raise NotImplementedError("the residual merge is applied to every layer or to none (core/models.py:260-276)")
    input_dropout, proj, noise = None, None, 0.0
    if isinstance(t.producer, Dropout):
        input_dropout, t = t.producer.p, t.parents[0]
    if isinstance(t.producer, TimeDistributed):
        proj, t = t.producer.layer, t.parents[0]
        decays.append(proj.weight_decay)
    if isinstance(t.producer, GaussianNoise):
        noise, t = t.producer.sigma, t.parents[0]
    if t is not inputs or t.producer is not None:
        raise NotImplementedError("unsupported layer between the input and the first Bidirectional(LSTM): %r"
                                  % type(t.producer).__name__)
    layers = [l for l, _ in lstms]
    for l in layers:
        decays += [0.0 if l.W_regularizer is None else l.W_regularizer.l2 if isinstance(l.W_regularizer, l2) else float(l.W_regularizer),
                   0.0 if l.U_regularizer is None else l.U_regularizer.l2 if isinstance(l.U_regularizer, l2) else float(l.U_regularizer)]
    if len(set(decays)) != 1:
        raise NotImplementedError("the l2 weight decay is tied across the LSTM and Dense layers (core/models.py:227-230)")
    dps = {(l.dropout_W, l.dropout_U) for l in layers}
    if len(dps) != 1 or len(set(next(iter(dps)))) != 1:
        raise NotImplementedError("dropout_W and dropout_U are tied and equal across layers (core/models.py:229-230)")
    sw = {(l.zoneout_h, l.layer_norm, l.mi) for l in layers}
    if len(sw) != 1:
        raise NotImplementedError("zoneout / layer_norm / mi are tied across layers (core/models.py:260-271)")
    zo, ln, mi = sw.pop()
    dp = float(layers[0].dropout_W)
    if input_dropout is not None and input_dropout != dp and dp > 0:
        raise NotImplementedError("the input Dropout shares the layers' dropout level (core/models.py:257-258)")
    widths = [l.output_dim for l in layers]
    num_features = inputs.shape[-1]
    if residual is not None and (proj is None or proj.output_dim != 2 * widths[0]):
        raise NotImplementedError("the residual stack needs the TimeDistributed(Dense(2 * num_hiddens)) input projection "
                                  "(core/models.py:253-255)")
    fields = dict(num_features=int(num_features), num_hiddens=widths[0], num_layers=len(layers), num_classes=int(num_classes),
                  weight_decay=float(decays[0]), dropout=dp if input_dropout is None else float(input_dropout or dp),
                  zoneout=zo, layer_norm=ln, mi=mi, residual=residual, input_dropout=input_dropout is not None and input_dropout > 0,
                  layer_hiddens=tuple(widths) if len(set(widths)) > 1 else None,
                  input_dense=(proj.output_dim if (proj is not None and residual is None) else None))
    return fields, noise


def ctc_model(inputs, output, **kwargs):
    """core/models.py:31-52: given the acoustic net's input tensor [N, T, F] and its logits tensor [N, T, C], returns the
    model [inputs, labels, inputs_length] -> [CTC loss per utterance, greedy decode].  Keyword arguments (device, seed,
    is_greedy, beam_width, merge_repeated, name) configure the engine / decoder."""
    fields, noise = _lower(inputs, output)
    spec = ModelSpec(name=kwargs.pop("name", "ctc_model"), **fields)
    kwargs.setdefault("input_std_noise", noise)
    return CTCModel(spec, **kwargs)


def graves2006(num_features=26, num_hiddens=100, num_classes=28, std=.6, **kw):
    """core/models.py:55-73: GaussianNoise(std) -> Bidirectional(LSTM(H)) -> TimeDistributed(Dense(C))."""
    x = Input(name="inputs", shape=(None, num_features))
    o = GaussianNoise(std)(x)
    o = Bidirectional(LSTM(num_hiddens, return_sequences=True, consume_less="gpu"))(o)
    o = TimeDistributed(Dense(num_classes))(o)
    return ctc_model(x, o, name="graves2006", **kw)


def eyben(num_features=39, num_hiddens=[78, 120, 27], num_classes=28, **kw):
    """core/models.py:76-103: TimeDistributed(Dense(n0)) -> Bidirectional(LSTM(n1)) -> Bidirectional(LSTM(n2)) ->
    TimeDistributed(Dense(C)); a zero entry drops that layer (:90-99).  Heterogeneous widths run on the general-cell
    engine (any H <= 1024)."""
    assert len(num_hiddens) == 3
    x = Input(name="inputs", shape=(None, num_features))
    o = x
    if num_hiddens[0]:
        o = TimeDistributed(Dense(num_hiddens[0]))(o)
    for n in num_hiddens[1:]:
        if n:
            o = Bidirectional(LSTM(n, return_sequences=True, consume_less="gpu"))(o)
    o = TimeDistributed(Dense(num_classes))(o)
    return ctc_model(x, o, name="eyben", **kw)


def deep_speech2(num_features=40, num_classes=28, num_hiddens=800, num_layers=5, dropout=0.2, weight_decay=1e-4,
                 conv_front=((32, 11, 41, 2, 2), (32, 11, 21, 1, 2)), conv_clip=20.0, **kw):
    """BASELINE configs[3]: DeepSpeech2-style 2 x Conv (32 channels, 41 x 11 and 21 x 11 frequency x time kernels,
    strides (2, 2) and (2, 1), clipped ReLU at 20; batch normalisation left out) in front of num_layers x BiLSTM and the
    Dense / CTC head of brsmv1.  NOT in the reference (README.md:118 lists Deep Speech 2 as future work): the conv
    semantics are those of include/asr_b200.h and oracle/conv.py."""
    spec = ModelSpec(int(num_features), int(num_hiddens), int(num_layers), int(num_classes), float(weight_decay),
                     "deep_speech2", float(dropout), conv_front=tuple(tuple(int(v) for v in l) for l in conv_front),
                     conv_clip=float(conv_clip))
    return CTCModel(spec, **kw)


def maas(*a, **k):
    raise NotImplementedError("maas is a SimpleRNN model outside the BiLSTM hot path (and broken in the reference)")


def deep_speech(*a, **k):
    raise NotImplementedError("deep_speech is a SimpleRNN model outside the BiLSTM hot path (and broken in the reference)")


def brsmv1(num_features=39, num_classes=28, num_hiddens=256, num_layers=5, dropout=0.2, zoneout=0.,
           input_dropout=False, input_std_noise=.0, weight_decay=1e-4, residual=None, layer_norm=None, mi=None,
           activation='tanh', **kw):
    """core/models.py:217-281: N x BiLSTM + Dense trunk with l2(weight_decay) and variational dropout
    (dropout_W = dropout_U = dropout, masks constant over time) on the tensor-core engines; zoneout and mi are
    template switches of the same kernels, layer_norm / residual='sum' (with its TimeDistributed(Dense(2H)) input
    projection) / input_dropout run on the general-cell engine (csrc/lstm_cell.cu)."""
    x = Input(name="inputs", shape=(None, num_features))
    o = x
    if input_std_noise is not None:
        o = GaussianNoise(input_std_noise)(o)
    if residual is not None:
        o = TimeDistributed(Dense(num_hiddens * 2, W_regularizer=l2(weight_decay)))(o)
    if input_dropout:
        o = Dropout(dropout)(o)
    for _ in range(num_layers):
        new_o = Bidirectional(LSTM(num_hiddens, return_sequences=True, W_regularizer=l2(weight_decay),
                                   U_regularizer=l2(weight_decay), dropout_W=dropout, dropout_U=dropout,
                                   zoneout_c=zoneout, zoneout_h=zoneout, mi=mi, layer_norm=layer_norm,
                                   activation=activation))(o)
        o = merge([new_o, o], mode=residual) if residual is not None else new_o
    o = TimeDistributed(Dense(num_classes, W_regularizer=l2(weight_decay)))(o)
    return ctc_model(x, o, name="brsmv1", **kw)
