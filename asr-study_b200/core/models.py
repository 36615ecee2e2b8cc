"""Model factories with the reference's surface (core/models.py:31-281) on the CUDA engine.

``graves2006``, ``eyben`` (BiLSTM part), ``brsmv1`` and ``ctc_model`` keep their names, keyword
arguments and defaults; they return a ``CTCModel`` that answers the Keras calls train.py /
eval.py / predict.py make on it (compile, fit_generator, evaluate_generator, predict,
train_on_batch, optimizer.lr, metrics_names, get_layer, save) — see SURVEY.md 8(b).
``maas`` / ``deep_speech`` are SimpleRNN + clipped-ReLU stacks that cannot even be constructed in
the reference (un-imported names, core/models.py:122,129): out of scope, they raise.
"""
from __future__ import annotations

import pickle
import time

import numpy as np
import torch

from ..engine import AcousticEngine, ModelSpec, pack_labels
from .layers import LSTM

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
            import os
            device = "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
        self.spec = spec
        self.engine = AcousticEngine(spec, device=device, seed=seed)
        self.device = self.engine.device
        self.optimizer = Adam()
        self.decoder = dict(is_greedy=is_greedy, beam_width=beam_width, merge_repeated=merge_repeated)
        self.input_std_noise = float(input_std_noise or 0.0)
        self.allreduce = None
        self.world_size = 1
        self.history = {}
        self._noise_rng = torch.Generator(device=self.device)
        self._noise_rng.manual_seed(1234)

    # ---- Keras-facing plumbing --------------------------------------------------
    def compile(self, loss=None, optimizer=None, metrics=None, loss_weights=None, **kw):
        if optimizer is not None:
            self.optimizer = optimizer
        return self

    def get_layer(self, name=None, index=None):
        if name in ("inputs", "labels", "inputs_length", "decoder", "ctc", "beam_search"):
            return name
        raise ValueError("No such layer: %s" % name)

    def set_data_parallel(self, allreduce, world_size):
        """allreduce(grad_slice) must SUM in place across ranks; it is called on slices that tile the flat gradient
        bucket exactly once per step, each as soon as it is complete, and may return a handle with .wait()."""
        self.allreduce, self.world_size = allreduce, int(world_size)

    # ---- batches ------------------------------------------------------------------
    def _device_batch(self, x, x_len, labels, training):
        x = np.asarray(x, dtype=np.float32)
        N, T, F = x.shape
        Np = (N + _PAD - 1) // _PAD * _PAD
        xt = torch.zeros(T, Np, F, dtype=torch.float32, device=self.device)
        xt[:, :N] = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2))).to(self.device)
        if training and self.input_std_noise > 0:           # GaussianNoise(std), core/models.py:67,251
            xt[:, :N] += self.input_std_noise * torch.randn(T, N, F, device=self.device, generator=self._noise_rng)
        lens = np.zeros(Np, np.int32)
        lens[:N] = np.asarray(x_len).reshape(-1)[:N]
        rows = _label_rows(labels)
        packed = None
        if rows is not None:
            rows = rows + [np.zeros(0, np.int32)] * (Np - N)
            packed = pack_labels(rows, self.device)
        return xt, torch.as_tensor(lens, device=self.device), packed, N

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
        return [float(v) for v in self._train_stats(x).tolist()]

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
        return self._stats(loss, self.engine.last_logits, lens, flat, off, mx, N)

    def test_on_batch(self, x, y=None):
        feats, labels, x_len = x[0], x[1], x[2]
        xt, lens, (flat, off, mx), N = self._device_batch(feats, x_len, labels, False)
        logits = self.engine.forward(xt, training=False)
        loss, _ = self.engine.ctc(logits, lens, flat, off, mx, want_grad=False)
        return [float(v) for v in self._stats(loss, logits, lens, flat, off, mx, N).tolist()]

    def predict(self, x, batch_size=None, verbose=0):
        """x = [inputs, inputs_length] (predict mode, utils/core_utils.py:76-93) -> -1-padded label matrix."""
        feats, x_len = x[0], x[1]
        xt, lens, _, N = self._device_batch(feats, x_len, None, False)
        logits = self.engine.forward(xt, training=False)
        return self._decode(logits, lens)[:N].cpu().numpy()

    predict_on_batch = predict

    def logits(self, feats, x_len):
        """batch-major logits [N,T,C] (what the reference's TimeDistributed(Dense) emits)."""
        xt, lens, _, N = self._device_batch(feats, x_len, None, False)
        return self.engine.forward(xt, training=False)[:, :N].transpose(0, 1).contiguous()

    def _reg_dev(self):
        """sum of the l2(weight_decay) regularisers (core/models.py:263-264,279) as a device scalar — a reported metric,
        not part of the step (the optimiser folds the l2 gradient in itself)."""
        P = self.engine.params
        wd = self.spec.weight_decay
        if not wd:
            return torch.zeros((), dtype=torch.float32, device=self.device)
        if getattr(self, "_decayf", None) is None:
            self._decayf = P.decay.float()
        return wd * torch.dot(P.flat * self._decayf, P.flat)

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
        import os
        eng, dev = self.engine, self.device
        beam = not self.decoder["is_greedy"]
        G = int(decode_group or (16 if beam else 4))
        main = torch.cuda.current_stream(dev)
        dec = torch.cuda.Stream(device=dev) if G > 1 else main
        # the search CTAs spread over all SMs: let the forward kernels share an SM with them (csrc/lstm_tc2.cu
        # exclusive_smem, csrc/api.cu engine selection) for the duration of this call
        saved = {k: os.environ.get(k) for k in ("ASR_LSTM_EXCLUSIVE", "ASR_B200_GEMM")}
        if beam and G > 1:
            os.environ["ASR_LSTM_EXCLUSIVE"] = "0"
            os.environ.setdefault("ASR_B200_GEMM", "tc1")
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
                prepared = [self._device_batch(x[0], x[2], x[1], False) for x in batches]
                Tmax = max(int(b[0].shape[0]) for b in prepared)
                Ntot = sum(int(b[0].shape[1]) for b in prepared)
                big = eng._buf("eval_logits%d" % p, (Tmax, Ntot, C), torch.float32)
                big_len = eng._buf("eval_len%d" % p, (Ntot,), torch.int32)
                rows, real, c0 = [], [], 0
                for x, (xt, lens, (flat, off, mx), N) in zip(batches, prepared):
                    logits = eng.forward(xt, training=False)
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
            for key, val in saved.items():
                if val is None:
                    os.environ.pop(key, None)
                else:
                    os.environ[key] = val
        n_real = sum(int(l.numel()) for l in losses)
        ctc = float(torch.cat(losses).sum().item()) / max(n_real, 1)
        return [ctc + self._reg(), ctc, 0.0, float(ler_sum.item()) / max(n_real, 1)]

    # ---- checkpoint (weights + optimiser state + meta; the .h5 wire format needs h5py: next row) ----
    def save(self, path, meta=None):
        P = self.engine.params
        blob = dict(spec=self.spec.__dict__, params=P.export("flat"), m=P.export("m"), v=P.export("v"),
                    step=self.engine.step_count, optimizer=self.optimizer.__dict__, decoder=self.decoder,
                    meta=meta or {})
        with open(path, "wb") as f:
            pickle.dump(blob, f)

    @classmethod
    def load(cls, path, device=None, **kw):
        with open(path, "rb") as f:
            blob = pickle.load(f)
        m = cls(ModelSpec(**blob["spec"]), device=device, **kw)
        P = m.engine.params
        P.load(blob["params"])
        for name in ("m", "v"):
            P.load(blob[name], which=name)
        m.engine.step_count = blob["step"]
        o = blob["optimizer"]
        m.optimizer = Adam(o["lr"], o["beta_1"], o["beta_2"], o["epsilon"], o["clipnorm"]) if o["kind"] == "adam" \
            else SGD(o["lr"], o["momentum"], o["clipnorm"])
        m.decoder = blob["decoder"]
        return m, blob.get("meta", {})


# -------------------------------------------------------------------------------------------
# factories
# -------------------------------------------------------------------------------------------
def ctc_model(inputs, output, **kwargs):
    """core/models.py:31-52.  ``inputs`` = num_features, ``output`` = list of LSTM layer records followed by
    the number of classes: ``ctc_model(26, [LSTM(100), 28])`` is graves2006."""
    *layers, num_classes = output
    input_dense = kwargs.pop("input_dense", None)
    widths = [l.output_dim for l in layers]
    hs = set(widths)
    if not layers or not all(isinstance(l, LSTM) for l in layers):
        raise NotImplementedError("the engine stacks BiLSTM layers")
    wd = kwargs.pop("weight_decay", 0.0)
    dps = {(l.dropout_W, l.dropout_U) for l in layers}
    if len(dps) != 1 or len(set(dps.pop())) != 1:
        raise NotImplementedError("dropout_W and dropout_U are tied and equal across layers (core/models.py:229-230)")
    sw = {(l.zoneout_h, l.layer_norm, l.mi) for l in layers}
    if len(sw) != 1:
        raise NotImplementedError("zoneout / layer_norm / mi are tied across layers (core/models.py:260-271)")
    zo, ln, mi = sw.pop()
    spec = ModelSpec(int(inputs), widths[0], len(layers), int(num_classes), float(wd), kwargs.pop("name", "ctc_model"),
                     float(layers[0].dropout_W), zoneout=zo, layer_norm=ln, mi=mi, residual=kwargs.pop("residual", None),
                     input_dropout=bool(kwargs.pop("input_dropout", False)),
                     layer_hiddens=tuple(widths) if len(hs) > 1 else None, input_dense=input_dense)
    return CTCModel(spec, **kwargs)


def graves2006(num_features=26, num_hiddens=100, num_classes=28, std=.6, **kw):
    """core/models.py:55-73: GaussianNoise(std) -> Bidirectional(LSTM(H)) -> TimeDistributed(Dense(C))."""
    if num_hiddens % 64:
        # the kernels take any H through the fp32 engine; H=100 is the reference default
        pass
    return ctc_model(num_features, [LSTM(num_hiddens), num_classes], name="graves2006", input_std_noise=std, **kw)


def eyben(num_features=39, num_hiddens=[78, 120, 27], num_classes=28, **kw):
    """core/models.py:76-103: TimeDistributed(Dense(n0)) -> Bidirectional(LSTM(n1)) -> Bidirectional(LSTM(n2)) ->
    TimeDistributed(Dense(C)); a zero entry drops that layer (:90-99).  Heterogeneous widths run on the general-cell
    engine (any H <= 1024)."""
    assert len(num_hiddens) == 3
    layers = [LSTM(n) for n in num_hiddens[1:] if n]
    if not layers:
        raise NotImplementedError("eyben without a recurrent layer is a plain Dense stack, outside the BiLSTM hot path")
    return ctc_model(num_features, layers + [num_classes], name="eyben", input_dense=num_hiddens[0] or None, **kw)


def maas(*a, **k):
    raise NotImplementedError("maas is a SimpleRNN model outside the BiLSTM hot path (and broken in the reference)")


def deep_speech(*a, **k):
    raise NotImplementedError("deep_speech is a SimpleRNN model outside the BiLSTM hot path (and broken in the reference)")


def brsmv1(num_features=39, num_classes=28, num_hiddens=256, num_layers=5, dropout=0.2, zoneout=0.,
           input_dropout=False, input_std_noise=.0, weight_decay=1e-4, residual=None, layer_norm=None, mi=None,
           activation='tanh', **kw):
    """core/models.py:217-281: N x BiLSTM + Dense trunk with l2(weight_decay) and variational dropout
    (dropout_W = dropout_U = dropout, masks constant over time) on the tensor-core engines.  zoneout, layer_norm,
    mi, residual='sum' (with its TimeDistributed(Dense(2H)) input projection) and input_dropout switch the
    recurrence to the general-cell engine (csrc/lstm_cell.cu)."""
    if residual not in (None, "sum"):
        raise NotImplementedError("merge mode %r: only 'sum' keeps the layer width the next Bidirectional expects" % residual)
    layers = [LSTM(num_hiddens, zoneout_c=zoneout, zoneout_h=zoneout, mi=mi, layer_norm=layer_norm,
                   activation=activation, dropout_W=dropout, dropout_U=dropout) for _ in range(num_layers)]
    return ctc_model(num_features, layers + [num_classes], name="brsmv1", weight_decay=weight_decay,
                     input_std_noise=input_std_noise, residual=residual, input_dropout=input_dropout, **kw)
