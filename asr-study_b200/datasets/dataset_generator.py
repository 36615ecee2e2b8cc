"""Batch iterators with the reference's contract (datasets/dataset_generator.py:129-251):

    next(it) -> ([x f32 [N,Tmax,F] zero-padded 'post', labels scipy COO int32 [N,Lmax], x_len [N]],
                 [zeros [N], labels])                      (train / eval)
             -> [x, x_len]                                  (predict)

``input_parser`` is a preprocessing.audio Feature.  When it has the batched device entry
(``.batch``), the whole batch's features come from ONE fused-kernel launch on the GPU instead of the
reference's per-utterance python loop (dataset_generator.py:225); the padding, lengths and the
"index_array.sort()" quirk are kept.  HDF5 / JSON readers are out of scope (no h5py here).
"""
import threading

import numpy as np
import scipy.sparse


class DatasetIterator(object):
    """Data parallel: under torchrun every rank builds the SAME iterator (same seed => same permutation) and
    takes every world_size-th utterance of it, i.e. its share of each global batch of batch_size * world_size
    utterances; ``len`` is then the rank's share of the corpus, so train.py's samples_per_epoch = flow.len walks the
    corpus once per epoch across all ranks (the tail is wrapped so that every rank yields the same number of
    utterances and no rank ever sees an empty batch).  rank / world_size default to RANK / WORLD_SIZE."""

    def __init__(self, inputs, labels=None, batch_size=32, shuffle=False, seed=None, input_parser=None,
                 label_parser=None, mode="train", rank=None, world_size=None):
        if labels is not None and len(inputs) != len(labels):
            raise ValueError("inputs and labels should have the same length. Found: len(inputs) = %s, "
                             "len(labels) = %s" % (len(inputs), len(labels)))
        import os
        self.inputs, self.labels = list(inputs), (list(labels) if labels is not None else None)
        self.batch_size, self.shuffle = batch_size, shuffle
        self.input_parser, self.label_parser, self.mode = input_parser, label_parser, mode
        self.world_size = int(os.environ.get("WORLD_SIZE", "1")) if world_size is None else int(world_size)
        self.rank = int(os.environ.get("RANK", "0")) if rank is None else int(rank)
        if not 0 <= self.rank < self.world_size:
            raise ValueError("rank %d outside world of %d" % (self.rank, self.world_size))
        self.lock = threading.Lock()
        if seed is None and self.world_size > 1:
            seed = 1234                                       # the ranks must draw the same permutation
        self._rng = np.random.RandomState(None if seed is None else int(seed))
        self._order, self._pos = None, 0
        self._gpu = None

    @property
    def len(self):
        n, w = len(self.inputs), self.world_size
        return (n + w - 1) // w

    def __iter__(self):
        return self

    def _next_indices(self):
        n, w = len(self.inputs), self.world_size
        if self._order is None or self._pos >= len(self._order):
            order = self._rng.permutation(n) if self.shuffle else np.arange(n)
            if w > 1:
                pad = (-n) % w
                order = np.concatenate([order, order[:pad]])[self.rank::w]
            self._order, self._pos = order, 0
        idx = self._order[self._pos:self._pos + self.batch_size]
        self._pos += self.batch_size
        return np.sort(idx)                                   # dataset_generator.py:200

    def __next__(self):
        with self.lock:
            idx = self._next_indices()
        batch_inputs, batch_len = self._make_in([self.inputs[i] for i in idx])
        batch_labels = self._make_out([self.labels[i] for i in idx]) if self.labels is not None else None
        if batch_labels is None or self.mode == "predict":
            return [batch_inputs, batch_len]
        return ([batch_inputs, batch_labels, batch_len], [np.zeros((batch_inputs.shape[0],)), batch_labels])

    next = __next__

    def _make_in(self, inputs):
        p = self.input_parser
        if p is not None and any(isinstance(i, str) for i in inputs):      # file paths (audio.py:55-59)
            from ..preprocessing.audio import load_audio
            inputs = [load_audio(i, p.fs) if isinstance(i, str) else i for i in inputs]
        if p is not None and hasattr(p, "batch") and str(p) != "raw":
            import torch
            clips = [np.asarray(c, dtype=np.float32).reshape(-1) for c in inputs]
            lens_s = [len(c) for c in clips]
            off = np.zeros(len(clips) + 1, np.int64)
            off[1:] = np.cumsum(lens_s)
            dev = torch.device("cuda", torch.cuda.current_device())
            # the generator thread works on its OWN stream with pinned staging buffers: its copies and the K1 launch then
            # overlap the training step on the model's streams instead of queueing in front of it on the default stream
            st = self._gpu
            if st is None or st["dev"] != dev:
                st = self._gpu = dict(dev=dev, stream=torch.cuda.Stream(device=dev), pcm=None, out=None)
            total = int(off[-1])
            if st["pcm"] is None or st["pcm"].numel() < total:
                st["pcm"] = torch.empty(total, dtype=torch.float32).pin_memory()
            hp = st["pcm"][:total]
            hn = hp.numpy()
            for c, o in zip(clips, off[:-1]):
                hn[o:o + len(c)] = c
            with torch.cuda.stream(st["stream"]):
                pcm_d = hp.to(dev, non_blocking=True)
                off_d = torch.from_numpy(off).to(dev, non_blocking=True)
                t_max = max(p.num_frames(n) for n in lens_s)
                feats, lens = p.batch(pcm_d, off_d, t_max=t_max, time_major=False)
                need = feats.numel()
                if st["out"] is None or st["out"].numel() < need:
                    st["out"] = torch.empty(need, dtype=torch.float32).pin_memory()
                ho = st["out"][:need].view(feats.shape)
                ho.copy_(feats, non_blocking=True)
                lens_h = lens.to("cpu", non_blocking=False)
            st["stream"].synchronize()
            return ho.numpy().copy(), lens_h.numpy().astype(np.int64)
        if p is not None:
            inputs = [p(i) for i in inputs]
        lens = np.asarray([np.asarray(i).shape[0] for i in inputs])
        F = np.asarray(inputs[0]).shape[1]
        x = np.zeros((len(inputs), int(lens.max()), F), dtype=np.float32)          # pad_sequences(.., 'post')
        for k, f in enumerate(inputs):
            x[k, :lens[k]] = f
        return x, lens

    def _make_out(self, labels):
        if self.label_parser is not None:
            labels = [self.label_parser(l) for l in labels]
        rows, cols, data = [], [], []
        for r, lab in enumerate(labels):
            cols.extend(range(len(lab)))
            rows.extend(len(lab) * [r])
            data.extend(lab)
        return scipy.sparse.coo_matrix((data, (rows, cols)), shape=(len(labels), max(len(l) for l in labels)),
                                       dtype="int32")


class DatasetGenerator(object):
    """datasets/dataset_generator.py:24-126 for in-memory data (dict-of-lists, e.g. Dummy.to_dict_list())."""

    def __init__(self, input_parser=None, label_parser=None, batch_size=32, shuffle=True, seed=None, mode="train"):
        self.input_parser, self.label_parser = input_parser, label_parser
        self.batch_size, self.shuffle, self.seed, self.mode = batch_size, shuffle, seed, mode

    def flow(self, inputs, labels, rank=None, world_size=None):
        return DatasetIterator(inputs, labels, batch_size=self.batch_size, shuffle=self.shuffle, seed=self.seed,
                               input_parser=self.input_parser, label_parser=self.label_parser, mode=self.mode,
                               rank=rank, world_size=world_size)

    def flow_from_dl(self, dl, datasets=None):
        def pick(name):
            if name is None or "dataset" not in dl:
                return self.flow(dl["input"], dl["label"])
            keep = [i for i, d in enumerate(dl["dataset"]) if d == name]
            return self.flow([dl["input"][i] for i in keep], [dl["label"][i] for i in keep])
        if datasets is None:
            return pick(None)
        if isinstance(datasets, str):
            return pick(datasets)
        return [pick(d) for d in datasets]

    def flow_from_fname(self, fname, datasets=None):
        if isinstance(fname, str) and fname.startswith("dummy"):
            from .dummy import Dummy
            kw = {}
            if ":" in fname:                                   # dummy:num_speakers=2,num_utterances_per_speaker=4
                import ast
                import re
                for kv in re.split(r",(?![^\[]*\])", fname.split(":", 1)[1]):   # commas outside [...]
                    k, v = kv.split("=")
                    kw[k] = ast.literal_eval(v)
            return self.flow_from_dl(Dummy(**kw).to_dict_list(), datasets)
        raise NotImplementedError("HDF5 / JSON dataset files need h5py / corpora that are not available here; "
                                  "use 'dummy[:k=v,...]' or DatasetGenerator.flow(inputs, labels)")
