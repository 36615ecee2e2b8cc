"""Offline featuriser — the reference's ``python -m extras.make_dataset`` (extras/make_dataset.py:10-54 ->
datasets/dataset_parser.py:87-177 ``to_h5``) driving the fused K1 kernel in bulk.

The reference walks a corpus, featurises every utterance with ``input_parser`` on the CPU and writes
``inputs`` / ``labels`` / ``durations`` per split into an HDF5 file.  Here the corpus clips are batched and each
batch is ONE asr_mfcc_forward launch on the device (same numbers as ``input_parser(audio)`` per utterance);
h5py is not installable in this image, so the container is a NumPy ``.npz`` with the same groups
(``<split>/inputs`` as an object array of [T, F] float32 matrices, ``<split>/labels``, ``<split>/durations``,
plus the ``num_feats`` / ``input_parser`` / ``label_parser`` attributes).  Corpus parsers other than the
in-memory ``dummy`` (brsd, cslu, lapsbm, sid, voxforge) walk file trees that do not exist here: out of scope.

  python -m asr_study_b200.extras.make_dataset --parser dummy --parser_params num_speakers 2 \
      num_utterances_per_speaker 4 "split" "[.5,.25]" --input_parser mfcc --input_parser_params num_cep 13 dd False \
      --output_file /tmp/dummy_mfcc.npz
"""
from __future__ import absolute_import, division, print_function

import argparse
import os

import numpy as np

from ..utils import generic_utils as utils
from ..utils.hparams import HParams


def featurise(dl, input_parser, label_parser=None, batch_size=64):
    """dict-of-lists corpus (datasets/dataset_parser.py:48-85) -> {split: dict(inputs, labels, durations)}."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    splits = dl.get("dataset") or ["all"] * len(dl["input"])
    out = {}
    for name in sorted(set(splits)):
        idx = [i for i, d in enumerate(splits) if d == name]
        feats = []
        for b in range(0, len(idx), batch_size):
            clips = [np.ascontiguousarray(np.asarray(dl["input"][i], np.float32).reshape(-1)) for i in idx[b:b + batch_size]]
            if input_parser is None or not hasattr(input_parser, "batch"):
                feats += [np.asarray(c if input_parser is None else input_parser(c), np.float32) for c in clips]
                continue
            off = np.zeros(len(clips) + 1, np.int64)
            off[1:] = np.cumsum([len(c) for c in clips])
            x, lens = input_parser.batch(torch.from_numpy(np.concatenate(clips)).to(dev), torch.from_numpy(off).to(dev),
                                         time_major=False)
            x, lens = x.cpu().numpy(), lens.cpu().numpy()
            feats += [x[k, :lens[k]].copy() for k in range(len(clips))]
        labels = [dl["label"][i] for i in idx]
        if label_parser is not None:
            labels = [np.asarray(label_parser(l), np.int32) for l in labels]
        inputs = np.empty(len(feats), dtype=object)
        inputs[:] = feats
        lab = np.empty(len(labels), dtype=object)
        lab[:] = labels
        out[name] = dict(inputs=inputs, labels=lab, durations=np.asarray([dl["duration"][i] for i in idx], np.float64))
    return out


def main(argv=None):
    p = argparse.ArgumentParser(description="Generates a preprocessed dataset by providing the dataset and the correct parser.")
    p.add_argument("--dataset_dir", type=str, default=None)
    p.add_argument("--parser", type=str, required=True)
    p.add_argument("--parser_params", nargs="+", default=[])
    p.add_argument("--output_file", type=str, default=None)
    p.add_argument("--input_parser", type=str, default=None)
    p.add_argument("--input_parser_params", nargs="+", default=[])
    p.add_argument("--label_parser", type=str, default=None)
    p.add_argument("--label_parser_params", nargs="+", default=[])
    p.add_argument("--override", action="store_true")
    args = p.parse_args(argv)
    if args.parser.lower() != "dummy":
        raise NotImplementedError("corpus parser %r walks files that are not available here; 'dummy' is built" % args.parser)
    from ..datasets.dummy import Dummy
    input_parser = utils.get_from_module("preprocessing.audio", args.input_parser, params=args.input_parser_params)
    label_parser = utils.get_from_module("preprocessing.text", args.label_parser, params=args.label_parser_params)
    dataset = Dummy(**HParams().parse(args.parser_params).values())
    fname = args.output_file or "%s.npz" % dataset.name
    if os.path.exists(fname) and not args.override:
        raise IOError("Unable to create file %s (exists; use --override)" % fname)
    groups = featurise(dataset.to_dict_list(), input_parser, label_parser)
    flat = {"num_feats": np.int64(getattr(input_parser, "num_feats", 0) or 0), "input_parser": str(args.input_parser),
            "label_parser": str(args.label_parser)}
    for split, g in groups.items():
        for k, v in g.items():
            flat["%s/%s" % (split, k)] = v
    np.savez(fname, **flat)
    print("Dataset %s saved at %s" % (dataset.name, fname))
    return fname


if __name__ == "__main__":
    main()
