#!/usr/bin/env python
"""predict.py — the reference's prediction driver surface (predict.py:17-93): load a checkpoint, decode a dataset
subset or a single audio file, print ground truth / prediction pairs, optionally save them.

  python predict.py --model run/model.npz --dataset "dummy:split=[.5,.25]" --subset test
  python predict.py --model run/model.npz --file clip.wav --save out.json

Differences forced by this image: checkpoints are the pickle ``CTCModel.save`` writes (no h5py / Keras), the
``--no_decoder`` posteriors go to ``.npz`` instead of HDF5, ``--file`` takes RIFF/WAV (scipy, see
preprocessing.audio.load_audio).  Batch size stays 1 like the reference (predict.py:72; its "Keras' bug" note,
README "Known bugs") although the engine takes any batch."""
from __future__ import absolute_import, division, print_function

import argparse
import codecs
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from asr_study_b200.core import models as core_models                      # noqa: E402
from asr_study_b200.datasets.dataset_generator import DatasetGenerator, DatasetIterator   # noqa: E402
from asr_study_b200.utils import generic_utils as utils                      # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser(description="Predicting with an ASR system.")
    p.add_argument("--model", required=True, type=str)
    p.add_argument("--dataset", default=None, type=str)
    p.add_argument("--file", default=None, type=str)
    p.add_argument("--subset", type=str, default="test")
    p.add_argument("--input_parser", type=str, default=None)
    p.add_argument("--input_parser_params", nargs="+", default=[])
    p.add_argument("--label_parser", type=str, default="simple_char_parser")
    p.add_argument("--label_parser_params", nargs="+", default=[])
    p.add_argument("--no_decoder", action="store_true", default=False)
    p.add_argument("--gpu", default="0", type=str)
    p.add_argument("--allow_growth", default=False, action="store_true")
    p.add_argument("--save", default=None, type=str)
    p.add_argument("--override", default=False, action="store_true")
    args = p.parse_args(argv)
    if args.dataset is None and args.file is None:
        raise ValueError("dataset or file args must be set.")
    if args.dataset and args.file:
        print("Both dataset and file args was set. Ignoring file args.")
    model, meta = core_models.CTCModel.load(args.model, device="cuda:%s" % args.gpu.split(",")[0])
    targs = meta.get("training_args", {})
    ip = args.input_parser or targs.get("input_parser")
    ipp = args.input_parser_params or targs.get("input_parser_params", [])
    input_parser = utils.get_from_module("preprocessing.audio", ip, params=ipp)
    label_parser = utils.get_from_module("preprocessing.text", args.label_parser, params=args.label_parser_params)
    if args.dataset is not None:
        gen = DatasetGenerator(input_parser, label_parser, batch_size=1, seed=0, mode="predict", shuffle=False)
        flow = gen.flow_from_fname(args.dataset, datasets=args.subset)
        truths = list(flow.labels)
        names = ["%s[%d]" % (args.dataset, i) for i in range(flow.len)]
    else:
        flow = DatasetIterator([args.file], None, batch_size=1, input_parser=input_parser, label_parser=label_parser,
                               mode="predict", shuffle=False)
        truths, names = [u""], [args.file]
    results = []
    for index in range(flow.len):
        x, x_len = next(flow)
        if args.no_decoder:
            pred = model.logits(x, x_len)[0].cpu().numpy()                   # [T, C] linear outputs (softmax lives in CTC)
            shown = "logits %s" % (pred.shape,)
        else:
            ids = model.predict([x, x_len])[0]
            pred = label_parser.imap(ids)
            shown = pred
        results.append({"input": names[index], "label": truths[index], "best": pred})
        print("Ground Truth: %s" % truths[index])
        print("   Predicted: %s\n\n" % shown)
    if args.save is not None:
        if os.path.exists(args.save):
            if not args.override:
                raise IOError("Unable to create file")
            os.remove(args.save)
        if args.no_decoder:
            arr = np.empty(len(results), dtype=object)
            arr[:] = [r["best"].astype(np.float32) for r in results]
            np.savez(args.save, predictions=arr, labels=np.asarray([r["label"] for r in results]),
                     inputs=np.asarray([r["input"] for r in results]), num_labels=np.int64(results[0]["best"].shape[-1]))
        else:
            with codecs.open(args.save, "w", encoding="utf8") as f:
                json.dump(results, f)
    return results


if __name__ == "__main__":
    main()
