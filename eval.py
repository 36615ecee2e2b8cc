#!/usr/bin/env python
"""eval.py — the reference's evaluation driver surface (eval.py:26-82): load a checkpoint, switch the
decoder to beam search (width 400 as utils/core_utils.py:70-71 does for mode='eval'), evaluate a subset."""
from __future__ import absolute_import, division, print_function

import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from asr_study_b200.core import models as core_models                    # noqa: E402
from asr_study_b200.datasets.dataset_generator import DatasetGenerator     # noqa: E402
from asr_study_b200.utils import generic_utils as utils                    # noqa: E402


def main(argv=None):
    p = argparse.ArgumentParser(description="Evaluating an ASR system.")
    p.add_argument("--model", required=True, type=str)
    p.add_argument("--dataset", required=True, type=str)
    p.add_argument("--subset", type=str, default="test")
    p.add_argument("--batch_size", default=32, type=int)
    p.add_argument("--input_parser", type=str, default=None)
    p.add_argument("--input_parser_params", nargs="+", default=[])
    p.add_argument("--label_parser", type=str, default="simple_char_parser")
    p.add_argument("--label_parser_params", nargs="+", default=[])
    p.add_argument("--gpu", default="0", type=str)
    p.add_argument("--allow_growth", default=False, action="store_true")
    p.add_argument("--save_transcriptions", default=None, type=str)
    p.add_argument("--greedy", default=False, action="store_true", help="keep the training-time best-path decoder")
    p.add_argument("--beam_width", default=400, type=int)
    args = p.parse_args(argv)
    model, meta = core_models.CTCModel.load(args.model, device="cuda:%s" % args.gpu.split(",")[0])
    targs = meta.get("training_args", {})
    if not args.greedy:
        model.decoder = dict(is_greedy=False, beam_width=args.beam_width, merge_repeated=True)
    ip = args.input_parser or targs.get("input_parser")
    ipp = args.input_parser_params or targs.get("input_parser_params", [])
    input_parser = utils.get_from_module("preprocessing.audio", ip, params=ipp)
    label_parser = utils.get_from_module("preprocessing.text", args.label_parser, params=args.label_parser_params)
    data_gen = DatasetGenerator(input_parser, label_parser, batch_size=args.batch_size, seed=0)
    flow = data_gen.flow_from_fname(args.dataset, datasets=args.subset)
    m = model.evaluate_generator(flow, flow.len, max_q_size=10, nb_worker=1)
    for name, v in zip(model.metrics_names, m):
        print("%s: %4f" % (name, v))
    return m


if __name__ == "__main__":
    main()
