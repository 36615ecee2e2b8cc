#!/usr/bin/env python
"""train.py — the reference's training driver surface (train.py:46-235) on the B200 engine.

Same flags; `--dataset dummy[:k=v,...]` selects the in-memory synthetic corpus (datasets/dummy.py).
Launch under torchrun for data-parallel training (one NCCL all-reduce of the gradient bucket per step).

  python train.py --dataset "dummy:num_speakers=2,num_utterances_per_speaker=8,split=[.5,.25]" \
      --input_parser mfcc --input_parser_params num_cep 13 dd False \
      --model graves2006 --model_params num_features 26 --batch_size 2 --num_epochs 2
"""
from __future__ import absolute_import, division, print_function

import argparse
import datetime
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from asr_study_b200.core import models as core_models                     # noqa: E402
from asr_study_b200.core.ctc_utils import ctc_dummy_loss, decoder_dummy_loss  # noqa: E402
from asr_study_b200.core import metrics                                     # noqa: E402
from asr_study_b200.core.callbacks import LR_SCHEDULES, MetaCheckpoint      # noqa: E402
from asr_study_b200.datasets.dataset_generator import DatasetGenerator      # noqa: E402
from asr_study_b200.utils import generic_utils as utils                     # noqa: E402
from asr_study_b200.utils.hparams import HParams                            # noqa: E402


def build_parser():
    p = argparse.ArgumentParser(description="Training an ASR system.")
    p.add_argument("--load", default=None, type=str)
    p.add_argument("--model", default="brsmv1", type=str)
    p.add_argument("--model_params", nargs="+", default=[])
    p.add_argument("--num_epochs", default=100, type=int)
    p.add_argument("--lr", default=0.001, type=float)
    p.add_argument("--momentum", default=0.9, type=float)
    p.add_argument("--clipnorm", default=400, type=float)
    p.add_argument("--batch_size", default=32, type=int)
    p.add_argument("--opt", default="adam", type=str, choices=["sgd", "adam"])
    p.add_argument("--dataset", default=None, type=str, nargs="+")
    p.add_argument("--input_parser", type=str, default=None)
    p.add_argument("--input_parser_params", nargs="+", default=[])
    p.add_argument("--label_parser", type=str, default="simple_char_parser")
    p.add_argument("--label_parser_params", nargs="+", default=[])
    p.add_argument("--lr_schedule", default=None)
    p.add_argument("--lr_params", nargs="+", default=[])
    p.add_argument("--save", default=None, type=str)
    p.add_argument("--gpu", default="0", type=str)
    p.add_argument("--allow_growth", default=False, action="store_true")
    p.add_argument("--verbose", default=0, type=int)
    p.add_argument("--seed", default=None, type=float)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", args.gpu.split(",")[0] if world == 1 else "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    meta, epoch_offset = None, 0
    if args.load:
        # train.py:107-122: the saved training arguments are the defaults, flags given on this command line win
        defaults = build_parser().parse_args([])
        given = {k: v for k, v in vars(args).items() if v != getattr(defaults, k)}
        model, meta = core_models.CTCModel.load(args.load, device="cuda:%d" % local)
        merged = dict(meta.get("training_args", {}))
        merged.update(given)
        for k, v in merged.items():
            if hasattr(args, k):
                setattr(args, k, v)
        epoch_offset = len(meta.get("epochs", []))
        if "lr" in given:
            model.optimizer.lr = float(args.lr)
    else:
        model_fn = utils.get_from_module("core.models", args.model)
        model = model_fn(**(HParams().parse(args.model_params).values()), device="cuda:%d" % local,
                         **({"seed": int(args.seed)} if args.seed is not None else {}))
        if args.opt.strip().lower() == "sgd":
            opt = core_models.SGD(lr=args.lr, momentum=args.momentum, clipnorm=args.clipnorm)
        else:
            opt = core_models.Adam(lr=args.lr, clipnorm=args.clipnorm)
        model.compile(loss={"ctc": ctc_dummy_loss, "decoder": decoder_dummy_loss}, optimizer=opt,
                      metrics={"decoder": metrics.ler}, loss_weights=[1, 0])
    if world > 1:
        import torch.distributed as dist
        # called on per-layer slices of the flat gradient bucket as they complete (overlaps the BPTT recurrences)
        model.set_data_parallel(lambda g: dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True), world,
                                rank=int(os.environ.get("RANK", "0")))

    output_dir = args.save or os.path.join("results", "%s_%s" % (args.model, datetime.datetime.now()))
    os.makedirs(output_dir, exist_ok=True)

    input_parser = utils.get_from_module("preprocessing.audio", args.input_parser, params=args.input_parser_params)
    label_parser = utils.get_from_module("preprocessing.text", args.label_parser, params=args.label_parser_params)
    data_gen = DatasetGenerator(input_parser, label_parser, batch_size=args.batch_size, seed=args.seed)
    if not args.dataset:
        raise SystemExit("--dataset is required")
    test_flow = None
    if len(args.dataset) == 1:
        train_flow, valid_flow, test_flow = data_gen.flow_from_fname(args.dataset[0], datasets=["train", "valid", "test"])
    else:
        train_flow = data_gen.flow_from_fname(args.dataset[0])
        valid_flow = data_gen.flow_from_fname(args.dataset[1])
        if len(args.dataset) == 3:
            test_flow = data_gen.flow_from_fname(args.dataset[2])
    print(str(vars(args)))

    # train.py:153-170: model / best checkpoints (+ meta) and the optional learning-rate schedule
    callback_list = [MetaCheckpoint(os.path.join(output_dir, "model.npz"), training_args=args, meta=meta),
                     MetaCheckpoint(os.path.join(output_dir, "best.npz"), monitor="val_decoder_ler", save_best_only=True,
                                    mode="min", training_args=args, meta=meta)]
    if args.lr_schedule:
        fn = LR_SCHEDULES.get(str(args.lr_schedule).lower().strip())
        if fn is None:
            raise ValueError("Learning rate schedule unrecognized")
        callback_list.append(fn(**HParams().parse(args.lr_params).values()))

    model.fit_generator(train_flow, samples_per_epoch=train_flow.len, nb_epoch=args.num_epochs,
                        validation_data=valid_flow, nb_val_samples=valid_flow.len if valid_flow else 0, max_q_size=10,
                        nb_worker=1, callbacks=callback_list, verbose=1, initial_epoch=epoch_offset)
    if test_flow is not None and test_flow.len:
        # train.py:219-233: the best checkpoint, reloaded in 'eval' mode (beam search instead of the greedy decoder)
        best = os.path.join(output_dir, "best.npz")
        if world > 1:
            import torch.distributed as dist
            dist.barrier()                                      # rank 0 wrote it
        if os.path.exists(best):
            model, _ = core_models.CTCModel.load(best, device="cuda:%d" % local, mode="eval")
        m = model.evaluate_generator(test_flow, test_flow.len, max_q_size=10, nb_worker=1)
        msg = "Total loss: %.4f\nCTC Loss: %.4f\nLER: %.2f%%" % (m[0], m[1], m[3] * 100)
        if int(os.environ.get("RANK", "0")) == 0:
            with open(os.path.join(output_dir, "results.txt"), "w") as f:
                f.write(msg)
        print(msg)
    return model


if __name__ == "__main__":
    main()
