"""Data-parallel host logic on CPU: world_size-2 gloo.  Checks that sharding the batch across ranks with
grad_scale = 1/global_batch and ONE all-reduce(SUM) of the flat gradient bucket reproduces the full-batch
gradient, and that every rank then takes the identical optimiser step (replicas stay in sync without a
broadcast) — the contract bench.py / train.py rely on.  Arithmetic here is the oracle's; the
CUDA kernels are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    N, T, F, H, L, C = 4, 12, 6, 8, 2, 7
    params = om.init_params(F, H, L, C, seed=3)
    x = rng.randn(N, T, F).astype(np.float32)
    lens = np.array([12, 10, 12, 7])
    labels = [rng.randint(0, C - 1, size=3) for _ in range(N)]
    lo, hi = rank * N // world, (rank + 1) * N // world
    _, ctc, grads, _ = om.loss_and_grads(params, x[lo:hi], lens[lo:hi], labels[lo:hi], dtype=np.float64,
                                         global_batch=N)
    keys = sorted(grads)
    flat = torch.from_numpy(np.concatenate([grads[k].ravel() for k in keys]))      # the flat bucket
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)                                      # the ONE collective per step
    out, off = {}, 0
    for k in keys:
        n = grads[k].size
        out[k] = flat[off:off + n].numpy().reshape(grads[k].shape).astype(np.float32)
        off += n
    st = {}
    om.clip_adam_step(params, out, st, lr=1e-3, clipnorm=400.0)
    q.put((rank, out, {k: v.copy() for k, v in params.items()}))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_matches_full_batch_and_keeps_replicas_identical():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.RandomState(0)
    N, T, F, H, L, C = 4, 12, 6, 8, 2, 7
    params = om.init_params(F, H, L, C, seed=3)
    x = rng.randn(N, T, F).astype(np.float32)
    lens = np.array([12, 10, 12, 7])
    labels = [rng.randint(0, C - 1, size=3) for _ in range(N)]
    _, _, full, _ = om.loss_and_grads(params, x, lens, labels, dtype=np.float64)
    for k in full:
        np.testing.assert_allclose(res[0][1][k], full[k], rtol=1e-5, atol=1e-7)
        assert np.array_equal(res[0][1][k], res[1][1][k])            # same reduced gradient on both ranks
        assert np.array_equal(res[0][2][k], res[1][2][k])            # identical parameters after the step
