"""The data-parallel hook: engine.backward hands the gradient bucket to `allreduce` once (the default) or, with
`dp_slices`, in slices that tile it exactly
once (the Dense layer first, then per layer [Wf|Wb] behind the dW GEMMs and the rest behind the dU GEMMs; layer 0 whole), and waits for returned handles."""
import numpy as np
import pytest
import torch

from tests.util_gpu import dev

pytestmark = pytest.mark.gpu


def test_allreduce_slices_tile_the_bucket_once_and_scaling_matches():
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    N, T, F, H, L, C = 8, 16, 26, 64, 3, 28
    rng = np.random.RandomState(0)
    x = dev(rng.randn(T, N, F).astype(np.float32))
    lens = dev(np.full(N, T, np.int32))
    labels = [rng.randint(0, C - 1, size=3).astype(np.int32) for _ in range(N)]
    flat, off, mx = pack_labels(labels, "cuda")
    ref = AcousticEngine(ModelSpec(F, H, L, C), seed=1)
    ref.train_step(x, lens, flat, off, mx, global_batch=N, lr=0.0, clipnorm=0.0)
    g_ref = ref.params.grad.clone()

    class Handle:
        waited = 0

        def wait(self):
            Handle.waited += 1

    eng = AcousticEngine(ModelSpec(F, H, L, C), seed=1)
    eng.dp_slices = True
    seen = torch.zeros_like(eng.params.grad)
    base = eng.params.grad.data_ptr()

    def fake_allreduce(g):                       # "sum over 2 identical ranks": doubles the slice in place
        lo = (g.data_ptr() - base) // 4
        seen[lo:lo + g.numel()] += 1
        g.mul_(2.0)
        return Handle()

    eng.train_step(x, lens, flat, off, mx, global_batch=2 * N, allreduce=fake_allreduce, lr=0.0, clipnorm=0.0)
    torch.cuda.synchronize()
    assert bool((seen == 1).all()) and Handle.waited == 2 * L
    # two identical ranks at global batch 2N give exactly the single-rank gradient at batch N
    torch.testing.assert_close(eng.params.grad, g_ref, rtol=1e-5, atol=1e-7)


def test_small_bucket_default_is_one_collective_of_the_whole_bucket():
    from asr_study_b200.engine import AcousticEngine, ModelSpec, pack_labels
    N, T, F, H, L, C = 8, 16, 26, 64, 2, 28
    rng = np.random.RandomState(1)
    x = dev(rng.randn(T, N, F).astype(np.float32))
    lens = dev(np.full(N, T, np.int32))
    labels = [rng.randint(0, C - 1, size=3).astype(np.int32) for _ in range(N)]
    flat, off, mx = pack_labels(labels, "cuda")
    eng = AcousticEngine(ModelSpec(F, H, L, C), seed=1)
    calls = []

    def fake_allreduce(g):
        calls.append((g.data_ptr(), g.numel()))

    eng.train_step(x, lens, flat, off, mx, global_batch=N, allreduce=fake_allreduce, lr=0.0, clipnorm=0.0)
    torch.cuda.synchronize()
    assert calls == [(eng.params.grad.data_ptr(), eng.params.grad.numel())]
