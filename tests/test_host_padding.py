"""Host logic of the zero-padded tensor-core widths (engine.tc_width / ParamBucket.load / export) on CPU tensors:
the padded bucket holds the model's values in the leading H entries of every gate / direction block, zeros
elsewhere, and export() returns exactly what load() was given."""
import numpy as np
import torch

from asr_study_b200.engine import ModelSpec, ParamBucket, _pad_blocks, _unpad_blocks, tc_width, TC_WIDTHS


def test_tc_width_rule():
    assert [tc_width(h) for h in (64, 100, 128, 129, 200, 256, 500, 512, 513, 800, 832, 896, 897, 1024)] == \
           [64, 100, 128, 256, 256, 256, 512, 512, 640, 832, 832, 896, 897, 1024]
    assert all(tc_width(w) == w for w in TC_WIDTHS)


def test_pad_blocks_round_trip():
    rng = np.random.RandomState(0)
    a = rng.randn(6, 4 * 5).astype(np.float32)
    b = _pad_blocks(a, 1, 4, 5, 8)
    assert b.shape == (6, 32)
    for g in range(4):
        np.testing.assert_array_equal(b[:, 8 * g:8 * g + 5], a[:, 5 * g:5 * g + 5])
        assert not b[:, 8 * g + 5:8 * g + 8].any()
    np.testing.assert_array_equal(_unpad_blocks(b, 1, 4, 5, 8), a)
    c = rng.randn(2 * 5, 3).astype(np.float32)
    np.testing.assert_array_equal(_unpad_blocks(_pad_blocks(c, 0, 2, 5, 8), 0, 2, 5, 8), c)


def test_param_bucket_pads_every_tensor_per_block_and_exports_the_model_shapes():
    F, H, Hp, L, C = 7, 200, 256, 2, 5
    rng = np.random.RandomState(1)
    params, D = {}, F
    for l in range(L):
        for d in "fb":
            params[f"l{l}.W{d}"] = rng.randn(D, 4 * H).astype(np.float32)
            params[f"l{l}.U{d}"] = rng.randn(H, 4 * H).astype(np.float32)
            params[f"l{l}.b{d}"] = rng.randn(4 * H).astype(np.float32)
        for n in ("mi_alpha", "mi_beta1", "mi_beta2"):
            params[f"l{l}.{n}"] = rng.randn(2, 4 * H).astype(np.float32)
        D = 2 * H
    params["dense.W"] = rng.randn(D, C).astype(np.float32)
    params["dense.b"] = rng.randn(C).astype(np.float32)
    dev_spec = ModelSpec(F, Hp, L, C, mi=(1.0, 0.5, 0.5))
    P = ParamBucket(dev_spec, torch.device("cpu"), logical_h=H)
    assert set(P.shapes) == set(params) and P.shapes["l1.Wf"] == (2 * Hp, 4 * Hp)
    P.load(params)
    out = P.export("flat")
    for k, v in params.items():
        assert out[k].shape == v.shape
        np.testing.assert_array_equal(out[k], v)
    # the padding is zero: the padded bucket carries exactly the model's mass
    total = sum(float(np.abs(v).sum()) for v in params.values())
    assert abs(float(P.flat.abs().sum()) - total) <= 1e-4 * total
    # gate block g of W starts at g*Hp; direction block of a deeper layer's input rows at Hp
    W1 = P.p("l1.Wf").numpy()
    np.testing.assert_array_equal(W1[Hp:Hp + H, 2 * Hp:2 * Hp + H], params["l1.Wf"][H:2 * H, 2 * H:3 * H])
    assert not W1[H:Hp].any() and not W1[:, H:Hp].any()
    # Adam moments go through the same path (checkpoint load)
    P.load({k: np.ones_like(v) for k, v in params.items()}, which="m")
    assert all(np.array_equal(v, np.ones_like(v)) for v in P.export("m").values())
