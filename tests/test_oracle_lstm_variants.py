"""Pins for the oracle's restatement of LSTM.step with the brsmv1 switches (core/layers.py:432-469,
core/layers_utils.py:16-51): the reference ships no test for them, so the numpy forward/backward is pinned by
(1) central finite differences in fp64 and (2) an independent torch-autograd transcription of the step."""
import numpy as np
import torch

from oracle import lstm as ol
from oracle import model as om


def _case(seed=0, N=3, T=5, D=4, H=6):
    rng = np.random.RandomState(seed)
    x = rng.randn(N, T, D)
    W, U, b = rng.randn(D, 4 * H) * 0.5, rng.randn(H, 4 * H) * 0.5, rng.randn(4 * H) * 0.1
    ln = {"uh": (1 + 0.1 * rng.randn(4 * H), 0.1 * rng.randn(4 * H)), "wx": (1 + 0.1 * rng.randn(4 * H), 0.1 * rng.randn(4 * H)),
          "c": (1 + 0.1 * rng.randn(H), 0.1 * rng.randn(H))}
    v = ol.make_variant(H, mi=(1 + 0.1 * rng.randn(4 * H), 0.5 + 0.1 * rng.randn(4 * H), 0.5 + 0.1 * rng.randn(4 * H)),
                        layer_norm=ln, zoneout_h=0.3, zoneout_c=0.3, zmask_h=(rng.rand(T, H) > 0.3).astype(float),
                        zmask_c=(rng.rand(T, H) > 0.3).astype(float))
    mW, mU = (rng.rand(N, D) > 0.2) / 0.8, (rng.rand(N, H) > 0.2) / 0.8
    G = rng.randn(N, T, H)
    return x, W, U, b, v, mW, mU, G


def test_cell_backward_matches_finite_differences():
    x, W, U, b, v, mW, mU, G = _case()
    rng = np.random.RandomState(1)
    for rev in (False, True):
        def loss():
            return float((ol.lstm_cell_forward(x, W, U, b, v, rev, mW, mU)[0] * G).sum())
        _, cache = ol.lstm_cell_forward(x, W, U, b, v, rev, mW, mU)
        dx, gp, _ = ol.lstm_cell_backward(G, cache)
        pairs = [(x, dx), (W, gp["W"]), (U, gp["U"]), (b, gp["b"]), (v["mi"][0], gp["mi_alpha"]), (v["mi"][1], gp["mi_beta1"]),
                 (v["mi"][2], gp["mi_beta2"])] + [(v["ln"][k][i], gp[f"ln_{n}_{k}"]) for k in ("uh", "wx", "c")
                                                  for i, n in ((0, "gain"), (1, "bias"))]
        for arr, g in pairs:
            for _ in range(5):
                idx = tuple(rng.randint(0, s) for s in arr.shape)
                old, eps = arr[idx], 1e-6
                arr[idx] = old + eps
                lp = loss()
                arr[idx] = old - eps
                lm = loss()
                arr[idx] = old
                assert abs((lp - lm) / (2 * eps) - g[idx]) < 1e-7


def _torch_step_layer(x, W, U, b, v, rev, mW, mU):
    """independent transcription of core/layers.py:432-469 in torch (autograd provides the gradients)."""
    N, T, D = x.shape
    H = U.shape[0]

    def ln(t, g, bb):
        mean = t.mean(1, keepdim=True)
        var = ((t - mean) ** 2).mean(1, keepdim=True)
        return (t - mean) / torch.sqrt(var + v["eps"]) * g + bb

    hs = lambda z: torch.clamp(0.2 * z + 0.5, 0.0, 1.0)
    h = torch.zeros(N, H, dtype=torch.float64)
    c = torch.zeros(N, H, dtype=torch.float64)
    outs = [None] * T
    for t in (range(T - 1, -1, -1) if rev else range(T)):
        Uh = ln((h * mU) @ U, *v["t_ln"]["uh"])
        Wx = ln((x[:, t] * mW) @ W, *v["t_ln"]["wx"])
        a, b1, b2 = v["t_mi"]
        z = a * Wx * Uh + b1 * Uh + b2 * Wx + b
        i, f, o = hs(z[:, :H]), hs(z[:, H:2 * H]), hs(z[:, 3 * H:])
        cn = f * c + i * torch.tanh(z[:, 2 * H:3 * H])
        c = c + torch.as_tensor(v["zmask_c"][t]) * (cn - c)
        hn = o * torch.tanh(ln(c, *v["t_ln"]["c"]))
        h = h + torch.as_tensor(v["zmask_h"][t]) * (hn - h)
        outs[t] = h
    return torch.stack(outs, 1)


def test_cell_matches_torch_autograd():
    x, W, U, b, v, mW, mU, G = _case(seed=3)
    leaf = lambda a: torch.tensor(a, dtype=torch.float64, requires_grad=True)
    for rev in (False, True):
        tx, tW, tU, tb = leaf(x), leaf(W), leaf(U), leaf(b)
        v["t_mi"] = tuple(leaf(m) for m in v["mi"])
        v["t_ln"] = {k: (leaf(g), leaf(bb)) for k, (g, bb) in v["ln"].items()}
        out = _torch_step_layer(tx, tW, tU, tb, v, rev, torch.as_tensor(mW), torch.as_tensor(mU))
        (out * torch.as_tensor(G)).sum().backward()
        ref, cache = ol.lstm_cell_forward(x, W, U, b, v, rev, mW, mU)
        np.testing.assert_allclose(ref, out.detach().numpy(), atol=1e-12)
        dx, gp, _ = ol.lstm_cell_backward(G, cache)
        np.testing.assert_allclose(dx, tx.grad.numpy(), atol=1e-10)
        np.testing.assert_allclose(gp["W"], tW.grad.numpy(), atol=1e-10)
        np.testing.assert_allclose(gp["U"], tU.grad.numpy(), atol=1e-10)
        np.testing.assert_allclose(gp["b"], tb.grad.numpy(), atol=1e-10)
        for i, n in enumerate(("mi_alpha", "mi_beta1", "mi_beta2")):
            np.testing.assert_allclose(gp[n], v["t_mi"][i].grad.numpy(), atol=1e-10)
        for k in ("uh", "wx", "c"):
            np.testing.assert_allclose(gp["ln_gain_" + k], v["t_ln"][k][0].grad.numpy(), atol=1e-10)
            np.testing.assert_allclose(gp["ln_bias_" + k], v["t_ln"][k][1].grad.numpy(), atol=1e-10)


def test_switches_off_equals_default_step_and_inference_zoneout_blend():
    x, W, U, b, v, mW, mU, G = _case(seed=5)
    H = U.shape[0]
    o1, _ = ol.lstm_cell_forward(x, W, U, b, ol.make_variant(H), False, mW, mU)
    o2, _ = ol.lstm_forward(x, W, U, b, False, mW, mU, dtype=np.float64)
    np.testing.assert_allclose(o1, o2, atol=1e-14)
    # inference: h = h_prev + (1 - level) * (h_new - h_prev)  (core/layers_utils.py:38-41, K.in_train_phase else-branch)
    vz = ol.make_variant(H, zoneout_h=0.25, zoneout_c=0.0)
    o3, _ = ol.lstm_cell_forward(x[:, :1], W, U, b, vz, False)
    o4, _ = ol.lstm_cell_forward(x[:, :1], W, U, b, ol.make_variant(H), False)
    np.testing.assert_allclose(o3, 0.75 * o4, atol=1e-14)


def test_whole_model_with_switches_matches_finite_differences():
    F, H, L, C, N, T = 5, 8, 2, 6, 3, 7
    rng = np.random.RandomState(0)
    p = om.init_variant_params(om.init_params(F, H, L, C, seed=1), F, H, L, layer_norm=(1.0, 0.0), mi=(1.0, 0.5, 0.5), residual="sum")
    p = {k: v.astype(np.float64) + (0.05 * rng.randn(*v.shape) if ("mi_" in k or "ln_" in k) else 0) for k, v in p.items()}
    x = rng.randn(N, T, F)
    labels = [np.array([1, 2]), np.array([3]), np.array([0, 4, 1])]
    kw = dict(masks={l: {k + d: (rng.rand(N, w) >= 0.2) / 0.8 for d in "fb" for k, w in (("W", 2 * H), ("U", H))} for l in range(L)},
              zoneout=0.2, zmasks={l: {k: (rng.rand(T, H) >= 0.2).astype(float) for k in ("hf", "hb", "cf", "cb")} for l in range(L)},
              residual="sum", input_mask=(rng.rand(N, T, 2 * H) >= 0.2) / 0.8)
    f = lambda: om.loss_and_grads_general(p, x, [T] * N, labels, weight_decay=1e-3, **kw)
    g = f()[2]
    for k in p:
        for _ in range(2):
            idx = tuple(rng.randint(0, s) for s in p[k].shape)
            old, eps = p[k][idx], 1e-6
            p[k][idx] = old + eps
            a = f()[0]
            p[k][idx] = old - eps
            b = f()[0]
            p[k][idx] = old
            assert abs((a - b) / (2 * eps) - g[k][idx]) < 1e-6, k
