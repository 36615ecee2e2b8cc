"""Oracle (BiLSTM / model / optimiser) self-consistency.  CPU only."""
import numpy as np
import torch

from oracle import lstm as ol
from oracle import model as om


def test_hard_sigmoid_and_gate_order():
    assert np.allclose(ol.hard_sigmoid(np.array([-3., -2.5, 0., 2.5, 3.])), [0, 0, .5, 1, 1])
    # hand-rolled single step, gate order i,f,c,o (core/layers.py:447-450)
    rng = np.random.RandomState(0)
    D, H = 3, 2
    W, U, b = rng.randn(D, 4 * H), rng.randn(H, 4 * H), rng.randn(4 * H)
    x = rng.randn(1, 1, D)
    out, _ = ol.lstm_forward(x, W, U, b, dtype=np.float64)
    z = x[0, 0] @ W + b
    i, f, g, o = (ol.hard_sigmoid(z[:H]), ol.hard_sigmoid(z[H:2*H]), np.tanh(z[2*H:3*H]), ol.hard_sigmoid(z[3*H:]))
    np.testing.assert_allclose(out[0, 0], o * np.tanh(i * g), atol=1e-12)


def test_reverse_direction_runs_over_padding_first():
    rng = np.random.RandomState(1)
    D, H, T = 4, 3, 6
    W, U, b = ol.init_lstm(rng, D, H)
    b = b + 0.5 * rng.randn(4 * H).astype(np.float32)   # trained (non-zero) cell bias
    x = np.zeros((1, T, D), np.float32)
    x[0, :3] = rng.randn(3, D)
    out_pad, _ = ol.lstm_forward(x, W, U, b, reverse=True)
    out_trim, _ = ol.lstm_forward(x[:, :3], W, U, b, reverse=True)
    # no masking: state evolves on the zero padding (bias-driven) -> differs
    assert np.abs(out_pad[0, :3] - out_trim[0]).max() > 1e-4


def _fd_check(masks):
    rng = np.random.RandomState(2)
    N, T, D, H = 2, 5, 3, 4
    p = dict(zip(("Wf", "Uf", "bf"), ol.init_lstm(rng, D, H)))
    p.update(zip(("Wb", "Ub", "bb"), ol.init_lstm(rng, D, H)))
    p = {k: (v.astype(np.float64) + 0.3 * rng.randn(*v.shape)) for k, v in p.items()}
    x = rng.randn(N, T, D)
    r = rng.randn(N, T, 2 * H)

    def f(pp, xx):
        out, c = ol.bilstm_forward(xx, pp, masks, dtype=np.float64)
        return float((out * r).sum()), c

    _, caches = f(p, x)
    dx, g = ol.bilstm_backward(r, caches)
    eps = 1e-6
    for k in p:
        idx = tuple(rng.randint(0, s) for s in p[k].shape)
        pp = {kk: vv.copy() for kk, vv in p.items()}
        pp[k][idx] += eps
        pm = {kk: vv.copy() for kk, vv in p.items()}
        pm[k][idx] -= eps
        fd = (f(pp, x)[0] - f(pm, x)[0]) / (2 * eps)
        assert abs(fd - g[k][idx]) < 1e-6 * max(1, abs(fd)), (k, fd, g[k][idx])
    xp, xm = x.copy(), x.copy()
    xp[1, 2, 0] += eps
    xm[1, 2, 0] -= eps
    assert abs((f(p, xp)[0] - f(p, xm)[0]) / (2 * eps) - dx[1, 2, 0]) < 1e-6


def test_bptt_matches_finite_differences():
    _fd_check(None)


def test_bptt_with_variational_dropout_masks():
    rng = np.random.RandomState(9)
    m = {k: (rng.rand(2, n) > 0.2) / 0.8 for k, n in (("Wf", 3), ("Uf", 4), ("Wb", 3), ("Ub", 4))}
    _fd_check(m)


def test_full_model_grad_and_adam_vs_torch():
    rng = np.random.RandomState(3)
    N, T, F, H, C = 2, 9, 5, 4, 6
    params = om.init_params(F, H, 2, C, seed=1)
    x = rng.randn(N, T, F).astype(np.float32)
    lens = np.array([9, 7])
    labels = [np.array([1, 2, 2]), np.array([0, 4])]
    total, ctc, grads, logits = om.loss_and_grads(params, x, lens, labels, weight_decay=1e-4,
                                                  dtype=np.float64)
    assert logits.shape == (N, T, C) and ctc.shape == (N,)
    # finite-difference one weight per tensor against total loss
    eps = 1e-5
    for k in params:
        idx = tuple(rng.randint(0, s) for s in params[k].shape)
        pp = {kk: vv.astype(np.float64).copy() for kk, vv in params.items()}
        pm = {kk: vv.astype(np.float64).copy() for kk, vv in params.items()}
        pp[k][idx] += eps
        pm[k][idx] -= eps
        fd = (om.loss_and_grads(pp, x, lens, labels, 1e-4, dtype=np.float64)[0] -
              om.loss_and_grads(pm, x, lens, labels, 1e-4, dtype=np.float64)[0]) / (2 * eps)
        assert abs(fd - grads[k][idx]) < 2e-5 * max(1.0, abs(fd)), (k, fd, grads[k][idx])
    # Adam (no clipping active) == torch.optim.Adam; Keras' lr_t form is algebraically
    # the same up to eps placement: compare with eps tiny relative to sqrt(v)
    p1 = {k: v.copy() for k, v in params.items()}
    st = {}
    tp = {k: torch.tensor(v.copy(), requires_grad=True) for k, v in params.items()}
    opt = torch.optim.Adam(list(tp.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8)
    for _ in range(3):
        n = om.clip_adam_step(p1, grads, st, clipnorm=400.0)
        for k in tp:
            tp[k].grad = torch.tensor(grads[k].astype(np.float32))
        opt.step()
    assert n < 400
    for k in p1:
        big = np.abs(grads[k]) > 1e-4        # eps placement differs (Keras: sqrt(v)+eps un-corrected)
        np.testing.assert_allclose(p1[k][big], tp[k].detach().numpy()[big], atol=3e-5)


def test_clipnorm_scales_globally():
    g = {"a": np.full(4, 300.0, np.float32), "b": np.full(4, 400.0, np.float32)}
    p = {"a": np.zeros(4, np.float32), "b": np.zeros(4, np.float32)}
    st = {}
    n = om.clip_adam_step(p, g, st, clipnorm=400.0)
    assert abs(n - 1000.0) < 1e-3
    np.testing.assert_allclose(st["m"]["a"], 0.1 * 300.0 * 0.4, rtol=1e-6)


def test_recurrence_structure_matches_torch_lstm_when_the_inner_activation_is_swapped(monkeypatch):
    """Independent pin of everything in the BiLSTM restatement EXCEPT the inner activation: with hard_sigmoid swapped
    for the logistic function (and its derivative), Bidirectional(LSTM) must be torch.nn.LSTM(bidirectional=True) —
    same gate order (i, f, candidate, o: Keras-1 W/U column blocks = torch's weight row blocks), bias placement,
    zero initial state, reverse direction run from the last frame and written back in time order, [fwd | bwd]
    concatenation — forward outputs and every parameter / input gradient.  hard_sigmoid itself is pinned on its
    definition (Keras-1 / Theano: clip(0.2 x + 0.5, 0, 1)) in test_hard_sigmoid_and_gate_order."""
    import torch
    monkeypatch.setattr(ol, "hard_sigmoid", lambda z: 1.0 / (1.0 + np.exp(-z)))
    monkeypatch.setattr(ol, "_dhs", lambda a: a * (1.0 - a))
    rng = np.random.RandomState(11)
    N, T, D, H = 3, 9, 5, 7
    p = {}
    for d in "fb":
        W, U, b = ol.init_lstm(rng, D, H)
        p["W" + d], p["U" + d], p["b" + d] = W.astype(np.float64), U.astype(np.float64), (b + 0.1 * rng.randn(4 * H))
    x = rng.randn(N, T, D)
    dout = rng.randn(N, T, 2 * H)
    out, caches = ol.bilstm_forward(x, p, dtype=np.float64)
    dx, grads = ol.bilstm_backward(dout, caches)

    net = torch.nn.LSTM(D, H, batch_first=True, bidirectional=True).double()
    with torch.no_grad():
        for d, suf in (("f", ""), ("b", "_reverse")):
            getattr(net, "weight_ih_l0" + suf).copy_(torch.tensor(p["W" + d].T))
            getattr(net, "weight_hh_l0" + suf).copy_(torch.tensor(p["U" + d].T))
            getattr(net, "bias_ih_l0" + suf).copy_(torch.tensor(p["b" + d]))
            getattr(net, "bias_hh_l0" + suf).zero_()
    xt = torch.tensor(x, requires_grad=True)
    yt, _ = net(xt)
    np.testing.assert_allclose(out, yt.detach().numpy(), atol=1e-12)
    (yt * torch.tensor(dout)).sum().backward()
    np.testing.assert_allclose(dx, xt.grad.numpy(), atol=1e-11)
    for d, suf in (("f", ""), ("b", "_reverse")):
        np.testing.assert_allclose(grads["W" + d], getattr(net, "weight_ih_l0" + suf).grad.numpy().T, atol=1e-11)
        np.testing.assert_allclose(grads["U" + d], getattr(net, "weight_hh_l0" + suf).grad.numpy().T, atol=1e-11)
        np.testing.assert_allclose(grads["b" + d], getattr(net, "bias_ih_l0" + suf).grad.numpy(), atol=1e-11)
