"""The convolutional front-end oracle (oracle/conv.py; BASELINE configs[3], not in the reference) against an independent
implementation: torch.nn.functional.conv2d forward and autograd, fp64."""
import numpy as np
import torch

from oracle import conv as ocv


def _torch_front(params, x, layers, clip):
    a = x[:, None]                                            # [N, 1, T, F]
    for i, (co, kt, kf, st, sf) in enumerate(layers):
        W = params[f"conv{i}.W"]
        ci = a.shape[1]
        Wt = W.reshape(co, kt, kf, ci).permute(0, 3, 1, 2)    # [C_out, C_in, kt, kf]
        a = torch.nn.functional.conv2d(a, Wt, params[f"conv{i}.b"], stride=(st, sf), padding=((kt - 1) // 2, (kf - 1) // 2))
        a = torch.clamp(a, 0.0, clip)
    N, C, To, Fo = a.shape
    return a.permute(0, 2, 3, 1).reshape(N, To, Fo * C)        # channel fastest


def test_front_forward_backward_match_torch_conv2d():
    rng = np.random.RandomState(0)
    layers = ((4, 5, 7, 2, 2), (3, 3, 5, 1, 2))
    N, T, F = 2, 23, 17
    p = ocv.init_front(rng, F, layers)
    for k in p:
        p[k] = (p[k] + 0.3 * rng.randn(*p[k].shape)).astype(np.float64)
    x = rng.randn(N, T, F) * 2.0
    clip = 1.5
    y, cache = ocv.front_forward(p, x, layers, clip)
    tp = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in p.items()}
    tx = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    ty = _torch_front(tp, tx, layers, clip)
    np.testing.assert_allclose(y, ty.detach().numpy(), atol=1e-12)
    assert y.shape[1] == ocv.front_out_lengths([T], layers)[0]
    dout = rng.randn(*y.shape)
    (ty * torch.tensor(dout)).sum().backward()
    grads, dx = ocv.front_backward(p, dout, cache, layers, clip)
    for k in p:
        np.testing.assert_allclose(grads[k], tp[k].grad.numpy(), atol=1e-11, err_msg=k)
    np.testing.assert_allclose(dx, tx.grad.numpy(), atol=1e-11)


def test_ds2_geometry_on_the_config4_input():
    # 40 log-mel bins, 999 frames -> 32 x 20 x 500 -> 32 x 10 x 500 -> 320 features x 500 frames
    assert ocv.front_out_lengths([999, 500, 11]).tolist() == [500, 250, 6]
    assert ocv.out_len(40, 41, 2) == 20 and ocv.out_len(20, 21, 2) == 10
