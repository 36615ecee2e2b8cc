"""K3/K4 (general cell) parity: LSTM.step with layer-norm / multiplicative integration / zoneout
(core/layers.py:432-469, core/layers_utils.py:16-51) through the C ABI vs the fp64 oracle.

fp32 kernel vs fp64 oracle: activations 2e-5 norm-wise, gradients 2e-4 norm-wise (the LN backward
subtracts means of O(1) terms).  Switch combinations, ragged shapes (N not a multiple of the
4-sample CTA group, 4H not a multiple of the 256-thread column stride) and the dropout masks
are covered; the no-switch configuration must reproduce the default engine's oracle.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import lstm as ol
from tests.util_gpu import dev, norm_err

pytestmark = pytest.mark.gpu


def _setup(seed, N, T, D, H, mi, ln, zone, train_masks, dropout):
    rng = np.random.RandomState(seed)
    x = rng.randn(N, T, D)
    p, var = {}, {}
    for d in "fb":
        W, U, b = ol.init_lstm(rng, D, H)
        p["W" + d], p["U" + d] = W.astype(np.float64), U.astype(np.float64)
        p["b" + d] = b.astype(np.float64) + 0.1 * rng.randn(4 * H)
        kw = {}
        if mi:
            kw["mi"] = (1.0 + 0.2 * rng.randn(4 * H), 0.5 + 0.2 * rng.randn(4 * H), 0.5 + 0.2 * rng.randn(4 * H))
        if ln:
            kw["layer_norm"] = {"uh": (1 + 0.2 * rng.randn(4 * H), 0.1 * rng.randn(4 * H)),
                                "wx": (1 + 0.2 * rng.randn(4 * H), 0.1 * rng.randn(4 * H)),
                                "c": (1 + 0.2 * rng.randn(H), 0.1 * rng.randn(H))}
        if zone:
            kw.update(zoneout_h=zone, zoneout_c=zone)
            if train_masks:
                kw.update(zmask_h=(rng.rand(T, H) >= zone).astype(np.float64),
                          zmask_c=(rng.rand(T, H) >= zone).astype(np.float64))
        var[d] = ol.make_variant(H, **kw)
    masks = None
    if dropout:
        masks = {k + d: (rng.rand(N, w) >= dropout) / (1.0 - dropout) for d in "fb" for k, w in (("W", D), ("U", H))}
    return x, p, var, masks


def _pack2(f, b):
    return dev(np.stack([f, b]).astype(np.float32))


def _run_device(x, p, var, masks, dh):
    from asr_study_b200._lib import (LstmBwdArgs, LstmFwdArgs, LstmVariant, LstmVariantGrads, cur_stream, lib, ptr)
    N, T, D = x.shape
    H = p["Uf"].shape[0]
    R = T * N
    m = masks or {}
    zx = np.zeros((T, N, 2, 4 * H))
    for i, d in enumerate("fb"):
        xm = x if masks is None else x * m["W" + d][:, None, :]
        zx[:, :, i] = (xm.reshape(N * T, D) @ p["W" + d]).reshape(N, T, 4 * H).transpose(1, 0, 2)
    zx_d = dev(zx.astype(np.float32))
    bias, U = _pack2(p["bf"], p["bb"]), _pack2(p["Uf"], p["Ub"])
    mask_u = _pack2(m["Uf"], m["Ub"]) if masks is not None else None
    vf, vb = var["f"], var["b"]
    keep = []

    def vec(get):
        a, b = get(vf), get(vb)
        if a is None:
            return None
        t = _pack2(a, b)
        keep.append(t)
        return t.data_ptr()

    v = LstmVariant(mi_alpha=vec(lambda q: q["mi"] and q["mi"][0]), mi_beta1=vec(lambda q: q["mi"] and q["mi"][1]),
                    mi_beta2=vec(lambda q: q["mi"] and q["mi"][2]),
                    ln_gain_uh=vec(lambda q: q["ln"] and q["ln"]["uh"][0]), ln_bias_uh=vec(lambda q: q["ln"] and q["ln"]["uh"][1]),
                    ln_gain_wx=vec(lambda q: q["ln"] and q["ln"]["wx"][0]), ln_bias_wx=vec(lambda q: q["ln"] and q["ln"]["wx"][1]),
                    ln_gain_c=vec(lambda q: q["ln"] and q["ln"]["c"][0]), ln_bias_c=vec(lambda q: q["ln"] and q["ln"]["c"][1]),
                    ln_eps=vf["eps"], zoneout_h=vf["zoneout_h"], zoneout_c=vf["zoneout_c"],
                    zmask_h=vec(lambda q: q["zmask_h"]), zmask_c=vec(lambda q: q["zmask_c"]))
    f32 = dict(dtype=torch.float32, device="cuda")
    h32, gates, cell = torch.empty(R, 2 * H, **f32), torch.empty(R, 8 * H, **f32), torch.empty(R, 2 * H, **f32)
    h16 = torch.empty(R, 2 * H, dtype=torch.float16, device="cuda")
    hT16 = torch.empty(2 * H, R, dtype=torch.bfloat16, device="cuda")
    uh_raw = torch.empty(R, 8 * H, **f32)
    flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device="cuda")
    a = LstmFwdArgs(T=T, N=N, H=H, training=1, zx=ptr(zx_d).value, bias=ptr(bias).value, U=ptr(U).value, U16=None,
                    h16=ptr(h16).value, hT16=ptr(hT16).value, h32=ptr(h32).value, gates=ptr(gates).value,
                    cell=ptr(cell).value, flags=ptr(flags).value, mask_u=ptr(mask_u).value if mask_u is not None else None)
    lib.asr_lstm_cell_forward(C.byref(a), C.byref(v), ptr(uh_raw), cur_stream())
    dz32, duh, dbias = torch.empty(R, 8 * H, **f32), torch.empty(R, 8 * H, **f32), torch.empty(8 * H, **f32)
    g = {k: (torch.empty(2 * (H if k.endswith("_c") else 4 * H), **f32)) for k, _ in LstmVariantGrads._fields_}
    gs = LstmVariantGrads(**{k: t.data_ptr() for k, t in g.items()})
    dh_d = dev(dh.transpose(1, 0, 2).astype(np.float32))
    b = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh_d).value, gates=ptr(gates).value, cell=ptr(cell).value, U=ptr(U).value,
                    U16=None, dz16=None, dzT16=None, dz32=ptr(dz32).value, dbias=ptr(dbias).value, flags=ptr(flags).value,
                    mask_u=ptr(mask_u).value if mask_u is not None else None)
    lib.asr_lstm_cell_backward(C.byref(b), C.byref(v), ptr(zx_d), ptr(uh_raw), ptr(duh), C.byref(gs), cur_stream())
    torch.cuda.synchronize()
    out = dict(h=h32.cpu().numpy().reshape(T, N, 2, H), h16=h16.float().cpu().numpy().reshape(T, N, 2, H),
               hT=hT16.float().cpu().numpy().reshape(2, H, T, N),
               dwx=dz32.cpu().numpy().reshape(T, N, 2, 4 * H), duh=duh.cpu().numpy().reshape(T, N, 2, 4 * H),
               dbias=dbias.cpu().numpy().reshape(2, 4 * H))
    for k, t in g.items():
        out["g_" + k] = t.cpu().numpy().reshape(2, -1)
    return out


CASES = [
    # N, T, D, H, mi, ln, zoneout, train masks, dropout
    (5, 7, 6, 24, True, False, 0.0, False, 0.0),
    (5, 7, 6, 24, False, True, 0.0, False, 0.0),
    (5, 7, 6, 24, False, False, 0.3, True, 0.0),
    (5, 7, 6, 24, False, False, 0.3, False, 0.0),       # inference blend (1 - level)
    (6, 9, 5, 100, True, True, 0.25, True, 0.2),        # everything on, 4H = 400 (ragged column stride), dropout masks
    (3, 5, 4, 24, False, False, 0.0, False, 0.0),       # no switch: must equal the default step
    (9, 12, 7, 288, True, True, 0.1, True, 0.0),        # 4H = 1152: 5 column slots, 2 unit slots
]


@pytest.mark.parametrize("case", CASES)
def test_cell_forward_backward_vs_oracle(case):
    N, T, D, H, mi, ln, zone, train_masks, dropout = case
    x, p, var, masks = _setup(11, N, T, D, H, mi, ln, zone, train_masks, dropout)
    rng = np.random.RandomState(5)
    dh = rng.randn(N, T, 2 * H)
    m = masks or {}
    got = _run_device(x, p, var, masks, dh)
    for i, d in enumerate("fb"):
        out, cache = ol.lstm_cell_forward(x, p["W" + d], p["U" + d], p["b" + d], var[d], reverse=(d == "b"),
                                          mask_W=m.get("W" + d), mask_U=m.get("U" + d))
        assert norm_err(got["h"][:, :, i].transpose(1, 0, 2), out) < 2e-5
        assert norm_err(got["h16"][:, :, i].transpose(1, 0, 2), out) < 2e-3
        hm = out if masks is None else out * m["U" + d][:, None, :]
        assert norm_err(got["hT"][i].transpose(2, 1, 0), hm) < 1e-2          # bf16 transposed copy of h * B_U
        _, gp, (dwx, duh) = ol.lstm_cell_backward(dh[:, :, i * H:(i + 1) * H], cache)
        assert norm_err(got["dwx"][:, :, i].transpose(1, 0, 2), dwx) < 2e-4
        assert norm_err(got["duh"][:, :, i].transpose(1, 0, 2), duh) < 2e-4
        assert norm_err(got["dbias"][i], gp["b"]) < 2e-4
        for k in ("mi_alpha", "mi_beta1", "mi_beta2", "ln_gain_uh", "ln_bias_uh", "ln_gain_wx", "ln_bias_wx",
                  "ln_gain_c", "ln_bias_c"):
            if k in gp:
                assert norm_err(got["g_" + k][i], gp[k]) < 3e-4, k
    if not (mi or ln or zone):
        ref, _ = ol.bilstm_forward(x, {k: v for k, v in p.items()}, None, np.float64)
        assert norm_err(np.concatenate([got["h"][:, :, 0], got["h"][:, :, 1]], axis=2).transpose(1, 0, 2), ref) < 2e-5


def test_cell_rejects_partial_groups():
    from asr_study_b200._lib import AsrError, LstmFwdArgs, LstmVariant, cur_stream, lib, ptr
    t = torch.zeros(64, device="cuda")
    a = LstmFwdArgs(T=1, N=1, H=4, training=0, zx=ptr(t).value, bias=ptr(t).value, U=ptr(t).value, U16=None, h16=None,
                    hT16=None, h32=ptr(t).value, gates=None, cell=None, flags=ptr(t).value, mask_u=None)
    v = LstmVariant(mi_alpha=t.data_ptr(), ln_eps=1e-5)          # beta1 / beta2 missing
    with pytest.raises(AsrError):
        lib.asr_lstm_cell_forward(C.byref(a), C.byref(v), None, cur_stream())
