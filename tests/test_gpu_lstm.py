"""K3/K4 parity: persistent BiLSTM forward / BPTT (C ABI) vs the oracle.

fp32 engine: 1e-5 (exact arithmetic up to summation order).
tensor-core engine (fp16 forward operands / bf16 backward operands): activations within 1e-3
norm-wise (north_star tolerance), gradients within 1e-2 norm-wise (bf16 operands).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import lstm as ol
from tests.util_gpu import dev, norm_err

pytestmark = pytest.mark.gpu


def _params(rng, D, H, scale=1.0):
    p = {}
    for d in "fb":
        W, U, b = ol.init_lstm(rng, D, H)
        p["W" + d], p["U" + d] = W * scale, U * scale
        p["b" + d] = b + 0.1 * rng.randn(4 * H).astype(np.float32)
    return p


def _device_fwd(p, x_ntd, training=True, engine=None):
    """x [N,T,D] -> runs zx GEMM in torch fp32 (operand prep is not under test here) + asr_lstm_forward."""
    from asr_study_b200._lib import LSTM_PIN_FP32, LstmFwdArgs, lib, ptr, cur_stream
    opts = LSTM_PIN_FP32 if engine == "fp32" else 0          # explicit engine flag of the C ABI (no environment switch)
    N, T, D = x_ntd.shape
    H = p["Uf"].shape[0]
    x = dev(x_ntd.transpose(1, 0, 2)).reshape(T * N, D)
    Wcat = dev(np.concatenate([p["Wf"], p["Wb"]], axis=1))
    torch.backends.cuda.matmul.allow_tf32 = False
    zx = (x @ Wcat).contiguous()
    bias = dev(np.concatenate([p["bf"], p["bb"]]))
    U = dev(np.stack([p["Uf"], p["Ub"]]))
    UT16 = dev(np.stack([p["Uf"].T, p["Ub"].T])).half().contiguous()
    R = T * N
    out = dict(h16=torch.empty(R, 2 * H, dtype=torch.float16, device="cuda"),
               h32=torch.empty(R, 2 * H, dtype=torch.float32, device="cuda"),
               hT16=torch.empty(2 * H, R, dtype=torch.bfloat16, device="cuda"),
               gates=torch.empty(R, 8 * H, dtype=torch.float32, device="cuda"),
               cell=torch.empty(R, 2 * H, dtype=torch.float32, device="cuda"))
    flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device="cuda")
    a = LstmFwdArgs(T=T, N=N, H=H, training=int(training), zx=ptr(zx).value, bias=ptr(bias).value,
                    U=ptr(U).value, U16=ptr(UT16).value, h16=ptr(out["h16"]).value, hT16=ptr(out["hT16"]).value,
                    h32=ptr(out["h32"]).value, gates=ptr(out["gates"]).value, cell=ptr(out["cell"]).value,
                    flags=ptr(flags).value, opts=opts)
    lib.asr_lstm_forward(C.byref(a), cur_stream())
    torch.cuda.synchronize()
    assert int(flags[64]) == 0, "persistent-kernel watchdog fired"
    return out, dict(U=U, flags=flags)


def _device_bwd(p, fwd, aux, dout_ntd, engine=None):
    from asr_study_b200._lib import LSTM_PIN_FP32, LstmBwdArgs, lib, ptr, cur_stream
    opts = LSTM_PIN_FP32 if engine == "fp32" else 0
    N, T, H2 = dout_ntd.shape
    H = H2 // 2
    R = T * N
    dh = dev(np.ascontiguousarray(dout_ntd.transpose(1, 0, 2))).reshape(R, H2)
    U16 = aux["U"].to(torch.bfloat16).contiguous()
    out = dict(dz16=torch.empty(R, 8 * H, dtype=torch.bfloat16, device="cuda"),
               dzT16=torch.empty(8 * H, R, dtype=torch.bfloat16, device="cuda"),
               dz32=torch.empty(R, 8 * H, dtype=torch.float32, device="cuda"),
               dbias=torch.zeros(8 * H, dtype=torch.float32, device="cuda"))
    a = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh).value, gates=ptr(fwd["gates"]).value, cell=ptr(fwd["cell"]).value,
                    U=ptr(aux["U"]).value, U16=ptr(U16).value, dz16=ptr(out["dz16"]).value,
                    dzT16=ptr(out["dzT16"]).value, dz32=ptr(out["dz32"]).value, dbias=ptr(out["dbias"]).value,
                    flags=ptr(aux["flags"]).value, opts=opts)
    lib.asr_lstm_backward(C.byref(a), cur_stream())
    torch.cuda.synchronize()
    assert int(aux["flags"][64]) == 0, "persistent-kernel watchdog fired"
    return out


def _oracle(p, x, dout):
    out, caches = ol.bilstm_forward(x, dict(Wf=p["Wf"], Uf=p["Uf"], bf=p["bf"], Wb=p["Wb"], Ub=p["Ub"], bb=p["bb"]),
                                    dtype=np.float64)
    H = p["Uf"].shape[0]
    _, _, _, dbf, dzf = ol.lstm_backward(dout[:, :, :H].astype(np.float64), caches[0])
    _, _, _, dbb, dzb = ol.lstm_backward(dout[:, :, H:].astype(np.float64), caches[1])
    return out, caches, np.concatenate([dzf, dzb], axis=2), np.concatenate([dbf, dbb])


SHAPES = [(8, 12, 5, 64), (16, 33, 26, 104), (32, 40, 26, 512), (5, 9, 7, 100)]


@pytest.mark.parametrize("N,T,D,H", SHAPES)
def test_fp32_engine_forward_backward_vs_oracle(N, T, D, H):
    rng = np.random.RandomState(N * 1000 + T)
    p = _params(rng, D, H, scale=2.0)
    x = rng.randn(N, T, D).astype(np.float32)
    dout = rng.randn(N, T, 2 * H).astype(np.float32)
    ref, caches, ref_dz, ref_db = _oracle(p, x, dout)
    fwd, aux = _device_fwd(p, x, engine="fp32")
    h = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h, ref) < 2e-5
    h16 = fwd["h16"].float().cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h16, ref) < 1e-3
    hT = fwd["hT16"].float().cpu().numpy().reshape(2 * H, T, N).transpose(2, 1, 0)
    assert norm_err(hT, ref) < 8e-3                                   # bf16 storage
    g = fwd["gates"].cpu().numpy().reshape(T, N, 2, 4 * H)
    for d in range(2):
        assert norm_err(g[:, :, d].transpose(1, 0, 2), caches[d]["gates"]) < 2e-5
    c = fwd["cell"].cpu().numpy().reshape(T, N, 2, H)
    for d in range(2):
        assert norm_err(c[:, :, d].transpose(1, 0, 2), caches[d]["cs"]) < 2e-5
    bwd = _device_bwd(p, fwd, aux, dout, engine="fp32")
    dz = bwd["dz32"].cpu().numpy().reshape(T, N, 8 * H).transpose(1, 0, 2)
    assert norm_err(dz, ref_dz) < 5e-5
    assert norm_err(bwd["dbias"].cpu().numpy(), ref_db) < 5e-5
    dz16 = bwd["dz16"].float().cpu().numpy().reshape(T, N, 8 * H).transpose(1, 0, 2)
    assert norm_err(dz16, ref_dz) < 8e-3
    dzT = bwd["dzT16"].float().cpu().numpy().reshape(8 * H, T, N).transpose(2, 1, 0)
    assert norm_err(dzT, ref_dz) < 8e-3


def test_reverse_direction_consumes_padding_first():
    """no-mask semantics: a zero-padded tail changes the reverse direction's outputs."""
    rng = np.random.RandomState(0)
    N, T, D, H = 8, 10, 6, 64
    p = _params(rng, D, H)
    p["bb"] = p["bb"] + 0.5
    x = rng.randn(N, T, D).astype(np.float32)
    x[:, 6:] = 0
    ref, *_ = _oracle(p, x, np.zeros((N, T, 2 * H), np.float32))
    fwd, _ = _device_fwd(p, x, engine="fp32")
    h = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h, ref) < 2e-5
    ref_trim, *_ = _oracle(p, x[:, :6], np.zeros((N, 6, 2 * H), np.float32))
    assert np.abs(ref[:, :6, H:] - ref_trim[:, :, H:]).max() > 1e-3


def test_full_length_sequence_T999_stability():
    """BASELINE size in time (T=999) on a narrower layer: drift of the fp32 engine vs the fp64 oracle."""
    rng = np.random.RandomState(1)
    N, T, D, H = 8, 999, 26, 128
    p = _params(rng, D, H, scale=1.5)
    x = rng.randn(N, T, D).astype(np.float32)
    ref, *_ = _oracle(p, x, np.zeros((N, T, 2 * H), np.float32))
    fwd, _ = _device_fwd(p, x, training=False, engine="fp32")
    h = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h, ref) < 1e-4


TC_SHAPES = [(16, 12, 26, 64), (16, 25, 20, 256), (32, 40, 26, 512), (48, 7, 26, 128), (24, 15, 26, 384), (64, 9, 26, 256),
             (96, 6, 26, 128), (128, 6, 26, 512), (160, 5, 26, 256)]     # the last two: several launches over batch groups


@pytest.mark.parametrize("N,T,D,H", TC_SHAPES)
def test_tensor_core_engine_forward_backward_vs_oracle(N, T, D, H):
    """tcgen05 persistent recurrence: fp16 forward operands (1e-3 activation bar), bf16 backward operands."""
    rng = np.random.RandomState(N * 100 + T)
    p = _params(rng, D, H, scale=1.5)
    x = rng.randn(N, T, D).astype(np.float32)
    dout = rng.randn(N, T, 2 * H).astype(np.float32)
    ref, caches, ref_dz, ref_db = _oracle(p, x, dout)
    fwd, aux = _device_fwd(p, x, engine="tc")
    h16 = fwd["h16"].float().cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h16, ref) < 1e-3, norm_err(h16, ref)
    h32 = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h32, ref) < 1e-3
    g = fwd["gates"].cpu().numpy().reshape(T, N, 2, 4 * H)
    c = fwd["cell"].cpu().numpy().reshape(T, N, 2, H)
    for d in range(2):
        assert norm_err(g[:, :, d].transpose(1, 0, 2), caches[d]["gates"]) < 1e-3
        assert norm_err(c[:, :, d].transpose(1, 0, 2), caches[d]["cs"]) < 1e-3
    hT = fwd["hT16"].float().cpu().numpy().reshape(2 * H, T, N).transpose(2, 1, 0)
    assert norm_err(hT, ref) < 8e-3
    # fp32 engine on the same inputs: the two device engines must agree to the same bar
    if N <= 32:
        fwd32, _ = _device_fwd(p, x, engine="fp32")
        assert norm_err(fwd["h32"].cpu().numpy(), fwd32["h32"].cpu().numpy()) < 1e-3
    bwd = _device_bwd(p, fwd, aux, dout, engine="tc")
    dz = bwd["dz32"].cpu().numpy().reshape(T, N, 8 * H).transpose(1, 0, 2)
    assert norm_err(dz, ref_dz) < 1e-2, norm_err(dz, ref_dz)
    assert norm_err(bwd["dbias"].cpu().numpy(), ref_db) < 1e-2
    dzT = bwd["dzT16"].float().cpu().numpy().reshape(8 * H, T, N).transpose(2, 1, 0)
    assert norm_err(dzT, ref_dz) < 1e-2


def test_tensor_core_engine_T999_drift():
    """T = 999 (BASELINE length), H = 512: drift of the fp16-operand recurrence vs the fp64 oracle."""
    rng = np.random.RandomState(5)
    N, T, D, H = 16, 999, 26, 512
    p = _params(rng, D, H)
    x = rng.randn(N, T, D).astype(np.float32)
    ref, *_ = _oracle(p, x, np.zeros((N, T, 2 * H), np.float32))
    fwd, _ = _device_fwd(p, x, training=False, engine="tc")
    h = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h, ref) < 1e-3, norm_err(h, ref)


def _fp16_storage_fwd_bwd(p, x_ntd, dout_ntd, with_masks=False, seed=0, opts=0):
    """The 16-bit-storage / TMA engine (csrc/lstm_tc4.cu) through the C ABI: zx16 in, gates16 / cell16 saved, every side
    output a TMA tile store; optional variational-dropout masks with the fused masked copies."""
    from asr_study_b200._lib import LstmBwdArgs, LstmFwdArgs, lib, ptr, cur_stream
    N, T, D = x_ntd.shape
    H = p["Uf"].shape[0]
    R = T * N
    assert lib.asr_lstm_fp16_storage(T, N, H, 0) == 1
    rng = np.random.RandomState(seed)
    mW = mU = mNext = None
    if with_masks:
        mW = {d: ((rng.rand(N, D) >= 0.2) / 0.8).astype(np.float32) for d in "fb"}
        mU = np.stack([((rng.rand(N, H) >= 0.2) / 0.8).astype(np.float32) for _ in "fb"])
        mNext = np.stack([((rng.rand(N, 2 * H) >= 0.2) / 0.8).astype(np.float32) for _ in "fb"])
    torch.backends.cuda.matmul.allow_tf32 = False
    xt = dev(x_ntd.transpose(1, 0, 2)).reshape(T * N, D)
    zs = []
    for d in "fb":
        xm = xt if mW is None else (xt.view(T, N, D) * dev(mW[d])[None]).reshape(R, D)
        zs.append(xm @ dev(p["W" + d]))
    zx16 = torch.stack(zs, dim=1).reshape(R, 8 * H).half().contiguous()          # [R, 2, 4H]
    bias = dev(np.concatenate([p["bf"], p["bb"]]))
    U = dev(np.stack([p["Uf"], p["Ub"]]))
    UT16 = dev(np.stack([p["Uf"].T, p["Ub"].T])).half().contiguous()
    f = dict(h16=torch.zeros(R, 2 * H, dtype=torch.float16, device="cuda"),
             hT16=torch.zeros(2 * H, R, dtype=torch.bfloat16, device="cuda"),
             gates16=torch.zeros(R, 8 * H, dtype=torch.float16, device="cuda"),
             cell16=torch.zeros(R, 2 * H, dtype=torch.float16, device="cuda"),
             hm16=torch.zeros(2, R, 2 * H, dtype=torch.float16, device="cuda"),
             hmT16=torch.zeros(2, 2 * H, R, dtype=torch.bfloat16, device="cuda"),
             hT16u=torch.zeros(2 * H, R, dtype=torch.bfloat16, device="cuda"))
    flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device="cuda")
    kw = {}
    if with_masks:
        mu_d, mn_d = dev(mU), dev(mNext)
        kw = dict(mask_u=ptr(mu_d).value, mask_next=ptr(mn_d).value, hm16=ptr(f["hm16"]).value, hmT16=ptr(f["hmT16"]).value,
                  hT16u=ptr(f["hT16u"]).value)
    a = LstmFwdArgs(T=T, N=N, H=H, training=1, zx16=ptr(zx16).value, bias=ptr(bias).value, U=ptr(U).value,
                    U16=ptr(UT16).value, h16=ptr(f["h16"]).value, hT16=ptr(f["hT16"]).value,
                    gates16=ptr(f["gates16"]).value, cell16=ptr(f["cell16"]).value, flags=ptr(flags).value, opts=opts, **kw)
    lib.asr_lstm_forward(C.byref(a), cur_stream())
    torch.cuda.synchronize()
    assert int(flags[64]) == 0, "persistent-kernel watchdog fired"
    dh = dev(np.ascontiguousarray(dout_ntd.transpose(1, 0, 2))).reshape(R, 2 * H)
    b = dict(dz16=torch.zeros(R, 8 * H, dtype=torch.bfloat16, device="cuda"),
             dzT16=torch.zeros(8 * H, R, dtype=torch.bfloat16, device="cuda"),
             dbias=torch.zeros(8 * H, dtype=torch.float32, device="cuda"))
    U16 = U.to(torch.bfloat16).contiguous()
    ba = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh).value, gates16=ptr(f["gates16"]).value, cell16=ptr(f["cell16"]).value,
                     U=ptr(U).value, U16=ptr(U16).value, dz16=ptr(b["dz16"]).value, dzT16=ptr(b["dzT16"]).value,
                     dbias=ptr(b["dbias"]).value, flags=ptr(flags).value, mask_u=kw.get("mask_u"), opts=opts)
    lib.asr_lstm_backward(C.byref(ba), cur_stream())
    torch.cuda.synchronize()
    assert int(flags[64]) == 0, "persistent-kernel watchdog fired"
    return f, b, dict(mW=mW, mU=mU, mNext=mNext)


@pytest.mark.parametrize("masks", [False, True])
@pytest.mark.parametrize("N,T,D,H", [(32, 60, 26, 512), (16, 33, 26, 256), (16, 21, 26, 128), (8, 7, 12, 384), (64, 9, 26, 512)])
def test_fp16_storage_tma_engine_vs_oracle(masks, N, T, D, H):
    """csrc/lstm_tc4.cu: activations 1e-3 norm-wise, saved gates / cell 1e-3 (fp16 storage), dz 1e-2, every TMA-stored
    layout (row-major fp16, the two B_W-masked copies, the transposed bf16 copies, dz and dz^T) in the right place."""
    rng = np.random.RandomState(H + T + N)
    p = _params(rng, D, H, scale=1.5)
    x = rng.randn(N, T, D).astype(np.float32)
    dout = (rng.randn(N, T, 2 * H) * 0.1).astype(np.float32)
    f, b, m = _fp16_storage_fwd_bwd(p, x, dout, with_masks=masks, seed=N)
    lp = dict(Wf=p["Wf"], Uf=p["Uf"], bf=p["bf"], Wb=p["Wb"], Ub=p["Ub"], bb=p["bb"])
    om_ = None if not masks else dict(Wf=m["mW"]["f"], Wb=m["mW"]["b"], Uf=m["mU"][0], Ub=m["mU"][1])
    ref, caches = ol.bilstm_forward(x, lp, om_, dtype=np.float64)
    h = f["h16"].float().cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h, ref) < 1e-3, norm_err(h, ref)
    g = f["gates16"].float().cpu().numpy().reshape(T, N, 2, 4 * H)
    c = f["cell16"].float().cpu().numpy().reshape(T, N, 2, H)
    for d in range(2):
        assert norm_err(g[:, :, d].transpose(1, 0, 2), caches[d]["gates"]) < 1e-3
        assert norm_err(c[:, :, d].transpose(1, 0, 2), caches[d]["cs"]) < 1e-3
    hT = f["hT16"].float().cpu().numpy().reshape(2 * H, T, N).transpose(2, 1, 0)
    refU = ref if not masks else ref * np.concatenate([m["mU"][0], m["mU"][1]], axis=1)[:, None, :]
    assert norm_err(hT, refU) < 8e-3
    if masks:
        hm = f["hm16"].float().cpu().numpy().reshape(2, T, N, 2 * H).transpose(0, 2, 1, 3)
        hmT = f["hmT16"].float().cpu().numpy().reshape(2, 2 * H, T, N).transpose(0, 3, 2, 1)
        hTu = f["hT16u"].float().cpu().numpy().reshape(2 * H, T, N).transpose(2, 1, 0)
        for i in range(2):
            assert norm_err(hm[i], ref * m["mNext"][i][:, None, :]) < 1e-3
            assert norm_err(hmT[i], ref * m["mNext"][i][:, None, :]) < 8e-3
        assert norm_err(hTu, ref) < 8e-3
    dzs, dbs = [], []
    for d in range(2):
        _, _, _, db, dz = ol.lstm_backward(dout[:, :, d * H:(d + 1) * H].astype(np.float64), caches[d])
        dzs.append(dz)
        dbs.append(db)
    ref_dz = np.concatenate(dzs, axis=2)
    dz = b["dz16"].float().cpu().numpy().reshape(T, N, 8 * H).transpose(1, 0, 2)
    assert norm_err(dz, ref_dz) < 1e-2, norm_err(dz, ref_dz)
    dzT = b["dzT16"].float().cpu().numpy().reshape(8 * H, T, N).transpose(2, 1, 0)
    assert norm_err(dzT, ref_dz) < 1e-2
    assert norm_err(b["dbias"].cpu().numpy(), np.concatenate(dbs)) < 1e-2
