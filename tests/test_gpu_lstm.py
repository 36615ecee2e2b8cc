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
    from asr_study_b200._lib import LstmFwdArgs, lib, ptr, cur_stream
    os.environ.pop("ASR_B200_LSTM", None)
    if engine in ("fp32", "tc1", "tc3"):
        os.environ["ASR_B200_LSTM"] = engine
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
                    flags=ptr(flags).value)
    lib.asr_lstm_forward(C.byref(a), cur_stream())
    torch.cuda.synchronize()
    assert int(flags[64]) == 0, "persistent-kernel watchdog fired"
    os.environ.pop("ASR_B200_LSTM", None)
    return out, dict(U=U, flags=flags)


def _device_bwd(p, fwd, aux, dout_ntd, engine=None):
    from asr_study_b200._lib import LstmBwdArgs, lib, ptr, cur_stream
    os.environ.pop("ASR_B200_LSTM", None)
    if engine == "fp32":
        os.environ["ASR_B200_LSTM"] = "fp32"
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
                    flags=ptr(aux["flags"]).value)
    lib.asr_lstm_backward(C.byref(a), cur_stream())
    torch.cuda.synchronize()
    assert int(aux["flags"][64]) == 0, "persistent-kernel watchdog fired"
    os.environ.pop("ASR_B200_LSTM", None)
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


@pytest.mark.parametrize("engine", ["tc3", "tc"])
@pytest.mark.parametrize("N,T,D,H", [(32, 60, 26, 512), (16, 33, 26, 256), (16, 20, 26, 128)])
def test_forward_engines_agree_and_match_oracle(engine, N, T, D, H):
    """LL-ring engine (default 'tc') and cluster/DSMEM engine (tc3) against the oracle on the same inputs."""
    rng = np.random.RandomState(H + T)
    p = _params(rng, D, H, scale=1.5)
    x = rng.randn(N, T, D).astype(np.float32)
    ref, caches, _, _ = _oracle(p, x, np.zeros((N, T, 2 * H), np.float32))
    fwd, _ = _device_fwd(p, x, engine=engine)
    h32 = fwd["h32"].cpu().numpy().reshape(T, N, 2 * H).transpose(1, 0, 2)
    assert norm_err(h32, ref) < 1e-3, norm_err(h32, ref)
    c = fwd["cell"].cpu().numpy().reshape(T, N, 2, H)
    for d in range(2):
        assert norm_err(c[:, :, d].transpose(1, 0, 2), caches[d]["cs"]) < 1e-3
    hT = fwd["hT16"].float().cpu().numpy().reshape(2 * H, T, N).transpose(2, 1, 0)
    assert norm_err(hT, ref) < 8e-3
