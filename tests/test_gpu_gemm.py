"""K2/K5: TN GEMM (C ABI) vs torch fp32 matmul on the same 16-bit-rounded operands."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(128, 128, 64), (31968, 4096, 32), (4000, 2048, 1024), (26, 2048, 31968), (512, 2048, 4000),
          (1000, 28, 1024), (777, 1024, 32), (130, 72, 200),
          # the C2 projections on the persistent 128x256 engine (gemm_tc2.cu): dW / dU (split-K), dX, ragged M and N
          (1024, 2048, 8000), (512, 2048, 7984), (7992, 1024, 4096), (300, 320, 12000)]


def _gemm(din, dout, A, B, bias=None, alpha=1.0, acc=None, M=None, N=None, K=None, lda=None, ldb=None, flags=0):
    from asr_study_b200._lib import lib, ptr, cur_stream
    M = M or A.shape[0]
    N = N or B.shape[0]
    K = K or A.shape[1]
    odt = {0: torch.float32, 1: torch.float16, 2: torch.bfloat16}[dout]
    Cm = acc.clone() if acc is not None else torch.empty(M, N, dtype=odt, device="cuda")
    lib.asr_gemm_tn_ex(din, dout, M, N, K, ptr(A), lda or A.stride(0), ptr(B), ldb or B.stride(0), ptr(Cm), N, ptr(bias),
                       alpha, int(acc is not None), flags, cur_stream())
    torch.cuda.synchronize()
    return Cm


@pytest.mark.parametrize("flags", [0, 2, 1])        # default (persistent 128x256 where it fits), ASR_GEMM_TILE128, ASR_GEMM_BACKGROUND
@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("din", [0, 1])
def test_gemm_matches_fp32_matmul(flags, M, N, K, din):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dt = torch.float16 if din == 0 else torch.bfloat16
    A = torch.randn(M, K, device="cuda", generator=g).to(dt)
    B = torch.randn(N, K, device="cuda", generator=g).to(dt)
    bias = torch.randn(N, device="cuda", generator=g)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A.float() @ B.float().t()
    out = _gemm(din, 0, A, B, flags=flags)
    scale = ref.abs().max().item()
    assert (out - ref).abs().max().item() <= 2e-5 * scale * max(1, K / 1024)
    out2 = _gemm(din, 0, A, B, bias=bias, alpha=0.5, acc=torch.ones(M, N, device="cuda"), flags=flags)
    assert (out2 - (0.5 * ref + bias + 1)).abs().max().item() <= 3e-5 * scale * max(1, K / 1024)
    out16 = _gemm(din, 1 if din == 0 else 2, A, B, flags=flags)
    assert (out16.float() - ref).abs().max().item() <= 1e-2 * scale


def test_gemm_strided_views_with_k_offset():
    """the dU GEMMs read column-offset views of [rows, R] buffers (time shift of the recurrence)."""
    g = torch.Generator(device="cuda").manual_seed(0)
    R, H, N = 40 * 32, 64, 32
    hT = torch.randn(2 * H, R, device="cuda", generator=g).to(torch.bfloat16)
    dzT = torch.randn(4 * H, R, device="cuda", generator=g).to(torch.bfloat16)
    K = R - N
    A, B = hT[H:2 * H], dzT[:, N:]
    ref = A[:, :K].float() @ B[:, :K].float().t()
    from asr_study_b200._lib import lib, cur_stream, ptr
    out = torch.empty(H, 4 * H, device="cuda")
    lib.asr_gemm_tn(1, 0, H, 4 * H, K, C.c_void_p(A.data_ptr()), R, C.c_void_p(B.data_ptr()), R, ptr(out), 4 * H,
                    None, 1.0, 0, cur_stream())
    torch.cuda.synchronize()
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
