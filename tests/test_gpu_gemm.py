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


def test_cast_batch_matches_the_single_tensor_entries():
    """asr_cast_batch: every job of a list (row casts with K padding into a column sub-block, transposes, the fp16 rounding
    residual) in one launch, bit-identical to asr_cast_rows / asr_cast_transpose."""
    from asr_study_b200._lib import CastJob, cur_stream, lib, ptr
    g = torch.Generator(device="cuda").manual_seed(5)
    specs = [(37, 26, 0, 0), (512, 2048, 1, 0), (100, 28, 0, 1), (26, 2048, 0, 1), (64, 72, 16, 1), (9, 13, 16, 0), (300, 40, 1, 1)]
    jobs, checks = [], []
    for rows, cols, dtype, tr in specs:
        src = torch.randn(rows, cols + 3, device="cuda", generator=g)
        tdt = torch.bfloat16 if dtype == 1 else torch.float16
        ld = (rows + 8) if tr else ((cols + 7) // 8 * 8 + 8)
        shape = (cols, ld) if tr else (rows, ld)
        got = torch.full(shape, 7.0, dtype=tdt, device="cuda")
        ref = torch.full(shape, 7.0, dtype=tdt, device="cuda")
        jobs.append(CastJob(src.data_ptr(), cols + 3, got.data_ptr(), ld, rows, cols, dtype, tr))
        fn = lib.asr_cast_transpose if tr else lib.asr_cast_rows
        fn(ptr(src), cols + 3, ptr(ref), ld, rows, cols, dtype, cur_stream())
        checks.append((src, got, ref))
    lib.asr_cast_batch((CastJob * len(jobs))(*jobs), len(jobs), cur_stream())
    torch.cuda.synchronize()
    for src, got, ref in checks:
        assert torch.equal(got.view(torch.int16), ref.view(torch.int16))
    # mode 2: two column blocks of one matrix whose width is not a multiple of 8, no padding fill across the seam
    a, b = torch.randn(50, 108, device="cuda", generator=g), torch.randn(50, 108, device="cuda", generator=g)
    out = torch.zeros(50, 216, dtype=torch.bfloat16, device="cuda")
    two = [CastJob(a.data_ptr(), 108, out.data_ptr(), 216, 50, 108, 1, 2), CastJob(b.data_ptr(), 108, out[:, 108:].data_ptr(), 216, 50, 108, 1, 2)]
    lib.asr_cast_batch((CastJob * 2)(*two), 2, cur_stream())
    torch.cuda.synchronize()
    assert torch.equal(out, torch.cat([a, b], 1).to(torch.bfloat16))


@pytest.mark.parametrize("rows,ld,K,N", [(8192, 640, 7040, 320), (4096, 80, 448, 640), (300, 64, 512, 128)])
def test_gemm_overlapping_rows_view(rows, ld, K, N):
    """lda < K: row m of A starts lda elements after row m-1 (the convolution view of csrc/conv.cu) — the TMA tensor map
    takes a global stride smaller than the box; against the same contraction on torch's unfold view."""
    g = torch.Generator(device="cuda").manual_seed(11)
    flat = (torch.randn(rows * ld + K + 64, device="cuda", generator=g) * 0.5).half()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    got = _gemm(0, 0, flat, B, M=rows, N=N, K=K, lda=ld, ldb=K)
    ref = flat[:(rows - 1) * ld + K].unfold(0, K, ld).float() @ B.float().t()
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5
