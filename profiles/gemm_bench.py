"""TF/s of the C2 projection GEMMs on both tcgen05 engines (gemm_tc.cu = tc1, gemm_tc2.cu = default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from asr_study_b200._lib import lib, ptr, cur_stream

R = 999 * 32
SHAPES = [("zx l1/l2 (fp16)", 0, R, 4096, 1024), ("zx per-direction", 0, R, 2048, 1024), ("dX (bf16)", 1, R, 1024, 4096),
          ("dX per-direction", 1, R, 1024, 2048), ("dW per-direction", 1, 1024, 2048, R), ("dU per-direction", 1, 512, 2048, R - 32),
          ("zx l0 (K=32)", 0, R, 4096, 32)]
for name, din, M, N, K in SHAPES:
    dt = torch.float16 if din == 0 else torch.bfloat16
    A = torch.randn(M, K, device="cuda").to(dt)
    B = torch.randn(N, K, device="cuda").to(dt)
    Cm = torch.empty(M, N, device="cuda")
    for eng in ("tc1", "default"):
        os.environ.pop("ASR_B200_GEMM", None)
        if eng != "default":
            os.environ["ASR_B200_GEMM"] = eng
        for _ in range(3):
            lib.asr_gemm_tn(din, 0, M, N, K, ptr(A), K, ptr(B), K, ptr(Cm), N, None, 1.0, 0, cur_stream())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            lib.asr_gemm_tn(din, 0, M, N, K, ptr(A), K, ptr(B), K, ptr(Cm), N, None, 1.0, 0, cur_stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:22s} M={M:6d} N={N:5d} K={K:6d}  {eng:8s} {ms:7.3f} ms  {2.0 * M * N * K / ms / 1e9:8.1f} TF/s", flush=True)
