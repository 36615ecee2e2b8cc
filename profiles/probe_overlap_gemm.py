import sys, os, ctypes as C
sys.path.insert(0, '/root/repo')
import torch
from asr_study_b200._lib import lib, ptr, cur_stream
dev='cuda'
torch.manual_seed(0)
for (rows, ld, K, N) in [(8160, 640, 7040, 320), (8080, 80, 448, 640), (300, 64, 512, 128)]:
    flat = (torch.randn(rows * ld + K + 64, device=dev) * 0.5).half()
    B = (torch.randn(N, K, device=dev) * 0.05).half()
    out = torch.zeros(rows, N, device=dev)
    rc = lib.asr_gemm_tn(0, 0, rows, N, K, ptr(flat), ld, ptr(B), K, ptr(out), N, None, C.c_float(1.0), 0, cur_stream())
    torch.cuda.synchronize()
    A = flat[: (rows - 1) * ld + K].unfold(0, K, ld)       # [rows, K] overlapping view
    ref = A.float() @ B.float().t()
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        lib.asr_gemm_tn(0, 0, rows, N, K, ptr(flat), ld, ptr(B), K, ptr(out), N, None, C.c_float(1.0), 0, cur_stream())
    e1.record(); torch.cuda.synchronize()
    print(rows, ld, K, N, 'rc', rc, 'relerr', err, 'ms', e0.elapsed_time(e1) / 10)
