"""clock64() phase profile of the lstm_tc4.cu recurrences (build: make EXTRA=-DASR_LSTM_PROFILE BUILD=build_prof
LIB=../libasr_b200_prof.so; run with ASR_B200_LIB pointing at it).  C2 shape, fused dropout configuration."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from asr_study_b200._lib import LstmBwdArgs, LstmFwdArgs, cur_stream, lib, ptr  # noqa: E402

T, N, H = 999, 32, 512
R = T * N
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
zx16 = (torch.randn(R, 8 * H, device=dev, generator=g) * 0.5).half()
bias = torch.zeros(8 * H, device=dev)
U = torch.randn(2, H, 4 * H, device=dev, generator=g) * 0.04
UT16 = U.transpose(1, 2).contiguous().half()
Ub16 = U.to(torch.bfloat16).contiguous()
mask_u = ((torch.rand(2, N, H, device=dev, generator=g) >= 0.2) / 0.8).float()
mask_n = ((torch.rand(2, N, 2 * H, device=dev, generator=g) >= 0.2) / 0.8).float()
hm16 = torch.empty(2, R, 2 * H, dtype=torch.float16, device=dev)
hmT16 = torch.empty(2, 2 * H, R, dtype=torch.bfloat16, device=dev)
hT16 = torch.empty(2 * H, R, dtype=torch.bfloat16, device=dev)
gates16 = torch.empty(R, 8 * H, dtype=torch.float16, device=dev)
cell16 = torch.empty(R, 2 * H, dtype=torch.float16, device=dev)
flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device=dev)
dh = torch.randn(R, 2 * H, device=dev, generator=g) * 0.01
dh2 = torch.randn(R, 2 * H, device=dev, generator=g) * 0.01
dz16 = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
dzT16 = torch.empty(8 * H, R, dtype=torch.bfloat16, device=dev)
dbias = torch.zeros(8 * H, device=dev)
a = LstmFwdArgs(T=T, N=N, H=H, training=1, bias=ptr(bias).value, U=ptr(U).value, U16=ptr(UT16).value, hT16=ptr(hT16).value,
                flags=ptr(flags).value, mask_u=ptr(mask_u).value, mask_next=ptr(mask_n).value, hm16=ptr(hm16).value,
                hmT16=ptr(hmT16).value, zx16=ptr(zx16).value, gates16=ptr(gates16).value, cell16=ptr(cell16).value)
b = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh).value, dh2=ptr(dh2).value, mask_dh=ptr(mask_n).value, U=ptr(U).value,
                U16=ptr(Ub16).value, dz16=ptr(dz16).value, dzT16=ptr(dzT16).value, dbias=ptr(dbias).value,
                flags=ptr(flags).value, mask_u=ptr(mask_u).value, gates16=ptr(gates16).value, cell16=ptr(cell16).value)
fn = ["ring wait + input regs", "poll (delay + LL words)", "stage + MMA issue", "mma wait", "tmem ld + gate swap + bar", "gates + publish",
      "staging writes + arrive", "loop top"]
bn = ["ring wait + input regs", "poll + partial sums", "BPTT math + stage B", "bar + MMA issue + wait", "tmem ld + send", "dz staging + arrive", "-",
      "loop top"]
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.asr_lstm_forward(C.byref(a), cur_stream()); e1.record(); torch.cuda.synchronize()
    p = flags[1024:1024 + 64].view(torch.int64).cpu().numpy()
    print("fwd ms %.3f status %d" % (e0.elapsed_time(e1), int(flags[64])))
    print("  fwd cycles/step:", {n: int(v / T) for n, v in zip(fn, p[:8])}, "sum", int(p[:8].sum() / T))
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.asr_lstm_backward(C.byref(b), cur_stream()); e1.record(); torch.cuda.synchronize()
    p = flags[1024:1024 + 64].view(torch.int64).cpu().numpy()
    print("bwd ms %.3f status %d" % (e0.elapsed_time(e1), int(flags[64])))
    print("  bwd cycles/step:", {n: int(v / T) for n, v in zip(bn, p[8:16])}, "sum", int(p[8:16].sum() / T))
