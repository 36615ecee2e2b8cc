"""Wall time (CUDA events) of the two recurrence engines at the C2 shape (T = 999, N = 32, H = 512, training):
lstm_tc2.cu (fp32 storage, plain loads / stores) and lstm_tc4.cu (fp16 storage, TMA ring + TMA tile stores), the latter
for a sweep of the first-probe delay (opts bits 16..27 = cycles / 8; 0xFFF = no delay)."""
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
zx = torch.randn(R, 8 * H, device=dev, generator=g) * 0.5
zx16 = zx.half()
bias = torch.zeros(8 * H, device=dev)
U = torch.randn(2, H, 4 * H, device=dev, generator=g) * 0.04
UT16 = U.transpose(1, 2).contiguous().half()
Ub16 = U.to(torch.bfloat16).contiguous()
mask_u = ((torch.rand(2, N, H, device=dev, generator=g) >= 0.2) / 0.8).float()
mask_n = ((torch.rand(2, N, 2 * H, device=dev, generator=g) >= 0.2) / 0.8).float()
hm16 = torch.empty(2, R, 2 * H, dtype=torch.float16, device=dev)
hmT16 = torch.empty(2, 2 * H, R, dtype=torch.bfloat16, device=dev)
hT16 = torch.empty(2 * H, R, dtype=torch.bfloat16, device=dev)
gates = torch.empty(R, 8 * H, device=dev)
cell = torch.empty(R, 2 * H, device=dev)
gates16 = torch.empty(R, 8 * H, dtype=torch.float16, device=dev)
cell16 = torch.empty(R, 2 * H, dtype=torch.float16, device=dev)
flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device=dev)
dh = torch.randn(R, 2 * H, device=dev, generator=g) * 0.01
dh2 = torch.randn(R, 2 * H, device=dev, generator=g) * 0.01
dz16 = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev)
dzT16 = torch.empty(8 * H, R, dtype=torch.bfloat16, device=dev)
dbias = torch.zeros(8 * H, device=dev)
common = dict(T=T, N=N, H=H, training=1, bias=ptr(bias).value, U=ptr(U).value, U16=ptr(UT16).value, hT16=ptr(hT16).value,
              flags=ptr(flags).value, mask_u=ptr(mask_u).value, mask_next=ptr(mask_n).value, hm16=ptr(hm16).value,
              hmT16=ptr(hmT16).value)
bcommon = dict(T=T, N=N, H=H, dh=ptr(dh).value, dh2=ptr(dh2).value, mask_dh=ptr(mask_n).value, U=ptr(U).value,
               U16=ptr(Ub16).value, dz16=ptr(dz16).value, dzT16=ptr(dzT16).value, dbias=ptr(dbias).value,
               flags=ptr(flags).value, mask_u=ptr(mask_u).value)


def timed(fn, reps=8):
    fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1))
    assert int(flags[64]) == 0
    best.sort()
    return best[len(best) // 2]


a2 = LstmFwdArgs(zx=ptr(zx).value, gates=ptr(gates).value, cell=ptr(cell).value, **common)
b2 = LstmBwdArgs(gates=ptr(gates).value, cell=ptr(cell).value, **bcommon)
print("tc2 (fp32 storage): fwd %.3f ms  bwd %.3f ms" % (timed(lambda: lib.asr_lstm_forward(C.byref(a2), cur_stream())),
                                                        timed(lambda: lib.asr_lstm_backward(C.byref(b2), cur_stream()))))
for delay in (None, 150, 300, 600):
    opts = 0 if delay is None else ((delay // 8) << 16)
    a4 = LstmFwdArgs(zx16=ptr(zx16).value, gates16=ptr(gates16).value, cell16=ptr(cell16).value, opts=opts, **common)
    b4 = LstmBwdArgs(gates16=ptr(gates16).value, cell16=ptr(cell16).value, opts=opts, **bcommon)
    print("tc4 (fp16 storage + TMA), probe delay %s: fwd %.3f ms  bwd %.3f ms" % ( "default (none)" if delay is None else ("none" if delay == 0xFFF else str(delay)),
        timed(lambda: lib.asr_lstm_forward(C.byref(a4), cur_stream())),
        timed(lambda: lib.asr_lstm_backward(C.byref(b4), cur_stream()))))
