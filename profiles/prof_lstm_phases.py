import sys, os, ctypes as C, numpy as np, torch
sys.path.insert(0, "/root/repo")
from asr_study_b200._lib import LstmFwdArgs, LstmBwdArgs, lib, ptr, cur_stream
T, N, H = 999, 32, 512
R = T * N
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
zx = torch.randn(R, 8 * H, device=dev, generator=g) * 0.5
bias = torch.zeros(8 * H, device=dev)
U = torch.randn(2, H, 4 * H, device=dev, generator=g) * 0.04
UT16 = U.transpose(1, 2).contiguous().half()
Ub16 = U.to(torch.bfloat16).contiguous()
h16 = torch.empty(R, 2 * H, dtype=torch.float16, device=dev)
hT16 = torch.empty(2 * H, R, dtype=torch.bfloat16, device=dev)
gates = torch.empty(R, 8 * H, device=dev); cell = torch.empty(R, 2 * H, device=dev)
flags = torch.zeros(lib.asr_lstm_flags_bytes() // 4, dtype=torch.int32, device=dev)
a = LstmFwdArgs(T=T, N=N, H=H, training=1, zx=ptr(zx).value, bias=ptr(bias).value, U=ptr(U).value, U16=ptr(UT16).value,
                h16=ptr(h16).value, hT16=ptr(hT16).value, h32=None, gates=ptr(gates).value, cell=ptr(cell).value, flags=ptr(flags).value)
names = ["poll_LL", "smem+fence+sync", "issue", "mma_wait", "tmem_ld+xchg", "gates+publish", "side_stores", "loop_top"]
for rep in range(int(os.environ.get("REPS", "2"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.asr_lstm_forward(C.byref(a), cur_stream()); e1.record(); torch.cuda.synchronize()
    p = flags[1024:1024 + 64].view(torch.int64).cpu().numpy()
    print("fwd ms", e0.elapsed_time(e1), "status", int(flags[64]))
    print("  fwd cycles/step:", {n: int(v / T) for n, v in zip(names, p[:8])}, "sum", int(p[:8].sum() / T))
    print("  fwd issue split (fence_after, mma 0, mma 1-7, mma 8-15, mma 16-31, commit):", [int(v / T) for v in p[16:22]])
dh = torch.randn(R, 2 * H, device=dev, generator=g) * 0.01
dz16 = torch.empty(R, 8 * H, dtype=torch.bfloat16, device=dev); dzT16 = torch.empty(8 * H, R, dtype=torch.bfloat16, device=dev)
dbias = torch.zeros(8 * H, device=dev)
b = LstmBwdArgs(T=T, N=N, H=H, dh=ptr(dh).value, gates=ptr(gates).value, cell=ptr(cell).value, U=ptr(U).value, U16=ptr(Ub16).value,
                dz16=ptr(dz16).value, dzT16=ptr(dzT16).value, dz32=None, dbias=ptr(dbias).value, flags=ptr(flags).value)
for rep in range(int(os.environ.get("REPS", "2"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); lib.asr_lstm_backward(C.byref(b), cur_stream()); e1.record(); torch.cuda.synchronize()
    p = flags[1024:1024 + 64].view(torch.int64).cpu().numpy()
    print("bwd ms", e0.elapsed_time(e1), "status", int(flags[64]))
    bn = ["hop1_poll", "smem+fence+sync", "issue+mma_wait", "tmem_ld+hop2_send", "hop2_poll+sync", "bptt+publish", "side_stores", "loop_top"]
    print("  bwd cycles/step:", {n: int(v / T) for n, v in zip(bn, p[8:16])}, "sum", int(p[8:16].sum() / T))
