// K3 (tensor-core engine, v2) — persistent BiLSTM forward recurrence on tcgen05, sm_100a.
//
// What changed against lstm_tc.cu (kept as the v1 cross-check) and why — from the per-phase clock64
// profile in profiles/lstm_phases_r1.md (10.3 k cycles per step, of which MMA 2.3 k, L2 loads 3.2 k,
// release fence 2.3 k, flag wait 1.6 k):
//   * the CTA's [128 x H] slice of U^T lives in TENSOR MEMORY for the whole sequence (tcgen05.st once),
//     and the step product is issued as TS-mode tcgen05.mma (A from TMEM, B = h_{t-1} from smem): no
//     4 KB shared-memory re-read of A per N=16 MMA (that made SS-mode ~70 cycles per instruction);
//   * h_t is exchanged through an LL-style ring in L2: every 8-byte word is {2 x fp16 h, step tag},
//     written with one volatile 8-byte store and polled with volatile 8-byte loads, so data and
//     readiness arrive atomically in ONE L2 hop — no release fence, no separate counter, no second
//     round trip for the payload; all of a thread's words are in flight together.
// Grid = (H/32 CTAs, 2 directions, N/16 batch groups), cooperative launch; each CTA owns 32 hidden
// units = 128 gate rows, accumulates [128 x 16] in TMEM, keeps c_t in registers.
//
// Semantics: core/layers.py:432-469 under Keras-1 Bidirectional, no masking (see lstm_fp32.cu).
#include "common.cuh"
#include "tc.cuh"

namespace lstmtc2 {

constexpr int UPC = 32;
constexpr int NG = 16;
constexpr int THREADS = 128;
constexpr int STATUS_IDX = 64;
constexpr int HEADER_BYTES = 8192;
constexpr uint32_t D_COL = 0, A_COL = 32;
constexpr long long WATCHDOG_CYCLES = 2000000000LL;

__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_v2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}

template <int H>
__global__ void __launch_bounds__(THREADS, 1)
fwd_kernel(asr_lstm_fwd_args a, int* __restrict__ flags, uint2* __restrict__ xbuf) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int KC = H / 64;
  constexpr int B_CHUNK = NG * 128;
  constexpr int WORDS = NG * H / 2;                      // LL words per (dir, group, parity)
  constexpr int WPT = WORDS / THREADS;
  constexpr int NPT = NG / 4;
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int u0 = cta * UPC, n0 = grp * NG;

  uint8_t* sB = smem;                                    // KC chunks of [NG rows x 128 B], SW128 K-major
  float* sZ = reinterpret_cast<float*>(sB + KC * B_CHUNK);   // [4 gates][NG][32 units]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sZ + 4 * NG * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  __shared__ int s_dead;

  if (tid == 0) {
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  // ---- one-time: U^T slice -> TMEM.  TMEM lane r = g*32 + j holds row (g*H + u0 + j) of U^T, two fp16
  //      K-elements per 32-bit column (the kind::f16 A-operand layout). --------------------------------
  {
    const int g = warp, j = lane;
    const uint4* row = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.U16) +
                                                      ((size_t)dir * 4 * H + (size_t)g * H + u0 + j) * H);
#pragma unroll 1
    for (int c = 0; c < H / 2; c += 32) {
      uint32_t r[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint4 v = __ldg(row + c / 4 + q);
        r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
      }
      tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + A_COL + c, r);
    }
    tc::tmem_st_wait();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();

  const uint32_t idesc = tc::umma_idesc_f16(128, NG, 0);
  const uint32_t sB_addr = tc::smem_u32(sB);
  const int u = u0 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
  float c_state[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) c_state[i] = 0.0f;

  int* status = flags + STATUS_IDX;
  uint2* xb = xbuf + (size_t)(dir * G + grp) * 2 * WORDS;
  __half* h16 = reinterpret_cast<__half*>(a.h16);
  const size_t R = (size_t)T * N;

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    float zx[NPT][4];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const float* zr = a.zx + (((size_t)t * N + n0 + warp * NPT + i) * 2 + dir) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) zx[i][g] = __ldg(zr + g * H + u);
    }
    float z[NPT][4];
    if (s > 0) {
      // ---- pull h_{t-1} of this group: poll the LL words (data + tag in one 8-byte access) ----------
      const uint2* src = xb + (size_t)((s - 1) & 1) * WORDS + tid;
      const uint32_t tag = (uint32_t)s;
      uint2 w[WPT];
#pragma unroll
      for (int q = 0; q < WPT; ++q) w[q] = ld_volatile_v2(src + q * THREADS);
      bool ok;
      long long t0 = 0;
      do {
        ok = true;
#pragma unroll
        for (int q = 0; q < WPT; ++q)
          if (w[q].y != tag) {
            w[q] = ld_volatile_v2(src + q * THREADS);
            ok = false;
          }
        if (!ok) {
          if (t0 == 0) t0 = clock64();
          else if (clock64() - t0 > WATCHDOG_CYCLES) {
            atomicExch(status, 1);
            s_dead = 1;
            break;
          }
        }
      } while (!ok);
#pragma unroll
      for (int q = 0; q < WPT; ++q) {
        const int i = tid + q * THREADS;
        const int n = i / (H / 2), k = 2 * (i % (H / 2));
        *reinterpret_cast<uint32_t*>(sB + (k >> 6) * B_CHUNK + tc::sw128_offset(n, k & 63)) = w[q].x;
      }
      tc::fence_proxy_async_smem();
      __syncthreads();
      if (s_dead) break;
      if (tid == 0) {
        tc::tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < H / 16; ++kb) {
          const uint64_t bd = tc::umma_desc_sw128(sB_addr + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
          tc::umma_ts(tmem + D_COL, tmem + A_COL + kb * 8, bd, idesc, kb != 0);
        }
        tc::umma_commit(mma_bar);
      }
      if (!tc::mbar_wait(mma_bar, (uint32_t)((s - 1) & 1), WATCHDOG_CYCLES)) {
        atomicExch(status, 1);
        s_dead = 1;
      }
      tc::tcgen05_fence_after();
      {
        uint32_t r[NG];
        tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + D_COL, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < NG; ++n) sZ[(warp * NG + n) * 32 + lane] = __uint_as_float(r[n]);
      }
      tc::tcgen05_fence_before();
      __syncthreads();
      if (s_dead) break;
#pragma unroll
      for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) z[i][g] = sZ[(g * NG + warp * NPT + i) * 32 + lane];
    } else {
#pragma unroll
      for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) z[i][g] = 0.0f;
    }
    float gi[NPT], gf[NPT], gg[NPT], go[NPT], hv[NPT];
    uint2* xo = xb + (size_t)(s & 1) * WORDS;
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      gi[i] = asr::hard_sigmoid(z[i][0] + zx[i][0] + bias[0]);
      gf[i] = asr::hard_sigmoid(z[i][1] + zx[i][1] + bias[1]);
      gg[i] = tanhf(z[i][2] + zx[i][2] + bias[2]);
      go[i] = asr::hard_sigmoid(z[i][3] + zx[i][3] + bias[3]);
      c_state[i] = gf[i] * c_state[i] + gi[i] * gg[i];
      hv[i] = go[i] * tanhf(c_state[i]);
      // publish: even lanes pack (unit, unit+1) into one LL word {half2, tag = s+1}
      const float other = __shfl_down_sync(0xffffffffu, hv[i], 1);
      if (!(lane & 1)) {
        const __half2 pk = __floats2half2_rn(hv[i], other);
        uint2 wv;
        wv.x = *reinterpret_cast<const uint32_t*>(&pk);
        wv.y = (uint32_t)(s + 1);
        st_volatile_v2(xo + (size_t)(warp * NPT + i) * (H / 2) + (u >> 1), wv);
      }
    }
    // side outputs (not on the recurrence's critical path)
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      h16[row * 2 * H + dir * H + u] = __float2half_rn(hv[i]);
      if (a.h32) a.h32[row * 2 * H + dir * H + u] = hv[i];
      if (a.training) {
        float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gp[u] = gi[i]; gp[H + u] = gf[i]; gp[2 * H + u] = gg[i]; gp[3 * H + u] = go[i];
        a.cell[(row * 2 + dir) * H + u] = c_state[i];
        if (a.hT16) reinterpret_cast<__nv_bfloat16*>(a.hT16)[(size_t)(dir * H + u) * R + row] = __float2bfloat16_rn(hv[i]);
      }
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------------------------
static bool shape_ok(int T, int N, int H) {
  return T >= 1 && (H == 128 || H == 256 || H == 384 || H == 512) && N >= NG && N % NG == 0 && N / NG <= 8 &&
         (H / UPC) * 2 * (N / NG) <= 148;
}
bool supports_fwd(const asr_lstm_fwd_args* a) { return a->U16 && a->h16 && shape_ok(a->T, a->N, a->H); }
size_t scratch_bytes(int) { return HEADER_BYTES + (size_t)2 * 8 * 2 * NG * (512 / 2) * sizeof(uint2); }

template <int H>
static int32_t launch_fwd(const asr_lstm_fwd_args* a, cudaStream_t st) {
  constexpr int KC = H / 64;
  const size_t smem = 1024 + (size_t)KC * NG * 128 + 4 * NG * 32 * 4 + 64;
  const int G = a->N / NG;
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const size_t xbytes = (size_t)2 * G * 2 * NG * (H / 2) * sizeof(uint2);
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, HEADER_BYTES + xbytes, st));
  asr_lstm_fwd_args args = *a;
  int* flags = a->flags;
  uint2* xbuf = reinterpret_cast<uint2*>(reinterpret_cast<char*>(a->flags) + HEADER_BYTES);
  void* kargs[] = {&args, &flags, &xbuf};
  ASR_CUDA(cudaLaunchCooperativeKernel((void*)fwd_kernel<H>, dim3(H / UPC, 2, G), dim3(THREADS), kargs, smem, st));
  asr::count_launch();
  return ASR_OK;
}

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  switch (a->H) {
    case 128: return launch_fwd<128>(a, st);
    case 256: return launch_fwd<256>(a, st);
    case 384: return launch_fwd<384>(a, st);
    case 512: return launch_fwd<512>(a, st);
  }
  asr::set_error("lstmtc2: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

}  // namespace lstmtc2
