// K3/K4 (tensor-core engine) — persistent BiLSTM recurrence on tcgen05, sm_100a.
//
// One cooperative launch runs both directions and all batch groups: grid = (H/32, 2, N/NG).
// Forward: each CTA owns 32 hidden units = 128 gate rows (i,f,c,o x 32) of one direction and keeps
//   its [128 x H] fp16 slice of U^T RESIDENT in shared memory (UMMA K-major, 128B-swizzled) for all T
//   steps.  Per step it pulls h_{t-1} [NG x H] (fp16, written by its peers straight into the layer
//   output tensor, which doubles as the exchange buffer) from L2 into the UMMA B layout, one elected
//   thread issues H/16 tcgen05.mma (M=128, N=NG, K=16) into a TMEM accumulator, and the 4 warps pull
//   their TMEM lane quarter (= one gate each), swap through smem so one thread owns the 4 gates of a
//   (unit, sample), apply hard-sigmoid/tanh, update the cell state it keeps in registers, and write
//   h_t / saved activations.  Peers are released through a per-(direction, group) step counter
//   (red.release / ld.acquire at gpu scope); there is no grid-wide barrier.
// Backward: same skeleton with A = the CTA's 32 ROWS of U ([32 x 4H] bf16, issued as M=64 tiles whose
//   upper half aliases the next K-chunk and is ignored), B = dz_{t+1} [NG x 4H] (bf16) streamed in 4
//   K-quarters so the MMAs of quarter q overlap the L2 loads of quarter q+1.
//
// Semantics: core/layers.py:432-469 under Keras-1 Bidirectional, no masking (see lstm_fp32.cu).
#include "common.cuh"
#include "tc.cuh"

namespace lstmtc {

#ifdef ASR_LSTM_PROFILE
#define PROF_DECL long long pt0 = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { const long long now = clock64(); pacc[i] += now - pt0; pt0 = now; } while (0)
#define PROF_DUMP(base) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) for (int i = 0; i < 8; ++i) reinterpret_cast<long long*>(flags + 1024)[(base) + i] = pacc[i]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_DUMP(base)
#endif

constexpr int UPC = 32;                 // hidden units per CTA
constexpr int FLAG_BASE = 128;          // ints; [0,128) belongs to the fp32 engine's header (status at 64)
constexpr int STATUS_IDX = 64;
constexpr long long WATCHDOG_CYCLES = 2000000000LL;   // ~1 s

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void wait_step(const int* flag, int target, int* status) {
  if (threadIdx.x == 0) {
    if (ld_acquire(flag) < target) {
      const long long t0 = clock64();
      while (ld_acquire(flag) < target) {
        if (clock64() - t0 > WATCHDOG_CYCLES) {
          atomicExch(status, 1);
          break;
        }
      }
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(128, 1)
fwd_kernel(asr_lstm_fwd_args a, int* __restrict__ flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int T = a.T, N = a.N, H = a.H;
  const int KC = H / 64;                               // 64-element K chunks
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z;
  const int u0 = cta * UPC, n0 = grp * NG;
  constexpr int A_CHUNK = 128 * 128;                   // bytes: 128 rows x 128 B
  constexpr int B_CHUNK = NG * 128;

  uint8_t* sA = smem;                                  // KC * 16 KB
  uint8_t* sB = sA + (size_t)KC * A_CHUNK;             // KC * NG*128 B
  float* sZ = reinterpret_cast<float*>(sB + (size_t)KC * B_CHUNK);   // [4][NG][32]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sZ + 4 * NG * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  // ---- one-time: resident U^T slice.  row r = g*32 + j  <->  U16[dir][g*H + u0 + j][:] ----------
  {
    const __half* Ut = reinterpret_cast<const __half*>(a.U16) + (size_t)dir * 4 * H * H;
    const int chunks_per_row = H / 8;
    for (int i = tid; i < 128 * chunks_per_row; i += 128) {
      const int r = i / chunks_per_row, c16 = i % chunks_per_row;
      const int g = r >> 5, j = r & 31;
      const uint4 v = *reinterpret_cast<const uint4*>(Ut + ((size_t)(g * H + u0 + j)) * H + c16 * 8);
      const int kc = c16 >> 3, cc = c16 & 7;
      *reinterpret_cast<uint4*>(sA + (size_t)kc * A_CHUNK + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4)) = v;
    }
  }
  if (tid == 0) {
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 32);
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = tc::umma_idesc_f16(128, NG, 0);
  const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);

  // epilogue ownership: unit j = lane, samples n = warp*NPT .. +NPT-1
  constexpr int NPT = NG / 4;
  const int u = u0 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
  float c_state[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) c_state[i] = 0.0f;

  int* flag = flags + FLAG_BASE + (dir * 8 + grp) * 32;
  int* status = flags + STATUS_IDX;
  const int nctas = gridDim.x;
  __half* h16 = reinterpret_cast<__half*>(a.h16);
  const size_t R = (size_t)T * N;

  __shared__ int s_dead;
  PROF_DECL;
  for (int s = 0; s < T; ++s) {
    PROF(7);
    if ((s & 15) == 15) {                              // bail out together if any CTA's watchdog fired
      if (tid == 0) s_dead = ld_acquire(status);
      __syncthreads();
      if (s_dead) break;
    }
    const int t = dir ? (T - 1 - s) : s;
    const int tp = dir ? (t + 1) : (t - 1);
    // input projection for my (unit, samples): issued before the wait, consumed after the MMA
    float zx[NPT][4];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int n = n0 + warp * NPT + i;
      const float* zr = a.zx + (((size_t)t * N + n) * 2 + dir) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) zx[i][g] = __ldg(zr + g * H + u);
    }
    float z[NPT][4];
    if (s > 0) {
      wait_step(flag, nctas * s, status);
      PROF(0);
      // h_{t-1} rows of this group -> UMMA B layout (K-major, SW128)
      {
        const int chunks_per_row = H / 8;
        const __half* src = h16 + ((size_t)tp * N + n0) * 2 * H + dir * H;
        for (int i = tid; i < NG * chunks_per_row; i += 128) {
          const int n = i / chunks_per_row, c16 = i % chunks_per_row;
          const uint4 v = ld_cg_v4(src + (size_t)n * 2 * H + c16 * 8);
          const int kc = c16 >> 3, cc = c16 & 7;
          *reinterpret_cast<uint4*>(sB + (size_t)kc * B_CHUNK + (n >> 3) * 1024 + (n & 7) * 128 + ((cc ^ (n & 7)) << 4)) = v;
        }
      }
      tc::fence_proxy_async_smem();
      __syncthreads();
      PROF(1);
      if (tid == 0) {
        tc::tcgen05_fence_after();
        for (int kc = 0; kc < KC; ++kc) {
          const uint64_t ad = tc::umma_desc_sw128(sA_addr + kc * A_CHUNK);
          const uint64_t bd = tc::umma_desc_sw128(sB_addr + kc * B_CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k) tc::umma_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, (kc | k) != 0);
        }
        tc::umma_commit(mma_bar);
      }
      PROF(2);
      if (!tc::mbar_wait(mma_bar, (uint32_t)((s - 1) & 1), WATCHDOG_CYCLES)) atomicExch(status, 1);
      tc::tcgen05_fence_after();
      PROF(3);
      {
        uint32_t r[NG];
        if constexpr (NG == 16) tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
        else tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < NG; ++n) sZ[(warp * NG + n) * 32 + lane] = __uint_as_float(r[n]);
      }
      tc::tcgen05_fence_before();
      __syncthreads();
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
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      gi[i] = asr::hard_sigmoid(z[i][0] + zx[i][0] + bias[0]);
      gf[i] = asr::hard_sigmoid(z[i][1] + zx[i][1] + bias[1]);
      gg[i] = tanhf(z[i][2] + zx[i][2] + bias[2]);
      go[i] = asr::hard_sigmoid(z[i][3] + zx[i][3] + bias[3]);
      c_state[i] = gf[i] * c_state[i] + gi[i] * gg[i];
      hv[i] = go[i] * tanhf(c_state[i]);
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      h16[row * 2 * H + dir * H + u] = __float2half_rn(hv[i]);      // exchange-critical store first
    }
    PROF(4);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release(flag, 1);
    }
    PROF(5);
    // non-critical outputs after the release: they overlap the peers' next step
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      if (a.h32) a.h32[row * 2 * H + dir * H + u] = hv[i];
      if (a.training) {
        float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gp[u] = gi[i]; gp[H + u] = gf[i]; gp[2 * H + u] = gg[i]; gp[3 * H + u] = go[i];
        a.cell[(row * 2 + dir) * H + u] = c_state[i];
        if (a.hT16) reinterpret_cast<__nv_bfloat16*>(a.hT16)[(size_t)(dir * H + u) * R + row] = __float2bfloat16_rn(hv[i]);
      }
    }
    PROF(6);
  }
  PROF_DUMP(0);
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// ------------------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------------------
template <int NG>
__global__ void __launch_bounds__(256, 1)
bwd_kernel(asr_lstm_bwd_args a, int* __restrict__ flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int T = a.T, N = a.N, H = a.H, K4 = 4 * a.H;
  const int KC = K4 / 64;                              // K chunks over the 4H gate columns
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z;
  const int u0 = cta * UPC, n0 = grp * NG;
  constexpr int A_CHUNK = 32 * 128;                    // 32 real rows x 128 B (M=64 tiles alias the next chunk)
  constexpr int B_CHUNK = NG * 128;

  uint8_t* sA = smem;                                  // KC * 4 KB (+ 4 KB tail read by the last M=64 tile)
  uint8_t* sB = sA + (size_t)KC * A_CHUNK;             // KC * NG*128 B   (the tail read lands here: finite data)
  float* sD = reinterpret_cast<float*>(sB + (size_t)KC * B_CHUNK);   // [32 units][NG]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sD + 32 * NG);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);

  {
    const __nv_bfloat16* Ub = reinterpret_cast<const __nv_bfloat16*>(a.U16) + (size_t)dir * H * K4;
    const int chunks_per_row = K4 / 8;
    for (int i = tid; i < 32 * chunks_per_row; i += 256) {
      const int r = i / chunks_per_row, c16 = i % chunks_per_row;
      const uint4 v = *reinterpret_cast<const uint4*>(Ub + (size_t)(u0 + r) * K4 + c16 * 8);
      const int kc = c16 >> 3, cc = c16 & 7;
      *reinterpret_cast<uint4*>(sA + (size_t)kc * A_CHUNK + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4)) = v;
    }
    // make the aliased tail finite before the first MMA ever reads it
    for (int i = tid; i < KC * B_CHUNK / 16; i += 256) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 32);
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = tc::umma_idesc_f16(64, NG, 1);
  const uint32_t sA_addr = tc::smem_u32(sA), sB_addr = tc::smem_u32(sB);

  // elementwise ownership (threads 0..127): unit j = lane, samples n = warp*NPT.. ; warps 4..7 only help load
  constexpr int NPT = NG / 4;
  const bool ew = warp < 4;
  const int u = u0 + lane;
  float dc_carry[NPT], db[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NPT; ++i) dc_carry[i] = 0.0f;

  int* flag = flags + FLAG_BASE + (dir * 8 + grp) * 32;
  int* status = flags + STATUS_IDX;
  const int nctas = gridDim.x;
  __nv_bfloat16* dz16 = reinterpret_cast<__nv_bfloat16*>(a.dz16);
  const size_t R = (size_t)T * N;

  __shared__ int s_dead;
  PROF_DECL;
  for (int s = 0; s < T; ++s) {
    PROF(7);
    if ((s & 15) == 15) {
      if (tid == 0) s_dead = ld_acquire(status);
      __syncthreads();
      if (s_dead) break;
    }
    const int t = dir ? s : (T - 1 - s);               // BPTT walks the forward order backwards
    const int t_fprev = dir ? (t + 1) : (t - 1);       // forward-order predecessor (holds c_{prev})
    const bool has_fprev = dir ? (t + 1 < T) : (t > 0);
    const int t_bprev = dir ? (t - 1) : (t + 1);       // step processed just before this one in BPTT
    float dho[NPT], gi[NPT], gf[NPT], gg[NPT], go[NPT], cc[NPT], cp[NPT];
    if (ew) {
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const size_t row = (size_t)t * N + n0 + warp * NPT + i;
        dho[i] = __ldg(a.dh + row * 2 * H + dir * H + u);
        const float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gi[i] = __ldg(gp + u); gf[i] = __ldg(gp + H + u); gg[i] = __ldg(gp + 2 * H + u); go[i] = __ldg(gp + 3 * H + u);
        cc[i] = __ldg(a.cell + (row * 2 + dir) * H + u);
        cp[i] = has_fprev ? __ldg(a.cell + ((((size_t)t_fprev * N + n0 + warp * NPT + i) * 2 + dir) * H + u)) : 0.0f;
      }
    }
    float dh_rec[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) dh_rec[i] = 0.0f;
    if (s > 0) {
      wait_step(flag, nctas * s, status);
      PROF(0);
      const int chunks_per_row = K4 / 8;               // 16-byte chunks per dz row (this direction)
      const __nv_bfloat16* src = dz16 + (((size_t)t_bprev * N + n0) * 2 + dir) * K4;
      const int per_q = NG * chunks_per_row / 4;
      for (int q = 0; q < 4; ++q) {                    // K quarters: MMAs of quarter q overlap loads of q+1
        for (int i = tid; i < per_q; i += 256) {
          const int n = i / (chunks_per_row / 4), c16 = q * (chunks_per_row / 4) + i % (chunks_per_row / 4);
          const uint4 v = ld_cg_v4(src + (size_t)n * 2 * K4 + c16 * 8);
          const int kc = c16 >> 3, c8 = c16 & 7;
          *reinterpret_cast<uint4*>(sB + (size_t)kc * B_CHUNK + (n >> 3) * 1024 + (n & 7) * 128 + ((c8 ^ (n & 7)) << 4)) = v;
        }
        tc::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          tc::tcgen05_fence_after();
          for (int kc = q * (KC / 4); kc < (q + 1) * (KC / 4); ++kc) {
            const uint64_t ad = tc::umma_desc_sw128(sA_addr + kc * A_CHUNK);
            const uint64_t bd = tc::umma_desc_sw128(sB_addr + kc * B_CHUNK);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc::umma_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, (kc | k) != 0);
          }
          if (q == 3) tc::umma_commit(mma_bar);
        }
      }
      PROF(1);
      if (!tc::mbar_wait(mma_bar, (uint32_t)((s - 1) & 1), WATCHDOG_CYCLES)) atomicExch(status, 1);
      tc::tcgen05_fence_after();
      PROF(3);
      if (warp < 4) {
        uint32_t r[NG];
        if constexpr (NG == 16) tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16), r);
        else tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tc::tmem_ld_wait();
        // M=64 accumulator: row m lives in lane (m % 16) + 32 * (m / 16); rows 0..31 are real
        if (warp < 2 && lane < 16) {
#pragma unroll
          for (int n = 0; n < NG; ++n) sD[(warp * 16 + lane) * NG + n] = __uint_as_float(r[n]);
        }
      }
      tc::tcgen05_fence_before();
      __syncthreads();
      if (ew) {
#pragma unroll
        for (int i = 0; i < NPT; ++i) dh_rec[i] = sD[lane * NG + warp * NPT + i];
      }
    }
    float dz[NPT][4];
    if (ew) {
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const float dh = dho[i] + dh_rec[i];
        const float tch = tanhf(cc[i]);
        const float d_o = dh * tch * asr::hard_sigmoid_grad(go[i]);
        const float dc = dc_carry[i] + dh * go[i] * (1.0f - tch * tch);
        dz[i][0] = dc * gg[i] * asr::hard_sigmoid_grad(gi[i]);
        dz[i][1] = dc * cp[i] * asr::hard_sigmoid_grad(gf[i]);
        dz[i][2] = dc * gi[i] * (1.0f - gg[i] * gg[i]);
        dz[i][3] = d_o;
        dc_carry[i] = dc * gf[i];
        const size_t row = (size_t)t * N + n0 + warp * NPT + i;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          db[g] += dz[i][g];
          dz16[(row * 2 + dir) * K4 + g * H + u] = __float2bfloat16_rn(dz[i][g]);   // exchange-critical
        }
      }
    }
    PROF(4);
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release(flag, 1);
    }
    PROF(5);
    if (ew) {
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const size_t row = (size_t)t * N + n0 + warp * NPT + i;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (a.dz32) a.dz32[(row * 2 + dir) * K4 + g * H + u] = dz[i][g];
          if (a.dzT16)
            reinterpret_cast<__nv_bfloat16*>(a.dzT16)[(size_t)(dir * K4 + g * H + u) * R + row] = __float2bfloat16_rn(dz[i][g]);
        }
      }
    }
    PROF(6);
  }
  PROF_DUMP(8);
  if (ew) {
#pragma unroll
    for (int g = 0; g < 4; ++g) atomicAdd(a.dbias + (size_t)dir * K4 + g * H + u, db[g]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// ---- host ---------------------------------------------------------------------------------------
constexpr int NG = 16;

static bool shape_ok(int T, int N, int H) {
  return T >= 1 && H >= 64 && H <= 512 && H % 64 == 0 && N >= NG && N % NG == 0 && N / NG <= 8 &&
         (H / UPC) * 2 * (N / NG) <= 148;
}
bool supports_fwd(const asr_lstm_fwd_args* a) { return a->U16 && a->h16 && !a->mask_u && shape_ok(a->T, a->N, a->H); }
bool supports_bwd(const asr_lstm_bwd_args* a) { return a->U16 && a->dz16 && !a->mask_u && shape_ok(a->T, a->N, a->H); }
size_t scratch_bytes(int) { return 8192; }

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  const int H = a->H, KC = H / 64;
  const size_t smem = 1024 + (size_t)KC * 128 * 128 + (size_t)KC * NG * 128 + 4 * NG * 32 * 4 + 64;
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, 4096, st));
  asr_lstm_fwd_args args = *a;
  int* flags = a->flags;
  void* kargs[] = {&args, &flags};
  ASR_CUDA(cudaLaunchCooperativeKernel((void*)fwd_kernel<NG>, dim3(H / UPC, 2, a->N / NG), dim3(128), kargs, smem, st));
  asr::count_launch();
  return ASR_OK;
}

int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st) {
  const int H = a->H, KC = 4 * H / 64;
  const size_t smem = 1024 + (size_t)KC * 32 * 128 + (size_t)KC * NG * 128 + 32 * NG * 4 + 64;
  ASR_CUDA(cudaFuncSetAttribute(bwd_kernel<NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, 4096, st));
  ASR_CUDA(cudaMemsetAsync(a->dbias, 0, (size_t)2 * 4 * H * sizeof(float), st));
  asr_lstm_bwd_args args = *a;
  int* flags = a->flags;
  void* kargs[] = {&args, &flags};
  ASR_CUDA(cudaLaunchCooperativeKernel((void*)bwd_kernel<NG>, dim3(H / UPC, 2, a->N / NG), dim3(256), kargs, smem, st));
  asr::count_launch();
  return ASR_OK;
}

}  // namespace lstmtc
