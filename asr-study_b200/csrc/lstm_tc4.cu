// K3/K4 (tensor-core engine, v4) — the persistent BiLSTM recurrences of lstm_tc2.cu with every per-step tile that is
// NOT the exchange moved onto the TMA engine and stored in 16 bits:
//   * inputs  (forward: zx_t; BPTT: gates_t, c_t, dL/dh_t) arrive through a 4-deep ring of cp.async.bulk.tensor loads
//     issued S - 1 steps ahead by a dedicated DMA warp: no HBM-latency load ever sits in the load/store queue in front
//     of the LL-ring polls of the exchange;
//   * outputs (forward: gates, cell, the fp16 layer output / its two B_W-masked copies and the transposed bf16 copies;
//     BPTT: dz and dz^T) are written to a double-buffered shared-memory staging area and leave as TMA tile stores from
//     the DMA warp: the compute warps issue no global store except the 8-byte LL words of the exchange;
//   * zx, the saved gates and the cell state are fp16 in HBM (the forward's only large read and both recurrences' largest
//     write / read: 8.2 GB -> 4.3 GB per C2 step); the arithmetic stays fp32.
// The exchange itself (LL ring through L2, per-warp MMA issue, U resident in TMEM) is lstm_tc2.cu's, unchanged — see the
// measurements in profiles/lstm_phases_r2.md for what was tried on it this round.
//
// Grid = (H/32 CTAs, 2 directions, batch groups), cooperative launch; forward: 288 threads (eight compute warps + the DMA
// warp), BPTT: 160 threads (four compute warps, one TMEM lane quarter each, + the DMA warp).
// Semantics: core/layers.py:432-469 under Keras-1 Bidirectional, default branch + variational dropout masks.
#include "common.cuh"
#include "tc.cuh"
#include <mutex>
#include <cudaTypedefs.h>

namespace lstmtc4 {

constexpr int UPC = 32;
constexpr int NM = 16;
constexpr int CTHREADS = 128;                              // compute threads (4 warps)
constexpr int BTHREADS = 192;                              // BPTT: 4 compute warps + the DMA warp + a warp that only issues the MMAs
constexpr int STATUS_IDX = 64;
constexpr int HEADER_BYTES = 8192;
constexpr uint32_t D_COL = 0, A_COL = 64;
constexpr long long WATCHDOG_CYCLES = 2000000000LL;
constexpr int S = 4;                                       // input ring depth (loads run S - 1 steps ahead)

#ifdef ASR_LSTM_PROFILE
// per-phase clock64() accumulators of thread 0 of CTA (0,0,0), dumped behind the header (profiles/prof_lstm_phases_r2.py)
#define PROF_DECL long long pt0 = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { const long long now = clock64(); pacc[i] += now - pt0; pt0 = now; } while (0)
#define PROF_DUMP(base) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) for (int i = 0; i < 8; ++i) reinterpret_cast<long long*>(flags + 1024)[(base) + i] = pacc[i]; } while (0)
#else
#define PROF_DECL
#define PROF(i)
#define PROF_DUMP(base)
#endif

struct FwdMaps { CUtensorMap zx, gates, cell, h, hm, hT, hmT, hTu; };
struct BwdMaps { CUtensorMap gates, cell, dh, dh2, dz, dzT; };

__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {      // two LL words per access
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_v2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// fp16 image of a hard_sigmoid gate that keeps its saturation state: BPTT reads d hard_sigmoid = 0.2 on the open
// interval (0, 1) and 0 on the clips off the STORED value, so a gate of 0.9999 must not round to 1.0 (nor 1e-9 to 0):
// inside the interval the image is clamped to the nearest fp16 inside it (0x3BFF = 0.99951, 0x0001 = 6e-8)
__device__ __forceinline__ __half gate_to_half(float g) {
  __half h = __float2half_rn(g);
  const unsigned short bits = __half_as_ushort(h);
  if (g < 1.0f && bits == 0x3C00u) h = __ushort_as_half((unsigned short)0x3BFFu);
  if (g > 0.0f && bits == 0x0000u) h = __ushort_as_half((unsigned short)0x0001u);
  return h;
}
__device__ __forceinline__ void spin_until(long long t_end) {
  while (clock64() < t_end) {}
}

// ------------------------------------------------------------------------------------------------------------------
// forward: EIGHT compute warps (+ the DMA warp).  The per-step work of a CTA that sits on the chain between "the last LL
// word landed" and "my words are published" — staging the B operand, the accumulator read-out, the gate math, the LL
// sends — is per-thread work of a handful of dependent instructions; with four warps (one per scheduler) it issues at one
// instruction per ~5 cycles.  Eight warps halve every thread's share (one sample per thread at NB = 8) and give each
// scheduler a second warp to hide latencies under: 1.56 -> 1.44 ms per layer at C2 (profiles/lstm_phases_r2.md).  Warp w:
// TMEM lane quarter q = w & 3 (hardware restriction), sample half hs = w >> 2; eight accumulators, one per issuing warp.
// (The BPTT kernel below stays at four compute warps: its eight-warp variant measured 3-6 % slower — the reduce-scatter
// receive then needs a cross-warp combine — and was removed.)
// ------------------------------------------------------------------------------------------------------------------
constexpr int CTHREADS8 = 256;
constexpr int THREADS8 = 288;
constexpr uint32_t A_COL8 = 128;                           // eight accumulators of NM columns in front of the U slice

template <int H, int NB>
__global__ void __launch_bounds__(THREADS8, 1)
fwd_kernel(asr_lstm_fwd_args a, int* __restrict__ flags, uint2* __restrict__ xbuf, int grp0, int probe_delay,
            const __grid_constant__ FwdMaps M) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int KC = H / 64;
  constexpr int B_CHUNK = NM * 128;
  constexpr int WORDS = NB * H / 2;                      // LL words per (dir, group, parity)
  constexpr int NPT = NB / 8;                            // samples per thread
  constexpr int NH = NB / 2;                             // samples per accumulator read-out half
  constexpr int TILE = NB * 64;
  static_assert(H % 128 == 0 && H <= 768 && (NPT == 1 || NPT == 2), "shape");
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int u0 = cta * UPC, n0 = (grp0 + grp) * NB;

  uint8_t* sB = smem;
  float* sZ = reinterpret_cast<float*>(sB + KC * B_CHUNK);   // [4 gates][NB][32 units]
  uint8_t* ring = reinterpret_cast<uint8_t*>(sZ + 4 * NB * 32);
  uint8_t* stage = ring + S * 4 * TILE;
  constexpr int STAGE_BYTES = 12 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * STAGE_BYTES);
  uint64_t* mma_bar = bars;
  uint64_t* full = bars + 1;
  uint64_t* sfull = full + S;
  uint64_t* sfree = sfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree + 2);
  __shared__ volatile int s_dead;

  if (tid == 0) {
    tc::mbar_init(mma_bar, 8);
    for (int k = 0; k < S; ++k) tc::mbar_init(full + k, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(sfull + b, CTHREADS8); tc::mbar_init(sfree + b, 1); }
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < KC * B_CHUNK / 16; i += THREADS8) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  if (*tmem_slot != 0u) { if (tid == 0) atomicExch(flags + STATUS_IDX, 2); }
  constexpr uint32_t tmem = 0u;
  int* status = flags + STATUS_IDX;
  const int q = warp & 3, hs = (warp >> 2) & 1;

  if (warp < 8) {
    // one-time: U^T slice -> TMEM; warp (q, hs) writes column half hs of lane quarter q
    const uint4* urow = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(a.U16) +
                                                       ((size_t)dir * 4 * H + (size_t)q * H + u0 + lane) * H);
#pragma unroll 1
    for (int c = hs * (H / 4); c < (hs + 1) * (H / 4); c += 32) {
      uint32_t r[32];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint4 v = __ldg(urow + c / 4 + k);
        r[4 * k] = v.x; r[4 * k + 1] = v.y; r[4 * k + 2] = v.z; r[4 * k + 3] = v.w;
      }
      tc::tmem_st32(tmem + ((uint32_t)(q * 32) << 16) + A_COL8 + c, r);
    }
    tc::tmem_st_wait();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();

  if (warp == 8) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&M.zx);
      auto load = [&](int k) {
        const int slot = k % S, tk = dir ? (T - 1 - k) : k;
        tc::mbar_expect_tx(full + slot, 4 * TILE);
        tc::tma_load_4d(ring + slot * 4 * TILE, &M.zx, full + slot, u0, 0, dir, tk * N + n0);
      };
      for (int k = 0; k < S && k < T; ++k) load(k);
      for (int s = 0; s < T; ++s) {
        const int b = s & 1;
        if (!tc::mbar_wait(sfull + b, (uint32_t)((s >> 1) & 1), WATCHDOG_CYCLES) || s_dead) {
          atomicExch(status, 1);
          s_dead = 1;
          break;
        }
        const int t = dir ? (T - 1 - s) : s;
        const int row0 = t * N + n0, col0 = dir * H + u0;
        uint8_t* st = stage + b * STAGE_BYTES;
        if (a.training) {
          tc::tma_store_4d(&M.gates, st, u0, 0, dir, row0);
          tc::tma_store_3d(&M.cell, st + 4 * TILE, u0, dir, row0);
        }
        if (a.h16) tc::tma_store_2d(&M.h, st + 5 * TILE, col0, row0);
        if (a.hm16) {
          tc::tma_store_3d(&M.hm, st + 6 * TILE, col0, row0, 0);
          tc::tma_store_3d(&M.hm, st + 7 * TILE, col0, row0, 1);
        }
        if (a.training && a.hT16) tc::tma_store_2d(&M.hT, st + 8 * TILE, row0, col0);
        if (a.training && a.hmT16) {
          tc::tma_store_3d(&M.hmT, st + 9 * TILE, row0, col0, 0);
          tc::tma_store_3d(&M.hmT, st + 10 * TILE, row0, col0, 1);
        }
        if (a.training && a.hT16u) tc::tma_store_2d(&M.hTu, st + 11 * TILE, row0, col0);
        tc::bulk_commit();
        if (s + S < T) load(s + S);
        tc::bulk_wait_read<0>();
        tc::mbar_arrive(sfree + b);
      }
      tc::bulk_wait<0>();
    }
  } else {
    const uint32_t idesc = tc::umma_idesc_f16(128, NM, 0);
    const uint32_t sB_addr = tc::smem_u32(sB);
    const int u = u0 + lane;
    float bias[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
    float c_state[NPT], mu[NPT], mn0[NPT], mn1[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const int n = warp * NPT + i;
      c_state[i] = 0.0f;
      mu[i] = a.mask_u ? a.mask_u[((size_t)dir * N + n0 + n) * H + u] : 1.0f;
      mn0[i] = a.mask_next ? a.mask_next[((size_t)0 * N + n0 + n) * 2 * H + dir * H + u] : 1.0f;
      mn1[i] = a.mask_next ? a.mask_next[((size_t)1 * N + n0 + n) * 2 * H + dir * H + u] : 1.0f;
    }
    uint2* xb = xbuf + (size_t)(dir * G + grp) * 2 * WORDS;
    long long t_pub = clock64();
    PROF_DECL;

    for (int s = 0; s < T; ++s) {
      PROF(7);
      if (!tc::mbar_wait(full + (s % S), (uint32_t)((s / S) & 1), WATCHDOG_CYCLES)) {
        atomicExch(status, 1);
        s_dead = 1;
        break;
      }
      float zxv[NPT][4];
      {
        const __half* zt = reinterpret_cast<const __half*>(ring + (s % S) * 4 * TILE);
#pragma unroll
        for (int i = 0; i < NPT; ++i)
#pragma unroll
          for (int g = 0; g < 4; ++g) zxv[i][g] = __half2float(zt[((warp * NPT + i) * 4 + g) * 32 + lane]) + bias[g];
      }
      PROF(0);
      float z[NPT][4];
      if (s > 0) {
        if (probe_delay > 0) spin_until(t_pub + probe_delay);
        // warp w polls (and stages) the K range [H/8 * w, H/8 * (w + 1)) of all NB samples
        const uint4* src = reinterpret_cast<const uint4*>(xb + (size_t)((s - 1) & 1) * WORDS);
        const uint32_t tag = (uint32_t)s;
        constexpr int V4W = H / 32;                        // 16-byte accesses (4 K columns) per (sample, warp K eighth)
        constexpr int QPT = NB * V4W / 32;
        static_assert((NB * V4W) % 32 == 0, "per-warp poll set must fill whole warp accesses");
        int vidx[QPT];
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
          const int f = k * 32 + lane;
          vidx[k] = (f / V4W) * (H / 4) + warp * V4W + (f % V4W);
        }
        uint4 w[QPT];
#pragma unroll
        for (int k = 0; k < QPT; ++k) w[k] = ld_volatile_v4(src + vidx[k]);
        bool ok;
        long long t0 = 0;
        do {
          ok = true;
#pragma unroll
          for (int k = 0; k < QPT; ++k)
            if (w[k].y != tag || w[k].w != tag) {
              w[k] = ld_volatile_v4(src + vidx[k]);
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
        PROF(1);
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
          const int f = k * 32 + lane;
          const int n = f / V4W, kk = 4 * (warp * V4W + (f % V4W));
          *reinterpret_cast<uint2*>(sB + (kk >> 6) * B_CHUNK + tc::sw128_offset(n, kk & 63)) = make_uint2(w[k].x, w[k].z);
        }
        tc::fence_proxy_async_smem();
        __syncwarp();
        if (tc::elect_one_sync()) {
          tc::tcgen05_fence_after();
          constexpr int KBW = H / 16 / 8;                    // MMAs (K = 16 each) per warp
#pragma unroll
          for (int j = 0; j < KBW; ++j) {
            const int kb = warp * KBW + j;
            const uint64_t bd = tc::umma_desc_sw128(sB_addr + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
            tc::umma_ts(tmem + D_COL + warp * NM, tmem + A_COL8 + kb * 8, bd, idesc, j > 0);
          }
          tc::umma_commit(mma_bar);
        }
        PROF(2);
        if (!tc::mbar_wait(mma_bar, (uint32_t)((s - 1) & 1), WATCHDOG_CYCLES)) {
          atomicExch(status, 1);
          s_dead = 1;
        }
        tc::tcgen05_fence_after();
        PROF(3);
        {
          // lane quarter q of all eight accumulators, sample half hs
          uint32_t r[8][NH];
          const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + D_COL + hs * NH;
#pragma unroll
          for (int k = 0; k < 8; ++k) tc::tmem_ldn(tq + k * NM, r[k]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int n = 0; n < NH; ++n) {
            const float v = ((__uint_as_float(r[0][n]) + __uint_as_float(r[1][n])) + (__uint_as_float(r[2][n]) + __uint_as_float(r[3][n]))) +
                            ((__uint_as_float(r[4][n]) + __uint_as_float(r[5][n])) + (__uint_as_float(r[6][n]) + __uint_as_float(r[7][n])));
            sZ[(q * NB + hs * NH + n) * 32 + lane] = v;
          }
        }
        tc::tcgen05_fence_before();
        tc::named_bar_sync(1, CTHREADS8);
        if (s_dead) break;
#pragma unroll
        for (int i = 0; i < NPT; ++i)
#pragma unroll
          for (int g = 0; g < 4; ++g) z[i][g] = sZ[(g * NB + warp * NPT + i) * 32 + lane];
        PROF(4);
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
        const int n = warp * NPT + i;
        float pre[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) pre[g] = z[i][g] + zxv[i][g];
        gi[i] = asr::hard_sigmoid(pre[0]);
        gf[i] = asr::hard_sigmoid(pre[1]);
        gg[i] = asr::tanh_fast(pre[2]);
        go[i] = asr::hard_sigmoid(pre[3]);
        c_state[i] = gf[i] * c_state[i] + gi[i] * gg[i];
        hv[i] = go[i] * asr::tanh_fast(c_state[i]);
        const float hm = hv[i] * mu[i];
        const float other = __shfl_down_sync(0xffffffffu, hm, 1);
        if (!(lane & 1)) {
          const __half2 pk = __floats2half2_rn(hm, other);
          uint2 wv;
          wv.x = *reinterpret_cast<const uint32_t*>(&pk);
          wv.y = (uint32_t)(s + 1);
          st_volatile_v2(xo + (size_t)n * (H / 2) + (u >> 1), wv);
        }
      }
      t_pub = clock64();
      PROF(5);
      const int b = s & 1;
      if (s >= 2 && !tc::mbar_wait(sfree + b, (uint32_t)(((s >> 1) - 1) & 1), WATCHDOG_CYCLES)) {
        atomicExch(status, 1);
        s_dead = 1;
        break;
      }
      uint8_t* st = stage + b * STAGE_BYTES;
      __half* gt = reinterpret_cast<__half*>(st);
      __half* ct = reinterpret_cast<__half*>(st + 4 * TILE);
      __half* ht = reinterpret_cast<__half*>(st + 5 * TILE);
      __half* hm0 = reinterpret_cast<__half*>(st + 6 * TILE);
      __half* hm1 = reinterpret_cast<__half*>(st + 7 * TILE);
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const int n = warp * NPT + i;
        if (a.training) {
          gt[(n * 4 + 0) * 32 + lane] = gate_to_half(gi[i]);
          gt[(n * 4 + 1) * 32 + lane] = gate_to_half(gf[i]);
          gt[(n * 4 + 2) * 32 + lane] = __float2half_rn(gg[i]);
          gt[(n * 4 + 3) * 32 + lane] = gate_to_half(go[i]);
          ct[n * 32 + lane] = __float2half_rn(c_state[i]);
        }
        if (a.h16) ht[n * 32 + lane] = __float2half_rn(hv[i]);
        if (a.hm16) {
          hm0[n * 32 + lane] = __float2half_rn(hv[i] * mn0[i]);
          hm1[n * 32 + lane] = __float2half_rn(hv[i] * mn1[i]);
        }
      }
      auto storeT = [&](uint8_t* tile, const float (&m)[NPT]) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(tile) + lane * NB + warp * NPT;
        if constexpr (NPT == 2) {
          *reinterpret_cast<__nv_bfloat162*>(dst) = __floats2bfloat162_rn(hv[0] * m[0], hv[1] * m[1]);
        } else {
          *dst = __float2bfloat16_rn(hv[0] * m[0]);
        }
      };
      if (a.training && a.hT16) storeT(st + 8 * TILE, mu);
      if (a.training && a.hmT16) {
        storeT(st + 9 * TILE, mn0);
        storeT(st + 10 * TILE, mn1);
      }
      if (a.training && a.hT16u) {
        float one[NPT];
#pragma unroll
        for (int i = 0; i < NPT; ++i) one[i] = 1.0f;
        storeT(st + 11 * TILE, one);
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(sfull + b);
      PROF(6);
    }
    PROF_DUMP(0);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// backward through time: the single-exchange kernel of lstm_tc2.cu (bwd3: CTA j multiplies its own dz by its
// [H units x 128 own gate columns] slice of U in TMEM and reduce-scatters bf16 partials through the LL ring), H <= 512
// ------------------------------------------------------------------------------------------------------------------
template <int H, int NB>
__global__ void __launch_bounds__(BTHREADS, 1)
bwd_kernel(asr_lstm_bwd_args a, int* __restrict__ flags, uint2* __restrict__ xbuf, int grp0, int probe_delay,
           const __grid_constant__ BwdMaps M) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int K4 = 4 * H, NCTA = H / UPC;
  constexpr int NBLK = H / 128;
  constexpr int HB = (NBLK + 1) / 2;                     // the accumulator blocks are committed in two halves
  static_assert(H % 128 == 0 && NBLK >= 1 && NBLK <= 4, "U slice: NBLK x 64 TMEM columns, one accumulator per block");
  constexpr int B_CHUNK = NM * 128;
  constexpr int NPT = NB / 4, PPT = NPT / 2, NP = NB / 2;
  constexpr int SLOT = NP * NCTA * 32;                   // LL words per (parity, destination): [pair][source][unit]
  constexpr int TILE = NB * 64;                          // bytes of one [NB x 32] 16-bit tile
  constexpr int IN_BYTES = 4 * TILE + TILE + 2 * TILE + 2 * TILE;   // gates, cell, dh (fp32), dh2 (fp32)
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int u0 = cta * UPC, n0 = (grp0 + grp) * NB;

  uint8_t* sB = smem;                                    // 2 chunks of [NM rows x 128 B]: K = 128 own gate columns
  uint8_t* ring = sB + 2 * B_CHUNK;                      // S x {gates [NB][4][32] f16, cell [NB][32] f16, dh / dh2 [NB][32] f32}
  uint8_t* stage = ring + S * IN_BYTES;                  // 2 x {dz [NB][4][32] bf16, dzT [4][32][NB] bf16}
  constexpr int STAGE_BYTES = 8 * TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage + 2 * STAGE_BYTES);
  uint64_t* mma_bar = bars;                              // [4], two used: the first / second half of the M blocks
  uint64_t* bready = bars + 4;                           // the B operand (dz_t of all samples) is staged: one arrival per compute warp
  uint64_t* full = bars + 5;
  uint64_t* sfull = full + S;
  uint64_t* sfree = sfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sfree + 2);
  __shared__ volatile int s_dead;

  if (tid == 0) {
    for (int b = 0; b < 4; ++b) tc::mbar_init(mma_bar + b, 1);
    tc::mbar_init(bready, 4);
    for (int k = 0; k < S; ++k) tc::mbar_init(full + k, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(sfull + b, CTHREADS); tc::mbar_init(sfree + b, 1); }
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * B_CHUNK / 16; i += BTHREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  if (*tmem_slot != 0u) { if (tid == 0) atomicExch(flags + STATUS_IDX, 2); }
  constexpr uint32_t tmem = 0u;
  int* status = flags + STATUS_IDX;

  if (warp < 4) {
    // one-time: U slice -> TMEM.  block b, lane m <-> unit 128b + m; K index k = g*32 + i <-> gate column g*H + u0 + i
    const int ctid = tid;
#pragma unroll 1
    for (int b = 0; b < NBLK; ++b) {
      const __nv_bfloat16* Ub = reinterpret_cast<const __nv_bfloat16*>(a.U16) + (size_t)dir * H * K4 +
                                (size_t)(128 * b + ctid) * K4 + u0;
#pragma unroll 1
      for (int hs = 0; hs < 2; ++hs) {
        uint32_t rr[32];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const uint4* src = reinterpret_cast<const uint4*>(Ub + (2 * hs + gg) * H);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = __ldg(src + q);
            rr[gg * 16 + 4 * q] = v.x; rr[gg * 16 + 4 * q + 1] = v.y; rr[gg * 16 + 4 * q + 2] = v.z; rr[gg * 16 + 4 * q + 3] = v.w;
          }
        }
        tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + A_COL + b * 64 + hs * 32, rr);
      }
    }
    tc::tmem_st_wait();
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();

  const bool has_dh2 = a.dh2 != nullptr;
  if (warp == 4) {
    if (lane == 0) {
      tc::tma_prefetch_desc(&M.gates);
      tc::tma_prefetch_desc(&M.cell);
      tc::tma_prefetch_desc(&M.dh);
      auto load = [&](int k) {
        const int slot = k % S, tk = dir ? k : (T - 1 - k);
        const int row0 = tk * N + n0;
        uint8_t* dst = ring + slot * IN_BYTES;
        tc::mbar_expect_tx(full + slot, (uint32_t)(4 * TILE + TILE + 2 * TILE + (has_dh2 ? 2 * TILE : 0)));
        tc::tma_load_4d(dst, &M.gates, full + slot, u0, 0, dir, row0);
        tc::tma_load_3d(dst + 4 * TILE, &M.cell, full + slot, u0, dir, row0);
        tc::tma_load_2d(dst + 5 * TILE, &M.dh, full + slot, dir * H + u0, row0);
        if (has_dh2) tc::tma_load_2d(dst + 7 * TILE, &M.dh2, full + slot, dir * H + u0, row0);
      };
      for (int k = 0; k < S - 1 && k < T; ++k) load(k);    // slot (s + 1) % S also holds c_{prev} of step s: S - 1 in flight
      for (int s = 0; s < T; ++s) {
        const int b = s & 1;
        // step s reads slots s % S and (s + 1) % S; slot (s - 1) % S = (s + S - 1) % S is free once step s - 1 is done
        // (released by sfull of step s - 1, waited for in the previous iteration; at s = 0 the slot was never used)
        if (s + S - 1 < T) load(s + S - 1);
        if (!tc::mbar_wait(sfull + b, (uint32_t)((s >> 1) & 1), WATCHDOG_CYCLES) || s_dead) {
          atomicExch(status, 1);
          s_dead = 1;
          break;
        }
        const int t = dir ? s : (T - 1 - s);
        const int row0 = t * N + n0;
        uint8_t* st = stage + b * STAGE_BYTES;
        tc::tma_store_4d(&M.dz, st, u0, 0, dir, row0);
        if (a.dzT16) tc::tma_store_3d(&M.dzT, st + 4 * TILE, row0, u0, dir * 4);
        tc::bulk_commit();
        tc::bulk_wait_read<0>();
        tc::mbar_arrive(sfree + b);
      }
      tc::bulk_wait<0>();
    }
  } else if (warp == 5) {
    // MMA issue, off the compute warps: P = U[:, own gate columns] . dz_t, M block by M block (block b = the rows of CTAs
    // 4b .. 4b+3), one tcgen05.commit per block, so the compute warps read out and SEND block b while the tensor pipe is
    // on block b + 1 — the send of the last block is all that follows the last MMA
    const uint32_t idesc = tc::umma_idesc_f16(128, NM, 1);
    const uint32_t sB_addr = tc::smem_u32(sB);
    for (int s = 0; s + 1 < T; ++s) {
      if (!tc::mbar_wait(bready, (uint32_t)(s & 1), WATCHDOG_CYCLES) || s_dead) {
        atomicExch(status, 1);
        s_dead = 1;
        break;
      }
      if (tc::elect_one_sync()) {
        tc::tcgen05_fence_after();
#pragma unroll
        for (int b = 0; b < NBLK; ++b) {
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const uint64_t bd = tc::umma_desc_sw128(sB_addr + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
            tc::umma_ts(tmem + D_COL + b * NM, tmem + A_COL + b * 64 + kb * 8, bd, idesc, kb > 0);
          }
          if (b == HB - 1) tc::umma_commit(mma_bar + 0);          // first half of the blocks: read out and sent under the second
          else if (b == NBLK - 1) tc::umma_commit(mma_bar + 1);
        }
      }
      __syncwarp();
    }
  } else {
    const int u = u0 + lane;
    float dc_carry[NPT], mu[NPT], md0[NPT], md1[NPT], db[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      dc_carry[i] = 0.0f;
      mu[i] = a.mask_u ? a.mask_u[((size_t)dir * N + n0 + warp * NPT + i) * H + u] : 1.0f;
      md0[i] = a.mask_dh ? a.mask_dh[((size_t)0 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
      md1[i] = a.mask_dh ? a.mask_dh[((size_t)1 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
    }
    uint2* xb = xbuf + (size_t)(dir * G + grp) * 2 * NCTA * SLOT;     // [(dir,grp)][parity][dst][pair][src][unit]
    long long t_pub = clock64();
    PROF_DECL;

    for (int s = 0; s < T; ++s) {
      PROF(7);
      // inputs of this step (slot s % S) and c of the step the forward pass ran BEFORE it (= the next BPTT step, slot
      // (s + 1) % S): into registers now, under the flight time of the exchange
      if (!tc::mbar_wait(full + (s % S), (uint32_t)((s / S) & 1), WATCHDOG_CYCLES) ||
          (s + 1 < T && !tc::mbar_wait(full + ((s + 1) % S), (uint32_t)(((s + 1) / S) & 1), WATCHDOG_CYCLES))) {
        atomicExch(status, 1);
        s_dead = 1;
        break;
      }
      float gi[NPT], gf[NPT], gg[NPT], go[NPT], cc[NPT], cp[NPT], dhx[NPT];
      {
        const uint8_t* in = ring + (s % S) * IN_BYTES;
        const __half* gt = reinterpret_cast<const __half*>(in);
        const __half* ct = reinterpret_cast<const __half*>(in + 4 * TILE);
        const float* dht = reinterpret_cast<const float*>(in + 5 * TILE);
        const float* dh2t = reinterpret_cast<const float*>(in + 7 * TILE);
        const __half* cpt = reinterpret_cast<const __half*>(ring + ((s + 1) % S) * IN_BYTES + 4 * TILE);
#pragma unroll
        for (int i = 0; i < NPT; ++i) {
          const int n = warp * NPT + i;
          gi[i] = __half2float(gt[(n * 4 + 0) * 32 + lane]);
          gf[i] = __half2float(gt[(n * 4 + 1) * 32 + lane]);
          gg[i] = __half2float(gt[(n * 4 + 2) * 32 + lane]);
          go[i] = __half2float(gt[(n * 4 + 3) * 32 + lane]);
          cc[i] = __half2float(ct[n * 32 + lane]);
          cp[i] = (s + 1 < T) ? __half2float(cpt[n * 32 + lane]) : 0.0f;
          const float dho = dht[n * 32 + lane];
          const float dho2 = has_dh2 ? dh2t[n * 32 + lane] : 0.0f;
          dhx[i] = fmaf(dho2, md1[i], dho * md0[i]);         // dL/d(output): the two masked dX partials of the layer above
        }
      }
      PROF(0);
      float dh_rec[NPT];
#pragma unroll
      for (int i = 0; i < NPT; ++i) dh_rec[i] = 0.0f;
      if (s > 0) {
        if (probe_delay > 0) spin_until(t_pub + probe_delay);
        const uint32_t tag = (uint32_t)s;
        const uint2* src = xb + ((size_t)((s - 1) & 1) * NCTA + cta) * SLOT + (size_t)(warp * PPT) * NCTA * 32 + lane;
        uint2 w[PPT * NCTA];
#pragma unroll
        for (int q = 0; q < PPT * NCTA; ++q) w[q] = ld_volatile_v2(src + q * 32);
        bool ok;
        long long t0 = 0;
        do {
          ok = true;
#pragma unroll
          for (int q = 0; q < PPT * NCTA; ++q)
            if (w[q].y != tag) {
              w[q] = ld_volatile_v2(src + q * 32);
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
        for (int pp = 0; pp < PPT; ++pp) {
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int j = 0; j < NCTA; ++j) {
            const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(&w[pp * NCTA + j].x);
            s0 += __low2float(v);
            s1 += __high2float(v);
          }
          dh_rec[2 * pp] = mu[2 * pp] * s0;
          dh_rec[2 * pp + 1] = mu[2 * pp + 1] * s1;
        }
      }
      PROF(1);
      float dz[NPT][4];
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const int n = warp * NPT + i;
        const float dh = dhx[i] + dh_rec[i];
        const float tch = asr::tanh_fast(cc[i]);
        const float d_o = dh * tch * asr::hard_sigmoid_grad(go[i]);
        const float dc = dc_carry[i] + dh * go[i] * (1.0f - tch * tch);
        dz[i][0] = dc * gg[i] * asr::hard_sigmoid_grad(gi[i]);
        dz[i][1] = dc * cp[i] * asr::hard_sigmoid_grad(gf[i]);
        dz[i][2] = dc * gi[i] * (1.0f - gg[i] * gg[i]);
        dz[i][3] = d_o;
        dc_carry[i] = dc * gf[i];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int k = g * 32 + lane;                       // K index inside my 128 gate columns
          *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 6) * B_CHUNK + tc::sw128_offset(n, k & 63)) = __float2bfloat16_rn(dz[i][g]);
          db[g] += dz[i][g];
        }
      }
      tc::fence_proxy_async_smem();
      PROF(2);
      if (s + 1 < T) {
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(bready);            // my warp's samples of the B operand are in shared memory
        // send: my warp's rows of block b belong to CTA 4b + warp
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16) + D_COL;
        const uint32_t tg = (uint32_t)(s + 1);
        uint2* out = xb + (size_t)(s & 1) * NCTA * SLOT + (size_t)cta * 32 + lane;      // + dst*SLOT + pair*NCTA*32
#pragma unroll
        for (int half = 0; half < (NBLK > 1 ? 2 : 1); ++half) {
          const int b0 = half ? HB : 0, b1 = half ? NBLK : HB;
          if (!tc::mbar_wait(mma_bar + half, (uint32_t)(s & 1), WATCHDOG_CYCLES)) {
            atomicExch(status, 1);
            s_dead = 1;
          }
          tc::tcgen05_fence_after();
          if (half == 0) { PROF(3); }
          uint32_t rb[HB][NB];
#pragma unroll
          for (int bb = b0; bb < b1; ++bb) tc::tmem_ldn(tq + bb * NM, rb[bb - b0]);
          tc::tmem_ld_wait();
#pragma unroll
          for (int np = 0; np < NP; ++np) {
#pragma unroll
            for (int bb = b0; bb < b1; ++bb) {
              const __nv_bfloat162 q = __floats2bfloat162_rn(__uint_as_float(rb[bb - b0][2 * np]), __uint_as_float(rb[bb - b0][2 * np + 1]));
              st_volatile_v2(out + (size_t)(bb * 4 + warp) * SLOT + (size_t)np * NCTA * 32,
                             make_uint2(*reinterpret_cast<const uint32_t*>(&q), tg));
            }
          }
        }
        if (s_dead) break;
        tc::tcgen05_fence_before();
        t_pub = clock64();
        PROF(4);
      }
      // side outputs (dz for the dW / dU / dX GEMMs) -> staging buffer s & 1, off the chain: after the send
      const int b = s & 1;
      if (s >= 2 && !tc::mbar_wait(sfree + b, (uint32_t)(((s >> 1) - 1) & 1), WATCHDOG_CYCLES)) {
        atomicExch(status, 1);
        s_dead = 1;
        break;
      }
      uint8_t* st = stage + b * STAGE_BYTES;
      __nv_bfloat16* dzt = reinterpret_cast<__nv_bfloat16*>(st);
      __nv_bfloat16* dzTt = reinterpret_cast<__nv_bfloat16*>(st + 4 * TILE);
#pragma unroll
      for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) dzt[((warp * NPT + i) * 4 + g) * 32 + lane] = __float2bfloat16_rn(dz[i][g]);
      if (a.dzT16) {                                       // [4 gates][32 units][NB samples]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          __nv_bfloat16* dst = dzTt + (g * 32 + lane) * NB + warp * NPT;
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(dz[0][g], dz[1][g]);
          if constexpr (NPT == 4) {
            const __nv_bfloat162 p1 = __floats2bfloat162_rn(dz[2][g], dz[3][g]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&p0);
            pk.y = *reinterpret_cast<const uint32_t*>(&p1);
            *reinterpret_cast<uint2*>(dst) = pk;
          } else {
            *reinterpret_cast<__nv_bfloat162*>(dst) = p0;
          }
        }
      }
      tc::fence_proxy_async_smem();
      tc::mbar_arrive(sfull + b);                          // dz staged; the input slots of this step were read long ago
      PROF(5);
    }
    PROF_DUMP(8);
#pragma unroll
    for (int g = 0; g < 4; ++g) atomicAdd(a.dbias + (size_t)dir * K4 + g * H + u, db[g]);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host ------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_once;

static bool encode(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                   const cuuint64_t* strides, const cuuint32_t* box) {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  });
  if (!g_encode) return false;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return g_encode(tm, dt, rank, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// [R][2][4][H] 16-bit (zx, gates, dz): box = 32 units x 4 gates x 1 direction x NB rows
static bool map_r24h(CUtensorMap* tm, const void* p, CUtensorMapDataType dt, int64_t R, int H, int NB) {
  const cuuint64_t dims[4] = {(cuuint64_t)H, 4, 2, (cuuint64_t)R};
  const cuuint64_t str[3] = {(cuuint64_t)H * 2, (cuuint64_t)4 * H * 2, (cuuint64_t)8 * H * 2};
  const cuuint32_t box[4] = {32, 4, 1, (cuuint32_t)NB};
  return encode(tm, dt, 4, p, dims, str, box);
}
// [R][2][H] fp16 (cell): box = 32 x 1 x NB
static bool map_r2h(CUtensorMap* tm, const void* p, int64_t R, int H, int NB) {
  const cuuint64_t dims[3] = {(cuuint64_t)H, 2, (cuuint64_t)R};
  const cuuint64_t str[2] = {(cuuint64_t)H * 2, (cuuint64_t)2 * H * 2};
  const cuuint32_t box[3] = {32, 1, (cuuint32_t)NB};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, p, dims, str, box);
}
// [copies][R][W] (h16, hm16, dh): box = 32 x NB (x 1)
static bool map_rows(CUtensorMap* tm, const void* p, CUtensorMapDataType dt, int esz, int64_t R, int W, int NB, int copies) {
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)R, (cuuint64_t)copies};
  const cuuint64_t str[2] = {(cuuint64_t)W * esz, (cuuint64_t)R * W * esz};
  const cuuint32_t box[3] = {32, (cuuint32_t)NB, 1};
  return encode(tm, dt, copies > 1 ? 3 : 2, p, dims, str, box);
}
// transposed bf16 [copies][W][R]: box = NB samples (contiguous) x 32 units (x 1 copy or `third` rows of a 3rd axis)
static bool map_T(CUtensorMap* tm, const void* p, int64_t R, int W, int NB, int copies) {
  const cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)W, (cuuint64_t)copies};
  const cuuint64_t str[2] = {(cuuint64_t)R * 2, (cuuint64_t)W * R * 2};
  const cuuint32_t box[3] = {(cuuint32_t)NB, 32, 1};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, copies > 1 ? 3 : 2, p, dims, str, box);
}
// dz^T bf16 [2*4][H][R] seen as (R, H, 8): box = NB x 32 units x 4 gates
static bool map_dzT(CUtensorMap* tm, const void* p, int64_t R, int H, int NB) {
  const cuuint64_t dims[3] = {(cuuint64_t)R, (cuuint64_t)H, 8};
  const cuuint64_t str[2] = {(cuuint64_t)R * 2, (cuuint64_t)H * R * 2};
  const cuuint32_t box[3] = {(cuuint32_t)NB, 32, 4};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, p, dims, str, box);
}

static int max_groups(int H) { return 148 / (2 * (H / UPC)); }
static int group_size(int N, int H, int flags) {
  if (flags & ASR_LSTM_GROUP16) return (N % 16 == 0) ? 16 : 0;
  if (N % 8 == 0 && N / 8 <= max_groups(H)) return 8;
  if (N % 16 == 0) return 16;
  return (N % 8 == 0) ? 8 : 0;
}
static bool width_ok(int H) { return H == 128 || H == 256 || H == 384 || H == 512; }
bool shape_supported(int T, int N, int H, int flags) {
  return T >= 1 && width_ok(H) && N >= 8 && group_size(N, H, flags) != 0 && ((int64_t)T * N) % 8 == 0;
}
bool supports_fwd(const asr_lstm_fwd_args* a) {
  return a->zx16 && a->U16 && (a->h16 || a->hm16) && !a->h32 && !a->mi && a->zoneout == 0.0f &&
         (!a->training || (a->gates16 && a->cell16)) && shape_supported(a->T, a->N, a->H, a->opts);
}
bool supports_bwd(const asr_lstm_bwd_args* a) {
  return a->gates16 && a->cell16 && a->U16 && a->dz16 && !a->dz32 && !a->mi && a->zoneout == 0.0f &&
         shape_supported(a->T, a->N, a->H, a->opts);
}
static size_t fwd_ring_bytes(int H, int NB, int G) { return (size_t)2 * G * 2 * NB * (H / 2) * sizeof(uint2); }
static size_t bwd_ring_bytes(int H, int NB, int G) {
  const size_t nc = H / UPC;
  return (size_t)2 * G * 2 * nc * (NB / 2) * nc * 32 * sizeof(uint2);
}
size_t scratch_bytes() {
  size_t m = 0;
  for (int H = 128; H <= 512; H += 128)
    for (int NB = 8; NB <= 16; NB += 8) {
      const size_t f = fwd_ring_bytes(H, NB, max_groups(H)), b = bwd_ring_bytes(H, NB, max_groups(H));
      m = f > m ? f : m;
      m = b > m ? b : m;
    }
  return HEADER_BYTES + m;
}
static size_t exclusive_smem(size_t need, int flags) {
  const size_t want = 200 * 1024;
  return (!(flags & ASR_LSTM_SHARED_SM) && need < want) ? want : need;
}
static int probe_delay_of(int flags, int dflt) {           // bits 16..27: cycles / 8 (0 = the default)
  const int v = (flags >> 16) & 0xFFF;
  return v ? (v == 0xFFF ? 0 : v * 8) : dflt;
}

template <int H, int NB>
static int32_t launch_fwd(const asr_lstm_fwd_args* a, cudaStream_t st) {
  constexpr int KC = H / 64, TILE = NB * 64;
  const size_t need = 1024 + (size_t)KC * NM * 128 + 4 * NB * 32 * 4 + (size_t)S * 4 * TILE + 2 * 12 * TILE + 128;
  const size_t smem = exclusive_smem(need, a->opts);
  const int Gall = a->N / NB, gm = max_groups(H);
  const int64_t R = (int64_t)a->T * a->N;
  FwdMaps M;
  memset(&M, 0, sizeof(M));
  const CUtensorMapDataType F16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  bool ok = map_r24h(&M.zx, a->zx16, F16, R, H, NB);
  if (a->training) ok = ok && map_r24h(&M.gates, a->gates16, F16, R, H, NB) && map_r2h(&M.cell, a->cell16, R, H, NB);
  if (a->h16) ok = ok && map_rows(&M.h, a->h16, F16, 2, R, 2 * H, NB, 1);
  if (a->hm16) ok = ok && map_rows(&M.hm, a->hm16, F16, 2, R, 2 * H, NB, 2);
  if (a->training && a->hT16) ok = ok && map_T(&M.hT, a->hT16, R, 2 * H, NB, 1);
  if (a->training && a->hmT16) ok = ok && map_T(&M.hmT, a->hmT16, R, 2 * H, NB, 2);
  if (a->training && a->hT16u) ok = ok && map_T(&M.hTu, a->hT16u, R, 2 * H, NB, 1);
  if (!ok) {
    asr::set_error("lstmtc4 forward: cuTensorMapEncodeTiled failed (T=%d N=%d H=%d)", a->T, a->N, H);
    return ASR_ERR_CUDA;
  }
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  asr_lstm_fwd_args args = *a;
  int* flags = a->flags;
  uint2* xbuf = reinterpret_cast<uint2*>(reinterpret_cast<char*>(a->flags) + HEADER_BYTES);
  int delay = probe_delay_of(a->opts, 0);
  for (int grp0 = 0; grp0 < Gall; grp0 += gm) {
    const int G = Gall - grp0 < gm ? Gall - grp0 : gm;
    ASR_CUDA(cudaMemsetAsync(a->flags, 0, HEADER_BYTES + fwd_ring_bytes(H, NB, G), st));
    void* kargs[] = {&args, &flags, &xbuf, &grp0, &delay, &M};
    ASR_CUDA(cudaLaunchCooperativeKernel((void*)fwd_kernel<H, NB>, dim3(H / UPC, 2, G), dim3(THREADS8), kargs, smem, st));
    asr::count_launch();
  }
  return ASR_OK;
}

template <int H, int NB>
static int32_t launch_bwd(const asr_lstm_bwd_args* a, cudaStream_t st) {
  constexpr int NCTA = H / UPC, TILE = NB * 64;
  const size_t need = 1024 + (size_t)2 * NM * 128 + (size_t)S * 9 * TILE + 2 * 8 * TILE + 128;
  const size_t smem = exclusive_smem(need, a->opts);
  const int Gall = a->N / NB, gm = max_groups(H);
  const int64_t R = (int64_t)a->T * a->N;
  BwdMaps M;
  memset(&M, 0, sizeof(M));
  bool ok = map_r24h(&M.gates, a->gates16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, R, H, NB) && map_r2h(&M.cell, a->cell16, R, H, NB) &&
            map_rows(&M.dh, a->dh, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, R, 2 * H, NB, 1) &&
            map_r24h(&M.dz, a->dz16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, R, H, NB);
  if (a->dh2) ok = ok && map_rows(&M.dh2, a->dh2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, R, 2 * H, NB, 1);
  if (a->dzT16) ok = ok && map_dzT(&M.dzT, a->dzT16, R, H, NB);
  if (!ok) {
    asr::set_error("lstmtc4 backward: cuTensorMapEncodeTiled failed (T=%d N=%d H=%d)", a->T, a->N, H);
    return ASR_ERR_CUDA;
  }
  ASR_CUDA(cudaFuncSetAttribute(bwd_kernel<H, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->dbias, 0, (size_t)2 * 4 * a->H * sizeof(float), st));
  asr_lstm_bwd_args args = *a;
  int* flags = a->flags;
  uint2* xbuf = reinterpret_cast<uint2*>(reinterpret_cast<char*>(a->flags) + HEADER_BYTES);
  int delay = probe_delay_of(a->opts, 0);
  for (int grp0 = 0; grp0 < Gall; grp0 += gm) {
    const int G = Gall - grp0 < gm ? Gall - grp0 : gm;
    ASR_CUDA(cudaMemsetAsync(a->flags, 0, HEADER_BYTES + bwd_ring_bytes(H, NB, G), st));
    void* kargs[] = {&args, &flags, &xbuf, &grp0, &delay, &M};
    ASR_CUDA(cudaLaunchCooperativeKernel((void*)bwd_kernel<H, NB>, dim3(NCTA, 2, G), dim3(BTHREADS), kargs, smem, st));
    asr::count_launch();
  }
  return ASR_OK;
}

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  const bool g8 = group_size(a->N, a->H, a->opts) == 8;
  switch (a->H) {
    case 128: return g8 ? launch_fwd<128, 8>(a, st) : launch_fwd<128, 16>(a, st);
    case 256: return g8 ? launch_fwd<256, 8>(a, st) : launch_fwd<256, 16>(a, st);
    case 384: return g8 ? launch_fwd<384, 8>(a, st) : launch_fwd<384, 16>(a, st);
    case 512: return g8 ? launch_fwd<512, 8>(a, st) : launch_fwd<512, 16>(a, st);
  }
  asr::set_error("lstmtc4: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st) {
  const bool g8 = group_size(a->N, a->H, a->opts) == 8;
  switch (a->H) {
    case 128: return g8 ? launch_bwd<128, 8>(a, st) : launch_bwd<128, 16>(a, st);
    case 256: return g8 ? launch_bwd<256, 8>(a, st) : launch_bwd<256, 16>(a, st);
    case 384: return g8 ? launch_bwd<384, 8>(a, st) : launch_bwd<384, 16>(a, st);
    case 512: return g8 ? launch_bwd<512, 8>(a, st) : launch_bwd<512, 16>(a, st);
  }
  asr::set_error("lstmtc4: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

}  // namespace lstmtc4
