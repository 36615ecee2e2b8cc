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
#include <stdlib.h>
#include <string.h>

namespace lstmtc2 {

#ifdef ASR_LSTM_PROFILE
#define PROF_DECL long long pt0 = clock64(), pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PROF(i) do { const long long now = clock64(); pacc[i] += now - pt0; pt0 = now; } while (0)
#define PROF_DUMP(base) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) for (int i = 0; i < 8; ++i) reinterpret_cast<long long*>(flags + 1024)[(base) + i] = pacc[i]; } while (0)
// finer split of the MMA issue phase (elected lane only): slots 16.. of the dump area
#define PROF2_DECL long long qacc[6] = {0, 0, 0, 0, 0, 0}, qt0 = 0
#define PROF2_START qt0 = clock64()
#define PROF2(i) do { const long long now = clock64(); qacc[i] += now - qt0; qt0 = now; } while (0)
#define PROF2_DUMP(base) do { if (tid == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) for (int i = 0; i < 6; ++i) reinterpret_cast<long long*>(flags + 1024)[(base) + i] = qacc[i]; } while (0)
#else
#define PROF2_DECL
#define PROF2_START
#define PROF2(i)
#define PROF2_DUMP(base)
#define PROF_DECL
#define PROF(i)
#define PROF_DUMP(base)
#endif

constexpr int UPC = 32;
constexpr int NM = 16;                                    // MMA N (M = 128 needs a multiple of 16); sample rows of the B operand
constexpr int THREADS = 128;
constexpr int STATUS_IDX = 64;
constexpr int HEADER_BYTES = 8192;
constexpr int NACC = 4;                                   // TMEM accumulators: one per warp (its K quarter), summed in the epilogue
static_assert(NACC == THREADS / 32, "one accumulator per warp");
constexpr uint32_t D_COL = 0, A_COL = 64;
constexpr long long WATCHDOG_CYCLES = 2000000000LL;

__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint4* p) {      // two LL words per access
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_v2(uint2* p, uint2 v) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
// NB = samples per CTA group (8 or 16).  NB = 8 spreads N = 32 over 128 CTAs instead of 64: the MMA still runs at
// N = 16 (rows 8..15 of the B operand stay zero, their accumulator columns are never read) and costs the same issue
// slots, while the LL exchange, the shared-memory staging and the gate math per CTA are halved.
// VAR = the element-wise brsmv1 switches (core/layers.py:441-467): multiplicative integration
// z = alpha*Wx*Uh + beta1*Uh + beta2*Wx + b and zoneout on c and h; off in the default instantiation.
template <int H, int NB, bool VAR>
__global__ void __launch_bounds__(THREADS, 1)
fwd_kernel(asr_lstm_fwd_args a, int* __restrict__ flags, uint2* __restrict__ xbuf, int delay1, int grp0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int KC = H / 64;
  constexpr int B_CHUNK = NM * 128;
  constexpr int WORDS = NB * H / 2;                      // LL words per (dir, group, parity)
  constexpr int NPT = NB / 4;
  static_assert(H % 64 == 0, "K chunks of 64; each warp owns H / 4 K columns = H / 64 MMAs");
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int u0 = cta * UPC, n0 = (grp0 + grp) * NB;      // grp0: first batch group of this launch (large batches: several launches)

  uint8_t* sB = smem;                                    // KC chunks of [NM rows x 128 B], SW128 K-major
  float* sZ = reinterpret_cast<float*>(sB + KC * B_CHUNK);   // [4 gates][NB][32 units]
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sZ + 4 * NB * 32);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  __shared__ int s_dead;

  if (tid == 0) {
    tc::mbar_init(mma_bar, 4);   // one tcgen05.commit per warp
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < KC * B_CHUNK / 16; i += THREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  // the CTA owns all 512 TMEM columns, so the allocation starts at column 0: use compile-time TMEM
  // addresses (a run-time base forces an ELECT/R2UR waterfall loop around every tcgen05.mma — measured
  // ~70 cycles per MMA, see profiles/lstm_phases_r1.md)
  if (*tmem_slot != 0u) { if (tid == 0) atomicExch(flags + STATUS_IDX, 2); }
  constexpr uint32_t tmem = 0u;

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

  const uint32_t idesc = tc::umma_idesc_f16(128, NM, 0);
  const uint32_t sB_addr = tc::smem_u32(sB);
  const int u = u0 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
  float mia[4], mib1[4], mib2[4];                        // multiplicative integration (VAR): alpha, beta1, beta2 per gate
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const size_t o = (size_t)dir * 4 * H + g * H + u;
    mia[g] = (VAR && a.mi) ? a.mi[o] : 0.0f;
    mib1[g] = (VAR && a.mi) ? a.mi[(size_t)8 * H + o] : 1.0f;
    mib2[g] = (VAR && a.mi) ? a.mi[(size_t)16 * H + o] : 1.0f;
  }
  const bool zone = VAR && a.zoneout > 0.0f;
  float h_prev[NPT];                                     // raw h_{t-1} for zoneout (the exchanged copy carries B_U)
  float c_state[NPT], mu[NPT], mn0[NPT], mn1[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    c_state[i] = 0.0f;
    h_prev[i] = 0.0f;
    mu[i] = a.mask_u ? a.mask_u[((size_t)dir * N + n0 + warp * NPT + i) * H + u] : 1.0f;   // B_U, constant over time
    // B_W of the NEXT layer's two directions (core/layers.py:439 applies it to this layer's output): the masked
    // operand copies are written here, as side stores, instead of by separate mask kernels
    mn0[i] = a.mask_next ? a.mask_next[((size_t)0 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
    mn1[i] = a.mask_next ? a.mask_next[((size_t)1 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
  }

  int* status = flags + STATUS_IDX;
  uint2* xb = xbuf + (size_t)(dir * G + grp) * 2 * WORDS;
  __half* h16 = reinterpret_cast<__half*>(a.h16);
  const size_t R = (size_t)T * N;

  // side outputs of a finished step (h16 for the next layer, saved activations, transposed copy); issued while the
  // NEXT step's LL poll loads are in flight, so they are off the recurrence's critical path
  auto side_stores = [&](int t, const float (&hv)[NPT], const float (&gi)[NPT], const float (&gf)[NPT],
                         const float (&gg)[NPT], const float (&go)[NPT], const float (&cs)[NPT]) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      if (h16) h16[row * 2 * H + dir * H + u] = __float2half_rn(hv[i]);
      if (a.h32) a.h32[row * 2 * H + dir * H + u] = hv[i];
      if (a.hm16) {
        __half* hm = reinterpret_cast<__half*>(a.hm16);
        hm[row * 2 * H + dir * H + u] = __float2half_rn(hv[i] * mn0[i]);
        hm[(R + row) * 2 * H + dir * H + u] = __float2half_rn(hv[i] * mn1[i]);
      }
      if (a.training) {
        float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gp[u] = gi[i]; gp[H + u] = gf[i]; gp[2 * H + u] = gg[i]; gp[3 * H + u] = go[i];
        a.cell[(row * 2 + dir) * H + u] = cs[i];
      }
    }
    // transposed bf16 copies: the thread's NPT samples are contiguous, one 4- or 8-byte store per copy
    static_assert(NPT == 4 || NPT == 2, "packed transposed store: 2 or 4 samples per thread");
    auto storeT = [&](void* base, size_t plane, const float (&m)[NPT]) {
      __nv_bfloat16* dstT = reinterpret_cast<__nv_bfloat16*>(base) + plane + (size_t)(dir * H + u) * R + (size_t)t * N + n0 + warp * NPT;
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(hv[0] * m[0], hv[1] * m[1]);
      if constexpr (NPT == 4) {
        const __nv_bfloat162 p1 = __floats2bfloat162_rn(hv[2] * m[2], hv[3] * m[3]);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&p0);
        pk.y = *reinterpret_cast<const uint32_t*>(&p1);
        *reinterpret_cast<uint2*>(dstT) = pk;
      } else {
        *reinterpret_cast<__nv_bfloat162*>(dstT) = p0;
      }
    };
    if (a.training && a.hT16) storeT(a.hT16, 0, mu);                       // h * B_U: the dU operand
    if (a.training && a.hmT16) {                                           // h * B_W(next layer, dir 0 / 1): its dW operands
      storeT(a.hmT16, 0, mn0);
      storeT(a.hmT16, (size_t)2 * H * R, mn1);
    }
    if (a.training && a.hT16u) {                                           // unmasked (the Dense kernel's dW operand)
      float one[NPT];
#pragma unroll
      for (int i = 0; i < NPT; ++i) one[i] = 1.0f;
      storeT(a.hT16u, 0, one);
    }
  };
  float p_hv[NPT], p_gi[NPT], p_gf[NPT], p_gg[NPT], p_go[NPT], p_cs[NPT];
  int p_t = -1;

  PROF_DECL;
  PROF2_DECL;
  for (int s = 0; s < T; ++s) {
    PROF(7);
    const int t = dir ? (T - 1 - s) : s;
    float zx[NPT][4];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const float* zr = a.zx + (((size_t)t * N + n0 + warp * NPT + i) * 2 + dir) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) zx[i][g] = __ldg(zr + g * H + u);
    }
    // zoneout keep coefficients of this step: the train-phase mask (one per time step and unit, shared by the batch) or
    // the inference blend 1 - level (core/layers_utils.py:34-42); zmask = [h | c][2 dirs][T][H]
    float kh = 1.0f, kc = 1.0f;
    if (zone) {
      kh = a.zmask ? __ldg(a.zmask + ((size_t)(0 * 2 + dir) * T + t) * H + u) : 1.0f - a.zoneout;
      kc = a.zmask ? __ldg(a.zmask + ((size_t)(1 * 2 + dir) * T + t) * H + u) : 1.0f - a.zoneout;
    }
    float z[NPT][4];
    if (s > 0) {
      // ---- pull h_{t-1} of this group: poll the LL words (data + tag in one 8-byte access) ----------
      // side outputs of the previous step first: the extra ~350 cycles let the peers' LL words land in L2, so
      // the first probe usually hits (a probe that races the store costs a second full L2 round trip)
      // warp w polls (and stages) the K range [H/4 * w, H/4 * (w + 1)) of all NB samples: V4W 16-byte accesses
      // (2 LL words = 4 K columns each) per sample, consecutive lanes on consecutive addresses
      const uint4* src = reinterpret_cast<const uint4*>(xb + (size_t)((s - 1) & 1) * WORDS);
      const uint32_t tag = (uint32_t)s;
      constexpr int V4W = H / 16;                         // 16-byte accesses per (sample, warp K range)
      constexpr int QPT = NB * V4W / 32;                  // per lane
      static_assert((NB * V4W) % 32 == 0, "per-warp poll set must fill whole warp accesses");
      int vidx[QPT];
#pragma unroll
      for (int q = 0; q < QPT; ++q) {
        const int f = q * 32 + lane;
        vidx[q] = (f / V4W) * (H / 4) + warp * V4W + (f % V4W);
      }
      if (p_t >= 0) side_stores(p_t, p_hv, p_gi, p_gf, p_gg, p_go, p_cs);
      if (delay1 > 0) __nanosleep(delay1);
      uint4 w[QPT];
#pragma unroll
      for (int q = 0; q < QPT; ++q) w[q] = ld_volatile_v4(src + vidx[q]);
      bool ok;
      long long t0 = 0;
      do {
        ok = true;
#pragma unroll
        for (int q = 0; q < QPT; ++q)
          if (w[q].y != tag || w[q].w != tag) {
            w[q] = ld_volatile_v4(src + vidx[q]);
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
      PROF(0);
#pragma unroll
      for (int q = 0; q < QPT; ++q) {
        const int f = q * 32 + lane;
        const int n = f / V4W, k = 4 * (warp * V4W + (f % V4W));   // both fp16 pairs sit in one 8-byte smem slot
        *reinterpret_cast<uint2*>(sB + (k >> 6) * B_CHUNK + tc::sw128_offset(n, k & 63)) = make_uint2(w[q].x, w[q].z);
      }
      // Each warp staged exactly the K range [KW * warp, KW * (warp + 1)) of the B operand (its lanes polled those LL
      // words), so it issues the MMAs of that range itself, into its own TMEM accumulator, as soon as ITS words
      // have landed: no block-wide barrier between "data arrived" and "MMA issued", and the MMAs of the early
      // warps run under the polling of the late ones.  The four commits complete mma_bar (count 4).
      tc::fence_proxy_async_smem();
      __syncwarp();
      PROF(1);
      if (tc::elect_one_sync()) {
        tc::tcgen05_fence_after();
        constexpr int KBW = H / 16 / 4;                      // MMAs (K = 16 each) per warp
#pragma unroll
        for (int j = 0; j < KBW; ++j) {
          const int kb = warp * KBW + j;
          const uint64_t bd = tc::umma_desc_sw128(sB_addr + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
          tc::umma_ts(tmem + D_COL + warp * NM, tmem + A_COL + kb * 8, bd, idesc, j > 0);
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
        uint32_t r0[NB], r1[NB], r2[NB], r3[NB];
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16) + D_COL;
        tc::tmem_ldn(tq, r0);
        tc::tmem_ldn(tq + NM, r1);
        tc::tmem_ldn(tq + 2 * NM, r2);
        tc::tmem_ldn(tq + 3 * NM, r3);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < NB; ++n)
          sZ[(warp * NB + n) * 32 + lane] = (__uint_as_float(r0[n]) + __uint_as_float(r1[n])) +
                                            (__uint_as_float(r2[n]) + __uint_as_float(r3[n]));
      }
      tc::tcgen05_fence_before();
      __syncthreads();
      PROF(4);
      if (s_dead) break;
#pragma unroll
      for (int i = 0; i < NPT; ++i)
#pragma unroll
        for (int g = 0; g < 4; ++g) z[i][g] = sZ[(g * NB + warp * NPT + i) * 32 + lane];
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
      float pre[4];
#pragma unroll
      for (int g = 0; g < 4; ++g)
        pre[g] = VAR ? fmaf(mia[g] * zx[i][g], z[i][g], fmaf(mib1[g], z[i][g], fmaf(mib2[g], zx[i][g], bias[g])))
                     : z[i][g] + zx[i][g] + bias[g];
      if (VAR && a.uh && a.training) {                     // raw recurrent product: the backward pass of MI needs it
        float* up = a.uh + (((size_t)t * N + n0 + warp * NPT + i) * 2 + dir) * 4 * H;
#pragma unroll
        for (int g = 0; g < 4; ++g) up[g * H + u] = z[i][g];
      }
      gi[i] = asr::hard_sigmoid(pre[0]);
      gf[i] = asr::hard_sigmoid(pre[1]);
      gg[i] = asr::tanh_fast(pre[2]);
      go[i] = asr::hard_sigmoid(pre[3]);
      const float c_new = gf[i] * c_state[i] + gi[i] * gg[i];
      c_state[i] = VAR ? fmaf(kc, c_new - c_state[i], c_state[i]) : c_new;
      const float h_new = go[i] * asr::tanh_fast(c_state[i]);
      hv[i] = VAR ? fmaf(kh, h_new - h_prev[i], h_prev[i]) : h_new;
      h_prev[i] = hv[i];
      // publish h * B_U: even lanes pack (unit, unit+1) into one LL word {half2, tag = s+1}
      const float hm = hv[i] * mu[i];
      const float other = __shfl_down_sync(0xffffffffu, hm, 1);
      if (!(lane & 1)) {
        const __half2 pk = __floats2half2_rn(hm, other);
        uint2 wv;
        wv.x = *reinterpret_cast<const uint32_t*>(&pk);
        wv.y = (uint32_t)(s + 1);
        st_volatile_v2(xo + (size_t)(warp * NPT + i) * (H / 2) + (u >> 1), wv);
      }
    }
    PROF(5);
    // stash this step's side outputs; they are written after the next step's poll loads have been issued
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      p_hv[i] = hv[i]; p_gi[i] = gi[i]; p_gf[i] = gf[i]; p_gg[i] = gg[i]; p_go[i] = go[i]; p_cs[i] = c_state[i];
    }
    p_t = t;
    PROF(6);
  }
  if (p_t >= 0 && !s_dead) side_stores(p_t, p_hv, p_gi, p_gf, p_gg, p_go, p_cs);   // last step
  PROF_DUMP(0);
  PROF2_DUMP(16);
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// backward through time: ONE exchange per step.  H a multiple of 64 up to 896: H / 32 CTAs per (direction,
//   batch group); the figures below are for H = 512 (16 CTAs, 4 blocks).  Wider layers (5 to 7 blocks: the U slice
//   fills 448 of the 512 TMEM columns, which leaves room for four accumulators) run the product in two rounds of
//   at most four blocks; the rows of a last half block (H = 832) beyond H are zero and are never sent.
//   CTA j owns the 32 hidden units [32j, 32j + 32) for the element-wise BPTT step, i.e. the 128 gate columns
//   {g*H + 32j + i}, and keeps the [H units x 128 own gate columns] slice of U (bf16) in TMEM as H / 128
//   M = 128 blocks (warps >= H / 128 issue no MMA; every warp still sends its lane quarter of every block).  Its own dz_t is the B operand — written to shared memory locally, no gather — and
//       P_j[u][n] = sum_{k in own columns} U[u][k] * dz_t[n][k]            (4 blocks x 8 TS-mode tcgen05.mma)
//   is its partial contribution to dh_rec of ALL 512 units.  Warp w of CTA j holds, for block b, the rows of the
//   units owned by CTA 4b + w and sends them there (reduce-scatter through the LL ring); every CTA sums the 16
//   partials that arrive for its own units.  A row-partitioned kernel (round 1's first BPTT) needs two dependent
//   exchanges per step (gather dz of a column block, then reduce partials along a row: ~2.0 k + ~1.1 k cycles of
//   4.5 k); this one needs one, with the same wire bytes per CTA as the forward all-gather (partials travel as bf16 pairs — the
//   operands of the product are bf16 already, see DESIGN.md for the error budget).
// ------------------------------------------------------------------------------------------------
template <int H, int NB, bool VAR>
__global__ void __launch_bounds__(THREADS, 1)
bwd3_kernel(asr_lstm_bwd_args a, int* __restrict__ flags, uint2* __restrict__ xbuf, int grp0) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int K4 = 4 * H, NCTA = H / UPC;
  constexpr int NBLK = (H + 127) / 128;                  // M = 128 blocks of the U slice
  constexpr int NR = (NBLK + 3) / 4;                     // rounds of <= 4 blocks (one issuing warp, one accumulator each)
  static_assert(H % 64 == 0 && NBLK >= 1 && NBLK <= 7, "U slice: NBLK x 64 TMEM columns beside four accumulators");
  constexpr int B_CHUNK = NM * 128;                      // one 64-wide K chunk of the B operand
  constexpr int NPT = NB / 4;                            // samples per thread
  constexpr int PPT = NPT / 2;                           // sample pairs per thread
  constexpr int NP = NB / 2;                             // sample pairs per group
  constexpr int SLOT = NP * NCTA * 32;                   // LL words per (parity, destination): [pair][source][unit]
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z, G = gridDim.z;
  const int u0 = cta * UPC, n0 = (grp0 + grp) * NB;

  uint8_t* sB = smem;                                    // 2 chunks of [NM rows x 128 B]: K = 128 own gate columns
  uint64_t* mma_bar = reinterpret_cast<uint64_t*>(sB + 2 * B_CHUNK);      // one per round
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 2);
  __shared__ int s_dead;

  if (tid == 0) {
    tc::mbar_init(mma_bar, NBLK < 4 ? NBLK : 4);           // one tcgen05.commit per issuing warp
    if constexpr (NR > 1) tc::mbar_init(mma_bar + 1, (uint32_t)(NBLK - 4));
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < 2 * B_CHUNK / 16; i += THREADS) reinterpret_cast<uint4*>(sB)[i] = make_uint4(0u, 0u, 0u, 0u);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  if (*tmem_slot != 0u) { if (tid == 0) atomicExch(flags + STATUS_IDX, 2); }
  constexpr uint32_t tmem = 0u;

  // one-time: U slice -> TMEM.  block b, lane m <-> unit 128b + m; K index k = g*32 + i <-> gate column g*H + u0 + i
  {
#pragma unroll 1
    for (int b = 0; b < NBLK; ++b) {
      const bool row_ok = 128 * b + tid < H;             // a last half block: the rows beyond H are zero
      const __nv_bfloat16* Ub = reinterpret_cast<const __nv_bfloat16*>(a.U16) + (size_t)dir * H * K4 +
                                (size_t)(row_ok ? 128 * b + tid : 0) * K4 + u0;
#pragma unroll 1
      for (int hs = 0; hs < 2; ++hs) {                   // two gates (64 K elements = 32 columns) per tcgen05.st
        uint32_t rr[32];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const uint4* src = reinterpret_cast<const uint4*>(Ub + (2 * hs + gg) * H);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 v = row_ok ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
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

  const uint32_t idesc = tc::umma_idesc_f16(128, NM, 1);
  const uint32_t sB_addr = tc::smem_u32(sB);
  const int u = u0 + lane;
  float mia[4], mib1[4], mib2[4], gal[4] = {0, 0, 0, 0}, gb1[4] = {0, 0, 0, 0}, gb2[4] = {0, 0, 0, 0};
  const bool mi = VAR && a.mi != nullptr, zone = VAR && a.zoneout > 0.0f;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const size_t o = (size_t)dir * K4 + g * H + u;
    mia[g] = mi ? a.mi[o] : 0.0f;
    mib1[g] = mi ? a.mi[(size_t)2 * K4 + o] : 1.0f;
    mib2[g] = mi ? a.mi[(size_t)4 * K4 + o] : 1.0f;
  }
  float dh_zone[NPT];                                    // zoneout: (1 - k_h) * dh passes straight to h_{t-1}
  float dc_carry[NPT], mu[NPT], md0[NPT], md1[NPT], db[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    dc_carry[i] = 0.0f;
    dh_zone[i] = 0.0f;
    mu[i] = a.mask_u ? a.mask_u[((size_t)dir * N + n0 + warp * NPT + i) * H + u] : 1.0f;
    md0[i] = a.mask_dh ? a.mask_dh[((size_t)0 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
    md1[i] = a.mask_dh ? a.mask_dh[((size_t)1 * N + n0 + warp * NPT + i) * 2 * H + dir * H + u] : 1.0f;
  }

  int* status = flags + STATUS_IDX;
  uint2* xb = xbuf + (size_t)(dir * G + grp) * 2 * NCTA * SLOT;     // [(dir,grp)][parity][dst][pair][src][unit]
  __nv_bfloat16* dz16 = reinterpret_cast<__nv_bfloat16*>(a.dz16);
  const size_t R = (size_t)T * N;

  auto side_stores = [&](int t, const float (&dz)[NPT][4], const float (&du)[NPT][4]) {
    if (mi && a.duhT16) {                                  // dL/d(uh), transposed: the dU operand when MI splits the two paths
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __nv_bfloat16* dstT = reinterpret_cast<__nv_bfloat16*>(a.duhT16) + (size_t)(dir * K4 + g * H + u) * ((size_t)T * N) + (size_t)t * N + n0 + warp * NPT;
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(du[0][g], du[1][g]);
        if constexpr (NPT == 4) {
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(du[2][g], du[3][g]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&p0);
          pk.y = *reinterpret_cast<const uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(dstT) = pk;
        } else {
          *reinterpret_cast<__nv_bfloat162*>(dstT) = p0;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        dz16[(row * 2 + dir) * K4 + g * H + u] = __float2bfloat16_rn(dz[i][g]);
        if (a.dz32) a.dz32[(row * 2 + dir) * K4 + g * H + u] = dz[i][g];
      }
    }
    if (a.dzT16) {
      static_assert(NPT == 4 || NPT == 2, "packed transposed store: 2 or 4 samples per thread");
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        __nv_bfloat16* dstT = reinterpret_cast<__nv_bfloat16*>(a.dzT16) + (size_t)(dir * K4 + g * H + u) * R + (size_t)t * N + n0 + warp * NPT;
        const __nv_bfloat162 p0 = __floats2bfloat162_rn(dz[0][g], dz[1][g]);
        if constexpr (NPT == 4) {
          const __nv_bfloat162 p1 = __floats2bfloat162_rn(dz[2][g], dz[3][g]);
          uint2 pk;
          pk.x = *reinterpret_cast<const uint32_t*>(&p0);
          pk.y = *reinterpret_cast<const uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(dstT) = pk;
        } else {
          *reinterpret_cast<__nv_bfloat162*>(dstT) = p0;
        }
      }
    }
  };
  float p_dz[NPT][4], p_du[NPT][4];      // p_du only lives in the VAR instantiation (never read otherwise)
  int p_t = -1;

  PROF_DECL;
  for (int s = 0; s < T; ++s) {
    PROF(7);
    const int t = dir ? s : (T - 1 - s);
    const int t_fprev = dir ? (t + 1) : (t - 1);
    const bool has_fprev = dir ? (t + 1 < T) : (t > 0);
    float dho[NPT], dho2[NPT], gi[NPT], gf[NPT], gg[NPT], go[NPT], cc[NPT], cp[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      dho[i] = __ldg(a.dh + row * 2 * H + dir * H + u);
      dho2[i] = a.dh2 ? __ldg(a.dh2 + row * 2 * H + dir * H + u) : 0.0f;
      const float* gp = a.gates + (row * 2 + dir) * 4 * H;
      gi[i] = __ldg(gp + u); gf[i] = __ldg(gp + H + u); gg[i] = __ldg(gp + 2 * H + u); go[i] = __ldg(gp + 3 * H + u);
      cc[i] = __ldg(a.cell + (row * 2 + dir) * H + u);
      cp[i] = has_fprev ? __ldg(a.cell + ((((size_t)t_fprev * N + n0 + warp * NPT + i) * 2 + dir) * H + u)) : 0.0f;
    }
    float wxv[VAR ? NPT : 1][4], uhv[VAR ? NPT : 1][4], kh = 1.0f, kc = 1.0f;
    if (mi) {
#pragma unroll
      for (int i = 0; i < NPT; ++i) {
        const size_t o = (((size_t)t * N + n0 + warp * NPT + i) * 2 + dir) * K4 + u;
#pragma unroll
        for (int g = 0; g < 4; ++g) { wxv[VAR ? i : 0][g] = __ldg(a.zx + o + g * H); uhv[VAR ? i : 0][g] = __ldg(a.uh + o + g * H); }
      }
    }
    if (zone) {
      kh = a.zmask ? __ldg(a.zmask + ((size_t)(0 * 2 + dir) * T + t) * H + u) : 1.0f - a.zoneout;
      kc = a.zmask ? __ldg(a.zmask + ((size_t)(1 * 2 + dir) * T + t) * H + u) : 1.0f - a.zoneout;
    }
    float dh_rec[NPT];
#pragma unroll
    for (int i = 0; i < NPT; ++i) dh_rec[i] = 0.0f;
    if (s > 0) {
      // ---- receive: NCTA partials (bf16 pairs of two samples) for each of my (unit, sample pair) --------------
      const uint32_t tag = (uint32_t)s;
      const uint2* src = xb + ((size_t)((s - 1) & 1) * NCTA + cta) * SLOT + (size_t)(warp * PPT) * NCTA * 32 + lane;
      if (p_t >= 0) side_stores(p_t, p_dz, p_du);   // first: gives the peers' LL words time to land in L2
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
      PROF(0);
    }
    // ---- element-wise BPTT -> dz; my dz is the B operand of my own product ------------------------------------
    float dz[NPT][4], du[VAR ? NPT : 1][4];
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      float dh = fmaf(dho2[i], md1[i], fmaf(dho[i], md0[i], dh_rec[i]));
      if (VAR) {                                           // h_t = h_{t-1} + k_h (h_new - h_{t-1})
        dh += dh_zone[i];
        dh_zone[i] = (1.0f - kh) * dh;
        dh *= kh;
      }
      const float tch = asr::tanh_fast(cc[i]);
      const float d_o = dh * tch * asr::hard_sigmoid_grad(go[i]);
      float dc = dc_carry[i] + dh * go[i] * (1.0f - tch * tch);
      float dc_pass = 0.0f;
      if (VAR) {                                           // c_t = c_{t-1} + k_c (c_new - c_{t-1})
        dc_pass = (1.0f - kc) * dc;
        dc *= kc;
      }
      dz[i][0] = dc * gg[i] * asr::hard_sigmoid_grad(gi[i]);
      dz[i][1] = dc * cp[i] * asr::hard_sigmoid_grad(gf[i]);
      dz[i][2] = dc * gi[i] * (1.0f - gg[i] * gg[i]);
      dz[i][3] = d_o;
      dc_carry[i] = dc * gf[i] + dc_pass;
      const int n = warp * NPT + i;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v = dz[i][g];                                // dL/dz; with MI the recurrent path sees dz * (alpha*Wx + beta1)
        if (mi) {
          const float wx = wxv[VAR ? i : 0][g], uh = uhv[VAR ? i : 0][g];
          gal[g] = fmaf(dz[i][g] * wx, uh, gal[g]);
          gb1[g] = fmaf(dz[i][g], uh, gb1[g]);
          gb2[g] = fmaf(dz[i][g], wx, gb2[g]);
          db[g] += dz[i][g];
          v = dz[i][g] * fmaf(mia[g], wx, mib1[g]);
          du[VAR ? i : 0][g] = v;
          dz[i][g] *= fmaf(mia[g], uh, mib2[g]);           // dL/d(zx): what dW and dX consume
        }
        const int k = g * 32 + lane;                       // K index inside my 128 gate columns
        *reinterpret_cast<__nv_bfloat16*>(sB + (k >> 6) * B_CHUNK + tc::sw128_offset(n, k & 63)) = __float2bfloat16_rn(v);
      }
    }
    PROF(1);
    if (s + 1 < T) {
      tc::fence_proxy_async_smem();
      __syncthreads();                                     // the whole B operand (all samples) is staged
      if (s_dead) break;
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        constexpr int B0[2] = {0, 4};
        const int b0 = B0[r], cnt = (NBLK - b0 < 4) ? NBLK - b0 : 4;       // blocks b0 .. b0 + cnt - 1 in this round
        if (r > 0) {                                         // every warp has read the accumulators of the previous round
          tc::tcgen05_fence_before();
          __syncthreads();
        }
        if (warp < cnt && tc::elect_one_sync()) {            // warp w issues M block b0 + w (units 128 (b0 + w) ..)
          tc::tcgen05_fence_after();
#pragma unroll
          for (int kb = 0; kb < 8; ++kb) {
            const uint64_t bd = tc::umma_desc_sw128(sB_addr + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
            tc::umma_ts(tmem + D_COL + warp * NM, tmem + A_COL + (b0 + warp) * 64 + kb * 8, bd, idesc, kb > 0);
          }
          tc::umma_commit(mma_bar + r);
        }
        if (!tc::mbar_wait(mma_bar + r, (uint32_t)(s & 1), WATCHDOG_CYCLES)) {
          atomicExch(status, 1);
          s_dead = 1;
        }
        tc::tcgen05_fence_after();
        if (r == 0) PROF(2);
        // ---- send: my warp's rows of block b belong to CTA 4b + warp --------------------------------------------
        uint32_t rb[4][NB];
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16) + D_COL;
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (b < cnt) tc::tmem_ldn(tq + b * NM, rb[b]);
        tc::tmem_ld_wait();
        const uint32_t tg = (uint32_t)(s + 1);
        uint2* out = xb + (size_t)(s & 1) * NCTA * SLOT + (size_t)cta * 32 + lane;      // + dst*SLOT + pair*NCTA*32
#pragma unroll
        for (int np = 0; np < NP; ++np) {
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            if (b < cnt && (H % 128 == 0 || (b0 + b) * 4 + warp < NCTA)) {     // a last half block has two destinations
              const __nv_bfloat162 q = __floats2bfloat162_rn(__uint_as_float(rb[b][2 * np]), __uint_as_float(rb[b][2 * np + 1]));
              st_volatile_v2(out + (size_t)((b0 + b) * 4 + warp) * SLOT + (size_t)np * NCTA * 32, make_uint2(*reinterpret_cast<const uint32_t*>(&q), tg));
            }
          }
        }
      }
      tc::tcgen05_fence_before();
      PROF(3);
    }
    // stash the side outputs (dz for the dW/dU/dX GEMMs); written while the next poll is in flight
#pragma unroll
    for (int i = 0; i < NPT; ++i)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (!mi) db[g] += dz[i][g];
        p_dz[i][g] = dz[i][g];
        if (VAR) p_du[i][g] = mi ? du[i][g] : dz[i][g];
      }
    p_t = t;
    PROF(6);
  }
  if (p_t >= 0 && !s_dead) side_stores(p_t, p_dz, p_du);
  if (mi) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const size_t o = (size_t)dir * K4 + g * H + u;
      atomicAdd(a.dmi + o, gal[g]);
      atomicAdd(a.dmi + (size_t)2 * K4 + o, gb1[g]);
      atomicAdd(a.dmi + (size_t)4 * K4 + o, gb2[g]);
    }
  }
  PROF_DUMP(8);
#pragma unroll
  for (int g = 0; g < 4; ++g) atomicAdd(a.dbias + (size_t)dir * K4 + g * H + u, db[g]);
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------------------------
// batch groups one cooperative launch holds: every CTA owns its SM
static int max_groups(int H) { return 148 / (2 * (H / UPC)); }
// samples per CTA group: 8 when the batch then still fits one cooperative wave (more SMs, half the exchange per CTA),
// else 16; a batch of more than max_groups() groups runs as several launches over consecutive groups (grp0)
static int group_size(int N, int H, int opts) {
  if (opts & ASR_LSTM_GROUP16) return (N % 16 == 0) ? 16 : 0;
  if (N % 8 == 0 && N / 8 <= max_groups(H)) return 8;
  if (N % 16 == 0) return 16;
  return (N % 8 == 0) ? 8 : 0;
}
// widths with an instantiation: the U^T slice of the forward kernel needs H / 2 TMEM columns beside 64 of accumulators
// (H <= 896), the U slice of the BPTT kernel ceil(H / 128) x 64 columns (<= 448).  Other widths are zero-padded up to
// the next of these by the host engine (engine.py: padded units stay exactly zero in both passes).
static bool width_ok(int H) {
  return H == 128 || H == 256 || H == 384 || H == 512 || H == 640 || H == 768 || H == 832 || H == 896;
}
static bool shape_ok(int T, int N, int H, int opts) { return T >= 1 && width_ok(H) && N >= 8 && group_size(N, H, opts) != 0; }
bool shape_supported(int T, int N, int H, int opts) { return shape_ok(T, N, H, opts); }
bool supports_fwd(const asr_lstm_fwd_args* a) { return a->zx && a->U16 && (a->h16 || a->hm16) && (!a->training || (a->gates && a->cell)) && shape_ok(a->T, a->N, a->H, a->opts); }
bool supports_bwd(const asr_lstm_bwd_args* a) { return a->gates && a->cell && a->U16 && a->dz16 && shape_ok(a->T, a->N, a->H, a->opts); }
static size_t fwd_ring_bytes(int H, int NB, int G) { return (size_t)2 * G * 2 * NB * (H / 2) * sizeof(uint2); }
// [(dir, grp)][parity][dst][pair][src][unit]
static size_t bwd3_ring_bytes(int H, int NB, int G) {
  const size_t nc = H / UPC;
  return (size_t)2 * G * 2 * nc * (NB / 2) * nc * 32 * sizeof(uint2);
}
size_t scratch_bytes(int) {                              // the largest ring any launch clears
  size_t m = 0;
  for (int H = 128; H <= 896; H += 64) {
    if (!width_ok(H)) continue;
    for (int NB = 8; NB <= 16; NB += 8) {
      const size_t f = fwd_ring_bytes(H, NB, max_groups(H)), b = bwd3_ring_bytes(H, NB, max_groups(H));
      m = f > m ? f : m;
      m = b > m ? b : m;
    }
  }
  return HEADER_BYTES + m;
}

// The recurrence CTAs own their SM: they hold all 512 TMEM columns, and any co-resident CTA would steal issue
// slots from a latency-bound chain.  Requesting (almost) the whole shared memory keeps every other kernel's CTAs
// (e.g. the gradient GEMMs of the previous layer running on a low-priority side stream) on the SMs this grid
// does not use.  ASR_LSTM_SHARED_SM in the argument record's opts turns it off.
static size_t exclusive_smem(size_t need, int opts) {
  const size_t want = 200 * 1024;
  return (!(opts & ASR_LSTM_SHARED_SM) && need < want) ? want : need;
}

template <int H, int NB, bool VAR>
static int32_t launch_fwd(const asr_lstm_fwd_args* a, cudaStream_t st) {
  constexpr int KC = H / 64;
  const size_t smem = exclusive_smem(1024 + (size_t)KC * NM * 128 + 4 * NB * 32 * 4 + 64, a->opts);
  const int Gall = a->N / NB, gm = max_groups(H);
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<H, NB, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  asr_lstm_fwd_args args = *a;
  int* flags = a->flags;
  uint2* xbuf = reinterpret_cast<uint2*>(reinterpret_cast<char*>(a->flags) + HEADER_BYTES);
  int delay1 = 0;
  for (int grp0 = 0; grp0 < Gall; grp0 += gm) {          // one launch per max_groups() batch groups (C2: a single launch)
    const int G = Gall - grp0 < gm ? Gall - grp0 : gm;
    ASR_CUDA(cudaMemsetAsync(a->flags, 0, HEADER_BYTES + fwd_ring_bytes(H, NB, G), st));
    void* kargs[] = {&args, &flags, &xbuf, &delay1, &grp0};
    ASR_CUDA(cudaLaunchCooperativeKernel((void*)fwd_kernel<H, NB, VAR>, dim3(H / UPC, 2, G), dim3(THREADS), kargs, smem, st));
    asr::count_launch();
  }
  return ASR_OK;
}

template <int H, int NB, bool VAR>
static int32_t launch_bwd3(const asr_lstm_bwd_args* a, cudaStream_t st) {
  constexpr int NCTA = H / UPC;
  const int Gall = a->N / NB, gm = max_groups(H);
  const size_t smem = exclusive_smem(1024 + (size_t)2 * NM * 128 + 64, a->opts);
  ASR_CUDA(cudaFuncSetAttribute(bwd3_kernel<H, NB, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->dbias, 0, (size_t)2 * 4 * a->H * sizeof(float), st));      // accumulated over all groups
  if (VAR && a->mi) ASR_CUDA(cudaMemsetAsync(a->dmi, 0, (size_t)3 * 2 * 4 * a->H * sizeof(float), st));
  asr_lstm_bwd_args args = *a;
  int* flags = a->flags;
  uint2* xbuf = reinterpret_cast<uint2*>(reinterpret_cast<char*>(a->flags) + HEADER_BYTES);
  for (int grp0 = 0; grp0 < Gall; grp0 += gm) {
    const int G = Gall - grp0 < gm ? Gall - grp0 : gm;
    ASR_CUDA(cudaMemsetAsync(a->flags, 0, HEADER_BYTES + bwd3_ring_bytes(H, NB, G), st));
    void* kargs[] = {&args, &flags, &xbuf, &grp0};
    ASR_CUDA(cudaLaunchCooperativeKernel((void*)bwd3_kernel<H, NB, VAR>, dim3(NCTA, 2, G), dim3(THREADS), kargs, smem, st));
    asr::count_launch();
  }
  return ASR_OK;
}

int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st) {
  const bool g8 = group_size(a->N, a->H, a->opts) == 8;
  const bool var = a->mi != nullptr || a->zoneout > 0.0f;
  if (var) ASR_CHECK_ARG(!a->mi || (a->zx && a->uh && a->dmi && a->duhT16), "lstmtc2 backward: MI needs zx, uh, dmi and duhT16");
#define ASR_BWD3_CASE(HH)                                                                             \
  case HH:                                                                                            \
    if (var) return g8 ? launch_bwd3<HH, 8, true>(a, st) : launch_bwd3<HH, 16, true>(a, st);          \
    return g8 ? launch_bwd3<HH, 8, false>(a, st) : launch_bwd3<HH, 16, false>(a, st);
  switch (a->H) {
    ASR_BWD3_CASE(128)
    ASR_BWD3_CASE(256)
    ASR_BWD3_CASE(384)
    ASR_BWD3_CASE(512)
    ASR_BWD3_CASE(640)
    ASR_BWD3_CASE(768)
    ASR_BWD3_CASE(832)
    ASR_BWD3_CASE(896)
  }
#undef ASR_BWD3_CASE
  asr::set_error("lstmtc2: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  const bool g8 = group_size(a->N, a->H, a->opts) == 8;
  const bool var = a->mi != nullptr || a->zoneout > 0.0f;
  if (var) ASR_CHECK_ARG(!a->mi || !a->training || a->uh, "lstmtc2 forward: training with MI needs the uh buffer");
#define ASR_FWD_CASE(HH)                                                                              \
  case HH:                                                                                            \
    if (var) return g8 ? launch_fwd<HH, 8, true>(a, st) : launch_fwd<HH, 16, true>(a, st);            \
    return g8 ? launch_fwd<HH, 8, false>(a, st) : launch_fwd<HH, 16, false>(a, st);
  switch (a->H) {
    ASR_FWD_CASE(128)
    ASR_FWD_CASE(256)
    ASR_FWD_CASE(384)
    ASR_FWD_CASE(512)
    ASR_FWD_CASE(640)
    ASR_FWD_CASE(768)
    ASR_FWD_CASE(832)
    ASR_FWD_CASE(896)
  }
#undef ASR_FWD_CASE
  asr::set_error("lstmtc2: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

}  // namespace lstmtc2
