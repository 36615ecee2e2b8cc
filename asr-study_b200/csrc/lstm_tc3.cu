// K3 (tensor-core engine, v3) — persistent BiLSTM forward recurrence, thread-block-cluster edition.
//
// Same arithmetic and TMEM layout as lstm_tc2.cu (U^T slice resident in tensor memory, TS-mode
// tcgen05.mma into 4 accumulators, gates swapped through smem, c_t in registers).  What changes is the
// h_t exchange: the H/32 CTAs of one (direction, batch group) are ONE thread-block cluster, and every
// CTA writes its 32 x 16 fp16 slice of h_t straight into the shared-memory B operand of all its peers
// (st.shared::cluster, already in the 128B-swizzled UMMA layout) and then arrives on each peer's
// mbarrier with release.cluster semantics.  The consumer's elected thread waits on its LOCAL mbarrier
// (acquire.cluster), fences the async proxy and issues the MMAs: no L2 round trip, no polling traffic,
// no block-wide barrier between "data arrived" and "MMA issued".
#include "common.cuh"
#include "tc.cuh"

namespace lstmtc3 {

constexpr int UPC = 32;
constexpr int NG = 16;
constexpr int THREADS = 128;
constexpr int STATUS_IDX = 64;
constexpr int NACC = 4;
constexpr uint32_t D_COL = 0, A_COL = 64;
constexpr long long WATCHDOG_CYCLES = 2000000000LL;

__device__ __forceinline__ uint32_t mapa(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(tc::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_wait_cluster(uint64_t* bar, uint32_t parity, long long limit) {
  if (mbar_try_wait_cluster(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity))
    if (clock64() - t0 > limit) return false;
  return true;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int H>
__global__ void __launch_bounds__(THREADS, 1)
fwd_kernel(asr_lstm_fwd_args a, int* __restrict__ flags) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int KC = H / 64;
  constexpr int NCTA = H / UPC;
  constexpr int B_CHUNK = NG * 128;
  constexpr int SB_BYTES = KC * B_CHUNK;
  constexpr int NPT = NG / 4;
  const int T = a.T, N = a.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cta = blockIdx.x, dir = blockIdx.y, grp = blockIdx.z;
  const int u0 = cta * UPC, n0 = grp * NG;

  uint8_t* sB = smem;                                                  // [2 parity][KC][NG x 128 B]
  float* sZ = reinterpret_cast<float*>(sB + 2 * SB_BYTES);              // [4][NG][32]
  __half* sP = reinterpret_cast<__half*>(sZ + 4 * NG * 32);             // [NG][32] staging of my h slice
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sP + NG * 32);       // [2]
  uint64_t* mma_bar = full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_bar + 1);
  __shared__ int s_dead;

  if (tid == 0) {
    tc::mbar_init(full_bar + 0, NCTA * 4);                               // one arrival per producer warp
    tc::mbar_init(full_bar + 1, NCTA * 4);
    tc::mbar_init(mma_bar, 1);
    tc::fence_mbar_init();
    s_dead = 0;
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, 512);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  int* status = flags + STATUS_IDX;
  if (*tmem_slot != 0u) { if (tid == 0) atomicExch(status, 2); }
  constexpr uint32_t tmem = 0u;                  // whole TMEM allocated -> base column 0 (compile-time addresses)

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
  cluster_sync_all();                            // every peer's barriers are initialised before any remote access

  const uint32_t idesc = tc::umma_idesc_f16(128, NG, 0);
  const uint32_t sB_addr = tc::smem_u32(sB);
  const uint32_t bar_addr = tc::smem_u32(full_bar);
  const int u = u0 + lane;
  float bias[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
  float c_state[NPT], mu[NPT];
#pragma unroll
  for (int i = 0; i < NPT; ++i) {
    c_state[i] = 0.0f;
    mu[i] = a.mask_u ? a.mask_u[((size_t)dir * N + n0 + warp * NPT + i) * H + u] : 1.0f;
  }
  __half* h16 = reinterpret_cast<__half*>(a.h16);
  const size_t R = (size_t)T * N;

  // my send slot: chunk (row i of my warp's 4 samples, 16-byte column c4) -> 8 of the NCTA destinations
  const int s_i = (lane & 15) >> 2, s_c4 = lane & 3, s_half = lane >> 4;
  const int s_n = warp * NPT + s_i;                                       // sample row inside the group
  const uint32_t s_off = (uint32_t)((cta >> 1) * B_CHUNK + (s_n >> 3) * 1024 + (s_n & 7) * 128 +
                                    (((4 * (cta & 1) + s_c4) ^ (s_n & 7)) << 4));

  auto side_stores = [&](int t, const float (&hv)[NPT], const float (&gi)[NPT], const float (&gf)[NPT],
                         const float (&gg)[NPT], const float (&go)[NPT], const float (&cs)[NPT]) {
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      const size_t row = (size_t)t * N + n0 + warp * NPT + i;
      h16[row * 2 * H + dir * H + u] = __float2half_rn(hv[i]);
      if (a.h32) a.h32[row * 2 * H + dir * H + u] = hv[i];
      if (a.training) {
        float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gp[u] = gi[i]; gp[H + u] = gf[i]; gp[2 * H + u] = gg[i]; gp[3 * H + u] = go[i];
        a.cell[(row * 2 + dir) * H + u] = cs[i];
      }
    }
    if (a.training && a.hT16) {
      static_assert(NPT == 4, "packed transposed store assumes 4 samples per thread");
      const __nv_bfloat162 p0 = __floats2bfloat162_rn(hv[0] * mu[0], hv[1] * mu[1]),
                           p1 = __floats2bfloat162_rn(hv[2] * mu[2], hv[3] * mu[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&p0);
      pk.y = *reinterpret_cast<const uint32_t*>(&p1);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(a.hT16) + (size_t)(dir * H + u) * R +
                                (size_t)t * N + n0 + warp * NPT) = pk;
    }
  };

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
      const int par = (s - 1) & 1;
      if (warp == 0 && tc::elect_one_sync()) {
        // h_{t-1} has been pushed into sB[par] by all peers once the 4*NCTA warp arrivals are in
        if (!mbar_wait_cluster(full_bar + par, (uint32_t)(((s - 1) >> 1) & 1), WATCHDOG_CYCLES)) {
          atomicExch(status, 1);
          s_dead = 1;
        }
        tc::fence_proxy_async_smem();
        tc::tcgen05_fence_after();
#pragma unroll
        for (int kb = 0; kb < H / 16; ++kb) {
          const uint64_t bd = tc::umma_desc_sw128(sB_addr + par * SB_BYTES + (kb >> 2) * B_CHUNK) + 2 * (kb & 3);
          tc::umma_ts(tmem + D_COL + (kb % NACC) * NG, tmem + A_COL + kb * 8, bd, idesc, kb >= NACC);
        }
        tc::umma_commit(mma_bar);
      }
      if (!tc::mbar_wait(mma_bar, (uint32_t)((s - 1) & 1), 2 * WATCHDOG_CYCLES)) {
        atomicExch(status, 1);
        s_dead = 1;
      }
      tc::tcgen05_fence_after();
      {
        uint32_t r0[NG], r1[NG], r2[NG], r3[NG];
        const uint32_t tq = tmem + ((uint32_t)(warp * 32) << 16) + D_COL;
        tc::tmem_ld16(tq, r0);
        tc::tmem_ld16(tq + NG, r1);
        tc::tmem_ld16(tq + 2 * NG, r2);
        tc::tmem_ld16(tq + 3 * NG, r3);
        tc::tmem_ld_wait();
#pragma unroll
        for (int n = 0; n < NG; ++n)
          sZ[(warp * NG + n) * 32 + lane] = (__uint_as_float(r0[n]) + __uint_as_float(r1[n])) +
                                            (__uint_as_float(r2[n]) + __uint_as_float(r3[n]));
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
#pragma unroll
    for (int i = 0; i < NPT; ++i) {
      gi[i] = asr::hard_sigmoid(z[i][0] + zx[i][0] + bias[0]);
      gf[i] = asr::hard_sigmoid(z[i][1] + zx[i][1] + bias[1]);
      gg[i] = asr::tanh_fast(z[i][2] + zx[i][2] + bias[2]);
      go[i] = asr::hard_sigmoid(z[i][3] + zx[i][3] + bias[3]);
      c_state[i] = gf[i] * c_state[i] + gi[i] * gg[i];
      hv[i] = go[i] * asr::tanh_fast(c_state[i]);
      sP[(warp * NPT + i) * 32 + lane] = __float2half_rn(hv[i] * mu[i]);       // h * B_U, fp16
    }
    __syncwarp();
    if (s + 1 < T) {
      // push my warp's 4 x 64-byte rows into every peer's B operand (buffer s & 1), then announce them
      const uint4 v = *reinterpret_cast<const uint4*>(sP + s_n * 32 + s_c4 * 8);
      const uint32_t dst_local = sB_addr + (uint32_t)((s & 1) * SB_BYTES) + s_off;
#pragma unroll
      for (int it = 0; it < NCTA / 2; ++it) st_cluster_v4(mapa(dst_local, (uint32_t)(2 * it + s_half)), v);
      __syncwarp();
      if (lane < NCTA) mbar_arrive_remote(mapa(bar_addr + (uint32_t)((s & 1) * 8), (uint32_t)lane));
    }
    side_stores(t, hv, gi, gf, gg, go, c_state);
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                            // nobody leaves while a peer may still write into its smem
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// ---- host ---------------------------------------------------------------------------------------
bool supports_fwd(const asr_lstm_fwd_args* a) {
  return a->U16 && a->h16 && a->T >= 1 && (a->H == 512 || a->H == 256 || a->H == 128) && a->N >= NG &&
         a->N % NG == 0 && a->N / NG <= 8;
}

template <int H>
static int32_t launch_fwd(const asr_lstm_fwd_args* a, cudaStream_t st) {
  constexpr int KC = H / 64;
  const size_t smem = 1024 + (size_t)2 * KC * NG * 128 + 4 * NG * 32 * 4 + NG * 32 * 2 + 64;
  const int G = a->N / NG;
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaFuncSetAttribute(fwd_kernel<H>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, 1024, st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(H / UPC, 2, G);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = H / UPC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  asr_lstm_fwd_args args = *a;
  int* flags = a->flags;
  ASR_CUDA(cudaLaunchKernelEx(&cfg, fwd_kernel<H>, args, flags));
  asr::count_launch();
  return ASR_OK;
}

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  switch (a->H) {
    case 128: return launch_fwd<128>(a, st);
    case 256: return launch_fwd<256>(a, st);
    case 512: return launch_fwd<512>(a, st);
  }
  asr::set_error("lstmtc3: unsupported H=%d", a->H);
  return ASR_ERR_INVALID;
}

}  // namespace lstmtc3
