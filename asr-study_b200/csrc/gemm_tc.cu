// K2/K5 — TN GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//   C[M,N] (+)= alpha * A[M,K] * B[N,K]^T + bias        A,B 16-bit K-major, fp32 accumulate
//
// One 128x128 output tile per CTA, BK = 64 (one 128-byte swizzle atom), 3-stage TMA->smem ring,
// two CTAs resident per SM so one CTA's epilogue overlaps the other's main loop.
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = TMEM allocator + MMA issuer
// (one lane), warps 2..5 = epilogue (tcgen05.ld -> registers -> global), one TMEM lane quarter each.
// Optional split-K over gridDim.z (fp32 C, red.global.add) for short-M/N, long-K gradient GEMMs.
#include "common.cuh"
#include "tc.cuh"
#include <cudaTypedefs.h>
#include <mutex>

namespace gemm_tc {

constexpr int BM = 128, BN = 128, BK = 64, STAGES = 3, THREADS = 192;
constexpr int STAGE_BYTES = (BM + BN) * BK * 2;                    // 32 KB
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 128;

struct Params {
  int M, N, K;
  void* C;
  int64_t ldc;
  const float* bias;
  float alpha;
  int dtype_out, accumulate, ab_fmt;
  int kb_per_split;   // k-blocks per gridDim.z slice
};

__global__ void __launch_bounds__(THREADS, 2)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb_total = (p.K + BK - 1) / BK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int nkb = min(p.kb_per_split, nkb_total - kb0);

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tmA);
    tc::tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(full + s, 1);
      tc::mbar_init(empty + s, 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::fence_mbar_init();
  }
  if (warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tcgen05_fence_before();
  __syncthreads();
  tc::tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (tc::elect_one_sync()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!tc::mbar_wait(empty + s, ph ^ 1)) __trap();
        uint8_t* a = smem + s * STAGE_BYTES;
        uint8_t* b = a + BM * BK * 2;
        tc::mbar_expect_tx(full + s, STAGE_BYTES);
        tc::tma_load_2d(a, &tmA, full + s, (kb0 + kb) * BK, m0);
        tc::tma_load_2d(b, &tmB, full + s, (kb0 + kb) * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one_sync()) {
      const uint32_t idesc = tc::umma_idesc_f16(BM, BN, p.ab_fmt);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        if (!tc::mbar_wait(full + s, ph)) __trap();
        tc::tcgen05_fence_after();
        const uint32_t a = tc::smem_u32(smem + s * STAGE_BYTES);
        const uint64_t ad = tc::umma_desc_sw128(a), bd = tc::umma_desc_sw128(a + BM * BK * 2);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)   // +32 bytes per UMMA_K inside the swizzle atom
          tc::umma_ss(tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0);
        tc::umma_commit(empty + s);          // frees the smem stage once these MMAs have read it
      }
      tc::umma_commit(tmem_full);            // accumulator complete
    }
  } else {
    // ---- epilogue: warps 2..5 own TMEM lane quarters (warp % 4) ----
    // TMEM -> registers (thread = one accumulator row) -> per-warp smem transpose -> COALESCED global stores
    // (8 lanes x 16 B = one 128-byte row segment; 4 rows per instruction).  The staging area is pipeline stage 0,
    // free by now: the accumulator-complete barrier implies every TMA load landed and every MMA has read it.
    const int q = warp & 3;
    if (!tc::mbar_wait(tmem_full, 0)) __trap();
    tc::tcgen05_fence_after();
    float* stg = reinterpret_cast<float*>(smem) + q * (32 * 36);     // [32 rows][36] (16-byte aligned, conflict-light)
    const bool split = gridDim.z > 1;
    const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
    const int sub = lane >> 3, l8 = lane & 7;                         // 4 rows per pass, 8 lanes per row
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
      if (n0 + c0 >= p.N) continue;                                   // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(stg + lane * 36 + j) =
            (nkb > 0) ? make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3]))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncwarp();
      const int col = n0 + c0 + l8 * 4;
      float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_bias) {
        if (col < p.N) bv.x = __ldg(p.bias + col);
        if (col + 1 < p.N) bv.y = __ldg(p.bias + col + 1);
        if (col + 2 < p.N) bv.z = __ldg(p.bias + col + 2);
        if (col + 3 < p.N) bv.w = __ldg(p.bias + col + 3);
      }
#pragma unroll
      for (int pass = 0; pass < 8; ++pass) {
        const int rl = pass * 4 + sub;
        const int row = m0 + q * 32 + rl;
        const float4 a = *reinterpret_cast<const float4*>(stg + rl * 36 + l8 * 4);
        float v[4] = {p.alpha * a.x + bv.x, p.alpha * a.y + bv.y, p.alpha * a.z + bv.z, p.alpha * a.w + bv.w};
        if (row >= p.M || col >= p.N) continue;
        const int64_t o = (int64_t)row * p.ldc + col;
        const int nc = min(4, p.N - col);
        if (p.dtype_out == 0) {
          float* cp = reinterpret_cast<float*>(p.C) + o;
          if (split) {
            for (int j = 0; j < nc; ++j) atomicAdd(cp + j, v[j]);
          } else if (nc == 4 && !p.accumulate && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0)) {
            *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
            for (int j = 0; j < nc; ++j) cp[j] = p.accumulate ? cp[j] + v[j] : v[j];
          }
        } else if (p.dtype_out == 1) {
          __half* cp = reinterpret_cast<__half*>(p.C) + o;
          if (nc == 4 && ((reinterpret_cast<uintptr_t>(cp) & 7) == 0)) {     // one 8-byte store per thread
            const __half2 lo = __floats2half2_rn(v[0], v[1]), hi = __floats2half2_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(cp) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
          } else {
            for (int j = 0; j < nc; ++j) cp[j] = __float2half_rn(v[j]);
          }
        } else {
          __nv_bfloat16* cp = reinterpret_cast<__nv_bfloat16*>(p.C) + o;
          if (nc == 4 && ((reinterpret_cast<uintptr_t>(cp) & 7) == 0)) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            *reinterpret_cast<uint2*>(cp) = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
          } else {
            for (int j = 0; j < nc; ++j) cp[j] = __float2bfloat16_rn(v[j]);
          }
        }
      }
      __syncwarp();
    }
  }
  tc::tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// ---- host --------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static std::once_flag g_once;

static bool make_map(CUtensorMap* tm, const void* base, int64_t rows, int64_t cols, int64_t ld, int ab_fmt) {
  std::call_once(g_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  });
  if (!g_encode) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = ab_fmt ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  return g_encode(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool supports(int dtype_in, int dtype_out, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc) {
  (void)dtype_in; (void)dtype_out; (void)ldc;
  return M >= 1 && N >= 1 && K >= 8 && (lda % 8 == 0) && (ldb % 8 == 0) && ((M + BM - 1) / BM) <= 65535;
}

int32_t run(int dtype_in, int dtype_out, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
            void* C, int64_t ldc, const float* bias, float alpha, int accumulate, cudaStream_t st) {
  CUtensorMap tmA, tmB;
  if (!make_map(&tmA, A, M, K, lda, dtype_in) || !make_map(&tmB, B, N, K, ldb, dtype_in)) {
    asr::set_error("gemm_tc: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%lld ldb=%lld)", M, N, K, (long long)lda,
                   (long long)ldb);
    return ASR_ERR_CUDA;
  }
  static bool attr_set = false;
  if (!attr_set) {
    ASR_CUDA(cudaFuncSetAttribute(gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  const int nkb = (K + BK - 1) / BK;
  // split-K when the tile grid cannot fill 148 SMs x 2 CTAs and K is long (fp32 C only)
  int split = 1;
  if (dtype_out == 0 && gx * gy < 148 && nkb >= 64) {
    split = (2 * 148 + gx * gy - 1) / (gx * gy);
    if (split > nkb / 16) split = nkb / 16;
    if (split < 1) split = 1;
  }
  Params p;
  p.M = M; p.N = N; p.K = K; p.C = C; p.ldc = ldc; p.bias = bias; p.alpha = alpha;
  p.dtype_out = dtype_out; p.accumulate = accumulate; p.ab_fmt = dtype_in;
  p.kb_per_split = (nkb + split - 1) / split;
  split = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  if (split > 1 && !accumulate) {   // partial sums are reduced with red.global.add into a zeroed C
    if (ldc == N) {
      ASR_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
    } else {
      ASR_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, st));
    }
  }
  gemm_kernel<<<dim3(gx, gy, split), THREADS, SMEM_BYTES, st>>>(tmA, tmB, p);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

}  // namespace gemm_tc
