// TN GEMM, legacy tensor path (mma.sync m16n8k16 + ldmatrix + cp.async).
//   C[M,N] (+)= alpha * A[M,K] * B[N,K]^T + bias
// This is the bring-up / cross-check GEMM: simple, shape-agnostic, used by the
// tests as an on-device second opinion for the tcgen05 GEMM (gemm_tc.cu) and as
// the engine for shapes the tcgen05 kernel does not take (K % 64 != 0 tails are
// handled by that kernel itself; this one only needs K % 8 == 0).
#include "common.cuh"

namespace gemm_mma {

constexpr int BM = 128, BN = 128, BK = 32, PAD = 8, THREADS = 256;
constexpr int LDS = BK + PAD;   // halves per smem row

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void ldmatrix_x4(unsigned (&r)[4], const void* p) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(s));
}
template <bool BF16>
__device__ __forceinline__ void mma16816(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  if constexpr (BF16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <bool BF16>
__global__ void __launch_bounds__(THREADS)
gemm_kernel(int M, int N, int K, const uint16_t* __restrict__ A, int64_t lda, const uint16_t* __restrict__ B,
            int64_t ldb, void* __restrict__ C, int64_t ldc, int dtype_out, const float* __restrict__ bias,
            float alpha, int accumulate) {
  __shared__ __align__(16) uint16_t As[2][BM][LDS];
  __shared__ __align__(16) uint16_t Bs[2][BN][LDS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 2, wn = warp & 3;           // 2 x 4 warps; warp tile 64 x 32
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  auto load_tile = [&](int stage, int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = tid + i * THREADS;     // 512 chunks of 16 B per operand tile
      const int r = c >> 2, kc = (c & 3) * 8;
      {
        const int gr = m0 + r, gk = k0 + kc;
        const bool ok = (gr < M) && (gk < K);
        cp_async16(&As[stage][r][kc], A + (ok ? (int64_t)gr * lda + gk : 0), ok ? 16 : 0);
      }
      {
        const int gr = n0 + r, gk = k0 + kc;
        const bool ok = (gr < N) && (gk < K);
        cp_async16(&Bs[stage][r][kc], B + (ok ? (int64_t)gr * ldb + gk : 0), ok ? 16 : 0);
      }
    }
  };

  float acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.0f;

  const int nk = (K + BK - 1) / BK;
  load_tile(0, 0);
  cp_async_commit();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tile((kt + 1) & 1, (kt + 1) * BK);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int st = kt & 1;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 16) {
      unsigned af[4][4], bf[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        ldmatrix_x4(af[i], &As[st][wm * 64 + i * 16 + (lane & 15)][kk + (lane >> 4) * 8]);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        ldmatrix_x4(bf[j], &Bs[st][wn * 32 + j * 16 + (lane & 7) + ((lane >> 4) << 3)][kk + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma16816<BF16>(acc[i][j], af[i], bf[j >> 1][(j & 1) * 2], bf[j >> 1][(j & 1) * 2 + 1]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = m0 + wm * 64 + i * 16 + (lane >> 2) + ((e >> 1) << 3);
        const int c = n0 + wn * 32 + j * 8 + (lane & 3) * 2 + (e & 1);
        if (r < M && c < N) {
          float v = alpha * acc[i][j][e] + (bias ? bias[c] : 0.0f);
          const int64_t o = (int64_t)r * ldc + c;
          if (dtype_out == 0) {
            float* cp = reinterpret_cast<float*>(C);
            cp[o] = accumulate ? cp[o] + v : v;
          } else if (dtype_out == 1) {
            reinterpret_cast<__half*>(C)[o] = __float2half_rn(v);
          } else {
            reinterpret_cast<__nv_bfloat16*>(C)[o] = __float2bfloat16_rn(v);
          }
        }
      }
}

int32_t run(int dtype_in, int dtype_out, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
            void* C, int64_t ldc, const float* bias, float alpha, int accumulate, cudaStream_t st) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  ASR_CHECK_ARG(grid.y <= 65535, "gemm: M too large for this grid");
  if (dtype_in == 1)
    gemm_kernel<true><<<grid, THREADS, 0, st>>>(M, N, K, (const uint16_t*)A, lda, (const uint16_t*)B, ldb, C, ldc,
                                                dtype_out, bias, alpha, accumulate);
  else
    gemm_kernel<false><<<grid, THREADS, 0, st>>>(M, N, K, (const uint16_t*)A, lda, (const uint16_t*)B, ldb, C, ldc,
                                                 dtype_out, bias, alpha, accumulate);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

}  // namespace gemm_mma
