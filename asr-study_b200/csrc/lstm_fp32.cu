// K3/K4 (fp32 engine) — persistent BiLSTM recurrence, forward and BPTT, sm_100a.
//
// Both directions run in ONE cooperative launch: grid = (H/8 CTAs, 2 dirs).
// Each CTA owns 8 hidden units of one direction, keeps its slice of the
// recurrent matrix U resident in shared memory for all T steps, keeps the cell
// state c (forward) / the dc, db carries (backward) in registers, and exchanges
// h_t (forward) / dz_t (backward) with its 63 peers through a small
// double-buffered, L2-resident global buffer guarded by a release/acquire step
// counter per direction.  The per-step panel product is a register-tiled
// (4 batch x 8 column) fp32 FMA loop with K split across thread groups.
//
// This engine computes in exact fp32 (it is the numerics reference the
// tensor-core engine in lstm_tc.cu is tested against, and the product path for
// hidden sizes / batch sizes the tensor-core engine does not cover).
//
// Semantics: core/layers.py:432-469 (LSTM.step; i,f,c,o; hard_sigmoid; tanh),
// Keras-1 Bidirectional without masking (core/models.py:68-70, 261-271).
#include "common.cuh"
#include <cooperative_groups.h>

namespace lstm32 {

constexpr int UPC = 8;            // hidden units per CTA
constexpr int NB = 32;            // batch rows per CTA (N <= 32 per launch group)
constexpr int THREADS = 256;
constexpr long long WATCHDOG_CYCLES = 6000000000LL;   // ~3 s

struct Scratch {        // lives in args.flags (zeroed by the host wrapper)
  int step_flag[2][32]; // [dir][0] used; padded to separate cache lines
  int status;           // 0 ok, 1 watchdog fired
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// wait until *flag >= target; returns false if the watchdog fired
__device__ __forceinline__ bool wait_flag(const int* flag, int target, int* status) {
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    while (ld_acquire(flag) < target) {
      if (clock64() - t0 > WATCHDOG_CYCLES || ld_acquire(status) != 0) {
        atomicExch(status, 1);
        break;
      }
    }
  }
  __syncthreads();
  return true;
}

// partial[n][c] (+)= sum_k A[k][n] * B[k][c] over this thread's K slice.
// A: smem [K][NB] (batch contiguous), B: smem [K][CC] (columns contiguous).
// thread tile: 4 batch rows x 8 columns.
template <int CC>
__device__ __forceinline__ void panel_fma(const float* __restrict__ A, const float* __restrict__ B, int k_begin,
                                          int k_end, int nq, int cg, float (&acc)[4][8]) {
  for (int k = k_begin; k < k_end; ++k) {
    const float4 a = *reinterpret_cast<const float4*>(A + (size_t)k * NB + nq * 4);
    const float4 b0 = *reinterpret_cast<const float4*>(B + (size_t)k * CC + cg * 8);
    const float4 b1 = *reinterpret_cast<const float4*>(B + (size_t)k * CC + cg * 8 + 4);
    const float av[4] = {a.x, a.y, a.z, a.w};
    const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
lstm_fwd_kernel(asr_lstm_fwd_args a, float* __restrict__ xbuf /* [2 dir][2 parity][H][NB] */, Scratch* sc) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int dir = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x;
  const int T = a.T, N = a.N, H = a.H;
  const int u0 = cta * UPC;
  constexpr int CC = 4 * UPC;  // 32 gate columns, c = u_local*4 + gate

  float* sU = smem;                       // [H][CC]
  float* sH = sU + (size_t)H * CC;        // [H][NB]
  float* sR = sH + (size_t)H * NB;        // [8 kslices][NB][CC]

  // one-time: U slice, gather Keras layout U[dir][k][g*H + u] -> sU[k][u_l*4+g]
  const float* Ug = a.U + (size_t)dir * H * 4 * H;
  for (int i = tid; i < H * CC; i += THREADS) {
    const int k = i / CC, c = i % CC, ul = c >> 2, g = c & 3;
    const int u = u0 + ul;
    sU[i] = (u < H) ? Ug[(size_t)k * 4 * H + g * H + u] : 0.0f;
  }

  // epilogue ownership: thread -> (n, u_local)
  const int en = tid >> 3, eu = tid & 7;
  const int u = u0 + eu;
  const bool own = (en < N) && (u < H);
  float bias[4] = {0, 0, 0, 0};
  if (own)
#pragma unroll
    for (int g = 0; g < 4; ++g) bias[g] = a.bias[(size_t)dir * 4 * H + g * H + u];
  float c_state = 0.0f;
  const float mu = (own && a.mask_u) ? a.mask_u[((size_t)dir * N + en) * H + u] : 1.0f;   // B_U, constant over time

  // matmul ownership: 8 K-slices x (8 n-quads x 4 column groups)
  const int ks = tid >> 5, lane = tid & 31, nq = lane & 7, cg = lane >> 3;
  const int kper = (H + 7) / 8;
  const int kb = min(ks * kper, H), ke = min(kb + kper, H);

  int* flag = &sc->step_flag[dir][0];
  float* xb = xbuf + (size_t)dir * 2 * H * NB;
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    // prefetch the input projection for this step
    float zx[4] = {0, 0, 0, 0};
    if (own) {
      const float* zr = a.zx + (((size_t)t * N + en) * 2 + dir) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) zx[g] = __ldg(zr + g * H + u);
    }
    float z[4] = {0, 0, 0, 0};
    if (s > 0) {
      wait_flag(flag, nctas * s, &sc->status);
      // load h_{prev} [H][NB] from the exchange buffer (written by all CTAs of this direction)
      const float4* src = reinterpret_cast<const float4*>(xb + (size_t)((s - 1) & 1) * H * NB);
      float4* dst = reinterpret_cast<float4*>(sH);
      for (int i = tid; i < H * NB / 4; i += THREADS) dst[i] = __ldcg(src + i);
      __syncthreads();
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      panel_fma<CC>(sH, sU, kb, ke, nq, cg, acc);
      float* r = sR + (size_t)ks * NB * CC;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4* rp = reinterpret_cast<float4*>(r + (size_t)(nq * 4 + i) * CC + cg * 8);
        rp[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        rp[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 v = *reinterpret_cast<const float4*>(sR + (size_t)q * NB * CC + (size_t)en * CC + eu * 4);
        z[0] += v.x; z[1] += v.y; z[2] += v.z; z[3] += v.w;
      }
    }
    float h = 0.0f;
    if (own) {
      const float gi = asr::hard_sigmoid(z[0] + zx[0] + bias[0]);
      const float gf = asr::hard_sigmoid(z[1] + zx[1] + bias[1]);
      const float gg = tanhf(z[2] + zx[2] + bias[2]);
      const float go = asr::hard_sigmoid(z[3] + zx[3] + bias[3]);
      c_state = gf * c_state + gi * gg;
      h = go * tanhf(c_state);
      const size_t row = (size_t)t * N + en;
      if (a.h32) a.h32[row * 2 * H + dir * H + u] = h;
      if (a.h16) reinterpret_cast<__half*>(a.h16)[row * 2 * H + dir * H + u] = __float2half_rn(h);
      if (a.training) {
        float* gp = a.gates + (row * 2 + dir) * 4 * H;
        gp[u] = gi; gp[H + u] = gf; gp[2 * H + u] = gg; gp[3 * H + u] = go;
        a.cell[(row * 2 + dir) * H + u] = c_state;
        if (a.hT16)
          reinterpret_cast<__nv_bfloat16*>(a.hT16)[(size_t)(dir * H + u) * ((size_t)T * N) + row] =
              __float2bfloat16_rn(h * mu);
      }
    }
    if (u < H && en < NB) xb[(size_t)(s & 1) * H * NB + (size_t)u * NB + en] = own ? h * mu : 0.0f;
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release(flag, 1);
    }
  }
}

// ------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
lstm_bwd_kernel(asr_lstm_bwd_args a, float* __restrict__ xbuf /* [2 dir][2 parity][4H][NB] */, Scratch* sc) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int dir = blockIdx.y, cta = blockIdx.x, nctas = gridDim.x;
  const int T = a.T, N = a.N, H = a.H, K4 = 4 * a.H;
  const int u0 = cta * UPC;
  constexpr int CC = UPC;   // 8 output columns = this CTA's units

  float* sU = smem;                        // [4H][8]  sU[k'][u_l] = U[dir][u0+u_l][k']
  float* sD = sU + (size_t)K4 * CC;        // [KCH][NB] dz chunk
  const int KCH = H;                       // chunk of gate columns staged at a time
  float* sR = sD + (size_t)KCH * NB;       // [32 kslices][NB][8]

  const float* Ug = a.U + (size_t)dir * H * K4;
  for (int i = tid; i < K4 * CC; i += THREADS) {
    const int k = i / CC, ul = i % CC;
    const int u = u0 + ul;
    sU[i] = (u < H) ? Ug[(size_t)u * K4 + k] : 0.0f;
  }
  const int en = tid >> 3, eu = tid & 7;
  const int u = u0 + eu;
  const bool own = (en < N) && (u < H);
  float dc_carry = 0.0f;
  float db[4] = {0, 0, 0, 0};
  const float mu = (own && a.mask_u) ? a.mask_u[((size_t)dir * N + en) * H + u] : 1.0f;

  // matmul: 32 K-slices x 8 n-quads, one column group
  const int ks = tid >> 3, nq = tid & 7;
  const int kper = (KCH + 31) / 32;
  const int kb = min(ks * kper, KCH), ke = min(kb + kper, KCH);

  int* flag = &sc->step_flag[dir][0];
  float* xb = xbuf + (size_t)dir * 2 * K4 * NB;
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    // forward order for dir 0 is t = 0..T-1, for dir 1 is t = T-1..0; BPTT walks it backwards
    const int t = dir ? s : (T - 1 - s);
    const int t_prev = dir ? (t + 1) : (t - 1);     // forward-order predecessor (holds c_{prev})
    const bool has_prev = dir ? (t + 1 < T) : (t > 0);
    float dho = 0.0f, gi = 0, gf = 0, gg = 0, go = 0, c = 0, cp = 0;
    const size_t row = (size_t)t * N + en;
    if (own) {
      dho = __ldg(a.dh + row * 2 * H + dir * H + u);
      const float* gp = a.gates + (row * 2 + dir) * 4 * H;
      gi = __ldg(gp + u); gf = __ldg(gp + H + u); gg = __ldg(gp + 2 * H + u); go = __ldg(gp + 3 * H + u);
      c = __ldg(a.cell + (row * 2 + dir) * H + u);
      if (has_prev) cp = __ldg(a.cell + (((size_t)t_prev * N + en) * 2 + dir) * H + u);
    }
    float dh_rec = 0.0f;
    if (s > 0) {
      wait_flag(flag, nctas * s, &sc->status);
      float acc[4][8];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
      const float* xsrc = xb + (size_t)((s - 1) & 1) * K4 * NB;
      for (int ch = 0; ch < 4; ++ch) {
        const float4* src = reinterpret_cast<const float4*>(xsrc + (size_t)ch * KCH * NB);
        float4* dst = reinterpret_cast<float4*>(sD);
        for (int i = tid; i < KCH * NB / 4; i += THREADS) dst[i] = __ldcg(src + i);
        __syncthreads();
        panel_fma<CC>(sD, sU + (size_t)ch * KCH * CC, kb, ke, nq, 0, acc);
        __syncthreads();
      }
      float* r = sR + (size_t)ks * NB * CC;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4* rp = reinterpret_cast<float4*>(r + (size_t)(nq * 4 + i) * CC);
        rp[0] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        rp[1] = make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]);
      }
      __syncthreads();
      for (int q = 0; q < 32; ++q) dh_rec += sR[(size_t)q * NB * CC + (size_t)en * CC + eu];
      dh_rec *= mu;
    }
    float dz[4] = {0, 0, 0, 0};
    if (own) {
      const float dh = dho + dh_rec;
      const float tc = tanhf(c);
      const float d_o = dh * tc * asr::hard_sigmoid_grad(go);
      const float dc = dc_carry + dh * go * (1.0f - tc * tc);
      dz[0] = dc * gg * asr::hard_sigmoid_grad(gi);
      dz[1] = dc * cp * asr::hard_sigmoid_grad(gf);
      dz[2] = dc * gi * (1.0f - gg * gg);
      dz[3] = d_o;
      dc_carry = dc * gf;
#pragma unroll
      for (int g = 0; g < 4; ++g) db[g] += dz[g];
      const size_t zrow = (row * 2 + dir) * 4 * H;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (a.dz32) a.dz32[zrow + g * H + u] = dz[g];
        if (a.dz16) reinterpret_cast<__nv_bfloat16*>(a.dz16)[zrow + g * H + u] = __float2bfloat16_rn(dz[g]);
        if (a.dzT16)
          reinterpret_cast<__nv_bfloat16*>(a.dzT16)[(size_t)(dir * K4 + g * H + u) * ((size_t)T * N) + row] =
              __float2bfloat16_rn(dz[g]);
      }
    }
    if (u < H && en < NB) {
      float* xo = xb + (size_t)(s & 1) * K4 * NB;
#pragma unroll
      for (int g = 0; g < 4; ++g) xo[(size_t)(g * H + u) * NB + en] = dz[g];
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      red_release(flag, 1);
    }
  }
  // bias gradient: reduce over the batch rows of this CTA
  float* red = smem;  // reuse
  __syncthreads();
#pragma unroll
  for (int g = 0; g < 4; ++g) red[(g * NB + en) * UPC + eu] = own ? db[g] : 0.0f;
  __syncthreads();
  if (tid < 4 * UPC) {
    const int g = tid / UPC, ul = tid % UPC;
    if (u0 + ul < H) {
      float sum = 0.0f;
      for (int n = 0; n < NB; ++n) sum += red[(g * NB + n) * UPC + ul];
      a.dbias[(size_t)dir * K4 + g * H + u0 + ul] = sum;
    }
  }
}

}  // namespace lstm32

// host-side dispatch lives in lstm_api.cu
namespace lstm32 {

size_t scratch_bytes(int H) {
  // Scratch header (256-aligned) + the larger of the two exchange buffers
  return 1024 + (size_t)2 * 2 * 4 * H * NB * sizeof(float);
}

int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st) {
  const int H = a->H;
  const int nctas = (H + UPC - 1) / UPC;
  const size_t smem = ((size_t)H * 4 * UPC + (size_t)H * NB + (size_t)8 * NB * 4 * UPC) * sizeof(float);
  ASR_CHECK_ARG(a->N <= NB, "lstm fp32 engine: N=%d > %d per launch", a->N, NB);
  ASR_CHECK_ARG(2 * nctas <= 148 && smem <= 227 * 1024, "lstm fp32 engine: H=%d does not fit (ctas=%d, smem=%zu)", H,
                2 * nctas, smem);
  ASR_CHECK_ARG(a->h32 || a->h16, "lstm forward: no output buffer");
  ASR_CUDA(cudaFuncSetAttribute(lstm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, 1024, st));
  Scratch* sc = reinterpret_cast<Scratch*>(a->flags);
  float* xbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(a->flags) + 1024);
  asr_lstm_fwd_args args = *a;
  void* kargs[] = {&args, &xbuf, &sc};
  ASR_CUDA(cudaLaunchCooperativeKernel((void*)lstm_fwd_kernel, dim3(nctas, 2), dim3(THREADS), kargs, smem, st));
  asr::count_launch();
  return ASR_OK;
}

int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st) {
  const int H = a->H;
  const int nctas = (H + UPC - 1) / UPC;
  const size_t smem = ((size_t)4 * H * UPC + (size_t)H * NB + (size_t)32 * NB * UPC) * sizeof(float);
  ASR_CHECK_ARG(a->N <= NB, "lstm fp32 engine: N=%d > %d per launch", a->N, NB);
  ASR_CHECK_ARG(2 * nctas <= 148 && smem <= 227 * 1024, "lstm fp32 engine: H=%d does not fit (ctas=%d, smem=%zu)", H,
                2 * nctas, smem);
  ASR_CUDA(cudaFuncSetAttribute(lstm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ASR_CUDA(cudaMemsetAsync(a->flags, 0, 1024, st));
  Scratch* sc = reinterpret_cast<Scratch*>(a->flags);
  float* xbuf = reinterpret_cast<float*>(reinterpret_cast<char*>(a->flags) + 1024);
  asr_lstm_bwd_args args = *a;
  void* kargs[] = {&args, &xbuf, &sc};
  ASR_CUDA(cudaLaunchCooperativeKernel((void*)lstm_bwd_kernel, dim3(nctas, 2), dim3(THREADS), kargs, smem, st));
  asr::count_launch();
  return ASR_OK;
}

}  // namespace lstm32
