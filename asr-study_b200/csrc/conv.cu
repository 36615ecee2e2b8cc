// Convolutional front end of BASELINE configs[3] ("DeepSpeech2-style 2 x Conv + 5 x BiLSTM-800"; not in the reference,
// README.md:118 lists it as future work) as implicit GEMM: the patch gather (im2col), its transposed copy for the weight
// gradient, the gather-form col2im of the input gradient and the clipped-ReLU epilogues live here; the contractions run
// on the tcgen05 GEMM engines (asr_gemm_tn).  All of it is HBM-bound layout work: 16-byte accesses along the channel /
// kernel-frequency axis, grids of a few CTAs per SM.
//
// Activations are time-major [T, N, F, C] (frame, utterance, frequency bin, channel), so the output of the last layer,
// rows (t', n, f') x channels, IS the [T', N, F' * C] input of the first BiLSTM without a copy.
// Patch matrix: rows (t', n, f'), columns (kt, kf, c) with c fastest; K is padded to a multiple of 8 (16-byte GEMM rows).
#include "common.cuh"

namespace {

struct ConvGeom {
  int T, N, F, C;        // input
  int To, Fo;            // output frames / bins
  int kt, kf, st, sf, pt, pf;
  int K, Kp;             // kt * kf * C and its padding to 8
};

__device__ __forceinline__ float load_in(const void* x, int dtype, size_t i) {
  return dtype == 2 ? reinterpret_cast<const float*>(x)[i] : __half2float(reinterpret_cast<const __half*>(x)[i]);
}

// row-major patches: one thread per (row, tap, 8-channel group) — 16-byte loads / stores, consecutive threads on
// consecutive 16-byte chunks of a patch row — or per (row, tap) element when C is not a multiple of 8 (C = 1: the first layer)
__global__ void __launch_bounds__(256)
im2col_kernel(const void* __restrict__ x, int x_dtype, ConvGeom g, __half* __restrict__ out, int64_t ld) {
  const int64_t M = (int64_t)g.To * g.N * g.Fo;
  const int runs = g.kt * g.kf;
  const bool vec = (g.C % 8 == 0) && x_dtype == 0;
  const int cg = vec ? g.C / 8 : g.C;
  const int64_t total = M * runs * cg;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cg);
    const int r = (int)((i / cg) % runs);
    const int64_t row = i / ((int64_t)cg * runs);
    const int fo = (int)(row % g.Fo);
    const int n = (int)((row / g.Fo) % g.N);
    const int to = (int)(row / ((int64_t)g.Fo * g.N));
    const int dkt = r / g.kf, dkf = r - dkt * g.kf;
    const int t = to * g.st - g.pt + dkt, f = fo * g.sf - g.pf + dkf;
    const bool inside = t >= 0 && t < g.T && f >= 0 && f < g.F;
    const size_t src = (((size_t)(inside ? t : 0) * g.N + n) * g.F + (inside ? f : 0)) * g.C;
    if (vec) {
      const uint4 v = inside ? reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(x) + src)[c] : make_uint4(0u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(out + (size_t)row * ld + (size_t)r * g.C + 8 * c) = v;
    } else {
      out[(size_t)row * ld + (size_t)r * g.C + c] = __float2half_rn(inside ? load_in(x, x_dtype, src + c) : 0.0f);
    }
  }
}

// fp16 [rows, ld] -> bf16 [cols, ldT] through 32 x 32 shared-memory tiles (the K-major operand of the dW GEMM)
__global__ void __launch_bounds__(256)
transpose16_kernel(const __half* __restrict__ in, int64_t ld, int64_t rows, int cols, __nv_bfloat16* __restrict__ out, int64_t ldT) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  for (int64_t tIdx = blockIdx.x; tIdx < tiles_c * tiles_r; tIdx += gridDim.x) {
    const int64_t r0 = (tIdx / tiles_c) * 32;
    const int c0 = (int)(tIdx % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + ty + 8 * j;
      const int c = c0 + tx;
      tile[ty + 8 * j][tx] = (r < rows && c < cols) ? __half2float(in[r * ld + c]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + ty + 8 * j;
      const int64_t r = r0 + tx;
      if (c < cols && r < rows) out[(size_t)c * ldT + r] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
    }
    __syncthreads();
  }
}

// zero the K padding columns of the row-major patch matrix
__global__ void __launch_bounds__(256) pad_cols_kernel(__half* out, int64_t M, int K, int Kp, int64_t ld) {
  const int w = Kp - K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M * w; i += (int64_t)gridDim.x * blockDim.x)
    out[(i / w) * ld + K + (i % w)] = __float2half_rn(0.0f);
}

// dx[t, n, f, c] = sum over the taps (kt, kf) that saw this input position of dpatch[(t', n, f'), (kt, kf, c)]
__global__ void __launch_bounds__(256)
col2im_kernel(const __nv_bfloat16* __restrict__ dp, int64_t ld, ConvGeom g, float* __restrict__ dx) {
  const int64_t total = (int64_t)g.T * g.N * g.F * g.C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    const int f = (int)((i / g.C) % g.F);
    const int n = (int)((i / ((int64_t)g.C * g.F)) % g.N);
    const int t = (int)(i / ((int64_t)g.C * g.F * g.N));
    float acc = 0.0f;
    for (int dkt = 0; dkt < g.kt; ++dkt) {
      const int tn = t + g.pt - dkt;
      if (tn < 0 || tn % g.st) continue;
      const int to = tn / g.st;
      if (to >= g.To) continue;
      for (int dkf = 0; dkf < g.kf; ++dkf) {
        const int fn = f + g.pf - dkf;
        if (fn < 0 || fn % g.sf) continue;
        const int fo = fn / g.sf;
        if (fo >= g.Fo) continue;
        const size_t row = ((size_t)to * g.N + n) * g.Fo + fo;
        acc += __bfloat162float(dp[row * ld + (size_t)(dkt * g.kf + dkf) * g.C + c]);
      }
    }
    dx[i] = acc;
  }
}

// y = min(max(z, 0), clip): fp32 GEMM output (bias already added) -> fp32 and / or fp16 activation
__global__ void __launch_bounds__(256)
crelu_fwd_kernel(const float* __restrict__ z, int64_t n, float clip, float* __restrict__ y32, __half* __restrict__ y16) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = fminf(fmaxf(z[i], 0.0f), clip);
    if (y32) y32[i] = v;
    if (y16) y16[i] = __float2half_rn(v);
  }
}

// g_in = g_out * [0 < y < clip]  ->  bf16 [M, C] (dX operand) and bf16 [C, M] (dW operand); y fp32 or fp16
__global__ void __launch_bounds__(256)
crelu_bwd_kernel(const float* __restrict__ gout, const void* __restrict__ y, int y_dtype, int64_t M, int C, float clip,
                 __nv_bfloat16* __restrict__ g16, __nv_bfloat16* __restrict__ gT16, float* __restrict__ g32) {
  const int64_t n = M * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = load_in(y, y_dtype, (size_t)i);
    const float v = (a > 0.0f && a < clip) ? gout[i] : 0.0f;
    const int64_t row = i / C;
    const int c = (int)(i - row * C);
    if (g16) g16[i] = __float2bfloat16_rn(v);
    if (gT16) gT16[(size_t)c * M + row] = __float2bfloat16_rn(v);
    if (g32) g32[i] = v;
  }
}

inline int grid_1d(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 148 * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

bool make_geom(const asr_conv_geom* a, ConvGeom* g) {
  g->T = a->T; g->N = a->N; g->F = a->F; g->C = a->C;
  g->kt = a->kt; g->kf = a->kf; g->st = a->st; g->sf = a->sf; g->pt = a->pt; g->pf = a->pf;
  if (g->T < 1 || g->N < 1 || g->F < 1 || g->C < 1 || g->kt < 1 || g->kf < 1 || g->st < 1 || g->sf < 1 || g->pt < 0 || g->pf < 0)
    return false;
  if (g->T + 2 * g->pt < g->kt || g->F + 2 * g->pf < g->kf) return false;
  g->To = (g->T + 2 * g->pt - g->kt) / g->st + 1;
  g->Fo = (g->F + 2 * g->pf - g->kf) / g->sf + 1;
  g->K = g->kt * g->kf * g->C;
  g->Kp = (g->K + 7) / 8 * 8;
  return true;
}

}  // namespace

extern "C" int32_t asr_conv_out_shape(const asr_conv_geom* geom, int32_t* t_out, int32_t* f_out, int32_t* k, int32_t* k_padded) {
  ConvGeom g;
  ASR_CHECK_ARG(geom && make_geom(geom, &g), "asr_conv_out_shape: bad geometry");
  if (t_out) *t_out = g.To;
  if (f_out) *f_out = g.Fo;
  if (k) *k = g.K;
  if (k_padded) *k_padded = g.Kp;
  return ASR_OK;
}

extern "C" int32_t asr_conv_im2col(const void* x, int32_t x_dtype, const asr_conv_geom* geom, void* patches16, int64_t ld,
                                   void* patchesT16, int64_t ldT, void* stream) {
  ConvGeom g;
  ASR_CHECK_ARG(x && geom && make_geom(geom, &g), "asr_conv_im2col: bad geometry");
  ASR_CHECK_ARG(x_dtype == 0 || x_dtype == 2, "asr_conv_im2col: input is fp16 (0) or fp32 (2)");
  const int64_t M = (int64_t)g.To * g.N * g.Fo;
  cudaStream_t st = (cudaStream_t)stream;
  ASR_CHECK_ARG(patches16 && ld >= g.Kp && ld % 8 == 0, "asr_conv_im2col: patches16 with ld >= padded K (multiple of 8) is required");
  const int cg = (g.C % 8 == 0 && x_dtype == 0) ? g.C / 8 : g.C;
  im2col_kernel<<<grid_1d(M * g.kt * g.kf * cg), 256, 0, st>>>(x, x_dtype, g, reinterpret_cast<__half*>(patches16), ld);
  ASR_LAUNCH_CHECK();
  if (g.Kp > g.K) {
    pad_cols_kernel<<<grid_1d(M * (g.Kp - g.K)), 256, 0, st>>>(reinterpret_cast<__half*>(patches16), M, g.K, g.Kp, ld);
    ASR_LAUNCH_CHECK();
  }
  if (patchesT16) {                                        // bf16 [K, rows]: the K-major operand of the weight-gradient GEMM
    ASR_CHECK_ARG(ldT >= M && ldT % 8 == 0, "asr_conv_im2col: ldT >= rows (multiple of 8)");
    const int64_t tiles = ((M + 31) / 32) * ((g.K + 31) / 32);
    transpose16_kernel<<<(int)(tiles < 148 * 16 ? tiles : 148 * 16), 256, 0, st>>>(
        reinterpret_cast<const __half*>(patches16), ld, M, g.K, reinterpret_cast<__nv_bfloat16*>(patchesT16), ldT);
    ASR_LAUNCH_CHECK();
  }
  return ASR_OK;
}

extern "C" int32_t asr_conv_col2im(const void* dpatches16, int64_t ld, const asr_conv_geom* geom, float* dx, void* stream) {
  ConvGeom g;
  ASR_CHECK_ARG(dpatches16 && dx && geom && make_geom(geom, &g) && ld >= g.K, "asr_conv_col2im: bad argument");
  col2im_kernel<<<grid_1d((int64_t)g.T * g.N * g.F * g.C), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dpatches16), ld, g, dx);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_clipped_relu(const float* z, int64_t n, float clip, float* y32, void* y16, void* stream) {
  ASR_CHECK_ARG(z && n > 0 && clip > 0.0f && (y32 || y16), "asr_clipped_relu: bad argument");
  crelu_fwd_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(z, n, clip, y32, reinterpret_cast<__half*>(y16));
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_clipped_relu_backward(const float* gout, const void* y, int32_t y_dtype, int64_t rows, int32_t cols,
                                             float clip, void* g16, void* gT16, float* g32, void* stream) {
  ASR_CHECK_ARG(gout && y && rows > 0 && cols > 0 && (y_dtype == 0 || y_dtype == 2), "asr_clipped_relu_backward: bad argument");
  crelu_bwd_kernel<<<grid_1d(rows * cols), 256, 0, (cudaStream_t)stream>>>(
      gout, y, y_dtype, rows, cols, clip, reinterpret_cast<__nv_bfloat16*>(g16), reinterpret_cast<__nv_bfloat16*>(gT16), g32);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
