// Convolutional front end of BASELINE configs[3] ("DeepSpeech2-style 2 x Conv + 5 x BiLSTM-800"; not in the reference,
// README.md:118 lists it as future work) as an implicit GEMM WITHOUT a patch matrix.
//
// Activations live batch-major and zero-padded in time, fp16: xp [N, Tp, W] with W = F * C (channel fastest) and the T
// valid frames of every utterance at rows [pt, pt + T).  Seen as one flat array, the kt consecutive frames a convolution
// window covers are CONTIGUOUS, and the window of output frame t' + 1 starts st * W elements after the window of t': the
// matrix "row (n, t') = the kt * W input values under that window" is a view of xp with row stride st * W < row length —
// overlapping rows, which a TMA tensor map expresses directly (global stride < box width).  The frequency axis is folded
// into the weights instead: Wt[(f', co), (dkt, f, c)] = W[co, dkt, f - sf * f' + pf, c] (zero outside the kernel: a banded
// block-Toeplitz matrix, ~60 % dense for a 21-bin kernel on 20 bins), so one tcgen05 GEMM
//     z[(n, t'), (f', co)] = view(xp)[(n, t'), (dkt, f, c)] . Wt^T
// is the whole layer: no im2col copy (the 41 x 11 / 21 x 11 kernels would expand the input 231-fold), no col2im.  The
// weight gradient contracts the same view with dL/dz (through a transposed copy, 11-fold in time only), the input
// gradient is the mirrored GEMM over the zero-padded dL/dz.  This file holds the layout kernels around those GEMMs: pack,
// weight expansion, activation (+ re-padding), activation backward, unfold-transpose, Toeplitz gradient reduction.
//
// GEMM rows: M = N * Rn with Rn >= T_out rows per utterance (Rn a multiple of 8, Tp = st * Rn >= T + 2 pt); the rows
// t' >= T_out of an utterance are computed on whatever follows in memory and dropped by the activation kernel.
#include "common.cuh"

namespace {

struct Geom {
  int T, N, F, C;             // input frames, utterances, bins, channels
  int kt, kf, st, sf, pt, pf;
  int To, Fo;                 // output frames / bins
  int Rn, Tp;                 // GEMM rows per utterance, padded frames per utterance (= st * Rn)
  int W;                      // F * C
  int K, Kp;                  // kt * W and its padding to 8
};

bool make_geom(const asr_conv_geom* a, Geom* g) {
  g->T = a->T; g->N = a->N; g->F = a->F; g->C = a->C;
  g->kt = a->kt; g->kf = a->kf; g->st = a->st; g->sf = a->sf; g->pt = a->pt; g->pf = a->pf;
  if (g->T < 1 || g->N < 1 || g->F < 1 || g->C < 1 || g->kt < 1 || g->kf < 1 || g->st < 1 || g->sf < 1 || g->pt < 0 || g->pf < 0)
    return false;
  if (g->T + 2 * g->pt < g->kt || g->F + 2 * g->pf < g->kf) return false;
  g->To = (g->T + 2 * g->pt - g->kt) / g->st + 1;
  g->Fo = (g->F + 2 * g->pf - g->kf) / g->sf + 1;
  const int need = (g->T + 2 * g->pt + g->st - 1) / g->st;
  g->Rn = ((need > g->To ? need : g->To) + 7) / 8 * 8;
  g->Tp = g->st * g->Rn;
  g->W = g->F * g->C;
  g->K = g->kt * g->W;
  g->Kp = (g->K + 7) / 8 * 8;
  return (g->st * g->W) % 8 == 0;          // 16-byte row stride of the overlapping view
}

inline int grid_1d(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = 148 * 8;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

// x f32 [T, N, W] (time-major, what the feature kernel emits) -> xp fp16 [N, Tp, W] rows [pt, pt + T); the padding rows are
// zeroed by the caller once (nothing ever writes them)
__global__ void __launch_bounds__(256)
pack_kernel(const float* __restrict__ x, int T, int N, int W, __half* __restrict__ xp, int Tp, int pt) {
  const int64_t total = (int64_t)T * N * W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const int n = (int)((i / W) % N);
    const int t = (int)(i / ((int64_t)W * N));
    xp[((size_t)n * Tp + pt + t) * W + w] = __float2half_rn(x[i]);
  }
}

// W f32 [Co, kt, kf, C] -> wt fp16 [Fo * Co, ldw] (forward B operand), w2 bf16 [F * C, kt * Fo * Co] (input-gradient B
// operand, taps mirrored in time), bias tiled over the output bins
__global__ void __launch_bounds__(256)
toeplitz_kernel(const float* __restrict__ w, const float* __restrict__ b, Geom g, int Co, __half* __restrict__ wt, int64_t ldw,
                __nv_bfloat16* __restrict__ w2, float* __restrict__ bias_t) {
  const int Wo = g.Fo * Co;
  const int64_t total = (int64_t)Wo * g.Kp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % g.Kp);
    const int j = (int)(i / g.Kp);
    const int co = j % Co, fo = j / Co;
    float v = 0.0f;
    int dkt = 0, f = 0, c = 0;
    if (k < g.K) {
      c = k % g.C;
      f = (k / g.C) % g.F;
      dkt = k / g.W;
      const int dkf = f - g.sf * fo + g.pf;
      if (dkf >= 0 && dkf < g.kf) v = w[(((size_t)co * g.kt + dkt) * g.kf + dkf) * g.C + c];
    }
    wt[(size_t)j * ldw + k] = __float2half_rn(v);
    if (w2 && k < g.K) w2[(size_t)(f * g.C + c) * ((size_t)g.kt * Wo) + (size_t)(g.kt - 1 - dkt) * Wo + j] = __float2bfloat16_rn(v);
    if (k == 0) bias_t[j] = b[co];
  }
}

// z f32 [N * Rn, Wo] (GEMM output, bias added) -> clipped ReLU -> fp16 rows [row0, row0 + To) of a [N, rows_per_utt, Wo]
// buffer (the next layer's padded input; rows t' >= To are dropped) and / or f32 time-major [To, N, Wo]
__global__ void __launch_bounds__(256)
act_kernel(const float* __restrict__ z, int N, int Rn, int To, int Wo, float clip, __half* __restrict__ y16, int y_rows, int y_row0,
           float* __restrict__ y32) {
  const int64_t total = (int64_t)N * To * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Wo);
    const int t = (int)((i / Wo) % To);
    const int n = (int)(i / ((int64_t)Wo * To));
    const float v = fminf(fmaxf(z[((size_t)n * Rn + t) * Wo + j], 0.0f), clip);
    if (y16) y16[((size_t)n * y_rows + y_row0 + t) * Wo + j] = __float2half_rn(v);
    if (y32) y32[((size_t)t * N + n) * Wo + j] = v;
  }
}

// g = gout * [0 < y < clip] in GEMM-row space (row n * Rn + t', zero for t' >= To): bf16 [N * Rn, Wo], bf16 transposed
// [Wo, N * Rn], f32.  gout element (n, t', j) sits at row t' * g_ts + n * g_ns + g_row0 (time-major: N, 1, 0; batch-major
// padded: 1, Tp, pt).
__global__ void __launch_bounds__(256)
act_bwd_kernel(const float* __restrict__ gout, int64_t g_ts, int64_t g_ns, int64_t g_row0, const __half* __restrict__ y16, int y_rows,
               int y_row0, int N, int Rn, int To, int Wo, float clip, __nv_bfloat16* __restrict__ g16, __nv_bfloat16* __restrict__ gT16,
               float* __restrict__ g32) {
  const int64_t M = (int64_t)N * Rn, total = M * Wo;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % Wo);
    const int64_t r = i / Wo;
    const int t = (int)(r % Rn), n = (int)(r / Rn);
    float v = 0.0f;
    if (t < To) {
      const float a = __half2float(y16[((size_t)n * y_rows + y_row0 + t) * Wo + j]);
      if (a > 0.0f && a < clip) v = gout[((size_t)t * g_ts + (size_t)n * g_ns + g_row0) * Wo + j];
    }
    if (g16) g16[i] = __float2bfloat16_rn(v);
    if (gT16) gT16[(size_t)j * M + r] = __float2bfloat16_rn(v);
    if (g32) g32[i] = v;
  }
}

// out bf16 [K, ldT]: out[m, r] = xp_flat[r * lda + m] — the overlapping view, transposed (K-major operand of the
// weight-gradient GEMM), through 32 x 32 shared-memory tiles
__global__ void __launch_bounds__(256)
unfold_t_kernel(const __half* __restrict__ xp, int64_t lda, int64_t rows, int K, __nv_bfloat16* __restrict__ out, int64_t ldT) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int64_t tiles_c = (K + 31) / 32, tiles_r = (rows + 31) / 32;
  for (int64_t tIdx = blockIdx.x; tIdx < tiles_c * tiles_r; tIdx += gridDim.x) {
    const int64_t r0 = (tIdx / tiles_c) * 32;
    const int c0 = (int)(tIdx % tiles_c) * 32;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t r = r0 + ty + 8 * j;
      const int c = c0 + tx;
      tile[ty + 8 * j][tx] = (r < rows && c < K) ? __half2float(xp[r * lda + c]) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = c0 + ty + 8 * j;
      const int64_t r = r0 + tx;
      if (c < K && r < rows) out[(size_t)c * ldT + r] = __float2bfloat16_rn(tile[tx][ty + 8 * j]);
    }
    __syncthreads();
  }
}

// dW[co, dkt, dkf, c] = sum over the output bins f' of dwt[(f', co), (dkt, f = sf * f' - pf + dkf, c)];  db[co] = sum_f' cs[(f', co)]
__global__ void __launch_bounds__(256)
toeplitz_grad_kernel(const float* __restrict__ dwt, int64_t ld, const float* __restrict__ cs, Geom g, int Co, float* __restrict__ dw,
                     float* __restrict__ db) {
  const int64_t total = (int64_t)Co * g.kt * g.kf * g.C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % g.C);
    const int dkf = (int)((i / g.C) % g.kf);
    const int dkt = (int)((i / ((int64_t)g.C * g.kf)) % g.kt);
    const int co = (int)(i / ((int64_t)g.C * g.kf * g.kt));
    float acc = 0.0f;
    for (int fo = 0; fo < g.Fo; ++fo) {
      const int f = g.sf * fo - g.pf + dkf;
      if (f >= 0 && f < g.F) acc += dwt[(size_t)(fo * Co + co) * ld + (size_t)dkt * g.W + f * g.C + c];
    }
    dw[i] = acc;
    if (i < Co) {
      float s = 0.0f;
      for (int fo = 0; fo < g.Fo; ++fo) s += cs[fo * Co + (int)i];
      db[i] = s;
    }
  }
}

}  // namespace

extern "C" int32_t asr_conv_plan_for(const asr_conv_geom* geom, asr_conv_plan* plan) {
  Geom g;
  ASR_CHECK_ARG(geom && plan, "asr_conv_plan_for: null argument");
  ASR_CHECK_ARG(make_geom(geom, &g), "asr_conv_plan_for: bad geometry (st * F * C must be a multiple of 8)");
  plan->t_out = g.To; plan->f_out = g.Fo; plan->rows = g.Rn; plan->t_padded = g.Tp; plan->k = g.K; plan->k_padded = g.Kp;
  return ASR_OK;
}

extern "C" int32_t asr_conv_pack(const float* x, const asr_conv_geom* geom, void* xp16, void* stream) {
  Geom g;
  ASR_CHECK_ARG(x && xp16 && geom && make_geom(geom, &g), "asr_conv_pack: bad argument");
  pack_kernel<<<grid_1d((int64_t)g.T * g.N * g.W), 256, 0, (cudaStream_t)stream>>>(x, g.T, g.N, g.W, reinterpret_cast<__half*>(xp16),
                                                                                   g.Tp, g.pt);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_conv_toeplitz(const float* w, const float* b, const asr_conv_geom* geom, int32_t c_out, void* wt16,
                                     int64_t ldw, void* w2_16, float* bias_t, void* stream) {
  Geom g;
  ASR_CHECK_ARG(w && b && wt16 && bias_t && geom && make_geom(geom, &g) && c_out >= 1, "asr_conv_toeplitz: bad argument");
  ASR_CHECK_ARG(ldw >= g.Kp && ldw % 8 == 0, "asr_conv_toeplitz: ldw >= padded K (multiple of 8)");
  toeplitz_kernel<<<grid_1d((int64_t)g.Fo * c_out * g.Kp), 256, 0, (cudaStream_t)stream>>>(
      w, b, g, c_out, reinterpret_cast<__half*>(wt16), ldw, reinterpret_cast<__nv_bfloat16*>(w2_16), bias_t);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_conv_act(const float* z, const asr_conv_geom* geom, int32_t c_out, float clip, void* y16, int32_t y_rows,
                                int32_t y_row0, float* y32_tm, void* stream) {
  Geom g;
  ASR_CHECK_ARG(z && geom && make_geom(geom, &g) && c_out >= 1 && clip > 0.0f && (y16 || y32_tm), "asr_conv_act: bad argument");
  ASR_CHECK_ARG(!y16 || (y_row0 >= 0 && y_rows >= y_row0 + g.To), "asr_conv_act: the fp16 destination holds fewer than row0 + T_out rows");
  act_kernel<<<grid_1d((int64_t)g.N * g.To * g.Fo * c_out), 256, 0, (cudaStream_t)stream>>>(
      z, g.N, g.Rn, g.To, g.Fo * c_out, clip, reinterpret_cast<__half*>(y16), y_rows, y_row0, y32_tm);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_conv_act_backward(const float* gout, int64_t g_t_stride, int64_t g_n_stride, int64_t g_row0, const void* y16,
                                         int32_t y_rows, int32_t y_row0, const asr_conv_geom* geom, int32_t c_out, float clip,
                                         void* g16, void* gT16, float* g32, void* stream) {
  Geom g;
  ASR_CHECK_ARG(gout && y16 && geom && make_geom(geom, &g) && c_out >= 1 && clip > 0.0f, "asr_conv_act_backward: bad argument");
  act_bwd_kernel<<<grid_1d((int64_t)g.N * g.Rn * g.Fo * c_out), 256, 0, (cudaStream_t)stream>>>(
      gout, g_t_stride, g_n_stride, g_row0, reinterpret_cast<const __half*>(y16), y_rows, y_row0, g.N, g.Rn, g.To, g.Fo * c_out, clip,
      reinterpret_cast<__nv_bfloat16*>(g16), reinterpret_cast<__nv_bfloat16*>(gT16), g32);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_conv_unfold_t(const void* xp16, const asr_conv_geom* geom, void* out16, int64_t ldT, void* stream) {
  Geom g;
  ASR_CHECK_ARG(xp16 && out16 && geom && make_geom(geom, &g), "asr_conv_unfold_t: bad argument");
  const int64_t rows = (int64_t)g.N * g.Rn;
  ASR_CHECK_ARG(ldT >= rows && ldT % 8 == 0, "asr_conv_unfold_t: ldT >= N * rows (multiple of 8)");
  const int64_t tiles = ((rows + 31) / 32) * ((g.K + 31) / 32);
  unfold_t_kernel<<<(int)(tiles < 148 * 16 ? tiles : 148 * 16), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const __half*>(xp16), (int64_t)g.st * g.W, rows, g.K, reinterpret_cast<__nv_bfloat16*>(out16), ldT);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_conv_toeplitz_grad(const float* dwt, int64_t ld, const float* colsum, const asr_conv_geom* geom, int32_t c_out,
                                          float* dw, float* db, void* stream) {
  Geom g;
  ASR_CHECK_ARG(dwt && colsum && dw && db && geom && make_geom(geom, &g) && c_out >= 1 && ld >= g.K, "asr_conv_toeplitz_grad: bad argument");
  toeplitz_grad_kernel<<<grid_1d((int64_t)c_out * g.kt * g.kf * g.C), 256, 0, (cudaStream_t)stream>>>(dwt, ld, colsum, g, c_out, dw, db);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
