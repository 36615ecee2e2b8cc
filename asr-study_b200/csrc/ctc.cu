// K6 / K7 — CTC loss + gradient and best-path decode for sm_100a.
//
// asr_ctc_loss_grad, label lengths <= 63 (2L+1 <= 128 lattice states) — three launches:
//   ctc_lse_kernel         a warp per frame over all T * N frames: log-sum-exp of the logits (softmax normaliser)
//   ctc_lattice_mw_kernel  one CTA per utterance: four warps run the alpha lattice forward in time, four the beta
//                          lattice backward, concurrently, in the log domain (lane = lattice state, log-sum-exp of
//                          the 3 predecessors, rows renormalised against fp64 offsets), rows parked in an
//                          L2-resident workspace
//   ctc_grad_mw_kernel     a warp per frame: reduce alpha+beta per class, write dloss/dlogits = softmax - occupancy
// longer labels: ctc_loss_grad_kernel<16>, one CTA per utterance running the same three phases in one launch.
// Replaces tf.nn.ctc_loss (core/ctc_utils.py:68-70): softmax inside, standard
// merge-repeated topology, zero gradient for t >= seq_len.
// asr_ctc_greedy replaces tf.nn.ctc_greedy_decoder (core/ctc_utils.py:42).
#include "common.cuh"
#include <math_constants.h>

namespace {

constexpr int CTC_THREADS = 256;
constexpr int CTC_WARPS = CTC_THREADS / 32;

__device__ __forceinline__ float lse2(float a, float b) {
  const float m = fmaxf(a, b);
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + log1pf(expf(fminf(a, b) - m));
}
// lattice recursion: the single alpha / beta warp is instruction-bound (4 states per lane x 3 exp + log per
// frame), so the recursion uses the SFU approximations (ex2/lg2, rel. error ~2^-21 — far inside the 5e-4
// gradient bar; rows are renormalised against fp64 offsets every frame so arguments stay O(10)).
__device__ __forceinline__ float lse3(float a, float b, float c) {
  const float m = fmaxf(a, fmaxf(b, c));
  if (m == -CUDART_INF_F) return -CUDART_INF_F;
  return m + __logf(__expf(a - m) + __expf(b - m) + __expf(c - m));
}

// per-utterance workspace, in floats: alpha[T][s_max], beta[T][s_max], lse[T] (f32, padded to an even count), then
// offA[T], offB[T] and log p(l|x) (f64)
__host__ __device__ __forceinline__ size_t ws_f32_part(int T, int s_max) {
  return ((size_t)2 * T * s_max + T + 1) & ~(size_t)1;
}
__host__ __device__ __forceinline__ size_t ws_stride(int T, int s_max) {
  return ws_f32_part(T, s_max) + 4 * (size_t)T + 2;
}

// NJ = states per lane (S <= 32*NJ)
template <int NJ>
__global__ void __launch_bounds__(CTC_THREADS)
ctc_loss_grad_kernel(const float* __restrict__ logits, int T, int N, int C,
                     const int* __restrict__ in_len, const int* __restrict__ labels,
                     const int* __restrict__ label_off, int s_max, int blank,
                     float grad_scale, float* __restrict__ loss, float* __restrict__ grad,
                     float* __restrict__ ws) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = min(max(in_len[n], 0), T);
  const int l0 = label_off[n], L = label_off[n + 1] - l0;
  const int S = 2 * L + 1;
  const float NEG = -CUDART_INF_F;

  // per-utterance workspace: alpha[T][s_max], beta[T][s_max], lse[T] (f32) then offA[T], offB[T] (f64).
  // alpha/beta rows are stored RELATIVE to a per-frame offset (row max = 0) that is accumulated in
  // fp64, so fp32 resolution applies to O(1) magnitudes instead of O(T) log-probabilities.
  const size_t fl_per = ws_f32_part(T, s_max);
  float* w_alpha = ws + (size_t)n * ws_stride(T, s_max);
  float* w_beta = w_alpha + (size_t)T * s_max;
  float* w_lse = w_beta + (size_t)T * s_max;
  double* w_offA = reinterpret_cast<double*>(w_alpha + fl_per);
  double* w_offB = w_offA + T;

  int* ext = reinterpret_cast<int*>(sm);            // [32*NJ]
  float* rowA = sm + 32 * NJ;                        // [32*NJ + 2] alpha exchange (2 pad in front)
  float* rowB = rowA + 32 * NJ + 2;                  // [32*NJ + 2] beta exchange (2 pad at end)
  float* acc = rowB + 32 * NJ + 2;                   // [CTC_WARPS][C]
  __shared__ double s_logp;

  for (int s = tid; s < 32 * NJ; s += CTC_THREADS)
    ext[s] = (s < S && (s & 1)) ? labels[l0 + (s >> 1)] : blank;

  // ---- phase 0: softmax normaliser per frame ------------------------------------
  for (int t = warp; t < len; t += CTC_WARPS) {
    const float* row = logits + ((size_t)t * N + n) * C;
    float m = NEG;
    for (int k = lane; k < C; k += 32) m = fmaxf(m, row[k]);
    m = asr::warp_max(m);
    float s = 0.0f;
    for (int k = lane; k < C; k += 32) s += expf(row[k] - m);
    s = asr::warp_sum(s);
    if (lane == 0) w_lse[t] = m + logf(s);
  }
  __syncthreads();

  if (len == 0 || S > 32 * NJ) {   // degenerate: no frames (or label too long for this build)
    if (tid == 0) loss[n] = CUDART_INF_F;
    for (int i = tid; i < T * C; i += CTC_THREADS) grad[((size_t)(i / C) * N + n) * C + (i % C)] = 0.0f;
    return;
  }

  // ---- phase 1: alpha (warp 0) and beta (warp 1) --------------------------------
  if (warp == 0) {
    bool skip[NJ];
    int lab[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int s = j * 32 + lane;
      lab[j] = ext[s];
      skip[j] = (s >= 2) && (s < S) && (lab[j] != blank) && (lab[j] != ext[s - 2]);
    }
    float* prev = rowA + 2;
    if (lane < 2) rowA[lane] = NEG;
    float a[NJ], x[NJ];
    {
      const float* row = logits + (size_t)n * C;
      const float z = w_lse[0];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int s = j * 32 + lane;
        a[j] = (s < 2 && s < S) ? row[lab[j]] - z : NEG;
      }
    }
    double offA = 0.0;
    {
      float m = NEG;
#pragma unroll
      for (int j = 0; j < NJ; ++j) m = fmaxf(m, a[j]);
      m = asr::warp_max(m);
      if (m > NEG) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) a[j] -= m;
        offA += (double)m;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j * 32 + lane < S) w_alpha[j * 32 + lane] = a[j];
      if (lane == 0) w_offA[0] = offA;
    }
    float z = 0.0f;
    if (len > 1) {
      const float* row = logits + ((size_t)1 * N + n) * C;
#pragma unroll
      for (int j = 0; j < NJ; ++j) x[j] = row[lab[j]];
      z = w_lse[1];
    }
    for (int t = 1; t < len; ++t) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) prev[j * 32 + lane] = a[j];
      __syncwarp();
      float xn[NJ], zn = 0.0f;          // next frame's operands are fetched a whole frame ahead
      if (t + 1 < len) {
        const float* row = logits + ((size_t)(t + 1) * N + n) * C;
#pragma unroll
        for (int j = 0; j < NJ; ++j) xn[j] = row[lab[j]];
        zn = w_lse[t + 1];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int s = j * 32 + lane;
        const float p1 = prev[s - 1];
        const float p2 = skip[j] ? prev[s - 2] : NEG;
        const float v = lse3(a[j], p1, p2) + (x[j] - z);
        a[j] = (s < S) ? v : NEG;
      }
      {
        float m = NEG;
#pragma unroll
        for (int j = 0; j < NJ; ++j) m = fmaxf(m, a[j]);
        m = asr::warp_max(m);
        if (m > NEG) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) a[j] -= m;
          offA += (double)m;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (j * 32 + lane < S) w_alpha[(size_t)t * s_max + j * 32 + lane] = a[j];
        if (lane == 0) w_offA[t] = offA;
      }
      __syncwarp();
      if (t + 1 < len) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[j] = xn[j];
        z = zn;
      }
    }
    // log p(l|x) = lse(alpha[len-1][S-1], alpha[len-1][S-2])
#pragma unroll
    for (int j = 0; j < NJ; ++j) prev[j * 32 + lane] = a[j];
    __syncwarp();
    if (lane == 0) {
      const float tail = (S > 1) ? lse2(prev[S - 1], prev[S - 2]) : prev[S - 1];
      s_logp = (tail > NEG) ? offA + (double)tail : -(double)CUDART_INF;
    }
  } else if (warp == 1) {
    bool skip[NJ];
    int lab[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int s = j * 32 + lane;
      lab[j] = ext[s];
      // transition s -> s+2 allowed
      skip[j] = (s + 2 < S) && (ext[s + 2] != blank) && (ext[s + 2] != lab[j]);
    }
    float* nxt = rowB;
    if (lane < 2) rowB[32 * NJ + lane] = NEG;
    float b[NJ], x[NJ];
    {
      const float* row = logits + ((size_t)(len - 1) * N + n) * C;
      const float z = w_lse[len - 1];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int s = j * 32 + lane;
        b[j] = (s < S && s >= S - 2) ? row[lab[j]] - z : NEG;
      }
    }
    double offB = 0.0;
    {
      float m = NEG;
#pragma unroll
      for (int j = 0; j < NJ; ++j) m = fmaxf(m, b[j]);
      m = asr::warp_max(m);
      if (m > NEG) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) b[j] -= m;
        offB += (double)m;
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (j * 32 + lane < S) w_beta[(size_t)(len - 1) * s_max + j * 32 + lane] = b[j];
      if (lane == 0) w_offB[len - 1] = offB;
    }
    float z = 0.0f;
    if (len > 1) {
      const float* row = logits + ((size_t)(len - 2) * N + n) * C;
#pragma unroll
      for (int j = 0; j < NJ; ++j) x[j] = row[lab[j]];
      z = w_lse[len - 2];
    }
    for (int t = len - 2; t >= 0; --t) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) nxt[j * 32 + lane] = b[j];
      __syncwarp();
      float xn[NJ], zn = 0.0f;
      if (t > 0) {
        const float* row = logits + ((size_t)(t - 1) * N + n) * C;
#pragma unroll
        for (int j = 0; j < NJ; ++j) xn[j] = row[lab[j]];
        zn = w_lse[t - 1];
      }
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int s = j * 32 + lane;
        const float p1 = nxt[s + 1];
        const float p2 = skip[j] ? nxt[s + 2] : NEG;
        const float v = lse3(b[j], p1, p2) + (x[j] - z);
        b[j] = (s < S) ? v : NEG;
      }
      {
        float m = NEG;
#pragma unroll
        for (int j = 0; j < NJ; ++j) m = fmaxf(m, b[j]);
        m = asr::warp_max(m);
        if (m > NEG) {
#pragma unroll
          for (int j = 0; j < NJ; ++j) b[j] -= m;
          offB += (double)m;
        }
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (j * 32 + lane < S) w_beta[(size_t)t * s_max + j * 32 + lane] = b[j];
        if (lane == 0) w_offB[t] = offB;
      }
      __syncwarp();
      if (t > 0) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) x[j] = xn[j];
        z = zn;
      }
    }
  }
  __syncthreads();

  const double logp = s_logp;
  if (tid == 0) loss[n] = (float)(-logp);
  const bool feasible = logp > -(double)CUDART_INF;

  // ---- phase 2: gradient ----------------------------------------------------------
  float* my = acc + warp * C;
  for (int t = warp; t < T; t += CTC_WARPS) {
    float* g = grad + ((size_t)t * N + n) * C;
    if (t >= len || !feasible) {
      for (int k = lane; k < C; k += 32) g[k] = 0.0f;
      continue;
    }
    const float* row = logits + ((size_t)t * N + n) * C;
    const float z = w_lse[t];
    const float kf = (float)(w_offA[t] + w_offB[t] - logp);   // frame constant, O(1) after the fp64 cancellation
    for (int k = lane; k < C; k += 32) my[k] = 0.0f;
    float v[NJ], m = NEG;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int s = j * 32 + lane;
      v[j] = (s < S) ? w_alpha[(size_t)t * s_max + s] + w_beta[(size_t)t * s_max + s] : NEG;
      m = fmaxf(m, v[j]);
    }
    m = asr::warp_max(m);
    __syncwarp();
    float bsum = 0.0f;
    if (m > NEG) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int s = j * 32 + lane;
        if (s < S) {
          const float e = expf(v[j] - m);
          if (s & 1) atomicAdd(my + ext[s], e);   // label states (odd s)
          else bsum += e;                          // blank states (even s)
        }
      }
    }
    bsum = asr::warp_sum(bsum);
    if (lane == 0) atomicAdd(my + blank, bsum);
    __syncwarp();
    for (int k = lane; k < C; k += 32) {
      const float lp = row[k] - z;
      float occ = 0.0f;
      if (my[k] > 0.0f) occ = my[k] * expf(m + kf - lp);
      g[k] = grad_scale * (expf(lp) - occ);
    }
    __syncwarp();
  }
}

// ---- multi-warp lattice variant (S <= 128): warps 0-3 walk alpha, warps 4-7 walk beta, one lattice state per
// lane; the previous frame's row lives in double-buffered shared memory, one 128-thread named barrier per frame.
// Rows are stored un-normalised for ONE frame (the row maximum is published next to the row and subtracted by
// the readers), and the running sum of maxima is the fp64 offset: alpha_t(s) = stored + offA[t].
constexpr int PF = 4;         // emission prefetch distance of the multi-warp lattice (frames)
// warp maximum in ONE instruction (REDUX.MAX over an order-preserving integer image of the floats) instead of a
// five-stage shuffle tree: the row maximum of the lattice sits on the 999-step dependent chain
__device__ __forceinline__ int f2ord(float x) {
  const int i = __float_as_int(x);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }
__device__ __forceinline__ float warp_max_redux(float v) { return ord2f(__reduce_max_sync(0xffffffffu, f2ord(v))); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_lattice_mw_kernel(const float* __restrict__ logits, int T, int N, int C, const int* __restrict__ in_len,
                      const int* __restrict__ labels, const int* __restrict__ label_off, int s_max, int blank,
                      float* __restrict__ loss, float* __restrict__ ws) {
  extern __shared__ float sm[];
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = min(max(in_len[n], 0), T);
  const int l0 = label_off[n], L = label_off[n + 1] - l0;
  const int S = 2 * L + 1;
  const float NEG = -CUDART_INF_F;
  const size_t fl_per = ws_f32_part(T, s_max);
  float* w_alpha = ws + (size_t)n * ws_stride(T, s_max);
  float* w_beta = w_alpha + (size_t)T * s_max;
  float* w_lse = w_beta + (size_t)T * s_max;
  double* w_offA = reinterpret_cast<double*>(w_alpha + fl_per);
  double* w_offB = w_offA + T;

  int* ext = reinterpret_cast<int*>(sm);              // [132]
  float* rowA = sm + 132;                             // [2][132]: 2 pads in front
  float* rowB = rowA + 2 * 132;                       // [2][132]: states at [0,128), 2 pads at [128,130)
  float* mxA = rowB + 2 * 132;                        // [2][4]
  float* mxB = mxA + 8;                               // [2][4]
  __shared__ double s_logp;

  for (int s = tid; s < 132; s += CTC_THREADS) ext[s] = (s < S && (s & 1)) ? labels[l0 + (s >> 1)] : blank;
  if (tid < 2) { rowA[tid] = NEG; rowA[132 + tid] = NEG; rowB[128 + tid] = NEG; rowB[132 + 128 + tid] = NEG; }
  __syncthreads();
  if (len == 0) {                                     // no frames: infinite loss, and the gradient kernel writes zeros
    if (tid == 0) { loss[n] = CUDART_INF_F; w_offB[T] = -(double)CUDART_INF; }
    return;
  }

  if (warp < 4) {                                     // ---------------- alpha ----------------
    const int s = warp * 32 + lane;
    const int lab = ext[s];
    const bool skip = (s >= 2) && (s < S) && (lab != blank) && (lab != ext[s - 2]);
    double off = 0.0;
    float a = (s < 2 && s < S) ? logits[(size_t)n * C + lab] - w_lse[0] : NEG;
    {
      if (s < S) w_alpha[s] = a;
      (rowA + 2)[s] = a;
      const float wm = warp_max_redux(a);
      if (lane == 0) mxA[warp] = wm;
      if (tid == 0) w_offA[0] = 0.0;
    }
    // the emission x_t(lab) - lse_t of the next PF frames is prefetched into registers: one frame of lattice work
    // (~400 cycles) does not cover an L2 round trip, and the load sits on the 999-step dependent chain otherwise
    float xq[PF], zq[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int tt = 1 + j;
      xq[j] = (tt < len) ? logits[((size_t)tt * N + n) * C + lab] : 0.0f;
      zq[j] = (tt < len) ? w_lse[tt] : 0.0f;
    }
    named_bar_sync(1, 128);
    for (int t = 1; t < len; ++t) {
      const float* prev = rowA + ((t - 1) & 1) * 132 + 2;
      const float* pm = mxA + ((t - 1) & 1) * 4;
      const float M = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
      const float x = xq[0], z = zq[0];
#pragma unroll
      for (int j = 0; j + 1 < PF; ++j) { xq[j] = xq[j + 1]; zq[j] = zq[j + 1]; }
      xq[PF - 1] = zq[PF - 1] = 0.0f;
      if (t + PF < len) { xq[PF - 1] = logits[((size_t)(t + PF) * N + n) * C + lab]; zq[PF - 1] = w_lse[t + PF]; }
      const float p0 = prev[s], p1 = prev[s - 1], p2 = skip ? prev[s - 2] : NEG;
      float v = lse3(p0, p1, p2);
      v = (v > NEG && M > NEG) ? v - M + (x - z) : NEG;
      a = (s < S) ? v : NEG;
      off += (M > NEG) ? (double)M : 0.0;
      if (s < S) w_alpha[(size_t)t * s_max + s] = a;
      (rowA + (t & 1) * 132 + 2)[s] = a;
      const float wm = warp_max_redux(a);
      if (lane == 0) mxA[(t & 1) * 4 + warp] = wm;
      if (tid == 0) w_offA[t] = off;
      named_bar_sync(1, 128);
    }
    if (tid == 0) {
      const float* last = rowA + ((len - 1) & 1) * 132 + 2;
      const float tail = (S > 1) ? lse2(last[S - 1], last[S - 2]) : last[S - 1];
      s_logp = (tail > NEG) ? off + (double)tail : -(double)CUDART_INF;
    }
  } else {                                            // ---------------- beta ----------------
    const int w4 = warp - 4;
    const int s = w4 * 32 + lane;
    const int lab = ext[s];
    const bool skip = (s + 2 < S) && (ext[s + 2] != blank) && (ext[s + 2] != lab);
    double off = 0.0;
    float b = (s < S && s >= S - 2) ? logits[((size_t)(len - 1) * N + n) * C + lab] - w_lse[len - 1] : NEG;
    {
      if (s < S) w_beta[(size_t)(len - 1) * s_max + s] = b;
      (rowB + ((len - 1) & 1) * 132)[s] = b;
      const float wm = warp_max_redux(b);
      if (lane == 0) mxB[((len - 1) & 1) * 4 + w4] = wm;
      if (tid == 128) w_offB[len - 1] = 0.0;
    }
    float xq[PF], zq[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int tt = len - 2 - j;
      xq[j] = (tt >= 0) ? logits[((size_t)tt * N + n) * C + lab] : 0.0f;
      zq[j] = (tt >= 0) ? w_lse[tt] : 0.0f;
    }
    named_bar_sync(2, 128);
    for (int t = len - 2; t >= 0; --t) {
      const float* nxt = rowB + ((t + 1) & 1) * 132;
      const float* pm = mxB + ((t + 1) & 1) * 4;
      const float M = fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3]));
      const float x = xq[0], z = zq[0];
#pragma unroll
      for (int j = 0; j + 1 < PF; ++j) { xq[j] = xq[j + 1]; zq[j] = zq[j + 1]; }
      xq[PF - 1] = zq[PF - 1] = 0.0f;
      if (t - PF >= 0) { xq[PF - 1] = logits[((size_t)(t - PF) * N + n) * C + lab]; zq[PF - 1] = w_lse[t - PF]; }
      const float p0 = nxt[s], p1 = nxt[s + 1], p2 = skip ? nxt[s + 2] : NEG;
      float v = lse3(p0, p1, p2);
      v = (v > NEG && M > NEG) ? v - M + (x - z) : NEG;
      b = (s < S) ? v : NEG;
      off += (M > NEG) ? (double)M : 0.0;
      if (s < S) w_beta[(size_t)t * s_max + s] = b;
      (rowB + (t & 1) * 132)[s] = b;
      const float wm = warp_max_redux(b);
      if (lane == 0) mxB[(t & 1) * 4 + w4] = wm;
      if (tid == 128) w_offB[t] = off;
      named_bar_sync(2, 128);
    }
  }
  __syncthreads();

  if (tid == 0) {
    loss[n] = (float)(-s_logp);
    w_offB[T] = s_logp;
  }
}

// ---- K6 at label lengths <= 63 runs as three launches: the two embarrassingly parallel phases (the per-frame softmax
// normaliser before the lattice, softmax - occupancy after it) get the whole GPU — a warp per frame over T * N frames —
// instead of the eight warps of the utterance's own CTA (which spent 19 % + 37 % of the fused kernel's time on them), and
// only the 999-step dependent chain stays on one CTA per utterance.
__global__ void __launch_bounds__(CTC_THREADS)
ctc_lse_kernel(const float* __restrict__ logits, int T, int N, int C, const int* __restrict__ in_len, int s_max,
               float* __restrict__ ws) {
  const int lane = threadIdx.x & 31;
  const long long f = (long long)blockIdx.x * CTC_WARPS + (threadIdx.x >> 5);      // frame (t, n) of the time-major logits
  if (f >= (long long)T * N) return;
  const int t = (int)(f / N), n = (int)(f % N);
  if (t >= min(max(in_len[n], 0), T)) return;
  const float* row = logits + (size_t)f * C;
  float m = -CUDART_INF_F;
  for (int k = lane; k < C; k += 32) m = fmaxf(m, row[k]);
  m = asr::warp_max(m);
  float sx = 0.0f;
  for (int k = lane; k < C; k += 32) sx += expf(row[k] - m);
  sx = asr::warp_sum(sx);
  if (lane == 0) (ws + (size_t)n * ws_stride(T, s_max) + (size_t)2 * T * s_max)[t] = m + logf(sx);
}

__global__ void __launch_bounds__(CTC_THREADS)
ctc_grad_mw_kernel(const float* __restrict__ logits, int T, int N, int C, const int* __restrict__ in_len,
                   const int* __restrict__ labels, const int* __restrict__ label_off, int s_max, int blank,
                   float grad_scale, float* __restrict__ grad, const float* __restrict__ ws) {
  extern __shared__ float sm[];                         // [CTC_WARPS][C] per-class occupancy of the warp's frame
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long f = (long long)blockIdx.x * CTC_WARPS + warp;
  if (f >= (long long)T * N) return;
  const int t = (int)(f / N), n = (int)(f % N);
  const float NEG = -CUDART_INF_F;
  const int len = min(max(in_len[n], 0), T);
  const float* w_alpha = ws + (size_t)n * ws_stride(T, s_max);
  const float* w_beta = w_alpha + (size_t)T * s_max;
  const float* w_lse = w_beta + (size_t)T * s_max;
  const double* w_offA = reinterpret_cast<const double*>(w_alpha + ws_f32_part(T, s_max));
  const double* w_offB = w_offA + T;
  const double logp = w_offB[T];
  float* g = grad + (size_t)f * C;
  if (t >= len || !(logp > -(double)CUDART_INF)) {
    for (int k = lane; k < C; k += 32) g[k] = 0.0f;
    return;
  }
  const int l0 = label_off[n], S = 2 * (label_off[n + 1] - l0) + 1;
  const float* row = logits + (size_t)f * C;
  const float z = w_lse[t];
  const float kf = (float)(w_offA[t] + w_offB[t] - logp);
  float* my = sm + warp * C;
  for (int k = lane; k < C; k += 32) my[k] = 0.0f;
  float v[4], m = NEG;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int s = j * 32 + lane;
    v[j] = (s < S) ? w_alpha[(size_t)t * s_max + s] + w_beta[(size_t)t * s_max + s] : NEG;
    m = fmaxf(m, v[j]);
  }
  m = asr::warp_max(m);
  __syncwarp();
  float bsum = 0.0f;
  if (m > NEG) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int s = j * 32 + lane;
      if (s < S) {
        const float e = expf(v[j] - m);
        if (s & 1) atomicAdd(my + labels[l0 + (s >> 1)], e);
        else bsum += e;
      }
    }
  }
  bsum = asr::warp_sum(bsum);
  if (lane == 0) atomicAdd(my + blank, bsum);
  __syncwarp();
  for (int k = lane; k < C; k += 32) {
    const float lp = row[k] - z;
    float occ = 0.0f;
    if (my[k] > 0.0f) occ = my[k] * expf(m + kf - lp);
    g[k] = grad_scale * (expf(lp) - occ);
  }
}

// ---- best path ----------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ctc_greedy_kernel(const float* __restrict__ logits, int T, int N, int C,
                  const int* __restrict__ in_len, int blank, int merge,
                  int* __restrict__ out_labels, int* __restrict__ out_len) {
  __shared__ int s_arg[256];
  __shared__ int s_wsum[8];
  __shared__ int s_base, s_carry;
  const int n = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int len = min(max(in_len[n], 0), T);
  int* outp = out_labels + (size_t)n * T;
  if (tid == 0) { s_base = 0; s_carry = -1; }
  __syncthreads();
  for (int c0 = 0; c0 < len; c0 += 256) {
    const int t = c0 + tid;
    int arg = -1;
    if (t < len) {
      const float* row = logits + ((size_t)t * N + n) * C;
      float best = row[0];
      arg = 0;
      for (int k = 1; k < C; ++k) {
        const float v = row[k];
        if (v > best) { best = v; arg = k; }   // first maximum wins
      }
    }
    s_arg[tid] = arg;
    __syncthreads();
    const int prev = (tid > 0) ? s_arg[tid - 1] : s_carry;
    const int keep = (t < len) && (arg != blank) && !(merge && arg == prev);
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    const int wpre = __popc(bal & ((1u << lane) - 1));
    if (lane == 0) s_wsum[warp] = __popc(bal);
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int w = 0; w < 8; ++w) {
      if (w < warp) wbase += s_wsum[w];
      tot += s_wsum[w];
    }
    const int base = s_base;
    if (keep) outp[base + wbase + wpre] = arg;
    __syncthreads();
    if (tid == 0) { s_base = base + tot; s_carry = s_arg[min(255, len - 1 - c0)]; }
    __syncthreads();
  }
  const int total = s_base;
  for (int i = total + tid; i < T; i += 256) outp[i] = -1;
  if (tid == 0) out_len[n] = total;
}

}  // namespace

extern "C" size_t asr_ctc_workspace_bytes(int32_t T, int32_t N, int32_t max_label_len) {
  if (T <= 0 || N <= 0 || max_label_len < 0) return 0;
  const size_t s_max = 2 * (size_t)max_label_len + 1;
  return (size_t)N * ws_stride(T, (int)s_max) * sizeof(float);
}

extern "C" int32_t asr_ctc_loss_grad(const float* logits, int32_t T, int32_t N, int32_t C, const int32_t* in_len,
                                     const int32_t* labels, const int32_t* label_off, int32_t max_label_len,
                                     int32_t blank, float grad_scale, float* loss, float* grad, void* ws,
                                     void* stream) {
  ASR_CHECK_ARG(logits && in_len && labels && label_off && loss && grad && ws, "asr_ctc_loss_grad: null argument");
  ASR_CHECK_ARG(T >= 1 && N >= 1 && C >= 2 && blank >= 0 && blank < C, "asr_ctc_loss_grad: bad shape T=%d N=%d C=%d", T, N, C);
  ASR_CHECK_ARG(max_label_len >= 0 && max_label_len <= 255, "asr_ctc_loss_grad: max_label_len %d > 255", max_label_len);
  const int s_max = 2 * max_label_len + 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (s_max <= 128) {
    const int frame_ctas = (int)(((long long)T * N + CTC_WARPS - 1) / CTC_WARPS);
    ctc_lse_kernel<<<frame_ctas, CTC_THREADS, 0, st>>>(logits, T, N, C, in_len, s_max, (float*)ws);
    ASR_LAUNCH_CHECK();
    ctc_lattice_mw_kernel<<<N, CTC_THREADS, (size_t)(132 + 4 * 132 + 16) * sizeof(float), st>>>(
        logits, T, N, C, in_len, labels, label_off, s_max, blank, loss, (float*)ws);
    ASR_LAUNCH_CHECK();
    ctc_grad_mw_kernel<<<frame_ctas, CTC_THREADS, (size_t)CTC_WARPS * C * sizeof(float), st>>>(
        logits, T, N, C, in_len, labels, label_off, s_max, blank, grad_scale, grad, (const float*)ws);
  } else {
    const size_t smem = (size_t)(3 * 32 * 16 + 4 + CTC_WARPS * C) * sizeof(float);
    ctc_loss_grad_kernel<16><<<N, CTC_THREADS, smem, st>>>(logits, T, N, C, in_len, labels, label_off, s_max, blank,
                                                            grad_scale, loss, grad, (float*)ws);
  }
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

// ---- K10: label error rate ---------------------------------------------------------------------
// core/metrics.py:4-8: tf.edit_distance(hyp, truth, normalize=True) per utterance (Levenshtein distance over label
// ids, divided by the truth length; an empty truth gives 0 for an empty hypothesis and +inf otherwise, as TF does).
// One warp per utterance, anti-diagonal wavefront: cell (i, j) of the DP table depends on (i-1, j), (i, j-1) and
// (i-1, j-1), so all cells of one anti-diagonal are independent; lane l owns the truth positions j = l + 1, l + 33, ...
// and three rotating diagonals live in shared memory.  A few dozen labels per utterance: latency, not bandwidth.
__global__ void __launch_bounds__(32)
edit_distance_kernel(const int* __restrict__ hyp, int hyp_stride, const int* __restrict__ hyp_len,
                     const int* __restrict__ truth, const int* __restrict__ truth_off, int normalize,
                     float* __restrict__ out) {
  extern __shared__ int sd[];                           // 3 diagonals of (n + 1) cells, indexed by j
  const int u = blockIdx.x, lane = threadIdx.x;
  const int* a = hyp + (size_t)u * hyp_stride;          // hypothesis, length m
  int m = 0;
  if (hyp_len) m = max(hyp_len[u], 0);
  else {                                                // -1 padded row: count the leading non-negative labels
    for (int base = 0; base < hyp_stride; base += 32) {
      const int i = base + lane;
      const unsigned neg = __ballot_sync(0xffffffffu, i >= hyp_stride || a[i] < 0);
      if (neg) { m = base + __ffs(neg) - 1; break; }
      m = min(base + 32, hyp_stride);
    }
  }
  m = min(m, hyp_stride);
  const int t0 = truth_off[u], n = truth_off[u + 1] - t0;
  const int* b = truth + t0;
  float res;
  if (n == 0) {
    res = (m == 0) ? 0.0f : (normalize ? CUDART_INF_F : (float)m);
  } else if (m == 0) {
    res = normalize ? 1.0f : (float)n;
  } else {
    int* d0 = sd;                                       // diagonal k - 2
    int* d1 = sd + (n + 1);                             // diagonal k - 1
    int* d2 = sd + 2 * (n + 1);                         // diagonal k   (cells (i, j) with i + j = k)
    // D(i, 0) = i, D(0, j) = j; diagonal k holds D(k - j, j) at index j
    for (int j = lane; j <= n; j += 32) { d0[j] = 0; d1[j] = 0; }
    if (lane == 0) { d0[0] = 0; d1[0] = 1; d1[1] = 1; }   // k = 0: D(0,0) = 0;  k = 1: D(1,0) = 1, D(0,1) = 1
    __syncwarp();
    for (int k = 2; k <= m + n; ++k) {
      const int jlo = max(0, k - m), jhi = min(n, k);
      for (int j = jlo + lane; j <= jhi; j += 32) {
        const int i = k - j;
        int v;
        if (j == 0) v = i;
        else if (i == 0) v = j;
        else {
          const int sub = d0[j - 1] + (a[i - 1] != b[j - 1]);      // D(i-1, j-1) on diagonal k-2
          const int del = d1[j] + 1;                               // D(i-1, j)   on diagonal k-1
          const int ins = d1[j - 1] + 1;                           // D(i, j-1)   on diagonal k-1
          v = min(sub, min(del, ins));
        }
        d2[j] = v;
      }
      __syncwarp();
      int* tmp = d0; d0 = d1; d1 = d2; d2 = tmp;
    }
    const int dist = d1[n];                             // diagonal m + n holds D(m, n) at index n
    res = normalize ? (float)dist / (float)n : (float)dist;
  }
  if (lane == 0) out[u] = res;
}

extern "C" int32_t asr_edit_distance(const int32_t* hyp, int32_t N, int32_t hyp_stride, const int32_t* hyp_len,
                                     const int32_t* truth, const int32_t* truth_off, int32_t max_truth_len,
                                     int32_t normalize, float* out, void* stream) {
  ASR_CHECK_ARG(hyp && truth && truth_off && out, "asr_edit_distance: null argument");
  ASR_CHECK_ARG(N >= 1 && hyp_stride >= 1 && max_truth_len >= 0, "asr_edit_distance: bad shape");
  const size_t smem = (size_t)3 * (max_truth_len + 1) * sizeof(int);
  ASR_CHECK_ARG(smem <= 200 * 1024, "asr_edit_distance: truth of %d labels needs %zu B of shared memory", max_truth_len, smem);
  if (smem > 48 * 1024)
    ASR_CUDA(cudaFuncSetAttribute(edit_distance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  edit_distance_kernel<<<N, 32, smem, (cudaStream_t)stream>>>(hyp, hyp_stride, hyp_len, truth, truth_off, normalize, out);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_ctc_greedy(const float* logits, int32_t T, int32_t N, int32_t C, const int32_t* in_len,
                                  int32_t blank, int32_t merge_repeated, int32_t* out_labels, int32_t* out_len,
                                  void* stream) {
  ASR_CHECK_ARG(logits && in_len && out_labels && out_len, "asr_ctc_greedy: null argument");
  ASR_CHECK_ARG(T >= 1 && N >= 1 && C >= 1, "asr_ctc_greedy: bad shape");
  ctc_greedy_kernel<<<N, 256, 0, (cudaStream_t)stream>>>(logits, T, N, C, in_len, blank, merge_repeated, out_labels,
                                                          out_len);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
