// K3/K4 (general-cell engine) — LSTM.step with the brsmv1 switches, forward and BPTT, fp32, sm_100a.
//
// Reference: core/layers.py:432-469 with layer normalisation (core/layers.py:407-430,
// core/layers_utils.py:16-19: moments over the feature axis, sqrt(var + eps)), multiplicative
// integration (core/layers.py:441-443) and zoneout (core/layers.py:457-467, layers_utils.py:34-42:
// one keep mask per time step shared by the batch).  The default (no-switch) step runs on the
// tensor-core engines; this engine takes over when any switch is on, because LN needs whole-row
// statistics of Uh [4H], Wx [4H] and c [H] at every time step.
//
// Layout: one CTA = NS samples of one direction, ALL hidden units, so every LN reduction is CTA-local
// and there is no inter-CTA exchange at all.  U (fp32, [H,4H]) is streamed from L2 every step (it is
// shared by all CTAs of a direction and stays L2-resident).  grid = (ceil(N/NS), 2).
//
// Phases of a forward step (threads own gate COLUMNS in the product phases and UNITS in the cell phase):
//   A  uh_raw[n][c] = sum_k (h_{t-1} * B_U)[n][k] * U[k][c]                       (column owner)
//   B  LN statistics of uh_raw and of zx_t (two-pass mean / variance), MI, bias, gate activations
//   C  c_new, zoneout(c), LN(c), h_new, zoneout(h); outputs                       (unit owner)
// The backward step mirrors it: C' (unit owner: dh -> dz), B' (column owner: MI / LN backward ->
// dWx, dUh + parameter gradients), A' dh_rec[n][k] = sum_c dUh[n][c] * U[k][c] (warp per k).
#include "common.cuh"

namespace lstmcell {

constexpr int NS = 4;              // samples per CTA
constexpr int THREADS = 256;
constexpr int JMAX = 16;           // column slots per thread (template NJ <= JMAX): 4H <= NJ * THREADS  (H <= 1024)
constexpr int MMAX = 4;            // unit slots per thread   (template NM <= MMAX): H  <= NM * THREADS

// sum over the block of v[0..n) (n <= 16); every thread gets the totals.  Two barriers.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = asr::warp_sum(v[i]);
  __syncthreads();                 // buf may still be read from the previous reduction
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) buf[warp * NV + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < THREADS / 32; ++w) s += buf[w * NV + i];
    v[i] = s;
  }
}

// per-sample mean and 1/sqrt(var + eps) of val[j][n] over the thread-distributed columns (two-pass)
template <int NJ>
__device__ __forceinline__ void row_stats(const float (&val)[NJ][NS], int nj, int width, float eps, float* buf,
                                          float (&mean)[NS], float (&rstd)[NS]) {
  float s[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) s[n] = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (j < nj && (int)threadIdx.x + j * THREADS < width)
#pragma unroll
      for (int n = 0; n < NS; ++n) s[n] += val[j][n];
  block_sum<NS>(s, buf);
#pragma unroll
  for (int n = 0; n < NS; ++n) mean[n] = s[n] / (float)width;
#pragma unroll
  for (int n = 0; n < NS; ++n) s[n] = 0.0f;
#pragma unroll
  for (int j = 0; j < NJ; ++j)
    if (j < nj && (int)threadIdx.x + j * THREADS < width)
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float d = val[j][n] - mean[n];
        s[n] = fmaf(d, d, s[n]);
      }
  block_sum<NS>(s, buf);
#pragma unroll
  for (int n = 0; n < NS; ++n) rstd[n] = rsqrtf(s[n] / (float)width + eps);
}
template <int NM>
__device__ __forceinline__ void unit_stats(const float (&val)[NM][NS], int nm, int width, float eps, float* buf,
                                           float (&mean)[NS], float (&rstd)[NS]) {
  float s[NS];
#pragma unroll
  for (int n = 0; n < NS; ++n) s[n] = 0.0f;
#pragma unroll
  for (int m = 0; m < NM; ++m)
    if (m < nm && (int)threadIdx.x + m * THREADS < width)
#pragma unroll
      for (int n = 0; n < NS; ++n) s[n] += val[m][n];
  block_sum<NS>(s, buf);
#pragma unroll
  for (int n = 0; n < NS; ++n) mean[n] = s[n] / (float)width;
#pragma unroll
  for (int n = 0; n < NS; ++n) s[n] = 0.0f;
#pragma unroll
  for (int m = 0; m < NM; ++m)
    if (m < nm && (int)threadIdx.x + m * THREADS < width)
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float d = val[m][n] - mean[n];
        s[n] = fmaf(d, d, s[n]);
      }
  block_sum<NS>(s, buf);
#pragma unroll
  for (int n = 0; n < NS; ++n) rstd[n] = rsqrtf(s[n] / (float)width + eps);
}

__device__ __forceinline__ float zone_coeff(float level, const float* mask, int dir, int T, int H, int t, int u) {
  if (!(level > 0.0f && level < 1.0f)) return 1.0f;
  return mask ? mask[((size_t)dir * T + t) * H + u] : 1.0f - level;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int NJ, int NM>
__global__ void __launch_bounds__(THREADS, 1)
cell_fwd_kernel(asr_lstm_fwd_args a, asr_lstm_variant v, float* __restrict__ uh_raw_out) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, N = a.N, H = a.H, H4 = 4 * a.H;
  const int tid = threadIdx.x, dir = blockIdx.y, n0 = blockIdx.x * NS;
  const int nj = (H4 + THREADS - 1) / THREADS, nm = (H + THREADS - 1) / THREADS;
  float* sHm = smem;                       // [H][NS]   h_{t-1} * B_U
  float* sG = sHm + (size_t)H * NS;        // [4H][NS]  activated gates of this step
  float* sRed = sG + (size_t)H4 * NS;      // [8][16]
  const bool ln = v.ln_gain_uh != nullptr, mi = v.mi_alpha != nullptr;
  const float* Ug = a.U + (size_t)dir * H * H4;

  // column-owner constants
  float bias[NJ], al[NJ], b1[NJ], b2[NJ], gu[NJ], bu[NJ], gw[NJ], bw[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = tid + j * THREADS;
    const bool okc = j < nj && c < H4;
    const size_t o = (size_t)dir * H4 + c;
    bias[j] = okc ? a.bias[o] : 0.0f;
    al[j] = (okc && mi) ? v.mi_alpha[o] : 0.0f;
    b1[j] = (okc && mi) ? v.mi_beta1[o] : 1.0f;
    b2[j] = (okc && mi) ? v.mi_beta2[o] : 1.0f;
    gu[j] = (okc && ln) ? v.ln_gain_uh[o] : 1.0f;
    bu[j] = (okc && ln) ? v.ln_bias_uh[o] : 0.0f;
    gw[j] = (okc && ln) ? v.ln_gain_wx[o] : 1.0f;
    bw[j] = (okc && ln) ? v.ln_bias_wx[o] : 0.0f;
  }
  // unit-owner state
  float c_state[NM][NS], h_state[NM][NS], mu[NM][NS], gc[NM], bc[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const int u = tid + m * THREADS;
    const bool oku = m < nm && u < H;
    gc[m] = (oku && ln) ? v.ln_gain_c[(size_t)dir * H + u] : 1.0f;
    bc[m] = (oku && ln) ? v.ln_bias_c[(size_t)dir * H + u] : 0.0f;
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      c_state[m][n] = 0.0f;
      h_state[m][n] = 0.0f;
      mu[m][n] = (oku && a.mask_u && n0 + n < N) ? a.mask_u[((size_t)dir * N + n0 + n) * H + u] : 1.0f;
    }
  }
  for (int i = tid; i < H * NS; i += THREADS) sHm[i] = 0.0f;
  __syncthreads();
  const size_t R = (size_t)T * N;

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    // ---- A: uh_raw = (h * B_U) . U ------------------------------------------------------------
    float uh[NJ][NS];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int n = 0; n < NS; ++n) uh[j][n] = 0.0f;
    if (s > 0) {
#pragma unroll 2
      for (int k = 0; k < H; ++k) {
        const float4 hv = *reinterpret_cast<const float4*>(sHm + (size_t)k * NS);
        const float* ur = Ug + (size_t)k * H4 + tid;
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (j < nj && tid + j * THREADS < H4) {
            const float w = __ldg(ur + j * THREADS);
            uh[j][0] = fmaf(hv.x, w, uh[j][0]);
            uh[j][1] = fmaf(hv.y, w, uh[j][1]);
            uh[j][2] = fmaf(hv.z, w, uh[j][2]);
            uh[j][3] = fmaf(hv.w, w, uh[j][3]);
          }
      }
    }
    float wx[NJ][NS];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = tid + j * THREADS;
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const bool ok = j < nj && c < H4 && n0 + n < N;
        wx[j][n] = ok ? __ldg(a.zx + (((size_t)t * N + n0 + n) * 2 + dir) * H4 + c) : 0.0f;
        if (ok && uh_raw_out) uh_raw_out[(((size_t)t * N + n0 + n) * 2 + dir) * H4 + c] = uh[j][n];
      }
    }
    // ---- B: LN(Uh), LN(Wx), MI, activations --------------------------------------------------------
    if (ln) {
      float mean[NS], rstd[NS];
      row_stats(uh, nj, H4, v.ln_eps, sRed, mean, rstd);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) uh[j][n] = (uh[j][n] - mean[n]) * rstd[n] * gu[j] + bu[j];
      row_stats(wx, nj, H4, v.ln_eps, sRed, mean, rstd);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) wx[j][n] = (wx[j][n] - mean[n]) * rstd[n] * gw[j] + bw[j];
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = tid + j * THREADS;
      if (j < nj && c < H4) {
        const int g = c / H;
        float4 out;
        float* o = &out.x;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float z = mi ? (al[j] * wx[j][n] * uh[j][n] + b1[j] * uh[j][n] + b2[j] * wx[j][n] + bias[j])
                             : (wx[j][n] + uh[j][n] + bias[j]);
          o[n] = (g == 2) ? tanhf(z) : asr::hard_sigmoid(z);
        }
        *reinterpret_cast<float4*>(sG + (size_t)c * NS) = out;
      }
    }
    __syncthreads();
    // ---- C: cell update (unit owner) ---------------------------------------------------------------
    float cn[NM][NS], go[NM][NS];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int u = tid + m * THREADS;
      if (m < nm && u < H) {
        const float4 gi = *reinterpret_cast<const float4*>(sG + (size_t)u * NS);
        const float4 gf = *reinterpret_cast<const float4*>(sG + (size_t)(H + u) * NS);
        const float4 gg = *reinterpret_cast<const float4*>(sG + (size_t)(2 * H + u) * NS);
        const float4 gov = *reinterpret_cast<const float4*>(sG + (size_t)(3 * H + u) * NS);
        const float* pi = &gi.x; const float* pf = &gf.x; const float* pg = &gg.x; const float* po = &gov.x;
        const float kc = zone_coeff(v.zoneout_c, v.zmask_c, dir, T, H, t, u);
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float c_new = pf[n] * c_state[m][n] + pi[n] * pg[n];
          c_state[m][n] = c_state[m][n] + kc * (c_new - c_state[m][n]);
          cn[m][n] = c_state[m][n];
          go[m][n] = po[n];
          if (a.training && n0 + n < N) {
            const size_t row = (size_t)t * N + n0 + n;
            float* gp = a.gates + (row * 2 + dir) * H4;
            gp[u] = pi[n]; gp[H + u] = pf[n]; gp[2 * H + u] = pg[n]; gp[3 * H + u] = po[n];
            a.cell[(row * 2 + dir) * H + u] = c_state[m][n];
          }
        }
      } else {
#pragma unroll
        for (int n = 0; n < NS; ++n) { cn[m][n] = 0.0f; go[m][n] = 0.0f; }
      }
    }
    if (ln) {
      float mean[NS], rstd[NS];
      unit_stats(cn, nm, H, v.ln_eps, sRed, mean, rstd);
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int n = 0; n < NS; ++n) cn[m][n] = (cn[m][n] - mean[n]) * rstd[n] * gc[m] + bc[m];
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int u = tid + m * THREADS;
      if (m < nm && u < H) {
        const float kh = zone_coeff(v.zoneout_h, v.zmask_h, dir, T, H, t, u);
        float4 hm;
        float* ph = &hm.x;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float h_new = go[m][n] * tanhf(cn[m][n]);
          h_state[m][n] = h_state[m][n] + kh * (h_new - h_state[m][n]);
          ph[n] = h_state[m][n] * mu[m][n];
          if (n0 + n < N) {
            const size_t row = (size_t)t * N + n0 + n;
            if (a.h32) a.h32[row * 2 * H + dir * H + u] = h_state[m][n];
            if (a.h16) reinterpret_cast<__half*>(a.h16)[row * 2 * H + dir * H + u] = __float2half_rn(h_state[m][n]);
            if (a.training && a.hT16)
              reinterpret_cast<__nv_bfloat16*>(a.hT16)[(size_t)(dir * H + u) * R + row] = __float2bfloat16_rn(ph[n]);
          }
        }
        *reinterpret_cast<float4*>(sHm + (size_t)u * NS) = hm;
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// backward through time
// ------------------------------------------------------------------------------------------------
template <int NJ, int NM>
__global__ void __launch_bounds__(THREADS, 1)
cell_bwd_kernel(asr_lstm_bwd_args a, asr_lstm_variant v, const float* __restrict__ zx, const float* __restrict__ uh_raw,
                float* __restrict__ duh_out, asr_lstm_variant_grads pg) {
  extern __shared__ __align__(16) float smem[];
  const int T = a.T, N = a.N, H = a.H, H4 = 4 * a.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, dir = blockIdx.y, n0 = blockIdx.x * NS;
  const int nj = (H4 + THREADS - 1) / THREADS, nm = (H + THREADS - 1) / THREADS;
  float* sDz = smem;                       // [4H][NS]  dz (unit phase) then dUh (column phase)
  float* sDh = sDz + (size_t)H4 * NS;      // [H][NS]   dh_rec * B_U from the previous backward step
  float* sRed = sDh + (size_t)H * NS;      // [8][16]
  const bool ln = v.ln_gain_uh != nullptr, mi = v.mi_alpha != nullptr;
  const float* Ug = a.U + (size_t)dir * H * H4;

  float al[NJ], b1[NJ], b2[NJ], gu[NJ], bu[NJ], gw[NJ], bw[NJ];
  float g_b[NJ], g_al[NJ], g_b1[NJ], g_b2[NJ], g_gu[NJ], g_bu[NJ], g_gw[NJ], g_bw[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = tid + j * THREADS;
    const bool okc = j < nj && c < H4;
    const size_t o = (size_t)dir * H4 + c;
    al[j] = (okc && mi) ? v.mi_alpha[o] : 0.0f;
    b1[j] = (okc && mi) ? v.mi_beta1[o] : 1.0f;
    b2[j] = (okc && mi) ? v.mi_beta2[o] : 1.0f;
    gu[j] = (okc && ln) ? v.ln_gain_uh[o] : 1.0f;
    bu[j] = (okc && ln) ? v.ln_bias_uh[o] : 0.0f;
    gw[j] = (okc && ln) ? v.ln_gain_wx[o] : 1.0f;
    bw[j] = (okc && ln) ? v.ln_bias_wx[o] : 0.0f;
    g_b[j] = g_al[j] = g_b1[j] = g_b2[j] = g_gu[j] = g_bu[j] = g_gw[j] = g_bw[j] = 0.0f;
  }
  float dc_carry[NM][NS], dh_zone[NM][NS], mu[NM][NS], gc[NM], bc[NM], g_gc[NM], g_bc[NM];
#pragma unroll
  for (int m = 0; m < NM; ++m) {
    const int u = tid + m * THREADS;
    const bool oku = m < nm && u < H;
    gc[m] = (oku && ln) ? v.ln_gain_c[(size_t)dir * H + u] : 1.0f;
    bc[m] = (oku && ln) ? v.ln_bias_c[(size_t)dir * H + u] : 0.0f;
    g_gc[m] = g_bc[m] = 0.0f;
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      dc_carry[m][n] = 0.0f;
      dh_zone[m][n] = 0.0f;
      mu[m][n] = (oku && a.mask_u && n0 + n < N) ? a.mask_u[((size_t)dir * N + n0 + n) * H + u] : 1.0f;
    }
  }
  for (int i = tid; i < H * NS; i += THREADS) sDh[i] = 0.0f;
  __syncthreads();

  for (int s = 0; s < T; ++s) {
    const int t = dir ? s : (T - 1 - s);                  // reverse of the forward order
    const int t_prev = dir ? (t + 1) : (t - 1);           // the step the forward pass took before t
    const bool has_prev = dir ? (t + 1 < T) : (t > 0);
    // ---- C': unit owner: dh -> dz ------------------------------------------------------------------
    float cv[NM][NS], dh_new[NM][NS], gov[NM][NS];
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int u = tid + m * THREADS;
      const bool oku = m < nm && u < H;
      const float kh = oku ? zone_coeff(v.zoneout_h, v.zmask_h, dir, T, H, t, u) : 1.0f;
      float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
      if (oku) rec = *reinterpret_cast<const float4*>(sDh + (size_t)u * NS);
      const float* pr = &rec.x;
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const bool ok = oku && n0 + n < N;
        const size_t row = (size_t)t * N + n0 + n;
        const float dh = ok ? (__ldg(a.dh + row * 2 * H + dir * H + u) + pr[n] + dh_zone[m][n]) : 0.0f;
        dh_new[m][n] = kh * dh;
        dh_zone[m][n] = (1.0f - kh) * dh;
        cv[m][n] = ok ? __ldg(a.cell + (row * 2 + dir) * H + u) : 0.0f;
        gov[m][n] = ok ? __ldg(a.gates + (row * 2 + dir) * H4 + 3 * H + u) : 0.0f;
      }
    }
    float xhat_c[NM][NS], rstd_c[NS], dnc[NM][NS], d_o[NM][NS];
    if (ln) {
      float mean[NS];
      unit_stats(cv, nm, H, v.ln_eps, sRed, mean, rstd_c);
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int n = 0; n < NS; ++n) xhat_c[m][n] = (cv[m][n] - mean[n]) * rstd_c[n];
    }
#pragma unroll
    for (int m = 0; m < NM; ++m)
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const float nc = ln ? (xhat_c[m][n] * gc[m] + bc[m]) : cv[m][n];
        const float tnc = tanhf(nc);
        d_o[m][n] = dh_new[m][n] * tnc;
        dnc[m][n] = dh_new[m][n] * gov[m][n] * (1.0f - tnc * tnc);
      }
    if (ln) {   // LN backward on c: dc = rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
      float r[2 * NS];
#pragma unroll
      for (int n = 0; n < NS; ++n) r[n] = r[NS + n] = 0.0f;
#pragma unroll
      for (int m = 0; m < NM; ++m) {
        const bool oku = m < nm && tid + m * THREADS < H;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const float dxh = oku ? dnc[m][n] * gc[m] : 0.0f;
          r[n] += dxh;
          r[NS + n] += dxh * xhat_c[m][n];
          if (oku && n0 + n < N) { g_gc[m] += dnc[m][n] * xhat_c[m][n]; g_bc[m] += dnc[m][n]; }
        }
      }
      block_sum<2 * NS>(r, sRed);
#pragma unroll
      for (int m = 0; m < NM; ++m)
#pragma unroll
        for (int n = 0; n < NS; ++n)
          dnc[m][n] = rstd_c[n] * (dnc[m][n] * gc[m] - r[n] / (float)H - xhat_c[m][n] * r[NS + n] / (float)H);
    }
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int u = tid + m * THREADS;
      if (m < nm && u < H) {
        const float kc = zone_coeff(v.zoneout_c, v.zmask_c, dir, T, H, t, u);
        float4 zi, zf, zg, zo;
        float* pzi = &zi.x; float* pzf = &zf.x; float* pzg = &zg.x; float* pzo = &zo.x;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const bool ok = n0 + n < N;
          const size_t row = (size_t)t * N + n0 + n;
          const float* gp = a.gates + (row * 2 + dir) * H4;
          const float gi = ok ? __ldg(gp + u) : 0.0f, gf = ok ? __ldg(gp + H + u) : 0.0f, gg = ok ? __ldg(gp + 2 * H + u) : 0.0f;
          const float cp = (ok && has_prev) ? __ldg(a.cell + (((size_t)t_prev * N + n0 + n) * 2 + dir) * H + u) : 0.0f;
          const float dc = dc_carry[m][n] + dnc[m][n];
          const float dc_new = kc * dc;
          pzi[n] = dc_new * gg * asr::hard_sigmoid_grad(gi);
          pzf[n] = dc_new * cp * asr::hard_sigmoid_grad(gf);
          pzg[n] = dc_new * gi * (1.0f - gg * gg);
          pzo[n] = d_o[m][n] * asr::hard_sigmoid_grad(gov[m][n]);
          dc_carry[m][n] = (1.0f - kc) * dc + dc_new * gf;
        }
        *reinterpret_cast<float4*>(sDz + (size_t)u * NS) = zi;
        *reinterpret_cast<float4*>(sDz + (size_t)(H + u) * NS) = zf;
        *reinterpret_cast<float4*>(sDz + (size_t)(2 * H + u) * NS) = zg;
        *reinterpret_cast<float4*>(sDz + (size_t)(3 * H + u) * NS) = zo;
      }
    }
    __syncthreads();
    // ---- B': column owner: MI / LN backward -> dWx (global), dUh (global + smem) ----------------------
    float dz[NJ][NS], uh[NJ][NS], wx[NJ][NS];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = tid + j * THREADS;
      const bool okc = j < nj && c < H4;
      float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (okc) d4 = *reinterpret_cast<const float4*>(sDz + (size_t)c * NS);
      const float* pd = &d4.x;
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        const bool ok = okc && n0 + n < N;
        const size_t o = (((size_t)t * N + n0 + n) * 2 + dir) * H4 + c;
        dz[j][n] = ok ? pd[n] : 0.0f;
        wx[j][n] = ok ? __ldg(zx + o) : 0.0f;
        uh[j][n] = ok ? __ldg(uh_raw + o) : 0.0f;
        if (ok) g_b[j] += dz[j][n];
      }
    }
    float xh_u[NJ][NS], xh_w[NJ][NS], rs_u[NS], rs_w[NS];
    if (ln) {
      float mean[NS];
      row_stats(uh, nj, H4, v.ln_eps, sRed, mean, rs_u);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          xh_u[j][n] = (uh[j][n] - mean[n]) * rs_u[n];
          uh[j][n] = xh_u[j][n] * gu[j] + bu[j];
        }
      row_stats(wx, nj, H4, v.ln_eps, sRed, mean, rs_w);
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          xh_w[j][n] = (wx[j][n] - mean[n]) * rs_w[n];
          wx[j][n] = xh_w[j][n] * gw[j] + bw[j];
        }
    }
    float duh[NJ][NS], dwx[NJ][NS];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        if (mi) {
          g_al[j] += dz[j][n] * wx[j][n] * uh[j][n];
          g_b1[j] += dz[j][n] * uh[j][n];
          g_b2[j] += dz[j][n] * wx[j][n];
          duh[j][n] = dz[j][n] * (al[j] * wx[j][n] + b1[j]);
          dwx[j][n] = dz[j][n] * (al[j] * uh[j][n] + b2[j]);
        } else {
          duh[j][n] = dz[j][n];
          dwx[j][n] = dz[j][n];
        }
      }
    if (ln) {
      float r[4 * NS];
#pragma unroll
      for (int i = 0; i < 4 * NS; ++i) r[i] = 0.0f;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          g_gu[j] += duh[j][n] * xh_u[j][n];
          g_bu[j] += duh[j][n];
          g_gw[j] += dwx[j][n] * xh_w[j][n];
          g_bw[j] += dwx[j][n];
          const float du = duh[j][n] * gu[j], dw = dwx[j][n] * gw[j];
          r[n] += du;
          r[NS + n] += du * xh_u[j][n];
          r[2 * NS + n] += dw;
          r[3 * NS + n] += dw * xh_w[j][n];
        }
      block_sum<4 * NS>(r, sRed);
      const float inv = 1.0f / (float)H4;
#pragma unroll
      for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          duh[j][n] = rs_u[n] * (duh[j][n] * gu[j] - r[n] * inv - xh_u[j][n] * r[NS + n] * inv);
          dwx[j][n] = rs_w[n] * (dwx[j][n] * gw[j] - r[2 * NS + n] * inv - xh_w[j][n] * r[3 * NS + n] * inv);
        }
    }
    __syncthreads();                                       // everyone has read sDz
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = tid + j * THREADS;
      if (j < nj && c < H4) {
        float4 o4;
        float* po = &o4.x;
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          po[n] = duh[j][n];
          if (n0 + n < N) {
            const size_t o = (((size_t)t * N + n0 + n) * 2 + dir) * H4 + c;
            a.dz32[o] = dwx[j][n];
            duh_out[o] = duh[j][n];
          }
        }
        *reinterpret_cast<float4*>(sDz + (size_t)c * NS) = o4;
      }
    }
    __syncthreads();
    // ---- A': dh_rec[n][k] = sum_c dUh[n][c] * U[k][c]   (warp per k, lanes over c) ---------------------
    for (int k = warp; k < H; k += THREADS / 32) {
      const float* ur = Ug + (size_t)k * H4;
      float acc[NS] = {0.f, 0.f, 0.f, 0.f};
      for (int c = lane; c < H4; c += 32) {
        const float w = __ldg(ur + c);
        const float4 d = *reinterpret_cast<const float4*>(sDz + (size_t)c * NS);
        acc[0] = fmaf(d.x, w, acc[0]); acc[1] = fmaf(d.y, w, acc[1]);
        acc[2] = fmaf(d.z, w, acc[2]); acc[3] = fmaf(d.w, w, acc[3]);
      }
#pragma unroll
      for (int n = 0; n < NS; ++n) acc[n] = asr::warp_sum(acc[n]);
      if (lane == 0) {
        float4 o4;
        float* po = &o4.x;
#pragma unroll
        for (int n = 0; n < NS; ++n)
          po[n] = acc[n] * ((a.mask_u && n0 + n < N) ? a.mask_u[((size_t)dir * N + n0 + n) * H + k] : 1.0f);
        *reinterpret_cast<float4*>(sDh + (size_t)k * NS) = o4;
      }
    }
    __syncthreads();
  }
  // ---- parameter gradients: sum over the CTAs of a direction ------------------------------------------
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = tid + j * THREADS;
    if (j < nj && c < H4) {
      const size_t o = (size_t)dir * H4 + c;
      atomicAdd(a.dbias + o, g_b[j]);
      if (mi) { atomicAdd(pg.mi_alpha + o, g_al[j]); atomicAdd(pg.mi_beta1 + o, g_b1[j]); atomicAdd(pg.mi_beta2 + o, g_b2[j]); }
      if (ln) {
        atomicAdd(pg.ln_gain_uh + o, g_gu[j]); atomicAdd(pg.ln_bias_uh + o, g_bu[j]);
        atomicAdd(pg.ln_gain_wx + o, g_gw[j]); atomicAdd(pg.ln_bias_wx + o, g_bw[j]);
      }
    }
  }
  if (ln)
#pragma unroll
    for (int m = 0; m < NM; ++m) {
      const int u = tid + m * THREADS;
      if (m < nm && u < H) {
        atomicAdd(pg.ln_gain_c + (size_t)dir * H + u, g_gc[m]);
        atomicAdd(pg.ln_bias_c + (size_t)dir * H + u, g_bc[m]);
      }
    }
}

// ---- host ---------------------------------------------------------------------------------------
static int32_t check_variant(const asr_lstm_variant* v, int H) {
  ASR_CHECK_ARG(v, "lstm_cell: null variant");
  ASR_CHECK_ARG(H >= 1 && 4 * H <= JMAX * THREADS && H <= MMAX * THREADS, "lstm_cell: H=%d out of range (<= 1024)", H);
  const bool mi_all = v->mi_alpha && v->mi_beta1 && v->mi_beta2, mi_none = !v->mi_alpha && !v->mi_beta1 && !v->mi_beta2;
  ASR_CHECK_ARG(mi_all || mi_none, "lstm_cell: mi_alpha/beta1/beta2 must be given together");
  const bool ln_all = v->ln_gain_uh && v->ln_bias_uh && v->ln_gain_wx && v->ln_bias_wx && v->ln_gain_c && v->ln_bias_c;
  const bool ln_none = !v->ln_gain_uh && !v->ln_bias_uh && !v->ln_gain_wx && !v->ln_bias_wx && !v->ln_gain_c && !v->ln_bias_c;
  ASR_CHECK_ARG(ln_all || ln_none, "lstm_cell: the six layer-norm vectors must be given together");
  ASR_CHECK_ARG(v->zoneout_h >= 0.0f && v->zoneout_h < 1.0f && v->zoneout_c >= 0.0f && v->zoneout_c < 1.0f,
                "lstm_cell: zoneout level must be in [0, 1)");
  ASR_CHECK_ARG(!ln_all || v->ln_eps > 0.0f, "lstm_cell: ln_eps must be > 0");
  return ASR_OK;
}

int32_t forward(const asr_lstm_fwd_args* a, const asr_lstm_variant* v, float* uh_raw, cudaStream_t st) {
  if (int32_t rc = check_variant(v, a->H)) return rc;
  ASR_CHECK_ARG(a->U, "lstm_cell forward: needs the fp32 U");
  ASR_CHECK_ARG(!a->training || uh_raw, "lstm_cell forward: training needs the uh_raw buffer");
  const size_t smem = ((size_t)a->H * NS + (size_t)4 * a->H * NS + 8 * 16) * sizeof(float);
  const dim3 grid((a->N + NS - 1) / NS, 2);
  float* up = a->training ? uh_raw : nullptr;
  const int need = (4 * a->H + THREADS - 1) / THREADS;
#define CELL_FWD(NJ_, NM_)                                                                                              \
  do {                                                                                                                  \
    ASR_CUDA(cudaFuncSetAttribute(cell_fwd_kernel<NJ_, NM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cell_fwd_kernel<NJ_, NM_><<<grid, THREADS, smem, st>>>(*a, *v, up);                                               \
  } while (0)
  if (need <= 2) CELL_FWD(2, 1);
  else if (need <= 4) CELL_FWD(4, 1);
  else if (need <= 8) CELL_FWD(8, 2);
  else CELL_FWD(16, 4);
#undef CELL_FWD
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

int32_t backward(const asr_lstm_bwd_args* a, const asr_lstm_variant* v, const float* zx, const float* uh_raw, float* duh,
                 const asr_lstm_variant_grads* g, cudaStream_t st) {
  if (int32_t rc = check_variant(v, a->H)) return rc;
  ASR_CHECK_ARG(a->U && a->dz32 && zx && uh_raw && duh && g, "lstm_cell backward: null argument");
  ASR_CHECK_ARG(!v->mi_alpha || (g->mi_alpha && g->mi_beta1 && g->mi_beta2), "lstm_cell backward: MI gradient buffers missing");
  ASR_CHECK_ARG(!v->ln_gain_uh || (g->ln_gain_uh && g->ln_bias_uh && g->ln_gain_wx && g->ln_bias_wx && g->ln_gain_c && g->ln_bias_c),
                "lstm_cell backward: LN gradient buffers missing");
  const int H = a->H;
  ASR_CUDA(cudaMemsetAsync(a->dbias, 0, (size_t)2 * 4 * H * sizeof(float), st));
  if (v->mi_alpha) {
    ASR_CUDA(cudaMemsetAsync(g->mi_alpha, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->mi_beta1, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->mi_beta2, 0, (size_t)8 * H * sizeof(float), st));
  }
  if (v->ln_gain_uh) {
    ASR_CUDA(cudaMemsetAsync(g->ln_gain_uh, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->ln_bias_uh, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->ln_gain_wx, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->ln_bias_wx, 0, (size_t)8 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->ln_gain_c, 0, (size_t)2 * H * sizeof(float), st));
    ASR_CUDA(cudaMemsetAsync(g->ln_bias_c, 0, (size_t)2 * H * sizeof(float), st));
  }
  const size_t smem = ((size_t)4 * H * NS + (size_t)H * NS + 8 * 16) * sizeof(float);
  const dim3 grid((a->N + NS - 1) / NS, 2);
  const int need = (4 * H + THREADS - 1) / THREADS;
#define CELL_BWD(NJ_, NM_)                                                                                              \
  do {                                                                                                                  \
    ASR_CUDA(cudaFuncSetAttribute(cell_bwd_kernel<NJ_, NM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    cell_bwd_kernel<NJ_, NM_><<<grid, THREADS, smem, st>>>(*a, *v, zx, uh_raw, duh, *g);                               \
  } while (0)
  if (need <= 2) CELL_BWD(2, 1);
  else if (need <= 4) CELL_BWD(4, 1);
  else if (need <= 8) CELL_BWD(8, 2);
  else CELL_BWD(16, 4);
#undef CELL_BWD
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

}  // namespace lstmcell
