// K1 — fused MFCC / log-mel front end for sm_100a.
//
// One launch does, per utterance chunk of 64 frames (+halo):
//   HBM-coalesced pcm -> smem staging -> pre-emphasis -> framing + Hamming ->
//   512-pt real FFT (256-pt complex radix-2 DIF: 3 in-register stages + 5
//   warp-shuffle stages, one frame per warp) -> power spectrum -> mel filterbank
//   -> log -> DCT-II(+lifter) -> energy -> delta / delta-delta -> [N,T,F] write,
// accumulates per-utterance CMVN statistics (fp64 atomics) and lets the LAST
// chunk CTA of each utterance normalise it in place (threadfence-reduction
// pattern), so the whole of Feature.__call__ is a single kernel.
//
// Reference: preprocessing/audio.py:41-75,223-253,339-388,419-442 and
// preprocessing/audio_utils.py:17-50,98-120,143-173 (arithmetic there is fp64
// numpy; here fp32 with fp64 CMVN statistics — parity tolerance 1e-3, tests/).
#include "common.cuh"
#include <math.h>
#include <vector>
#include <string.h>

namespace {

constexpr int CHUNK = 64;       // frames per CTA
constexpr int WARPS = 8;
constexpr int THREADS = WARPS * 32;
constexpr int NFFT = 512;
constexpr int NBIN = NFFT / 2 + 1;
constexpr int MAX_FILT = 64;
constexpr int MAX_CEP = 32;
constexpr int ZPAD = 288;       // 256 + 256/8 padding (bank-conflict-free SoA)

struct Tables {
  const float* window;    // [NFFT] (zeros past frame_len)
  const float2* tw256;    // [128]  W_256^j = (cos, -sin)
  const float2* tw512;    // [257]  W_512^k
  const int* fb_start;    // [MAX_FILT]
  const int* fb_len;      // [MAX_FILT]
  const int* fb_off;      // [MAX_FILT]
  const float* fb_w;      // packed triangle weights
  const float* dct;       // [num_cep][num_filt] with lifter folded in
};

struct Params {
  Tables tb;
  int frame_len, frame_step;
  int num_filt, num_cep, kind, append_energy, d, dd;
  int mean_norm, var_norm, stride;
  int base_dim, feat_dim, halo;
  float pre_emph, eps;
  int sig_cap;  // floats reserved for the staged signal
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ int zidx(int k) { return k + (k >> 3); }

// 256-point complex FFT, decimation in frequency; lane l holds n = l + 32a.
// On return z[a] holds Z[k], k = 8*bitrev5(l) + bitrev3(a).
__device__ __forceinline__ void fft256(float2 (&z)[8], const float2 (&twA)[4],
                                       const float2 (&twB)[2], float2 twC,
                                       const float2 (&twX)[4], int lane) {
  // stage M=256: pairs (a, a+4), twiddle W256^(l+32a)
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    float2 u = z[a], v = z[a + 4];
    z[a] = make_float2(u.x + v.x, u.y + v.y);
    z[a + 4] = cmul(make_float2(u.x - v.x, u.y - v.y), twA[a]);
  }
  // stage M=128: pairs (a, a+2) within each half, twiddle W128^(l+32(a&1))
#pragma unroll
  for (int hb = 0; hb < 8; hb += 4) {
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      float2 u = z[hb + a], v = z[hb + a + 2];
      z[hb + a] = make_float2(u.x + v.x, u.y + v.y);
      z[hb + a + 2] = cmul(make_float2(u.x - v.x, u.y - v.y), twB[a]);
    }
  }
  // stage M=64: pairs (a, a+1), twiddle W64^l
#pragma unroll
  for (int a = 0; a < 8; a += 2) {
    float2 u = z[a], v = z[a + 1];
    z[a] = make_float2(u.x + v.x, u.y + v.y);
    z[a + 1] = cmul(make_float2(u.x - v.x, u.y - v.y), twC);
  }
  // stages M=32..2 across lanes
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int off = 16 >> s;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      float px = __shfl_xor_sync(0xffffffffu, z[a].x, off);
      float py = __shfl_xor_sync(0xffffffffu, z[a].y, off);
      if (!upper) {
        z[a] = make_float2(z[a].x + px, z[a].y + py);
      } else {
        float2 dlt = make_float2(px - z[a].x, py - z[a].y);
        z[a] = (s < 4) ? cmul(dlt, twX[s]) : dlt;
      }
    }
  }
}

__global__ void __launch_bounds__(THREADS)
mfcc_kernel(Params p, const float* __restrict__ pcm, const int64_t* __restrict__ offsets,
            int n_utt, int t_max, float* __restrict__ out, int* __restrict__ out_len,
            int time_major, int* __restrict__ counters, double* __restrict__ stats) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = blockIdx.y;
  const int f0 = blockIdx.x * CHUNK;
  const int64_t off = offsets[n];
  const int slen = (int)(offsets[n + 1] - off);
  const int F = p.feat_dim, BD = p.base_dim;

  int nf_raw = 1;
  if (slen > p.frame_len) nf_raw = 1 + (slen - p.frame_len + p.frame_step - 1) / p.frame_step;
  const int t_out = (nf_raw + p.stride - 1) / p.stride;
  if (blockIdx.x == 0 && tid == 0) out_len[n] = t_out;

  auto out_row = [&](int row) -> float* {
    return time_major ? out + ((int64_t)row * n_utt + n) * F : out + ((int64_t)n * t_max + row) * F;
  };

  if (f0 >= nf_raw) {  // pure padding chunk: zero-fill 'post'
    for (int i = tid; i < CHUNK * F; i += THREADS) {
      int t = f0 + i / F, c = i % F;
      if (t % p.stride == 0 && t / p.stride < t_max) out_row(t / p.stride)[c] = 0.0f;
    }
    return;
  }

  const int lo = max(f0 - p.halo, 0);
  const int hi = min(f0 + CHUNK + p.halo, nf_raw);
  const int nb = hi - lo;
  const int s0 = lo * p.frame_step - 1;  // first staged sample (one extra for pre-emphasis)
  const int ns = (nb - 1) * p.frame_step + p.frame_len + 1;

  float* sig = smem;                                   // [sig_cap]
  float* base = sig + p.sig_cap;                       // [CHUNK+8][BD]
  float* dl = base + (CHUNK + 8) * BD;                 // [CHUNK+4][BD]
  float* wscr = dl + (CHUNK + 4) * BD;                 // per-warp scratch
  float* zre = wscr + warp * (2 * ZPAD + 264 + MAX_FILT);
  float* zim = zre + ZPAD;
  float* pw = zim + ZPAD;                              // [264]
  float* lm = pw + 264;                                // [MAX_FILT]

  // ---- A. coalesced stage of the raw samples ---------------------------------
  for (int i = tid; i < ns; i += THREADS) {
    int g = s0 + i;
    sig[i] = (g >= 0 && g < slen) ? __ldg(pcm + off + g) : 0.0f;
  }
  __syncthreads();

  // per-lane twiddles (fixed across frames)
  float2 twA[4], twB[2], twC, twX[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) twA[a] = p.tb.tw256[lane + 32 * a];
#pragma unroll
  for (int a = 0; a < 2; ++a) twB[a] = p.tb.tw256[2 * (lane + 32 * a)];
  twC = p.tb.tw256[4 * lane];
  twX[0] = p.tb.tw256[8 * (lane & 15)];
  twX[1] = p.tb.tw256[16 * (lane & 7)];
  twX[2] = p.tb.tw256[32 * (lane & 3)];
  twX[3] = p.tb.tw256[64 * (lane & 1)];
  const int rev5 = __brev((unsigned)lane) >> 27;

  // ---- B. one frame per warp ---------------------------------------------------
  for (int fl = warp; fl < nb; fl += WARPS) {
    const int t = lo + fl;
    const int i0 = t * p.frame_step - s0;     // index of sample 0 of this frame in sig (>= 1)
    const int g0 = t * p.frame_step;          // global sample index of sample 0
    float2 z[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int j = 2 * (lane + 32 * a);
      float v[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int jj = j + e;
        float y = 0.0f;
        if (jj < p.frame_len) {
          const int g = g0 + jj;
          if (g < slen) {
            const float x = sig[i0 + jj];
            // numpy: x[n] - coeff*x[n-1] with two fp32 roundings (no FMA contraction)
            y = (g == 0) ? x : __fsub_rn(x, __fmul_rn(p.pre_emph, sig[i0 + jj - 1]));
          }
          y *= __ldg(p.tb.window + jj);
        }
        v[e] = y;
      }
      z[a] = make_float2(v[0], v[1]);
    }
    fft256(z, twA, twB, twC, twX, lane);
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      const int k = 8 * rev5 + ((a & 1) << 2 | (a & 2) | (a >> 2));
      zre[zidx(k)] = z[a].x;
      zim[zidx(k)] = z[a].y;
    }
    __syncwarp();
    // real-FFT split + power spectrum
    float esum = 0.0f;
    for (int k = lane; k < NBIN; k += 32) {
      const int ka = k & 255, kb = (256 - k) & 255;
      const float ar = zre[zidx(ka)], ai = zim[zidx(ka)];
      const float br = zre[zidx(kb)], bi = -zim[zidx(kb)];
      const float er = 0.5f * (ar + br), ei = 0.5f * (ai + bi);
      const float orr = 0.5f * (ar - br), oi = 0.5f * (ai - bi);
      const float2 w = p.tb.tw512[k];
      const float xr = er + (w.x * oi + w.y * orr);
      const float xi = ei - (w.x * orr - w.y * oi);
      const float pk = (xr * xr + xi * xi) * (1.0f / NFFT);
      pw[k] = pk;
      esum += pk;
    }
    esum = asr::warp_sum(esum);
    if (esum == 0.0f) esum = 2.220446049250313e-16f;
    __syncwarp();
    // mel filterbank
    for (int j = lane; j < p.num_filt; j += 32) {
      const int st = p.tb.fb_start[j], ln = p.tb.fb_len[j];
      const float* w = p.tb.fb_w + p.tb.fb_off[j];
      float acc = 0.0f;
      for (int i = 0; i < ln; ++i) acc = fmaf(__ldg(w + i), pw[st + i], acc);
      if (acc == 0.0f) acc = 2.220446049250313e-16f;
      if (p.kind == 0) lm[j] = logf(acc);
      else base[fl * BD + j] = (p.kind == 1) ? logf(acc) : acc;
    }
    __syncwarp();
    if (p.kind == 0) {
      if (lane < p.num_cep) {
        const float* dr = p.tb.dct + lane * p.num_filt;
        float acc = 0.0f;
        for (int m = 0; m < p.num_filt; ++m) acc = fmaf(__ldg(dr + m), lm[m], acc);
        if (lane == 0 && p.append_energy) acc = logf(esum + p.eps);
        base[fl * BD + lane] = acc;
      }
    } else if (p.kind == 1 && p.append_energy) {
      if (lane == 0) base[fl * BD + p.num_filt] = logf(esum + p.eps);
    }
    __syncwarp();
  }
  __syncthreads();

  // ---- C. deltas (audio_utils.py:153-173; edge-replicated, /10) -------------------
  auto bidx = [&](int t) { return (min(max(t, 0), nf_raw - 1) - lo) * BD; };
  const int dlo = p.dd ? max(f0 - 2, 0) : f0;
  const int dhi = p.dd ? min(f0 + CHUNK + 2, nf_raw) : min(f0 + CHUNK, nf_raw);
  if (p.d) {
    for (int i = tid; i < (dhi - dlo) * BD; i += THREADS) {
      const int t = dlo + i / BD, k = i % BD;
      const float v = 2.0f * (base[bidx(t + 2) + k] - base[bidx(t - 2) + k]) +
                      (base[bidx(t + 1) + k] - base[bidx(t - 1) + k]);
      dl[(t - dlo) * BD + k] = v / 10.0f;
    }
    __syncthreads();
  }
  auto didx = [&](int t) { return (min(max(t, 0), nf_raw - 1) - dlo) * BD; };

  // ---- D. emit rows + CMVN statistics ---------------------------------------------
  const int tend = min(f0 + CHUNK, nf_raw);
  auto feat_at = [&](int t, int c) -> float {
    if (c < BD) return base[(t - lo) * BD + c];
    if (c < 2 * BD) return dl[(t - dlo) * BD + (c - BD)];
    const int k = c - 2 * BD;
    return (2.0f * (dl[didx(t + 2) + k] - dl[didx(t - 2) + k]) +
            (dl[didx(t + 1) + k] - dl[didx(t - 1) + k])) / 10.0f;
  };
  for (int i = tid; i < (tend - f0) * F; i += THREADS) {
    const int t = f0 + i / F, c = i % F;
    if (t % p.stride == 0 && t / p.stride < t_max) out_row(t / p.stride)[c] = feat_at(t, c);
  }
  for (int i = tid; i < (f0 + CHUNK - tend) * F; i += THREADS) {   // 'post' padding in a straddling chunk
    const int t = tend + i / F, c = i % F;
    if (t % p.stride == 0 && t / p.stride < t_max) out_row(t / p.stride)[c] = 0.0f;
  }
  const bool norm = p.mean_norm || p.var_norm;
  if (!norm) return;
  for (int c = tid; c < F; c += THREADS) {
    double s = 0.0, ss = 0.0;
    for (int t = f0; t < tend; ++t) {
      if (t % p.stride) continue;
      const double v = (double)feat_at(t, c);
      s += v;
      ss += v * v;
    }
    atomicAdd(stats + ((int64_t)n * 2 + 0) * F + c, s);
    atomicAdd(stats + ((int64_t)n * 2 + 1) * F + c, ss);
  }

  // ---- E. last chunk CTA of the utterance normalises it ---------------------------
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int nactive = (nf_raw + CHUNK - 1) / CHUNK;
    s_last = (atomicAdd(counters + n, 1) == nactive - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double* mean = reinterpret_cast<double*>(smem);   // reuse staging area: [F] mean, [F] inv
  double* inv = mean + F;
  for (int c = tid; c < F; c += THREADS) {
    const double s = __ldcg(stats + ((int64_t)n * 2 + 0) * F + c);
    const double ss = __ldcg(stats + ((int64_t)n * 2 + 1) * F + c);
    const double m = s / t_out;
    double var = ss / t_out - m * m;
    if (var < 0.0) var = 0.0;
    mean[c] = p.mean_norm ? m : 0.0;
    inv[c] = p.var_norm ? 1.0 / (sqrt(var) + (double)p.eps) : 1.0;
    stats[((int64_t)n * 2 + 0) * F + c] = 0.0;     // leave the workspace zeroed
    stats[((int64_t)n * 2 + 1) * F + c] = 0.0;
  }
  if (tid == 0) counters[n] = 0;
  __syncthreads();
  for (int i = tid; i < min(t_out, t_max) * F; i += THREADS) {
    const int r = i / F, c = i % F;
    float* q = out_row(r) + c;
    *q = (float)(((double)__ldcg(q) - mean[c]) * inv[c]);
  }
}

// ---- +-num_context frames (Feature._postprocessing, audio.py:88-150) + CMVN over the widened matrix ------------
// raw: un-normalised (strided) features [N, t_max, F] from the fused kernel.  Output column j = (k, f) with
// k = j / F - ctx is column f shifted by k frames, zeros outside [0, len) ("empty_mfcc", audio.py:97-98); every
// output column is then standardised on its own over the utterance's len rows (audio.py:70-75 runs after
// _postprocessing, audio.py:65).  One CTA per (utterance, output column): two-pass mean / deviation in fp64.
__global__ void __launch_bounds__(256)
mfcc_context_kernel(const float* __restrict__ raw, const int* __restrict__ out_len, int n_utt, int t_max, int F, int ctx,
                    int mean_norm, int var_norm, float eps, float* __restrict__ out, int time_major) {
  const int n = blockIdx.y, j = blockIdx.x, tid = threadIdx.x;
  const int Fx = F * (1 + 2 * ctx), k = j / F - ctx, f = j % F;
  const int len = min(out_len[n], t_max);
  const float* col = raw + (size_t)n * t_max * F + f;
  auto at = [&](int t) -> float { const int ts = t + k; return (ts >= 0 && ts < len) ? col[(size_t)ts * F] : 0.0f; };
  __shared__ double red[256];
  auto block_sum = [&](double v) -> double {
    red[tid] = v;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) red[tid] += red[tid + o];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  double s = 0.0;
  for (int t = tid; t < len; t += 256) s += (double)at(t);
  const double mean = len > 0 ? block_sum(s) / len : 0.0;
  double ss = 0.0;
  for (int t = tid; t < len; t += 256) { const double d = (double)at(t) - mean; ss += d * d; }
  const double var = len > 0 ? block_sum(ss) / len : 0.0;
  const double m = mean_norm ? mean : 0.0, inv = var_norm ? 1.0 / (sqrt(var) + (double)eps) : 1.0;
  for (int t = tid; t < t_max; t += 256) {
    float* q = time_major ? out + ((size_t)t * n_utt + n) * Fx + j : out + ((size_t)n * t_max + t) * Fx + j;
    *q = (t < len) ? (float)(((double)at(t) - m) * inv) : 0.0f;
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
// host side: plan (constant tables computed in fp64 like the reference's init code)
// ------------------------------------------------------------------------------------
struct asr_mfcc_plan {
  asr_mfcc_config cfg;
  Params p;
  void* dev_tables = nullptr;
  size_t smem_bytes = 0;
};

static int round_half_up(double x) { return (int)floor(x + 0.5); }  // audio_utils.py:11-14 (x >= 0)

extern "C" int32_t asr_mfcc_plan_create(const asr_mfcc_config* cfg, asr_mfcc_plan** out) {
  ASR_CHECK_ARG(cfg && out, "asr_mfcc_plan_create: null argument");
  ASR_CHECK_ARG(cfg->nfft == NFFT, "mfcc: only nfft=512 is built (got %d)", cfg->nfft);
  ASR_CHECK_ARG(cfg->num_context >= 0 && cfg->num_context <= 64, "mfcc: num_context out of range");
  ASR_CHECK_ARG(cfg->high_freq <= cfg->fs / 2, "high_freq must be less or equal than fs/2");  // audio.py:186
  ASR_CHECK_ARG(cfg->num_filt >= 1 && cfg->num_filt <= MAX_FILT, "mfcc: num_filt out of range");
  ASR_CHECK_ARG(cfg->kind >= 0 && cfg->kind <= 2, "mfcc: kind must be 0..2");
  ASR_CHECK_ARG(cfg->kind != 0 || (cfg->num_cep >= 1 && cfg->num_cep <= MAX_CEP && cfg->num_cep <= cfg->num_filt),
                "mfcc: num_cep out of range");
  ASR_CHECK_ARG(cfg->stride >= 1, "mfcc: stride must be >= 1");
  const int frame_len = round_half_up((double)cfg->win_len * cfg->fs);
  const int frame_step = round_half_up((double)cfg->win_step * cfg->fs);
  ASR_CHECK_ARG(frame_len >= 2 && frame_len <= NFFT && frame_step >= 1, "mfcc: frame_len must be in [2, nfft]");

  asr_mfcc_plan* pl = new asr_mfcc_plan();
  pl->cfg = *cfg;
  Params& p = pl->p;
  memset(&p, 0, sizeof(p));
  p.frame_len = frame_len;
  p.frame_step = frame_step;
  p.num_filt = cfg->num_filt;
  p.num_cep = cfg->num_cep;
  p.kind = cfg->kind;
  p.append_energy = cfg->append_energy ? 1 : 0;
  p.d = cfg->d ? 1 : 0;
  p.dd = (cfg->d && cfg->dd) ? 1 : 0;  // audio.py:360-365: dd only inside `if self.d`
  p.mean_norm = cfg->mean_norm ? 1 : 0;
  p.var_norm = cfg->var_norm ? 1 : 0;
  p.stride = cfg->stride;
  p.pre_emph = cfg->pre_emph;
  p.eps = cfg->eps;
  p.base_dim = (cfg->kind == 0) ? cfg->num_cep : cfg->num_filt + ((cfg->kind == 1 && cfg->append_energy) ? 1 : 0);
  if (cfg->kind == 2) { p.d = p.dd = 0; }
  p.feat_dim = p.base_dim * (1 + p.d + p.dd);
  p.halo = p.dd ? 4 : (p.d ? 2 : 0);
  p.sig_cap = ((CHUNK + 2 * 4 - 1) * frame_step + frame_len + 1 + 3) & ~3;

  // ---- tables (fp64 -> fp32) ----
  const double PI = 3.14159265358979323846;
  std::vector<float> window(NFFT, 0.0f);
  for (int i = 0; i < frame_len; ++i)  // scipy.signal.hamming (symmetric), audio.py:182
    window[i] = (float)(0.54 - 0.46 * cos(2.0 * PI * i / (frame_len - 1)));
  std::vector<float2> tw256(128), tw512(NBIN);
  for (int j = 0; j < 128; ++j) tw256[j] = make_float2((float)cos(2 * PI * j / 256), (float)-sin(2 * PI * j / 256));
  for (int k = 0; k < NBIN; ++k) tw512[k] = make_float2((float)cos(2 * PI * k / 512), (float)-sin(2 * PI * k / 512));
  // mel filterbank, audio.py:201-203, 255-277
  const int nf = cfg->num_filt;
  std::vector<double> bins(nf + 2);
  {
    auto hz2mel = [](double hz) { return 2595.0 * log10(1.0 + hz / 700.0); };
    auto mel2hz = [](double mel) { return 700.0 * (pow(10.0, mel / 2595.0) - 1.0); };
    const double lo = hz2mel(cfg->low_freq), hi = hz2mel(cfg->high_freq);
    for (int i = 0; i < nf + 2; ++i) {
      const double mel = lo + (hi - lo) * i / (nf + 1);   // np.linspace
      bins[i] = floor((NFFT + 1) * mel2hz(mel) / cfg->fs);
    }
  }
  std::vector<int> fb_start(MAX_FILT, 0), fb_len(MAX_FILT, 0), fb_off(MAX_FILT, 0);
  std::vector<float> fb_w;
  for (int j = 0; j < nf; ++j) {
    const int b0 = (int)bins[j], b1 = (int)bins[j + 1], b2 = (int)bins[j + 2];
    fb_start[j] = b0;
    fb_len[j] = (b2 > b0) ? (b2 - b0) : 0;
    fb_off[j] = (int)fb_w.size();
    if (b2 > NBIN || b0 < 0) { delete pl; asr::set_error("mfcc: filterbank bin out of range"); return ASR_ERR_INVALID; }
    for (int i = b0; i < b1; ++i) fb_w.push_back((float)((i - bins[j]) / (bins[j + 1] - bins[j])));
    for (int i = b1; i < b2; ++i) fb_w.push_back((float)((bins[j + 2] - i) / (bins[j + 2] - bins[j + 1])));
  }
  if (fb_w.empty()) fb_w.push_back(0.0f);
  // DCT-II ortho (scipy.fftpack.dct norm='ortho', audio.py:353) with lifter (audio.py:381-385) folded in
  const int nc = (cfg->kind == 0) ? cfg->num_cep : 1;
  std::vector<float> dct((size_t)nc * nf, 0.0f);
  if (cfg->kind == 0) {
    for (int k = 0; k < nc; ++k) {
      const double lift = (cfg->cep_lifter > 0) ? 1.0 + (cfg->cep_lifter / 2.0) * sin(PI * k / cfg->cep_lifter) : 1.0;
      const double sc = (k == 0) ? sqrt(1.0 / nf) : sqrt(2.0 / nf);
      for (int m = 0; m < nf; ++m) dct[(size_t)k * nf + m] = (float)(lift * sc * cos(PI * k * (2 * m + 1) / (2.0 * nf)));
    }
  }
  // pack into one device allocation
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o_win = 0, o_t256 = al(o_win + NFFT * 4), o_t512 = al(o_t256 + 128 * 8), o_fs = al(o_t512 + NBIN * 8),
         o_fl = al(o_fs + MAX_FILT * 4), o_fo = al(o_fl + MAX_FILT * 4), o_fw = al(o_fo + MAX_FILT * 4),
         o_dct = al(o_fw + fb_w.size() * 4), total = al(o_dct + dct.size() * 4);
  std::vector<char> host(total, 0);
  memcpy(&host[o_win], window.data(), NFFT * 4);
  memcpy(&host[o_t256], tw256.data(), 128 * 8);
  memcpy(&host[o_t512], tw512.data(), NBIN * 8);
  memcpy(&host[o_fs], fb_start.data(), MAX_FILT * 4);
  memcpy(&host[o_fl], fb_len.data(), MAX_FILT * 4);
  memcpy(&host[o_fo], fb_off.data(), MAX_FILT * 4);
  memcpy(&host[o_fw], fb_w.data(), fb_w.size() * 4);
  memcpy(&host[o_dct], dct.data(), dct.size() * 4);
  cudaError_t e = cudaMalloc(&pl->dev_tables, total);
  if (e == cudaSuccess) e = cudaMemcpy(pl->dev_tables, host.data(), total, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    asr::set_error("mfcc plan: %s", cudaGetErrorString(e));
    if (pl->dev_tables) cudaFree(pl->dev_tables);
    delete pl;
    return ASR_ERR_CUDA;
  }
  char* d = (char*)pl->dev_tables;
  p.tb.window = (const float*)(d + o_win);
  p.tb.tw256 = (const float2*)(d + o_t256);
  p.tb.tw512 = (const float2*)(d + o_t512);
  p.tb.fb_start = (const int*)(d + o_fs);
  p.tb.fb_len = (const int*)(d + o_fl);
  p.tb.fb_off = (const int*)(d + o_fo);
  p.tb.fb_w = (const float*)(d + o_fw);
  p.tb.dct = (const float*)(d + o_dct);

  const size_t fl = (size_t)p.sig_cap + (size_t)(CHUNK + 8) * p.base_dim + (size_t)(CHUNK + 4) * p.base_dim +
                    (size_t)WARPS * (2 * ZPAD + 264 + MAX_FILT);
  pl->smem_bytes = fl * sizeof(float);
  if (pl->smem_bytes > 200 * 1024 || (size_t)p.sig_cap * 4 < (size_t)p.feat_dim * 16) {
    asr::set_error("mfcc: configuration needs %zu B of shared memory", pl->smem_bytes);
    cudaFree(pl->dev_tables);
    delete pl;
    return ASR_ERR_INVALID;
  }
  e = cudaFuncSetAttribute(mfcc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl->smem_bytes);
  if (e != cudaSuccess) {
    asr::set_error("mfcc plan: %s", cudaGetErrorString(e));
    cudaFree(pl->dev_tables);
    delete pl;
    return ASR_ERR_CUDA;
  }
  *out = pl;
  return ASR_OK;
}

extern "C" void asr_mfcc_plan_destroy(asr_mfcc_plan* plan) {
  if (!plan) return;
  if (plan->dev_tables) cudaFree(plan->dev_tables);
  delete plan;
}

extern "C" int32_t asr_mfcc_num_feats(const asr_mfcc_plan* plan) {
  return plan ? plan->p.feat_dim * (1 + 2 * plan->cfg.num_context) : ASR_ERR_INVALID;   // audio.py:146
}

extern "C" int32_t asr_mfcc_num_frames(const asr_mfcc_plan* plan, int64_t num_samples) {
  if (!plan) return ASR_ERR_INVALID;
  int64_t nf = 1;
  if (num_samples > plan->p.frame_len)
    nf = 1 + (num_samples - plan->p.frame_len + plan->p.frame_step - 1) / plan->p.frame_step;
  return (int32_t)((nf + plan->p.stride - 1) / plan->p.stride);
}

static size_t stats_bytes(const asr_mfcc_plan* plan, int32_t n) {
  return (size_t)n * 2 * plan->p.feat_dim * sizeof(double) + (((size_t)n * sizeof(int) + 15) & ~(size_t)15);
}
extern "C" size_t asr_mfcc_workspace_bytes(const asr_mfcc_plan* plan, int32_t n) {
  if (!plan || n <= 0) return 0;
  return stats_bytes(plan, n);
}
// with num_context > 0 the fused kernel's un-normalised [n, t_max, F] features are staged in the workspace too
extern "C" size_t asr_mfcc_workspace_bytes_ex(const asr_mfcc_plan* plan, int32_t n, int32_t t_max) {
  if (!plan || n <= 0 || t_max <= 0) return 0;
  size_t b = (stats_bytes(plan, n) + 255) & ~(size_t)255;
  if (plan->cfg.num_context > 0) b += (size_t)n * t_max * plan->p.feat_dim * sizeof(float);
  return b;
}

extern "C" int32_t asr_mfcc_forward(const asr_mfcc_plan* plan, const float* pcm, const int64_t* offsets, int32_t n,
                                    int32_t t_max, float* out, int32_t* out_len, int32_t time_major, void* ws,
                                    void* stream) {
  ASR_CHECK_ARG(plan && pcm && offsets && out && out_len && ws, "asr_mfcc_forward: null argument");
  ASR_CHECK_ARG(n >= 1 && n <= 65535 && t_max >= 1, "asr_mfcc_forward: bad n=%d / t_max=%d", n, t_max);
  double* stats = (double*)ws;
  int* counters = (int*)((char*)ws + (size_t)n * 2 * plan->p.feat_dim * sizeof(double));
  const int raw = t_max * plan->p.stride;
  dim3 grid((raw + CHUNK - 1) / CHUNK, n);
  const int ctx = plan->cfg.num_context;
  if (ctx == 0) {
    mfcc_kernel<<<grid, THREADS, plan->smem_bytes, (cudaStream_t)stream>>>(plan->p, pcm, offsets, n, t_max, out, out_len,
                                                                            time_major, counters, stats);
    ASR_LAUNCH_CHECK();
    return ASR_OK;
  }
  // num_context > 0 (audio.py:88-150): un-normalised features into the workspace (sized by
  // asr_mfcc_workspace_bytes_ex), then the context gather + CMVN kernel
  float* rawf = (float*)((char*)ws + ((stats_bytes(plan, n) + 255) & ~(size_t)255));
  Params p = plan->p;
  p.mean_norm = p.var_norm = 0;
  mfcc_kernel<<<grid, THREADS, plan->smem_bytes, (cudaStream_t)stream>>>(p, pcm, offsets, n, t_max, rawf, out_len, 0, counters,
                                                                          stats);
  ASR_LAUNCH_CHECK();
  const int Fx = plan->p.feat_dim * (1 + 2 * ctx);
  mfcc_context_kernel<<<dim3(Fx, n), 256, 0, (cudaStream_t)stream>>>(rawf, out_len, n, t_max, plan->p.feat_dim, ctx,
                                                                      plan->p.mean_norm, plan->p.var_norm, plan->p.eps, out,
                                                                      time_major);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_mfcc_forward_host(const asr_mfcc_plan* plan, const float* pcm_host, int64_t num_samples,
                                         float* out_host) {
  ASR_CHECK_ARG(plan && pcm_host && out_host && num_samples > 1, "asr_mfcc_forward_host: bad argument");
  const int T = asr_mfcc_num_frames(plan, num_samples), F = asr_mfcc_num_feats(plan);
  const size_t wsb = asr_mfcc_workspace_bytes_ex(plan, 1, T);
  char* dev = nullptr;
  const size_t o_off = ((size_t)num_samples * 4 + 255) & ~(size_t)255, o_out = o_off + 256,
               o_len = o_out + (((size_t)T * F * 4 + 255) & ~(size_t)255), o_ws = o_len + 256, total = o_ws + wsb;
  ASR_CUDA(cudaMalloc(&dev, total));
  int64_t offs[2] = {0, num_samples};
  cudaError_t e = cudaMemcpy(dev, pcm_host, (size_t)num_samples * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(dev + o_off, offs, sizeof(offs), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(dev + o_ws, 0, wsb);
  int32_t rc = ASR_OK;
  if (e == cudaSuccess)
    rc = asr_mfcc_forward(plan, (const float*)dev, (const int64_t*)(dev + o_off), 1, T, (float*)(dev + o_out),
                          (int32_t*)(dev + o_len), 0, dev + o_ws, nullptr);
  if (e == cudaSuccess && rc == ASR_OK) e = cudaMemcpy(out_host, dev + o_out, (size_t)T * F * 4, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) {
    asr::set_error("asr_mfcc_forward_host: %s", cudaGetErrorString(e));
    return ASR_ERR_CUDA;
  }
  return rc;
}
