// K9 — global-norm clip + Adam / SGD-momentum on one flat fp32 bucket (sm_100a).
// Replaces Keras-1.2.2 Adam(lr, clipnorm) / SGD(lr, momentum, clipnorm) as
// configured at train.py:133-137; the l2(weight_decay) regularisers of
// core/models.py:263-264,279 are folded in as g += 2*wd*p on masked elements
// BEFORE the norm, like Keras (gradients of the total loss are clipped).
#include "common.cuh"

namespace {

__device__ __forceinline__ float eff_grad(const float* g, const float* p, const uint8_t* mask, int64_t i, float gs,
                                          float wd) {
  float v = gs * g[i];
  if (mask && mask[i]) v = fmaf(2.0f * wd, p[i], v);
  return v;
}

// 16-byte accesses over the body of the bucket (the bucket and every tensor in it are 16-byte aligned), scalar tail
__device__ __forceinline__ float4 eff_grad4(const float* g, const float* p, const uint8_t* mask, int64_t q, float gs,
                                            float wd) {
  const float4 gv = __ldg(reinterpret_cast<const float4*>(g) + q);
  float4 r = make_float4(gs * gv.x, gs * gv.y, gs * gv.z, gs * gv.w);
  if (mask) {
    const uchar4 mk = __ldg(reinterpret_cast<const uchar4*>(mask) + q);
    if (mk.x | mk.y | mk.z | mk.w) {
      const float4 pv = reinterpret_cast<const float4*>(p)[q];
      if (mk.x) r.x = fmaf(2.0f * wd, pv.x, r.x);
      if (mk.y) r.y = fmaf(2.0f * wd, pv.y, r.y);
      if (mk.z) r.z = fmaf(2.0f * wd, pv.z, r.z);
      if (mk.w) r.w = fmaf(2.0f * wd, pv.w, r.w);
    }
  }
  return r;
}
__device__ __forceinline__ bool aligned16(const void* a, const void* b, const void* c, const void* d, const void* e) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
           reinterpret_cast<uintptr_t>(d)) & 15) == 0 && (reinterpret_cast<uintptr_t>(e) & 3) == 0;
}

__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ p, const uint8_t* __restrict__ mask, int64_t n,
              float gs, float wd, double* __restrict__ out) {
  double acc = 0.0;
  const int64_t nq = aligned16(g, p, g, p, mask) ? n / 4 : 0;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = eff_grad4(g, p, mask, q, gs, wd);
    acc += ((double)v.x * (double)v.x + (double)v.y * (double)v.y) + ((double)v.z * (double)v.z + (double)v.w * (double)v.w);
  }
  for (int64_t i = 4 * nq + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = eff_grad(g, p, mask, i, gs, wd);
    acc += (double)v * (double)v;
  }
  acc = asr::warp_sum(acc);
  __shared__ double s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += s[k];
    atomicAdd(out, v);
  }
}

__device__ __forceinline__ float clip_scale(const double* sqnorm, float clipnorm) {
  if (clipnorm <= 0.0f || sqnorm == nullptr) return 1.0f;
  const double nrm = sqrt(*sqnorm);
  return (nrm >= (double)clipnorm) ? (float)((double)clipnorm / nrm) : 1.0f;   // Keras clip_norm
}

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const uint8_t* __restrict__ mask, int64_t n, float gs, float wd, const double* __restrict__ sqnorm,
            float clipnorm, float lr_t, float b1, float b2, float eps) {
  const float sc = clip_scale(sqnorm, clipnorm);
  const int64_t nq = aligned16(p, g, m, v, mask) ? n / 4 : 0;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    const float4 ge = eff_grad4(g, p, mask, q, gs, wd);
    float4 pv = reinterpret_cast<float4*>(p)[q], mv = reinterpret_cast<float4*>(m)[q], vv = reinterpret_cast<float4*>(v)[q];
    const float gi[4] = {ge.x * sc, ge.y * sc, ge.z * sc, ge.w * sc};
    float* pe = &pv.x; float* me = &mv.x; float* ve = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {                          // same arithmetic, element by element, as the scalar tail
      const float mi = b1 * me[k] + (1.0f - b1) * gi[k];
      const float vi = b2 * ve[k] + (1.0f - b2) * gi[k] * gi[k];
      me[k] = mi;
      ve[k] = vi;
      pe[k] = pe[k] - lr_t * mi / (sqrtf(vi) + eps);
    }
    reinterpret_cast<float4*>(m)[q] = mv;
    reinterpret_cast<float4*>(v)[q] = vv;
    reinterpret_cast<float4*>(p)[q] = pv;
  }
  for (int64_t i = 4 * nq + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = eff_grad(g, p, mask, i, gs, wd) * sc;
    const float mi = b1 * m[i] + (1.0f - b1) * gi;
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] - lr_t * mi / (sqrtf(vi) + eps);
  }
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ mom,
           const uint8_t* __restrict__ mask, int64_t n, float gs, float wd, const double* __restrict__ sqnorm,
           float clipnorm, float lr, float momentum) {
  const float sc = clip_scale(sqnorm, clipnorm);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = eff_grad(g, p, mask, i, gs, wd) * sc;
    const float vi = momentum * mom[i] - lr * gi;   // Keras-1 SGD: v = mom*m - lr*g ; p += v
    mom[i] = vi;
    p[i] = p[i] + vi;
  }
}

inline int grid_for(int64_t n) {
  int64_t b = (n / 4 + 255) / 256;                         // one 16-byte quad per thread and trip
  const int64_t cap = 148 * 8;   // 8 resident 256-thread CTAs per SM on 148 SMs
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

extern "C" int32_t asr_grad_sqnorm(const float* grad, const float* param, const uint8_t* decay_mask, int64_t n,
                                   float grad_scale, float weight_decay, double* sqnorm, void* stream) {
  ASR_CHECK_ARG(grad && param && sqnorm && n > 0, "asr_grad_sqnorm: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  ASR_CUDA(cudaMemsetAsync(sqnorm, 0, sizeof(double), st));
  sqnorm_kernel<<<grid_for(n), 256, 0, st>>>(grad, param, decay_mask, n, grad_scale, weight_decay, sqnorm);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_adam_step(float* param, const float* grad, float* m, float* v, const uint8_t* decay_mask,
                                 int64_t n, float grad_scale, float weight_decay, const double* sqnorm, float clipnorm,
                                 float lr, float beta1, float beta2, float eps, int32_t step, void* stream) {
  ASR_CHECK_ARG(param && grad && m && v && n > 0 && step >= 1, "asr_adam_step: bad argument");
  // Keras-1.2.2 Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, decay_mask, n, grad_scale, weight_decay,
                                                              sqnorm, clipnorm, (float)lr_t, beta1, beta2, eps);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_sgd_step(float* param, const float* grad, float* mom, const uint8_t* decay_mask, int64_t n,
                                float grad_scale, float weight_decay, const double* sqnorm, float clipnorm, float lr,
                                float momentum, void* stream) {
  ASR_CHECK_ARG(param && grad && mom && n > 0, "asr_sgd_step: bad argument");
  sgd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, mom, decay_mask, n, grad_scale, weight_decay,
                                                             sqnorm, clipnorm, lr, momentum);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
