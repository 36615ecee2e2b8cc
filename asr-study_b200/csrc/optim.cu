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

__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ g, const float* __restrict__ p, const uint8_t* __restrict__ mask, int64_t n,
              float gs, float wd, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
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
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
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
  int64_t b = (n + 255) / 256;
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
