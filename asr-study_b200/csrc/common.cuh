// Shared host/device helpers for libasr_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/asr_b200.h"

namespace asr {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define ASR_CHECK_ARG(cond, ...)                       \
  do {                                                 \
    if (!(cond)) {                                     \
      asr::set_error(__VA_ARGS__);                     \
      return ASR_ERR_INVALID;                          \
    }                                                  \
  } while (0)

#define ASR_CUDA(call)                                                        \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) {                                                  \
      asr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,            \
                     cudaGetErrorString(_e));                                 \
      return ASR_ERR_CUDA;                                                    \
    }                                                                         \
  } while (0)

#define ASR_LAUNCH_CHECK()                                                    \
  do {                                                                        \
    cudaError_t _e = cudaGetLastError();                                      \
    if (_e != cudaSuccess) {                                                  \
      asr::set_error("%s:%d: kernel launch -> %s", __FILE__, __LINE__,        \
                     cudaGetErrorString(_e));                                 \
      return ASR_ERR_CUDA;                                                    \
    }                                                                         \
    asr::count_launch();                                                      \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// tanh through one ex2.approx and one fast divide: |abs error| ~1e-6 (the libdevice tanhf costs ~10x more
// instructions and sits on the recurrence's critical path twice per step)
__device__ __forceinline__ float tanh_fast(float x) {
  const float e = __expf(2.0f * x);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

// Keras-1 hard_sigmoid: clip(0.2x + 0.5, 0, 1)
__device__ __forceinline__ float hard_sigmoid(float x) {
  return fminf(fmaxf(fmaf(0.2f, x, 0.5f), 0.0f), 1.0f);
}
// derivative expressed on the activated value
__device__ __forceinline__ float hard_sigmoid_grad(float a) {
  return (a > 0.0f && a < 1.0f) ? 0.2f : 0.0f;
}

}  // namespace asr
