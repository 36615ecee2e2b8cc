// precision / layout helpers between the hot-path kernels (sm_100a).
#include "common.cuh"

namespace {

template <typename T> __device__ __forceinline__ T cvt(float v);
template <> __device__ __forceinline__ __half cvt<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// LO = true writes the rounding residual cvt(v - float(cvt(v))): with it a 16-bit operand pair (hi, lo) carries
// ~22 mantissa bits, and hi*hi + hi*lo + lo*hi on the tensor cores reproduces the fp32 product to ~1e-6
// (used where layer normalisation amplifies operand rounding, see engine._forward_general)
template <typename T, bool LO> __device__ __forceinline__ T cvt2(float v) {
  if (LO) return cvt<T>(v - (float)cvt<T>(v));
  return cvt<T>(v);
}

template <typename T, bool LO = false>
__global__ void __launch_bounds__(256)
cast_rows_kernel(const float* __restrict__ src, int64_t ld_src, T* __restrict__ dst, int64_t ld_dst, int64_t rows,
                 int cols) {
  // writes cols columns per row and zero-fills up to the next multiple of 8 (the 16-byte K padding
  // the GEMM operands need); never touches anything beyond that, so dst may be a column sub-block.
  const int cp = min((int64_t)((cols + 7) / 8 * 8), ld_dst);
  const int64_t total = rows * cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cp;
    const int c = (int)(i - r * cp);
    dst[r * ld_dst + c] = cvt2<T, LO>(c < cols ? src[r * ld_src + c] : 0.0f);
  }
}

// 32x32 tiles through shared memory: coalesced on both sides
template <typename T, bool LO = false>
__global__ void __launch_bounds__(256)
cast_transpose_kernel(const float* __restrict__ src, int64_t ld_src, T* __restrict__ dst, int64_t ld_dst, int64_t rows,
                      int cols) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = r0 + k;
    const int c = c0 + tx;
    tile[k][tx] = (r < rows && c < cols) ? src[r * ld_src + c] : 0.0f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k;
    const int64_t r = r0 + tx;
    if (c < cols && r < rows) dst[(int64_t)c * ld_dst + r] = cvt2<T, LO>(tile[tx][k]);
  }
}

// every 16-bit operand copy of a step in ONE launch: up to CAST_JOBS jobs (asr_cast_rows / asr_cast_transpose semantics
// each), the grid = the 32 x 32 tiles of all jobs back to back.  The ~20 per-tensor launches this replaces took 0.4 ms of
// small-kernel time beside the first recurrence and slowed it by 0.15 ms.
constexpr int CAST_JOBS = 32;
struct CastBatch {
  asr_cast_job job[CAST_JOBS];
  int tile0[CAST_JOBS + 1];       // first tile of job i in the grid
  int n;
};

template <typename T, bool LO>
__device__ __forceinline__ void cast_tile(const asr_cast_job& j, int64_t r0, int c0, float (*tile)[33]) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  T* dst = reinterpret_cast<T*>(j.dst16);
  if (j.transpose == 1) {
    for (int k = ty; k < 32; k += 8) {
      const int64_t r = r0 + k;
      const int c = c0 + tx;
      tile[k][tx] = (r < j.rows && c < j.cols) ? j.src[r * j.ld_src + c] : 0.0f;
    }
    __syncthreads();
    for (int k = ty; k < 32; k += 8) {
      const int c = c0 + k;
      const int64_t r = r0 + tx;
      if (c < j.cols && r < j.rows) dst[(int64_t)c * j.ld_dst + r] = cvt2<T, LO>(tile[tx][k]);
    }
  } else {
    // mode 0 zero-fills the K padding up to the next multiple of 8 like asr_cast_rows; mode 2 writes exactly cols columns
    // (a column sub-block whose right neighbour is another job of the same launch: the fill would race with its data)
    const int cp = j.transpose == 2 ? j.cols : (int)min((int64_t)((j.cols + 7) / 8 * 8), j.ld_dst);
    for (int k = ty; k < 32; k += 8) {
      const int64_t r = r0 + k;
      const int c = c0 + tx;
      if (r < j.rows && c < cp) dst[r * j.ld_dst + c] = cvt2<T, LO>(c < j.cols ? j.src[r * j.ld_src + c] : 0.0f);
    }
  }
}

__global__ void __launch_bounds__(256) cast_batch_kernel(const __grid_constant__ CastBatch b) {
  __shared__ float tile[32][33];
  int i = 0;
  while (i + 1 < b.n && (int)blockIdx.x >= b.tile0[i + 1]) ++i;
  const asr_cast_job& j = b.job[i];
  const int t = blockIdx.x - b.tile0[i];
  const int tiles_c = (((j.cols + 7) / 8 * 8) + 31) / 32;
  const int64_t r0 = (int64_t)(t / tiles_c) * 32;
  const int c0 = (t % tiles_c) * 32;
  if (j.dtype == 16) cast_tile<__half, true>(j, r0, c0, tile);
  else if (j.dtype == 0) cast_tile<__half, false>(j, r0, c0, tile);
  else cast_tile<__nv_bfloat16, false>(j, r0, c0, tile);
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ src, int64_t ld, int64_t rows, int cols, float* __restrict__ out) {
  // each CTA: 32 columns x a slab of rows; 8 warps stride the rows; atomics to out (pre-zeroed)
  __shared__ float part[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + tx;
  const int64_t slab = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t rb = blockIdx.y * slab, re = min(rows, rb + slab);
  float acc = 0.0f;
  if (c < cols)
    for (int64_t r = rb + ty; r < re; r += 8) acc += src[r * ld + c];
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && c < cols) {
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k][tx];
    atomicAdd(out + c, s);
  }
}

template <typename S> __device__ __forceinline__ float ldf(const S* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <typename S, typename T>
__global__ void __launch_bounds__(256)
mask_rows_kernel(const S* __restrict__ src, int64_t ld_src, const float* __restrict__ mask, int nb, T* __restrict__ dst,
                 int64_t ld_dst, int64_t rows, int cols) {
  const int cp = min((int64_t)((cols + 7) / 8 * 8), ld_dst);
  const int64_t total = rows * cp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cp;
    const int c = (int)(i - r * cp);
    float v = 0.0f;
    if (c < cols) v = ldf<S>(src + r * ld_src + c) * mask[(int64_t)(r % nb) * cols + c];
    dst[r * ld_dst + c] = cvt<T>(v);
  }
}

// 16-bit source, 16-bit destination, cols % 8 == 0, 16-byte aligned rows: 8 elements per thread
template <typename S, typename T>
__global__ void __launch_bounds__(256)
mask_rows_vec8_kernel(const S* __restrict__ src, int64_t ld_src, const float* __restrict__ mask, int nb,
                      T* __restrict__ dst, int64_t ld_dst, int rows, int cols) {
  const int c8n = cols >> 3;
  const int total = rows * c8n;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / c8n, c = (i - r * c8n) << 3;
    const uint4 v = *reinterpret_cast<const uint4*>(src + (int64_t)r * ld_src + c);
    const float4 m0 = *reinterpret_cast<const float4*>(mask + (int64_t)(r % nb) * cols + c);
    const float4 m1 = *reinterpret_cast<const float4*>(mask + (int64_t)(r % nb) * cols + c + 4);
    const S* e = reinterpret_cast<const S*>(&v);
    const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
    uint4 o;
    T* oe = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int k = 0; k < 8; ++k) oe[k] = cvt<T>(ldf<S>(e + k) * mm[k]);
    *reinterpret_cast<uint4*>(dst + (int64_t)r * ld_dst + c) = o;
  }
}

template <typename S, typename T>
__global__ void __launch_bounds__(256)
mask_transpose_kernel(const S* __restrict__ src, int64_t ld_src, const float* __restrict__ mask, int nb,
                      T* __restrict__ dst, int64_t ld_dst, int64_t rows, int cols) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = r0 + k;
    const int c = c0 + tx;
    tile[k][tx] = (r < rows && c < cols) ? ldf<S>(src + r * ld_src + c) * mask[(int64_t)(r % nb) * cols + c] : 0.0f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int c = c0 + k;
    const int64_t r = r0 + tx;
    if (c < cols && r < rows) dst[(int64_t)c * ld_dst + r] = cvt<T>(tile[tx][k]);
  }
}

__global__ void __launch_bounds__(256)
mask_combine_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ ma,
                    const float* __restrict__ mb, int nb, float* __restrict__ out, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int64_t mi = (r % nb) * cols + (i - r * cols);
    out[i] = a[i] * ma[mi] + b[i] * mb[mi];
  }
}

inline int grid_1d(int64_t total) {
  const int64_t b = (total + 255) / 256;
  return (int)(b < 148 * 8 ? (b < 1 ? 1 : b) : 148 * 8);
}

}  // namespace

extern "C" int32_t asr_cast_rows(const float* src, int64_t ld_src, void* dst16, int64_t ld_dst, int64_t rows,
                                 int32_t cols, int32_t dtype, void* stream) {
  ASR_CHECK_ARG(src && dst16 && rows > 0 && cols > 0 && ld_dst >= cols && ld_src >= cols, "asr_cast_rows: bad argument");
  const int grid = grid_1d(rows * ((cols + 7) / 8 * 8));
  if (dtype == 16)          // fp16 rounding residual (the "lo" half of a split-precision operand)
    cast_rows_kernel<__half, true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__half*)dst16, ld_dst, rows, cols);
  else if (dtype == 0)
    cast_rows_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__half*)dst16, ld_dst, rows, cols);
  else
    cast_rows_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst16, ld_dst,
                                                                            rows, cols);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_cast_batch(const asr_cast_job* jobs, int32_t n_jobs, void* stream) {
  ASR_CHECK_ARG(jobs && n_jobs >= 1, "asr_cast_batch: no jobs");
  for (int base = 0; base < n_jobs; base += CAST_JOBS) {
    CastBatch b;
    b.n = n_jobs - base < CAST_JOBS ? n_jobs - base : CAST_JOBS;
    int tiles = 0;
    for (int i = 0; i < b.n; ++i) {
      const asr_cast_job& j = jobs[base + i];
      ASR_CHECK_ARG(j.src && j.dst16 && j.rows > 0 && j.cols > 0 && j.ld_src >= j.cols &&
                        (j.transpose == 1 ? j.ld_dst >= j.rows : j.ld_dst >= j.cols) && (j.dtype == 0 || j.dtype == 1 || j.dtype == 16) &&
                        j.transpose >= 0 && j.transpose <= 2,
                    "asr_cast_batch: bad job %d", base + i);
      b.job[i] = j;
      b.tile0[i] = tiles;
      const int64_t t = ((j.rows + 31) / 32) * ((((j.cols + 7) / 8 * 8) + 31) / 32);
      ASR_CHECK_ARG(t < (1 << 24), "asr_cast_batch: job %d is too large", base + i);
      tiles += (int)t;
    }
    b.tile0[b.n] = tiles;
    cast_batch_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>(b);
    ASR_LAUNCH_CHECK();
  }
  return ASR_OK;
}

extern "C" int32_t asr_cast_transpose(const float* src, int64_t ld_src, void* dst16, int64_t ld_dst, int64_t rows,
                                      int32_t cols, int32_t dtype, void* stream) {
  ASR_CHECK_ARG(src && dst16 && rows > 0 && cols > 0 && ld_dst >= rows && ld_src >= cols,
                "asr_cast_transpose: bad argument");
  dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  ASR_CHECK_ARG(grid.y <= 65535, "asr_cast_transpose: too many columns");
  if (dtype == 16)
    cast_transpose_kernel<__half, true><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__half*)dst16, ld_dst, rows, cols);
  else if (dtype == 0)
    cast_transpose_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__half*)dst16, ld_dst, rows, cols);
  else
    cast_transpose_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, ld_src, (__nv_bfloat16*)dst16,
                                                                                 ld_dst, rows, cols);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

template <typename S>
static int32_t mask_cast_dispatch(const S* src, int64_t ld_src, const float* mask, int nb, void* dst16, int dtype,
                                  int64_t ld_dst, int64_t rows, int cols, int transpose, cudaStream_t st) {
  if (transpose) {
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    if (dtype == 0) mask_transpose_kernel<S, __half><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__half*)dst16, ld_dst, rows, cols);
    else mask_transpose_kernel<S, __nv_bfloat16><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__nv_bfloat16*)dst16, ld_dst, rows, cols);
  } else if (sizeof(S) == 2 && cols % 8 == 0 && ld_src % 8 == 0 && ld_dst % 8 == 0 && rows * (int64_t)(cols / 8) < (1LL << 31) &&
             ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst16) | reinterpret_cast<uintptr_t>(mask)) & 15) == 0) {
    const int grid = grid_1d(rows * (int64_t)(cols / 8));
    if (dtype == 0) mask_rows_vec8_kernel<S, __half><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__half*)dst16, ld_dst, (int)rows, cols);
    else mask_rows_vec8_kernel<S, __nv_bfloat16><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__nv_bfloat16*)dst16, ld_dst, (int)rows, cols);
  } else {
    const int grid = grid_1d(rows * ((cols + 7) / 8 * 8));
    if (dtype == 0) mask_rows_kernel<S, __half><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__half*)dst16, ld_dst, rows, cols);
    else mask_rows_kernel<S, __nv_bfloat16><<<grid, 256, 0, st>>>(src, ld_src, mask, nb, (__nv_bfloat16*)dst16, ld_dst, rows, cols);
  }
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_mask_cast(const void* src, int32_t src_dtype, int64_t ld_src, const float* mask, int32_t n_batch,
                                 void* dst16, int32_t dtype, int64_t ld_dst, int64_t rows, int32_t cols,
                                 int32_t transpose, void* stream) {
  ASR_CHECK_ARG(src && mask && dst16 && rows > 0 && cols > 0 && n_batch > 0 && ld_src >= cols, "asr_mask_cast: bad argument");
  ASR_CHECK_ARG(transpose ? ld_dst >= rows : ld_dst >= cols, "asr_mask_cast: ld_dst too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (src_dtype == 0) return mask_cast_dispatch((const __half*)src, ld_src, mask, n_batch, dst16, dtype, ld_dst, rows, cols, transpose, st);
  if (src_dtype == 1) return mask_cast_dispatch((const __nv_bfloat16*)src, ld_src, mask, n_batch, dst16, dtype, ld_dst, rows, cols, transpose, st);
  return mask_cast_dispatch((const float*)src, ld_src, mask, n_batch, dst16, dtype, ld_dst, rows, cols, transpose, st);
}

extern "C" int32_t asr_mask_combine(const float* a, const float* b, const float* mask_a, const float* mask_b,
                                    int32_t n_batch, float* out, int64_t rows, int32_t cols, void* stream) {
  ASR_CHECK_ARG(a && b && mask_a && mask_b && out && rows > 0 && cols > 0 && n_batch > 0, "asr_mask_combine: bad argument");
  mask_combine_kernel<<<grid_1d(rows * cols), 256, 0, (cudaStream_t)stream>>>(a, b, mask_a, mask_b, n_batch, out, rows, cols);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

extern "C" int32_t asr_colsum(const float* src, int64_t ld, int64_t rows, int32_t cols, float* out, void* stream) {
  ASR_CHECK_ARG(src && out && rows > 0 && cols > 0 && ld >= cols, "asr_colsum: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  ASR_CUDA(cudaMemsetAsync(out, 0, (size_t)cols * sizeof(float), st));
  int gy = (int)((rows + 511) / 512);
  if (gy > 296) gy = 296;
  dim3 grid((cols + 31) / 32, gy);
  colsum_kernel<<<grid, 256, 0, st>>>(src, ld, rows, cols, out);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

// out[r, c] = (a[r, c] + (b ? b[r, c] : 0)) * (mask ? mask[r % n_batch, c] : 1): the residual merge(mode='sum')
// of core/models.py:273-274, the element-wise input Dropout of :257-258 (n_batch = rows: one mask entry per
// element) and its backward.  out may alias a.
__global__ void add_mask_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ mask,
                                int n_batch, float* __restrict__ out, int64_t rows, int cols) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols;
    const int c = (int)(i - r * cols);
    float v = a[i];
    if (b) v += b[i];
    if (mask) v *= mask[(r % n_batch) * cols + c];
    out[i] = v;
  }
}

extern "C" int32_t asr_add_mask(const float* a, const float* b, const float* mask, int64_t n_batch, float* out,
                                int64_t rows, int32_t cols, void* stream) {
  ASR_CHECK_ARG(a && out && rows > 0 && cols > 0 && (!mask || n_batch > 0), "asr_add_mask: bad argument");
  ASR_CHECK_ARG(n_batch <= 2147483647LL, "asr_add_mask: n_batch too large");
  add_mask_kernel<<<grid_1d(rows * cols), 256, 0, (cudaStream_t)stream>>>(a, b, mask, (int)(mask ? n_batch : 1), out, rows, cols);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

// Bernoulli keep masks scaled by 1/(1-p) (K.dropout, Keras-1): out[i] = u_i >= p ? 1/(1-p) : 0 with u_i from a
// counter-based generator (splitmix64 of seed and the element counter), so ONE launch fills every mask of a step
__global__ void dropout_mask_kernel(float* __restrict__ out, int64_t n, float p, float scale, uint64_t seed, uint64_t offset) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(offset + (uint64_t)i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);      // 24 random bits -> [0, 1)
    out[i] = u >= p ? scale : 0.0f;
  }
}

extern "C" int32_t asr_dropout_mask(float* out, int64_t n, float p, uint64_t seed, uint64_t offset, void* stream) {
  ASR_CHECK_ARG(out && n > 0 && p >= 0.0f && p < 1.0f, "asr_dropout_mask: bad argument");
  dropout_mask_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(out, n, p, 1.0f / (1.0f - p), seed, offset);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

// same generator, caller-chosen keep value: zoneout keep masks are 0 / 1 (core/layers_utils.py:34-42 multiplies the
// K.dropout result by (1 - level) again), the element-wise input Dropout of core/models.py:257-258 is 0 / 1/(1-p)
extern "C" int32_t asr_bernoulli_mask(float* out, int64_t n, float p, float keep_value, uint64_t seed, uint64_t offset,
                                      void* stream) {
  ASR_CHECK_ARG(out && n > 0 && p >= 0.0f && p < 1.0f, "asr_bernoulli_mask: bad argument");
  dropout_mask_kernel<<<grid_1d(n), 256, 0, (cudaStream_t)stream>>>(out, n, p, keep_value, seed, offset);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}

// GaussianNoise(std) of core/models.py:67,251 (train phase): x[r, c] += std * n, n ~ N(0, 1) by Box-Muller on two
// 24-bit uniforms of the counter-based generator above
__global__ void gaussian_noise_kernel(float* __restrict__ x, int64_t rows, int cols, int64_t ld, float stdv, uint64_t seed,
                                      uint64_t offset) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(offset + (uint64_t)i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u1 = ((float)(uint32_t)(z >> 40) + 1.0f) * (1.0f / 16777216.0f);      // (0, 1]
    const float u2 = (float)(uint32_t)((z >> 16) & 0xFFFFFFu) * (1.0f / 16777216.0f);
    const float g = sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
    const int64_t r = i / cols;
    x[r * ld + (i - r * cols)] += stdv * g;
  }
}

extern "C" int32_t asr_add_gaussian_noise(float* x, int64_t rows, int32_t cols, int64_t ld, float stdv, uint64_t seed,
                                          uint64_t offset, void* stream) {
  ASR_CHECK_ARG(x && rows > 0 && cols > 0 && ld >= cols && stdv >= 0.0f, "asr_add_gaussian_noise: bad argument");
  gaussian_noise_kernel<<<grid_1d(rows * cols), 256, 0, (cudaStream_t)stream>>>(x, rows, cols, ld, stdv, seed, offset);
  ASR_LAUNCH_CHECK();
  return ASR_OK;
}
