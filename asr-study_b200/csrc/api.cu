// C-ABI dispatch for the GEMM and BiLSTM entry points (see include/asr_b200.h).
#include "common.cuh"

namespace gemm_tc {
bool supports(int dtype_in, int dtype_out, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc);
int32_t run(int dtype_in, int dtype_out, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
            void* C, int64_t ldc, const float* bias, float alpha, int accumulate, cudaStream_t st);
}
namespace gemm_tc2 {
bool supports(int dtype_in, int dtype_out, int M, int N, int K, int64_t lda, int64_t ldb, int64_t ldc);
int32_t run(int dtype_in, int dtype_out, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
            void* C, int64_t ldc, const float* bias, float alpha, int accumulate, cudaStream_t st);
}
namespace lstm32 {
size_t scratch_bytes(int H);
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st);
}
namespace lstmcell {
int32_t forward(const asr_lstm_fwd_args* a, const asr_lstm_variant* v, float* uh_raw, cudaStream_t st);
int32_t backward(const asr_lstm_bwd_args* a, const asr_lstm_variant* v, const float* zx, const float* uh_raw, float* duh,
                 const asr_lstm_variant_grads* g, cudaStream_t st);
}
namespace lstmtc2 {
bool shape_supported(int T, int N, int H, int opts);
bool supports_fwd(const asr_lstm_fwd_args* a);
bool supports_bwd(const asr_lstm_bwd_args* a);
int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st);
size_t scratch_bytes(int H);
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
}
namespace lstmtc4 {
bool shape_supported(int T, int N, int H, int opts);
bool supports_fwd(const asr_lstm_fwd_args* a);
bool supports_bwd(const asr_lstm_bwd_args* a);
int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st);
size_t scratch_bytes();
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
}

// Engine selection is explicit (flag arguments), never an environment variable, and never a fallback to another
// instruction set: both GEMM engines and both recurrence engines are tcgen05; the fp32 CUDA-core recurrence serves the
// widths below one tensor-core tile (H <= 128 of graves2006) and is the exact-arithmetic parity reference.
extern "C" int32_t asr_gemm_tn_ex(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N, int32_t K, const void* A,
                                  int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                                  float alpha, int32_t accumulate, int32_t flags, void* stream) {
  ASR_CHECK_ARG(A && B && C, "asr_gemm_tn: null operand");
  ASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "asr_gemm_tn: bad shape %dx%dx%d", M, N, K);
  ASR_CHECK_ARG(dtype_in >= 0 && dtype_in <= 1 && dtype_out >= 0 && dtype_out <= 2, "asr_gemm_tn: bad dtype");
  // lda < K is allowed: rows of A then overlap (row m starts lda elements after row m-1), a Toeplitz view that turns a
  // convolution over the leading axis into a plain GEMM without an im2col copy (engine.py:_conv_forward)
  ASR_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= 8 && ldb >= K && ldc >= N,
                "asr_gemm_tn: K, lda, ldb must be multiples of 8 (16-byte rows)");
  ASR_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "asr_gemm_tn: operands must be 16-byte aligned");
  ASR_CHECK_ARG(!accumulate || dtype_out == 0, "asr_gemm_tn: accumulate needs fp32 C");
  cudaStream_t st = (cudaStream_t)stream;
  // persistent 128x256 engine for the large regular projections; a GEMM that is meant to run BESIDE a persistent
  // recurrence (ASR_GEMM_BACKGROUND: only the SMs that kernel leaves idle are free) keeps the non-persistent tiling,
  // whose CTAs are scheduled one by one as SMs free up
  if (!(flags & (ASR_GEMM_BACKGROUND | ASR_GEMM_TILE128)) && gemm_tc2::supports(dtype_in, dtype_out, M, N, K, lda, ldb, ldc))
    return gemm_tc2::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
  if (gemm_tc::supports(dtype_in, dtype_out, M, N, K, lda, ldb, ldc))
    return gemm_tc::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
  asr::set_error("asr_gemm_tn: no tcgen05 engine takes M=%d N=%d K=%d lda=%lld ldb=%lld", M, N, K, (long long)lda,
                 (long long)ldb);
  return ASR_ERR_UNSUPPORTED;
}

extern "C" int32_t asr_gemm_tn(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N, int32_t K, const void* A,
                               int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                               float alpha, int32_t accumulate, void* stream) {
  return asr_gemm_tn_ex(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, 0, stream);
}

extern "C" size_t asr_lstm_flags_bytes(void) {
  // sized for the largest supported hidden size of any engine
  size_t m = lstm32::scratch_bytes(1024);
  if (lstmtc2::scratch_bytes(1024) > m) m = lstmtc2::scratch_bytes(1024);
  if (lstmtc4::scratch_bytes() > m) m = lstmtc4::scratch_bytes();
  return m;
}

static int32_t check_common(int T, int N, int H) {
  ASR_CHECK_ARG(T >= 1 && N >= 1 && H >= 1 && H <= 1024, "lstm: bad shape T=%d N=%d H=%d", T, N, H);
  return ASR_OK;
}

// 1 when one of the persistent engines (tensor-core or fp32) takes this shape; 0 = only the general-cell engine
// (asr_lstm_cell_forward / backward: any N, H <= 1024) does
extern "C" int32_t asr_lstm_persistent_supported(int32_t T, int32_t N, int32_t H, int32_t training, int32_t opts) {
  if (T < 1 || N < 1 || H < 1) return 0;
  const bool fp32_ok = N <= 32 && 2 * ((H + 7) / 8) <= 148 &&
                       ((size_t)H * 4 * 8 + (size_t)H * 32 + (size_t)8 * 32 * 4 * 8) * sizeof(float) <= 227 * 1024;
  if (fp32_ok) return 1;
  if (opts & ASR_LSTM_PIN_FP32) return 0;
  (void)training;
  return lstmtc2::shape_supported(T, N, H, opts) ? 1 : 0;
}

// the fused dropout fields are implemented by the tensor-core engines only
extern "C" int32_t asr_lstm_fuses_masks(int32_t T, int32_t N, int32_t H, int32_t opts) {
  if (opts & ASR_LSTM_PIN_FP32) return 0;
  return lstmtc2::shape_supported(T, N, H, opts) ? 1 : 0;
}

extern "C" int32_t asr_lstm_fuses_variants(int32_t T, int32_t N, int32_t H, int32_t opts) {
  return asr_lstm_fuses_masks(T, N, H, opts);
}

extern "C" int32_t asr_lstm_fp16_storage(int32_t T, int32_t N, int32_t H, int32_t opts) {
  if (opts & ASR_LSTM_PIN_FP32) return 0;
  return lstmtc4::shape_supported(T, N, H, opts) ? 1 : 0;
}

extern "C" int32_t asr_lstm_forward(const asr_lstm_fwd_args* a, void* stream) {
  ASR_CHECK_ARG(a && (a->zx || a->zx16) && a->bias && a->flags, "asr_lstm_forward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  ASR_CHECK_ARG(!a->training || (a->gates && a->cell) || (a->gates16 && a->cell16), "asr_lstm_forward: training needs gates/cell buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = a->mask_next || a->hm16 || a->hmT16 || a->hT16u || a->mi || a->zoneout > 0.0f;
  ASR_CHECK_ARG(!a->mask_next == !a->hm16, "asr_lstm_forward: mask_next and hm16 go together");
  ASR_CHECK_ARG(a->zoneout >= 0.0f && a->zoneout < 1.0f, "asr_lstm_forward: zoneout must be in [0, 1)");
  if (!(a->opts & ASR_LSTM_PIN_FP32)) {
    if (a->zx16) {
      ASR_CHECK_ARG(lstmtc4::supports_fwd(a), "asr_lstm_forward: zx16 needs the engine asr_lstm_fp16_storage() reports "
                    "(no mi / zoneout / h32, gates16 + cell16 when training)");
      return lstmtc4::forward(a, st);
    }
    if (lstmtc2::supports_fwd(a)) return lstmtc2::forward(a, st);      // LL ring through L2, fp32 storage
  }
  ASR_CHECK_ARG(!fused, "asr_lstm_forward: fused dropout outputs need the engine asr_lstm_fuses_masks() reports");
  ASR_CHECK_ARG(a->zx && a->U && a->h16 && (!a->training || (a->gates && a->cell)), "asr_lstm_forward: fp32 engine needs zx, U, h16 (and gates / cell)");
  return lstm32::forward(a, st);
}

extern "C" int32_t asr_lstm_backward(const asr_lstm_bwd_args* a, void* stream) {
  ASR_CHECK_ARG(a && a->dh && ((a->gates && a->cell) || (a->gates16 && a->cell16)) && a->dbias && a->flags, "asr_lstm_backward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = a->dh2 || a->mask_dh || a->mi || a->zoneout > 0.0f;
  ASR_CHECK_ARG(!a->dh2 || a->mask_dh, "asr_lstm_backward: dh2 needs mask_dh");
  if (!(a->opts & ASR_LSTM_PIN_FP32)) {
    if (a->gates16) {
      ASR_CHECK_ARG(lstmtc4::supports_bwd(a), "asr_lstm_backward: gates16 needs the engine asr_lstm_fp16_storage() reports");
      return lstmtc4::backward(a, st);
    }
    if (lstmtc2::supports_bwd(a)) return lstmtc2::backward(a, st);
  }
  ASR_CHECK_ARG(!fused, "asr_lstm_backward: fused dropout inputs need the engine asr_lstm_fuses_masks() reports");
  ASR_CHECK_ARG(a->U && a->gates && a->cell, "asr_lstm_backward: fp32 engine needs U, gates, cell");
  return lstm32::backward(a, st);
}

extern "C" int32_t asr_lstm_cell_forward(const asr_lstm_fwd_args* a, const asr_lstm_variant* v, float* uh_raw, void* stream) {
  ASR_CHECK_ARG(a && a->zx && a->bias && (a->h32 || a->h16), "asr_lstm_cell_forward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  ASR_CHECK_ARG(!a->training || (a->gates && a->cell), "asr_lstm_cell_forward: training needs gates/cell buffers");
  return lstmcell::forward(a, v, uh_raw, (cudaStream_t)stream);
}

extern "C" int32_t asr_lstm_cell_backward(const asr_lstm_bwd_args* a, const asr_lstm_variant* v, const float* zx,
                                          const float* uh_raw, float* duh, const asr_lstm_variant_grads* g, void* stream) {
  ASR_CHECK_ARG(a && a->dh && a->gates && a->cell && a->dbias, "asr_lstm_cell_backward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  return lstmcell::backward(a, v, zx, uh_raw, duh, g, (cudaStream_t)stream);
}
