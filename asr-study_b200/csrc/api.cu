// C-ABI dispatch for the GEMM and BiLSTM entry points (see include/asr_b200.h).
#include "common.cuh"
#include <stdlib.h>
#include <string.h>

namespace gemm_mma {
int32_t run(int dtype_in, int dtype_out, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
            void* C, int64_t ldc, const float* bias, float alpha, int accumulate, cudaStream_t st);
}
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
namespace lstmtc3 {
bool supports_fwd(const asr_lstm_fwd_args* a);
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
}
namespace lstmtc2 {
bool shape_supported(int T, int N, int H, bool bwd);
bool supports_fwd(const asr_lstm_fwd_args* a);
bool supports_bwd(const asr_lstm_bwd_args* a);
int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st);
size_t scratch_bytes(int H);
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
}
namespace lstmtc {
bool supports_fwd(const asr_lstm_fwd_args* a);
bool supports_bwd(const asr_lstm_bwd_args* a);
size_t scratch_bytes(int H);
int32_t forward(const asr_lstm_fwd_args* a, cudaStream_t st);
int32_t backward(const asr_lstm_bwd_args* a, cudaStream_t st);
}

// Engine selection is a *build/debug* switch, not a fallback: ASR_B200_GEMM=mma / ASR_B200_LSTM=fp32 pin
// the cross-check engines (used by the tests); default is the tcgen05 engines wherever they take the shape.
static int env_is(const char* name, const char* val) {
  const char* e = getenv(name);
  return e && strcmp(e, val) == 0;
}

extern "C" int32_t asr_gemm_tn_ex(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N, int32_t K, const void* A,
                                  int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                                  float alpha, int32_t accumulate, int32_t flags, void* stream) {
  ASR_CHECK_ARG(A && B && C, "asr_gemm_tn: null operand");
  ASR_CHECK_ARG(M > 0 && N > 0 && K > 0, "asr_gemm_tn: bad shape %dx%dx%d", M, N, K);
  ASR_CHECK_ARG(dtype_in >= 0 && dtype_in <= 1 && dtype_out >= 0 && dtype_out <= 2, "asr_gemm_tn: bad dtype");
  ASR_CHECK_ARG(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= K && ldb >= K && ldc >= N,
                "asr_gemm_tn: K, lda, ldb must be multiples of 8 (16-byte rows)");
  ASR_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "asr_gemm_tn: operands must be 16-byte aligned");
  ASR_CHECK_ARG(!accumulate || dtype_out == 0, "asr_gemm_tn: accumulate needs fp32 C");
  cudaStream_t st = (cudaStream_t)stream;
  if (env_is("ASR_B200_GEMM", "mma"))
    return gemm_mma::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
  // persistent 128x256 engine for the large regular projections; a GEMM that is meant to run BESIDE a persistent
  // recurrence (ASR_GEMM_BACKGROUND: only the SMs that kernel leaves idle are free) keeps the non-persistent tiling,
  // whose CTAs are scheduled one by one as SMs free up
  if (!(flags & ASR_GEMM_BACKGROUND) && !env_is("ASR_B200_GEMM", "tc1") &&
      gemm_tc2::supports(dtype_in, dtype_out, M, N, K, lda, ldb, ldc))
    return gemm_tc2::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
  if (gemm_tc::supports(dtype_in, dtype_out, M, N, K, lda, ldb, ldc))
    return gemm_tc::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
  return gemm_mma::run(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, st);
}

extern "C" int32_t asr_gemm_tn(int32_t dtype_in, int32_t dtype_out, int32_t M, int32_t N, int32_t K, const void* A,
                               int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                               float alpha, int32_t accumulate, void* stream) {
  return asr_gemm_tn_ex(dtype_in, dtype_out, M, N, K, A, lda, B, ldb, C, ldc, bias, alpha, accumulate, 0, stream);
}

extern "C" size_t asr_lstm_flags_bytes(void) {
  // sized for the largest supported hidden size of either engine
  size_t m = lstm32::scratch_bytes(1024);
  if (lstmtc::scratch_bytes(1024) > m) m = lstmtc::scratch_bytes(1024);
  if (lstmtc2::scratch_bytes(1024) > m) m = lstmtc2::scratch_bytes(1024);
  return m;
}

static int32_t check_common(int T, int N, int H) {
  ASR_CHECK_ARG(T >= 1 && N >= 1 && H >= 1 && H <= 1024, "lstm: bad shape T=%d N=%d H=%d", T, N, H);
  return ASR_OK;
}

// 1 when one of the persistent engines (tensor-core or fp32) takes this shape; 0 = only the general-cell engine
// (asr_lstm_cell_forward / backward: any N, H <= 1024) does, e.g. the 5 x BiLSTM-800 stack of BASELINE config 4
extern "C" int32_t asr_lstm_persistent_supported(int32_t T, int32_t N, int32_t H, int32_t training) {
  if (T < 1 || N < 1 || H < 1) return 0;
  const bool fp32_ok = N <= 32 && 2 * ((H + 7) / 8) <= 148 &&
                       ((size_t)H * 4 * 8 + (size_t)H * 32 + (size_t)8 * 32 * 4 * 8) * sizeof(float) <= 227 * 1024;
  if (fp32_ok) return 1;
  if (env_is("ASR_B200_LSTM", "fp32")) return 0;
  return lstmtc2::shape_supported(T, N, H, false) && (!training || lstmtc2::shape_supported(T, N, H, true)) ? 1 : 0;
}

// the fused dropout fields are implemented by the default tensor-core engine (lstm_tc2.cu) only
extern "C" int32_t asr_lstm_fuses_masks(int32_t T, int32_t N, int32_t H) {
  if (env_is("ASR_B200_LSTM", "fp32") || env_is("ASR_B200_LSTM", "tc1") || env_is("ASR_B200_LSTM", "tc3")) return 0;
  return lstmtc2::shape_supported(T, N, H, false) && lstmtc2::shape_supported(T, N, H, true) ? 1 : 0;
}

extern "C" int32_t asr_lstm_fuses_variants(int32_t T, int32_t N, int32_t H) {
  return asr_lstm_fuses_masks(T, N, H);
}

extern "C" int32_t asr_lstm_forward(const asr_lstm_fwd_args* a, void* stream) {
  ASR_CHECK_ARG(a && a->zx && a->bias && a->flags, "asr_lstm_forward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  ASR_CHECK_ARG(!a->training || (a->gates && a->cell), "asr_lstm_forward: training needs gates/cell buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = a->mask_next || a->hm16 || a->hmT16 || a->hT16u || a->mi || a->zoneout > 0.0f;
  ASR_CHECK_ARG(!a->mask_next == !a->hm16, "asr_lstm_forward: mask_next and hm16 go together");
  ASR_CHECK_ARG(a->zoneout >= 0.0f && a->zoneout < 1.0f, "asr_lstm_forward: zoneout must be in [0, 1)");
  if (!env_is("ASR_B200_LSTM", "fp32")) {
    const bool pin1 = env_is("ASR_B200_LSTM", "tc1"), pin3 = env_is("ASR_B200_LSTM", "tc3");
    if (pin3 && !fused && lstmtc3::supports_fwd(a)) return lstmtc3::forward(a, st);    // cluster / DSMEM exchange (same speed, kept selectable)
    if (!pin1 && !pin3 && lstmtc2::supports_fwd(a)) return lstmtc2::forward(a, st);    // LL ring through L2 (default)
    if (!fused && lstmtc::supports_fwd(a)) return lstmtc::forward(a, st);
  }
  ASR_CHECK_ARG(!fused, "asr_lstm_forward: fused dropout outputs need the engine asr_lstm_fuses_masks() reports");
  ASR_CHECK_ARG(a->U && a->h16, "asr_lstm_forward: fp32 engine needs U and h16");
  return lstm32::forward(a, st);
}

extern "C" int32_t asr_lstm_backward(const asr_lstm_bwd_args* a, void* stream) {
  ASR_CHECK_ARG(a && a->dh && a->gates && a->cell && a->dbias && a->flags, "asr_lstm_backward: null argument");
  if (int32_t rc = check_common(a->T, a->N, a->H)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const bool fused = a->dh2 || a->mask_dh || a->mi || a->zoneout > 0.0f;
  ASR_CHECK_ARG(!a->dh2 || a->mask_dh, "asr_lstm_backward: dh2 needs mask_dh");
  if (!env_is("ASR_B200_LSTM", "fp32")) {
    if (!env_is("ASR_B200_LSTM", "tc1") && lstmtc2::supports_bwd(a)) return lstmtc2::backward(a, st);
    if (!fused && lstmtc::supports_bwd(a)) return lstmtc::backward(a, st);
  }
  ASR_CHECK_ARG(!fused, "asr_lstm_backward: fused dropout inputs need the engine asr_lstm_fuses_masks() reports");
  ASR_CHECK_ARG(a->U, "asr_lstm_backward: fp32 engine needs U");
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
