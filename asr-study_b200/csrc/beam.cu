// K8 placeholder (prefix beam search lands after the training path): loud failure, no fallback.
#include "common.cuh"
extern "C" size_t asr_ctc_beam_workspace_bytes(int32_t, int32_t, int32_t, int32_t) { return 0; }
extern "C" int32_t asr_ctc_beam(const float*, int32_t, int32_t, int32_t, const int32_t*, int32_t, int32_t, int32_t,
                                int32_t*, int32_t*, void*, void*) {
  asr::set_error("asr_ctc_beam: not built yet");
  return ASR_ERR_UNSUPPORTED;
}
