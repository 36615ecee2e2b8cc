// placeholder until the tcgen05 persistent BiLSTM lands: declines every shape.
#include "common.cuh"
namespace lstmtc {
bool supports_fwd(const asr_lstm_fwd_args*) { return false; }
bool supports_bwd(const asr_lstm_bwd_args*) { return false; }
size_t scratch_bytes(int) { return 0; }
int32_t forward(const asr_lstm_fwd_args*, cudaStream_t) { asr::set_error("lstm_tc: not built"); return ASR_ERR_UNSUPPORTED; }
int32_t backward(const asr_lstm_bwd_args*, cudaStream_t) { asr::set_error("lstm_tc: not built"); return ASR_ERR_UNSUPPORTED; }
}  // namespace lstmtc
