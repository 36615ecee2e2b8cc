// error reporting, version, launch accounting for libasr_b200.
#include "common.cuh"

namespace asr {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace asr

extern "C" const char* asr_last_error(void) { return asr::g_err; }
extern "C" int32_t asr_version(void) { return 100; }
extern "C" int64_t asr_launch_count(void) { return asr::g_launches.load(); }
