// ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include <string.h>

namespace tlt {
static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace tlt

extern "C" int tlt_version(void) { return TLT_ABI_VERSION; }
extern "C" const char* tlt_last_error_string(void) { return tlt::g_err; }
extern "C" uint64_t tlt_launch_count(void) { return tlt::g_launches.load(); }
