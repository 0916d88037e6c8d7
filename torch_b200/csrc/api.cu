// ABI plumbing: version, thread-local error string, launch counter.
#include "common.cuh"
#include <string.h>
#include <map>
#include <mutex>
#include <utility>

namespace tlt {
static thread_local char g_err[1024] = "";
std::atomic<uint64_t> g_launches{0};

int ensure_smem(const void* kernel, size_t smem) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> high;
  if (smem <= 48 * 1024) return TLT_OK;
  int dev = 0;
  TLT_CHECK_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  size_t& h = high[std::make_pair(dev, kernel)];
  if (smem > h) {
    TLT_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h = smem;
  }
  return TLT_OK;
}

int device_sms() {
  static std::mutex mu;
  static std::map<int, int> sms;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  std::lock_guard<std::mutex> lock(mu);
  int& n = sms[dev];
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

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
