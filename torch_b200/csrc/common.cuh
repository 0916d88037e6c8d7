// Shared helpers for the tltorch CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/tltorch_cuda.h"

namespace tlt {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define TLT_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::tlt::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return TLT_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define TLT_REQUIRE(cond, ...)              \
  do {                                      \
    if (!(cond)) {                          \
      ::tlt::set_error(__VA_ARGS__);        \
      return TLT_ERR_INVALID;               \
    }                                       \
  } while (0)

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("launch of %s failed: %s", what, cudaGetErrorString(e));
    return TLT_ERR_CUDA;
  }
  count_launch();
  return TLT_OK;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

inline size_t dtype_size(int dt) { return dt == TLT_BF16 ? 2 : 4; }

// Opt-in to `smem` bytes of dynamic shared memory for `kernel` on the CURRENT device.  The attribute is per device and per
// function, so the high-water mark is kept per (device, function) under a mutex (api.cu): safe with several GPUs in one
// process and with launches from several host threads.  Returns a tlt_status.
int ensure_smem(const void* kernel, size_t smem);
// SM count of the current device (cached per device).
int device_sms();
#define TLT_ENSURE_SMEM(kernel, smem)                                   \
  do {                                                                  \
    int _rc = ::tlt::ensure_smem((const void*)(kernel), (size_t)(smem)); \
    if (_rc != TLT_OK) return _rc;                                      \
  } while (0)

}  // namespace tlt
