// Shared host-side helpers of libikflow_b200 (error reporting, launch accounting).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/ikflow_b200.h"

namespace ikf {

std::string& last_error_ref();
extern std::atomic<uint64_t> g_launch_count;

inline int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define IKF_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      return ::ikf::fail(_e == cudaErrorMemoryAllocation ? IKF_ENOMEM : IKF_ECUDA, "%s failed: %s (%s:%d)", #expr, \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                                   \
  } while (0)

// RAII device switch: every entry point runs on the handle's device and restores the caller's.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace ikf
