// Error reporting and device queries for the tcow_b200 C ABI (include/tcow_b200.h).
#include <stdarg.h>
#include <stdio.h>

#include "tcow_internal.h"

namespace tcow {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(TCOW_ERR_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  int& c = cached[dev & 63];
  if (c == 0) cudaDeviceGetAttribute(&c, cudaDevAttrMultiProcessorCount, dev);
  return c > 0 ? c : 148;
}

}  // namespace tcow

extern "C" int tcow_abi_version(void) { return 1; }
extern "C" const char* tcow_last_error(void) { return tcow::g_err; }
extern "C" int tcow_check_device(void) {
  int dev = 0, major = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return tcow::set_error(TCOW_ERR_CUDA, "no usable CUDA device: %s", cudaGetErrorString(e));
  if (major != 10)
    return tcow::set_error(TCOW_ERR_ARCH, "tcow_b200 requires a Blackwell sm_100 device (found sm_%d0); no fallback", major);
  return 0;
}
