// C-ABI plumbing shared by every entry point: thread-local last-error string, device-property cache.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/spatten_b200.h"
#include "common.cuh"

namespace spatten {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
  set_error("%s: CUDA error %d (%s)", where, static_cast<int>(e), cudaGetErrorString(e));
  return SPATTEN_ERR_CUDA;
}

int sm_count() {
  static int cached[64] = {};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaGetDevice"); return -1; }
  if (dev < 64 && cached[dev] > 0) return cached[dev];
  int n = 0;
  e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) { cuda_fail(e, "cudaDeviceGetAttribute(SM count)"); return -1; }
  if (dev < 64) cached[dev] = n;
  return n;
}

}  // namespace spatten

extern "C" int spatten_abi_version(void) { return SPATTEN_ABI_VERSION; }

extern "C" const char* spatten_last_error(void) { return spatten::g_err; }

extern "C" int spatten_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return spatten::cuda_fail(e, "cudaGetDevice");
  int n = 0, maj = 0, mnr = 0;
  if ((e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev)) != cudaSuccess ||
      (e = cudaDeviceGetAttribute(&mnr, cudaDevAttrComputeCapabilityMinor, dev)) != cudaSuccess)
    return spatten::cuda_fail(e, "cudaDeviceGetAttribute");
  if (sm_count_out) *sm_count_out = n;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = mnr;
  return SPATTEN_OK;
}
