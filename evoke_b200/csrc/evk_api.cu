// Error state, version and device queries of the C ABI (include/evoke_b200.h).
#include "evk_common.cuh"

#include <mutex>

namespace {
thread_local char g_err[1024] = "";
}

int evk_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int evk_sm_count() {
  static int cached[64];
  static std::mutex mu;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  std::lock_guard<std::mutex> lock(mu);
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

bool evk_is_sm100() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return false;
  return major == 10;
}

extern "C" {

int evk_version(void) { return EVK_ABI_VERSION; }

const char* evk_last_error(void) { return g_err; }

int evk_device_info(int device, int* sm_count, int* cc_major, int* cc_minor) {
  int v = 0;
  if (sm_count) {
    EVK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device));
    *sm_count = v;
  }
  if (cc_major) {
    EVK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, device));
    *cc_major = v;
  }
  if (cc_minor) {
    EVK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, device));
    *cc_minor = v;
  }
  return EVK_OK;
}

}  // extern "C"
