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

// ---- sizes of caller-owned buffers: the single source of truth for the layouts the kernels assume ------------
int64_t evk_stats_workspace_bytes(int64_t n_rows, int64_t n_cols) {
  if (n_rows <= 0 || n_cols < 0) return 0;
  const int64_t n = n_cols > n_rows ? n_cols : n_rows;
  const int64_t b = 16 + 24 * ((n + 31) / 32);          // ticket + 3 statistics x one fp64 partial per 32-element CTA
  return (b + 15) / 16 * 16;
}

int64_t evk_shard_finish_workspace_bytes(int64_t n_cols) {
  if (n_cols <= 0) return 0;
  const int64_t b = 16 + 8 * ((n_cols + 255) / 256);
  return (b + 15) / 16 * 16;
}

int64_t evk_posmask_ld_words(int64_t n_cols) {
  if (n_cols <= 0) return 0;
  const int64_t w = (n_cols + 255) / 256 * 8;           // whole 256-column tiles
  return (w + 7) / 8 * 8;
}

int64_t evk_mpce_rowpart_rows(int64_t n_cols) {
  if (n_cols <= 0) return 0;
  return (n_cols + 255) / 256 * (int64_t)evk_mpce_row_parts();
}

int64_t evk_mpce_colpart_rows(int64_t n_rows) { return n_rows <= 0 ? 0 : (n_rows + 127) / 128; }

int64_t evk_mpce_strip_ld(int64_t n_cols) { return n_cols <= 0 ? 0 : (n_cols + 63) / 64 * 64; }

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
