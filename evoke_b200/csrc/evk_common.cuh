// Shared host/device helpers for libevoke_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/evoke_b200.h"

#define EVK_ABI_VERSION 3
#define EVK_NORM_EPS 1e-12f       // F.normalize default eps

// ---- error state (thread local; see evk_api.cu) ------------------------------------------
int evk_set_error(int code, const char* fmt, ...);

#define EVK_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return evk_set_error(EVK_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define EVK_CUDA(call)                                                                 \
  do {                                                                                 \
    cudaError_t e__ = (call);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return evk_set_error(EVK_ERR_CUDA, "%s failed: %s (%s:%d)", #call,               \
                           cudaGetErrorString(e__), __FILE__, __LINE__);               \
  } while (0)

#define EVK_CHECK_LAUNCH(name)                                                         \
  do {                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                              \
    if (e__ != cudaSuccess)                                                            \
      return evk_set_error(EVK_ERR_CUDA, "launch of %s failed: %s", name,              \
                           cudaGetErrorString(e__));                                   \
  } while (0)

static inline bool evk_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// number of SMs of the current device (cached per device)
int evk_sm_count();
// true when the current device is compute capability 10.x
bool evk_is_sm100();

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float load_as_float(const void* p, int dtype, int64_t idx) {
  if (dtype == EVK_DTYPE_F32) return reinterpret_cast<const float*>(p)[idx];
  if (dtype == EVK_DTYPE_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[idx]);
  return __half2float(reinterpret_cast<const __half*>(p)[idx]);
}

__device__ __forceinline__ void store_from_float(void* p, int dtype, int64_t idx, float v) {
  if (dtype == EVK_DTYPE_F32) reinterpret_cast<float*>(p)[idx] = v;
  else if (dtype == EVK_DTYPE_BF16) reinterpret_cast<__nv_bfloat16*>(p)[idx] = __float2bfloat16_rn(v);
  else reinterpret_cast<__half*>(p)[idx] = __float2half_rn(v);
}
