// f4: exact inner-product top-k for the "patient-specific knowledge" retrieval of PretrainTester.predict
// (reference modules/multiview/trainer.py:543-653: faiss IndexIVFFlat, METRIC_INNER_PRODUCT, search(x, k)).
// The similarity contraction runs on the tcgen05 main loop (evk_tc_gemm_nt, tc_engine.cu) in chunks
// [queries x corpus block]; this kernel folds each chunk of scores into the running top-k of every query.
// At the reference's feature size (d = output_dim * 50 = 38400) the contraction does 76.8 kFLOP per score
// against the 8 bytes the score costs to write and read back, so keeping the epilogue out of the GEMM is free.
//
// One warp per query.  The list (k <= 64 entries, sorted by score descending, ties by corpus index ascending =
// numpy's stable argsort of the negated scores) lives in registers, kPerLane consecutive positions per lane.
// A chunk row is scanned 32 scores at a time; only scores above the current k-th value are inserted.
#include "evk_common.cuh"

#include <math_constants.h>

namespace {

constexpr int kWarps = 8;

template <int kPerLane>
__global__ void __launch_bounds__(kWarps * 32)
topk_update_kernel(const float* __restrict__ scores, int64_t ld, int64_t n_q, int64_t n_c, int64_t col_offset,
                   const int32_t* __restrict__ q_group, const int32_t* __restrict__ c_group, int k,
                   float* __restrict__ best_val, int32_t* __restrict__ best_idx, int init) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
  if (q >= n_q) return;
  float val[kPerLane];
  int32_t idx[kPerLane];
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int p = lane * kPerLane + j;
    val[j] = (!init && p < k) ? best_val[q * k + p] : -CUDART_INF_F;
    idx[j] = (!init && p < k) ? best_idx[q * k + p] : -1;
  }
  auto kth = [&]() {                          // value at position k-1: the admission threshold
    float v = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < kPerLane; ++j)
      if (lane * kPerLane + j == k - 1) v = val[j];
    return __shfl_sync(0xffffffffu, v, (k - 1) / kPerLane);
  };
  float thr = kth();
  const int32_t qg = q_group ? __ldg(q_group + q) : 0;
  const float* row = scores + q * ld;
  for (int64_t c0 = 0; c0 < n_c; c0 += 32) {
    const int64_t c = c0 + lane;
    float s = c < n_c ? __ldg(row + c) : -CUDART_INF_F;
    if (c < n_c && q_group && __ldg(c_group + col_offset + c) == qg) s = -CUDART_INF_F;   // same study: not a candidate
    if (s != s) s = -CUDART_INF_F;
    uint32_t pass = __ballot_sync(0xffffffffu, s > thr);
    while (pass) {
      const int b = __ffs(pass) - 1;
      pass &= pass - 1;
      const float sv = __shfl_sync(0xffffffffu, s, b);
      if (!(sv > thr)) continue;              // the threshold may have risen since the ballot
      const int32_t si = (int32_t)(c0 + b + col_offset);
      // insertion position = number of entries ordered before the new one
      int before = 0;
#pragma unroll
      for (int j = 0; j < kPerLane; ++j) before += (val[j] > sv || (val[j] == sv && idx[j] >= 0 && idx[j] < si)) ? 1 : 0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
      // shift the tail down by one position
      const float up_v = __shfl_up_sync(0xffffffffu, val[kPerLane - 1], 1);
      const int32_t up_i = __shfl_up_sync(0xffffffffu, idx[kPerLane - 1], 1);
#pragma unroll
      for (int j = kPerLane - 1; j >= 0; --j) {
        const int p = lane * kPerLane + j;
        if (p > before) {
          val[j] = j > 0 ? val[j - 1] : up_v;
          idx[j] = j > 0 ? idx[j - 1] : up_i;
        } else if (p == before) {
          val[j] = sv;
          idx[j] = si;
        }
      }
      thr = kth();
    }
  }
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int p = lane * kPerLane + j;
    if (p < k) {
      best_val[q * k + p] = val[j];
      best_idx[q * k + p] = idx[j];
    }
  }
}

}  // namespace

extern "C" int evk_topk_update(const float* scores, int64_t ld, int64_t n_q, int64_t n_c, int64_t col_offset,
                               const int32_t* q_group, const int32_t* c_group, int k, float* best_val,
                               int32_t* best_idx, int init, evk_stream_t stream) {
  EVK_REQUIRE(scores && best_val && best_idx, "evk_topk_update: null pointer");
  EVK_REQUIRE(n_q > 0 && n_c > 0 && ld >= n_c && col_offset >= 0, "evk_topk_update: bad shape");
  EVK_REQUIRE(k >= 1 && k <= 64, "evk_topk_update: k=%d outside 1..64", k);
  EVK_REQUIRE((q_group == nullptr) == (c_group == nullptr), "evk_topk_update: q_group / c_group must both be set or both null");
  EVK_REQUIRE(col_offset + n_c < (1ll << 31), "evk_topk_update: corpus index does not fit int32");
  const int64_t blocks = (n_q + kWarps - 1) / kWarps;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (k <= 32)
    topk_update_kernel<1><<<(unsigned)blocks, kWarps * 32, 0, s>>>(scores, ld, n_q, n_c, col_offset, q_group, c_group, k,
                                                                 best_val, best_idx, init);
  else
    topk_update_kernel<2><<<(unsigned)blocks, kWarps * 32, 0, s>>>(scores, ld, n_q, n_c, col_offset, q_group, c_group, k,
                                                                 best_val, best_idx, init);
  EVK_CHECK_LAUNCH("topk_update");
  return EVK_OK;
}
