// O(N) statistics kernels: deterministic reduction of per-tile partial sums, and the
// statistics -> (loss, a_row, b_col) step that closes the forward pass.
// Together with K3 these replace F.cross_entropy(logits/temp, labels) x2 and the final
// (l1 + l2)/2 of models/model_pretrain_finetune_v0520.py:501-503 (and :443 for MPC).
#include "evk_common.cuh"

namespace {

__global__ void reduce_partials_kernel(const float* __restrict__ part, int64_t parts, int64_t ld, int64_t n,
                                       float* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int64_t p = 0; p < parts; ++p) acc += part[p * ld + j];   // fixed order: deterministic
    out[j] = acc;
  }
}

__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int kFinThreads = 1024;

// Single CTA: N <= a few 1e5 elements, fixed summation order, fp64 accumulation.
__global__ void __launch_bounds__(kFinThreads)
finalize_kernel(const float* __restrict__ row_sum, const float* __restrict__ row_pos,
                const int32_t* __restrict__ counts, int64_t n_rows, const float* __restrict__ col_sum,
                int64_t n_cols, int64_t col_lo, int64_t col_hi, float shift, float pos_weight, double inv_count,
                float* __restrict__ a_row, float* __restrict__ b_col, float* __restrict__ loss_out) {
  __shared__ double s_part[kFinThreads / 32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n_rows; i += kFinThreads) {
    const float r = row_sum[i];
    const int c = counts[i];
    if (a_row) a_row[i] = 1.f / r;
    const double pos = c > 0 ? (double)row_pos[i] / (double)c : 0.0;
    acc += (double)shift + (double)logf(r) - (double)pos_weight * pos;
  }
  if (col_sum) {
    for (int64_t j = threadIdx.x; j < n_cols; j += kFinThreads) {
      const float c = col_sum[j];
      if (b_col) b_col[j] = 1.f / c;
      if (j >= col_lo && j < col_hi) acc += (double)shift + (double)logf(c);
    }
  }
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kFinThreads / 32; ++w) t += s_part[w];
    loss_out[0] = (float)(t * inv_count);
  }
}

}  // namespace

extern "C" int evk_reduce_partials(const float* part, int64_t parts, int64_t ld, int64_t n, float* out,
                                   evk_stream_t stream) {
  EVK_REQUIRE(part && out && parts >= 1 && ld >= n && n >= 0, "evk_reduce_partials: bad arguments");
  if (n == 0) return EVK_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)evk_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(part, parts, ld, n, out);
  EVK_CHECK_LAUNCH("reduce_partials");
  return EVK_OK;
}

extern "C" int evk_mpce_finalize(const float* row_sum, const float* row_pos, const int32_t* counts, int64_t n_rows,
                                 const float* col_sum, int64_t n_cols, int64_t col_lo, int64_t col_hi, float shift,
                                 float pos_weight, double inv_count, float* a_row, float* b_col, float* loss_out,
                                 evk_stream_t stream) {
  EVK_REQUIRE(row_sum && row_pos && counts && loss_out && n_rows > 0, "evk_mpce_finalize: null pointer or n_rows <= 0");
  EVK_REQUIRE(!col_sum || n_cols > 0, "evk_mpce_finalize: col_sum given with n_cols <= 0");
  finalize_kernel<<<1, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      row_sum, row_pos, counts, n_rows, col_sum, n_cols, col_lo, col_hi, shift, pos_weight, inv_count, a_row, b_col,
      loss_out);
  EVK_CHECK_LAUNCH("mpce_finalize");
  return EVK_OK;
}
