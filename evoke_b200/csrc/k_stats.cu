// O(N) statistics kernels: deterministic reduction of per-tile partial sums, and the
// statistics -> (loss, a_row, b_col) step that closes the forward pass.
// Together with K3 these replace F.cross_entropy(logits/temp, labels) x2 and the final
// (l1 + l2)/2 of models/model_pretrain_finetune_v0520.py:501-503 (and :443 for MPC).
#include "evk_common.cuh"

#include <string.h>

namespace {

__global__ void reduce_partials_kernel(const float* __restrict__ part, int64_t parts, int64_t ld, int64_t n,
                                       const int32_t* __restrict__ divisor, float* __restrict__ out) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int64_t p = 0; p < parts; ++p) acc += part[p * ld + j];   // fixed order: deterministic
    if (divisor) {
      const int c = divisor[j];
      acc = c > 0 ? acc / (float)c : 0.f;
    }
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

// 'Averaged positive logit' rule, one softmax direction (PretrainNewMulPos.global_alignment_loss :783-811,
// multi_pos_contra_images_v0404 :691-705).  With row_neg[i] = sum over the NEGATIVES of exp(S_ij - shift) and
// row_pos[i] = sum over the positives of S_ij:
//   pbar = row_pos/c,  u = exp(pbar - shift),  Z = u + row_neg,  l_i = -pbar + shift + ln Z
//   a_row[i] = 1/Z (weight of a negative: E_ij/Z),  pos_row[i] = (u/Z - 1)/c (weight of each positive)
// Rows without positives (c = 0: single-view rows of v0404, :685) contribute nothing and get zero weights.
__global__ void __launch_bounds__(kFinThreads)
finalize_avgpos_kernel(const float* __restrict__ row_neg, const float* __restrict__ row_pos,
                       const int32_t* __restrict__ counts, int64_t n_rows, float shift, double inv_count,
                       float* __restrict__ a_row, float* __restrict__ pos_row, float* __restrict__ loss_out,
                       int accumulate) {
  __shared__ double s_part[kFinThreads / 32];
  double acc = 0.0;
  for (int64_t i = threadIdx.x; i < n_rows; i += kFinThreads) {
    const int c = counts[i];
    float a = 0.f, pw = 0.f;
    if (c > 0) {
      const float pbar = row_pos[i] / (float)c;
      const float u = expf(pbar - shift);
      const float z = u + row_neg[i];
      a = 1.f / z;
      pw = (u * a - 1.f) / (float)c;
      acc += (double)shift - (double)pbar + (double)logf(z);
    }
    a_row[i] = a;
    pos_row[i] = pw;
  }
  acc = warp_sum_f64(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kFinThreads / 32; ++w) t += s_part[w];
    const float l = (float)(t * inv_count);
    loss_out[0] = accumulate ? loss_out[0] + l : l;
  }
}

// Fused single-GPU form: reduce the per-tile partials, emit a_row / b_col and the loss in ONE
// multi-CTA launch.  Each CTA owns 256 rows (and the same 256 columns), adds its terms in fp64,
// publishes a per-CTA partial, and the last CTA to finish (atomic ticket) sums the partials in
// index order, so the result is deterministic.
// Sharded form: instead of b_col = 1/C_j the RAW partial column sums of this rank's rows, and instead of the
// final loss this rank's row-side loss term, are stored into the statistics slot of EVERY rank (peer-mapped
// pointers), i.e. the reduce + finalize + all-gather of the sharded forward in one launch.
struct StatPush {
  float* dst[16];
  int n;
  int64_t offset;        // element offset of this rank's slot inside each destination
};

constexpr int kStatThreads = 256;
constexpr int kStatElems = 32;                       // rows (and columns) per CTA
constexpr int kStatGroups = kStatThreads / kStatElems;   // partial-sum groups per element

// sum of part[p*ld + i] over p = g, g+G, ... with eight independent loads in flight
__device__ __forceinline__ float strided_partial_sum(const float* __restrict__ part, int64_t parts, int64_t ld,
                                                     int64_t i, int g) {
  float s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = 0.f;
  int64_t p = g;
  for (; p + 7 * kStatGroups < parts; p += 8 * kStatGroups) {
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += part[(p + k * kStatGroups) * ld + i];
  }
  for (; p < parts; p += kStatGroups) s[0] += part[p * ld + i];
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
}

// blockIdx.y selects the statistic a CTA reduces - 0: row exp-sums (a_row, ln R_i), 1: positive-logit sums
// (- pos_weight pos_i / c_i), 2: column exp-sums (b_col, ln C_j) - the loss terms are additive, so the three
// run side by side instead of one after the other in the same threads.
__global__ void __launch_bounds__(kStatThreads)
stats_fused_kernel(const float* __restrict__ rs_part, int64_t row_parts, int64_t ld_row,
                   const float* __restrict__ rp_part, int64_t pos_parts, int64_t ld_pos,
                   const int32_t* __restrict__ counts, int64_t n_rows,
                   const float* __restrict__ cs_part, int64_t col_parts, int64_t ld_col, int64_t n_cols,
                   int64_t col_lo, int64_t col_hi, float shift, float pos_weight, double inv_count,
                   float* __restrict__ a_row, float* __restrict__ b_col, float* __restrict__ loss_out,
                   double* __restrict__ cta_partial, unsigned int* __restrict__ ticket, const StatPush push) {
  __shared__ float s_sum[kStatGroups][kStatElems];
  __shared__ double s_part[kStatElems / 32];
  __shared__ bool s_last;
  const int e = threadIdx.x % kStatElems, g = threadIdx.x / kStatElems;
  const int64_t i = (int64_t)blockIdx.x * kStatElems + e;
  const int which = blockIdx.y;
  float v = 0.f;
  if (which == 0 && i < n_rows) v = strided_partial_sum(rs_part, row_parts, ld_row, i, g);
  if (which == 1 && i < n_rows) v = strided_partial_sum(rp_part, pos_parts, ld_pos, i, g);
  if (which == 2 && cs_part && i < n_cols) v = strided_partial_sum(cs_part, col_parts, ld_col, i, g);
  s_sum[g][e] = v;
  __syncthreads();
  if (threadIdx.x < kStatElems) {
    double acc = 0.0;
    float t = 0.f;
#pragma unroll
    for (int gg = 0; gg < kStatGroups; ++gg) t += s_sum[gg][e];          // fixed order: deterministic
    if (which == 0 && i < n_rows) {
      a_row[i] = 1.f / t;
      acc = (double)shift + (double)logf(t);
    } else if (which == 1 && i < n_rows) {
      const int cnt = counts ? counts[i] : 1;
      acc = -(double)pos_weight * (cnt > 0 ? (double)t / (double)cnt : 0.0);
    } else if (which == 2 && cs_part && i < n_cols) {
      if (push.n > 0) {
        for (int p = 0; p < push.n; ++p) push.dst[p][push.offset + i] = t;
      } else {
        b_col[i] = 1.f / t;
        if (i >= col_lo && i < col_hi) acc = (double)shift + (double)logf(t);
      }
    }
    acc = warp_sum_f64(acc);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kStatElems / 32; ++w) t += s_part[w];
    cta_partial[blockIdx.y * gridDim.x + blockIdx.x] = t;
    __threadfence();
    const unsigned int done = atomicAdd(ticket, 1u);
    s_last = (done == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {
    __threadfence();
    // lane l sums partials l, l+32, ... ; lanes are then combined in a fixed butterfly order
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x * gridDim.y; b += 32)
      t += reinterpret_cast<volatile double*>(cta_partial)[b];
    t = warp_sum_f64(t);
    if (threadIdx.x == 0) {
      const float l = (float)(t * inv_count);
      if (push.n > 0) {
        for (int p = 0; p < push.n; ++p) push.dst[p][push.offset + n_cols] = l;
      } else {
        loss_out[0] = l;
      }
      *ticket = 0u;                                  // a persistent workspace is ready for the next call
    }
  }
}

}  // namespace

namespace {
int stats_fused_impl(const float* rs_part, int64_t row_parts, int64_t ld_row, const float* rp_part,
                     int64_t pos_parts, int64_t ld_pos, const int32_t* counts, int64_t n_rows,
                     const float* cs_part, int64_t col_parts,
                     int64_t ld_col, int64_t n_cols, int64_t col_lo, int64_t col_hi, float shift,
                     float pos_weight, double inv_count, float* a_row, float* b_col, float* loss_out,
                     void* workspace, int64_t workspace_bytes, const StatPush& push, evk_stream_t stream,
                     bool workspace_persistent = false) {
  EVK_REQUIRE(rs_part && rp_part && a_row && (loss_out || push.n > 0) && workspace && n_rows > 0 && row_parts >= 1 &&
                  ld_row >= n_rows && pos_parts >= 1 && ld_pos >= n_rows, "evk_mpce_stats_fused: bad row arguments");
  EVK_REQUIRE(!cs_part || ((b_col || push.n > 0) && col_parts >= 1 && ld_col >= n_cols && n_cols > 0),
              "evk_mpce_stats_fused: bad column arguments");
  const int64_t n = (cs_part && n_cols > n_rows) ? n_cols : n_rows;
  const int64_t blocks = (n + kStatElems - 1) / kStatElems;
  EVK_REQUIRE(workspace_bytes >= 16 + 24 * blocks && evk_aligned16(workspace),
              "evk_mpce_stats_fused: workspace needs %lld bytes, 16-byte aligned", (long long)(16 + 24 * blocks));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* ticket = static_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 16);
  if (!workspace_persistent) EVK_CUDA(cudaMemsetAsync(ticket, 0, 16, s));
  stats_fused_kernel<<<dim3((unsigned)blocks, 3), kStatThreads, 0, s>>>(rs_part, row_parts, ld_row, rp_part, pos_parts, ld_pos,
                                                              counts, n_rows,
                                                              cs_part, col_parts, ld_col, cs_part ? n_cols : 0, col_lo,
                                                              col_hi, shift, pos_weight, inv_count, a_row, b_col,
                                                              loss_out, partial, ticket, push);
  EVK_CHECK_LAUNCH("mpce_stats_fused");
  return EVK_OK;
}
}  // namespace

extern "C" int evk_mpce_stats_fused(const float* rs_part, int64_t row_parts, int64_t ld_row, const float* rp_part,
                                    int64_t pos_parts, int64_t ld_pos, const int32_t* counts, int64_t n_rows,
                                    const float* cs_part, int64_t col_parts,
                                    int64_t ld_col, int64_t n_cols, int64_t col_lo, int64_t col_hi, float shift,
                                    float pos_weight, double inv_count, float* a_row, float* b_col, float* loss_out,
                                    void* workspace, int64_t workspace_bytes, evk_stream_t stream) {
  StatPush push;
  memset(&push, 0, sizeof(push));
  return stats_fused_impl(rs_part, row_parts, ld_row, rp_part, pos_parts, ld_pos, counts, n_rows, cs_part, col_parts, ld_col,
                          n_cols, col_lo, col_hi, shift, pos_weight, inv_count, a_row, b_col, loss_out, workspace,
                          workspace_bytes, push, stream);
}

extern "C" int evk_mpce_shard_stats_push(const float* rs_part, int64_t row_parts, int64_t ld_row, const float* rp_part,
                                         int64_t pos_parts, int64_t ld_pos, const int32_t* counts, int64_t n_rows,
                                         const float* cs_part, int64_t col_parts, int64_t ld_col, int64_t n_cols,
                                         float shift, float pos_weight, double inv_count, float* a_row,
                                         const uint64_t* slot_ptrs, int n_dst, int64_t slot_offset, void* workspace,
                                         int64_t workspace_bytes, int workspace_persistent, evk_stream_t stream) {
  EVK_REQUIRE(slot_ptrs && n_dst >= 1 && n_dst <= 16 && slot_offset >= 0 && cs_part, "evk_mpce_shard_stats_push: bad destinations");
  StatPush push;
  memset(&push, 0, sizeof(push));
  push.n = n_dst;
  push.offset = slot_offset;
  for (int p = 0; p < n_dst; ++p) {
    push.dst[p] = reinterpret_cast<float*>(slot_ptrs[p]);
    EVK_REQUIRE(push.dst[p], "evk_mpce_shard_stats_push: null destination");
  }
  return stats_fused_impl(rs_part, row_parts, ld_row, rp_part, pos_parts, ld_pos, counts, n_rows, cs_part, col_parts, ld_col,
                          n_cols, 0, 0, shift, pos_weight, inv_count, a_row, nullptr, nullptr, workspace, workspace_bytes,
                          push, stream, workspace_persistent != 0);
}

extern "C" int evk_reduce_partials(const float* part, int64_t parts, int64_t ld, int64_t n, const int32_t* divisor,
                                   float* out, evk_stream_t stream) {
  EVK_REQUIRE(part && out && parts >= 1 && ld >= n && n >= 0, "evk_reduce_partials: bad arguments");
  if (n == 0) return EVK_OK;
  int64_t blocks = (n + 255) / 256;
  const int64_t cap = (int64_t)evk_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  reduce_partials_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(part, parts, ld, n, divisor, out);
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

extern "C" int evk_mpce_finalize_avgpos(const float* row_neg, const float* row_pos, const int32_t* counts, int64_t n_rows,
                                        float shift, double inv_count, float* a_row, float* pos_row, float* loss_out,
                                        int accumulate, evk_stream_t stream) {
  EVK_REQUIRE(row_neg && row_pos && counts && a_row && pos_row && loss_out && n_rows > 0,
              "evk_mpce_finalize_avgpos: null pointer or n_rows <= 0");
  finalize_avgpos_kernel<<<1, kFinThreads, 0, static_cast<cudaStream_t>(stream)>>>(row_neg, row_pos, counts, n_rows, shift,
                                                                                 inv_count, a_row, pos_row, loss_out,
                                                                                 accumulate);
  EVK_CHECK_LAUNCH("mpce_finalize_avgpos");
  return EVK_OK;
}
