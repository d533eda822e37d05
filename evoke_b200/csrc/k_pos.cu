// Positive-logit sums from the K2 bit mask:  row_pos[i] = sum_j M_ij * (q_i . k_j) * inv_tau.
// This is the "sum_j Y_ij S_ij" term of the soft-target cross entropy
// (models/model_pretrain_finetune_v0520.py:501-502, :443) before the division by c_i.  Positives
// are sparse (about one to three per row), so the term is O(N*D) work and does not belong in the
// N^2 tile epilogue: it runs on a side stream next to K3.
//
// One warp per row: the row's mask words are scanned with coalesced loads, every set bit triggers
// a warp-cooperative bf16 dot product (128-bit loads, fp32 accumulate) over the same bf16 operands
// the tensor cores see (hi, and hi/lo pairs in fp32-parity mode).
// Algorithmic bytes: ld_words*4 per row (mask) + 2*D per positive.
#include "evk_common.cuh"

namespace {

constexpr int kWarpsPerBlock = 8;

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = __bfloat1622float2(pa[e]);
    const float2 fb = __bfloat1622float2(pb[e]);
    s = fmaf(fa.x, fb.x, s);
    s = fmaf(fa.y, fb.y, s);
  }
  return s;
}

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pos_kernel(const __nv_bfloat16* __restrict__ q_hi, const __nv_bfloat16* __restrict__ q_lo, int64_t ld_q,
           const __nv_bfloat16* __restrict__ k_hi, const __nv_bfloat16* __restrict__ k_lo, int64_t ld_k,
           int64_t n_rows, int64_t n_cols, int d_vec /* ceil(d/8) */, const uint32_t* __restrict__ bits,
           int64_t ld_words, float inv_tau, float* __restrict__ row_pos) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  const int64_t words = (n_cols + 31) >> 5;
  for (int64_t i = warp0; i < n_rows; i += nwarps) {
    const uint32_t* mrow = bits + i * ld_words;
    const uint4* qh = reinterpret_cast<const uint4*>(q_hi + i * ld_q);
    const uint4* ql = q_lo ? reinterpret_cast<const uint4*>(q_lo + i * ld_q) : nullptr;
    float acc = 0.f;
    for (int64_t w0 = 0; w0 < words; w0 += 32) {
      const int64_t w = w0 + lane;
      uint32_t m = w < words ? __ldg(mrow + w) : 0u;
      uint32_t any = __ballot_sync(0xffffffffu, m != 0u);
      while (any) {
        const int src = __ffs(any) - 1;
        any &= any - 1;
        uint32_t mw = __shfl_sync(0xffffffffu, m, src);
        const int64_t jbase = (w0 + src) << 5;
        while (mw) {
          const int b = __ffs(mw) - 1;
          mw &= mw - 1;
          const int64_t j = jbase + b;
          const uint4* kh = reinterpret_cast<const uint4*>(k_hi + j * ld_k);
          const uint4* kl = k_lo ? reinterpret_cast<const uint4*>(k_lo + j * ld_k) : nullptr;
          float s = 0.f;
          for (int c = lane; c < d_vec; c += 32) {
            const uint4 a = __ldg(qh + c), bb = __ldg(kh + c);
            s += dot8_bf16(a, bb);
            if (ql) {                                  // (qh+ql).(kh+kl) without the lo.lo term, as the MMA segments
              const uint4 al = __ldg(ql + c), bl = __ldg(kl + c);
              s += dot8_bf16(a, bl) + dot8_bf16(al, bb);
            }
          }
          acc += s;
        }
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) row_pos[i] = acc * inv_tau;
  }
}

// raw dot products of the listed positives: pos_dot[i, s] = q_i . k_{pos_idx[i, s]} over the bf16 operands the
// tensor cores see, fp32 accumulate.  One warp per row, the query row stays in registers (kQVec 128-bit words
// per lane).  Runs on a side stream next to K3; the backward turns the values into exact W entries.
template <int kQVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pos_logits_kernel(const __nv_bfloat16* __restrict__ q_hi, int64_t ld_q, const __nv_bfloat16* __restrict__ k_hi,
                  int64_t ld_k, int64_t n_rows, int d_vec, const int32_t* __restrict__ pos_idx,
                  const int32_t* __restrict__ counts, int pos_slots, float* __restrict__ pos_dot) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n_rows) return;
  const int c = min(__ldg(counts + i), pos_slots);
  if (c <= 0) return;
  const int32_t myj = lane < c ? __ldg(pos_idx + i * pos_slots + lane) : 0;
  const uint4* qa = reinterpret_cast<const uint4*>(q_hi + i * ld_q);
  uint4 qv[kQVec];
#pragma unroll
  for (int t = 0; t < kQVec; ++t) qv[t] = (lane + 32 * t < d_vec) ? __ldg(qa + lane + 32 * t) : make_uint4(0u, 0u, 0u, 0u);
  for (int s = 0; s < c; ++s) {
    const int64_t j = __shfl_sync(0xffffffffu, myj, s);
    const uint4* kb = reinterpret_cast<const uint4*>(k_hi + j * ld_k);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < kQVec; ++t)
      if (lane + 32 * t < d_vec) acc += dot8_bf16(qv[t], __ldg(kb + lane + 32 * t));
    acc = warp_sum(acc);
    if (lane == 0) pos_dot[i * pos_slots + s] = acc;
  }
}

// Positive-logit sums WITHOUT the dense mask: row_pos[i] = inv_tau * sum over the positives of row i of q_i . k_j.
// Rows whose positives all fit the K2 list (counts[i] <= pos_slots: every row of a realistic batch) just add up
// their pos_dot entries; a row with more positives finds them by scanning the column ids (4-8 bytes per column
// instead of one mask bit, but only for such rows) and computes the dot products itself.  One warp per row.
template <int kQVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
pos_from_lists_kernel(const __nv_bfloat16* __restrict__ q_hi, int64_t ld_q, const __nv_bfloat16* __restrict__ k_hi,
                      int64_t ld_k, int64_t n_rows, int64_t n_cols, int d_vec, const int32_t* __restrict__ ids_row,
                      const int32_t* __restrict__ ids2_row, const int32_t* __restrict__ ids_col,
                      const int32_t* __restrict__ ids2_col, int64_t diag_offset, int clear_diag,
                      const int32_t* __restrict__ counts, const float* __restrict__ pos_dot, int pos_slots, float inv_tau,
                      float* __restrict__ row_pos) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (i >= n_rows) return;
  const int c = __ldg(counts + i);
  if (c <= pos_slots) {
    float v = lane < c ? __ldg(pos_dot + i * pos_slots + lane) : 0.f;
    v = warp_sum(v);
    if (lane == 0) row_pos[i] = v * inv_tau;
    return;
  }
  const int32_t key = __ldg(ids_row + i);
  const int32_t key2 = ids2_row ? __ldg(ids2_row + i) : 0;
  const int64_t diag = clear_diag ? i + diag_offset : -1;
  const uint4* qa = reinterpret_cast<const uint4*>(q_hi + i * ld_q);
  uint4 qv[kQVec];
#pragma unroll
  for (int t = 0; t < kQVec; ++t) qv[t] = (lane + 32 * t < d_vec) ? __ldg(qa + lane + 32 * t) : make_uint4(0u, 0u, 0u, 0u);
  float acc = 0.f;
  for (int64_t j0 = 0; j0 < n_cols; j0 += 32) {
    const int64_t j = j0 + lane;
    bool hit = j < n_cols && __ldg(ids_col + j) == key && j != diag;
    if (hit && ids2_col) hit = __ldg(ids2_col + j) == key2;
    uint32_t any = __ballot_sync(0xffffffffu, hit);
    while (any) {
      const int b = __ffs(any) - 1;
      any &= any - 1;
      const uint4* kb = reinterpret_cast<const uint4*>(k_hi + (j0 + b) * ld_k);
      float s = 0.f;
#pragma unroll
      for (int t = 0; t < kQVec; ++t)
        if (lane + 32 * t < d_vec) s += dot8_bf16(qv[t], __ldg(kb + lane + 32 * t));
      acc += s;
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) row_pos[i] = acc * inv_tau;
}

}  // namespace

extern "C" int evk_mpce_pos_from_lists(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t n_rows,
                                       int64_t n_cols, int64_t d, const int32_t* ids_row, const int32_t* ids2_row,
                                       const int32_t* ids_col, const int32_t* ids2_col, int64_t diag_offset, int clear_diag,
                                       const int32_t* counts, const float* pos_dot, int pos_slots, float inv_tau,
                                       float* row_pos, evk_stream_t stream) {
  EVK_REQUIRE(q_hi && k_hi && ids_row && ids_col && counts && pos_dot && row_pos, "evk_mpce_pos_from_lists: null pointer");
  EVK_REQUIRE((ids2_row == nullptr) == (ids2_col == nullptr), "evk_mpce_pos_from_lists: ids2_row/ids2_col must both be set or both null");
  EVK_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0 && d <= 4096 && pos_slots >= 1 && pos_slots <= 32,
              "evk_mpce_pos_from_lists: bad shape (d <= 4096, 1..32 slots)");
  EVK_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_q >= d && ld_k >= d && evk_aligned16(q_hi) && evk_aligned16(k_hi),
              "evk_mpce_pos_from_lists: operands need 16-byte aligned rows (ld %% 8 == 0)");
  const int64_t blocks = (n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int d_vec = (int)((d + 7) / 8);
  auto* qp = static_cast<const __nv_bfloat16*>(q_hi);
  auto* kp = static_cast<const __nv_bfloat16*>(k_hi);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d_vec <= 128)
    pos_from_lists_kernel<4><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(qp, ld_q, kp, ld_k, n_rows, n_cols, d_vec, ids_row,
                                                                            ids2_row, ids_col, ids2_col, diag_offset, clear_diag,
                                                                            counts, pos_dot, pos_slots, inv_tau, row_pos);
  else
    pos_from_lists_kernel<16><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(qp, ld_q, kp, ld_k, n_rows, n_cols, d_vec, ids_row,
                                                                             ids2_row, ids_col, ids2_col, diag_offset, clear_diag,
                                                                             counts, pos_dot, pos_slots, inv_tau, row_pos);
  EVK_CHECK_LAUNCH("mpce_pos_from_lists");
  return EVK_OK;
}

extern "C" int evk_mpce_pos_logits(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t n_rows,
                                   int64_t d, const int32_t* pos_idx, const int32_t* counts, int pos_slots,
                                   float* pos_dot, evk_stream_t stream) {
  EVK_REQUIRE(q_hi && k_hi && pos_idx && counts && pos_dot, "evk_mpce_pos_logits: null pointer");
  EVK_REQUIRE(n_rows > 0 && d > 0 && d <= 4096 && pos_slots >= 1 && pos_slots <= 32, "evk_mpce_pos_logits: bad shape (d <= 4096, 1..32 slots)");
  EVK_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_q >= d && ld_k >= d && evk_aligned16(q_hi) && evk_aligned16(k_hi),
              "evk_mpce_pos_logits: operands need 16-byte aligned rows (ld %% 8 == 0)");
  const int64_t blocks = (n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int d_vec = (int)((d + 7) / 8);
  auto* qp = static_cast<const __nv_bfloat16*>(q_hi);
  auto* kp = static_cast<const __nv_bfloat16*>(k_hi);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d_vec <= 128)
    pos_logits_kernel<4><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(qp, ld_q, kp, ld_k, n_rows, d_vec, pos_idx, counts,
                                                                        pos_slots, pos_dot);
  else
    pos_logits_kernel<16><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(qp, ld_q, kp, ld_k, n_rows, d_vec, pos_idx, counts,
                                                                         pos_slots, pos_dot);
  EVK_CHECK_LAUNCH("mpce_pos_logits");
  return EVK_OK;
}

extern "C" int evk_mpce_pos(const void* q_hi, const void* q_lo, int64_t ld_q, const void* k_hi, const void* k_lo,
                            int64_t ld_k, int64_t n_rows, int64_t n_cols, int64_t d, const uint32_t* bits,
                            int64_t ld_words, float inv_tau, float* row_pos, evk_stream_t stream) {
  EVK_REQUIRE(q_hi && k_hi && bits && row_pos, "evk_mpce_pos: null pointer");
  EVK_REQUIRE((q_lo == nullptr) == (k_lo == nullptr), "evk_mpce_pos: q_lo/k_lo must both be set or both null");
  EVK_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0, "evk_mpce_pos: empty problem");
  EVK_REQUIRE(ld_q % 8 == 0 && ld_k % 8 == 0 && ld_q >= d && ld_k >= d && evk_aligned16(q_hi) && evk_aligned16(k_hi) &&
                  (!q_lo || (evk_aligned16(q_lo) && evk_aligned16(k_lo))),
              "evk_mpce_pos: operands need 16-byte aligned rows (ld %% 8 == 0)");
  EVK_REQUIRE(ld_words >= (n_cols + 31) / 32, "evk_mpce_pos: ld_words too small");
  // rows are padded to a multiple of 8 elements by K1, which writes zeros into the pad columns
  int64_t blocks = (n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)evk_sm_count() * 4;
  if (blocks > cap) blocks = cap;
  pos_kernel<<<(unsigned)blocks, kWarpsPerBlock * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(q_hi), static_cast<const __nv_bfloat16*>(q_lo), ld_q,
      static_cast<const __nv_bfloat16*>(k_hi), static_cast<const __nv_bfloat16*>(k_lo), ld_k, n_rows, n_cols,
      (int)((d + 7) / 8), bits, ld_words, inv_tau, row_pos);
  EVK_CHECK_LAUNCH("mpce_pos");
  return EVK_OK;
}
