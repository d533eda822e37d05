// Small path: fp32 SIMT fused similarity + masked multi-positive softmax statistics (forward)
// and the row-direction gradient contraction (backward) for reference-sized batches (the
// reference trains with 32 studies per step, run_cxr_pt_224.sh:14).  At these sizes the work is
// launch-latency bound; a 128x256 tcgen05 tile would be mostly padding.  Exact fp32 FMA
// arithmetic makes this the tightest-parity path (fp32 loss ~1e-6 rel).
//
// One CTA owns kTM query rows and streams all key columns in tiles of 256 (one column per
// thread).  Nothing of size n_rows x n_cols is stored; the backward recomputes the S tile.
// Replaces, for one softmax direction, the mm + /temp + cross_entropy of
// models/model_pretrain_finetune_v0520.py:499-502 (G) and :437-443 (MPC).
#include "evk_common.cuh"

namespace {

constexpr int kThreads = 256;        // = key columns per tile
constexpr int kTM = 8;               // query rows per CTA
constexpr int kDK = 32;              // feature chunk staged through shared memory
constexpr int kMaxSlots = 8;         // backward: feature columns per thread per pass (8*256 = 2048)

template <bool kBwd>
__global__ void __launch_bounds__(kThreads, kBwd ? 2 : 3)
small_rows_kernel(const float* __restrict__ q, int64_t ld_q, const float* __restrict__ k, int64_t ld_k,
                  int64_t n_rows, int64_t n_cols, int d, int d_pad, const uint32_t* __restrict__ bits,
                  int64_t ld_words, const int32_t* __restrict__ counts, const float* __restrict__ a_row,
                  const float* __restrict__ b_col, float inv_tau, int flags, int64_t diag_offset,
                  float* __restrict__ row_sum, float* __restrict__ row_pos, float* __restrict__ dq,
                  int64_t ld_dq, const float* __restrict__ pos_row, const float* __restrict__ pos_col,
                  int64_t bs_q, int64_t bs_k, int64_t bs_vec, int64_t bs_dq) {
  // blockIdx.y = batch entry (independent problems of the same shape sharing mask and counts: the per-sample
  // token-level InfoNCE of local_text_token_alignment_loss :518-525); all strides are 0 for a single problem
  q += blockIdx.y * bs_q;
  k += blockIdx.y * bs_k;
  if (a_row) a_row += blockIdx.y * bs_vec;
  if (b_col) b_col += blockIdx.y * bs_vec;
  if (row_sum) row_sum += blockIdx.y * bs_vec;
  if (row_pos) row_pos += blockIdx.y * bs_vec;
  if (dq) dq += blockIdx.y * bs_dq;
  if (pos_row) pos_row += blockIdx.y * bs_vec;
  if (pos_col) pos_col += blockIdx.y * bs_vec;
  extern __shared__ __align__(16) float smem[];
  float* qs = smem;                                  // [kTM][d_pad]
  float* kt = qs + (size_t)kTM * d_pad;              // [256][kDK+1]
  float* ws = kt + kThreads * (kDK + 1);             // [256][kTM]   (backward) / reduction scratch

  const int t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  // 128-bit loads of the key rows need 16-byte aligned rows in every batch entry
  const bool k_vec = ((reinterpret_cast<uintptr_t>(k) & 15u) == 0) && (ld_k % 4 == 0);
  const int64_t r0 = (int64_t)blockIdx.x * kTM;
  const float shift = inv_tau;                       // |S| <= inv_tau for unit rows
  const bool excl = (flags & EVK_FLAG_EXCLUDE_DIAG) != 0;
  // 'averaged positive logit' rule (PretrainNewMulPos :748-815, :670-708): the positives enter the softmax as ONE
  // logit, so the exp-sum runs over the negatives only and a positive's weight has no E term
  const bool avg = (flags & EVK_FLAG_AVGPOS) != 0;

  for (int i = t; i < kTM * d_pad; i += kThreads) {
    const int r = i / d_pad, c = i - r * d_pad;
    qs[i] = (r0 + r < n_rows && c < d) ? q[(r0 + r) * ld_q + c] : 0.f;
  }

  float a_i[kTM], neg2_over_c[kTM];             // neg2_over_c: the row's share of a positive's weight in AVGPOS mode
  if (kBwd) {
#pragma unroll
    for (int r = 0; r < kTM; ++r) {
      const bool ok = r0 + r < n_rows;
      a_i[r] = ok ? a_row[r0 + r] : 0.f;
      if (avg) {
        neg2_over_c[r] = ok ? pos_row[r0 + r] : 0.f;
      } else {
        const int c = ok ? counts[r0 + r] : 1;
        neg2_over_c[r] = -2.f / (float)(c > 0 ? c : 1);
      }
    }
  }

  float rs[kTM], rp[kTM];
#pragma unroll
  for (int r = 0; r < kTM; ++r) rs[r] = rp[r] = 0.f;

  const int n_pass = kBwd ? (d + kMaxSlots * kThreads - 1) / (kMaxSlots * kThreads) : 1;
  for (int pass = 0; pass < n_pass; ++pass) {
    const int dbase = pass * kMaxSlots * kThreads;
    float acc[kBwd ? kMaxSlots : 1][kTM];
    if (kBwd) {
#pragma unroll
      for (int m = 0; m < kMaxSlots; ++m)
#pragma unroll
        for (int r = 0; r < kTM; ++r) acc[m][r] = 0.f;
    }

    for (int64_t j0 = 0; j0 < n_cols; j0 += kThreads) {
      // ---- phase A: S[kTM, 256] tile, one key column per thread ----
      float s[kTM];
#pragma unroll
      for (int r = 0; r < kTM; ++r) s[r] = 0.f;
      for (int dk0 = 0; dk0 < d_pad; dk0 += kDK) {
        __syncthreads();                             // kt (and, first time, qs) hazards
        {
          // a warp fills its 32 rows x 32 floats with eight 128-bit loads per lane, four in flight at a time (the
          // tile load is latency-bound): lane -> (row lane>>3 of a group of four, 16-byte piece lane&7)
          const int piece = (lane & 7) * 4;
#pragma unroll
          for (int half = 0; half < 2; ++half) {                 // two rounds of four loads: registers vs loads in flight
            float4 tmp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int row = warp * 32 + (half * 4 + i) * 4 + (lane >> 3);
              const int64_t j = j0 + row;
              const int c = dk0 + piece;
              tmp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (j < n_cols) {
                const float* src = k + j * ld_k + c;
                if (k_vec && c + 3 < d) {
                  tmp[i] = __ldg(reinterpret_cast<const float4*>(src));
                } else {
                  if (c < d) tmp[i].x = __ldg(src);
                  if (c + 1 < d) tmp[i].y = __ldg(src + 1);
                  if (c + 2 < d) tmp[i].z = __ldg(src + 2);
                  if (c + 3 < d) tmp[i].w = __ldg(src + 3);
                }
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              float* dst = kt + (warp * 32 + (half * 4 + i) * 4 + (lane >> 3)) * (kDK + 1) + piece;
              dst[0] = tmp[i].x; dst[1] = tmp[i].y; dst[2] = tmp[i].z; dst[3] = tmp[i].w;
            }
          }
        }
        __syncthreads();
#pragma unroll
        for (int d4 = 0; d4 < kDK / 4; ++d4) {
          const float k0 = kt[t * (kDK + 1) + d4 * 4 + 0];
          const float k1 = kt[t * (kDK + 1) + d4 * 4 + 1];
          const float k2 = kt[t * (kDK + 1) + d4 * 4 + 2];
          const float k3 = kt[t * (kDK + 1) + d4 * 4 + 3];
#pragma unroll
          for (int r = 0; r < kTM; ++r) {
            const float4 qv = *reinterpret_cast<const float4*>(qs + r * d_pad + dk0 + d4 * 4);
            s[r] = fmaf(qv.x, k0, s[r]);
            s[r] = fmaf(qv.y, k1, s[r]);
            s[r] = fmaf(qv.z, k2, s[r]);
            s[r] = fmaf(qv.w, k3, s[r]);
          }
        }
      }
      // ---- softmax statistics / W for this tile ----
      const int64_t j = j0 + t;
      const bool col_ok = j < n_cols;
      const float b_j = (kBwd && col_ok) ? b_col[j] : 0.f;
      const float p_j = (kBwd && avg && col_ok) ? pos_col[j] : 0.f;
#pragma unroll
      for (int r = 0; r < kTM; ++r) {
        const int64_t i = r0 + r;
        const bool ok = col_ok && i < n_rows;
        const float sv = s[r] * inv_tau;
        float e = ok ? expf(sv - shift) : 0.f;
        if (excl && j == i + diag_offset) e = 0.f;
        const bool pos = ok && ((bits[i * ld_words + (j >> 5)] >> (j & 31)) & 1u);
        if (!kBwd) {
          if (pass == 0) {
            rs[r] += (avg && pos) ? 0.f : e;
            rp[r] += pos ? sv : 0.f;
          }
        } else {
          float w = avg ? (pos ? neg2_over_c[r] + p_j : e * (a_i[r] + b_j))
                        : e * (a_i[r] + b_j) + (pos ? neg2_over_c[r] : 0.f);
          if (excl && j == i + diag_offset) w = 0.f;
          ws[t * kTM + r] = w;
        }
      }
      if (kBwd) {
        // ---- phase B: dq[kTM, dbase + slots] += W[kTM, tile] . k[tile, :] ----
        __syncthreads();
        const int jn = (int)((n_cols - j0) < kThreads ? (n_cols - j0) : kThreads);
        for (int jj = 0; jj < jn; ++jj) {
          const float4 w0 = *reinterpret_cast<const float4*>(ws + jj * kTM);
          const float4 w1 = *reinterpret_cast<const float4*>(ws + jj * kTM + 4);
          const float* kr = k + (j0 + jj) * ld_k + dbase + t;
#pragma unroll
          for (int m = 0; m < kMaxSlots; ++m) {
            const int c = dbase + t + m * kThreads;
            if (c < d) {
              const float kv = __ldg(kr + m * kThreads);
              acc[m][0] = fmaf(w0.x, kv, acc[m][0]);
              acc[m][1] = fmaf(w0.y, kv, acc[m][1]);
              acc[m][2] = fmaf(w0.z, kv, acc[m][2]);
              acc[m][3] = fmaf(w0.w, kv, acc[m][3]);
              acc[m][4] = fmaf(w1.x, kv, acc[m][4]);
              acc[m][5] = fmaf(w1.y, kv, acc[m][5]);
              acc[m][6] = fmaf(w1.z, kv, acc[m][6]);
              acc[m][7] = fmaf(w1.w, kv, acc[m][7]);
            }
          }
        }
      }
    }

    if (kBwd) {
#pragma unroll
      for (int m = 0; m < kMaxSlots; ++m) {
        const int c = dbase + t + m * kThreads;
        if (c < d) {
#pragma unroll
          for (int r = 0; r < kTM; ++r)
            if (r0 + r < n_rows) dq[(r0 + r) * ld_dq + c] = acc[m][r];
        }
      }
    }
  }

  if (!kBwd) {
    // deterministic block reduction of the per-thread (per-column-residue) partial sums
    __syncthreads();
#pragma unroll
    for (int r = 0; r < kTM; ++r) {
      const float a = warp_sum(rs[r]);
      const float b = warp_sum(rp[r]);
      if (lane == 0) {
        ws[(warp * kTM + r) * 2 + 0] = a;
        ws[(warp * kTM + r) * 2 + 1] = b;
      }
    }
    __syncthreads();
    if (t < kTM && r0 + t < n_rows) {
      float a = 0.f, b = 0.f;
      for (int wv = 0; wv < kThreads / 32; ++wv) {
        a += ws[(wv * kTM + t) * 2 + 0];
        b += ws[(wv * kTM + t) * 2 + 1];
      }
      row_sum[r0 + t] = a;
      row_pos[r0 + t] = b;
    }
  }
}

size_t small_smem_bytes(int d_pad) {
  return sizeof(float) * ((size_t)kTM * d_pad + (size_t)kThreads * (kDK + 1) + (size_t)kThreads * kTM);
}

int small_check(const float* q, const float* k, int64_t n_rows, int64_t n_cols, int64_t d, const uint32_t* bits,
                int64_t ld_words, int64_t ld_q, int64_t ld_k, float inv_tau) {
  EVK_REQUIRE(q && k && bits, "evk_mpce_small: null pointer");
  // same domain as the tcgen05 path: with the fixed shift E = exp(S - 1/tau) >= exp(-2/tau) must stay a normal
  // fp32 number, or a row of weakly aligned pairs could sum to 0 (ln 0, 1/0)
  EVK_REQUIRE(inv_tau > 0.f && inv_tau <= EVK_MAX_INV_TAU, "evk_mpce_small: 1/tau=%g outside (0, %g] (temperature >= %g)",
              inv_tau, EVK_MAX_INV_TAU, 1.0 / EVK_MAX_INV_TAU);
  EVK_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0, "evk_mpce_small: empty problem (%lld x %lld x %lld)",
              (long long)n_rows, (long long)n_cols, (long long)d);
  EVK_REQUIRE(d <= 4096, "evk_mpce_small: d=%lld > 4096 is not supported by the small path", (long long)d);
  EVK_REQUIRE(ld_q >= d && ld_k >= d, "evk_mpce_small: row pitch smaller than d");
  EVK_REQUIRE(ld_words >= (n_cols + 31) / 32, "evk_mpce_small: ld_words too small");
  return EVK_OK;
}

}  // namespace

namespace {
int small_fwd_impl(const float* q, int64_t ld_q, const float* k, int64_t ld_k, int64_t n_rows,
                   int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words, float inv_tau,
                   int flags, int64_t diag_offset, float* row_sum, float* row_pos, int64_t batch, int64_t bs_q,
                   int64_t bs_k, int64_t bs_vec, evk_stream_t stream) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = small_check(q, k, n_rows, n_cols, d, bits, ld_words, ld_q, ld_k, inv_tau);
  if (rc != EVK_OK) return rc;
  EVK_REQUIRE(row_sum && row_pos, "evk_mpce_small_fwd: null output");
  EVK_REQUIRE(batch >= 1 && batch <= 65535, "evk_mpce_small_fwd: batch must be in 1..65535");
  const int d_pad = (int)((d + kDK - 1) / kDK * kDK);
  const size_t smem = small_smem_bytes(d_pad);
  EVK_CUDA(cudaFuncSetAttribute(small_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)((n_rows + kTM - 1) / kTM), (unsigned)batch);
  small_rows_kernel<false><<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      q, ld_q, k, ld_k, n_rows, n_cols, (int)d, d_pad, bits, ld_words, nullptr, nullptr, nullptr, inv_tau, flags,
      diag_offset, row_sum, row_pos, nullptr, 0, nullptr, nullptr, bs_q, bs_k, bs_vec, 0);
  EVK_CHECK_LAUNCH("mpce_small_fwd");
  return EVK_OK;
}
}  // namespace

extern "C" int evk_mpce_small_fwd(const float* q, int64_t ld_q, const float* k, int64_t ld_k, int64_t n_rows,
                                  int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words, float inv_tau,
                                  int flags, int64_t diag_offset, float* row_sum, float* row_pos,
                                  evk_stream_t stream) {
  return small_fwd_impl(q, ld_q, k, ld_k, n_rows, n_cols, d, bits, ld_words, inv_tau, flags, diag_offset, row_sum, row_pos,
                        1, 0, 0, 0, stream);
}

extern "C" int evk_mpce_small_fwd_batched(const float* q, int64_t ld_q, int64_t bs_q, const float* k, int64_t ld_k,
                                          int64_t bs_k, int64_t batch, int64_t n_rows, int64_t n_cols, int64_t d,
                                          const uint32_t* bits, int64_t ld_words, float inv_tau, int flags,
                                          float* row_sum, float* row_pos, int64_t bs_vec, evk_stream_t stream) {
  return small_fwd_impl(q, ld_q, k, ld_k, n_rows, n_cols, d, bits, ld_words, inv_tau, flags, 0, row_sum, row_pos, batch,
                        bs_q, bs_k, bs_vec, stream);
}

namespace {
int small_bwd_impl(const float* q, int64_t ld_q, const float* k, int64_t ld_k, int64_t n_rows,
                   int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words,
                   const int32_t* counts, const float* a_row, const float* b_col, float inv_tau,
                   int flags, int64_t diag_offset, float* dq, int64_t ld_dq, const float* pos_row,
                   const float* pos_col, int64_t batch, int64_t bs_q, int64_t bs_k, int64_t bs_vec, int64_t bs_dq,
                   evk_stream_t stream) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = small_check(q, k, n_rows, n_cols, d, bits, ld_words, ld_q, ld_k, inv_tau);
  if (rc != EVK_OK) return rc;
  EVK_REQUIRE(batch >= 1 && batch <= 65535, "evk_mpce_small_bwd: batch must be in 1..65535");
  EVK_REQUIRE(a_row && b_col && dq && ld_dq >= d, "evk_mpce_small_bwd: null pointer or ld_dq < d");
  if (flags & EVK_FLAG_AVGPOS) EVK_REQUIRE(pos_row && pos_col, "evk_mpce_small_bwd: EVK_FLAG_AVGPOS needs pos_row and pos_col");
  else EVK_REQUIRE(counts, "evk_mpce_small_bwd: counts is null");
  const int d_pad = (int)((d + kDK - 1) / kDK * kDK);
  const size_t smem = small_smem_bytes(d_pad);
  EVK_CUDA(cudaFuncSetAttribute(small_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid((unsigned)((n_rows + kTM - 1) / kTM), (unsigned)batch);
  small_rows_kernel<true><<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      q, ld_q, k, ld_k, n_rows, n_cols, (int)d, d_pad, bits, ld_words, counts, a_row, b_col, inv_tau, flags,
      diag_offset, nullptr, nullptr, dq, ld_dq, pos_row, pos_col, bs_q, bs_k, bs_vec, bs_dq);
  EVK_CHECK_LAUNCH("mpce_small_bwd");
  return EVK_OK;
}
}  // namespace

extern "C" int evk_mpce_small_bwd(const float* q, int64_t ld_q, const float* k, int64_t ld_k, int64_t n_rows,
                                  int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words,
                                  const int32_t* counts, const float* a_row, const float* b_col, float inv_tau,
                                  int flags, int64_t diag_offset, float* dq, int64_t ld_dq, const float* pos_row,
                                  const float* pos_col, evk_stream_t stream) {
  return small_bwd_impl(q, ld_q, k, ld_k, n_rows, n_cols, d, bits, ld_words, counts, a_row, b_col, inv_tau, flags,
                        diag_offset, dq, ld_dq, pos_row, pos_col, 1, 0, 0, 0, 0, stream);
}

extern "C" int evk_mpce_small_bwd_batched(const float* q, int64_t ld_q, int64_t bs_q, const float* k, int64_t ld_k,
                                          int64_t bs_k, int64_t batch, int64_t n_rows, int64_t n_cols, int64_t d,
                                          const uint32_t* bits, int64_t ld_words, const int32_t* counts,
                                          const float* a_row, const float* b_col, int64_t bs_vec, float inv_tau,
                                          int flags, float* dq, int64_t ld_dq, int64_t bs_dq, evk_stream_t stream) {
  return small_bwd_impl(q, ld_q, k, ld_k, n_rows, n_cols, d, bits, ld_words, counts, a_row, b_col, inv_tau, flags, 0, dq,
                        ld_dq, nullptr, nullptr, batch, bs_q, bs_k, bs_vec, bs_dq, stream);
}
