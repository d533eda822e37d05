// K4t: E strip -> W strip, in place (bf16 mode of the backward).
//
// K3 launched through evk_mpce_fwd_store leaves E_ij = exp(S_ij - 1/tau) as bf16 in a row strip.
// Once the row / column statistics are known, the weights of the two gradient contractions are an
// elementwise function of that strip:
//     W_ij = E_ij (a_i + b_j) - 2 M_ij / c_i          a_i = 1/R_i,  b_j = 1/C_j
// (the softmax - target terms of both cross entropies, models/model_pretrain_finetune_v0520.py:501-503
// and :443, times N resp. M'), so the S tiles need not be recomputed on the tensor cores: the pass is
// HBM-bound instead (reads and writes the strip once: 4 bytes per pair of the N x N problem, plus one
// mask bit per pair), and the step executes 6 N^2 D FLOP instead of 8 N^2 D.
//
// Two launches: the dense scale pass over every entry, then a sparse pass over the positives.
#include "evk_common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;
constexpr int kPatchWarps = 8;

__device__ __forceinline__ uint32_t scale_pair(uint32_t e2, float s0, float s1) {
  const float e0 = __uint_as_float(e2 << 16), e1 = __uint_as_float(e2 & 0xffff0000u);
  __nv_bfloat162 t = __floats2bfloat162_rn(e0 * s0, e1 * s1);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Pass 1 (all entries): strip <- bf16(E (a_i + b_j)).  Thread = 8 consecutive columns (one 128-bit word)
// walking down the rows: b_j for its columns stays in registers; per row one 128-bit load + store.
__global__ void __launch_bounds__(kThreads)
w_scale_kernel(uint4* __restrict__ strip, int64_t ld_vec, int64_t n_rows, int64_t n_cols,
               const float* __restrict__ a_row, const float* __restrict__ b_col) {
  const int64_t c8 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t j0 = c8 * 8;
  if (j0 >= n_cols) return;
  float b[8];
  if (j0 + 8 <= n_cols) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(b_col + j0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(b_col + j0 + 4));
    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) b[k] = (j0 + k < n_cols) ? __ldg(b_col + j0 + k) : 0.f;
  }
  const int64_t step = gridDim.y;
  for (int64_t r0 = blockIdx.y; r0 < n_rows; r0 += step * kUnroll) {
    uint4 v[kUnroll];
    float a[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t r = r0 + u * step;
      if (r < n_rows) {
        v[u] = __ldcs(strip + r * ld_vec + c8);
        a[u] = __ldg(a_row + r);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t r = r0 + u * step;
      if (r < n_rows) {
        const float ai = a[u];
        uint4 o;
        o.x = scale_pair(v[u].x, ai + b[0], ai + b[1]);
        o.y = scale_pair(v[u].y, ai + b[2], ai + b[3]);
        o.z = scale_pair(v[u].z, ai + b[4], ai + b[5]);
        o.w = scale_pair(v[u].w, ai + b[6], ai + b[7]);
        strip[r * ld_vec + c8] = o;
      }
    }
  }
}

__device__ __forceinline__ float dot2_bf16(uint32_t a, uint32_t b, float s) {
  s = fmaf(__uint_as_float(a << 16), __uint_as_float(b << 16), s);
  return fmaf(__uint_as_float(a & 0xffff0000u), __uint_as_float(b & 0xffff0000u), s);
}

__device__ __forceinline__ float dot8_bf16(const uint4& a, const uint4& b, float s) {
  s = dot2_bf16(a.x, b.x, s);
  s = dot2_bf16(a.y, b.y, s);
  s = dot2_bf16(a.z, b.z, s);
  return dot2_bf16(a.w, b.w, s);
}

// Pass 2 (positives only, about c_i per row): W_ij = E_ij (a_i + b_j) - 2 / c_i.
// Positives are where softmax and target cancel (both terms are O(1/c) for a well-aligned pair, most of
// all at cold temperatures), so a bf16-rounded E is not good enough THERE: with the operands at hand the
// entry is recomputed from S_ij in fp32 (warp-cooperative dot product over the bf16 operands the tensor
// cores saw) and rounded to bf16 once, exactly like K4a does for every entry.
// One warp per row.  Latency-bound, so everything independent is issued at once: the row's mask words
// (128-bit loads, 512 words per sweep), the query row (kept in registers, kQVec 128-bit words per lane),
// and per positive all of the key row's loads.
// Pass 2, list form: thread = one listed positive.  pos_idx / pos_dot come from K2 and evk_mpce_pos_logits
// (forward, side stream), so neither the mask nor the operands are touched here.
__global__ void __launch_bounds__(256)
w_pos_list_kernel(__nv_bfloat16* __restrict__ strip, int64_t ld_e, int64_t n_rows, const int32_t* __restrict__ counts,
                  const float* __restrict__ a_row, const float* __restrict__ b_col, const int32_t* __restrict__ pos_idx,
                  const float* __restrict__ pos_dot, int pos_slots, float inv_tau) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / pos_slots;
  const int s = (int)(t - i * pos_slots);
  if (i >= n_rows) return;
  const int c = __ldg(counts + i);
  if (s >= c || c > pos_slots) return;                 // rows with more positives than slots: mask-scan kernel
  const int64_t j = __ldg(pos_idx + t);
  const float c1 = inv_tau * 1.4426950408889634f;
  const float e = exp2f(fmaf(__ldg(pos_dot + t), c1, -c1));
  strip[i * ld_e + j] = __float2bfloat16_rn(fmaf(e, __ldg(a_row + i) + __ldg(b_col + j), -2.f / (float)c));
}

template <int kQVec>
__global__ void __launch_bounds__(kPatchWarps * 32)
w_pos_kernel(__nv_bfloat16* __restrict__ strip, int64_t ld_e, int64_t n_rows, int64_t n_cols,
             const uint32_t* __restrict__ bits, int64_t ld_words, const int32_t* __restrict__ counts,
             const float* __restrict__ a_row, const float* __restrict__ b_col,
             const __nv_bfloat16* __restrict__ q_hi, int64_t ld_q, const __nv_bfloat16* __restrict__ k_hi,
             int64_t ld_k, int d_vec, float inv_tau, int mask_vec_ok, int skip_upto) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPatchWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPatchWarps;
  const int64_t words = (n_cols + 31) >> 5;
  const float c1 = inv_tau * 1.4426950408889634f;
  // a warp takes 32 consecutive rows at a time and keeps those that still need the mask scan (with the
  // positive lists at hand that is none, or the few rows with more positives than list slots)
  const int span = skip_upto > 0 ? 32 : 1;               // without lists every row is a candidate: one per warp
  for (int64_t rb = warp0 * span; rb < n_rows; rb += nwarps * span) {
   const int cnt_l = (lane < span && rb + lane < n_rows) ? __ldg(counts + rb + lane) : 0;
   uint32_t need = __ballot_sync(0xffffffffu, cnt_l > skip_upto);
   while (need) {
    const int rl = __ffs(need) - 1;
    need &= need - 1;
    const int64_t i = rb + rl;
    const int cnt = __shfl_sync(0xffffffffu, cnt_l, rl);
    const uint32_t* mrow = bits + i * ld_words;
    const float ai = __ldg(a_row + i);
    const float pc = -2.f / (float)max(cnt, 1);
    __nv_bfloat16* srow = strip + i * ld_e;
    uint4 qv[kQVec];
    if (q_hi) {
      const uint4* qa = reinterpret_cast<const uint4*>(q_hi + i * ld_q);
#pragma unroll
      for (int t = 0; t < kQVec; ++t) qv[t] = (lane + 32 * t < d_vec) ? __ldg(qa + lane + 32 * t) : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int64_t w0 = 0; w0 < words; w0 += 512) {
      uint32_t mv[16];                                   // lane holds words w0 + 128 b + 4 lane + e
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int64_t w = w0 + 128 * b + 4 * lane;
        if (mask_vec_ok && w + 4 <= ld_words) {          // rows are zero beyond n_cols, up to ld_words
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(mrow + w));
          mv[4 * b] = v.x; mv[4 * b + 1] = v.y; mv[4 * b + 2] = v.z; mv[4 * b + 3] = v.w;
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) mv[4 * b + e] = (w + e < words) ? __ldg(mrow + w + e) : 0u;
        }
      }
      uint32_t hb = 0u;                                  // which of the 16 word slots hold a positive in some lane
#pragma unroll
      for (int b = 0; b < 16; ++b)
        if (__any_sync(0xffffffffu, mv[b] != 0u)) hb |= 1u << b;
      while (hb) {                                       // rare path, kept out of the unrolled code
        const int b = __ffs(hb) - 1;
        hb &= hb - 1;
        uint32_t mine = 0u;
#pragma unroll
        for (int bb = 0; bb < 16; ++bb)
          if (bb == b) mine = mv[bb];
        uint32_t any = __ballot_sync(0xffffffffu, mine != 0u);
        while (any) {
          const int src = __ffs(any) - 1;
          any &= any - 1;
          uint32_t mw = __shfl_sync(0xffffffffu, mine, src);
          const int64_t jbase = (w0 + 128 * (b >> 2) + 4 * src + (b & 3)) << 5;
          while (mw) {
            const int bit = __ffs(mw) - 1;
            mw &= mw - 1;
            const int64_t j = jbase + bit;
            if (q_hi) {
              const uint4* kb = reinterpret_cast<const uint4*>(k_hi + j * ld_k);
              uint4 kv[kQVec];
#pragma unroll
              for (int t = 0; t < kQVec; ++t)
                kv[t] = (lane + 32 * t < d_vec) ? __ldg(kb + lane + 32 * t) : make_uint4(0u, 0u, 0u, 0u);
              float sdot = 0.f;
#pragma unroll
              for (int t = 0; t < kQVec; ++t) sdot = dot8_bf16(qv[t], kv[t], sdot);
              sdot = warp_sum(sdot);
              if (lane == 0) srow[j] = __float2bfloat16_rn(fmaf(exp2f(fmaf(sdot, c1, -c1)), ai + __ldg(b_col + j), pc));
            } else if (lane == 0) {
              srow[j] = __float2bfloat16_rn(__bfloat162float(srow[j]) + pc);
            }
          }
        }
      }
    }
   }
  }
}

// Pass 2 without the dense mask (bits == NULL): rows with more positives than list slots find them by scanning
// the column ids.  A warp takes 32 consecutive rows and keeps those with counts > skip_upto (normally none).
template <int kQVec>
__global__ void __launch_bounds__(kPatchWarps * 32)
w_pos_ids_kernel(__nv_bfloat16* __restrict__ strip, int64_t ld_e, int64_t n_rows, int64_t n_cols,
                 const int32_t* __restrict__ ids_row, const int32_t* __restrict__ ids2_row,
                 const int32_t* __restrict__ ids_col, const int32_t* __restrict__ ids2_col, int64_t diag_offset,
                 int clear_diag, const int32_t* __restrict__ counts, const float* __restrict__ a_row,
                 const float* __restrict__ b_col, const __nv_bfloat16* __restrict__ q_hi, int64_t ld_q,
                 const __nv_bfloat16* __restrict__ k_hi, int64_t ld_k, int d_vec, float inv_tau, int skip_upto) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kPatchWarps + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kPatchWarps;
  const float c1 = inv_tau * 1.4426950408889634f;
  for (int64_t rb = warp0 * 32; rb < n_rows; rb += nwarps * 32) {
    const int cnt_l = (rb + lane < n_rows) ? __ldg(counts + rb + lane) : 0;
    uint32_t need = __ballot_sync(0xffffffffu, cnt_l > skip_upto);
    while (need) {
      const int rl = __ffs(need) - 1;
      need &= need - 1;
      const int64_t i = rb + rl;
      const int cnt = __shfl_sync(0xffffffffu, cnt_l, rl);
      const int32_t key = __ldg(ids_row + i);
      const int32_t key2 = ids2_row ? __ldg(ids2_row + i) : 0;
      const int64_t diag = clear_diag ? i + diag_offset : -1;
      const float ai = __ldg(a_row + i);
      const float pc = -2.f / (float)max(cnt, 1);
      __nv_bfloat16* srow = strip + i * ld_e;
      const uint4* qa = reinterpret_cast<const uint4*>(q_hi + i * ld_q);
      uint4 qv[kQVec];
#pragma unroll
      for (int t = 0; t < kQVec; ++t) qv[t] = (lane + 32 * t < d_vec) ? __ldg(qa + lane + 32 * t) : make_uint4(0u, 0u, 0u, 0u);
      for (int64_t j0 = 0; j0 < n_cols; j0 += 32) {
        const int64_t jl = j0 + lane;
        bool hit = jl < n_cols && __ldg(ids_col + jl) == key && jl != diag;
        if (hit && ids2_col) hit = __ldg(ids2_col + jl) == key2;
        uint32_t any = __ballot_sync(0xffffffffu, hit);
        while (any) {
          const int b = __ffs(any) - 1;
          any &= any - 1;
          const int64_t j = j0 + b;
          const uint4* kb = reinterpret_cast<const uint4*>(k_hi + j * ld_k);
          float sdot = 0.f;
#pragma unroll
          for (int t = 0; t < kQVec; ++t)
            if (lane + 32 * t < d_vec) sdot = dot8_bf16(qv[t], __ldg(kb + lane + 32 * t), sdot);
          sdot = warp_sum(sdot);
          if (lane == 0) srow[j] = __float2bfloat16_rn(fmaf(exp2f(fmaf(sdot, c1, -c1)), ai + __ldg(b_col + j), pc));
        }
      }
    }
  }
}

}  // namespace

extern "C" int evk_mpce_w_from_e(void* strip, int64_t ld_e, int64_t n_rows, int64_t n_cols, const uint32_t* bits,
                                 int64_t ld_words, const int32_t* counts, const float* a_row, const float* b_col,
                                 const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t d, float inv_tau,
                                 const int32_t* pos_idx, const float* pos_dot, int pos_slots, const int32_t* ids_row,
                                 const int32_t* ids2_row, const int32_t* ids_col, const int32_t* ids2_col,
                                 int64_t diag_offset, int clear_diag, evk_stream_t stream) {
  EVK_REQUIRE(strip && counts && a_row && b_col, "evk_mpce_w_from_e: null pointer");
  EVK_REQUIRE(bits || (ids_row && ids_col && pos_idx && pos_dot && q_hi && k_hi),
              "evk_mpce_w_from_e: without the dense mask (bits == NULL) the positive lists, the operands and the ids are required");
  EVK_REQUIRE((ids2_row == nullptr) == (ids2_col == nullptr), "evk_mpce_w_from_e: ids2_row/ids2_col must both be set or both null");
  EVK_REQUIRE(n_rows > 0 && n_cols > 0, "evk_mpce_w_from_e: empty problem");
  EVK_REQUIRE(evk_aligned16(strip) && ld_e % 8 == 0 && ld_e >= ((n_cols + 7) / 8) * 8,
              "evk_mpce_w_from_e: strip needs a 16-byte aligned base and ld_e %% 8 == 0, ld_e >= n_cols rounded up to 8");
  EVK_REQUIRE(evk_aligned16(b_col), "evk_mpce_w_from_e: b_col must be 16-byte aligned");
  EVK_REQUIRE(!bits || ld_words >= (n_cols + 31) / 32, "evk_mpce_w_from_e: ld_words too small");
  if (q_hi) {
    EVK_REQUIRE(d <= 4096, "evk_mpce_w_from_e: d=%lld > 4096 is not supported by the exact-positives pass", (long long)d);
    EVK_REQUIRE(k_hi && d > 0 && ld_q % 8 == 0 && ld_k % 8 == 0 && ld_q >= d && ld_k >= d && evk_aligned16(q_hi) &&
                    evk_aligned16(k_hi) && inv_tau > 0.f,
                "evk_mpce_w_from_e: the operands of the exact positive entries need 16-byte aligned rows (ld %% 8 == 0)");
  }
  const int64_t vecs = (n_cols + 7) / 8;
  const int64_t gx = (vecs + kThreads - 1) / kThreads;
  int64_t gy = ((int64_t)evk_sm_count() * 8 + gx - 1) / gx;
  if (gy > n_rows) gy = n_rows;
  if (gy > 65535) gy = 65535;
  EVK_REQUIRE(gx <= 0x7fffffff, "evk_mpce_w_from_e: too many columns");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  w_scale_kernel<<<dim3((unsigned)gx, (unsigned)gy), kThreads, 0, s>>>(static_cast<uint4*>(strip), ld_e / 8, n_rows, n_cols,
                                                                      a_row, b_col);
  EVK_CHECK_LAUNCH("w_scale");

  int skip_upto = 0;                                  // rows with counts <= skip_upto need no mask scan
  if (pos_idx && pos_dot) {
    EVK_REQUIRE(pos_slots >= 1 && pos_slots <= 64, "evk_mpce_w_from_e: pos_slots must be in 1..64");
    const int64_t threads = n_rows * pos_slots;
    w_pos_list_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(static_cast<__nv_bfloat16*>(strip), ld_e, n_rows, counts,
                                                                       a_row, b_col, pos_idx, pos_dot, pos_slots, inv_tau);
    EVK_CHECK_LAUNCH("w_pos_list");
    skip_upto = pos_slots;
  }
  // with lists: 32 rows per warp-iteration and nearly every warp finds nothing to do
  const int64_t blocks = skip_upto > 0 ? (n_rows + 32 * kPatchWarps - 1) / (32 * kPatchWarps)
                                       : (n_rows + kPatchWarps - 1) / kPatchWarps;
  const int d_vec = (int)((d + 7) / 8);
  const int mask_vec_ok = (bits && ld_words % 4 == 0 && evk_aligned16(bits)) ? 1 : 0;
  auto* sp = static_cast<__nv_bfloat16*>(strip);
  auto* qp = static_cast<const __nv_bfloat16*>(q_hi);
  auto* kp = static_cast<const __nv_bfloat16*>(k_hi);
  if (!bits) {                                        // overflow rows find their positives by scanning the ids
    if (d_vec <= 128)
      w_pos_ids_kernel<4><<<(unsigned)blocks, kPatchWarps * 32, 0, s>>>(sp, ld_e, n_rows, n_cols, ids_row, ids2_row, ids_col,
                                                                     ids2_col, diag_offset, clear_diag, counts, a_row, b_col,
                                                                     qp, ld_q, kp, ld_k, d_vec, inv_tau, skip_upto);
    else
      w_pos_ids_kernel<16><<<(unsigned)blocks, kPatchWarps * 32, 0, s>>>(sp, ld_e, n_rows, n_cols, ids_row, ids2_row, ids_col,
                                                                      ids2_col, diag_offset, clear_diag, counts, a_row, b_col,
                                                                      qp, ld_q, kp, ld_k, d_vec, inv_tau, skip_upto);
    EVK_CHECK_LAUNCH("w_pos_ids");
    return EVK_OK;
  }
  if (d_vec <= 128)
    w_pos_kernel<4><<<(unsigned)blocks, kPatchWarps * 32, 0, s>>>(sp, ld_e, n_rows, n_cols, bits, ld_words, counts, a_row,
                                                               b_col, qp, ld_q, kp, ld_k, d_vec, inv_tau, mask_vec_ok, skip_upto);
  else
    w_pos_kernel<16><<<(unsigned)blocks, kPatchWarps * 32, 0, s>>>(sp, ld_e, n_rows, n_cols, bits, ld_words, counts, a_row,
                                                                b_col, qp, ld_q, kp, ld_k, d_vec, inv_tau, mask_vec_ok, skip_upto);
  EVK_CHECK_LAUNCH("w_pos");
  return EVK_OK;
}
