// K2: study/patient ids -> bit-packed positive mask (+ positives per row).
// Replaces the host numpy compare + H2D of a dense fp32 label matrix at
// models/model_pretrain_finetune_v0520.py:488-491 and :422-424/:430.
//
// Layout: bits[r, w] (uint32, row pitch ld_words), bit k of word w <=> column 32*w + k, which is
// np.packbits(M, axis=1, bitorder='little') read as little-endian uint32.  Each thread owns one
// word column: it keeps the 32 column keys of that word in registers and walks kRows rows, so
// a warp writes 128 contiguous bytes per row and the column keys are read once per kRows rows.
// Algorithmic bytes: ld_words*4 per row written + 4*(n_rows + n_cols) read; the kernel is bound
// by the n_rows*n_cols integer compares (1 ISETP + 1 predicated LOP per pair), not by HBM.
#include "evk_common.cuh"

namespace {

constexpr int kThreads = 128;   // words per CTA along a row
constexpr int kRows = 32;       // rows per CTA

template <bool kTwoKeys>
__global__ void __launch_bounds__(kThreads)
posmask_kernel(const int32_t* __restrict__ ids_row, const int32_t* __restrict__ ids2_row, int64_t n_rows,
               const int32_t* __restrict__ ids_col, const int32_t* __restrict__ ids2_col, int64_t n_cols,
               int64_t diag_offset, int clear_diag, uint32_t* __restrict__ bits, int64_t ld_words,
               int32_t* __restrict__ counts) {
  __shared__ int32_t s_row[kRows];
  __shared__ int32_t s_row2[kRows];
  const int64_t w = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  const int64_t r0 = (int64_t)blockIdx.y * kRows;
  if (threadIdx.x < kRows) {
    const int64_t r = r0 + threadIdx.x;
    s_row[threadIdx.x] = r < n_rows ? ids_row[r] : 0;
    if (kTwoKeys) s_row2[threadIdx.x] = r < n_rows ? ids2_row[r] : 0;
  }
  __syncthreads();

  const int64_t c0 = w * 32;
  const bool in_row = w < ld_words;
  // column keys of this word, and which of its 32 bits are real columns
  int32_t ck[32];
  int32_t ck2[kTwoKeys ? 32 : 1];
  uint32_t valid = 0u;
  if (in_row && c0 + 32 <= n_cols) {
    valid = 0xffffffffu;
    const int4* p = reinterpret_cast<const int4*>(ids_col + c0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int4 v = __ldg(p + q);
      ck[4 * q] = v.x; ck[4 * q + 1] = v.y; ck[4 * q + 2] = v.z; ck[4 * q + 3] = v.w;
    }
    if (kTwoKeys) {
      const int4* p2 = reinterpret_cast<const int4*>(ids2_col + c0);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int4 v = __ldg(p2 + q);
        ck2[4 * q] = v.x; ck2[4 * q + 1] = v.y; ck2[4 * q + 2] = v.z; ck2[4 * q + 3] = v.w;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const bool ok = in_row && (c0 + k < n_cols);
      ck[k] = ok ? __ldg(ids_col + c0 + k) : 0;
      if (kTwoKeys) ck2[k] = ok ? __ldg(ids2_col + c0 + k) : 0;
      valid |= ok ? (1u << k) : 0u;
    }
  }

  const int lane = threadIdx.x & 31;
  for (int rr = 0; rr < kRows; ++rr) {
    const int64_t r = r0 + rr;
    if (r >= n_rows) break;                      // uniform across the CTA
    const int32_t key = s_row[rr];
    uint32_t word = 0u;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      bool eq = (ck[k] == key);
      if (kTwoKeys) eq = eq && (ck2[k] == s_row2[rr]);
      word |= eq ? (1u << k) : 0u;
    }
    word &= valid;
    if (clear_diag) {
      const int64_t dc = r + diag_offset - c0;   // bit position of the diagonal in this word
      if (dc >= 0 && dc < 32) word &= ~(1u << (int)dc);
    }
    if (in_row) bits[r * ld_words + w] = word;
    int pc = __popc(word);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pc += __shfl_xor_sync(0xffffffffu, pc, o);
    if (lane == 0 && pc != 0) atomicAdd(counts + r, pc);
  }
}

}  // namespace

extern "C" int evk_posmask_build(const int32_t* ids_row, const int32_t* ids2_row, int64_t n_rows,
                                 const int32_t* ids_col, const int32_t* ids2_col, int64_t n_cols,
                                 int64_t diag_offset, int clear_diag, uint32_t* bits, int64_t ld_words,
                                 int32_t* counts, evk_stream_t stream) {
  EVK_REQUIRE(ids_row && ids_col && bits && counts, "evk_posmask_build: null pointer");
  EVK_REQUIRE((ids2_row == nullptr) == (ids2_col == nullptr), "evk_posmask_build: ids2_row/ids2_col must both be set or both null");
  EVK_REQUIRE(n_rows >= 0 && n_cols >= 0, "evk_posmask_build: negative size");
  EVK_REQUIRE(ld_words >= (n_cols + 31) / 32, "evk_posmask_build: ld_words=%lld < ceil(n_cols/32)", (long long)ld_words);
  EVK_REQUIRE(evk_aligned16(ids_col) && (!ids2_col || evk_aligned16(ids2_col)), "evk_posmask_build: column ids must be 16-byte aligned");
  if (n_rows == 0) return EVK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  EVK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_rows, s));
  if (ld_words == 0) return EVK_OK;
  const dim3 grid((unsigned)((ld_words + kThreads - 1) / kThreads), (unsigned)((n_rows + kRows - 1) / kRows));
  EVK_REQUIRE(grid.y <= 65535u, "evk_posmask_build: n_rows=%lld too large for one launch", (long long)n_rows);
  if (ids2_row)
    posmask_kernel<true><<<grid, kThreads, 0, s>>>(ids_row, ids2_row, n_rows, ids_col, ids2_col, n_cols, diag_offset,
                                                   clear_diag, bits, ld_words, counts);
  else
    posmask_kernel<false><<<grid, kThreads, 0, s>>>(ids_row, ids2_row, n_rows, ids_col, ids2_col, n_cols, diag_offset,
                                                    clear_diag, bits, ld_words, counts);
  EVK_CHECK_LAUNCH("posmask");
  return EVK_OK;
}
