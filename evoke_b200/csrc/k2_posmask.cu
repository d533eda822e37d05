// K2: study/patient ids -> bit-packed positive mask (+ positives per row).
// Replaces the host numpy compare + H2D of a dense fp32 label matrix at
// models/model_pretrain_finetune_v0520.py:488-491 and :422-424/:430.
//
// Layout: bits[r, w] (uint32, row pitch ld_words), bit k of word w <=> column 32*w + k, which is
// np.packbits(M, axis=1, bitorder='little') read as little-endian uint32.
//
// The mask is sparse (a study has a handful of views), so the N^2 compares of the reference are not
// executed.  A CTA owns a block of 256 columns x kRowsPerCta rows: it hashes the block's column keys
// into an open-addressing multiset in shared memory (duplicates take consecutive slots), then every
// lane looks its own row key up - each match is one bit set in a per-warp [32 rows x 8 words] tile in
// shared memory - and the warp streams the tile out with 128-bit stores.  Work per row and column block
// is a probe sequence (about two slots) instead of 256 compares, which leaves the kernel bound by
// writing the mask: algorithmic bytes ld_words*4 per row written + 4*(n_rows + n_cols) read.
// Degenerate inputs (every key equal) degrade gracefully to one compare per pair.  Exact: keys are
// compared in full, the hash only picks the starting slot.
//
// bits == NULL builds counts and lists only: the bf16 mode of the large path needs nothing of size N^2 from the ids
// (positive sums and exact W entries come from the lists; rows with more positives than slots are handled by an id
// scan, see evk_mpce_pos_from_lists / evk_mpce_w_from_e).
// Optional second output: pos_idx[r, s] = column of the s-th positive of row r (s < pos_slots; the order
// within a row is unspecified, rows with more positives keep only the first pos_slots - counts[r] tells).
// The O(N) consumers of the positives (exact W entries, K4t) then need no scan of the N^2/8-byte mask.
#include "evk_common.cuh"
#include "peer_sync.cuh"

namespace {

constexpr int kThreads = 256;                 // 8 warps; one column key inserted per thread
constexpr int kWarps = kThreads / 32;
constexpr int kColsPerCta = 256;              // 8 words
constexpr int kWordsPerCta = kColsPerCta / 32;
constexpr int kSlots = 512;                   // load factor <= 0.5
constexpr int kRowsPerCta = 512;              // 2 x 32 rows per warp: the table build is 1/3 of a CTA's work
constexpr int kTilePitch = 12;                // words per tile row (16-byte aligned rows, banks staggered)

__device__ __forceinline__ uint32_t slot_hash(int32_t k, int32_t k2) {
  uint32_t h = (uint32_t)k * 0x9E3779B1u;
  h ^= (uint32_t)k2 * 0x85EBCA6Bu;
  h ^= h >> 15;
  return h & (kSlots - 1);
}

template <bool kTwoKeys>
__global__ void __launch_bounds__(kThreads)
posmask_kernel(const int32_t* __restrict__ ids_row, const int32_t* __restrict__ ids2_row, int64_t n_rows,
               const int32_t* __restrict__ ids_col, const int32_t* __restrict__ ids2_col, int64_t n_cols,
               int64_t diag_offset, int clear_diag, uint32_t* __restrict__ bits, int64_t ld_words,
               int32_t* __restrict__ counts, int vec_ok, int32_t* __restrict__ pos_idx, int pos_slots,
               const PeerSyncDev sync) {
  __shared__ int32_t t_col[kSlots];                      // column (0..1023) held by the slot, -1 = empty
  __shared__ int32_t t_key[kSlots];
  __shared__ int32_t t_key2[kTwoKeys ? kSlots : 1];
  __shared__ __align__(16) uint32_t tile[kWarps][32 * kTilePitch];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t c0 = (int64_t)blockIdx.x * kColsPerCta;
  const int64_t w0 = c0 >> 5;
  const int64_t r_cta = (int64_t)blockIdx.y * kRowsPerCta;

  peer_sync_block(sync, blockIdx.x == 0 && blockIdx.y == 0);     // sharded: the column ids come from the peers
  for (int s = threadIdx.x; s < kSlots; s += kThreads) t_col[s] = -1;
  for (int k = lane; k < 32 * kTilePitch; k += 32) tile[warp][k] = 0u;
  __syncthreads();
  for (int c = threadIdx.x; c < kColsPerCta; c += kThreads) {
    if (c0 + c < n_cols) {
      const int32_t key = __ldg(ids_col + c0 + c);
      const int32_t key2 = kTwoKeys ? __ldg(ids2_col + c0 + c) : 0;
      uint32_t s = slot_hash(key, key2);
      while (atomicCAS(&t_col[s], -1, c) != -1) s = (s + 1) & (kSlots - 1);
      t_key[s] = key;
      if (kTwoKeys) t_key2[s] = key2;
    }
  }
  __syncthreads();

  uint32_t* my = tile[warp];
  for (int64_t rb = r_cta + warp * 32; rb < r_cta + kRowsPerCta && rb < n_rows; rb += kWarps * 32) {
    const int64_t r = rb + lane;
    if (r < n_rows) {
      const int32_t key = __ldg(ids_row + r);
      const int32_t key2 = kTwoKeys ? __ldg(ids2_row + r) : 0;
      const int64_t diag = clear_diag ? r + diag_offset - c0 : -1;    // column of this block to leave clear
      uint32_t s = slot_hash(key, key2);
      for (int probes = 0; probes < kSlots; ++probes) {
        const int32_t c = t_col[s];
        if (c < 0) break;
        if (t_key[s] == key && (!kTwoKeys || t_key2[s] == key2) && (int64_t)c != diag) {
          if (bits) my[lane * kTilePitch + (c >> 5)] |= 1u << (c & 31);
          // positives are rare: one atomic each; its return value is the entry's slot in the row's list
          const int slot = atomicAdd(counts + r, 1);
          if (pos_idx && slot < pos_slots) pos_idx[r * pos_slots + slot] = (int32_t)(c0 + c);
        }
        s = (s + 1) & (kSlots - 1);
      }
    }
    __syncwarp();
    if (!bits) continue;                         // list-only build: nothing of size N^2 is written
    // stream the 32 x 8-word tile out (and clear it for the next row group): 32 bytes per row
    if (vec_ok) {
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int row = it * 16 + (lane >> 1), wq = (lane & 1) * 4;
        uint4* src = reinterpret_cast<uint4*>(my + row * kTilePitch + wq);
        const uint4 v = *src;
        *src = make_uint4(0u, 0u, 0u, 0u);
        if (rb + row < n_rows && w0 + wq < ld_words)
          *reinterpret_cast<uint4*>(bits + (rb + row) * ld_words + w0 + wq) = v;
      }
    } else {
      for (int k = lane; k < 32 * kWordsPerCta; k += 32) {
        const int row = k / kWordsPerCta, wq = k % kWordsPerCta;
        const uint32_t v = my[row * kTilePitch + wq];
        my[row * kTilePitch + wq] = 0u;
        if (rb + row < n_rows && w0 + wq < ld_words) bits[(rb + row) * ld_words + w0 + wq] = v;
      }
    }
    __syncwarp();
  }
}

}  // namespace

extern "C" int evk_posmask_build(const int32_t* ids_row, const int32_t* ids2_row, int64_t n_rows,
                                 const int32_t* ids_col, const int32_t* ids2_col, int64_t n_cols,
                                 int64_t diag_offset, int clear_diag, uint32_t* bits, int64_t ld_words,
                                 int32_t* counts, int32_t* pos_idx, int pos_slots, int counts_zeroed,
                                 const evk_peer_sync_t* sync, evk_stream_t stream) {
  EVK_REQUIRE(ids_row && ids_col && counts && (bits || pos_idx), "evk_posmask_build: null pointer (bits may be NULL only with pos_idx)");
  EVK_REQUIRE((ids2_row == nullptr) == (ids2_col == nullptr), "evk_posmask_build: ids2_row/ids2_col must both be set or both null");
  EVK_REQUIRE(n_rows >= 0 && n_cols >= 0, "evk_posmask_build: negative size");
  if (!bits) ld_words = (n_cols + 31) / 32;      // list-only build: ld_words only sizes the grid
  EVK_REQUIRE(ld_words >= (n_cols + 31) / 32, "evk_posmask_build: ld_words=%lld < ceil(n_cols/32)", (long long)ld_words);
  EVK_REQUIRE(!pos_idx || (pos_slots >= 1 && pos_slots <= 64), "evk_posmask_build: pos_slots must be in 1..64");
  if (n_rows == 0) return EVK_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PeerSyncDev ps;
  int rc = peer_sync_from_host(sync, ps);
  if (rc != EVK_OK) return rc;
  if (!counts_zeroed) EVK_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * n_rows, s));
  if (ld_words == 0) return EVK_OK;
  const dim3 grid((unsigned)((ld_words * 32 + kColsPerCta - 1) / kColsPerCta), (unsigned)((n_rows + kRowsPerCta - 1) / kRowsPerCta));
  EVK_REQUIRE(grid.y <= 65535u, "evk_posmask_build: n_rows=%lld too large for one launch", (long long)n_rows);
  const int vec_ok = (bits && ld_words % 4 == 0 && evk_aligned16(bits)) ? 1 : 0;
  if (ids2_row)
    posmask_kernel<true><<<grid, kThreads, 0, s>>>(ids_row, ids2_row, n_rows, ids_col, ids2_col, n_cols, diag_offset,
                                                   clear_diag, bits, ld_words, counts, vec_ok, pos_idx, pos_slots, ps);
  else
    posmask_kernel<false><<<grid, kThreads, 0, s>>>(ids_row, ids2_row, n_rows, ids_col, ids2_col, n_cols, diag_offset,
                                                    clear_diag, bits, ld_words, counts, vec_ok, pos_idx, pos_slots, ps);
  EVK_CHECK_LAUNCH("posmask");
  return EVK_OK;
}
