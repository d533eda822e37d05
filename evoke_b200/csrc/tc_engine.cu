// Large path: one warp-specialised, persistent tcgen05 main loop (TMA -> shared memory ->
// tcgen05.mma -> TMEM, double-buffered accumulators) with five epilogues:
//
//   EPI_FWD      K3  S = Qhat Khat^T tile lives only in TMEM; the epilogue turns it into
//                    exp-sums per row / per column and the positive-logit sums (O(N) outputs).
//                    Replaces 2x mm + 2x `/temp` + 2x cross_entropy, v0520.py:499-502 (and :437-443).
//   EPI_FWD_E    K3  the same, and E = exp(S - 1/tau) leaves as a bf16 row strip (per-warp staging, TMA stores):
//                    the bf16-mode backward then needs no second sweep over the similarity tiles.
//   EPI_BWD_W    K4a recompute the S tile, form W = E (a_i + b_j) - 2 M / c_i, store it as bf16 (hi, lo)
//                    (TMA store) into the row-strip workspace: fp32-parity mode.
//   EPI_GEMM_TMA K4b dQhat = W X / dKhat = W^T X with MN-major operands, split-K; fp32 tiles leave through
//                    shared memory and TMA reduce-add (L2), or TMA stores into the owning rank's buffer on
//                    another GPU (the reduce-scatter of the sharded path, fused into the GEMM, tile by tile).
//   EPI_GEMM     K4b with per-lane red.global.add.v4 / st (single-CTA kernels, EVK_GEMM_EPI=red).
//                    K4t/K4a + K4b replace autograd's 4 mm + N^2 elementwise passes.
//
// Tile 128 x 256 (UMMA M=128, N=256, K=16), BK = 64 bf16 = one 128-byte swizzle row, so a
// stage is 16 KiB (A) + 32 KiB (B) (16 + 16 KiB per CTA of a pair).  TMEM: 2 accumulator stages x 256 fp32
// columns = all 512.  Warps: 0..7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31, i.e. one tile row per
// thread, and half (w/4) of the tile's columns), 8 = TMA producer, 9 = MMA issuer + TMEM allocator.  The two
// issuing roles run their loops with the WHOLE warp and issue under elect.sync, so that descriptors and barrier
// addresses stay in uniform registers (see elect_one() in tc_ptx.cuh).
//
// Split-bf16 (fp32-parity) mode is just a longer K loop over "segments": S = hi.hi + hi.lo + lo.hi.
#include "evk_common.cuh"
#include "tc_ptx.cuh"

#include <stdlib.h>
#include <string.h>

namespace {

using namespace tc;

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int kABytes = BM * BK * 2;          // 16 KiB: this CTA's 128 rows of A
constexpr int kEpiWarps = 8;                  // two per SMSP; warp w reads TMEM lanes 32*(w%4)..+31
// (+ one TMA producer warp + one MMA issuer warp per CTA; the sixteen-warp K3 variants are chosen at launch, see
// mpce_fwd_impl / k3_variant())
constexpr int kTmemCols = 512;
constexpr int kK3DefaultVariant = 0;          // see k3_variant()
constexpr int kMaxSegs = 3;

enum { EPI_FWD = 0, EPI_BWD_W = 1, EPI_GEMM = 2, EPI_FWD_E = 3, EPI_GEMM_TMA = 4 };
// Bits of TcParams::flags above the public EVK_FLAG_* set.  They are set by this file only: every extern "C" entry
// point masks the caller's flags with EVK_FLAG_PUBLIC_MASK first.
enum {
  kIntNoEpiPipe = 0x8000,  // K3 epilogue without the software-pipelined TMEM loads (A/B, EVK_K3_PIPE=0)
  kIntDropOut = 0x100,   // evk_tc_gemm_probe(variant & 8): run the main loop, drop the output (main-loop rate probe)
  kIntStore = 0x200,     // gradient contraction: whole-K units, plain stores instead of reduce-add
  kIntBf16Out = 0x2000,  // gradient contraction, store mode: bf16 [32 x 64] output boxes
  kIntVecMask = 0x4000,  // K3: the mask rows allow 128-bit loads
};
// EPI_FWD_E: K3 that also stores E as bf16.  EPI_GEMM_TMA: K4b whose fp32 tiles leave through shared memory and
// TMA (store, or reduce-add in L2): whole 128-byte lines per request instead of 16-byte red.add / st per lane,
// which is what peer memory over NVLink needs and also takes the accumulation off the epilogue warps.

// B bytes per stage held by ONE CTA: the whole 256-column tile, or half of it in CTA-pair mode
template <bool CTA2> constexpr int b_bytes() { return (CTA2 ? BN / 2 : BN) * BK * 2; }
template <bool CTA2> constexpr int stage_bytes() { return kABytes + b_bytes<CTA2>(); }

struct alignas(64) TcParams {
  CUtensorMap a_map[kMaxSegs];
  CUtensorMap b_map[kMaxSegs];
  CUtensorMap out_map[2];        // EPI_BWD_W: W hi / lo stores
  int m_tiles, n_tiles, splits;  // m_tiles counts 128-row tiles (1-CTA) or 256-row pair tiles (CTA2)
  int num_segs, kb_per_seg;      // k-blocks (of BK) per segment
  int64_t n_rows, n_cols;        // logical extent of the M x N problem (masking)
  float inv_tau;
  int flags;
  int64_t diag_offset;
  // EPI_FWD / EPI_BWD_W
  const uint32_t* bits; int64_t ld_words;
  float* row_sum_part; float* row_pos_part; int64_t ld_rowpart;
  float* col_sum_part; int64_t ld_colpart;
  const int32_t* counts; const float* a_row; const float* b_col;
  // EPI_FWD*, sharded: the key rows arrive per source rank while the sweep runs (evk_peer_push_shard).  The tiles
  // are then visited column-block-major starting at this rank's own columns (rot_tiles), and the producer waits for
  // landed[source] >= *step before its first load from a source's columns.
  const uint32_t* landed; const int* step; int* error; int64_t cols_per_source; int rot_tiles;
  int landed_per_step;           // landed[s] counts the push CTAs of source s: it has landed at landed_per_step * step
  int rot_m_tiles;               // EPI_GEMM_TMA scatter: first row block (pair tile) to visit, see Sched
  int cta_limit;                 // > 0: launch at most this many CTAs (two contractions sharing the GPU side by side)
  // EPI_GEMM
  float* out; int64_t ld_out; float alpha;
  // EPI_GEMM, reduce-scatter fused into the epilogue: output row i belongs to rank i / rows_per_owner and
  // is accumulated (red.add over NVLink) into that rank's buffer at row i % rows_per_owner
  float* out_peer[16]; int64_t rows_per_owner;
  // EPI_GEMM_TMA: fp32 output maps (box 32 rows x 32 columns), one per owner (entry 0 for a local output)
  CUtensorMap peer_map[16];
  // descriptor bases (see tc_ptx.cuh), filled by the host so the probe can try variants
  uint64_t desc_a, desc_b;
  uint32_t idesc;
  // L2 policies: operands reused by every CTA are kept, the N^2 W stream is not
  uint64_t policy_a, policy_b, policy_out;
};

template <int EPI>
constexpr int epi_smem_bytes() {
  return EPI == EPI_FWD ? 2 * 4 * BN * 4
         : (EPI == EPI_BWD_W ? kEpiWarps * 8192
            : (EPI == EPI_FWD_E ? 2 * 4 * BN * 4 : (EPI == EPI_GEMM_TMA ? kEpiWarps * 8192 : 0)));
}
// Split-role epilogue of K3 with the strip store (EPI_FWD_E, sixteen epilogue warps, two staging boxes): warps 0-7
// turn the accumulator into the statistics (exp, row / column sums, positives), warps 8-15 into the bf16 E strip
// (exp, pack, swizzled staging, TMA stores).  Both read the same TMEM tile (TMEM reads are free, the exp is done
// twice); the tile's epilogue time becomes max(statistics, store) instead of their sum.
template <int EPI, int EW, int SB> constexpr bool split_epi() { return EPI == EPI_FWD_E && EW == 16 && SB == 2; }
// warps that stage E boxes
template <int EPI, int EW, int SB> constexpr int store_warps() { return EPI != EPI_FWD_E ? 0 : (split_epi<EPI, EW, SB>() ? 8 : EW); }
// EW epilogue warps, SB staging boxes (4 KiB each) per warp for the E-strip store of EPI_FWD_E
template <int EPI, int STAGES, bool CTA2, int EW = kEpiWarps, int SB = 2>
constexpr int smem_bytes_total() {
  return 1024 /*align slack*/ + STAGES * stage_bytes<CTA2>() + epi_smem_bytes<EPI>() +
         store_warps<EPI, EW, SB>() * SB * 4096 + (2 * STAGES + 4) * 8 + 16;
}

// Work decomposition, identical in the three warp roles.
//  splits >= 1: units = (tile, split) pairs dealt round-robin to the groups (CTAs or CTA pairs); tiles are
//               numbered row-block-major (nb fastest).  Used by K3 / K4a (splits = 1) and as a fallback.
//  splits == 0: stream-K for the gradient contractions.  The tiles' k-blocks form one sequence of
//               tiles * total_kb items, cut into one contiguous, equally long range per group: every group
//               does the same number of MMAs (no wave quantisation) and the fp32 red.add epilogue runs once
//               per tile plus once per range boundary, instead of `splits` times per tile.  Tiles are
//               numbered column-block-major (mb fastest) so that the groups working on the same rows of W
//               at the same time are the ones with different nb: the strip is read from HBM once.
struct Sched {
  int splits, total_kb, m_tiles, n_tiles, num_groups, u, total_units, rot, rot_m;
  int64_t pos, end;
  // rot_ >= 0 (splits == 1 only): column-block-major order starting at column tile rot_ (sharded K3)
  // rot_m_ > 0 (row-block-major order only): the row blocks are visited starting at row block rot_m_ (fused
  // reduce-scatter: every rank starts with a different owner, so no GPU is the destination of all ranks at once)
  __device__ __forceinline__ void init(int splits_, int total_kb_, int m_tiles_, int n_tiles_, int group_id,
                                       int num_groups_, int rot_ = -1, int rot_m_ = 0) {
    splits = splits_; total_kb = total_kb_; m_tiles = m_tiles_; n_tiles = n_tiles_; num_groups = num_groups_;
    rot = rot_; rot_m = rot_m_;
    u = group_id;
    total_units = m_tiles * n_tiles * (splits > 0 ? splits : 1);
    const int64_t items = (int64_t)m_tiles * n_tiles * total_kb;
    pos = (items * group_id) / num_groups;
    end = (items * (group_id + 1)) / num_groups;
  }
  __device__ __forceinline__ bool next(int& mb, int& nb, int& kb0, int& kb1) {
    if (splits > 0) {
      if (u >= total_units) return false;
      const int tile = u / splits, sp = u - tile * splits;
      if (rot >= 0) {
        const int nbi = tile / m_tiles;
        mb = tile - nbi * m_tiles;
        nb = nbi + rot;
        if (nb >= n_tiles) nb -= n_tiles;
      } else {
        mb = tile / n_tiles; nb = tile - mb * n_tiles;
        mb += rot_m;
        if (mb >= m_tiles) mb -= m_tiles;
      }
      kb0 = (int)(((int64_t)sp * total_kb) / splits);
      kb1 = (int)(((int64_t)(sp + 1) * total_kb) / splits);
      u += num_groups;
      return true;
    }
    if (pos >= end) return false;
    const int tile = (int)(pos / total_kb);
    nb = tile / m_tiles; mb = tile - nb * m_tiles;
    kb0 = (int)(pos - (int64_t)tile * total_kb);
    const int64_t left = end - pos;
    kb1 = (left < (int64_t)(total_kb - kb0)) ? kb0 + (int)left : total_kb;
    pos += kb1 - kb0;
    return true;
  }
};

// lane l ends with the sum over the warp's 32 lanes of x[l] (x is destroyed)
__device__ __forceinline__ float warp_transpose_sum(float (&x)[32], int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = upper ? x[k] : x[k + s];
      const float keep = upper ? x[k + s] : x[k];
      x[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return x[0];
}

// CTA2 = true: the kernel is launched as clusters of two CTAs.  A pair owns a 256 x 256 tile:
// CTA r loads its own 128 rows of A and columns [128r, 128r+128) of B; the leader (r = 0) issues
// tcgen05.mma.cta_group::2 (M = 256), which reads A from each CTA and B from both halves, and its
// commits are multicast to the barriers of both CTAs.  Each CTA's TMEM holds the accumulator of
// its own 128 rows x 256 columns, so the epilogue is the same code in both modes.
// EW = epilogue warps, SB = 4 KiB staging boxes per warp for the E-strip store (EPI_FWD_E).
// (the contraction kernels are capped at 128 registers - bound 512 threads - so that K4t CTAs fit beside them)
template <int EPI, bool A_MN, bool B_MN, int STAGES, bool CTA2, int EW = kEpiWarps, int SB = 2>
__global__ void __launch_bounds__((EPI == EPI_GEMM || EPI == EPI_GEMM_TMA) ? 512 : (EW + 2) * 32, 1)
tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int kEW = EW;                             // epilogue warps of THIS instantiation
  constexpr int kProducerWarp = kEW, kMmaWarp = kEW + 1;
  constexpr int kStage = stage_bytes<CTA2>();
  constexpr int kBRows = CTA2 ? BN / 2 : BN;          // B rows (tile columns) loaded by this CTA
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + STAGES * kStage;
  uint8_t* epi_gen = smem_gen + STAGES * kStage;
  constexpr int kEpiBytes = epi_smem_bytes<EPI>() + store_warps<EPI, EW, SB>() * SB * 4096;
  constexpr bool kSplitEpi = split_epi<EPI, EW, SB>();
  const uint32_t bar_base = epi_base + kEpiBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(epi_gen + kEpiBytes + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = CTA2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int group_id = CTA2 ? (blockIdx.x >> 1) : blockIdx.x;       // scheduling unit: CTA or CTA pair
  const int num_groups = CTA2 ? (gridDim.x >> 1) : gridDim.x;

  if (warp == kProducerWarp && lane == 0) {
    for (int s = 0; s < p.num_segs; ++s) {
      tma_prefetch_desc(&p.a_map[s]);
      tma_prefetch_desc(&p.b_map[s]);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), kEW * (CTA2 ? 2 : 1));
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) {
    if (CTA2) { tmem_alloc_cta2(tmem_slot, kTmemCols); tmem_relinquish_cta2(); }
    else      { tmem_alloc(tmem_slot, kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int total_kb = p.num_segs * p.kb_per_seg;
  Sched sched;
  sched.init(p.splits, total_kb, p.m_tiles, p.n_tiles, group_id, num_groups,
             ((EPI == EPI_FWD || EPI == EPI_FWD_E) && p.landed) ? p.rot_tiles : -1,
             (EPI == EPI_GEMM_TMA || EPI == EPI_GEMM) ? p.rot_m_tiles : 0);
  int mb, nb, kb0, kb1;

  if (warp == kProducerWarp) {
    // ===================================================================== TMA producer (whole warp loops, one lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      int have_src = -1;
      while (sched.next(mb, nb, kb0, kb1)) {
        const int m0 = mb * (CTA2 ? 2 * BM : BM) + (int)cta_rank * BM;
        const int n0 = nb * BN + (int)cta_rank * kBRows * (CTA2 ? 1 : 0);
        if ((EPI == EPI_FWD || EPI == EPI_FWD_E) && p.landed && lane == 0) {
          // first tile of a source's columns: its rows must have landed (both CTAs of a pair wait on their own)
          const int src_lo = (int)(((int64_t)nb * BN) / p.cols_per_source);
          int64_t last_col = (int64_t)nb * BN + BN - 1;
          if (last_col >= p.n_cols) last_col = p.n_cols - 1;
          const int src_hi = (int)(last_col / p.cols_per_source);
          const int own = (int)(((int64_t)p.rot_tiles * BN) / p.cols_per_source);   // written by this GPU's prologue
          for (int src = src_lo; src <= src_hi; ++src) {
            if (src == have_src || src == own) continue;
            const uint32_t epoch = (uint32_t)*p.step * (uint32_t)p.landed_per_step;
            uint64_t t0 = 0;
            for (uint32_t spins = 0;; ++spins) {
              uint32_t v;
              asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.landed + src) : "memory");
              if ((int32_t)(v - epoch) >= 0) break;
              uint64_t now;
              asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
              if (spins == 0) t0 = now;
              if (now - t0 > 2000000000ull) {                    // 2 s: a dead peer must not hang the GPU
                atomicExch(p.error, 1);
                break;
              }
              __nanosleep(128);
            }
            have_src = src;
          }
          asm volatile("fence.proxy.async;" ::: "memory");       // the landed rows are read by TMA (async proxy)
        }
        __syncwarp();
        // segment / offset of the k-block are carried along (no division per k-block) and everything that does not
        // depend on the slot is computed BEFORE the wait: the time from "slot free" to "TMA issued" is on the critical
        // path of a four-stage ring (profiles/r2_ncu_k3_stalls.txt)
        int seg = kb0 / p.kb_per_seg, kin = kb0 - seg * p.kb_per_seg;
        for (int kb = kb0; kb < kb1; ++kb) {
          const int kk = kin * BK;
          const CUtensorMap* const a_map = &p.a_map[seg];
          const CUtensorMap* const b_map = &p.b_map[seg];
          const uint32_t a_dst = smem_base + stage * kStage, b_dst = a_dst + kABytes;
          if (++kin == p.kb_per_seg) { kin = 0; ++seg; }
          mbar_wait(empty_bar(stage), phase ^ 1u);
          if (elect_one()) {
          // the (leader's) full barrier expects the bytes of every CTA that feeds this stage
          if (leader) mbar_expect_tx(full_bar(stage), kStage * (CTA2 ? 2 : 1));
          auto load = [&](uint32_t dst, const CUtensorMap* m, int c0, int c1, uint64_t pol) {
            if (CTA2) tma_load_2d_cta2(dst, m, full_bar(stage), c0, c1, pol);
            else tma_load_2d(dst, m, full_bar(stage), c0, c1, pol);
          };
          if (!A_MN) {
            load(a_dst, a_map, kk, m0, p.policy_a);
          } else {
#pragma unroll
            for (int b = 0; b < BM / 64; ++b) load(a_dst + b * 8192, a_map, m0 + 64 * b, kk, p.policy_a);
          }
          if (!B_MN) {
            load(b_dst, b_map, kk, n0, p.policy_b);
          } else {
#pragma unroll
            for (int b = 0; b < kBRows / 64; ++b) load(b_dst + b * 8192, b_map, n0 + 64 * b, kk, p.policy_b);
          }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else if (warp == kMmaWarp) {
    // ===================================================================== MMA issuer (leader CTA; whole warp loops, one lane issues)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int lu = 0;
      const uint64_t a_desc0 = umma_smem_desc(p.desc_a, smem_base);
      const uint64_t b_desc0 = umma_smem_desc(p.desc_b, smem_base + kABytes);
      const uint32_t idesc = p.idesc;
      for (; sched.next(mb, nb, kb0, kb1); ++lu) {
        const int as = lu & 1;
        const uint32_t aphase = (lu >> 1) & 1;
        mbar_wait(tempty_bar(as), aphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          // descriptors differ from the stage-0 ones only in the 14-bit (address >> 4) field
          const uint64_t ad0 = a_desc0 + (uint64_t)((stage * kStage) >> 4);
          const uint64_t bd0 = b_desc0 + (uint64_t)((stage * kStage) >> 4);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              const uint64_t ad = ad0 + (uint64_t)(A_MN ? k * (2048 >> 4) : k * (32 >> 4));
              const uint64_t bd = bd0 + (uint64_t)(B_MN ? k * (2048 >> 4) : k * (32 >> 4));
              const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
              if (CTA2) umma_bf16_cta2(d_tmem, ad, bd, idesc, acc);
              else umma_bf16(d_tmem, ad, bd, idesc, acc);
            }
            // smem slot free (in both CTAs) once these MMAs retire
            if (CTA2) umma_commit_cta2(empty_bar(stage)); else umma_commit(empty_bar(stage));
            // last k-block of the unit: accumulator ready for the epilogue warps (of both CTAs)
            if (kb == kb1 - 1) { if (CTA2) umma_commit_cta2(tfull_bar(as)); else umma_commit(tfull_bar(as)); }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
    __syncwarp();
  } else {
    // ===================================================================== epilogue warps
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    // split roles: warps 0-7 statistics, 8-15 strip store; each group covers the tile like the eight-warp layout
    const int role = kSplitEpi ? (warp >> 3) : 0;
    const int gw = kSplitEpi ? (warp & 7) : warp; // warp index within its role group
    constexpr int kGroupWarps = kSplitEpi ? 8 : kEW;
    const int hh = gw >> 2;                       // which part (half / quarter) of the tile's columns this warp owns
    constexpr int kChunks = BN / 32 / (kGroupWarps / 4);  // 32-column chunks per warp: 4 (eight warps) or 2 (sixteen)
    const int row = q * 32 + lane;                // tile row owned by this thread
    const int et = warp * 32 + lane;              // 0 .. 32*kEW-1
    const int c_lo = hh * kChunks;                // first 32-column chunk of this warp
    const float c1 = p.inv_tau * 1.4426950408889634f;    // exp(s/tau - shift) = 2^(s*c1 - c1)
    const bool excl = (p.flags & EVK_FLAG_EXCLUDE_DIAG) != 0;
    auto release_tmem = [&](int as) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CTA2 && !leader) mbar_arrive_cluster(tempty_bar(as), 0);
        else mbar_arrive(tempty_bar(as));
      }
    };
    int lu = 0;
    for (; sched.next(mb, nb, kb0, kb1); ++lu) {
      const int m0 = mb * (CTA2 ? 2 * BM : BM) + (int)cta_rank * BM;
      const int n0 = nb * BN;
      const int as = lu & 1;
      const uint32_t aphase = (lu >> 1) & 1;
      const int64_t i = (int64_t)m0 + row;
      const bool row_ok = i < p.n_rows;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      if (EPI == EPI_FWD || EPI == EPI_FWD_E) {
        constexpr bool kStoreE = EPI == EPI_FWD_E;
        // [2][4][BN]: double-buffered by tile, so one CTA-level barrier per tile is enough (a warp cannot write
        // buffer b again before every warp has passed the next tile's barrier, i.e. finished reading b)
        float* colpart = reinterpret_cast<float*>(epi_gen) + (lu & 1) * 4 * BN;
        // EPI_FWD_E: this warp's private staging for its 32 rows x (32 kChunks) columns of E (bf16),
        // 128B-swizzled [32 x 64] boxes, stored with its own TMA stores (as in EPI_BWD_W)
        // SB staging buffers per warp: with fewer buffers than boxes a buffer is reused after its store has been read
        const uint32_t wstg = epi_base + 2 * 4 * BN * 4 + gw * (SB * 4096);
        const int m_warp = m0 + q * 32;
        const bool do_stats = !kSplitEpi || role == 0;            // this warp produces the statistics
        const bool do_store = kStoreE && (!kSplitEpi || role == 1);   // ... the E strip
        const bool want_col = do_stats && (p.flags & EVK_FLAG_NO_COLSUM) == 0;
        const bool want_pos = do_stats && (p.flags & EVK_FLAG_NO_POS) == 0;   // else evk_mpce_pos supplies the positive sums
        const uint32_t* mrow = want_pos ? p.bits + (row_ok ? i : 0) * p.ld_words + (n0 >> 5) : nullptr;
        const int64_t dcol = i + p.diag_offset - n0;              // diagonal column inside this tile?
        float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f, rp = 0.f;
        // the mask words of this warp's columns: one vector load issued BEFORE the wait for the accumulator
        // (its latency - the mask streams from HBM - used to be exposed once per 32-column chunk)
        uint32_t mw4[4] = {0u, 0u, 0u, 0u};
        if (want_pos && row_ok) {
          if ((p.flags & kIntVecMask) && kChunks == 4) {
            const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(mrow + c_lo));
            mw4[0] = t4.x; mw4[1] = t4.y; mw4[2] = t4.z; mw4[3] = t4.w;
          } else if ((p.flags & kIntVecMask) && kChunks == 2) {
            const uint2 t2 = __ldg(reinterpret_cast<const uint2*>(mrow + c_lo));
            mw4[0] = t2.x; mw4[1] = t2.y;
          } else {
#pragma unroll
            for (int k = 0; k < kChunks; ++k) mw4[k] = __ldg(mrow + c_lo + k);
          }
        }
        if (do_store) {
          if (lane == 0) tma_store_wait_read<0>();                // this warp's previous stores have left smem
          __syncwarp();
        }
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        if (kSplitEpi && role == 1) {
          // ---- store warps: accumulator -> exp -> bf16 -> swizzled staging -> TMA store; nothing else
#pragma unroll 1
          for (int cc = 0; cc < kChunks; ++cc) {
            const int c = c_lo + cc;
            float v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait(v);
            if (cc == kChunks - 1) release_tmem(as);              // last read of this warp: the stage may be refilled
            const int cbase = n0 + c * 32;
            uint32_t live = (cbase + 32 <= p.n_cols) ? 0xffffffffu
                            : (cbase >= p.n_cols ? 0u : ((1u << (int)(p.n_cols - cbase)) - 1u));
            if (!row_ok) live = 0u;
            if (excl && dcol >= 0 && (dcol >> 5) == c) live &= ~(1u << (int)(dcol & 31));
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = ex2_approx(fmaf(v[k], c1, -c1));
            if (live != 0xffffffffu) {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = ((live >> k) & 1u) ? v[k] : 0.f;
            }
            const int box = cc >> 1;
            const uint32_t row_st = wstg + (box % SB) * 4096 + lane * 128;
            const int u0 = (cc & 1) * 4;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const uint32_t h0 = pack_bf16x2(v[8 * uu + 0], v[8 * uu + 1]);
              const uint32_t h1 = pack_bf16x2(v[8 * uu + 2], v[8 * uu + 3]);
              const uint32_t h2 = pack_bf16x2(v[8 * uu + 4], v[8 * uu + 5]);
              const uint32_t h3 = pack_bf16x2(v[8 * uu + 6], v[8 * uu + 7]);
              const uint32_t off = static_cast<uint32_t>(((u0 + uu) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_st + off), "r"(h0), "r"(h1), "r"(h2),
                           "r"(h3) : "memory");
            }
            if (cc & 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&p.out_map[0], wstg + (box % SB) * 4096, n0 + (c - 1) * 32, m_warp, p.policy_out);
                tma_store_commit();
              }
            }
          }
          continue;
        }
        // one 32-column chunk of this warp's rows: v = the accumulator values (destroyed)
        auto chunk = [&](float (&v)[32], const int cc) {
          const int c = c_lo + cc;
          const uint32_t mword = cc == 0 ? mw4[0] : (cc == 1 ? mw4[1] : (cc == 2 ? mw4[2] : mw4[3]));
          const int cbase = n0 + c * 32;
          uint32_t live = (cbase + 32 <= p.n_cols) ? 0xffffffffu
                          : (cbase >= p.n_cols ? 0u : ((1u << (int)(p.n_cols - cbase)) - 1u));
          if (!row_ok) live = 0u;
          if (excl && dcol >= 0 && (dcol >> 5) == c) live &= ~(1u << (int)(dcol & 31));
          if (mword & live) {                                     // rare: a positive in this 32-col chunk
            const uint32_t m = mword & live;
#pragma unroll
            for (int k = 0; k < 32; ++k)
              if ((m >> k) & 1u) rp = fmaf(v[k], p.inv_tau, rp);
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) v[k] = ex2_approx(fmaf(v[k], c1, -c1));
          if (live != 0xffffffffu) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = ((live >> k) & 1u) ? v[k] : 0.f;
          }
#pragma unroll
          for (int k = 0; k < 32; k += 4) {
            rs0 += v[k]; rs1 += v[k + 1]; rs2 += v[k + 2]; rs3 += v[k + 3];
          }
          if (kStoreE && !kSplitEpi) {                            // E -> bf16 -> swizzled staging (dead entries are 0)
            const int box = cc >> 1;
            if ((cc & 1) == 0 && box >= SB) {                     // buffer reuse within the tile
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            }
            const uint32_t row_st = wstg + (box % SB) * 4096 + lane * 128;
            const int u0 = (cc & 1) * 4;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const uint32_t h0 = pack_bf16x2(v[8 * uu + 0], v[8 * uu + 1]);
              const uint32_t h1 = pack_bf16x2(v[8 * uu + 2], v[8 * uu + 3]);
              const uint32_t h2 = pack_bf16x2(v[8 * uu + 4], v[8 * uu + 5]);
              const uint32_t h3 = pack_bf16x2(v[8 * uu + 6], v[8 * uu + 7]);
              const uint32_t off = static_cast<uint32_t>(((u0 + uu) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_st + off), "r"(h0), "r"(h1), "r"(h2),
                           "r"(h3) : "memory");
            }
            if (cc & 1) {                                         // a [32 x 64] box is complete: store it now
              fence_proxy_async_smem();                           // generic-proxy writes -> async proxy
              __syncwarp();
              if (lane == 0) {                                    // OOB rows / columns are clipped by the map
                tma_store_2d(&p.out_map[0], wstg + (box % SB) * 4096, n0 + (c - 1) * 32, m_warp, p.policy_out);
                tma_store_commit();
              }
            }
          }
          if (want_col) {
            const float cs = warp_transpose_sum(v, lane);
            colpart[q * BN + c * 32 + lane] = cs;
          }
        };
        // The chunks are software-pipelined over two register buffers: the tcgen05.ld of chunk cc+1 is in flight
        // while chunk cc is processed, and the TMEM stage goes back to the MMA warp as soon as the LAST load has
        // completed - one chunk's worth of exp / pack / transpose earlier than after the loop (kChunks is 2 or 4).
        constexpr bool kPipe = kEW <= 8;                          // the sixteen-warp variants are capped at 96 registers
        if (!kPipe || (p.flags & kIntNoEpiPipe)) {                // (EVK_K3_PIPE=0: A/B switch) load, wait, process
          float v[32];
#pragma unroll 1
          for (int cc = 0; cc < kChunks; ++cc) {
            tmem_ld_32x32(taddr + (c_lo + cc) * 32, v);
            tmem_ld_wait(v);
            if (cc == kChunks - 1) release_tmem(as);
            chunk(v, cc);
          }
        } else {
          float va[32], vb[32];
          tmem_ld_32x32(taddr + c_lo * 32, va);
#pragma unroll 1
          for (int cc = 0; cc < kChunks; cc += 2) {
            tmem_ld_wait(va);
            tmem_ld_32x32(taddr + (c_lo + cc + 1) * 32, vb);
            chunk(va, cc);
            tmem_ld_wait(vb);
            if (cc + 2 < kChunks) tmem_ld_32x32(taddr + (c_lo + cc + 2) * 32, va);
            else release_tmem(as);                                // TMEM stage drained
            chunk(vb, cc + 1);
          }
        }
        if (row_ok) {
          const int64_t po = ((int64_t)nb * (kGroupWarps / 4) + hh) * p.ld_rowpart + i;
          p.row_sum_part[po] = (rs0 + rs1) + (rs2 + rs3);
          if (want_pos) p.row_pos_part[po] = rp;
        }
        if (want_col) {
          named_bar_sync(1, kGroupWarps * 32);                    // the statistics warps only
          if (et < BN) {
            const float s = (colpart[et] + colpart[BN + et]) + (colpart[2 * BN + et] + colpart[3 * BN + et]);
            // each CTA (128-row block) writes its own partial row: index = global 128-row block
            if (n0 + et < p.n_cols && m0 < p.n_rows) p.col_sum_part[(int64_t)(m0 / BM) * p.ld_colpart + n0 + et] = s;
          }
        }
      } else if (EPI == EPI_BWD_W) {
        // Every warp stages and stores its own 32 rows x 128 columns: private 8 KiB staging region,
        // own 32-row TMA stores, own bulk-group accounting -> no CTA-level barrier in this epilogue.
        const bool split = (p.flags & EVK_FLAG_SPLIT_BF16) != 0;
        const uint32_t wstg = epi_base + warp * 8192;             // [2 boxes][32 rows][128 B], 128B-swizzled
        const float a_i = row_ok ? __ldg(p.a_row + i) : 0.f;
        const float n2c = row_ok ? -2.f / (float)max(__ldg(p.counts + i), 1) : 0.f;
        const uint32_t* mrow = p.bits + (row_ok ? i : 0) * p.ld_words + (n0 >> 5);
        const int64_t dcol = i + p.diag_offset - n0;
        const int m_warp = m0 + q * 32;                           // first row of this warp's boxes
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
        // non-split: one round of 4 chunks: box 0 = chunks 0,1, box 1 = chunks 2,3 (hi)
        // split    : two rounds of 2 chunks: box 0 = hi, box 1 = lo
        const int rounds = split ? 2 : 1, per_round = split ? 2 : 4;
#pragma unroll 1
        for (int r = 0; r < rounds; ++r) {
          if (lane == 0) tma_store_wait_read<0>();                // this warp's previous stores have left smem
          __syncwarp();
#pragma unroll 1
          for (int cc = 0; cc < per_round; ++cc) {
            const int c = c_lo + r * per_round + cc;
            const uint32_t mword = row_ok ? __ldg(mrow + c) : 0u;
            float v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait(v);
            {
              const int cbase = n0 + c * 32;
              if (cbase + 32 <= p.n_cols) {
                const float4* bp = reinterpret_cast<const float4*>(p.b_col + cbase);   // warp-uniform: broadcast loads
#pragma unroll
                for (int k4 = 0; k4 < 8; ++k4) {
                  const float4 b4 = __ldg(bp + k4);
                  v[4 * k4 + 0] = ex2_approx(fmaf(v[4 * k4 + 0], c1, -c1)) * (a_i + b4.x);
                  v[4 * k4 + 1] = ex2_approx(fmaf(v[4 * k4 + 1], c1, -c1)) * (a_i + b4.y);
                  v[4 * k4 + 2] = ex2_approx(fmaf(v[4 * k4 + 2], c1, -c1)) * (a_i + b4.z);
                  v[4 * k4 + 3] = ex2_approx(fmaf(v[4 * k4 + 3], c1, -c1)) * (a_i + b4.w);
                }
              } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                  const float bj = (cbase + k < p.n_cols) ? __ldg(p.b_col + cbase + k) : 0.f;
                  v[k] = ex2_approx(fmaf(v[k], c1, -c1)) * (a_i + bj);
                }
              }
            }
            if (mword) {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if ((mword >> k) & 1u) v[k] += n2c;
            }
            if (excl && dcol >= 0 && (dcol >> 5) == c) {
              const int dk = (int)(dcol & 31);
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (k == dk) v[k] = 0.f;
            }
            // rows >= n_rows / columns >= n_cols hold garbage here; the TMA store clips them.
            const int box_hi = split ? 0 : (cc >> 1);
            const uint32_t row_hi = wstg + box_hi * 4096 + lane * 128;
            const int u0 = (cc & 1) * 4;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              uint32_t h0 = pack_bf16x2(v[8 * uu + 0], v[8 * uu + 1]);
              uint32_t h1 = pack_bf16x2(v[8 * uu + 2], v[8 * uu + 3]);
              uint32_t h2 = pack_bf16x2(v[8 * uu + 4], v[8 * uu + 5]);
              uint32_t h3 = pack_bf16x2(v[8 * uu + 6], v[8 * uu + 7]);
              const uint32_t off = static_cast<uint32_t>(((u0 + uu) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + off), "r"(h0), "r"(h1), "r"(h2),
                           "r"(h3) : "memory");
              if (split) {
                float l[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const uint32_t hw = e == 0 ? h0 : (e == 1 ? h1 : (e == 2 ? h2 : h3));
                  l[2 * e] = v[8 * uu + 2 * e] - __uint_as_float(hw << 16);
                  l[2 * e + 1] = v[8 * uu + 2 * e + 1] - __uint_as_float(hw & 0xffff0000u);
                }
                const uint32_t l0 = pack_bf16x2(l[0], l[1]), l1 = pack_bf16x2(l[2], l[3]);
                const uint32_t l2 = pack_bf16x2(l[4], l[5]), l3 = pack_bf16x2(l[6], l[7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_hi + 4096 + off), "r"(l0),
                             "r"(l1), "r"(l2), "r"(l3) : "memory");
              }
            }
          }
          if (r == rounds - 1) release_tmem(as);
          fence_proxy_async_smem();                               // generic-proxy writes -> async proxy
          __syncwarp();
          if (lane == 0) {
            const int cx = n0 + (c_lo + r * per_round) * 32;      // first column of this round
            if (split) {
              tma_store_2d(&p.out_map[0], wstg, cx, m_warp, p.policy_out);
              tma_store_2d(&p.out_map[1], wstg + 4096, cx, m_warp, p.policy_out);
            } else {
              tma_store_2d(&p.out_map[0], wstg, cx, m_warp, p.policy_out);
              tma_store_2d(&p.out_map[0], wstg + 4096, cx + 64, m_warp, p.policy_out);
            }
            tma_store_commit();
          }
        }
      } else if (EPI == EPI_GEMM_TMA) {
        // Each warp moves its 32 rows x 128 columns in four [32 x 32] fp32 boxes through a private,
        // double-buffered, 128B-swizzled staging area; one elected lane issues the TMA store / reduce-add.
        const uint32_t wstg = epi_base + warp * 8192;
        const int m_warp = m0 + q * 32;
        int owner = 0, row_in_owner = m_warp;
        if (p.rows_per_owner > 0) {
          owner = (int)(m_warp / p.rows_per_owner);
          row_in_owner = (int)(m_warp - (int64_t)owner * p.rows_per_owner);
        }
        const CUtensorMap* omap = &p.peer_map[m_warp < p.n_rows ? owner : 0];
        const bool reduce = (p.flags & kIntStore) == 0;
        const bool out16 = (p.flags & kIntBf16Out) != 0;              // bf16 output boxes [32 rows x 64 columns] (store mode only)
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BN / 64; ++cc) {
          const int c = c_lo + cc;
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait(v);
          if (cc == BN / 64 - 1) release_tmem(as);
          if (out16) {
            // two 32-column chunks share one 128-byte-row box; buffers alternate per box
            const int box = cc >> 1;
            if ((cc & 1) == 0) {
              if (lane == 0) tma_store_wait_read<1>();
              __syncwarp();
            }
            const uint32_t row_st = wstg + (box & 1) * 4096 + lane * 128;
            const int u0 = (cc & 1) * 4;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              const uint32_t h0 = pack_bf16x2(p.alpha * v[8 * uu + 0], p.alpha * v[8 * uu + 1]);
              const uint32_t h1 = pack_bf16x2(p.alpha * v[8 * uu + 2], p.alpha * v[8 * uu + 3]);
              const uint32_t h2 = pack_bf16x2(p.alpha * v[8 * uu + 4], p.alpha * v[8 * uu + 5]);
              const uint32_t h3 = pack_bf16x2(p.alpha * v[8 * uu + 6], p.alpha * v[8 * uu + 7]);
              const uint32_t off = static_cast<uint32_t>(((u0 + uu) ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_st + off), "r"(h0), "r"(h1), "r"(h2),
                           "r"(h3) : "memory");
            }
            if (cc & 1) {
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0 && m_warp < p.n_rows && !(p.flags & kIntDropOut))
                tma_store_2d(omap, wstg + (box & 1) * 4096, n0 + (c - 1) * 32, row_in_owner, p.policy_out);
              if (lane == 0) tma_store_commit();
            }
            continue;
          }
          if (lane == 0) tma_store_wait_read<1>();              // the buffer used two boxes ago has left smem
          __syncwarp();
          const uint32_t row_st = wstg + (cc & 1) * 4096 + lane * 128;
#pragma unroll
          for (int uu = 0; uu < 8; ++uu) {
            const uint32_t off = static_cast<uint32_t>((uu ^ (lane & 7)) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(row_st + off), "f"(p.alpha * v[4 * uu]),
                         "f"(p.alpha * v[4 * uu + 1]), "f"(p.alpha * v[4 * uu + 2]), "f"(p.alpha * v[4 * uu + 3]) : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && m_warp < p.n_rows && !(p.flags & kIntDropOut)) {
            // rows / columns beyond the output are clipped by the tensor map
            if (reduce) tma_reduce_add_2d(omap, wstg + (cc & 1) * 4096, n0 + c * 32, row_in_owner);
            else tma_store_2d(omap, wstg + (cc & 1) * 4096, n0 + c * 32, row_in_owner, p.policy_out);
          }
          if (lane == 0) tma_store_commit();
        }
      } else {  // EPI_GEMM
        float* orow;
        if (p.rows_per_owner > 0) {
          const int64_t owner = i / p.rows_per_owner;
          orow = p.out_peer[row_ok ? owner : 0] + (i - owner * p.rows_per_owner) * p.ld_out + n0;
        } else {
          orow = p.out + i * p.ld_out + n0;
        }
        mbar_wait(tfull_bar(as), aphase);
        tc_fence_after();
#pragma unroll 1
        for (int cc = 0; cc < BN / 64; ++cc) {
          const int c = c_lo + cc;
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          tmem_ld_wait(v);
          const int cbase = n0 + c * 32;
          if (row_ok && !(p.flags & kIntDropOut)) {
            if (p.flags & kIntStore) {                            // whole-K units: plain (posted) stores, no accumulation
              if (cbase + 32 <= p.n_cols) {
#pragma unroll
                for (int k = 0; k < 32; k += 4)
                  *reinterpret_cast<float4*>(orow + c * 32 + k) =
                      make_float4(p.alpha * v[k], p.alpha * v[k + 1], p.alpha * v[k + 2], p.alpha * v[k + 3]);
              } else {
#pragma unroll
                for (int k = 0; k < 32; ++k)
                  if (cbase + k < p.n_cols) orow[c * 32 + k] = p.alpha * v[k];
              }
            } else if (cbase + 32 <= p.n_cols) {
#pragma unroll
              for (int k = 0; k < 32; k += 4)
                red_add_v4(orow + c * 32 + k, p.alpha * v[k], p.alpha * v[k + 1], p.alpha * v[k + 2], p.alpha * v[k + 3]);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k)
                if (cbase + k < p.n_cols) atomicAdd(orow + c * 32 + k, p.alpha * v[k]);
            }
          }
        }
        release_tmem(as);
      }
    }
    if ((EPI == EPI_BWD_W || EPI == EPI_FWD_E || EPI == EPI_GEMM_TMA) && lane == 0) tma_store_wait_all<0>();   // smem must outlive each warp's bulk stores
  }

  tc_fence_before();
  if (CTA2) cluster_sync_all(); else __syncthreads();            // the peer may still arrive on our barriers
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (CTA2) tmem_dealloc_cta2(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// 2D bf16 row-major matrix [rows, cols] with pitch ld (elements); box = [box_rows, box_cols], 128B swizzle.
int make_map_bf16(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                  int box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return evk_set_error(EVK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  EVK_REQUIRE(evk_aligned16(base) && ld % 8 == 0, "bf16 operand needs a 16-byte aligned base and ld %% 8 == 0 (ld=%lld)",
              (long long)ld);
  EVK_REQUIRE(rows > 0 && cols > 0 && ld >= cols, "bad operand shape rows=%lld cols=%lld ld=%lld", (long long)rows,
              (long long)cols, (long long)ld);
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return evk_set_error(EVK_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return EVK_OK;
}

// fp32 row-major output [rows, cols] with pitch ld (elements); box = [32 rows, 32 columns] (128-byte rows), 128B swizzle
int make_map_f32_out(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return evk_set_error(EVK_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  EVK_REQUIRE(evk_aligned16(base) && ld % 4 == 0 && rows > 0 && cols > 0 && ld >= cols, "fp32 output needs a 16-byte aligned base and ld %% 4 == 0");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return evk_set_error(EVK_ERR_CUDA, "cuTensorMapEncodeTiled (fp32 output) failed with CUresult %d", (int)r);
  return EVK_OK;
}

bool use_tma_epilogue() {
  // EVK_GEMM_EPI=red selects the per-lane red.add / st epilogue of the gradient contractions
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVK_GEMM_EPI");
    v = (e && e[0] == 'r') ? 0 : 1;
  }
  return v != 0;
}

// K-major operand [mn_extent rows, k_extent cols]: one box of box_mn rows x 64 k.
// MN-major operand stored [k_extent rows, mn_extent cols]: boxes of 64 k rows x 64 mn.
int make_operand_map(CUtensorMap* map, const void* base, bool mn_major, int64_t mn_extent, int64_t k_extent,
                     int64_t ld, int box_mn) {
  if (!mn_major) return make_map_bf16(map, base, mn_extent, k_extent, ld, box_mn, BK);
  return make_map_bf16(map, base, k_extent, mn_extent, ld, BK, 64);
}

bool use_cta_pairs() {
  // CTA-pair (cta_group::2) kernels are the default; EVK_CTA_PAIR=0 selects the single-CTA kernels.
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVK_CTA_PAIR");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

int k3_variant() {
  // EVK_K3_VARIANT: 0 = 8 epilogue warps, 4 stages; 1 = 16 epilogue warps (32 x 64 columns each); 2 = 8 warps, 5 stages
  // with single-buffered E staging; 3 = 16 warps in two roles (8 statistics + 8 strip store), 4 stages.
  // evk_mpce_row_parts() tells the caller how many row-partial rows K3 writes.
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVK_K3_VARIANT");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : kK3DefaultVariant;
  }
  return v;
}

bool use_stream_k() {
  // EVK_STREAMK=1 selects the stream-K work decomposition of the gradient contractions.  Measured on B200 at
  // the bench shape it is 5% SLOWER than tile x split-K units (322 vs 305 us per contraction: the split-K
  // units of one row block run concurrently on different SMs and share the W rows through L2), so it is off.
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("EVK_STREAMK");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

void fill_descs(TcParams& p, bool a_mn, bool b_mn, int variant, bool cta2) {
  // K-major SW128: 8-row groups 1024 B apart (SBO); LBO unused.  MN-major SW128: 64-element MN
  // atoms 8192 B apart (LBO, one TMA box each), 8-row K groups 1024 B apart (SBO).
  const uint64_t kmaj = umma_smem_desc_base(16, 1024);
  const uint64_t mnmaj = (variant & 1) ? umma_smem_desc_base(1024, 8192) : umma_smem_desc_base(8192, 1024);
  p.desc_a = a_mn ? mnmaj : kmaj;
  p.desc_b = b_mn ? mnmaj : kmaj;
  p.idesc = umma_idesc_bf16(cta2 ? 2 * BM : BM, BN, a_mn, b_mn);
}

template <int EPI, bool A_MN, bool B_MN, int STAGES, bool CTA2, int EW = kEpiWarps, int SB = 2>
int launch(const TcParams& p, cudaStream_t s) {
  constexpr int smem = smem_bytes_total<EPI, STAGES, CTA2, EW, SB>();
  static_assert(smem <= 232448, "shared memory budget exceeded");
  auto kern = tc_kernel<EPI, A_MN, B_MN, STAGES, CTA2, EW, SB>;
  // per instantiation, per thread AND per device (the attribute belongs to the function on one device; a host
  // thread that drives several GPUs must opt in on each of them)
  static thread_local uint64_t attr_set = 0;
  int dev = 0;
  EVK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !((attr_set >> dev) & 1ull)) {
    EVK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) attr_set |= 1ull << dev;
  }
  int sms = evk_sm_count();
  if (p.cta_limit > 0 && p.cta_limit < sms) sms = p.cta_limit & ~1;
  if (sms < 2) sms = 2;
  // stream-K (splits == 0): every group gets a range; a range should hold at least a few k-blocks
  const int64_t items = (int64_t)p.m_tiles * p.n_tiles * p.num_segs * p.kb_per_seg;
  const int units = p.splits > 0 ? p.m_tiles * p.n_tiles * p.splits : (int)(items / 4 < sms ? (items / 4 > 0 ? items / 4 : 1) : sms);
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  if (CTA2) {
    const int groups = units < sms / 2 ? units : sms / 2;
    cfg.gridDim = dim3(2 * groups);
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3(units < sms ? units : sms);
  }
  cfg.blockDim = dim3((EW + 2) * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  EVK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  EVK_CHECK_LAUNCH("tc_kernel");
  return EVK_OK;
}

int check_device() {
  // cuTensorMapEncodeTiled is a driver call: it needs the primary context bound to THIS thread.
  // autograd runs backward on its own thread, whose first CUDA activity may be this library.
  static thread_local uint64_t ctx_bound = 0;         // per device: the thread may switch devices between calls
  int dev = 0;
  EVK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !((ctx_bound >> dev) & 1ull)) {
    // cudaSetDevice initialises the primary context and makes it current on this thread; unlike cudaFree(nullptr)
    // it is legal while another thread's stream is being captured
    EVK_CUDA(cudaSetDevice(dev));
    if (dev >= 0 && dev < 64) ctx_bound |= 1ull << dev;
  }
  if (!evk_is_sm100())
    return evk_set_error(EVK_ERR_UNSUPPORTED, "the tcgen05 path needs an sm_100 (B200) device; there is no fallback");
  return EVK_OK;
}

// smallest split factor whose last wave is at least ~94% full (else the best seen)
int choose_splits(int tiles, int total_kb, int groups) {
  int best = 1;
  double best_eff = 0.0;
  for (int s = 1; s <= 16; ++s) {
    if (total_kb / s < 8 && s > 1) break;
    const int units = tiles * s;
    const double eff = (double)units / (double)(((units + groups - 1) / groups) * groups);
    if (eff > best_eff + 1e-9) { best_eff = eff; best = s; }
    if (eff >= 0.94) { best = s; break; }
  }
  return best;
}

int setup_sim_operands(TcParams& p, const void* q_hi, const void* q_lo, int64_t ld_q, const void* k_hi,
                       const void* k_lo, int64_t ld_k, int64_t n_rows, int64_t n_cols, int64_t d, int flags,
                       bool cta2) {
  const bool split = (flags & EVK_FLAG_SPLIT_BF16) != 0;
  EVK_REQUIRE(q_hi && k_hi && (!split || (q_lo && k_lo)), "null operand pointer");
  EVK_REQUIRE(n_rows > 0 && n_cols > 0 && d > 0, "empty problem");
  EVK_REQUIRE(n_rows < (1ll << 30) && n_cols < (1ll << 30), "problem too large");
  p.num_segs = split ? 3 : 1;
  p.kb_per_seg = (int)((d + BK - 1) / BK);
  const void* qa[3] = {q_hi, q_hi, q_lo};
  const void* ka[3] = {k_hi, k_lo, k_hi};
  for (int s = 0; s < p.num_segs; ++s) {
    int rc = make_operand_map(&p.a_map[s], qa[s], false, n_rows, d, ld_q, BM);
    if (rc != EVK_OK) return rc;
    rc = make_operand_map(&p.b_map[s], ka[s], false, n_cols, d, ld_k, cta2 ? BN / 2 : BN);
    if (rc != EVK_OK) return rc;
  }
  const int tile_m = cta2 ? 2 * BM : BM;
  p.m_tiles = (int)((n_rows + tile_m - 1) / tile_m);
  p.n_tiles = (int)((n_cols + BN - 1) / BN);
  p.splits = 1;
  p.n_rows = n_rows;
  p.n_cols = n_cols;
  p.policy_a = p.policy_b = kEvictLast;       // Qhat / Khat are re-read by every tile of the sweep
  p.policy_out = kEvictFirst;                 // the W strip is written once and streamed back later
  if (const char* e = getenv("EVK_L2_HINTS")) {
    if (e[0] == '0') p.policy_a = p.policy_b = p.policy_out = kEvictNormal;
  }
  fill_descs(p, false, false, 0, cta2);
  return EVK_OK;
}

}  // namespace

namespace {
int mpce_fwd_impl(const void* q_hi, const void* q_lo, int64_t ld_q, const void* k_hi, const void* k_lo,
                  int64_t ld_k, int64_t n_rows, int64_t n_cols, int64_t d, const uint32_t* bits,
                  int64_t ld_words, float inv_tau, int flags, int64_t diag_offset, float* row_sum_part,
                  float* row_pos_part, int64_t ld_rowpart, float* col_sum_part, int64_t ld_colpart,
                  void* e_out, int64_t ld_e, evk_stream_t stream, const uint32_t* landed = nullptr,
                  const int* step = nullptr, int* error = nullptr, int64_t cols_per_source = 0, int64_t first_col = 0,
                  int landed_per_step = 1) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  const bool cta2 = use_cta_pairs();
  TcParams p;
  memset(&p, 0, sizeof(p));
  rc = setup_sim_operands(p, q_hi, q_lo, ld_q, k_hi, k_lo, ld_k, n_rows, n_cols, d, flags, cta2);
  if (rc != EVK_OK) return rc;
  const bool want_col = (flags & EVK_FLAG_NO_COLSUM) == 0;
  const bool want_pos = (flags & EVK_FLAG_NO_POS) == 0;
  EVK_REQUIRE(row_sum_part && (!want_pos || (bits && row_pos_part)) && (!want_col || col_sum_part), "evk_mpce_fwd: null pointer");
  EVK_REQUIRE(!want_pos || ld_words >= (int64_t)p.n_tiles * (BN / 32), "evk_mpce_fwd: ld_words=%lld must cover whole 256-column tiles (>= %lld)",
              (long long)ld_words, (long long)p.n_tiles * (BN / 32));
  EVK_REQUIRE(ld_rowpart >= n_rows && (!want_col || ld_colpart >= n_cols), "evk_mpce_fwd: partial pitches too small");
  EVK_REQUIRE(inv_tau > 0.f && inv_tau <= EVK_MAX_INV_TAU, "evk_mpce_fwd: 1/tau=%g outside (0, 40]: the fixed-shift softmax needs exp(-2/tau) to stay normal in fp32", inv_tau);
  p.inv_tau = inv_tau;
  p.flags = flags;
  if (bits && evk_aligned16(bits) && ld_words % 4 == 0) p.flags |= kIntVecMask;     // 128-bit mask loads
  {
    static const bool no_pipe = [] { const char* e = getenv("EVK_K3_PIPE"); return e && e[0] == '0'; }();
    if (no_pipe) p.flags |= kIntNoEpiPipe;
  }
  p.diag_offset = diag_offset;
  p.bits = bits;
  p.ld_words = ld_words;
  p.row_sum_part = row_sum_part;
  p.row_pos_part = row_pos_part;
  p.ld_rowpart = ld_rowpart;
  p.col_sum_part = col_sum_part;
  p.ld_colpart = ld_colpart;
  if (landed) {
    EVK_REQUIRE(step && error && cols_per_source > 0 && cols_per_source % BN == 0 && first_col >= 0 && first_col % BN == 0 &&
                    first_col < n_cols, "evk_mpce_fwd_store_gathered: cols_per_source / first_col must be multiples of 256");
    p.landed = landed;
    p.step = step;
    p.error = error;
    p.cols_per_source = cols_per_source;
    p.rot_tiles = (int)(first_col / BN);
    p.landed_per_step = landed_per_step > 0 ? landed_per_step : 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (e_out) {
    EVK_REQUIRE((flags & EVK_FLAG_SPLIT_BF16) == 0, "evk_mpce_fwd_store: the E strip is a bf16-mode feature (no split operands)");
    EVK_REQUIRE(ld_e >= n_cols && ld_e % 8 == 0, "evk_mpce_fwd_store: ld_e=%lld must be >= n_cols and a multiple of 8", (long long)ld_e);
    rc = make_map_bf16(&p.out_map[0], e_out, n_rows, n_cols, ld_e, 32, 64);
    if (rc != EVK_OK) return rc;
    if (!cta2) return launch<EPI_FWD_E, false, false, 3, false>(p, s);
    switch (k3_variant()) {
      case 1: return launch<EPI_FWD_E, false, false, 4, true, 16, 1>(p, s);   // 16 epilogue warps, 4 stages
      case 2: return launch<EPI_FWD_E, false, false, 5, true, 8, 1>(p, s);    // 8 warps, one staging box each, 5 stages
      case 3: return launch<EPI_FWD_E, false, false, 4, true, 16, 2>(p, s);   // 8 statistics + 8 store warps
      default: return launch<EPI_FWD_E, false, false, 4, true, 8, 2>(p, s);   // 8 warps, two boxes each, 4 stages
    }
  }
  if (cta2 && k3_variant() == 1) return launch<EPI_FWD, false, false, 6, true, 16, 1>(p, s);
  return cta2 ? launch<EPI_FWD, false, false, 6, true>(p, s) : launch<EPI_FWD, false, false, 4, false>(p, s);
}
}  // namespace

extern "C" int evk_mpce_row_parts(void) { return (use_cta_pairs() && k3_variant() == 1) ? 4 : 2; }

extern "C" int evk_mpce_fwd(const void* q_hi, const void* q_lo, int64_t ld_q, const void* k_hi, const void* k_lo,
                            int64_t ld_k, int64_t n_rows, int64_t n_cols, int64_t d, const uint32_t* bits,
                            int64_t ld_words, float inv_tau, int flags, int64_t diag_offset, float* row_sum_part,
                            float* row_pos_part, int64_t ld_rowpart, float* col_sum_part, int64_t ld_colpart,
                            evk_stream_t stream) {
  return mpce_fwd_impl(q_hi, q_lo, ld_q, k_hi, k_lo, ld_k, n_rows, n_cols, d, bits, ld_words, inv_tau, flags, diag_offset,
                       row_sum_part, row_pos_part, ld_rowpart, col_sum_part, ld_colpart, nullptr, 0, stream);
}

extern "C" int evk_mpce_fwd_store(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t n_rows,
                                  int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words, float inv_tau,
                                  int flags, int64_t diag_offset, float* row_sum_part, float* row_pos_part,
                                  int64_t ld_rowpart, float* col_sum_part, int64_t ld_colpart, void* e_out, int64_t ld_e,
                                  evk_stream_t stream) {
  EVK_REQUIRE(e_out && evk_aligned16(e_out), "evk_mpce_fwd_store: e_out must be a 16-byte aligned device pointer");
  return mpce_fwd_impl(q_hi, nullptr, ld_q, k_hi, nullptr, ld_k, n_rows, n_cols, d, bits, ld_words, inv_tau, flags,
                       diag_offset, row_sum_part, row_pos_part, ld_rowpart, col_sum_part, ld_colpart, e_out, ld_e, stream);
}

extern "C" int evk_mpce_fwd_store_gathered(const void* q_hi, int64_t ld_q, const void* k_hi, int64_t ld_k, int64_t n_rows,
                                           int64_t n_cols, int64_t d, const uint32_t* bits, int64_t ld_words,
                                           float inv_tau, int flags, int64_t diag_offset, float* row_sum_part,
                                           float* row_pos_part, int64_t ld_rowpart, float* col_sum_part,
                                           int64_t ld_colpart, void* e_out, int64_t ld_e, const uint32_t* landed,
                                           const int* step, int* error, int64_t cols_per_source, int64_t first_col,
                                           int landed_per_step, evk_stream_t stream) {
  EVK_REQUIRE(landed && step && error, "evk_mpce_fwd_store_gathered: null flag pointers");
  EVK_REQUIRE(!e_out || evk_aligned16(e_out), "evk_mpce_fwd_store_gathered: e_out must be 16-byte aligned");
  return mpce_fwd_impl(q_hi, nullptr, ld_q, k_hi, nullptr, ld_k, n_rows, n_cols, d, bits, ld_words, inv_tau, flags,
                       diag_offset, row_sum_part, row_pos_part, ld_rowpart, col_sum_part, ld_colpart, e_out, ld_e, stream,
                       landed, step, error, cols_per_source, first_col, landed_per_step);
}

extern "C" int evk_mpce_bwd_w(const void* q_hi, const void* q_lo, int64_t ld_q, const void* k_hi, const void* k_lo,
                              int64_t ld_k, int64_t n_rows, int64_t n_cols, int64_t d, const uint32_t* bits,
                              int64_t ld_words, const int32_t* counts, const float* a_row, const float* b_col,
                              float inv_tau, int flags, int64_t diag_offset, void* w_hi, void* w_lo, int64_t ld_w,
                              evk_stream_t stream) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  const bool cta2 = use_cta_pairs();
  TcParams p;
  memset(&p, 0, sizeof(p));
  rc = setup_sim_operands(p, q_hi, q_lo, ld_q, k_hi, k_lo, ld_k, n_rows, n_cols, d, flags, cta2);
  if (rc != EVK_OK) return rc;
  const bool split = (flags & EVK_FLAG_SPLIT_BF16) != 0;
  EVK_REQUIRE(bits && counts && a_row && b_col && w_hi && (!split || w_lo), "evk_mpce_bwd_w: null pointer");
  EVK_REQUIRE(evk_aligned16(b_col), "evk_mpce_bwd_w: b_col must be 16-byte aligned");
  EVK_REQUIRE(ld_words >= (int64_t)p.n_tiles * (BN / 32), "evk_mpce_bwd_w: ld_words must cover whole 256-column tiles");
  EVK_REQUIRE(ld_w >= n_cols && ld_w % 8 == 0, "evk_mpce_bwd_w: ld_w=%lld must be >= n_cols and a multiple of 8", (long long)ld_w);
  EVK_REQUIRE(inv_tau > 0.f && inv_tau <= EVK_MAX_INV_TAU, "evk_mpce_bwd_w: 1/tau=%g outside (0, 40]", inv_tau);
  rc = make_map_bf16(&p.out_map[0], w_hi, n_rows, n_cols, ld_w, 32, 64);
  if (rc != EVK_OK) return rc;
  if (split) {
    rc = make_map_bf16(&p.out_map[1], w_lo, n_rows, n_cols, ld_w, 32, 64);
    if (rc != EVK_OK) return rc;
  }
  p.inv_tau = inv_tau;
  p.flags = flags;
  p.diag_offset = diag_offset;
  p.bits = bits;
  p.ld_words = ld_words;
  p.counts = counts;
  p.a_row = a_row;
  p.b_col = b_col;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  return cta2 ? launch<EPI_BWD_W, false, false, 5, true>(p, s) : launch<EPI_BWD_W, false, false, 3, false>(p, s);
}

namespace {
int gemm_common(TcParams& p, const void* const* a_ptrs, int64_t lda, bool a_mn, const void* const* b_ptrs,
                int64_t ldb, bool b_mn, int nsegs, int64_t m, int64_t n, int64_t k, float alpha, float* out,
                int64_t ld_out, int variant, int force_splits, bool cta2, cudaStream_t s) {
  if (p.rows_per_owner > 0) out = p.out_peer[0];
  EVK_REQUIRE(m > 0 && n > 0 && k > 0 && out, "gemm: empty problem or null output");
  EVK_REQUIRE(m < (1ll << 30) && n < (1ll << 30) && k < (1ll << 30), "gemm: problem too large");
  EVK_REQUIRE(ld_out >= n && ld_out % ((p.flags & kIntBf16Out) ? 8 : 4) == 0 && evk_aligned16(out),
              "gemm: out needs 16-byte alignment and 16-byte rows");
  p.num_segs = nsegs;
  p.kb_per_seg = (int)((k + BK - 1) / BK);
  for (int sg = 0; sg < nsegs; ++sg) {
    int rc = make_operand_map(&p.a_map[sg], a_ptrs[sg], a_mn, m, k, lda, BM);
    if (rc != EVK_OK) return rc;
    rc = make_operand_map(&p.b_map[sg], b_ptrs[sg], b_mn, n, k, ldb, cta2 ? BN / 2 : BN);
    if (rc != EVK_OK) return rc;
  }
  const int tile_m = cta2 ? 2 * BM : BM;
  p.m_tiles = (int)((m + tile_m - 1) / tile_m);
  p.n_tiles = (int)((n + BN - 1) / BN);
  p.n_rows = m;
  p.n_cols = n;
  const int total_kb = p.num_segs * p.kb_per_seg;
  int sm_avail = evk_sm_count();
  if (p.cta_limit > 0 && p.cta_limit < sm_avail) sm_avail = p.cta_limit & ~1;
  const int groups = cta2 ? sm_avail / 2 : sm_avail;
  p.splits = force_splits > 0 ? force_splits : (use_stream_k() ? 0 : choose_splits(p.m_tiles * p.n_tiles, total_kb, groups));
  if (p.flags & kIntStore) p.splits = 1;      // store epilogue: every output element is written by exactly one unit
  if (p.splits > total_kb) p.splits = total_kb;
  p.out = out;
  p.ld_out = ld_out;
  p.alpha = alpha;
  p.policy_a = kEvictFirst;                   // W: N^2 stream, each element used by n_tiles (3) column tiles only
  p.policy_b = kEvictLast;                    // X: re-read by every row tile
  p.policy_out = kEvictNormal;
  if (const char* e = getenv("EVK_L2_HINTS")) {
    if (e[0] == '0') p.policy_a = p.policy_b = kEvictNormal;
  }
  fill_descs(p, a_mn, b_mn, variant, cta2);
  if (variant & 8) p.flags |= kIntDropOut;      // bring-up: run the main loop, drop the output
  if (cta2 && use_tma_epilogue() && !(variant & 16) && (b_mn) ) {
    // output tiles leave through TMA (store, or reduce-add when units are split along K)
    int rc;
    if (p.rows_per_owner > 0) {
      const int owners = (int)((m + p.rows_per_owner - 1) / p.rows_per_owner);
      for (int o = 0; o < owners; ++o) {
        const int64_t rows_o = (o + 1) * p.rows_per_owner <= m ? p.rows_per_owner : m - o * p.rows_per_owner;
        rc = (p.flags & kIntBf16Out) ? make_map_bf16(&p.peer_map[o], p.out_peer[o], rows_o, n, ld_out, 32, 64)
                                : make_map_f32_out(&p.peer_map[o], p.out_peer[o], rows_o, n, ld_out);
        if (rc != EVK_OK) return rc;
      }
    } else {
      rc = make_map_f32_out(&p.peer_map[0], out, m, n, ld_out);
      if (rc != EVK_OK) return rc;
    }
    if (a_mn) return launch<EPI_GEMM_TMA, true, true, 5, true>(p, s);
    return launch<EPI_GEMM_TMA, false, true, 5, true>(p, s);
  }
  if (cta2) {
    if (a_mn && b_mn) return launch<EPI_GEMM, true, true, 6, true>(p, s);
    if (a_mn) return launch<EPI_GEMM, true, false, 6, true>(p, s);
    if (b_mn) return launch<EPI_GEMM, false, true, 6, true>(p, s);
    return launch<EPI_GEMM, false, false, 6, true>(p, s);
  }
  if (a_mn && b_mn) return launch<EPI_GEMM, true, true, 4, false>(p, s);
  if (a_mn) return launch<EPI_GEMM, true, false, 4, false>(p, s);
  if (b_mn) return launch<EPI_GEMM, false, true, 4, false>(p, s);
  return launch<EPI_GEMM, false, false, 4, false>(p, s);
}
}  // namespace

extern "C" int evk_mpce_bwd_gemm(const void* w_hi, const void* w_lo, int64_t ld_w, int64_t n_rows, int64_t n_cols,
                                 int transpose_w, const void* x_hi, const void* x_lo, int64_t ld_x, int64_t d,
                                 float alpha, int flags, float* out, int64_t ld_out, int cta_limit,
                                 evk_stream_t stream) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  const bool split = (flags & EVK_FLAG_SPLIT_BF16) != 0;
  EVK_REQUIRE(w_hi && x_hi && (!split || (w_lo && x_lo)), "evk_mpce_bwd_gemm: null operand");
  EVK_REQUIRE(cta_limit >= 0, "evk_mpce_bwd_gemm: negative cta_limit");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.cta_limit = cta_limit;
  const void* a_ptrs[3] = {w_hi, w_hi, w_lo};
  const void* b_ptrs[3] = {x_hi, x_lo, x_hi};
  // W is stored [n_rows, n_cols].  Not transposed: A = W is K-major (K = columns).  Transposed:
  // A = W^T is MN-major (M = columns contiguous, K = rows).  X [K, d] is always MN-major for B.
  const int64_t m = transpose_w ? n_cols : n_rows;
  const int64_t k = transpose_w ? n_rows : n_cols;
  return gemm_common(p, a_ptrs, ld_w, transpose_w != 0, b_ptrs, ld_x, true, split ? 3 : 1, m, d, k, alpha, out, ld_out,
                     0, 0, use_cta_pairs(), static_cast<cudaStream_t>(stream));
}

extern "C" int evk_mpce_bwd_gemm_scatter(const void* w_hi, const void* w_lo, int64_t ld_w, int64_t n_rows, int64_t n_cols,
                                         const void* x_hi, const void* x_lo, int64_t ld_x, int64_t d, float alpha,
                                         int flags, const uint64_t* out_ptrs, int n_owners, int64_t rows_per_owner,
                                         int64_t ld_out, int store, int first_owner, int cta_limit,
                                         evk_stream_t stream) {
  flags &= EVK_FLAG_PUBLIC_MASK;
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  const bool split = (flags & EVK_FLAG_SPLIT_BF16) != 0;
  EVK_REQUIRE(w_hi && x_hi && (!split || (w_lo && x_lo)) && out_ptrs, "evk_mpce_bwd_gemm_scatter: null operand");
  EVK_REQUIRE(n_owners >= 1 && n_owners <= 16 && rows_per_owner > 0 && rows_per_owner % 128 == 0 &&
                  (int64_t)n_owners * rows_per_owner >= n_cols,
              "evk_mpce_bwd_gemm_scatter: need 1..16 owners of rows_per_owner %% 128 == 0 rows covering all %lld columns",
              (long long)n_cols);
  TcParams p;
  memset(&p, 0, sizeof(p));
  for (int r = 0; r < n_owners; ++r) {
    p.out_peer[r] = reinterpret_cast<float*>(out_ptrs[r]);
    EVK_REQUIRE(p.out_peer[r] && evk_aligned16(p.out_peer[r]), "evk_mpce_bwd_gemm_scatter: owner buffers must be 16-byte aligned");
  }
  p.rows_per_owner = rows_per_owner;
  p.cta_limit = cta_limit;
  EVK_REQUIRE(first_owner >= 0 && first_owner < n_owners && cta_limit >= 0, "evk_mpce_bwd_gemm_scatter: first_owner / cta_limit out of range");
  {
    // the row blocks (keys) of owner `first_owner` are computed - and sent - first
    const int64_t tile_m = use_cta_pairs() ? 2 * BM : BM;
    p.rot_m_tiles = (int)(((int64_t)first_owner * rows_per_owner) / tile_m);
  }
  if (store) p.flags |= kIntStore;
  if (store == 2) {
    EVK_REQUIRE(use_cta_pairs() && use_tma_epilogue(), "evk_mpce_bwd_gemm_scatter: bf16 partials need the TMA epilogue (CTA pairs)");
    p.flags |= kIntBf16Out;
  }
  const void* a_ptrs[3] = {w_hi, w_hi, w_lo};
  const void* b_ptrs[3] = {x_hi, x_lo, x_hi};
  // dKhat partial of this rank's rows: A = W^T (MN-major), K = n_rows; out rows = the n_cols keys
  return gemm_common(p, a_ptrs, ld_w, true, b_ptrs, ld_x, true, split ? 3 : 1, n_cols, d, n_rows, alpha, nullptr, ld_out,
                     0, 0, use_cta_pairs(), static_cast<cudaStream_t>(stream));
}

// f4 (retrieval, reference modules/multiview/trainer.py:543-653): scores = A B^T for the exact inner-product top-k.
// Plain stores (every element written once, c needs no zero fill); a_lo / b_lo select the 3-segment split mode.
extern "C" int evk_tc_gemm_nt(const void* a_hi, const void* a_lo, int64_t lda, const void* b_hi, const void* b_lo,
                              int64_t ldb, int64_t m, int64_t n, int64_t k, float* c, int64_t ldc, evk_stream_t stream) {
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  EVK_REQUIRE(a_hi && b_hi && c && ((a_lo == nullptr) == (b_lo == nullptr)), "evk_tc_gemm_nt: null operand (a_lo / b_lo both or neither)");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.flags = kIntStore;
  const void* a_ptrs[3] = {a_hi, a_hi, a_lo};
  const void* b_ptrs[3] = {b_hi, b_lo, b_hi};
  return gemm_common(p, a_ptrs, lda, false, b_ptrs, ldb, false, a_lo ? 3 : 1, m, n, k, 1.f, c, ldc, 0, 0, use_cta_pairs(),
                     static_cast<cudaStream_t>(stream));
}

extern "C" int evk_tc_gemm_probe(const void* a, int64_t lda, int a_major, const void* b, int64_t ldb, int b_major,
                                 int64_t m, int64_t n, int64_t k, float* c, int64_t ldc, int variant, int splits,
                                 evk_stream_t stream) {
  int rc = check_device();
  if (rc != EVK_OK) return rc;
  EVK_REQUIRE(a && b, "evk_tc_gemm_probe: null operand");
  TcParams p;
  memset(&p, 0, sizeof(p));
  const void* a_ptrs[3] = {a, nullptr, nullptr};
  const void* b_ptrs[3] = {b, nullptr, nullptr};
  // variant bit 0: alternative MN-major descriptor; bit 1: force CTA pairs; bit 2: force single CTA
  const bool cta2 = (variant & 2) ? true : ((variant & 4) ? false : use_cta_pairs());
  return gemm_common(p, a_ptrs, lda, a_major != 0, b_ptrs, ldb, b_major != 0, 1, m, n, k, 1.f, c, ldc, variant, splits,
                     cta2, static_cast<cudaStream_t>(stream));
}
