// Peer-memory (NVLink / NVSwitch) transport for the sharded path: kernels that write their result
// straight into the symmetric buffers of EVERY rank, so the all-gather is part of the producing
// kernel instead of a separate NCCL launch.
//
//   evk_shard_prologue   : K1 of both sides fused with the all-gather of the normalised bf16 key rows and the ids
//                          (F.normalize of v0520.py:495-496, then what a sharded run must exchange)
//   evk_peer_push_shard  : the same all-gather as TMA bulk copies next to the similarity sweep (opt-in)
//   evk_peer_barrier / evk_mpce_shard_finish / evk_peer_* : ordering, statistics exchange, symmetric buffers
//
// Destination pointers are the per-rank base addresses of a symmetric allocation (peer-mapped device
// pointers; the local one included).  Cross-rank ordering (all shards landed / buffers free again) is
// the caller's job: a symmetric-memory barrier between producer and consumer kernels.
#include "evk_common.cuh"
#include "peer_sync.cuh"
#include "tc_ptx.cuh"

#include <string.h>

namespace {

constexpr int kMaxPeers = 16;
constexpr int kWarpsPerBlock = 8;

struct PeerDst {
  void* hi[kMaxPeers];
  void* lo[kMaxPeers];
  int n;
};

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Prologue of a sharded step in ONE launch: K1 of the key rows (text) stored into every rank's buffer
// (all-gather), K1 of the query rows (image) kept local, the id shard pushed to every rank, and the zero fill
// of the split-K accumulator of the local gradient contraction (same [n, D] shape as the query rows).
struct Prologue {
  // the embeddings as the caller holds them: fp32 / bf16 / fp16, any row and feature strides (the reference hands over
  // [:,0,:] views of the permuted projection-head output: feature stride 1 + P, v0520.py:484,399).  `*_fast`: fp32,
  // unit feature stride, 16-byte aligned rows -> 128-bit loads.
  const void* text; int64_t text_stride, text_cs; int text_dtype, text_fast;
  const void* image; int64_t image_stride, image_cs; int image_dtype, image_fast;
  PeerDst khat;                         // destinations of the normalised key rows
  __nv_bfloat16* q_hi;                  // local normalised query rows
  float* k_norm; float* q_norm;
  const int32_t* ids; const int32_t* ids2;
  int32_t* ids_dst[kMaxPeers]; int32_t* ids2_dst[kMaxPeers];
  float* zero; int64_t ld_zero; int zero_width;
  int32_t* zero_i32; int64_t n_zero_i32;  // K2's counts (K2 then needs no memset node)
  int64_t n, ld, row_offset;
  int d;
  int* step;                            // optional device counter advanced once per launch (one step = one epoch)
  const int* error;                     // optional: sticky failure flag of the transport (a barrier timed out)
};

template <int kIters>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
shard_prologue_kernel(const Prologue p) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * kWarpsPerBlock;
  // a peer missed a barrier earlier: its buffers may still be in use, so nothing is written into peer memory any
  // more (the step's loss and gradients are poisoned downstream, the host raises at its next entry)
  if (p.error && *reinterpret_cast<const volatile int*>(p.error) != 0) return;
  if (p.step && blockIdx.x == 0 && threadIdx.x == 0) *p.step += 1;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.n_zero_i32; t += (int64_t)gridDim.x * blockDim.x)
    p.zero_i32[t] = 0;
  // ids: one element per thread, pushed to every rank
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < p.n; t += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = __ldg(p.ids + t);
    const int32_t v2 = p.ids2 ? __ldg(p.ids2 + t) : 0;
    for (int q = 0; q < kMaxPeers; ++q) {
      if (p.ids_dst[q] == nullptr) break;
      p.ids_dst[q][p.row_offset + t] = v;
      if (p.ids2) p.ids2_dst[q][p.row_offset + t] = v2;
    }
  }
  for (int64_t rr = warp0; rr < 2 * p.n; rr += nwarps) {
    const bool is_text = rr < p.n;
    const int64_t r = is_text ? rr : rr - p.n;
    const void* xb = is_text ? p.text : p.image;
    const int64_t rs = is_text ? p.text_stride : p.image_stride, cs = is_text ? p.text_cs : p.image_cs;
    const int dt = is_text ? p.text_dtype : p.image_dtype;
    const bool fast = (is_text ? p.text_fast : p.image_fast) != 0;
    float v[kIters][8];
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < p.d) {
        if (fast) {
          const float* xr = static_cast<const float*>(xb) + r * rs;
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(xr + c));
          const float4 p1 = __ldg(reinterpret_cast<const float4*>(xr + c + 4));
          v[it][0] = p0.x; v[it][1] = p0.y; v[it][2] = p0.z; v[it][3] = p0.w;
          v[it][4] = p1.x; v[it][5] = p1.y; v[it][6] = p1.z; v[it][7] = p1.w;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[it][e] = load_as_float(xb, dt, r * rs + (int64_t)(c + e) * cs);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(v[it][e], v[it][e], ss);
      }
    }
    ss = warp_sum(ss);
    const float nrm = sqrtf(ss);
    const float den = fmaxf(nrm, EVK_NORM_EPS);
    if (lane == 0) (is_text ? p.k_norm : p.q_norm)[r] = nrm;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int c = (it * 32 + lane) * 8;
      if (c < p.d) {
        uint4 hi;
        hi.x = pack2(v[it][0] / den, v[it][1] / den); hi.y = pack2(v[it][2] / den, v[it][3] / den);
        hi.z = pack2(v[it][4] / den, v[it][5] / den); hi.w = pack2(v[it][6] / den, v[it][7] / den);
        if (is_text) {
          const int64_t off = (p.row_offset + r) * p.ld + c;
          for (int q = 0; q < p.khat.n; ++q)       // one 128-bit store per destination rank (NVLink for peers)
            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.khat.hi[q]) + off) = hi;
        } else {
          *reinterpret_cast<uint4*>(p.q_hi + r * p.ld + c) = hi;
        }
      }
    }
    if (!is_text && p.zero) {
      float4* z = reinterpret_cast<float4*>(p.zero + r * p.ld_zero);
      for (int c = lane; c * 4 < p.zero_width; c += 32) z[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Cross-GPU barrier over a symmetric flag area: flags[r] on every rank is written by rank r only.
// Thread t signals peer t (release at system scope, after a system fence that publishes this GPU's
// earlier peer writes) and then waits for peer t's signal in its own flag area.  The epoch lives in
// device memory and is advanced by the kernel itself, so the launch can sit in a replayed CUDA graph.
// A peer that never arrives does not hang the GPU: after timeout_ns the kernel raises *error and returns.
struct PeerFlags {
  uint32_t* flags[kMaxPeers];      // flags[t] = base of rank t's flag area (kMaxPeers uint32 slots)
  int* err[kMaxPeers];             // err[t] = rank t's failure flag (symmetric memory; err[rank] is this rank's own)
  int n, rank;
};

__global__ void __launch_bounds__(32)
peer_barrier_kernel(PeerFlags pf, uint32_t* __restrict__ epoch, int* __restrict__ error_host, uint64_t timeout_ns) {
  const int t = threadIdx.x;
  const uint32_t e = *epoch + 1u;
  __syncwarp();
  if (t < pf.n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.flags[t] + pf.rank), "r"(e) : "memory");
    const uint32_t* mine = pf.flags[pf.rank] + t;
    const uint64_t t0 = global_timer_ns();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
      if ((int32_t)(v - e) >= 0) break;
      if (global_timer_ns() - t0 > timeout_ns) {
        // Sticky, and raised on EVERY rank (the late peer passes its own barriers at once - this rank is ahead of it -
        // and would otherwise consume this rank's out-of-step buffers without noticing).  The consumers of the step
        // (evk_mpce_shard_finish, evk_l2norm_bwd_parts) turn the flag into NaN loss / gradients, the prologue stops
        // writing into peer memory, and the host sees the mirror without a sync.
        for (int q = 0; q < pf.n; ++q)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(pf.err[q]), "r"(1) : "memory");
        if (error_host) *reinterpret_cast<volatile int*>(error_host) = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncwarp();
  if (t == 0) *epoch = e;
}

// Closes the sharded forward once every rank's statistics slot has landed (slot r = rank r's partial
// column exp-sums over its rows [n_cols floats] followed by its row-side loss term [1 float]):
//   b_col[j] = 1 / sum_r slot_r[j]        loss = sum_r slot_r[n_cols] + inv_count * sum_j (shift + ln C_j)
// Multi-CTA, fixed summation order (slots in rank order; per-CTA fp64 partials added up in index order by the
// last CTA to finish): every rank gets the same bits.
constexpr int kFinishThreads = 256;
__global__ void __launch_bounds__(kFinishThreads)
shard_finish_kernel(const float* __restrict__ slots, int n_slots, int64_t ld_slot, int64_t n_cols, float shift,
                    double inv_count, float* __restrict__ b_col, float* __restrict__ loss_out,
                    double* __restrict__ cta_partial, unsigned int* __restrict__ ticket, const int* __restrict__ error,
                    int* __restrict__ error_host, const PeerSyncDev sync) {
  __shared__ double s_part[kFinishThreads / 32];
  __shared__ bool s_last;
  peer_sync_block(sync, blockIdx.x == 0);          // the slots were stored by the peers' statistics kernels
  const int64_t j = (int64_t)blockIdx.x * kFinishThreads + threadIdx.x;
  // a barrier of this step timed out: some slot / key rows may be stale -> the result must not look valid
  const bool poisoned = error && *reinterpret_cast<const volatile int*>(error) != 0;
  if (poisoned && error_host && blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile int*>(error_host) = 1;
  double acc = 0.0;
  if (j < n_cols) {
    float c = 0.f;
    for (int r = 0; r < n_slots; ++r) c += slots[r * ld_slot + j];
    b_col[j] = poisoned ? __int_as_float(0x7fc00000) : 1.f / c;
    acc = (double)shift + (double)logf(c);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kFinishThreads / 32; ++w) t += s_part[w];
    cta_partial[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (s_last && threadIdx.x < 32) {                  // the last CTA adds the partials up in index order
    __threadfence();
    double t = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) t += reinterpret_cast<volatile double*>(cta_partial)[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) {
      t *= inv_count;
      for (int r = 0; r < n_slots; ++r) t += (double)slots[r * ld_slot + n_cols];
      loss_out[0] = poisoned ? __int_as_float(0x7fc00000) : (float)t;
      *ticket = 0u;                                  // a persistent workspace is ready for the next call
    }
  }
}

// All-gather that runs NEXT TO the similarity sweep: this rank's shard (already in its own buffer) is copied to the
// peers one destination at a time, in the order rank+1, rank+2, ... - at any moment every GPU receives from exactly
// one source, so the shards land in a known order and K3 (tc_engine.cu) consumes the column blocks in that order,
// waiting per source.  The copy must not compete with K3 for issue slots or registers: ONE thread per CTA drives TMA
// bulk copies (global -> 8 KiB of shared memory -> peer memory over NVLink, two stages), nothing else runs.  A CTA
// that has finished its chunks for a destination waits for its bulk stores to complete, fences at system scope and
// adds 1 to landed[source] on that destination; a source has landed when the counter reaches CTAs * step.
constexpr int kPushChunk = 8192;
constexpr int kPushStages = 2;

struct PushArgs {
  uint8_t* dst[kMaxPeers];
  uint32_t* landed[kMaxPeers];          // landed[t] = base of rank t's landed-counter area (entry s: source s)
  int n, rank;
};

__global__ void __launch_bounds__(32)
peer_push_bulk_kernel(const uint8_t* __restrict__ src, int64_t bytes, PushArgs a, int64_t dst_offset) {
  extern __shared__ __align__(128) uint8_t push_stage[];
  __shared__ __align__(8) uint64_t bars[kPushStages];
  if (threadIdx.x != 0) return;
  const uint32_t sbase = (tc::smem_u32(push_stage) + 127u) & ~127u;
  for (int st = 0; st < kPushStages; ++st) tc::mbar_init(tc::smem_u32(&bars[st]), 1);
  tc::fence_mbar_init();
  // this rank's own rows were written by the prologue, earlier in stream order
  if (blockIdx.x == 0) atomicAdd_system(a.landed[a.rank] + a.rank, gridDim.x);
  const int64_t n_chunks = (bytes + kPushChunk - 1) / kPushChunk;
  uint32_t phase = 0;                   // bit st = parity of stage st's barrier
  int it = 0;
  for (int k = 1; k < a.n; ++k) {
    const int t = (a.rank + k) % a.n;
    uint8_t* out = a.dst[t] + dst_offset;
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x, ++it) {
      const int st = it % kPushStages;
      const uint32_t sm = sbase + st * kPushChunk, bar = tc::smem_u32(&bars[st]);
      const int64_t off = c * kPushChunk;
      const uint32_t len = (uint32_t)((bytes - off) < kPushChunk ? (bytes - off) : kPushChunk);
      // the bulk store that used this stage two chunks ago has finished READING it
      asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPushStages - 1) : "memory");
      tc::mbar_expect_tx(bar, len);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sm), "l"(src + off), "r"(len), "r"(bar) : "memory");
      tc::mbar_wait(bar, (phase >> st) & 1u);
      phase ^= 1u << st;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + off), "r"(sm), "r"(len) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // this CTA's stores to destination t are complete
    __threadfence_system();
    atomicAdd_system(a.landed[t] + a.rank, 1u);
  }
}

// waits until every source's shard of this step has landed here (side-stream consumers of the gathered rows)
__global__ void __launch_bounds__(32)
peer_wait_landed_kernel(const uint32_t* __restrict__ landed, int n, const int* __restrict__ step, int per_step,
                        int* __restrict__ error, uint64_t timeout_ns) {
  const int t = threadIdx.x;
  if (t < n) {
    const uint32_t epoch = (uint32_t)*step * (uint32_t)per_step;     // landed[s] counts the push CTAs of source s
    const uint64_t t0 = global_timer_ns();
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(landed + t) : "memory");
      if ((int32_t)(v - epoch) >= 0) break;
      if (global_timer_ns() - t0 > timeout_ns) {
        atomicExch(error, 1);
        break;
      }
      __nanosleep(64);
    }
  }
}

int fill_dst(PeerDst& dst, int n_dst, const uint64_t* hi, const uint64_t* lo) {
  EVK_REQUIRE(n_dst >= 1 && n_dst <= kMaxPeers && hi, "peer destinations: need 1..%d base pointers", kMaxPeers);
  memset(&dst, 0, sizeof(dst));
  dst.n = n_dst;
  for (int p = 0; p < n_dst; ++p) {
    dst.hi[p] = reinterpret_cast<void*>(hi[p]);
    dst.lo[p] = lo ? reinterpret_cast<void*>(lo[p]) : nullptr;
    EVK_REQUIRE(dst.hi[p] && evk_aligned16(dst.hi[p]) && (!lo || (dst.lo[p] && evk_aligned16(dst.lo[p]))),
                "peer destinations must be non-null and 16-byte aligned");
  }
  return EVK_OK;
}

}  // namespace

// ---- symmetric buffers: allocation and CUDA-IPC exchange ------------------------------------------
extern "C" int evk_peer_alloc(int64_t bytes, void** ptr_out) {
  EVK_REQUIRE(bytes > 0 && ptr_out, "evk_peer_alloc: bad arguments");
  void* p = nullptr;
  EVK_CUDA(cudaMalloc(&p, (size_t)bytes));
  EVK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  *ptr_out = p;
  return EVK_OK;
}

extern "C" int evk_peer_free(void* ptr) {
  if (ptr) EVK_CUDA(cudaFree(ptr));
  return EVK_OK;
}

extern "C" int evk_peer_export(const void* ptr, void* handle_out) {
  EVK_REQUIRE(ptr && handle_out, "evk_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t h;
  EVK_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle_out, &h, sizeof(h));
  return EVK_OK;
}

extern "C" int evk_peer_open(const void* handle, void** ptr_out) {
  EVK_REQUIRE(handle && ptr_out, "evk_peer_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  EVK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return EVK_OK;
}

extern "C" int evk_peer_close(void* ptr) {
  if (ptr) EVK_CUDA(cudaIpcCloseMemHandle(ptr));
  return EVK_OK;
}

extern "C" int evk_peer_barrier(const uint64_t* flag_ptrs, const uint64_t* error_ptrs, int n_ranks, int rank,
                                uint32_t* epoch, int* error_host, int64_t timeout_ms, evk_stream_t stream) {
  EVK_REQUIRE(flag_ptrs && error_ptrs && epoch && n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks,
              "evk_peer_barrier: bad arguments (1..%d ranks)", kMaxPeers);
  PeerFlags pf;
  memset(&pf, 0, sizeof(pf));
  pf.n = n_ranks;
  pf.rank = rank;
  for (int t = 0; t < n_ranks; ++t) {
    pf.flags[t] = reinterpret_cast<uint32_t*>(flag_ptrs[t]);
    pf.err[t] = reinterpret_cast<int*>(error_ptrs[t]);
    EVK_REQUIRE(pf.flags[t] && pf.err[t], "evk_peer_barrier: null flag area");
  }
  const uint64_t timeout_ns = (uint64_t)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pf, epoch, error_host, timeout_ns);
  EVK_CHECK_LAUNCH("peer_barrier");
  return EVK_OK;
}

extern "C" int evk_mpce_shard_finish(const float* slots, int n_slots, int64_t ld_slot, int64_t n_cols, float shift,
                                     double inv_count, float* b_col, float* loss_out, void* workspace,
                                     int64_t workspace_bytes, int workspace_persistent, const int* error, int* error_host,
                                     const evk_peer_sync_t* sync, evk_stream_t stream) {
  PeerSyncDev ps;
  {
    int rc = peer_sync_from_host(sync, ps);
    if (rc != EVK_OK) return rc;
  }
  EVK_REQUIRE(slots && b_col && loss_out && n_slots >= 1 && n_cols > 0 && ld_slot > n_cols,
              "evk_mpce_shard_finish: bad arguments (ld_slot must exceed n_cols: the loss term follows the column sums)");
  const int64_t blocks = (n_cols + kFinishThreads - 1) / kFinishThreads;
  EVK_REQUIRE(workspace && evk_aligned16(workspace) && workspace_bytes >= 16 + 8 * blocks,
              "evk_mpce_shard_finish: workspace needs %lld bytes, 16-byte aligned", (long long)(16 + 8 * blocks));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned int* ticket = static_cast<unsigned int*>(workspace);
  double* partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + 16);
  if (!workspace_persistent) EVK_CUDA(cudaMemsetAsync(ticket, 0, 16, s));
  shard_finish_kernel<<<(unsigned)blocks, kFinishThreads, 0, s>>>(slots, n_slots, ld_slot, n_cols, shift, inv_count, b_col,
                                                                 loss_out, partial, ticket, error, error_host, ps);
  EVK_CHECK_LAUNCH("shard_finish");
  return EVK_OK;
}

extern "C" int evk_shard_prologue(const void* text, int text_dtype, int64_t text_stride, int64_t text_col_stride,
                                  const void* image, int image_dtype, int64_t image_stride, int64_t image_col_stride,
                                  int64_t n_rows, int64_t d, int n_dst, const uint64_t* khat_ptrs, int64_t ld_bf16,
                                  int64_t row_offset, float* k_norm, void* q_hi, float* q_norm, const int32_t* ids,
                                  const int32_t* ids2, int n_ids_dst, const uint64_t* ids_ptrs, const uint64_t* ids2_ptrs,
                                  float* zero_buf, int64_t ld_zero, int32_t* zero_i32, int64_t n_zero_i32,
                                  int* step_counter, const int* error, evk_stream_t stream) {
  EVK_REQUIRE(text && image && k_norm && q_hi && q_norm && ids && ids_ptrs && khat_ptrs && n_rows > 0 && d > 0,
              "evk_shard_prologue: null pointer or empty shape");
  EVK_REQUIRE(text_dtype >= EVK_DTYPE_F32 && text_dtype <= EVK_DTYPE_F16 && image_dtype >= EVK_DTYPE_F32 &&
                  image_dtype <= EVK_DTYPE_F16, "evk_shard_prologue: bad dtype");
  EVK_REQUIRE(d % 8 == 0 && d <= 2048 && evk_aligned16(q_hi) && ld_bf16 >= d && ld_bf16 % 8 == 0 && row_offset >= 0,
              "evk_shard_prologue: needs d %% 8 == 0, d <= 2048, 16-byte aligned bf16 outputs");
  auto fast = [](const void* x, int dt, int64_t rs, int64_t cs) {
    return dt == EVK_DTYPE_F32 && cs == 1 && rs % 4 == 0 && evk_aligned16(x);
  };
  EVK_REQUIRE((ids2 == nullptr) == (ids2_ptrs == nullptr), "evk_shard_prologue: ids2 / ids2_ptrs must both be set or both null");
  EVK_REQUIRE(!zero_buf || (ld_zero % 4 == 0 && ld_zero >= d && evk_aligned16(zero_buf)), "evk_shard_prologue: bad zero buffer");
  Prologue p;
  memset(&p, 0, sizeof(p));
  int rc = fill_dst(p.khat, n_dst, khat_ptrs, nullptr);
  if (rc != EVK_OK) return rc;
  EVK_REQUIRE(n_ids_dst >= 1 && n_ids_dst <= kMaxPeers, "evk_shard_prologue: 1..%d id destinations", kMaxPeers);
  p.step = step_counter;
  p.error = error;
  for (int q = 0; q < n_ids_dst; ++q) {
    p.ids_dst[q] = reinterpret_cast<int32_t*>(ids_ptrs[q]);
    p.ids2_dst[q] = ids2_ptrs ? reinterpret_cast<int32_t*>(ids2_ptrs[q]) : nullptr;
    EVK_REQUIRE(p.ids_dst[q] && (!ids2_ptrs || p.ids2_dst[q]), "evk_shard_prologue: null id destination");
  }
  p.text = text; p.text_stride = text_stride; p.text_cs = text_col_stride; p.text_dtype = text_dtype;
  p.text_fast = fast(text, text_dtype, text_stride, text_col_stride) ? 1 : 0;
  p.image = image; p.image_stride = image_stride; p.image_cs = image_col_stride; p.image_dtype = image_dtype;
  p.image_fast = fast(image, image_dtype, image_stride, image_col_stride) ? 1 : 0;
  p.q_hi = static_cast<__nv_bfloat16*>(q_hi); p.k_norm = k_norm; p.q_norm = q_norm;
  p.ids = ids; p.ids2 = ids2;
  p.zero = zero_buf; p.ld_zero = ld_zero; p.zero_width = (int)(((d + 3) / 4) * 4);
  p.zero_i32 = zero_i32; p.n_zero_i32 = zero_i32 ? n_zero_i32 : 0;
  p.n = n_rows; p.ld = ld_bf16; p.row_offset = row_offset; p.d = (int)d;
  int64_t blocks = (2 * n_rows + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)evk_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (d <= 1024) shard_prologue_kernel<4><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  else shard_prologue_kernel<8><<<(unsigned)blocks, kWarpsPerBlock * 32, 0, s>>>(p);
  EVK_CHECK_LAUNCH("shard_prologue");
  return EVK_OK;
}

extern "C" int evk_peer_push_shard(const void* src, int64_t bytes, int n_ranks, int rank, const uint64_t* dst_ptrs,
                                   int64_t dst_offset_bytes, const uint64_t* landed_ptrs, int n_ctas,
                                   evk_stream_t stream) {
  EVK_REQUIRE(src && dst_ptrs && landed_ptrs && bytes > 0 && bytes % 16 == 0 && dst_offset_bytes >= 0 &&
                  dst_offset_bytes % 16 == 0 && evk_aligned16(src),
              "evk_peer_push_shard: bad arguments (16-byte aligned multiples)");
  EVK_REQUIRE(n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks && n_ctas >= 1 && n_ctas <= 1024,
              "evk_peer_push_shard: bad rank / world / CTA count");
  PushArgs a;
  memset(&a, 0, sizeof(a));
  a.n = n_ranks;
  a.rank = rank;
  for (int t = 0; t < n_ranks; ++t) {
    a.dst[t] = reinterpret_cast<uint8_t*>(dst_ptrs[t]);
    a.landed[t] = reinterpret_cast<uint32_t*>(landed_ptrs[t]);
    EVK_REQUIRE(a.dst[t] && a.landed[t] && evk_aligned16(a.dst[t]), "evk_peer_push_shard: null / misaligned destination");
  }
  peer_push_bulk_kernel<<<(unsigned)n_ctas, 32, kPushStages * kPushChunk + 128, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(src), bytes, a, dst_offset_bytes);
  EVK_CHECK_LAUNCH("peer_push_shard");
  return EVK_OK;
}

extern "C" int evk_peer_wait_landed(const void* landed, int n_ranks, const int* step, int per_step, int* error,
                                    int64_t timeout_ms, evk_stream_t stream) {
  EVK_REQUIRE(landed && step && error && n_ranks >= 1 && n_ranks <= kMaxPeers && per_step >= 1, "evk_peer_wait_landed: bad arguments");
  const uint64_t timeout_ns = (uint64_t)(timeout_ms > 0 ? timeout_ms : 2000) * 1000000ull;
  peer_wait_landed_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint32_t*>(landed), n_ranks,
                                                                        step, per_step, error, timeout_ns);
  EVK_CHECK_LAUNCH("peer_wait_landed");
  return EVK_OK;
}
