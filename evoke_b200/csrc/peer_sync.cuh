// Cross-GPU synchronisation folded into the head of a consumer kernel (sharded path): instead of a separate
// barrier launch between the kernel that stored into peer memory and the first kernel that reads what the peers
// stored, the consumer itself signals ("everything this GPU wrote in earlier kernels of the stream is published")
// and waits for the same signal of every peer.  Same protocol as evk_peer_barrier (system-scope fence + release
// store into the peers' flag areas, acquire polling of the own area, timeout that raises the failure flag of EVERY
// rank), but the epoch is derived from the transport's step counter - per_step * (*step) + index - so that no kernel
// has to advance a counter that other CTAs of the same kernel still read.
#pragma once
#include "evk_common.cuh"

#include <string.h>

struct PeerSyncDev {
  uint32_t* flags[16];      // flags[t] = base of rank t's flag area (entry r is written by rank r only)
  int* err[16];             // failure flag of every rank
  int* err_host;            // pinned host mirror (may be null)
  const int* step;          // device step counter, advanced once per step by the prologue kernel
  int n, rank, index, per_step;
  unsigned long long timeout_ns;
};

// Called by ALL threads of a CTA (contains __syncthreads).  Exactly one CTA of the grid passes signaller = true.
__device__ __forceinline__ void peer_sync_block(const PeerSyncDev& ps, bool signaller) {
  if (ps.n <= 0) return;
  if (threadIdx.x < 32) {
    const int t = threadIdx.x;
    if (t < ps.n) {
      const uint32_t e = (uint32_t)ps.per_step * (uint32_t)(*ps.step) + (uint32_t)ps.index;
      if (signaller) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ps.flags[t] + ps.rank), "r"(e) : "memory");
      }
      const uint32_t* mine = ps.flags[ps.rank] + t;
      unsigned long long t0 = 0;
      for (uint32_t spins = 0;; ++spins) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
        if ((int32_t)(v - e) >= 0) break;
        unsigned long long now;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
        if (spins == 0) t0 = now;
        if (now - t0 > ps.timeout_ns) {
          for (int q = 0; q < ps.n; ++q)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ps.err[q]), "r"(1) : "memory");
          if (ps.err_host) *reinterpret_cast<volatile int*>(ps.err_host) = 1;
          break;
        }
        __nanosleep(64);
      }
    }
  }
  __syncthreads();
}

// host: evk_peer_sync_t (include/evoke_b200.h) -> device form; sync == nullptr disables it
static inline int peer_sync_from_host(const evk_peer_sync_t* sync, PeerSyncDev& out) {
  memset(&out, 0, sizeof(out));
  if (!sync) return EVK_OK;
  EVK_REQUIRE(sync->n_ranks >= 1 && sync->n_ranks <= 16 && sync->rank >= 0 && sync->rank < sync->n_ranks && sync->step &&
                  sync->per_step >= 1 && sync->index >= 1 && sync->index <= sync->per_step,
              "evk_peer_sync_t: bad ranks / step / index");
  for (int t = 0; t < sync->n_ranks; ++t) {
    out.flags[t] = reinterpret_cast<uint32_t*>(sync->flag_ptrs[t]);
    out.err[t] = reinterpret_cast<int*>(sync->err_ptrs[t]);
    EVK_REQUIRE(out.flags[t] && out.err[t], "evk_peer_sync_t: null flag / error pointer");
  }
  out.err_host = sync->err_host;
  out.step = sync->step;
  out.n = sync->n_ranks;
  out.rank = sync->rank;
  out.index = sync->index;
  out.per_step = sync->per_step;
  out.timeout_ns = (unsigned long long)(sync->timeout_ms > 0 ? sync->timeout_ms : 2000) * 1000000ull;
  return EVK_OK;
}
