"""Symmetric peer-mapped buffers for the sharded loss (one process per GPU, NVLink / NVSwitch).

Every rank allocates the same buffer layout with ``evk_peer_alloc`` (a plain cudaMalloc: CUDA-IPC
handles name whole allocations), the 64-byte IPC handles travel through ``torch.distributed`` once,
and each rank maps all peers' buffers.  From then on the exchange steps of the sharded path are
stores / fp32 red.add issued by this library's own kernels straight into peer memory, ordered by a
flag barrier kernel (``evk_peer_barrier``) - no NCCL call sits on the data path:

  khat   [N, ld]  bf16    K1 writes this rank's normalised rows into EVERY rank's copy (all-gather)
  ids    [N] (+ids2 [N])  the id shards, pushed the same way
  slots  [R, N+4] fp32    slot r = rank r's partial column exp-sums + its row-side loss term
  dk_parts [R, n, D] bf16 (or fp32) this rank's rows of dKhat, one partial per source rank: rank s's K4b epilogue
                          stores its tiles for these rows into part s (posted NVLink stores), K1b adds them up
  flags  [16]     uint32  barrier flags (entry r written by rank r only)
  err    [1]      int32   failure flag: a barrier that times out on ANY rank raises it on EVERY rank
  landed [16]     uint32  landed[s] = push CTAs of source s that have delivered their rows (monotonic; opt-in overlapped gather)
  sync   [16]     uint32  flags of the syncs folded into the consumer kernels (entry r written by rank r only)

Only torch.distributed's object all-gather is used, once per context, for the handles.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib

_ALIGN = 1024
# EVOKE_B200_FOLDED_SYNC=0: separate evk_peer_barrier launches between producer and consumer kernels (round 1)
FOLDED_SYNC = os.environ.get("EVOKE_B200_FOLDED_SYNC", "1") == "1"


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _RawCuda:
    """Zero-copy view of raw device memory for torch.as_tensor (CUDA array interface, uint8)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class PeerContext:
    def __init__(self, group, n_local: int, d: int, device: torch.device, two_keys: bool = False,
                 timeout_ms: int = 10000, exchange: str = "bf16"):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > 16:
            raise ValueError("the peer-memory transport supports up to 16 ranks")
        if n_local % 128 or d % 8:
            raise ValueError("the peer-memory transport needs n_local % 128 == 0 and d % 8 == 0")
        self.device = torch.device(device)
        self.n_local, self.d = n_local, d
        self.n_total = n_local * self.world
        self.ld = _round_up(d, 8)
        self.exchange = exchange                       # dtype of the dKhat partials that cross NVLink
        self.width = _round_up(d, 8) if exchange == "bf16" else _round_up(d, 4)
        esize = 2 if exchange == "bf16" else 4
        self.ld_slot = _round_up(self.n_total + 4, 4)
        self.timeout_ms = timeout_ms
        n, big_n, r = n_local, self.n_total, self.world
        off = 0
        self.off = {}
        for name, nbytes in (("khat", big_n * self.ld * 2), ("ids", big_n * 4), ("ids2", big_n * 4 if two_keys else 0),
                             ("slots", r * self.ld_slot * 4), ("dk_parts", r * n * self.width * esize), ("flags", 64), ("landed", 64), ("err", 64), ("sync", 64)):
            self.off[name] = off
            off += _round_up(nbytes, _ALIGN)
        self.nbytes = off
        # Collectives must be entered by every rank whatever fails locally: allocate/export (may fail), exchange
        # (always), map (may fail); get_context() then agrees on success with one all-reduce.
        self.base, self.bases, self.failure = 0, [], None
        lib = _lib.load()
        mine = None
        with torch.cuda.device(self.device):
            try:
                base = ctypes.c_void_p()
                _lib.check(lib.evk_peer_alloc(self.nbytes, ctypes.byref(base)), "evk_peer_alloc")
                self.base = int(base.value)
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.evk_peer_export(ctypes.c_void_p(self.base), handle), "evk_peer_export")
                mine = bytes(handle.raw)
            except (RuntimeError, ValueError) as e:
                self.failure = repr(e)
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=group)
            if self.failure is None and any(h is None for h in handles):
                self.failure = "a peer could not export its buffer"
            if self.failure is None:
                try:
                    for t, h in enumerate(handles):
                        if t == self.rank:
                            self.bases.append(self.base)
                            continue
                        p = ctypes.c_void_p()
                        _lib.check(lib.evk_peer_open(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), "evk_peer_open")
                        self.bases.append(int(p.value))
                except (RuntimeError, ValueError) as e:
                    self.failure = repr(e)
        if self.failure is not None:
            return
        self._raw = torch.as_tensor(_RawCuda(self.base, self.nbytes), device=self.device)
        self.khat = self._view("khat", big_n * self.ld * 2).view(torch.bfloat16).view(big_n, self.ld)
        self.ids = self._view("ids", big_n * 4).view(torch.int32)
        self.ids2 = self._view("ids2", big_n * 4).view(torch.int32) if two_keys else None
        self.slots = self._view("slots", r * self.ld_slot * 4).view(torch.float32).view(r, self.ld_slot)
        self.dk_parts = self._view("dk_parts", r * n * self.width * esize).view(
            torch.bfloat16 if exchange == "bf16" else torch.float32).view(r, n, self.width)
        self.landed = self._view("landed", 64).view(torch.int32)
        self.step = torch.zeros(1, dtype=torch.int32, device=self.device)          # advanced by the prologue kernel
        self.epoch = torch.zeros(1, dtype=torch.int32, device=self.device)
        # failure flag of the transport (a peer missed a barrier): sticky device int read by the step's closing kernels
        # (NaN loss / gradients), mirrored by the barrier kernel into pinned HOST memory so that the next step's entry
        # can raise without synchronising the device
        self.error = self._view("err", 64).view(torch.int32)[:1]       # in symmetric memory: any rank's barrier may raise it
        self.error_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        # persistent, zero-initialised workspaces of the statistics kernels (their tickets reset themselves)
        self.ws_stats = torch.zeros(_lib.size("evk_stats_workspace_bytes", n, big_n), dtype=torch.uint8, device=self.device)
        self.ws_finish = torch.zeros(_lib.size("evk_shard_finish_workspace_bytes", big_n), dtype=torch.uint8, device=self.device)
        self._syncs = {}
        self.ptrs = {name: (ctypes.c_uint64 * self.world)(*[b + o for b in self.bases]) for name, o in self.off.items()}
        # where THIS rank's partial for owner t goes: part `rank` of t's dk_parts
        mine = self.rank * n * self.width * esize
        self.ptrs["khat_local"] = (ctypes.c_uint64 * 1)(self.base + self.off["khat"])
        self.ptrs["dk_mine"] = (ctypes.c_uint64 * self.world)(*[b + self.off["dk_parts"] + mine for b in self.bases])
        torch.cuda.synchronize(self.device)

    def _view(self, name: str, nbytes: int) -> torch.Tensor:
        o = self.off[name]
        return self._raw[o:o + nbytes]

    def table(self, name: str):
        """HOST array of the per-rank device addresses of one buffer (argument of the C entry points)."""
        return self.ptrs[name]

    SYNCS_PER_STEP = 3

    def sync(self, index: int):
        """ctypes pointer to the evk_peer_sync_t of the step's index-th sync point (1: ids / key rows landed, 2:
        statistics slots landed, 3: gradient partials landed and the gathered rows are free again), folded into the
        head of its first consumer kernel instead of a barrier launch."""
        if not FOLDED_SYNC:
            return None
        hit = self._syncs.get(index)
        if hit is None:
            ps = _lib.PeerSync()
            for t in range(self.world):
                ps.flag_ptrs[t] = self.bases[t] + self.off["sync"]
                ps.err_ptrs[t] = self.bases[t] + self.off["err"]
            ps.err_host = self.error_host.data_ptr()
            ps.step = self.step.data_ptr()
            ps.n_ranks, ps.rank, ps.index, ps.per_step = self.world, self.rank, index, self.SYNCS_PER_STEP
            ps.timeout_ms = self.timeout_ms
            hit = (ps, ctypes.byref(ps))
            self._syncs[index] = hit
        return hit[1]

    def barrier(self) -> None:
        """Enqueue the cross-GPU barrier on the current stream."""
        _lib.call("evk_peer_barrier", self.ptrs["flags"], self.ptrs["err"], self.world, self.rank, self.epoch.data_ptr(),
                  self.error_host.data_ptr(), self.timeout_ms, torch.cuda.current_stream().cuda_stream)

    FAILED = ("evoke_b200: a peer did not reach a cross-GPU barrier within {:.1f} s; the loss and gradients of that step "
              "are NaN and this transport context is dead (the peers' buffers may be out of step): tear the process "
              "group down")

    def raise_if_failed(self) -> None:
        """Host-side check WITHOUT a device sync (reads the pinned mirror): called at the entry of every step."""
        if int(self.error_host[0]) != 0:
            raise RuntimeError(self.FAILED.format(self.timeout_ms / 1e3))

    def check(self) -> None:
        """Raise if a barrier timed out (synchronises the device)."""
        if int(self.error.item()) != 0 or int(self.error_host[0]) != 0:
            raise RuntimeError(self.FAILED.format(self.timeout_ms / 1e3))

    def close(self) -> None:
        lib = _lib.load()
        torch.cuda.synchronize(self.device)
        for t, b in enumerate(self.bases):
            if t != self.rank and b:
                lib.evk_peer_close(ctypes.c_void_p(b))
        self.bases = []
        if self.base:
            self._raw = self.khat = self.ids = self.ids2 = self.slots = self.dk_parts = self.landed = self.error = None
            lib.evk_peer_free(ctypes.c_void_p(self.base))
            self.base = 0


_CONTEXTS: dict = {}


def get_context(group, n_local: int, d: int, device: torch.device, two_keys: bool = False,
                exchange: str = "bf16", timeout_ms: Optional[int] = None) -> Optional[PeerContext]:
    """Cached context for (group, shard shape); None if peer mapping is not possible here (the caller
    then uses the NCCL transport).  Collective: every rank must call it with the same arguments."""
    key = (id(group), n_local, d, torch.device(device).index, two_keys, exchange)
    if key in _CONTEXTS:
        return _CONTEXTS[key]
    if timeout_ms is None:
        timeout_ms = int(os.environ.get("EVOKE_B200_PEER_TIMEOUT_MS", "10000"))
    ctx: Optional[PeerContext] = PeerContext(group, n_local, d, device, two_keys, timeout_ms=timeout_ms, exchange=exchange)
    if ctx.failure is not None:                                  # IPC refused (container policy, no P2P, ...)
        _CONTEXTS["last_error"] = ctx.failure
    # all or nothing; also the barrier after which every rank has mapped every buffer
    flag = torch.tensor([0 if ctx.failure is not None else 1], device=device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        ctx.close()
        ctx = None
    _CONTEXTS[key] = ctx
    return ctx


def close_all() -> None:
    for k, c in list(_CONTEXTS.items()):
        if isinstance(c, PeerContext):
            c.close()
        _CONTEXTS.pop(k, None)
