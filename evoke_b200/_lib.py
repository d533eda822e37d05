"""ctypes binding of libevoke_b200.so (the C ABI declared in include/evoke_b200.h).

The library is loaded lazily on first use and there is NO fallback: if it is missing or a
call fails, a RuntimeError/ValueError carrying ``evk_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libevoke_b200.so")

EVK_OK, EVK_ERR_INVALID, EVK_ERR_CUDA, EVK_ERR_UNSUPPORTED = 0, -1, -2, -3
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
FLAG_EXCLUDE_DIAG, FLAG_NO_COLSUM, FLAG_SPLIT_BF16, FLAG_NO_POS, FLAG_AVGPOS = 1, 2, 4, 8, 16
ABI_VERSION = 3

P, I, L, F, D = c_void_p, c_int, c_int64, c_float, c_double

# name -> argtypes, in header order.  tests/test_abi.py checks this table against the header.
SIGNATURES = {
    "evk_version": [],
    "evk_mpce_row_parts": [],
    "evk_last_error": [],
    "evk_device_info": [I, P, P, P],
    "evk_stats_workspace_bytes": [L, L],
    "evk_shard_finish_workspace_bytes": [L],
    "evk_posmask_ld_words": [L],
    "evk_mpce_rowpart_rows": [L],
    "evk_mpce_colpart_rows": [L],
    "evk_mpce_strip_ld": [L],
    "evk_l2norm_fwd": [P, I, L, L, L, L, P, P, L, P, P, L, P, P],
    "evk_l2norm_bwd": [P, I, L, L, L, L, P, P, P, L, P, F, P, I, L, I, P],
    "evk_posmask_build": [P, P, L, P, P, L, L, I, P, L, P, P, I, I, P, P],
    "evk_mpce_small_fwd": [P, L, P, L, L, L, L, P, L, F, I, L, P, P, P],
    "evk_mpce_small_bwd": [P, L, P, L, L, L, L, P, L, P, P, P, F, I, L, P, L, P, P, P],
    "evk_mpce_finalize_avgpos": [P, P, P, L, F, D, P, P, P, I, P],
    "evk_mpce_small_fwd_batched": [P, L, L, P, L, L, L, L, L, L, P, L, F, I, P, P, L, P],
    "evk_mpce_small_bwd_batched": [P, L, L, P, L, L, L, L, L, L, P, L, P, P, P, L, F, I, P, L, L, P],
    "evk_token_sim_fwd": [P, P, L, L, L, F, P, P, P, P, P],
    "evk_token_sim_bwd": [P, P, P, P, P, L, L, L, P, P, P],
    "evk_local_attend_fwd": [P, P, L, L, L, L, P, P, P],
    "evk_local_attend_bwd": [P, P, P, P, L, L, L, L, P, P, P, P],
    "evk_reduce_partials": [P, L, L, L, P, P, P],
    "evk_mpce_finalize": [P, P, P, L, P, L, L, L, F, F, D, P, P, P, P],
    "evk_mpce_stats_fused": [P, L, L, P, L, L, P, L, P, L, L, L, L, L, F, F, D, P, P, P, P, L, P],
    "evk_mpce_pos": [P, P, L, P, P, L, L, L, L, P, L, F, P, P],
    "evk_mpce_fwd": [P, P, L, P, P, L, L, L, L, P, L, F, I, L, P, P, L, P, L, P],
    "evk_mpce_bwd_w": [P, P, L, P, P, L, L, L, L, P, L, P, P, P, F, I, L, P, P, L, P],
    "evk_mpce_bwd_gemm": [P, P, L, L, L, I, P, P, L, L, F, I, P, L, I, P],
    "evk_mpce_fwd_store": [P, L, P, L, L, L, L, P, L, F, I, L, P, P, L, P, L, P, L, P],
    "evk_mpce_w_from_e": [P, L, L, L, P, L, P, P, P, P, L, P, L, L, F, P, P, I, P, P, P, P, L, I, P],
    "evk_mpce_pos_logits": [P, L, P, L, L, L, P, P, I, P, P],
    "evk_mpce_pos_from_lists": [P, L, P, L, L, L, L, P, P, P, P, L, I, P, P, I, F, P, P],
    "evk_shard_prologue": [P, I, L, L, P, I, L, L, L, L, I, P, L, L, P, P, P, P, P, I, P, P, P, L, P, L, P, P, P],
    "evk_peer_push_shard": [P, L, I, I, P, L, P, I, P],
    "evk_peer_wait_landed": [P, I, P, I, P, L, P],
    "evk_mpce_fwd_store_gathered": [P, L, P, L, L, L, L, P, L, F, I, L, P, P, L, P, L, P, L, P, P, P, L, L, I, P],
    "evk_mpce_shard_stats_push": [P, L, L, P, L, L, P, L, P, L, L, L, F, F, D, P, P, I, L, P, L, I, P],
    "evk_peer_alloc": [L, P],
    "evk_peer_free": [P],
    "evk_peer_export": [P, P],
    "evk_peer_open": [P, P],
    "evk_peer_close": [P],
    "evk_peer_barrier": [P, P, I, I, P, P, L, P],
    "evk_mpce_shard_finish": [P, I, L, L, F, D, P, P, P, L, I, P, P, P, P],
    "evk_mpce_bwd_gemm_scatter": [P, P, L, L, L, P, P, L, L, F, I, P, I, L, L, I, I, I, P],
    "evk_l2norm_bwd_parts": [P, I, L, L, L, L, P, P, P, I, L, I, L, P, F, P, I, L, I, P, P, P],
    "evk_tc_gemm_nt": [P, P, L, P, P, L, L, L, L, P, L, P],
    "evk_topk_update": [P, L, L, L, L, P, P, I, P, P, I, P],
    "evk_tc_gemm_probe": [P, L, I, P, L, I, L, L, L, P, L, I, I, P],
}

# host helpers that return a size instead of a status code
INT64_RESULT = {"evk_stats_workspace_bytes", "evk_shard_finish_workspace_bytes", "evk_posmask_ld_words",
                "evk_mpce_rowpart_rows", "evk_mpce_colpart_rows", "evk_mpce_strip_ld"}

_lib = None


class PeerSync(ctypes.Structure):
    """evk_peer_sync_t of include/evoke_b200.h (host struct, passed by pointer)."""
    _fields_ = [("flag_ptrs", ctypes.c_uint64 * 16), ("err_ptrs", ctypes.c_uint64 * 16), ("err_host", c_void_p),
                ("step", c_void_p), ("n_ranks", c_int), ("rank", c_int), ("index", c_int), ("per_step", c_int),
                ("timeout_ms", c_int64)]


class EvokeLibraryError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EvokeLibraryError(
            f"{LIB_PATH} is missing: build it with `python -m evoke_b200.build` "
            "(or __graft_entry__.build()).  evoke_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = ABI mismatch, fail loudly
        fn.argtypes = argtypes
        fn.restype = c_char_p if name == "evk_last_error" else (c_int64 if name in INT64_RESULT else c_int)
    ver = lib.evk_version()
    if ver != ABI_VERSION:
        raise EvokeLibraryError(f"libevoke_b200.so ABI version {ver} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def last_error() -> str:
    return (load().evk_last_error() or b"").decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc == EVK_OK:
        return
    msg = f"{what}: {last_error()} (code {rc})"
    if rc == EVK_ERR_INVALID:
        raise ValueError(msg)
    raise EvokeLibraryError(msg)


# kernels of this library launched by one call of each entry point (cudaMemsetAsync not counted)
KERNELS_PER_CALL = {
    "evk_l2norm_fwd": 1, "evk_l2norm_bwd": 1, "evk_posmask_build": 1, "evk_mpce_small_fwd": 1,
    "evk_mpce_small_bwd": 1, "evk_reduce_partials": 1, "evk_mpce_finalize": 1, "evk_mpce_stats_fused": 1, "evk_mpce_pos": 1, "evk_mpce_fwd": 1,
    "evk_mpce_bwd_w": 1, "evk_mpce_bwd_gemm": 1, "evk_tc_gemm_probe": 1,
    "evk_mpce_fwd_store": 1, "evk_mpce_w_from_e": 3, "evk_mpce_pos_logits": 1, "evk_mpce_pos_from_lists": 1, "evk_tc_gemm_nt": 1, "evk_topk_update": 1,
    "evk_peer_barrier": 1, "evk_mpce_small_fwd_batched": 1, "evk_mpce_small_bwd_batched": 1,
    "evk_local_attend_fwd": 1, "evk_local_attend_bwd": 2, "evk_token_sim_fwd": 1, "evk_token_sim_bwd": 1, "evk_peer_push_shard": 1, "evk_peer_wait_landed": 1, "evk_mpce_fwd_store_gathered": 1, "evk_mpce_finalize_avgpos": 1, "evk_shard_prologue": 1, "evk_mpce_shard_stats_push": 1, "evk_l2norm_bwd_parts": 1, "evk_mpce_shard_finish": 1, "evk_mpce_bwd_gemm_scatter": 1,
}
launch_count = 0          # running total, read by bench.py ("gpu_launches")
call_hook = None          # optional callable(name, phase) with phase in {"before", "after"} (bench.py timing)


def size(name: str, *args) -> int:
    """A buffer-size helper of the C ABI (evk_*_bytes / _ld_* / _rows): the library is the single source of the layouts."""
    v = int(getattr(load(), name)(*args))
    if v <= 0:
        raise ValueError(f"{name}{args}: bad arguments")
    return v


def call(name: str, *args) -> None:
    global launch_count
    hook = call_hook
    if hook is not None:
        hook(name, "before")
    check(getattr(load(), name)(*args), name)
    launch_count += KERNELS_PER_CALL.get(name, 0)
    if hook is not None:
        hook(name, "after")
