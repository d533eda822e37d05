"""Host orchestration of the contrastive hot path: torch.autograd Functions that enqueue the
sm_100a kernels of libevoke_b200.so on the current CUDA stream.

PyTorch is used for device memory (the caching allocator owns every buffer, the library
allocates nothing), streams and autograd plumbing only; every arithmetic step of the loss
runs in the library.  There is no CPU path and no PyTorch fallback.

Math (oracle/evoke_oracle.py has the fp64 restatement; reference file:line in the docstrings
of evoke_b200/loss.py):
    S = Qhat Khat^T / tau,   E = exp(S - 1/tau)          (|S| <= 1/tau for unit rows)
    R_i = sum_j E_ij,  C_j = sum_i E_ij,  pos_i = sum_j M_ij S_ij
    G   loss = 1/(2N) [ sum_i (1/tau + ln R_i) + sum_j (1/tau + ln C_j) - 2 sum_i pos_i/c_i ]
    MPC loss = 1/M'   [ sum_i (1/tau + ln R_i - pos_i/c_i) ]          (diagonal excluded)
    W_ij = E_ij (1/R_i + 1/C_j) - 2 M_ij / c_i        (MPC: C := R)
    dQhat = W Khat / (2 N tau),  dKhat = W^T Qhat / (2 N tau)         (MPC: dXhat = W Xhat / (M' tau))
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from ._lib import (DTYPE_BF16, DTYPE_F16, DTYPE_F32, FLAG_AVGPOS, FLAG_EXCLUDE_DIAG, FLAG_NO_COLSUM, FLAG_NO_POS,
                   FLAG_SPLIT_BF16)
from .ids import DeviceIds

SMALL_PATH_MAX = int(os.environ.get("EVOKE_B200_SMALL_MAX", "512"))   # rows/cols up to which the SIMT path is used
TILE_M, TILE_N = 128, 256
ROW_PARTS = 2          # row-statistic partials per 256-column tile (epilogue warps sharing a row); see _row_parts()
OVERLAP_STREAMS = os.environ.get("EVOKE_B200_OVERLAP", "0") == "1"   # side-stream overlap outside graph capture too
# bf16 mode: K3 stores E = exp(S - 1/tau) as a bf16 row strip and the backward turns it into W in place
# (HBM-bound K4t) instead of recomputing the similarity tiles (K4a): 6 N^2 D executed FLOP instead of 8.
E_STRIP = os.environ.get("EVOKE_B200_ESTRIP", "1") == "1"
# ... and then nothing of size N^2 is derived from the ids either: K2 builds only the per-row positive lists, the
# positive-logit sums and the exact W entries come from those (rows with more positives than list slots: id scan).
# EVOKE_B200_MASK_FREE=0 keeps the dense bit mask in the K3 epilogue (round-1 behaviour, for A/B measurements).
MASK_FREE = os.environ.get("EVOKE_B200_MASK_FREE", "1") == "1"


def _row_parts() -> int:
    """Row-partial rows K3 writes per 256-column tile (asked from the library: it depends on the kernel variant)."""
    global ROW_PARTS
    ROW_PARTS = int(_lib.load().evk_mpce_row_parts())
    return ROW_PARTS


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return DTYPE_F32
    if t.dtype == torch.bfloat16:
        return DTYPE_BF16
    if t.dtype == torch.float16:
        return DTYPE_F16
    raise TypeError(f"embeddings must be float32, bfloat16 or float16, got {t.dtype}")


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"evoke_b200: `{name}` lives on {t.device}; this library only runs on a CUDA "
                           "(sm_100a) device and has no CPU path")
    if t.dim() != 2:
        raise ValueError(f"evoke_b200: `{name}` must be 2-D [N, D], got shape {tuple(t.shape)}")


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


_SIDE_STREAMS: dict = {}


def _side_stream(device: torch.device, which: int = 0) -> torch.cuda.Stream:
    """One auxiliary stream per device for work that runs next to the tensor-bound kernels
    (mask builder, positive sums, zero fills, the second gradient contraction).  It is always
    forked from and joined back into the caller's stream, so the sequence stays capturable."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get((idx, which))
    if st is None:
        st = torch.cuda.Stream(device=idx)
        _SIDE_STREAMS[(idx, which)] = st
    return st


def _shared_with(stream: torch.cuda.Stream, *tensors):
    """Tell the caching allocator that `tensors` are also used on `stream`."""
    for t in tensors:
        if t is not None:
            t.record_stream(stream)


# ------------------------------------------------------------------------------------- kernels
@dataclass
class Normalized:
    """Output of K1 for one embedding matrix."""
    n: int
    d: int
    norm: torch.Tensor                      # [n] fp32, unclamped
    f32: Optional[torch.Tensor] = None      # [n, d] fp32 (small path)
    hi: Optional[torch.Tensor] = None       # [n, ld] bf16 (tc path)
    lo: Optional[torch.Tensor] = None       # [n, ld] bf16 (tc path, fp32-parity mode)
    ld: int = 0


def l2norm_fwd(x: torch.Tensor, *, want_f32: bool, want_hi: bool, want_lo: bool,
               gather: Optional[torch.Tensor] = None, out: Optional[Normalized] = None) -> Normalized:
    """K1: xhat = x / max(||x||, 1e-12) (F.normalize, v0520.py:495-496, :436), honouring the
    strides of ``x`` and an optional row gather.  ``out``: write into the buffers of an earlier result of the same
    shape (the static operands of a captured graph) instead of allocating."""
    n = int(gather.shape[0]) if gather is not None else int(x.shape[0])
    d = int(x.shape[1])
    dev = x.device
    if out is not None:
        if out.n != n or out.d != d:
            raise ValueError(f"l2norm_fwd: out holds {out.n} x {out.d}, input is {n} x {d}")
    else:
        out = Normalized(n=n, d=d, norm=torch.empty(n, dtype=torch.float32, device=dev))
        if want_f32:
            out.f32 = torch.empty((n, d), dtype=torch.float32, device=dev)
        if want_hi:
            out.ld = _round_up(d, 8)
            out.hi = torch.empty((n, out.ld), dtype=torch.bfloat16, device=dev)
            if want_lo:
                out.lo = torch.empty((n, out.ld), dtype=torch.bfloat16, device=dev)
    _lib.call("evk_l2norm_fwd", _ptr(x), _dtype_code(x), n, d, x.stride(0), x.stride(1), _ptr(gather),
              _ptr(out.f32), d, _ptr(out.hi), _ptr(out.lo), out.ld, _ptr(out.norm), _stream())
    return out


def l2norm_bwd(x: torch.Tensor, nrm: Normalized, g_hat: torch.Tensor, *, scale_dev: Optional[torch.Tensor],
               scale_host: float, gather: Optional[torch.Tensor] = None, parts=None,
               error: Optional[torch.Tensor] = None, sync=None) -> torch.Tensor:
    """Backward of K1 fused with the loss scale; returns dx with x's shape and dtype.
    parts = (n_parts, stride_in_elements): g_hat is the first of n_parts partial buffers to be summed.
    error: the sharded transport's failure flag (device int32): non-zero -> the gradients are NaN.
    sync: ctypes pointer to an evk_peer_sync_t - the kernel first waits for the peers (sharded path)."""
    if gather is not None:
        dx = torch.zeros(x.shape, dtype=x.dtype, device=x.device)       # filtered rows get zero gradient
    else:
        dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
    n_parts, part_stride = parts if parts is not None else (1, 0)
    _lib.call("evk_l2norm_bwd_parts", _ptr(x), _dtype_code(x), nrm.n, nrm.d, x.stride(0), x.stride(1), _ptr(gather),
              _ptr(nrm.norm), _ptr(g_hat), _dtype_code(g_hat), g_hat.stride(0), int(n_parts), int(part_stride), _ptr(scale_dev),
              float(scale_host), _ptr(dx), _dtype_code(dx), dx.stride(0), 0, _ptr(error), sync, _stream())
    return dx


POS_SLOTS = 8          # listed positives per row (rows with more fall back to scanning the mask)


def posmask_build(rows: DeviceIds, cols: DeviceIds, *, clear_diag: bool, diag_offset: int = 0, want_list: bool = False,
                  want_bits: bool = True, counts: Optional[torch.Tensor] = None, sync=None):
    """K2: bit-packed positive mask [n_rows, ld_words] (uint32 stored as int32) + counts[n_rows]
    (+ pos_idx [n_rows, POS_SLOTS], the first positives of every row, with want_list).
    want_bits=False (needs want_list): counts and lists only, bits is None.
    counts: a buffer the caller has already zeroed (sharded path: the prologue kernel does); sync: evk_peer_sync_t."""
    n_rows, n_cols = len(rows), len(cols)
    dev = rows.key.device
    ld_words = _lib.size("evk_posmask_ld_words", n_cols)                 # whole 256-column tiles
    bits = torch.empty((n_rows, ld_words), dtype=torch.int32, device=dev) if want_bits else None
    zeroed = counts is not None
    if counts is None:
        counts = torch.empty(n_rows, dtype=torch.int32, device=dev)
    pos_idx = torch.empty((n_rows, POS_SLOTS), dtype=torch.int32, device=dev) if want_list else None
    _lib.call("evk_posmask_build", _ptr(rows.key), _ptr(rows.key2), n_rows, _ptr(cols.key), _ptr(cols.key2),
              n_cols, diag_offset, int(clear_diag), _ptr(bits), ld_words, _ptr(counts), _ptr(pos_idx), POS_SLOTS,
              int(zeroed), sync, _stream())
    return (bits, counts, pos_idx) if want_list else (bits, counts)


def pos_from_lists(q: Normalized, k: Normalized, rows: DeviceIds, cols: DeviceIds, counts: torch.Tensor,
                   pos_dot: torch.Tensor, inv_tau: float, *, clear_diag: bool, diag_offset: int = 0) -> torch.Tensor:
    """row_pos[i] = sum_j M_ij S_ij from the K2 lists (id scan for rows with more positives than slots): the
    mask-free replacement of K3's positive sums."""
    row_pos = torch.empty(q.n, dtype=torch.float32, device=q.hi.device)
    _lib.call("evk_mpce_pos_from_lists", _ptr(q.hi), q.ld, _ptr(k.hi), k.ld, q.n, k.n, q.d, _ptr(rows.key), _ptr(rows.key2),
              _ptr(cols.key), _ptr(cols.key2), int(diag_offset), int(clear_diag), _ptr(counts), _ptr(pos_dot), POS_SLOTS,
              float(inv_tau), _ptr(row_pos), _stream())
    return row_pos


def pos_logits(q: Normalized, k: Normalized, pos_idx: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """pos_dot[i, s] = qhat_i . khat_{pos_idx[i, s]} for the listed positives (side-stream work next to K3)."""
    pos_dot = torch.empty((q.n, POS_SLOTS), dtype=torch.float32, device=q.hi.device)
    _lib.call("evk_mpce_pos_logits", _ptr(q.hi), q.ld, _ptr(k.hi), k.ld, q.n, q.d, _ptr(pos_idx), _ptr(counts),
              POS_SLOTS, _ptr(pos_dot), _stream())
    return pos_dot


def reduce_partials(part: torch.Tensor, parts: int, n: int, out: Optional[torch.Tensor] = None,
                    divisor: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[j] = sum_p part[p, j] (/ divisor[j] if given); `out` may be a slice of a larger buffer."""
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=part.device)
    _lib.call("evk_reduce_partials", _ptr(part), parts, part.stride(0), n, _ptr(divisor), _ptr(out), _stream())
    return out


def finalize(row_sum, row_pos, counts, col_sum, *, col_lo: int, col_hi: int, shift: float, pos_weight: float,
             inv_count: float, want_b: bool = True):
    n_rows = int(row_sum.shape[0])
    dev = row_sum.device
    a_row = torch.empty(n_rows, dtype=torch.float32, device=dev)
    n_cols = 0 if col_sum is None else int(col_sum.shape[0])
    b_col = torch.empty(n_cols, dtype=torch.float32, device=dev) if (col_sum is not None and want_b) else None
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    _lib.call("evk_mpce_finalize", _ptr(row_sum), _ptr(row_pos), _ptr(counts), n_rows, _ptr(col_sum), n_cols,
              col_lo, col_hi, float(shift), float(pos_weight), float(inv_count), _ptr(a_row), _ptr(b_col),
              _ptr(loss), _stream())
    return a_row, b_col, loss


# --- small path ----------------------------------------------------------------------------
def small_fwd(q: Normalized, k: Normalized, bits, inv_tau: float, flags: int, diag_offset: int = 0):
    dev = q.f32.device
    row_sum = torch.empty(q.n, dtype=torch.float32, device=dev)
    row_pos = torch.empty(q.n, dtype=torch.float32, device=dev)
    _lib.call("evk_mpce_small_fwd", _ptr(q.f32), q.f32.stride(0), _ptr(k.f32), k.f32.stride(0), q.n, k.n, q.d,
              _ptr(bits), bits.stride(0), float(inv_tau), flags, diag_offset, _ptr(row_sum), _ptr(row_pos), _stream())
    return row_sum, row_pos


def small_bwd(q: Normalized, k: Normalized, bits, counts, a_row, b_col, inv_tau: float, flags: int,
              diag_offset: int = 0, pos_row=None, pos_col=None) -> torch.Tensor:
    dq = torch.empty((q.n, q.d), dtype=torch.float32, device=q.f32.device)
    _lib.call("evk_mpce_small_bwd", _ptr(q.f32), q.f32.stride(0), _ptr(k.f32), k.f32.stride(0), q.n, k.n, q.d,
              _ptr(bits), bits.stride(0), _ptr(counts), _ptr(a_row), _ptr(b_col), float(inv_tau), flags,
              diag_offset, _ptr(dq), dq.stride(0), _ptr(pos_row), _ptr(pos_col), _stream())
    return dq


def finalize_avgpos(row_neg, row_pos, counts, *, shift: float, inv_count: float, loss: Optional[torch.Tensor] = None):
    """One direction of the averaged-positive rule -> (a_row, pos_row, loss[1]); ``loss`` given: accumulate into it."""
    n = int(row_neg.shape[0])
    dev = row_neg.device
    a_row = torch.empty(n, dtype=torch.float32, device=dev)
    pos_row = torch.empty(n, dtype=torch.float32, device=dev)
    acc = loss is not None
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=dev)
    _lib.call("evk_mpce_finalize_avgpos", _ptr(row_neg), _ptr(row_pos), _ptr(counts), n, float(shift), float(inv_count),
              _ptr(a_row), _ptr(pos_row), _ptr(loss), int(acc), _stream())
    return a_row, pos_row, loss


# --- tcgen05 path --------------------------------------------------------------------------
def alloc_partials(n_rows: int, n_cols: int, dev, want_pos: bool = True, want_col: bool = True):
    """K3's per-tile partial statistics (rs_part, rp_part | None, cs_part | None); the library states their row counts."""
    rows = _lib.size("evk_mpce_rowpart_rows", n_cols)
    rs_part = torch.empty((rows, n_rows), dtype=torch.float32, device=dev)
    rp_part = torch.empty((rows, n_rows), dtype=torch.float32, device=dev) if want_pos else None
    cs_part = torch.empty((_lib.size("evk_mpce_colpart_rows", n_rows), n_cols), dtype=torch.float32,
                          device=dev) if want_col else None
    return rs_part, rp_part, cs_part


def tc_fwd_partials(q: Normalized, k: Normalized, bits, inv_tau: float, flags: int, diag_offset: int = 0):
    """K3.  Returns the per-tile partials (rs_part, rp_part, cs_part | None) and their counts."""
    dev = q.hi.device
    want_col = not (flags & FLAG_NO_COLSUM)
    want_pos = not (flags & FLAG_NO_POS)
    rs_part, rp_part, cs_part = alloc_partials(q.n, k.n, dev, want_pos, want_col)
    _lib.call("evk_mpce_fwd", _ptr(q.hi), _ptr(q.lo), q.ld, _ptr(k.hi), _ptr(k.lo), k.ld, q.n, k.n, q.d,
              _ptr(bits) if want_pos else None, bits.stride(0) if want_pos else 0, float(inv_tau), flags, diag_offset,
              _ptr(rs_part), _ptr(rp_part), q.n, _ptr(cs_part), k.n, _stream())
    return rs_part, rp_part, cs_part


def tc_fwd_store(q: Normalized, k: Normalized, bits, inv_tau: float, flags: int, diag_offset: int = 0):
    """K3 that also writes the bf16 E strip [q.n, ld_e].  Returns (rs_part, rp_part, cs_part | None, e, ld_e)."""
    dev = q.hi.device
    want_col = not (flags & FLAG_NO_COLSUM)
    want_pos = not (flags & FLAG_NO_POS)
    rs_part, rp_part, cs_part = alloc_partials(q.n, k.n, dev, want_pos, want_col)
    ld_e = _lib.size("evk_mpce_strip_ld", k.n)
    e = torch.empty((q.n, ld_e), dtype=torch.bfloat16, device=dev)
    _lib.call("evk_mpce_fwd_store", _ptr(q.hi), q.ld, _ptr(k.hi), k.ld, q.n, k.n, q.d,
              _ptr(bits) if want_pos else None, bits.stride(0) if want_pos else 0, float(inv_tau), flags, int(diag_offset),
              _ptr(rs_part), _ptr(rp_part), q.n, _ptr(cs_part), k.n, _ptr(e), ld_e, _stream())
    return rs_part, rp_part, cs_part, e, ld_e


def tc_w_from_e(e: torch.Tensor, ld_e: int, n_cols: int, bits, counts, a_row, b_col, q: Optional[Normalized],
                k: Optional[Normalized], inv_tau: float, row0: int = 0, rows: Optional[int] = None,
                pos=None, ids=None, clear_diag: bool = False, diag_offset: int = 0) -> None:
    """K4t: rows [row0, row0+rows) of the E strip become W, in place.  q / k: the forward's operands, from
    which the positive entries are recomputed in fp32 (None: every entry from the strip).
    pos = (pos_idx, pos_dot): the forward's positive lists - then no mask scan / dot products are needed.
    bits = None (mask-free mode): ids = (row DeviceIds, column DeviceIds) serve the rows with more positives than slots."""
    rows = int(e.shape[0]) - row0 if rows is None else rows
    exact = q is not None and k is not None
    rid, cid = ids if ids is not None else (None, None)
    _lib.call("evk_mpce_w_from_e", _ptr(e[row0:]), ld_e, rows, n_cols, None if bits is None else _ptr(bits[row0:]),
              0 if bits is None else bits.stride(0),
              _ptr(counts[row0:]), _ptr(a_row[row0:]), _ptr(b_col),
              _ptr(q.hi[row0:]) if exact else None, q.ld if exact else 0, _ptr(k.hi) if exact else None,
              k.ld if exact else 0, q.d if exact else 0, float(inv_tau),
              _ptr(pos[0][row0:]) if pos is not None else None, _ptr(pos[1][row0:]) if pos is not None else None,
              POS_SLOTS, None if rid is None else _ptr(rid.key[row0:]),
              None if rid is None or rid.key2 is None else _ptr(rid.key2[row0:]),
              None if cid is None else _ptr(cid.key), None if cid is None or cid.key2 is None else _ptr(cid.key2),
              int(diag_offset) + row0, int(clear_diag), _stream())


def rows_of(x: Normalized, r0: int, r1: int) -> Normalized:
    """Row range of a normalised matrix (views, no copy)."""
    return Normalized(n=r1 - r0, d=x.d, norm=None if x.norm is None else x.norm[r0:r1],
                      f32=None if x.f32 is None else x.f32[r0:r1], hi=None if x.hi is None else x.hi[r0:r1],
                      lo=None if x.lo is None else x.lo[r0:r1], ld=x.ld)


def tc_pos(q: Normalized, k: Normalized, bits, inv_tau: float) -> torch.Tensor:
    """row_pos[i] = sum_j M_ij S_ij from the bit mask (for K3 launched with FLAG_NO_POS)."""
    row_pos = torch.empty(q.n, dtype=torch.float32, device=q.hi.device)
    _lib.call("evk_mpce_pos", _ptr(q.hi), _ptr(q.lo), q.ld, _ptr(k.hi), _ptr(k.lo), k.ld, q.n, k.n, q.d,
              _ptr(bits), bits.stride(0), float(inv_tau), _ptr(row_pos), _stream())
    return row_pos


def stats_fused(rs_part, rp_part, cs_part, counts, *, shift: float, pos_weight: float, inv_count: float,
                col_lo: int = 0, col_hi: Optional[int] = None):
    # counts=None: rp_part already holds pos_i / c_i
    """Partials -> (a_row, b_col | None, loss[1]) in one launch (single-GPU form).  rp_part is
    either K3's [parts, n] partials or the [n] vector written by tc_pos."""
    dev = rs_part.device
    if rs_part.dim() == 1:
        rs_part = rs_part.unsqueeze(0)
    if rp_part.dim() == 1:
        rp_part = rp_part.unsqueeze(0)
    if cs_part is not None and cs_part.dim() == 1:
        cs_part = cs_part.unsqueeze(0)
    n_rows = int(rs_part.shape[1])
    n_cols = 0 if cs_part is None else int(cs_part.shape[1])
    a_row = torch.empty(n_rows, dtype=torch.float32, device=dev)
    b_col = torch.empty(n_cols, dtype=torch.float32, device=dev) if cs_part is not None else None
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    ws = torch.empty(_lib.size("evk_stats_workspace_bytes", n_rows, n_cols), dtype=torch.uint8, device=dev)
    _lib.call("evk_mpce_stats_fused", _ptr(rs_part), int(rs_part.shape[0]), max(int(rs_part.stride(0)), n_rows),
              _ptr(rp_part), int(rp_part.shape[0]), max(int(rp_part.stride(0)), n_rows),
              _ptr(counts), n_rows, _ptr(cs_part), 0 if cs_part is None else int(cs_part.shape[0]),
              0 if cs_part is None else max(int(cs_part.stride(0)), n_cols), n_cols,
              int(col_lo), int(n_cols if col_hi is None else col_hi), float(shift), float(pos_weight), float(inv_count),
              _ptr(a_row), _ptr(b_col), _ptr(loss), _ptr(ws), ws.numel(), _stream())
    return a_row, b_col, loss


def tc_fwd(q: Normalized, k: Normalized, bits, inv_tau: float, flags: int, diag_offset: int = 0):
    """K3 + partial reduction (sharded path).  Returns (row_sum, row_pos, col_sum | None)."""
    dev = q.hi.device
    want_col = not (flags & FLAG_NO_COLSUM)
    rs_part, rp_part, cs_part = alloc_partials(q.n, k.n, dev, True, want_col)
    _lib.call("evk_mpce_fwd", _ptr(q.hi), _ptr(q.lo), q.ld, _ptr(k.hi), _ptr(k.lo), k.ld, q.n, k.n, q.d,
              _ptr(bits), bits.stride(0), float(inv_tau), flags, diag_offset,
              _ptr(rs_part), _ptr(rp_part), q.n, _ptr(cs_part), k.n, _stream())
    row_sum = reduce_partials(rs_part, int(rs_part.shape[0]), q.n)
    row_pos = reduce_partials(rp_part, int(rp_part.shape[0]), q.n)
    col_sum = reduce_partials(cs_part, int(cs_part.shape[0]), k.n) if want_col else None
    return row_sum, row_pos, col_sum


def tc_bwd_w(q: Normalized, k: Normalized, bits, counts, a_row, b_col, inv_tau: float, flags: int,
             diag_offset: int = 0):
    """K4a.  Returns the bf16 W strip (hi, lo|None, ld_w)."""
    dev = q.hi.device
    ld_w = _lib.size("evk_mpce_strip_ld", k.n)
    w_hi = torch.empty((q.n, ld_w), dtype=torch.bfloat16, device=dev)
    w_lo = torch.empty((q.n, ld_w), dtype=torch.bfloat16, device=dev) if (flags & FLAG_SPLIT_BF16) else None
    _lib.call("evk_mpce_bwd_w", _ptr(q.hi), _ptr(q.lo), q.ld, _ptr(k.hi), _ptr(k.lo), k.ld, q.n, k.n, q.d,
              _ptr(bits), bits.stride(0), _ptr(counts), _ptr(a_row), _ptr(b_col), float(inv_tau), flags,
              diag_offset, _ptr(w_hi), _ptr(w_lo), ld_w, _stream())
    return w_hi, w_lo, ld_w


def tc_bwd_gemm(w_hi, w_lo, ld_w: int, n_rows: int, n_cols: int, transpose_w: bool, x: Normalized, flags: int,
                out: Optional[torch.Tensor] = None, cta_limit: int = 0) -> torch.Tensor:
    """K4b: out (+)= W x  (or W^T x).  ``out`` fp32 [rows_out, d]; zero-initialised here if None.
    cta_limit > 0: use at most that many SMs (a second contraction runs beside this one on another stream)."""
    rows_out = n_cols if transpose_w else n_rows
    if out is None:
        out = torch.zeros((rows_out, _round_up(x.d, 4)), dtype=torch.float32, device=w_hi.device)
    _lib.call("evk_mpce_bwd_gemm", _ptr(w_hi), _ptr(w_lo), ld_w, n_rows, n_cols, int(transpose_w),
              _ptr(x.hi), _ptr(x.lo), x.ld, x.d, 1.0, flags & FLAG_SPLIT_BF16, _ptr(out), out.stride(0), int(cta_limit),
              _stream())
    return out


def tc_gemm_probe(a: torch.Tensor, b: torch.Tensor, a_major: int, b_major: int, m: int, n: int, k: int,
                  variant: int = 0, splits: int = 0) -> torch.Tensor:
    """Bring-up helper: C[m,n] = A B^T through the tcgen05 main loop (see the header)."""
    c = torch.zeros((m, _round_up(n, 4)), dtype=torch.float32, device=a.device)
    _lib.call("evk_tc_gemm_probe", _ptr(a), a.stride(0), a_major, _ptr(b), b.stride(0), b_major, m, n, k,
              _ptr(c), c.stride(0), variant, splits, _stream())
    return c[:, :n]


# ------------------------------------------------------------------------------------- autograd
@dataclass
class LossConfig:
    kind: str                       # "G" | "MPC"
    inv_tau: float
    precision: str                  # "fp32" | "bf16"
    path: str                       # "small" | "tc"
    row_ids: DeviceIds              # keys of the rows that take part (already truncated / filtered)
    gather: Optional[torch.Tensor] = None   # MPC: int32 indices of the kept rows
    n_keep: int = 0                 # AMPC: number of multi-view rows (the loss is their mean, :707)


def choose_path(path: str, n_rows: int, n_cols: int, d: int) -> str:
    if path in ("small", "tc"):
        return path
    if path != "auto":
        raise ValueError(f"path must be 'auto', 'small' or 'tc', got {path!r}")
    return "small" if (max(n_rows, n_cols) <= SMALL_PATH_MAX and d <= 4096) else "tc"


class _State:
    """What the backward needs from the forward (the role autograd's ctx plays; a plain object so that the same two
    functions can also be captured into CUDA graphs, evoke_b200/graphs.py)."""
    pass


def normalize_pair(cfg: LossConfig, image: torch.Tensor, text: Optional[torch.Tensor], out=None):
    """K1 of both sides for a loss configuration -> (qn, kn); ``out`` = an earlier result to overwrite."""
    small = cfg.path == "small"
    split = (not small) and cfg.precision == "fp32"
    kw = dict(want_f32=small, want_hi=not small, want_lo=split)
    qn = l2norm_fwd(image, gather=cfg.gather, out=None if out is None else out[0], **kw)
    kn = qn if cfg.kind == "MPC" else l2norm_fwd(text, out=None if out is None else out[1], **kw)
    return qn, kn


def mpce_forward(cfg: LossConfig, image: Optional[torch.Tensor], text: Optional[torch.Tensor], need_grad=(True, True),
                 pre=None):
    """Forward kernel sequence of the G / MPC loss.  Returns (loss [1] fp32, state).  need_grad = (image, text).
    pre = (qn, kn): the operands are already normalised (K1 ran outside: the captured-graph form)."""
    st = _State()
    st.e_strip = None
    st.pos = None
    small = cfg.path == "small"
    split = (not small) and cfg.precision == "fp32"
    flags = FLAG_SPLIT_BF16 if split else 0
    kw = dict(want_f32=small, want_hi=not small, want_lo=split)
    mpc = cfg.kind == "MPC"
    if mpc:
        flags |= FLAG_EXCLUDE_DIAG | FLAG_NO_COLSUM
    if small:
        qn, kn = pre if pre is not None else normalize_pair(cfg, image, text)
        n = qn.n
        pos_weight, inv_count = (1.0, 1.0 / n) if mpc else (2.0, 0.5 / n)
        bits, counts = posmask_build(cfg.row_ids, cfg.row_ids, clear_diag=mpc)
        row_sum, row_pos = small_fwd(qn, kn, bits, cfg.inv_tau, flags)
        col_sum = None if mpc else small_fwd(kn, qn, bits, cfg.inv_tau, flags)[0]
        a_row, b_col, loss = finalize(row_sum, row_pos, counts, col_sum, col_lo=0, col_hi=0 if mpc else n,
                                      shift=cfg.inv_tau, pos_weight=pos_weight, inv_count=inv_count)
    else:
        # K2 (integer-ALU bound) runs on a side stream next to the two HBM-bound K1 launches
        # when the sequence is being captured into a CUDA graph.  The positive-logit sums stay
        # in the K3 epilogue: ln R_i and pos_i must come from the same tensor-core accumulators
        # for their rounding to cancel in the loss (cold temperatures).
        overlap = OVERLAP_STREAMS or torch.cuda.is_current_stream_capturing()
        use_strip = E_STRIP and not split and any(need_grad)
        pos_idx = pos_dot = None

        mask_free = use_strip and MASK_FREE

        def build_mask():
            if use_strip:
                return posmask_build(cfg.row_ids, cfg.row_ids, clear_diag=mpc, want_list=True, want_bits=not mask_free)
            return posmask_build(cfg.row_ids, cfg.row_ids, clear_diag=mpc) + (None,)

        if overlap:
            main = torch.cuda.current_stream()
            side = _side_stream(image.device if image is not None else pre[0].norm.device)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                bits, counts, pos_idx = build_mask()
        else:
            bits, counts, pos_idx = build_mask()
        qn, kn = pre if pre is not None else normalize_pair(cfg, image, text)
        n = qn.n
        pos_weight, inv_count = (1.0, 1.0 / n) if mpc else (2.0, 0.5 / n)
        if overlap:
            main.wait_stream(side)
            _shared_with(main, bits, counts, pos_idx)
        if use_strip:
            # exact logits of the listed positives (O(N*D)), next to K3 on the side stream; the backward's
            # K4t needs them for the entries where softmax and target cancel
            def positives():
                pd = pos_logits(qn, kn, pos_idx, counts)
                rp = pos_from_lists(qn, kn, cfg.row_ids, cfg.row_ids, counts, pd, cfg.inv_tau,
                                    clear_diag=mpc) if mask_free else None
                return pd, rp

            if overlap:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    pos_dot, row_pos_l = positives()
                _shared_with(side, qn.hi, kn.hi, pos_idx, counts)
            else:
                pos_dot, row_pos_l = positives()
            rs_part, rp_part, cs_part, e_strip, ld_e = tc_fwd_store(qn, kn, bits, cfg.inv_tau,
                                                                    flags | (FLAG_NO_POS if mask_free else 0))
            if mask_free:
                if overlap:                     # the statistics need the positive sums computed next to K3
                    main.wait_stream(side)
                    _shared_with(main, pos_dot, row_pos_l)
                rp_part = row_pos_l
            st.e_strip = (e_strip, ld_e)
            st.pos = (pos_idx, pos_dot)
        else:
            rs_part, rp_part, cs_part = tc_fwd_partials(qn, kn, bits, cfg.inv_tau, flags)
        a_row, b_col, loss = stats_fused(rs_part, rp_part, cs_part, counts, shift=cfg.inv_tau,
                                         pos_weight=pos_weight, inv_count=inv_count)
        if use_strip and overlap:
            main.wait_stream(side)
            _shared_with(main, pos_dot)
    if mpc:
        b_col = a_row
    st.cfg, st.flags, st.qn, st.kn = cfg, flags, qn, kn
    st.aux = (bits, counts, a_row, b_col)
    st.image, st.text = image, text
    st.need_grad = (bool(need_grad[0]), bool(need_grad[1]) if (text is not None or (pre is not None and not mpc)) else False)
    return loss, st


def mpce_scale(st: _State) -> float:
    """Host part of the gradient scale: 1/(M' tau) for MPC, 1/(2 N tau) for G (the device part is the upstream gradient)."""
    return st.cfg.inv_tau / st.qn.n if st.cfg.kind == "MPC" else 0.5 * st.cfg.inv_tau / st.qn.n


def mpce_finish(st: _State, image: torch.Tensor, text: Optional[torch.Tensor], dq, dk, g: torch.Tensor):
    """K1b of both sides: (dQhat, dKhat) of mpce_backward(finish=False) -> gradients of the caller's tensors."""
    scale = mpce_scale(st)
    d_image = None if dq is None else l2norm_bwd(image, st.qn, dq, scale_dev=g, scale_host=scale, gather=st.cfg.gather)
    d_text = None if dk is None else l2norm_bwd(text, st.kn, dk, scale_dev=g, scale_host=scale)
    return d_image, d_text


def mpce_backward(st: _State, g: torch.Tensor, finish: bool = True):
    """Backward kernel sequence; g = upstream gradient, fp32 [1] on the device.  Returns (d_image, d_text), or with
    finish=False the gradients (dQhat, dKhat) of the NORMALISED operands (fp32 [n, ld]): K1b then runs outside, on
    the caller's tensors (the captured-graph form, mpce_finish)."""
    cfg, flags, qn, kn = st.cfg, st.flags, st.qn, st.kn
    bits, counts, a_row, b_col = st.aux
    image, text = st.image, st.text
    mpc = cfg.kind == "MPC"
    n = qn.n
    scale = cfg.inv_tau / n if mpc else 0.5 * cfg.inv_tau / n
    need_q = st.need_grad[0]
    need_k = (not mpc) and st.need_grad[1]
    d_image = d_text = dq = dk = None
    if cfg.path == "small":
        if need_q:
            dq = small_bwd(qn, kn, bits, counts, a_row, b_col, cfg.inv_tau, flags)
            if finish:
                d_image = l2norm_bwd(image, qn, dq, scale_dev=g, scale_host=scale, gather=cfg.gather)
        if need_k:
            dk = small_bwd(kn, qn, bits, counts, b_col, a_row, cfg.inv_tau, flags)   # M is symmetric
            if finish:
                d_text = l2norm_bwd(text, kn, dk, scale_dev=g, scale_host=scale)
    elif need_q or need_k:
        dev = qn.norm.device
        width = _round_up(qn.d, 4)
        overlap = OVERLAP_STREAMS or torch.cuda.is_current_stream_capturing()
        strip = st.e_strip
        if strip is not None:
            if strip[0] is None:
                raise RuntimeError("evoke_b200: backward called twice: the E strip saved by the forward is turned "
                                   "into W in place (set EVOKE_B200_ESTRIP=0 if the graph must be retained)")

        def weights():
            """W strip for the whole row block: K4t over the saved E strip, or K4a (recompute)."""
            if strip is None:
                return tc_bwd_w(qn, kn, bits, counts, a_row, b_col, cfg.inv_tau, flags)
            e, ld_e = strip
            tc_w_from_e(e, ld_e, kn.n, bits, counts, a_row, b_col, qn, kn, cfg.inv_tau, pos=st.pos,
                        ids=(cfg.row_ids, cfg.row_ids), clear_diag=mpc)
            return e, None, ld_e

        if not overlap:
            w_hi, w_lo, ld_w = weights()
            if need_q:
                dq = tc_bwd_gemm(w_hi, w_lo, ld_w, qn.n, kn.n, False, kn, flags)
                if finish:
                    d_image = l2norm_bwd(image, qn, dq, scale_dev=g, scale_host=scale, gather=cfg.gather)
            if need_k:
                dk = tc_bwd_gemm(w_hi, w_lo, ld_w, qn.n, kn.n, True, qn, flags)
                if finish:
                    d_text = l2norm_bwd(text, kn, dk, scale_dev=g, scale_host=scale)
        else:
            main = torch.cuda.current_stream()
            side = _side_stream(dev)
            # zero-filled split-K accumulators: filled on the side stream while K4a / K4t runs
            side.wait_stream(main)
            with torch.cuda.stream(side):
                dq = torch.zeros((qn.n, width), dtype=torch.float32, device=dev) if need_q else None
                dk = torch.zeros((kn.n, width), dtype=torch.float32, device=dev) if need_k else None
                filled = side.record_event()
            w_hi, w_lo, ld_w = weights()
            main.wait_event(filled)
            _shared_with(main, dq, dk)
            if need_q:
                tc_bwd_gemm(w_hi, w_lo, ld_w, qn.n, kn.n, False, kn, flags, out=dq)
            if need_k:
                # second contraction + its normalise-backward on the side stream: its CTAs take
                # over the SMs as the first contraction drains, and the image-side K1b overlaps it
                w_ready = main.record_event()
                with torch.cuda.stream(side):
                    side.wait_event(w_ready)
                    tc_bwd_gemm(w_hi, w_lo, ld_w, qn.n, kn.n, True, qn, flags, out=dk)
                    if finish:
                        d_text = l2norm_bwd(text, kn, dk, scale_dev=g, scale_host=scale)
                _shared_with(side, w_hi, w_lo, qn.hi, qn.lo, g, text, kn.norm)
            if need_q and finish:
                d_image = l2norm_bwd(image, qn, dq, scale_dev=g, scale_host=scale, gather=cfg.gather)
            if need_k:
                main.wait_stream(side)
                _shared_with(main, d_text, dk)
        if strip is not None and not torch.cuda.is_current_stream_capturing():
            st.e_strip = (None, 0)                  # drop the N^2 buffer as soon as it has been consumed
    return (d_image, d_text) if finish else (dq, dk)


class _MultiPositiveCE(torch.autograd.Function):
    """loss = f(image[, text]); cfg carries the non-differentiable pieces."""

    @staticmethod
    def forward(ctx, cfg: LossConfig, image: torch.Tensor, text: Optional[torch.Tensor]):
        need = (ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        loss, st = mpce_forward(cfg, image.detach(), None if text is None else text.detach(), need)
        ctx.st = st
        out = loss.reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out: torch.Tensor):
        g = grad_out.reshape(1).to(torch.float32).contiguous()
        d_image, d_text = mpce_backward(ctx.st, g)
        return None, d_image, d_text


def multi_positive_ce(cfg: LossConfig, image: torch.Tensor, text: Optional[torch.Tensor]) -> torch.Tensor:
    return _MultiPositiveCE.apply(cfg, image, text)


class _AvgPosCE(torch.autograd.Function):
    """The 'averaged positive logit' objectives of PretrainNewMulPos:
    kind "AG"   global_alignment_loss :748-815 - both directions, x0.5, / B, shape [1]
    kind "AMPC" multi_pos_contra_images_v0404 :670-708 - single-view rows leave the QUERIES only (:685), all
                rows stay keys, diagonal excluded, mean over the multi-view rows, shape [1]
    fp32 SIMT kernels (the small path): the class is not used by the reference's entry points and its own
    implementation is a Python loop over rows, so reference-sized batches are what matters here."""

    @staticmethod
    def forward(ctx, cfg: LossConfig, image: torch.Tensor, text: Optional[torch.Tensor]):
        ampc = cfg.kind == "AMPC"
        flags = FLAG_AVGPOS | (FLAG_EXCLUDE_DIAG if ampc else 0)
        kw = dict(want_f32=True, want_hi=False, want_lo=False)
        qn = l2norm_fwd(image, **kw)
        kn = qn if ampc else l2norm_fwd(text, **kw)
        n = qn.n
        bits, counts = posmask_build(cfg.row_ids, cfg.row_ids, clear_diag=ampc)
        row_neg, row_pos = small_fwd(qn, kn, bits, cfg.inv_tau, flags)
        if ampc:
            a_row, p_row, loss = finalize_avgpos(row_neg, row_pos, counts, shift=cfg.inv_tau, inv_count=1.0 / cfg.n_keep)
            b_col, p_col = a_row, p_row                      # (dS + dS^T) in one contraction: S is symmetric
        else:
            a_row, p_row, loss = finalize_avgpos(row_neg, row_pos, counts, shift=cfg.inv_tau, inv_count=0.5 / n)
            col_neg, col_pos = small_fwd(kn, qn, bits, cfg.inv_tau, flags)       # M is symmetric
            b_col, p_col, loss = finalize_avgpos(col_neg, col_pos, counts, shift=cfg.inv_tau, inv_count=0.5 / n, loss=loss)
        ctx.cfg, ctx.flags, ctx.qn, ctx.kn = cfg, flags, qn, kn
        ctx.aux = (bits, a_row, b_col, p_row, p_col)
        ctx.has_text = text is not None
        ctx.save_for_backward(image, text) if text is not None else ctx.save_for_backward(image)
        return loss if image.dtype == torch.float32 else loss.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out: torch.Tensor):
        cfg, flags, qn, kn = ctx.cfg, ctx.flags, ctx.qn, ctx.kn
        bits, a_row, b_col, p_row, p_col = ctx.aux
        saved = ctx.saved_tensors
        image = saved[0]
        text = saved[1] if ctx.has_text else None
        ampc = cfg.kind == "AMPC"
        g = grad_out.reshape(1).to(torch.float32).contiguous()
        scale = cfg.inv_tau / cfg.n_keep if ampc else 0.5 * cfg.inv_tau / qn.n
        d_image = d_text = None
        if ctx.needs_input_grad[1]:
            dq = small_bwd(qn, kn, bits, None, a_row, b_col, cfg.inv_tau, flags, pos_row=p_row, pos_col=p_col)
            d_image = l2norm_bwd(image, qn, dq, scale_dev=g, scale_host=scale)
        if not ampc and ctx.needs_input_grad[2]:
            dk = small_bwd(kn, qn, bits, None, b_col, a_row, cfg.inv_tau, flags, pos_row=p_col, pos_col=p_row)
            d_text = l2norm_bwd(text, kn, dk, scale_dev=g, scale_host=scale)
        return None, d_image, d_text


def avgpos_ce(cfg: LossConfig, image: torch.Tensor, text: Optional[torch.Tensor]) -> torch.Tensor:
    return _AvgPosCE.apply(cfg, image, text)


# ------------------------------------------------------------------------------------- f1: local token alignment
F1_TOKEN_SIM = os.environ.get("EVOKE_B200_F1_TOKEN_SIM", "1") == "1"   # per-sample token-similarity kernels (l <= 128)
_IDENTITY_CACHE: dict = {}


def _identity_targets(l: int, n: int, dev: torch.device):
    """Identity mask [l, ld_words] with its counts, and n ones (constants of f1, built once per shape and device)."""
    key = (l, n, dev.index if dev.index is not None else torch.cuda.current_device())
    hit = _IDENTITY_CACHE.get(key)
    if hit is None:
        eye_ids = DeviceIds(torch.arange(l, dtype=torch.int32, device=dev))
        bits, counts = posmask_build(eye_ids, eye_ids, clear_diag=False)
        hit = (bits, counts, torch.ones(n, dtype=torch.int32, device=dev))
        _IDENTITY_CACHE[key] = hit
    return hit


class _LocalTokenAlign(torch.autograd.Function):
    """Pretrain.local_text_token_alignment_loss (reference :506-526): text tokens attend over their sample's patch
    tokens (K `local_attend`), both sides are L2-normalised (K1) and an L x L token-level InfoNCE with identity targets
    is taken in both directions over all B*L rows - per sample that is the G loss with identity ids, so it runs on the
    batched small-path kernels with a shared identity mask.  fp32 throughout (the reference's sizes are tiny:
    B=32, L~99, P=49)."""

    @staticmethod
    def forward(ctx, inv_tau: float, image: torch.Tensor, text: torch.Tensor):
        b, p, d = (int(x) for x in image.shape)
        l = int(text.shape[1])
        dev = image.device
        v = image.detach().to(torch.float32).contiguous()
        t = text.detach().to(torch.float32).contiguous()
        att = torch.empty((b, l, p), dtype=torch.float32, device=dev)
        o = torch.empty((b, l, d), dtype=torch.float32, device=dev)
        _lib.call("evk_local_attend_fwd", _ptr(t), _ptr(v), b, l, p, d, _ptr(att), _ptr(o), _stream())
        t2, o2 = t.view(b * l, d), o.view(b * l, d)
        tn = l2norm_fwd(t2, want_f32=True, want_hi=False, want_lo=False)
        on = l2norm_fwd(o2, want_f32=True, want_hi=False, want_lo=False)
        n = b * l
        bits, counts, ones = _identity_targets(l, n, dev)                          # identity targets (:520), c_i = 1

        # l <= 128: register-blocked per-sample kernels (E kept for the backward: B*l*l floats); else the batched
        # small-path kernels
        token_sim = F1_TOKEN_SIM and l <= 128
        if token_sim:
            e = torch.empty((b, l, l), dtype=torch.float32, device=dev)
            row_sum = torch.empty(n, dtype=torch.float32, device=dev)
            row_pos = torch.empty(n, dtype=torch.float32, device=dev)
            col_sum = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.call("evk_token_sim_fwd", _ptr(tn.f32), _ptr(on.f32), b, l, d, float(inv_tau), _ptr(e), _ptr(row_sum),
                      _ptr(row_pos), _ptr(col_sum), _stream())

        def fwd(q, k):
            rs = torch.empty(n, dtype=torch.float32, device=dev)
            rp = torch.empty(n, dtype=torch.float32, device=dev)
            _lib.call("evk_mpce_small_fwd_batched", _ptr(q.f32), q.f32.stride(0), l * q.f32.stride(0), _ptr(k.f32),
                      k.f32.stride(0), l * k.f32.stride(0), b, l, l, d, _ptr(bits), bits.stride(0), float(inv_tau), 0,
                      _ptr(rs), _ptr(rp), l, _stream())
            return rs, rp

        if not token_sim:
            e = None
            row_sum, row_pos = fwd(tn, on)        # rows = text tokens (word_sim_1, :519-521)
            col_sum, _ = fwd(on, tn)              # rows = attended tokens (word_sim_2, :523-524)
        a_row, b_col, loss = finalize(row_sum, row_pos, ones, col_sum, col_lo=0, col_hi=n, shift=inv_tau, pos_weight=2.0,
                                      inv_count=0.5 / n)
        ctx.inv_tau, ctx.shape = inv_tau, (b, l, p, d)
        ctx.aux = (v, t, att, o2, tn, on, bits, counts, a_row, b_col, e)
        ctx.in_dtypes = (image.dtype, text.dtype)
        out = loss.reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out: torch.Tensor):
        b, l, p, d = ctx.shape
        v, t, att, o2, tn, on, bits, counts, a_row, b_col, e = ctx.aux
        dev = v.device
        inv_tau = ctx.inv_tau
        g = grad_out.reshape(1).to(torch.float32).contiguous()
        scale = 0.5 * inv_tau / (b * l)

        def bwd(q, k, a, bc):
            dq = torch.empty((b * l, d), dtype=torch.float32, device=dev)
            _lib.call("evk_mpce_small_bwd_batched", _ptr(q.f32), q.f32.stride(0), l * q.f32.stride(0), _ptr(k.f32),
                      k.f32.stride(0), l * k.f32.stride(0), b, l, l, d, _ptr(bits), bits.stride(0), _ptr(counts), _ptr(a),
                      _ptr(bc), l, float(inv_tau), 0, _ptr(dq), dq.stride(0), l * dq.stride(0), _stream())
            return dq

        if e is not None:
            d_th = torch.empty((b * l, d), dtype=torch.float32, device=dev)
            d_oh = torch.empty((b * l, d), dtype=torch.float32, device=dev)
            _lib.call("evk_token_sim_bwd", _ptr(tn.f32), _ptr(on.f32), _ptr(e), _ptr(a_row), _ptr(b_col), b, l, d, _ptr(d_th),
                      _ptr(d_oh), _stream())
        else:
            d_th = bwd(tn, on, a_row, b_col)
            d_oh = bwd(on, tn, b_col, a_row)
        d_t = l2norm_bwd(t.view(b * l, d), tn, d_th, scale_dev=g, scale_host=scale)     # through F.normalize (:515)
        d_o = l2norm_bwd(o2, on, d_oh, scale_dev=g, scale_host=scale)                   # through F.normalize (:514)
        ds = torch.empty((b, l, p), dtype=torch.float32, device=dev)
        d_v = torch.empty((b, p, d), dtype=torch.float32, device=dev)
        _lib.call("evk_local_attend_bwd", _ptr(t), _ptr(v), _ptr(att), _ptr(d_o), b, l, p, d, _ptr(ds), _ptr(d_t), _ptr(d_v),
                  _stream())
        d_image = d_v if ctx.needs_input_grad[1] else None
        d_text = d_t.view(b, l, d) if ctx.needs_input_grad[2] else None
        if d_image is not None and ctx.in_dtypes[0] != torch.float32:
            d_image = d_image.to(ctx.in_dtypes[0])
        if d_text is not None and ctx.in_dtypes[1] != torch.float32:
            d_text = d_text.to(ctx.in_dtypes[1])
        return None, d_image, d_text


def local_token_align(inv_tau: float, image: torch.Tensor, text: torch.Tensor) -> torch.Tensor:
    return _LocalTokenAlign.apply(inv_tau, image, text)
