"""Sharded (one process per GPU) form of the G loss: rank r of R owns rows
[r*n, (r+1)*n) of the image and text embeddings and the matching ids (a data-parallel batch
shard) and obtains the loss and gradients of the GLOBAL batch.

The reference has no sharded implementation (its only multi-GPU mode is nn.DataParallel,
modules/trainer_v0401.py:23-29, under which negatives would be per-replica); the result here is
defined as the single-device reference on the concatenated batch.

Exchange steps (torch.distributed, NCCL over NVLink on the GPU box, gloo in the CPU tests):
  forward   all-gather  That (bf16, N*D*2 B) and ids (4N B)   [asynchronous, under K1/K2]
            all-reduce  ONE packed vector: column exp-sums | row exp-sums | pos_i/c_i  (3N fp32)
  backward  reduce-scatter  partial dThat (N x D fp32) -> each rank's rows
The local dQhat contraction is issued while the reduce-scatter is in flight (side stream).

`ops` is the kernel namespace (evoke_b200.functional on a GPU).  The CPU gloo tests pass a
numpy-backed stand-in with the same function names so that the collective choreography can be
checked without a GPU; the product never does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
import torch.distributed as dist

from .ids import DeviceIds


def _all_gather_rows(t: torch.Tensor, group, async_op: bool = False):
    """[n, ...] per rank -> [R*n, ...] (rank order), equal n on every rank.  Returns (out, work)."""
    world = dist.get_world_size(group)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    work = dist.all_gather_into_tensor(out, t.contiguous(), group=group, async_op=async_op)
    return out, work


@dataclass
class _Saved:
    qn: object
    kn_local: object
    kn_all: object
    bits: torch.Tensor
    counts: torch.Tensor
    a_row: torch.Tensor
    b_col: torch.Tensor
    flags: int
    n_local: int
    n_total: int
    rank: int
    qn_all: object = None          # "sym" mode: all-gathered queries
    a_all: torch.Tensor = None     # "sym" mode: 1/R_i of every row of the global batch


class _ShardedG(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ops, group, inv_tau: float, precision: str, sym: bool, row_ids: DeviceIds,
                image: torch.Tensor, text: torch.Tensor):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = int(image.shape[0])
        n_total = n * world
        split = precision == "fp32"
        flags = ops.FLAG_SPLIT_BF16 if split else 0
        kw = dict(want_f32=False, want_hi=True, want_lo=split)
        # exchange 1 (asynchronous): ids first (tiny, K2 needs them), then the normalised keys; both
        # overlap the image-side K1 and K2
        key_all, w_ids = _all_gather_rows(row_ids.key, group, async_op=True)
        key2_all, w_ids2 = (None, None) if row_ids.key2 is None else _all_gather_rows(row_ids.key2, group, async_op=True)
        kn_local = ops.l2norm_fwd(text, **kw)
        k_hi_all, w_hi = _all_gather_rows(kn_local.hi, group, async_op=True)
        k_lo_all, w_lo = _all_gather_rows(kn_local.lo, group, async_op=True) if split else (None, None)
        qn = ops.l2norm_fwd(image, **kw)
        if sym:   # the key-side row block of the backward needs every query: gather them now, use them later
            q_hi_all, wq_hi = _all_gather_rows(qn.hi, group, async_op=True)
            q_lo_all, wq_lo = _all_gather_rows(qn.lo, group, async_op=True) if split else (None, None)
        w_ids.wait()
        if w_ids2 is not None:
            w_ids2.wait()
        ids_all = DeviceIds(key_all, key2_all)
        bits, counts = ops.posmask_build(row_ids, ids_all, clear_diag=False, diag_offset=rank * n)
        w_hi.wait()
        if w_lo is not None:
            w_lo.wait()
        kn_all = ops.Normalized(n=n_total, d=kn_local.d, norm=None, hi=k_hi_all, lo=k_lo_all, ld=kn_local.ld)
        rs_part, rp_part, cs_part = ops.tc_fwd_partials(qn, kn_all, bits, inv_tau, flags, rank * n)
        # exchange 2 - the only one after K3: ONE all-reduce of [column partial sums (N) | row sums (N) |
        # pos_i/c_i (N)], the last two placed in this rank's slice (zeros elsewhere, so the sum is a
        # gather).  The fixed shift makes the column sums additive across ranks.  Afterwards every rank
        # holds the statistics of the whole batch and finishes the loss locally, identically.
        packed = torch.zeros(3 * n_total, dtype=cs_part.dtype, device=cs_part.device)
        lo_, hi_ = rank * n, (rank + 1) * n
        ops.reduce_partials(cs_part, int(cs_part.shape[0]), n_total, out=packed[:n_total])
        ops.reduce_partials(rs_part, int(rs_part.shape[0]), n, out=packed[n_total + lo_: n_total + hi_])
        ops.reduce_partials(rp_part, int(rp_part.shape[0]), n, out=packed[2 * n_total + lo_: 2 * n_total + hi_],
                            divisor=counts)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        a_all, b_col, loss = ops.stats_fused(packed[n_total: 2 * n_total], packed[2 * n_total:], packed[:n_total],
                                             None, shift=inv_tau, pos_weight=2.0, inv_count=0.5 / n_total)
        a_row = a_all[lo_:hi_].contiguous()
        ctx.ops, ctx.group, ctx.inv_tau, ctx.sym = ops, group, inv_tau, sym
        ctx.sv = _Saved(qn, kn_local, kn_all, bits, counts, a_row, b_col, flags, n, n_total, rank)
        if sym:
            wq_hi.wait()
            if wq_lo is not None:
                wq_lo.wait()
            ctx.sv.qn_all = ops.Normalized(n=n_total, d=qn.d, norm=None, hi=q_hi_all, lo=q_lo_all, ld=qn.ld)
            ctx.sv.a_all = a_all
        ctx.save_for_backward(image, text)
        out = loss.reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        ops, group, inv_tau, sv = ctx.ops, ctx.group, ctx.inv_tau, ctx.sv
        image, text = ctx.saved_tensors
        g = grad_out.reshape(1).to(torch.float32).contiguous()
        scale = 0.5 * inv_tau / sv.n_total
        if ctx.sym:
            # CLIP-style: both gradients from row blocks this rank owns; no reduce-scatter.
            #   W  = K4a(Qhat_r, Khat_all; a = a_r,  b = b_all)   dQhat_r = W  Khat_all
            #   W' = K4a(Khat_r, Qhat_all; a = b_r,  b = a_all)   dKhat_r = W' Qhat_all
            # M is symmetric and M_ij = 1 implies c_i = c_j, so the mask and the counts are shared.
            lo, hi = sv.rank * sv.n_local, (sv.rank + 1) * sv.n_local
            w_hi, w_lo, ld_w = ops.tc_bwd_w(sv.qn, sv.kn_all, sv.bits, sv.counts, sv.a_row, sv.b_col, inv_tau,
                                            sv.flags, lo)
            dq = ops.tc_bwd_gemm(w_hi, w_lo, ld_w, sv.n_local, sv.n_total, False, sv.kn_all, sv.flags)
            d_image = ops.l2norm_bwd(image, sv.qn, dq, scale_dev=g, scale_host=scale)
            b_local = sv.b_col[lo:hi].contiguous()
            w_hi, w_lo, ld_w = ops.tc_bwd_w(sv.kn_local, sv.qn_all, sv.bits, sv.counts, b_local, sv.a_all, inv_tau,
                                            sv.flags, lo)
            dk = ops.tc_bwd_gemm(w_hi, w_lo, ld_w, sv.n_local, sv.n_total, False, sv.qn_all, sv.flags)
            d_text = ops.l2norm_bwd(text, sv.kn_local, dk, scale_dev=g, scale_host=scale)
            return None, None, None, None, None, None, d_image, d_text
        w_hi, w_lo, ld_w = ops.tc_bwd_w(sv.qn, sv.kn_all, sv.bits, sv.counts, sv.a_row, sv.b_col, inv_tau, sv.flags,
                                        sv.rank * sv.n_local)
        # partial dKhat for ALL columns from this rank's rows, then reduce-scatter to the owners
        dk_part = ops.tc_bwd_gemm(w_hi, w_lo, ld_w, sv.n_local, sv.n_total, True, sv.qn, sv.flags)
        dk_local = torch.empty((sv.n_local, dk_part.shape[1]), dtype=dk_part.dtype, device=dk_part.device)
        work = dist.reduce_scatter_tensor(dk_local, dk_part, op=dist.ReduceOp.SUM, group=group, async_op=True)
        # local dQhat while the reduce-scatter is in flight
        dq = ops.tc_bwd_gemm(w_hi, w_lo, ld_w, sv.n_local, sv.n_total, False, sv.kn_all, sv.flags)
        d_image = ops.l2norm_bwd(image, sv.qn, dq, scale_dev=g, scale_host=scale)
        work.wait()
        d_text = ops.l2norm_bwd(text, sv.kn_local, dk_local, scale_dev=g, scale_host=scale)
        return None, None, None, None, None, None, d_image, d_text


def global_alignment_sharded(image: torch.Tensor, text: torch.Tensor, ids_local, temp: float, *,
                             group=None, precision: str = "bf16", mode: str = "auto", ops=None) -> torch.Tensor:
    """G loss of the GLOBAL batch from this rank's shard (same row count on every rank).

    image, text: [n_local, D] on this rank's device; ids_local: the n_local ids of these rows
    (numpy array / int tensor / DeviceIds; string keys must be factorised consistently across
    ranks by the caller, e.g. with a shared vocabulary - ints are used as they are).
    Returns the global loss (identical on every rank); .backward() yields d(global loss)/d(local
    shard), so a DDP-style gradient average over ranks must not be applied to it twice.

    mode: "rs"  - partial dKhat for all N keys + reduce-scatter (8 nND FLOP per rank, N*D fp32 exchanged);
          "sym" - queries are all-gathered as well and the key-side row block is recomputed locally
                  (10 nND FLOP per rank, no reduce-scatter, no N*D buffer);
          "auto" - "sym" from 8 ranks up, where the N*D exchange dominates (measured: rs wins at 2-4, sym at 8).
    """
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if ops is None:
        from . import functional as ops  # the CUDA kernels
        ops._require_cuda(image, "image")
        ops._require_cuda(text, "text")
    if image.shape != text.shape:
        raise ValueError(f"image/text shapes differ: {tuple(image.shape)} vs {tuple(text.shape)}")
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    from . import ids as idmod
    if isinstance(ids_local, DeviceIds):
        row_ids = ids_local
    else:
        import numpy as np
        if isinstance(ids_local, np.ndarray):
            if ids_local.dtype.kind not in "iu":
                raise TypeError("sharded ids must be integers (factorise string keys with a vocabulary shared by all ranks)")
            # rank-local factorisation would give inconsistent codes: keep the raw integers (exact two-word keys)
            ids_local = torch.from_numpy(np.ascontiguousarray(ids_local.astype(np.int64)))
        row_ids, _ = idmod.to_device_ids(ids_local, image.device, n=int(image.shape[0]))
    temp = float(temp)
    if not temp > 0:
        raise ValueError("temperature must be positive")
    if mode not in ("auto", "rs", "sym"):
        raise ValueError(f"mode must be 'auto', 'rs' or 'sym', got {mode!r}")
    sym = mode == "sym" or (mode == "auto" and dist.get_world_size(group) >= 8)
    return _ShardedG.apply(ops, group, 1.0 / temp, precision, sym, row_ids, image, text)
