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

import os
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


def peer_forward(ops, pc, inv_tau: float, row_ids: DeviceIds, image: torch.Tensor, text: torch.Tensor, need_grad: bool = True):
    """Forward kernel sequence of the peer-memory sharded G loss.  Returns (loss [1] fp32, state)."""
    from . import _lib
    world, rank = pc.world, pc.rank
    n, n_total, d = pc.n_local, pc.n_total, pc.d
    lo_ = rank * n
    dev = image.device
    pc.raise_if_failed()                     # a barrier of an earlier step timed out (pinned mirror: no device sync)
    overlap = ops.OVERLAP_STREAMS or torch.cuda.is_current_stream_capturing()
    main = torch.cuda.current_stream()
    side = ops._side_stream(dev)
    stream = main.cuda_stream
    two = row_ids.key2 is not None
    # exchange 1: one launch normalises both sides (keys into this rank's own buffer rows), pushes the id shard
    # to every rank and zeroes dQhat; after the barrier the key rows travel to the peers on the side stream
    # WHILE K3 runs - K3 visits its own columns first and waits per source for the others (landed flags)
    k_norm = torch.empty(n, dtype=torch.float32, device=dev)
    q_norm = torch.empty(n, dtype=torch.float32, device=dev)
    q_hi = torch.empty((n, pc.ld), dtype=torch.bfloat16, device=dev)
    wq = _round_up(d, 4)
    dq = torch.empty((n, wq), dtype=torch.float32, device=dev) if need_grad else None
    counts = torch.empty(n, dtype=torch.int32, device=dev)          # zeroed by the prologue kernel
    overlap_gather = OVERLAP_GATHER and world > 1
    folded = pc.sync(1) is not None                                  # syncs folded into the consumer kernels
    # inputs as the caller holds them (strided [:,0,:] head views, bf16/fp16): the prologue's loader honours them
    _lib.call("evk_shard_prologue", text.data_ptr(), ops._dtype_code(text), text.stride(0), text.stride(1),
              image.data_ptr(), ops._dtype_code(image), image.stride(0), image.stride(1), n, d,
              1 if overlap_gather else world, pc.table("khat_local") if overlap_gather else pc.table("khat"), pc.ld, lo_,
              k_norm.data_ptr(), q_hi.data_ptr(), q_norm.data_ptr(),
              row_ids.key.data_ptr(), row_ids.key2.data_ptr() if two else None, world, pc.table("ids"),
              pc.table("ids2") if two else None, None if dq is None else dq.data_ptr(), wq, counts.data_ptr(), n,
              pc.step.data_ptr(), pc.error.data_ptr(), stream)
    if not folded:
        pc.barrier()
    sync1 = pc.sync(1) if folded else None
    qn = ops.Normalized(n=n, d=d, norm=q_norm, hi=q_hi, lo=None, ld=pc.ld)
    kn_all = ops.Normalized(n=n_total, d=d, norm=None, hi=pc.khat, lo=None, ld=pc.ld)
    kn_local = ops.Normalized(n=n, d=d, norm=k_norm, hi=pc.khat[lo_:lo_ + n], lo=None, ld=pc.ld)
    ids_all = DeviceIds(pc.ids, pc.ids2 if two else None)
    if overlap_gather:
        side.wait_stream(main)
        with torch.cuda.stream(side):
            _lib.call("evk_peer_push_shard", pc.khat[lo_:].data_ptr(), n * pc.ld * 2, world, rank, pc.table("khat"),
                      lo_ * pc.ld * 2, pc.table("landed"), PUSH_CTAS, side.cuda_stream)

    def sweep(bits, store):
        """K3 over this rank's row block (with the per-source waits when the gather is still in flight)."""
        if not overlap_gather:
            return ops.tc_fwd_store(qn, kn_all, bits, inv_tau, 0, lo_) if store else \
                ops.tc_fwd_partials(qn, kn_all, bits, inv_tau, 0, lo_) + (None, 0)
        rs, rp, cs = ops.alloc_partials(n, n_total, dev, bits is not None, True)
        ld_e = _lib.size("evk_mpce_strip_ld", n_total)
        e = torch.empty((n, ld_e), dtype=torch.bfloat16, device=dev) if store else None
        _lib.call("evk_mpce_fwd_store_gathered", q_hi.data_ptr(), pc.ld, pc.khat.data_ptr(), pc.ld, n, n_total, d,
                  None if bits is None else bits.data_ptr(), 0 if bits is None else bits.stride(0), float(inv_tau),
                  ops.FLAG_NO_POS if bits is None else 0, lo_, rs.data_ptr(), None if rp is None else rp.data_ptr(), n,
                  cs.data_ptr(), n_total, None if e is None else e.data_ptr(), ld_e, pc.landed.data_ptr(),
                  pc.step.data_ptr(), pc.error.data_ptr(), n, lo_, PUSH_CTAS, stream)
        return rs, rp, cs, e, (ld_e if store else 0)

    pos = None
    mask_free = need_grad and ops.MASK_FREE
    if need_grad:
        bits, counts, pos_idx = ops.posmask_build(row_ids, ids_all, clear_diag=False, diag_offset=lo_, want_list=True,
                                                  want_bits=not mask_free, counts=counts, sync=sync1)

        def positives():
            pd = ops.pos_logits(qn, kn_all, pos_idx, counts)
            rp = ops.pos_from_lists(qn, kn_all, row_ids, ids_all, counts, pd, inv_tau, clear_diag=False,
                                    diag_offset=lo_) if mask_free else None
            return pd, rp

        if overlap or overlap_gather:      # exact positive logits (O(n D)) next to K3, once every shard has landed
            side.wait_stream(main)
            with torch.cuda.stream(side):
                if overlap_gather:
                    _lib.call("evk_peer_wait_landed", pc.landed.data_ptr(), world, pc.step.data_ptr(), PUSH_CTAS,
                              pc.error.data_ptr(), pc.timeout_ms, side.cuda_stream)
                pos_dot, row_pos_l = positives()
            ops._shared_with(side, q_hi, pos_idx, counts)
        else:
            pos_dot, row_pos_l = positives()
        pos = (pos_idx, pos_dot)
        if mask_free and overlap_gather:
            rs_part, _, cs_part, e, ld_e = sweep(None, True)
            main.wait_stream(side)          # the statistics need the positive sums; the push is part of this step
            ops._shared_with(main, pos_dot, row_pos_l)
            rp_part = row_pos_l.unsqueeze(0)
        elif mask_free:
            rs_part, _, cs_part, e, ld_e = ops.tc_fwd_store(qn, kn_all, None, inv_tau, ops.FLAG_NO_POS, lo_)
            if overlap:
                main.wait_stream(side)      # the statistics need the positive sums computed next to K3
                ops._shared_with(main, pos_dot, row_pos_l)
            rp_part = row_pos_l.unsqueeze(0)
        else:
            rs_part, rp_part, cs_part, e, ld_e = sweep(bits, True)
    else:
        bits, counts = ops.posmask_build(row_ids, ids_all, clear_diag=False, diag_offset=lo_, counts=counts, sync=sync1)
        rs_part, rp_part, cs_part, e, ld_e = sweep(bits, False)
    # exchange 2 (one launch): partials -> a_row, and this rank's slot (partial column sums of its rows +
    # its row-side loss term) into every rank's slot buffer
    a_row = torch.empty(n, dtype=torch.float32, device=dev)
    _lib.call("evk_mpce_shard_stats_push", rs_part.data_ptr(), int(rs_part.shape[0]), n, rp_part.data_ptr(),
              int(rp_part.shape[0]), n, counts.data_ptr(), n, cs_part.data_ptr(), int(cs_part.shape[0]), n_total,
              n_total, float(inv_tau), 2.0, 0.5 / n_total, a_row.data_ptr(), pc.table("slots"), world,
              rank * pc.ld_slot, pc.ws_stats.data_ptr(), pc.ws_stats.numel(), 1, stream)
    if not folded:
        pc.barrier()
    b_col = torch.empty(n_total, dtype=torch.float32, device=dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    _lib.call("evk_mpce_shard_finish", pc.slots.data_ptr(), world, pc.ld_slot, n_total, float(inv_tau), 0.5 / n_total,
              b_col.data_ptr(), loss.data_ptr(), pc.ws_finish.data_ptr(), pc.ws_finish.numel(), 1, pc.error.data_ptr(),
              pc.error_host.data_ptr(), pc.sync(2) if folded else None, stream)
    if overlap_gather or (need_grad and overlap):
        main.wait_stream(side)              # the push (and the positives) are part of this step
        if need_grad:
            ops._shared_with(main, pos_dot)
    st = ops._State()
    st.ops, st.pc, st.inv_tau = ops, pc, inv_tau
    st.sv = (qn, kn_local, kn_all, bits, counts, a_row, b_col, e, ld_e, dq)
    st.pos = pos
    st.folded = folded
    st.ids = (row_ids, ids_all)
    st.image, st.text = image, text
    return loss, st



def peer_backward(st, g: torch.Tensor):
    """Backward kernel sequence; g = upstream gradient, fp32 [1] on the device.  Returns (d_image, d_text)."""
    from . import _lib
    ops, pc, inv_tau = st.ops, st.pc, st.inv_tau
    qn, kn_local, kn_all, bits, counts, a_row, b_col, e, ld_e, dq = st.sv
    if e is None:
        raise RuntimeError("evoke_b200: backward called twice on the sharded loss (the E strip was consumed)")
    image, text = st.image, st.text
    n, n_total = pc.n_local, pc.n_total
    scale = 0.5 * inv_tau / n_total
    overlap = ops.OVERLAP_STREAMS or torch.cuda.is_current_stream_capturing()
    main = torch.cuda.current_stream()
    ops.tc_w_from_e(e, ld_e, n_total, bits, counts, a_row, b_col, qn, kn_all, inv_tau, pos=st.pos, ids=st.ids,
                    clear_diag=False, diag_offset=pc.rank * n)
    # exchange 3, fused: the tiles of this rank's partial dKhat are stored straight into their owners'
    # per-source buffers (posted NVLink stores from the GEMM epilogue; no split-K, no zero fill)
    # with two streams the two contractions run side by side, half of the SMs each: the scattering one is bound by
    # NVLink (its tiles leave as remote stores), so it gives up SMs it cannot use while the local one fills them
    half = SIDE_BY_SIDE_CTAS if (overlap and pc.world > 1) else 0
    w_ready = main.record_event() if overlap else None       # W is complete: both contractions may start
    _lib.call("evk_mpce_bwd_gemm_scatter", e.data_ptr(), None, ld_e, n, n_total, qn.hi.data_ptr(), None, qn.ld, qn.d,
              1.0, 0, pc.table("dk_mine"), pc.world, n, pc.width, 2 if pc.exchange == "bf16" else 1,
              SCATTER_FIRST_OWNER(pc), half, main.cuda_stream)

    def image_side():
        ops.tc_bwd_gemm(e, None, ld_e, n, n_total, False, kn_all, 0, out=dq, cta_limit=half)   # dq was zeroed by the prologue
        return ops.l2norm_bwd(image, qn, dq, scale_dev=g, scale_host=scale, error=pc.error)

    if overlap:
        # the local contraction fills the SMs as the scattering one drains (it may be NVLink-bound)
        side = ops._side_stream(image.device)
        with torch.cuda.stream(side):
            side.wait_event(w_ready)
            d_image = image_side()
        ops._shared_with(side, e, dq, g, image, qn.norm)
    else:
        d_image = image_side()
    if overlap:
        # the barrier also tells the peers that this rank is done READING its gathered key rows (the local
        # contraction does), so that they may overwrite them in the next step: join before signalling
        main.wait_stream(side)
        ops._shared_with(main, d_image)
    if not st.folded:
        pc.barrier()                       # every rank's partial for these rows has landed
    d_text = ops.l2norm_bwd(text, kn_local, pc.dk_parts[0], scale_dev=g, scale_host=scale,
                            parts=(pc.world, n * pc.width), error=pc.error, sync=pc.sync(3) if st.folded else None)
    if not torch.cuda.is_current_stream_capturing():
        st.sv = (qn, kn_local, kn_all, bits, counts, a_row, b_col, None, 0, None)
    return d_image, d_text


class _ShardedGPeer(torch.autograd.Function):
    """Peer-memory form (bf16 mode): the three exchange steps are done by this library's own kernels over
    NVLink - the prologue kernel stores the normalised key rows and the ids into every rank's buffers
    (all-gather), the statistics slots are pushed the same way, and the key-side gradient contraction stores
    its tiles straight into the owning rank's per-source buffer, which K1b sums (reduce-scatter fused into the
    GEMM epilogue).  Three cross-GPU syncs per step order them, folded into the first consumer kernel of each
    exchange (K2, finish, the final K1b); NCCL is not on the data path.  The forward keeps E as a bf16 strip (K3 store
    variant), so the backward needs no second similarity sweep: 6 nND FLOP per rank.  Launches per step: prologue, K2
    (lists), positives x2, K3, statistics+push, finish | K4t x3, contraction+scatter beside the local contraction
    (half of the SMs each), K1b, K1b over the partials.  peer_forward / peer_backward are plain functions so that
    the same sequence is also captured into CUDA graphs (evoke_b200/graphs.py)."""

    @staticmethod
    def forward(ctx, ops, pc, inv_tau: float, row_ids: DeviceIds, image: torch.Tensor, text: torch.Tensor):
        loss, ctx.st = peer_forward(ops, pc, inv_tau, row_ids, image.detach(), text.detach(), any(ctx.needs_input_grad[4:]))
        out = loss.reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        d_image, d_text = peer_backward(ctx.st, grad_out.reshape(1).to(torch.float32).contiguous())
        return None, None, None, None, d_image, d_text


class _GatheredLoss(torch.autograd.Function):
    """Small global batches (reference-sized per-GPU batches: 32 studies per rank): sharding the N x N work would
    only add exchange latency, so every rank all-gathers the embeddings and ids ONCE, evaluates the whole loss with the
    single-device kernels, and keeps the gradient rows of its own shard - no collective in the backward at all.
    kind "G": (image, text) -> global_alignment_loss; kind "MPC": image only, rows filtered as on one device."""

    @staticmethod
    def forward(ctx, ops, group, cfg_of, image: torch.Tensor, text: Optional[torch.Tensor]):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n = int(image.shape[0])
        img_all, w1 = _all_gather_rows(image.detach(), group, async_op=True)
        txt_all, w2 = (None, None) if text is None else _all_gather_rows(text.detach(), group, async_op=True)
        w1.wait()
        if w2 is not None:
            w2.wait()
        loss, ctx.st = ops.mpce_forward(cfg_of(), img_all, txt_all, (True, text is not None))
        ctx.ops, ctx.rows = ops, (rank * n, (rank + 1) * n)
        out = loss.reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        d_i, d_t = ctx.ops.mpce_backward(ctx.st, grad_out.reshape(1).to(torch.float32).contiguous())
        lo, hi = ctx.rows
        return None, None, None, d_i[lo:hi], None if d_t is None else d_t[lo:hi]


# global batches up to this many rows are gathered and evaluated on every rank instead of being sharded
GATHER_ALL_MAX = int(os.environ.get("EVOKE_B200_GATHER_ALL_MAX", "2048"))


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


_FLOAT_DTYPES = (torch.float32, torch.bfloat16, torch.float16)


def peer_eligible(image: torch.Tensor, text: torch.Tensor, precision: str, world: int) -> bool:
    """Conditions of the peer-memory path.  They depend on shapes, dtypes and the precision mode only - never on
    strides or pointer alignment, which may differ between ranks: every rank must pick the same transport (the
    context set-up is collective), and the prologue kernel reads any strides / fp32, bf16, fp16 inputs itself."""
    n, d = int(image.shape[0]), int(image.shape[1])
    return (precision == "bf16" and image.is_cuda and text.is_cuda and world <= 16 and n % 128 == 0 and d % 8 == 0
            and d <= 2048 and image.dtype == text.dtype and image.dtype in _FLOAT_DTYPES)


def SCATTER_FIRST_OWNER(pc) -> int:
    """Owner whose rows the fused reduce-scatter contraction computes (and sends) first: rank + 1, so that at any
    moment every GPU is the destination of about one sender (EVOKE_B200_SCATTER_ROTATE=0: owner 0 on every rank, the
    round-1 order, kept for A/B measurements)."""
    return (pc.rank + 1) % pc.world if os.environ.get("EVOKE_B200_SCATTER_ROTATE", "1") == "1" else 0


# CTAs each of the two backward contractions of the sharded step may use when they run side by side (0: all SMs, one
# after the other as in round 1)
SIDE_BY_SIDE_CTAS = int(os.environ.get("EVOKE_B200_SIDE_BY_SIDE_CTAS", "74"))
PEER_EXCHANGE = os.environ.get("EVOKE_B200_PEER_EXCHANGE", "bf16")     # dtype of the dKhat partials on NVLink: bf16 | fp32
# EVOKE_B200_OVERLAP_GATHER=1: the key rows travel on a side stream WHILE K3 runs (TMA bulk copies driven by one
# thread per CTA, one destination at a time; per-source landed counters; K3 sweeps its own columns first and waits per
# source).  Round 1's version (an SM-resident copy kernel) lost 0.383 vs 0.321 ms at 8 GPUs; see profiles/r2_experiments.md.
OVERLAP_GATHER = os.environ.get("EVOKE_B200_OVERLAP_GATHER", "0") == "1"
PUSH_CTAS = int(os.environ.get("EVOKE_B200_PUSH_CTAS", "148"))          # one driving thread each


def global_alignment_sharded(image: torch.Tensor, text: torch.Tensor, ids_local, temp: float, *,
                             group=None, precision: str = "bf16", mode: str = "auto", ops=None,
                             graph: Optional[bool] = None) -> torch.Tensor:
    """G loss of the GLOBAL batch from this rank's shard (same row count on every rank).

    image, text: [n_local, D] on this rank's device; ids_local: the n_local ids of these rows
    (numpy array / int tensor / DeviceIds; string keys must be factorised consistently across
    ranks by the caller, e.g. with a shared vocabulary - ints are used as they are).
    Returns the global loss (identical on every rank); .backward() yields d(global loss)/d(local
    shard), so a DDP-style gradient average over ranks must not be applied to it twice.

    graph: peer transport only - replay the step from the CUDA-graph cache (see evoke_b200.loss.global_alignment;
    None: the EVOKE_B200_GRAPHS default, on).
    mode: "peer" - exchanges done by this library's kernels over peer-mapped memory (NVLink): K1 stores into every
                  rank's key buffer, K4b's epilogue red.adds into the owner's dKhat (6 nND FLOP per rank); bf16 mode;
          "rs"  - partial dKhat for all N keys + reduce-scatter (8 nND FLOP per rank, N*D fp32 exchanged);
          "sym" - queries are all-gathered as well and the key-side row block is recomputed locally
                  (10 nND FLOP per rank, no reduce-scatter, no N*D buffer);
          "auto" - "peer" when its conditions hold and CUDA-IPC mapping works, else "sym" from 8 ranks up, where the N*D exchange dominates (measured: rs wins at 2-4, sym at 8).
    """
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if ops is None:
        from . import functional as ops  # the CUDA kernels
        ops._require_cuda(image, "image")
        ops._require_cuda(text, "text")
        if image.dtype not in _FLOAT_DTYPES:
            raise TypeError(f"embeddings must be float32, bfloat16 or float16, got {image.dtype}")
    if image.shape != text.shape:
        raise ValueError(f"image/text shapes differ: {tuple(image.shape)} vs {tuple(text.shape)}")
    if image.dtype != text.dtype:
        raise TypeError(f"image/text embedding dtypes differ: {image.dtype} vs {text.dtype}")
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
    from .loss import _inv_tau
    inv_tau = _inv_tau(temp)
    if mode not in ("auto", "rs", "sym", "peer"):
        raise ValueError(f"mode must be 'auto', 'rs', 'sym' or 'peer', got {mode!r}")
    world = dist.get_world_size(group)
    n_total = int(image.shape[0]) * world
    if mode == "auto" and n_total <= GATHER_ALL_MAX and hasattr(ops, "mpce_forward"):
        key_all, _ = _all_gather_rows(row_ids.key, group)
        key2_all = None if row_ids.key2 is None else _all_gather_rows(row_ids.key2, group)[0]
        d = int(image.shape[1])
        path = ops.choose_path("auto", n_total, n_total, d)

        def cfg_of():
            return ops.LossConfig(kind="G", inv_tau=inv_tau, precision=precision, path=path,
                                  row_ids=DeviceIds(key_all, key2_all))
        return _GatheredLoss.apply(ops, group, cfg_of, image, text)
    if mode in ("auto", "peer") and hasattr(ops, "tc_fwd_store"):
        pc = None
        if peer_eligible(image, text, precision, world):
            from . import peer
            pc = peer.get_context(group, int(image.shape[0]), int(image.shape[1]), image.device,
                                  two_keys=row_ids.key2 is not None, exchange=PEER_EXCHANGE)
        if pc is not None:
            from . import graphs
            use_graph = graphs.DROPIN_GRAPHS if graph is None else bool(graph)
            if use_graph and not torch.cuda.is_current_stream_capturing():
                pc.raise_if_failed()          # (a replayed graph never re-enters peer_forward's own check)
                # collective: every rank captures at the same call (same shapes on every rank); the warm-up and the
                # captured sequences contain the flag barriers
                def fwd(im, tx, ids, need):
                    return peer_forward(ops, pc, inv_tau, ids, im, tx, any(need))
                return graphs.graphed_call(("Gpeer", id(pc), inv_tau), fwd, peer_backward, image, text, row_ids)
            return _ShardedGPeer.apply(ops, pc, inv_tau, row_ids, image, text)
        if mode == "peer":
            raise RuntimeError("evoke_b200: mode='peer' needs bf16 precision, n_local % 128 == 0, d % 8 == 0, d <= 2048 "
                               "and CUDA-IPC peer mapping between the ranks' GPUs")
    sym = mode == "sym" or (mode in ("auto", "peer") and world >= 8)
    return _ShardedG.apply(ops, group, inv_tau, precision, sym, row_ids, image, text)


# ------------------------------------------------------------------------------------- sharded MPC (a7)
class _ShardedMPC(torch.autograd.Function):
    """multi_pos_contra_images_v0401 (reference :421-446) of the GLOBAL set of views from this rank's shard.

    The objective is symmetric (one matrix on both sides), so one exchange suffices: the raw view
    rows are all-gathered once, every rank normalises the kept rows (views whose study has another view somewhere in
    the global batch, :424-429) and sweeps its own row block against all of them - E strip / statistics exactly as
    on one device, rectangular [n' x N'] with the diagonal at column offset k0.  The row exp-sums of all ranks are
    exchanged in one packed all-reduce (W needs a_j = 1/R_j of every column: W_ij = E_ij (a_i + a_j) - 2 M_ij / c_i
    carries both the query-side and the key-side term of dS + dS^T), after which dX_r = W_r Xhat / (M' tau) is a
    LOCAL contraction: no reduce-scatter, 4 n' N' D FLOP per rank."""

    @staticmethod
    def forward(ctx, ops, group, inv_tau: float, precision: str, codes_all, x: torch.Tensor):
        import numpy as np
        from . import ids as idmod
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        m = int(x.shape[0])
        lo = rank * m
        dev = x.device
        keep_all = idmod.multi_view_rows(codes_all)                      # sorted global indices of the kept rows
        n_keep = int(len(keep_all))
        k0, k1 = int(np.searchsorted(keep_all, lo)), int(np.searchsorted(keep_all, lo + m))
        split = precision == "fp32"
        flags = ops.FLAG_EXCLUDE_DIAG | ops.FLAG_NO_COLSUM | (ops.FLAG_SPLIT_BF16 if split else 0)
        xg = x.detach() if x.dtype == torch.float64 else x.detach().to(torch.float32)       # (fp64: CPU choreography tests)
        x_all, work = _all_gather_rows(xg, group, async_op=True)                             # [N, D] raw rows
        keep_dev = torch.from_numpy(np.ascontiguousarray(keep_all)).to(dev)
        ids_keep = DeviceIds(torch.from_numpy(np.ascontiguousarray(codes_all[keep_all])).to(dev))
        work.wait()
        kn = ops.l2norm_fwd(x_all, want_f32=False, want_hi=True, want_lo=split, gather=keep_dev)       # [N', ld]
        packed = torch.zeros(2 * n_keep, dtype=torch.float64 if x.dtype == torch.float64 else torch.float32, device=dev)
        st = None
        if k1 > k0:
            qn = ops.rows_of(kn, k0, k1)
            row_ids = DeviceIds(ids_keep.key[k0:k1])
            use_strip = hasattr(ops, "tc_fwd_store") and ops.E_STRIP and not split and ctx.needs_input_grad[5]
            mask_free = use_strip and ops.MASK_FREE
            if use_strip:
                bits, counts, pos_idx = ops.posmask_build(row_ids, ids_keep, clear_diag=True, diag_offset=k0, want_list=True,
                                                          want_bits=not mask_free)
                pos_dot = ops.pos_logits(qn, kn, pos_idx, counts)
                if mask_free:
                    rs_part, _, _, e, ld_e = ops.tc_fwd_store(qn, kn, None, inv_tau, flags | ops.FLAG_NO_POS, k0)
                    rp_part = ops.pos_from_lists(qn, kn, row_ids, ids_keep, counts, pos_dot, inv_tau, clear_diag=True,
                                                 diag_offset=k0).unsqueeze(0)
                else:
                    rs_part, rp_part, _, e, ld_e = ops.tc_fwd_store(qn, kn, bits, inv_tau, flags, k0)
                strip, pos = (e, ld_e), (pos_idx, pos_dot)
            else:
                bits, counts = ops.posmask_build(row_ids, ids_keep, clear_diag=True, diag_offset=k0)
                rs_part, rp_part, _ = ops.tc_fwd_partials(qn, kn, bits, inv_tau, flags, k0)
                strip = pos = None
            ops.reduce_partials(rs_part, int(rs_part.shape[0]), k1 - k0, out=packed[k0:k1])
            ops.reduce_partials(rp_part, int(rp_part.shape[0]), k1 - k0, out=packed[n_keep + k0: n_keep + k1], divisor=counts)
            st = (qn, row_ids, bits, counts, strip, pos)
        # the one exchange after the sweep: [row exp-sums | pos_i / c_i] of every kept row (own slice, zeros elsewhere)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        a_all, _, loss = ops.stats_fused(packed[:n_keep], packed[n_keep:], None, None, shift=inv_tau, pos_weight=1.0,
                                         inv_count=1.0 / n_keep)
        ctx.ops, ctx.inv_tau, ctx.flags = ops, inv_tau, flags
        ctx.sv = (st, kn, ids_keep, a_all, k0, k1, n_keep, keep_dev, lo)
        ctx.save_for_backward(x)
        out = loss.reshape(())
        return out if x.dtype == torch.float32 else out.to(x.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        ops, inv_tau, flags = ctx.ops, ctx.inv_tau, ctx.flags
        st, kn, ids_keep, a_all, k0, k1, n_keep, keep_dev, lo = ctx.sv
        (x,) = ctx.saved_tensors
        if st is None:                                   # no kept row in this shard
            return None, None, None, None, None, torch.zeros_like(x)
        qn, row_ids, bits, counts, strip, pos = st
        g = grad_out.reshape(1).to(torch.float32).contiguous()
        a_row = a_all[k0:k1].contiguous()
        if strip is not None:
            e, ld_e = strip
            ops.tc_w_from_e(e, ld_e, n_keep, bits, counts, a_row, a_all, qn, kn, inv_tau, pos=pos, ids=(row_ids, ids_keep),
                            clear_diag=True, diag_offset=k0)
            w_hi, w_lo, ld_w = e, None, ld_e
        else:
            w_hi, w_lo, ld_w = ops.tc_bwd_w(qn, kn, bits, counts, a_row, a_all, inv_tau, flags, k0)
        dq = ops.tc_bwd_gemm(w_hi, w_lo, ld_w, k1 - k0, n_keep, False, kn, flags)
        local_rows = (keep_dev[k0:k1] - lo).to(torch.int32).contiguous()
        d_x = ops.l2norm_bwd(x, qn, dq, scale_dev=g, scale_host=inv_tau / n_keep, gather=local_rows)
        return None, None, None, None, None, d_x


def multi_pos_contra_images_sharded(image: torch.Tensor, ids_local, temp: float, *, group=None, precision: str = "bf16",
                                    ops=None) -> torch.Tensor:
    """Image<->image multi-positive loss (reference :421-446) of the GLOBAL set of views from this rank's shard
    (same number of views on every rank).  ids_local: this rank's integer study ids (numpy / tensor; string keys
    must be factorised with a vocabulary shared by all ranks).  Returns the global loss (identical on every rank;
    tensor([0.0]) leaf of shape [1] when no study of the global batch has a second view, :427-428);
    .backward() yields d(global loss)/d(local shard).  One host round trip for the ids, as in the reference (:427)."""
    import numpy as np
    if not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    if ops is None:
        from . import functional as ops
        ops._require_cuda(image, "image")
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    from .loss import _inv_tau
    inv_tau = _inv_tau(temp)
    m = int(image.shape[0])
    if isinstance(ids_local, DeviceIds):
        if ids_local.key2 is not None:
            raise TypeError("sharded MPC takes one integer key per view (combine (patient, study) on the host)")
        ids_local = ids_local.key
    if isinstance(ids_local, np.ndarray):
        if ids_local.dtype.kind not in "iu":
            raise TypeError("sharded ids must be integers (factorise string keys with a vocabulary shared by all ranks)")
        ids_local = torch.from_numpy(np.ascontiguousarray(ids_local.astype(np.int64)))
    ids_t = torch.as_tensor(ids_local).reshape(-1).to(device=image.device, dtype=torch.int64)
    if int(ids_t.shape[0]) != m:
        raise ValueError(f"ids_local has {int(ids_t.shape[0])} entries for {m} views")
    world = dist.get_world_size(group)
    ids_all = torch.empty(world * m, dtype=torch.int64, device=image.device)
    dist.all_gather_into_tensor(ids_all, ids_t.contiguous(), group=group)
    from . import ids as idmod
    codes_all = idmod.factorize(ids_all.cpu().numpy())                 # the size-determining host sync (:427)
    keep_all = idmod.multi_view_rows(codes_all)
    if len(keep_all) == 0:
        return torch.tensor([0.0], requires_grad=True, device=image.device)
    if world * m <= GATHER_ALL_MAX and hasattr(ops, "mpce_forward"):
        n_keep = len(keep_all)
        gather = None if n_keep == world * m else torch.from_numpy(np.ascontiguousarray(keep_all)).to(image.device)
        codes_dev = torch.from_numpy(np.ascontiguousarray(codes_all if gather is None else codes_all[keep_all])).to(image.device)
        path = ops.choose_path("auto", n_keep, n_keep, int(image.shape[1]))

        def cfg_of():
            return ops.LossConfig(kind="MPC", inv_tau=inv_tau, precision=precision, path=path,
                                  row_ids=DeviceIds(codes_dev), gather=gather)
        return _GatheredLoss.apply(ops, group, cfg_of, image, None)
    return _ShardedMPC.apply(ops, group, inv_tau, precision, codes_all, image)
