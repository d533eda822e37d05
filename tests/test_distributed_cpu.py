"""T4 (CPU, gloo, world_size 2): the collective choreography of evoke_b200.distributed
(all-gather keys/ids -> row-block statistics -> all-reduce column sums -> scalar all-reduce;
backward: partial dK -> reduce-scatter) equals the single-device reference on the concatenated
batch.  The kernel namespace is replaced by a numpy/torch-CPU stand-in built on the oracle's
math - test infrastructure only; the product always uses the CUDA kernels.
"""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_ops():
    """Stand-in for evoke_b200.functional with the same function names (fp64 on CPU)."""
    from dataclasses import dataclass
    from typing import Optional
    from oracle import evoke_oracle as orc

    ops = types.SimpleNamespace()
    ops.FLAG_SPLIT_BF16, ops.FLAG_EXCLUDE_DIAG, ops.FLAG_NO_COLSUM, ops.FLAG_NO_POS = 4, 1, 2, 8
    ops.E_STRIP = ops.MASK_FREE = False

    @dataclass
    class Normalized:
        n: int
        d: int
        norm: Optional[torch.Tensor]
        f32: Optional[torch.Tensor] = None
        hi: Optional[torch.Tensor] = None
        lo: Optional[torch.Tensor] = None
        ld: int = 0

    ops.Normalized = Normalized

    def l2norm_fwd(x, *, want_f32, want_hi, want_lo, gather=None):
        xs = x.detach().double().numpy()
        if gather is not None:
            xs = xs[gather.numpy()]
        xh, nrm = orc.l2_normalize(xs)
        hi = torch.from_numpy(xh)                                  # "hi" carries the exact unit rows here
        lo = torch.zeros_like(hi) if want_lo else None
        return Normalized(n=xs.shape[0], d=xs.shape[1], norm=torch.from_numpy(nrm.reshape(-1)), hi=hi, lo=lo, ld=xs.shape[1])

    def rows_of(x, r0, r1):
        return Normalized(n=r1 - r0, d=x.d, norm=None if x.norm is None else x.norm[r0:r1], hi=x.hi[r0:r1],
                          lo=None if x.lo is None else x.lo[r0:r1], ld=x.ld)

    def posmask_build(rows, cols, *, clear_diag, diag_offset=0):
        m = orc.posmask_dense(rows.key.numpy(), cols.key.numpy(), clear_diag, diag_offset)
        return torch.from_numpy(m), torch.from_numpy(m.sum(1).astype(np.int32))

    def _e(q, k, inv_tau, flags=0, diag_offset=0):
        s = (q.hi + (q.lo if q.lo is not None else 0)) @ (k.hi + (k.lo if k.lo is not None else 0)).T * inv_tau
        e = torch.exp(s - inv_tau)
        if flags & ops.FLAG_EXCLUDE_DIAG:                          # column (row + diag_offset) leaves the softmax (:438)
            r = torch.arange(e.shape[0])
            e[r, r + diag_offset] = 0.0
        return s, e

    def tc_fwd(q, k, bits, inv_tau, flags, diag_offset=0):
        s, e = _e(q, k, inv_tau, flags, diag_offset)
        return e.sum(1), (s * bits).sum(1), e.sum(0)

    def tc_fwd_partials(q, k, bits, inv_tau, flags, diag_offset=0):
        s, e = _e(q, k, inv_tau, flags, diag_offset)
        # two row partials and three column partials, as the tiled kernel would produce
        half = e.shape[1] // 2
        rs = torch.stack([e[:, :half].sum(1), e[:, half:].sum(1)])
        rp = torch.stack([(s * bits)[:, :half].sum(1), (s * bits)[:, half:].sum(1)])
        third = max(e.shape[0] // 3, 1)
        cs = torch.stack([e[:third].sum(0), e[third:2 * third].sum(0), e[2 * third:].sum(0)])
        return rs, rp, cs

    def reduce_partials(part, parts, n, out=None, divisor=None):
        r = part[:parts, :n].sum(0)
        if divisor is not None:
            r = r / divisor
        if out is not None:
            out.copy_(r)
            return out
        return r

    def stats_fused(rs_part, rp_part, cs_part, counts, *, shift, pos_weight, inv_count, col_lo=0, col_hi=None):
        row_sum = rs_part if rs_part.dim() == 1 else rs_part.sum(0)
        row_pos = rp_part if rp_part.dim() == 1 else rp_part.sum(0)
        if counts is not None:
            row_pos = row_pos / counts
        acc = (shift + row_sum.log() - pos_weight * row_pos).sum()
        if cs_part is None:                                        # MPC: no column statistics
            return 1.0 / row_sum, None, (acc * inv_count).reshape(1)
        col_sum = cs_part if cs_part.dim() == 1 else cs_part.sum(0)
        col_hi = col_sum.shape[0] if col_hi is None else col_hi
        acc = acc + (shift + col_sum[col_lo:col_hi].log()).sum()
        return 1.0 / row_sum, 1.0 / col_sum, (acc * inv_count).reshape(1)

    def tc_bwd_w(q, k, bits, counts, a_row, b_col, inv_tau, flags, diag_offset=0):
        _, e = _e(q, k, inv_tau, flags, diag_offset)
        w = e * (a_row[:, None] + b_col[None, :]) - 2.0 * bits / counts[:, None]
        return w, None, w.shape[1]

    def tc_bwd_gemm(w_hi, w_lo, ld_w, n_rows, n_cols, transpose_w, x, flags, out=None):
        xm = x.hi + (x.lo if x.lo is not None else 0)
        return (w_hi.T @ xm) if transpose_w else (w_hi @ xm)

    def l2norm_bwd(x, nrm, g_hat, *, scale_dev, scale_host, gather=None):
        xs = x.detach().double().numpy()
        if gather is None:
            g = orc.l2_normalize_bwd(xs, g_hat.numpy())
        else:                                                      # filtered rows get zero gradient (:429)
            idx = gather.numpy()
            g = np.zeros_like(xs)
            g[idx] = orc.l2_normalize_bwd(xs[idx], g_hat.numpy())
        return (torch.from_numpy(g) * (scale_host * scale_dev.double().item())).to(x.dtype)

    for f in (l2norm_fwd, posmask_build, tc_fwd, tc_fwd_partials, reduce_partials, stats_fused, tc_bwd_w, tc_bwd_gemm,
              l2norm_bwd, rows_of):
        setattr(ops, f.__name__, f)
    return ops


def _worker(rank, world, port, n_total, d, tau, precision, mode, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evoke_b200 import synth
        from evoke_b200.distributed import global_alignment_sharded
        from evoke_b200.ids import DeviceIds
        ids = synth.make_study_ids(n_total, seed=21)
        xi = synth.make_embeddings(ids, d, seed=22).astype(np.float64)
        xt = synth.make_embeddings(ids, d, seed=23).astype(np.float64)
        n = n_total // world
        sl = slice(rank * n, (rank + 1) * n)
        image = torch.tensor(xi[sl], requires_grad=True)
        text = torch.tensor(xt[sl], requires_grad=True)
        row_ids = DeviceIds(torch.from_numpy(ids[sl].copy()))
        loss = global_alignment_sharded(image, text, row_ids, tau, precision=precision, mode=mode, ops=_oracle_ops())
        (2.5 * loss).backward()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.item(), d_image=image.grad.numpy(),
                 d_text=text.grad.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("precision,mode", [("bf16", "rs"), ("fp32", "rs"), ("bf16", "sym"), ("fp32", "sym")])
def test_sharded_choreography_equals_global_batch(tmp_path, precision, mode):
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 48, 24, 0.5
    port = 29700 + (os.getpid() % 200) + (1 if precision == "fp32" else 0) + (2 if mode == "sym" else 0)
    mp.spawn(_worker, args=(world, port, n_total, d, tau, precision, mode, str(tmp_path)), nprocs=world, join=True)
    ids = synth.make_study_ids(n_total, seed=21)
    xi = synth.make_embeddings(ids, d, seed=22).astype(np.float64)
    xt = synth.make_embeddings(ids, d, seed=23).astype(np.float64)
    want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    n = n_total // world
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert abs(float(got["loss"]) - want) < 1e-7 * abs(want)          # oracle labels are fp32-rounded (3e-8)
        sl = slice(r * n, (r + 1) * n)
        assert np.abs(got["d_image"] - 2.5 * d_i[sl]).max() < 1e-7 * np.abs(d_i).max() * 2.5
        assert np.abs(got["d_text"] - 2.5 * d_t[sl]).max() < 1e-7 * np.abs(d_t).max() * 2.5


def test_sharded_rejects_string_ids_and_uninitialised_group():
    from evoke_b200.distributed import global_alignment_sharded
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="not initialised"):
        global_alignment_sharded(x, x, np.arange(4), 0.5, ops=_oracle_ops())


def _mpc_worker(rank, world, port, n_total, d, tau, ids_kind, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from evoke_b200.distributed import multi_pos_contra_images_sharded
        ids = _mpc_ids(n_total, ids_kind)
        from evoke_b200 import synth
        x = synth.make_embeddings(ids, d, seed=42).astype(np.float64)
        m = n_total // world
        sl = slice(rank * m, (rank + 1) * m)
        xs = torch.tensor(x[sl], requires_grad=True)
        loss = multi_pos_contra_images_sharded(xs, ids[sl].copy(), tau, precision="fp32", ops=_oracle_ops())
        if loss.grad_fn is not None:
            (1.5 * loss).backward()
            grad = xs.grad.numpy()
        else:
            grad = np.zeros_like(x[sl])
        np.savez(os.path.join(out_dir, f"mpc{rank}.npz"), loss=loss.detach().numpy().reshape(-1), grad=grad, leaf=loss.grad_fn is None)
    finally:
        dist.destroy_process_group()


def _mpc_ids(n_total, kind):
    from evoke_b200 import synth
    if kind == "mixed":
        return synth.make_study_ids(n_total, seed=41)
    if kind == "single":                       # no study has a second view: the shape-[1] zero leaf (:427-428)
        return np.arange(n_total, dtype=np.int32)
    # "lopsided": every multi-view study lives on rank 0's rows only, rank 1 keeps nothing
    ids = np.arange(n_total, dtype=np.int32)
    ids[: n_total // 4] = ids[: n_total // 4] // 2
    return ids


@pytest.mark.parametrize("ids_kind", ["mixed", "lopsided", "single"])
def test_sharded_mpc_equals_global_batch(tmp_path, ids_kind):
    """multi_pos_contra_images_v0401 (:421-446) sharded over 2 ranks == the single-device reference on all views:
    cross-rank positives, rows dropped as queries AND keys, a rank without kept rows, the empty case."""
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 56, 24, 0.5
    port = 29500 + (os.getpid() % 150) + ["mixed", "lopsided", "single"].index(ids_kind) * 3
    mp.spawn(_mpc_worker, args=(world, port, n_total, d, tau, ids_kind, str(tmp_path)), nprocs=world, join=True)
    ids = _mpc_ids(n_total, ids_kind)
    x = synth.make_embeddings(ids, d, seed=42).astype(np.float64)
    want, dx = orc.mpc_closed_form(x, ids, tau)
    m = n_total // world
    for r in range(world):
        got = np.load(tmp_path / f"mpc{r}.npz")
        if want is None:
            assert bool(got["leaf"]) and got["loss"].shape == (1,) and float(got["loss"][0]) == 0.0
            continue
        assert abs(float(got["loss"][0]) - want) < 1e-9 * abs(want)
        assert np.abs(got["grad"] - 1.5 * dx[r * m:(r + 1) * m]).max() < 1e-7 * np.abs(dx).max() * 1.5
