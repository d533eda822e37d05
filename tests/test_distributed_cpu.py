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
    ops.FLAG_SPLIT_BF16 = 4

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
        xh, nrm = orc.l2_normalize(x.detach().double().numpy())
        hi = torch.from_numpy(xh)                                  # "hi" carries the exact unit rows here
        lo = torch.zeros_like(hi) if want_lo else None
        return Normalized(n=x.shape[0], d=x.shape[1], norm=torch.from_numpy(nrm.reshape(-1)), hi=hi, lo=lo, ld=x.shape[1])

    def posmask_build(rows, cols, *, clear_diag, diag_offset=0):
        m = orc.posmask_dense(rows.key.numpy(), cols.key.numpy(), clear_diag, diag_offset)
        return torch.from_numpy(m), torch.from_numpy(m.sum(1).astype(np.int32))

    def _e(q, k, inv_tau):
        s = (q.hi + (q.lo if q.lo is not None else 0)) @ (k.hi + (k.lo if k.lo is not None else 0)).T * inv_tau
        return s, torch.exp(s - inv_tau)

    def tc_fwd(q, k, bits, inv_tau, flags, diag_offset=0):
        s, e = _e(q, k, inv_tau)
        return e.sum(1), (s * bits).sum(1), e.sum(0)

    def tc_fwd_partials(q, k, bits, inv_tau, flags, diag_offset=0):
        s, e = _e(q, k, inv_tau)
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
        col_sum = cs_part if cs_part.dim() == 1 else cs_part.sum(0)
        col_hi = col_sum.shape[0] if col_hi is None else col_hi
        if counts is not None:
            row_pos = row_pos / counts
        acc = (shift + row_sum.log() - pos_weight * row_pos).sum()
        acc = acc + (shift + col_sum[col_lo:col_hi].log()).sum()
        return 1.0 / row_sum, 1.0 / col_sum, (acc * inv_count).reshape(1)

    def tc_bwd_w(q, k, bits, counts, a_row, b_col, inv_tau, flags, diag_offset=0):
        _, e = _e(q, k, inv_tau)
        w = e * (a_row[:, None] + b_col[None, :]) - 2.0 * bits / counts[:, None]
        return w, None, w.shape[1]

    def tc_bwd_gemm(w_hi, w_lo, ld_w, n_rows, n_cols, transpose_w, x, flags, out=None):
        xm = x.hi + (x.lo if x.lo is not None else 0)
        return (w_hi.T @ xm) if transpose_w else (w_hi @ xm)

    def l2norm_bwd(x, nrm, g_hat, *, scale_dev, scale_host, gather=None):
        g = orc.l2_normalize_bwd(x.detach().double().numpy(), g_hat.numpy())
        return (torch.from_numpy(g) * (scale_host * scale_dev.double().item())).to(x.dtype)

    for f in (l2norm_fwd, posmask_build, tc_fwd, tc_fwd_partials, reduce_partials, stats_fused, tc_bwd_w, tc_bwd_gemm,
              l2norm_bwd):
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
