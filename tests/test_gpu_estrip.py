"""GPU: the bf16-mode backward that keeps E = exp(S - 1/tau) from the forward (K3 with a bf16
strip store, then K4t in place) instead of recomputing the similarity tiles (K4a).

Checked: the strip against a torch fp32 evaluation of the same bf16 operands (bf16 rounding only),
its statistics against plain K3 (identical arithmetic), K4t against K4a, and loss + gradients through
the public API against the fp64 oracle for every backward variant (strip on/off, side-stream
overlap, CUDA graph).  Tolerance: gradients <= 2e-2 rel (BASELINE.json, bf16 mode).
"""
import numpy as np
import pytest
import torch

import evoke_b200
from evoke_b200 import functional as Fn
from evoke_b200 import synth
from evoke_b200.ids import to_device_ids
from gpu_util import DEV, TOL, rel_max
from oracle import evoke_oracle as orc

pytestmark = pytest.mark.gpu


def _operands(n_rows, n_cols, d, seed):
    n = max(n_rows, n_cols)
    ids = synth.make_study_ids(n, seed=seed)
    xq = torch.tensor(synth.make_embeddings(ids, d, seed=seed + 1)[:n_rows], device=DEV)
    xk = torch.tensor(synth.make_embeddings(ids, d, seed=seed + 2)[:n_cols], device=DEV)
    q = Fn.l2norm_fwd(xq, want_f32=False, want_hi=True, want_lo=False)
    k = Fn.l2norm_fwd(xk, want_f32=False, want_hi=True, want_lo=False)
    rid, _ = to_device_ids(ids[:n_rows], torch.device(DEV), n=n_rows)
    cid, _ = to_device_ids(ids[:n_cols], torch.device(DEV), n=n_cols)
    return q, k, rid, cid


@pytest.mark.parametrize("n_rows,n_cols,d,inv_tau,flags", [
    (256, 256, 64, 2.0, 0),
    (300, 777, 200, 2.0, 0),
    (1024, 1536, 768, 1.0 / 0.07, 0),
    (515, 515, 96, 2.0, Fn.FLAG_EXCLUDE_DIAG | Fn.FLAG_NO_COLSUM),
])
def test_strip_holds_exp_of_shifted_logits_and_statistics_match_plain_k3(n_rows, n_cols, d, inv_tau, flags):
    q, k, rid, cid = _operands(n_rows, n_cols, d, seed=n_rows + n_cols)
    excl = bool(flags & Fn.FLAG_EXCLUDE_DIAG)
    bits, counts = Fn.posmask_build(rid, cid, clear_diag=excl)
    rs0, rp0, cs0 = Fn.tc_fwd_partials(q, k, bits, inv_tau, flags)
    rs1, rp1, cs1, e, ld_e = Fn.tc_fwd_store(q, k, bits, inv_tau, flags)
    torch.cuda.synchronize()
    assert torch.equal(rs0, rs1) and torch.equal(rp0, rp1)
    assert (cs0 is None and cs1 is None) or torch.equal(cs0, cs1)
    s = (q.hi[:, :d].float() @ k.hi[:, :d].float().t()) * inv_tau
    want = torch.exp(s - inv_tau)
    if excl:
        want.fill_diagonal_(0.0)
    got = e[:, :n_cols].float()
    # bf16 storage (2^-9) + ex2.approx + fp32 accumulation-order differences in S (scaled by 1/tau)
    err = ((got - want).abs() / want.clamp_min(1e-30)).max().item() if not excl else \
        ((got - want).abs() / want.clamp_min(1e-30))[want > 0].max().item()
    assert err < 2.0 ** -8 + 4e-6 * inv_tau, err
    if excl:
        assert torch.all(got.diagonal() == 0)


@pytest.mark.parametrize("use_lists", [False, True], ids=["maskscan", "lists"])
@pytest.mark.parametrize("n_rows,n_cols,d", [(384, 640, 128), (1000, 1000, 768)])
def test_k4t_matches_k4a(n_rows, n_cols, d, use_lists):
    inv_tau = 2.0
    q, k, rid, cid = _operands(n_rows, n_cols, d, seed=7)
    bits, counts, pos_idx = Fn.posmask_build(rid, cid, clear_diag=False, want_list=True)
    pos = (pos_idx, Fn.pos_logits(q, k, pos_idx, counts)) if use_lists else None
    rs, rp, cs, e, ld_e = Fn.tc_fwd_store(q, k, bits, inv_tau, 0)
    a_row = 1.0 / Fn.reduce_partials(rs, int(rs.shape[0]), n_rows)
    b_col = 1.0 / Fn.reduce_partials(cs, int(cs.shape[0]), n_cols)
    w_ref, _, ld_w = Fn.tc_bwd_w(q, k, bits, counts, a_row, b_col, inv_tau, 0)
    # two row ranges (the entry point works on any row range of the strip)
    half = 256
    Fn.tc_w_from_e(e, ld_e, n_cols, bits, counts, a_row, b_col, q, k, inv_tau, 0, half, pos=pos)
    Fn.tc_w_from_e(e, ld_e, n_cols, bits, counts, a_row, b_col, q, k, inv_tau, half, n_rows - half, pos=pos)
    torch.cuda.synchronize()
    got, want = e[:, :n_cols].float(), w_ref[:, :n_cols].float()
    scale = want.abs().max().item()
    m = torch.zeros((n_rows, n_cols), dtype=torch.bool, device=DEV)
    ids_r, ids_c = rid.key.long(), cid.key.long()
    m = ids_r[:, None] == ids_c[None, :]
    # negatives: W from a bf16-rounded E, rounded again: 2 * 2^-9 relative
    assert ((got - want).abs()[~m] <= 2.0 ** -7 * want.abs()[~m] + 1e-9 * scale).all()
    # positives are recomputed from S in fp32 and rounded once, like K4a: one bf16 ulp at most
    assert ((got - want).abs()[m] <= 2.0 ** -8 * want.abs()[m] + 1e-6 * scale).all()


def _grads(n, d, tau, monkeypatch, *, strip, overlap=False, mask_free=True):
    monkeypatch.setattr(Fn, "E_STRIP", strip)
    monkeypatch.setattr(Fn, "OVERLAP_STREAMS", overlap)
    monkeypatch.setattr(Fn, "MASK_FREE", mask_free)
    ids = synth.make_study_ids(n, seed=n)
    xi = synth.make_embeddings(ids, d, seed=1)
    xt = synth.make_embeddings(ids, d, seed=2)
    image = torch.tensor(xi, device=DEV, requires_grad=True)
    text = torch.tensor(xt, device=DEV, requires_grad=True)
    out = evoke_b200.global_alignment(image, text, ids, tau, precision="bf16", path="tc")
    out.backward()
    x = torch.tensor(xi, device=DEV, requires_grad=True)
    outm = evoke_b200.multi_pos_contra_images(x, ids, tau, precision="bf16", path="tc")
    outm.backward()
    torch.cuda.synchronize()
    return (xi, xt, ids), out.item(), image.grad.cpu().numpy(), text.grad.cpu().numpy(), outm.item(), x.grad.cpu().numpy()


@pytest.mark.parametrize("variant", [dict(strip=False), dict(strip=True), dict(strip=True, overlap=True),
                                     dict(strip=True, mask_free=False), dict(strip=True, overlap=True, mask_free=False)],
                         ids=lambda v: "-".join(f"{k}{int(x)}" for k, x in v.items()))
@pytest.mark.parametrize("n,d,tau", [(4096, 768, 0.5), (2500, 512, 0.07)])
def test_backward_variants_against_oracle(n, d, tau, variant, monkeypatch):
    (xi, xt, ids), loss, d_i, d_t, lossm, d_x = _grads(n, d, tau, monkeypatch, **variant)
    want, w_i, w_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    tol = TOL["bf16"]
    assert abs(loss - want) <= tol["loss"] * abs(want)
    assert rel_max(d_i, w_i) <= tol["grad"]
    assert rel_max(d_t, w_t) <= tol["grad"]
    wantm, w_x = orc.mpc_closed_form(xi, ids, tau)
    assert abs(lossm - wantm) <= tol["loss"] * abs(wantm)
    assert rel_max(d_x, w_x) <= tol["grad"]


@pytest.mark.parametrize("mask_free", [True, False])
def test_rows_with_more_positives_than_list_slots_fall_back_to_the_mask(monkeypatch, mask_free):
    """Groups of 12 views (> POS_SLOTS) next to ordinary ones: list path and the overflow path (id scan in mask-free
    mode, mask scan otherwise) in one launch - forward positive sums and exact W entries."""
    monkeypatch.setattr(Fn, "E_STRIP", True)
    monkeypatch.setattr(Fn, "MASK_FREE", mask_free)
    n, d, tau = 1536, 128, 0.2
    rng = np.random.default_rng(3)
    ids = np.concatenate([np.arange(600) // 12, 1000 + np.arange(n - 600) // 2]).astype(np.int32)
    ids = ids[rng.permutation(n)]
    assert Fn.POS_SLOTS < 12
    xi = synth.make_embeddings(ids, d, seed=1)
    xt = synth.make_embeddings(ids, d, seed=2)
    image = torch.tensor(xi, device=DEV, requires_grad=True)
    text = torch.tensor(xt, device=DEV, requires_grad=True)
    out = evoke_b200.global_alignment(image, text, ids, tau, precision="bf16", path="tc")
    out.backward()
    want, w_i, w_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    assert abs(out.item() - want) <= TOL["bf16"]["loss"] * abs(want)
    assert rel_max(image.grad.cpu().numpy(), w_i) <= TOL["bf16"]["grad"]
    assert rel_max(text.grad.cpu().numpy(), w_t) <= TOL["bf16"]["grad"]
    x = torch.tensor(xi, device=DEV, requires_grad=True)              # MPC: diagonal excluded in the id scan too
    outm = evoke_b200.multi_pos_contra_images(x, ids, tau, precision="bf16", path="tc")
    outm.backward()
    wantm, w_x = orc.mpc_closed_form(xi, ids, tau)
    assert abs(outm.item() - wantm) <= TOL["bf16"]["loss"] * abs(wantm)
    assert rel_max(x.grad.cpu().numpy(), w_x) <= TOL["bf16"]["grad"]


def test_strip_and_recompute_backwards_agree_closely(monkeypatch):
    a = _grads(3000, 768, 0.5, monkeypatch, strip=False)
    b = _grads(3000, 768, 0.5, monkeypatch, strip=True, overlap=True)
    # same exp-sums; the positive sums come from the tensor-core accumulators (recompute mode: K3 epilogue) resp. from
    # fp32 dot products of the same bf16 operands (mask-free strip mode): equal to fp32 rounding
    assert abs(a[1] - b[1]) <= 1e-6 * abs(a[1]) and abs(a[4] - b[4]) <= 1e-6 * abs(a[4])
    for ga, gb in ((a[2], b[2]), (a[3], b[3]), (a[5], b[5])):
        assert rel_max(gb, ga) <= 4e-3                       # both carry bf16 W; they differ by one more rounding


def test_second_backward_over_a_consumed_strip_is_an_error(monkeypatch):
    monkeypatch.setattr(Fn, "E_STRIP", True)
    ids = synth.make_study_ids(1024, seed=5)
    x = torch.tensor(synth.make_embeddings(ids, 128, seed=3), device=DEV, requires_grad=True)
    y = torch.tensor(synth.make_embeddings(ids, 128, seed=4), device=DEV, requires_grad=True)
    out = evoke_b200.global_alignment(x, y, ids, 0.5, precision="bf16", path="tc")
    out.backward(retain_graph=True)
    with pytest.raises(RuntimeError, match="E strip"):
        out.backward()


def test_strip_backward_inside_a_cuda_graph(monkeypatch):
    monkeypatch.setattr(Fn, "E_STRIP", True)
    n, d = 4096, 256
    g = evoke_b200.GraphedGlobalAlignment(n, d, 0.5, precision="bf16", path="tc").capture()
    for seed in (1, 2):
        ids = synth.make_study_ids(n, seed=seed)
        xi = synth.make_embeddings(ids, d, seed=seed + 10)
        xt = synth.make_embeddings(ids, d, seed=seed + 20)
        g.load(torch.tensor(xi, device=DEV), torch.tensor(xt, device=DEV), torch.from_numpy(ids).to(DEV))
        loss = g.step()
        want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, 0.5)
        assert abs(loss.item() - want) <= TOL["bf16"]["loss"] * abs(want)
        assert rel_max(g.image.grad.cpu().numpy(), d_i) <= TOL["bf16"]["grad"]
        assert rel_max(g.text.grad.cpu().numpy(), d_t) <= TOL["bf16"]["grad"]
