"""T2 (GPU): K1 / K2 / statistics kernels through the C ABI against the oracle."""
import numpy as np
import pytest
import torch

from evoke_b200 import functional as Fn
from evoke_b200 import ids as idmod
from evoke_b200 import synth
from gpu_util import DEV, rel_max
from oracle import evoke_oracle as orc

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------- K1
@pytest.mark.parametrize("n,d", [(1, 8), (37, 768), (64, 2048), (33, 100), (300, 512), (5, 4100)])
def test_l2norm_fwd_matches_normalize(n, d):
    torch.manual_seed(n * 1000 + d)
    x = torch.randn(n, d, device=DEV) * 3.0
    if n > 2:
        x[2] = 0.0                                                   # eps clamp row
    out = Fn.l2norm_fwd(x, want_f32=True, want_hi=True, want_lo=True)
    ref = torch.nn.functional.normalize(x, dim=-1, p=2)
    # <= 2 ulp of fp32 (summation order of the norm differs from ATen's)
    assert torch.allclose(out.f32, ref, rtol=3e-7, atol=1e-37)
    assert torch.allclose(out.norm, x.norm(dim=-1), rtol=1e-6)
    hi = out.hi[:, :d].float()
    lo = out.lo[:, :d].float()
    assert torch.equal(out.hi[:, :d], out.f32.to(torch.bfloat16))     # hi is the RN bf16 of xhat
    assert (hi + lo - out.f32).abs().max().item() <= 2.0 ** -16       # split keeps ~16 mantissa bits of a unit vector


def test_l2norm_fwd_strided_view_like_the_projection_head():
    # the reference slices [:,0,:] of a permuted [B, D, 1+P] head output: strides (D*(1+P), 1+P)
    torch.manual_seed(0)
    b, d, p1 = 48, 768, 50
    head = torch.randn(b, d, p1, device=DEV)
    x = head.permute(0, 2, 1)[:, 0, :]
    assert x.stride() == (d * p1, p1)
    out = Fn.l2norm_fwd(x, want_f32=True, want_hi=False, want_lo=False)
    assert torch.allclose(out.f32, torch.nn.functional.normalize(x, dim=-1), rtol=3e-7, atol=1e-37)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_l2norm_fwd_half_inputs_and_gather(dtype):
    torch.manual_seed(1)
    x = torch.randn(40, 256, device=DEV).to(dtype)
    gather = torch.tensor([3, 7, 7, 39, 0], device=DEV, dtype=torch.int32)
    out = Fn.l2norm_fwd(x, want_f32=True, want_hi=False, want_lo=False, gather=gather)
    ref = torch.nn.functional.normalize(x.float()[gather.long()], dim=-1)
    assert torch.allclose(out.f32, ref, rtol=3e-7, atol=1e-37)


def test_l2norm_bwd_matches_autograd_including_clamped_row():
    torch.manual_seed(2)
    x = torch.randn(19, 96, device=DEV)
    x[5] = 0.0
    g = torch.randn(19, 96, device=DEV)
    scale = torch.tensor([0.37], device=DEV)
    nrm = Fn.l2norm_fwd(x, want_f32=False, want_hi=False, want_lo=False)
    dx = Fn.l2norm_bwd(x, nrm, g, scale_dev=scale, scale_host=2.0)
    want = orc.l2_normalize_bwd(x.double().cpu().numpy(), g.double().cpu().numpy()) * (0.37 * 2.0)
    assert rel_max(dx.cpu().numpy(), want) < 2e-6


# ------------------------------------------------------------------------------------- K2
@pytest.mark.parametrize("n", [1, 31, 32, 33, 255, 256, 257, 1000, 4100])
@pytest.mark.parametrize("clear_diag", [False, True])
def test_posmask_bit_exact(n, clear_diag):
    ids = synth.make_study_ids(n, seed=n)
    dev = idmod.DeviceIds(torch.from_numpy(ids).to(DEV))
    bits, counts = Fn.posmask_build(dev, dev, clear_diag=clear_diag)
    want_bits, want_counts = orc.posmask_packed(ids, clear_diag=clear_diag)
    got = bits.cpu().numpy().view(np.uint32)
    w = want_bits.shape[1]
    assert np.array_equal(got[:, :w], want_bits)
    assert not got[:, w:].any()                                       # pad words are zero
    assert np.array_equal(counts.cpu().numpy(), want_counts)


def test_posmask_rectangular_block_with_offset_and_two_keys():
    pat, stu = synth.make_patient_study_ids(700, seed=4)
    rows = slice(256, 420)
    dr = idmod.DeviceIds(torch.from_numpy(pat[rows].copy()).to(DEV), torch.from_numpy(stu[rows].copy()).to(DEV))
    dc = idmod.DeviceIds(torch.from_numpy(pat).to(DEV), torch.from_numpy(stu).to(DEV))
    bits, counts = Fn.posmask_build(dr, dc, clear_diag=True, diag_offset=256)
    key = idmod.combine_keys(pat, stu)
    want_bits, want_counts = orc.posmask_packed(key[rows], key, clear_diag=True, row_offset=256)
    got = bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[:, : want_bits.shape[1]], want_bits)
    assert np.array_equal(counts.cpu().numpy(), want_counts)


@pytest.mark.parametrize("clear_diag", [False, True])
def test_posmask_positive_lists_are_the_sparse_form_of_the_mask(clear_diag):
    """pos_idx[r, :min(c_r, slots)] = the set bits of row r (any order); degenerate ids (one big group) and
    a long row of equal keys exercise duplicate-heavy hash chains."""
    rng = np.random.default_rng(5)
    ids = np.concatenate([synth.make_study_ids(3000, seed=9), np.full(40, 777777, dtype=np.int32)])
    ids = ids[rng.permutation(len(ids))]
    dev = idmod.DeviceIds(torch.from_numpy(ids).to(DEV))
    bits, counts, pos_idx = Fn.posmask_build(dev, dev, clear_diag=clear_diag, want_list=True)
    want_bits, want_counts = orc.posmask_packed(ids, clear_diag=clear_diag)
    got = bits.cpu().numpy().view(np.uint32)
    assert np.array_equal(got[:, : want_bits.shape[1]], want_bits)
    cnt = counts.cpu().numpy()
    assert np.array_equal(cnt, want_counts)
    dense = np.unpackbits(want_bits.view(np.uint8), axis=1, bitorder="little")[:, : len(ids)]
    lists = pos_idx.cpu().numpy()
    for r in range(len(ids)):
        k = min(int(cnt[r]), Fn.POS_SLOTS)
        cols = lists[r, :k]
        assert len(set(cols.tolist())) == k and dense[r, cols].all()
        if cnt[r] <= Fn.POS_SLOTS:
            assert set(cols.tolist()) == set(np.nonzero(dense[r])[0].tolist())


def test_posmask_string_ids_equal_int_ids():
    ids = synth.make_study_ids(130, seed=8)
    d1, _ = idmod.to_device_ids(synth.ids_as_strings(ids), torch.device(DEV))
    d2, _ = idmod.to_device_ids(ids, torch.device(DEV))
    b1, c1 = Fn.posmask_build(d1, d1, clear_diag=False)
    b2, c2 = Fn.posmask_build(d2, d2, clear_diag=False)
    assert torch.equal(b1, b2) and torch.equal(c1, c2)


# ------------------------------------------------------------------------------------- stats
def test_reduce_partials_and_finalize():
    torch.manual_seed(3)
    part = torch.rand(7, 1000, device=DEV)
    got = Fn.reduce_partials(part, 7, 1000)
    assert torch.allclose(got, part.sum(0), rtol=1e-6)
    n = 333
    rs = torch.rand(n, device=DEV) + 0.5
    rp = torch.randn(n, device=DEV)
    cs = torch.rand(n, device=DEV) + 0.5
    cnt = torch.randint(1, 4, (n,), device=DEV, dtype=torch.int32)
    a, b, loss = Fn.finalize(rs, rp, cnt, cs, col_lo=0, col_hi=n, shift=2.0, pos_weight=2.0, inv_count=0.5 / n)
    want = (0.5 / n) * ((2.0 + rs.double().log() - 2.0 * rp.double() / cnt.double()).sum() + (2.0 + cs.double().log()).sum())
    assert abs(loss.item() - want.item()) < 1e-5 * abs(want.item())
    assert torch.allclose(a, 1.0 / rs) and torch.allclose(b, 1.0 / cs)


def test_pos_kernel_matches_mask_times_similarity():
    """evk_mpce_pos (optional O(N*D) form of the positive-logit sums) against a dense evaluation."""
    n, d = 700, 264
    ids = synth.make_study_ids(n, seed=3)
    x = torch.tensor(synth.make_embeddings(ids, d, seed=4), device=DEV)
    y = torch.tensor(synth.make_embeddings(ids, d, seed=5), device=DEV)
    dev = idmod.DeviceIds(torch.from_numpy(ids).to(DEV))
    bits, _ = Fn.posmask_build(dev, dev, clear_diag=True)
    for split in (False, True):
        q = Fn.l2norm_fwd(x, want_f32=False, want_hi=True, want_lo=split)
        k = Fn.l2norm_fwd(y, want_f32=False, want_hi=True, want_lo=split)
        got = Fn.tc_pos(q, k, bits, 2.0).cpu().numpy()
        qf = q.hi[:, :d].double() + (q.lo[:, :d].double() if split else 0)
        kf = k.hi[:, :d].double() + (k.lo[:, :d].double() if split else 0)
        m = torch.from_numpy(orc.posmask_dense(ids, clear_diag=True)).to(DEV)
        want = ((qf @ kf.t()) * 2.0 * m).sum(1).cpu().numpy()
        assert rel_max(got, want) < (2e-6 if not split else 2e-5)
