"""T1: host-side logic (no GPU): id factorisation, row filter, path choice, loud failures."""
import numpy as np
import pytest
import torch

import evoke_b200
from evoke_b200 import functional as Fn
from evoke_b200 import ids as idmod
from evoke_b200 import synth
from oracle import evoke_oracle as orc


def test_factorize_preserves_equality_structure():
    rng = np.random.default_rng(0)
    raw = rng.integers(0, 50, size=300)
    strs = np.array([f"p{v // 3}_s{v}" for v in raw])
    for arr in (raw, strs, raw.astype(np.int64) * 10**12, -raw):
        codes = idmod.factorize(arr)
        assert codes.dtype == np.int32
        assert np.array_equal(codes[:, None] == codes[None, :], arr[:, None] == arr[None, :])


def test_combine_keys_is_conjunction():
    pat, stu = synth.make_patient_study_ids(500, seed=3)
    codes = idmod.combine_keys(pat, stu)
    want = (pat[:, None] == pat[None, :]) & (stu[:, None] == stu[None, :])
    assert np.array_equal(codes[:, None] == codes[None, :], want)
    ref_strings = np.array([f"p{p}_s{s}" for p, s in zip(pat, stu)])      # dataloaders_v0401.py:83
    assert np.array_equal(ref_strings[:, None] == ref_strings[None, :], want)


def test_multi_view_rows_matches_oracle_filter():
    for seed in range(5):
        ids = synth.make_study_ids(97, seed=seed)
        assert np.array_equal(idmod.multi_view_rows(ids), orc.mpc_kept_rows(ids))
    assert len(idmod.multi_view_rows(np.arange(10))) == 0


def test_int64_tensor_keys_are_exact():
    big = torch.tensor([2**40 + 1, 2**40 + 1, 1, 2**32 + 1, -5, -5], dtype=torch.int64)
    k = idmod._tensor_to_keys(big, torch.device("cpu"))
    eq = (k.key[:, None] == k.key[None, :]) & (k.key2[:, None] == k.key2[None, :])
    assert torch.equal(eq, big[:, None] == big[None, :])


def test_path_choice():
    assert Fn.choose_path("auto", 64, 64, 768) == "small"
    assert Fn.choose_path("auto", 16384, 16384, 768) == "tc"
    assert Fn.choose_path("tc", 8, 8, 16) == "tc"
    with pytest.raises(ValueError):
        Fn.choose_path("cpu", 8, 8, 16)


def test_cpu_tensors_fail_loudly():
    x = torch.randn(4, 8)
    with pytest.raises(RuntimeError, match="no CPU path"):
        evoke_b200.global_alignment(x, x, np.arange(4), 0.5)
    with pytest.raises(RuntimeError, match="no CPU path"):
        evoke_b200.multi_pos_contra_images(x, np.arange(4), 0.5)


def test_synth_is_deterministic_and_sized():
    a = synth.make_study_ids(1000, seed=5)
    b = synth.make_study_ids(1000, seed=5)
    assert np.array_equal(a, b) and len(a) == 1000
    sizes = np.unique(a, return_counts=True)[1]
    assert sizes.max() <= 4 and 1.8 < sizes.mean() < 2.6
    x = synth.make_embeddings(a, 32, seed=1)
    assert x.shape == (1000, 32) and x.dtype == np.float32


def test_lm_loss_boundary_signature():
    torch.manual_seed(0)
    logp = torch.log_softmax(torch.randn(3, 6, 11), -1)
    ids_ = torch.randint(0, 11, (3, 8))
    mask = (torch.rand(3, 8) > 0.3).float()
    got = evoke_b200.compute_lm_loss(logp, ids_, mask)
    tgt, msk = ids_[:, 1:][:, :6], mask[:, 1:][:, :6]
    want = -(logp.gather(2, tgt.unsqueeze(2)).squeeze(2) * msk).sum() / msk.sum()
    assert torch.allclose(got, want)
