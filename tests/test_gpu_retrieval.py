"""f4 (GPU): exact inner-product top-k (evoke_b200.retrieval) against the numpy oracle (oracle/retrieval_oracle.py),
which restates the search of PretrainTester.predict (reference modules/multiview/trainer.py:543-653) for an exact
index.  Indices must be identical wherever the oracle's score gap to the next candidate exceeds the arithmetic's
resolution; scores within 1e-5 (fp32 mode) / 1e-2 (bf16 mode) relative."""
import numpy as np
import pytest
import torch

from evoke_b200 import retrieval
from oracle import retrieval_oracle as ro

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _data(n_q, n_c, d, seed):
    rng = np.random.default_rng(seed)
    centers = rng.standard_normal((max(n_c // 7, 2), d)).astype(np.float32)
    corpus = (centers[rng.integers(0, len(centers), n_c)] + 0.7 * rng.standard_normal((n_c, d))).astype(np.float32)
    queries = (centers[rng.integers(0, len(centers), n_q)] + 0.7 * rng.standard_normal((n_q, d))).astype(np.float32)
    return queries, corpus


def _check(val, idx, want_val, want_idx, all_scores, rtol):
    val, idx = val.cpu().numpy(), idx.cpu().numpy()
    scale = np.abs(want_val[np.isfinite(want_val)]).max()
    fin = np.isfinite(want_val)
    assert np.array_equal(np.isfinite(val), fin)
    assert np.abs(val[fin] - want_val[fin]).max() <= rtol * scale
    assert np.array_equal(idx[~fin], want_idx[~fin])
    # positions whose score is separated from both neighbours in the ranking by more than the resolution must agree
    gap_ok = np.ones_like(fin)
    srt = -np.sort(-all_scores, axis=1)[:, : want_val.shape[1] + 1]
    d_next = np.abs(srt[:, :-1] - srt[:, 1:])
    d_prev = np.concatenate([np.full((srt.shape[0], 1), np.inf), d_next[:, :-1]], axis=1)
    gap_ok = fin & (np.nan_to_num(d_next, nan=np.inf) > 4 * rtol * scale) & (np.nan_to_num(d_prev, nan=np.inf) > 4 * rtol * scale)
    if rtol <= 1e-4:
        assert gap_ok.mean() > 0.5               # (in bf16 mode most neighbouring scores are closer than its resolution)
    assert np.array_equal(idx[gap_ok], want_idx[gap_ok])
    # and every returned index is a genuine top-k member up to the resolution
    kth = want_val[:, -1:]
    got_scores = np.take_along_axis(all_scores, np.maximum(idx, 0), axis=1)
    assert (got_scores[fin] >= (kth - 4 * rtol * scale).repeat(want_val.shape[1], 1)[fin]).all()


@pytest.mark.parametrize("precision,rtol", [("fp32", 1e-5), ("bf16", 1e-2)])
@pytest.mark.parametrize("n_q,n_c,d,k", [(37, 1000, 96, 5), (300, 5000, 768, 13), (130, 700, 1544, 40)])
def test_topk_matches_the_exact_search(n_q, n_c, d, k, precision, rtol):
    queries, corpus = _data(n_q, n_c, d, seed=n_q + n_c)
    val, idx = retrieval.topk_inner_product(torch.tensor(queries, device=DEV), torch.tensor(corpus, device=DEV), k,
                                            precision=precision, chunk_q=128, chunk_c=1024)       # several chunks each way
    want_val, want_idx = ro.topk_inner_product(queries, corpus, k)
    _check(val, idx, want_val, want_idx, queries.astype(np.float64) @ corpus.astype(np.float64).T, rtol)


def test_own_study_is_not_a_candidate_and_short_lists_are_padded():
    n_q, n_c, d, k = 64, 48, 64, 20
    queries, corpus = _data(n_q, n_c, d, seed=3)
    qg = np.arange(n_q) % 6
    cg = np.arange(n_c) % 3                       # groups 0..2: a query of group g < 3 loses a third of the corpus
    cg[:40] = 0                                   # ... and group 0 nearly all of it (fewer than k candidates left)
    val, idx = retrieval.topk_inner_product(torch.tensor(queries, device=DEV), torch.tensor(corpus, device=DEV), k,
                                            query_groups=torch.tensor(qg, device=DEV), corpus_groups=torch.tensor(cg, device=DEV))
    want_val, want_idx = ro.topk_inner_product(queries, corpus, k, qg, cg)
    s = np.where(qg[:, None] == cg[None, :], -np.inf, queries.astype(np.float64) @ corpus.astype(np.float64).T)
    _check(val, idx, want_val, want_idx, s, 1e-5)
    idx = idx.cpu().numpy()
    assert (idx[qg == 0] == -1).sum() > 0
    hit = idx >= 0
    assert not (cg[np.maximum(idx, 0)][hit] == np.repeat(qg[:, None], k, 1)[hit]).any()


def test_faiss_style_index_on_the_reference_feature_shape():
    """d = 50 tokens x output_dim 768 = 38400 (trainer.py:545), the calls of predict(): train / add (two halves) / search."""
    d, k = 50 * 768, 13
    queries, corpus = _data(24, 320, d, seed=9)
    index = retrieval.FlatIPIndex(d)
    index.train(corpus[:160])
    index.add(corpus[:160])
    index.add(corpus[160:])
    assert index.ntotal == 320
    dist, ind = index.search(queries, k)
    want_val, want_idx = ro.topk_inner_product(queries, corpus, k)
    assert dist.shape == ind.shape == (24, k)
    # 38400-term dot products: the tensor cores accumulate in fp32 with truncation, 7200 accumulation steps deep here,
    # which biases every score low by ~3e-4 relative (measured; the ranking is unaffected: the bias is proportional)
    _check(torch.tensor(dist), torch.tensor(ind), want_val, want_idx, queries.astype(np.float64) @ corpus.astype(np.float64).T, 1e-3)
    assert (ind == want_idx).mean() >= 0.95          # the rest are near-ties inside that resolution (checked by _check)
