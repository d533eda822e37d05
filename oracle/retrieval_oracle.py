"""CPU oracle of the retrieval row (f4).  TEST INFRASTRUCTURE ONLY.

Restates what ``train_index.search(x, k)`` means in PretrainTester.predict (reference
modules/multiview/trainer.py:543-653) for an EXACT inner-product index - faiss ``IndexFlatIP`` semantics: for every
query the k corpus rows of largest inner product, best first.  The reference itself uses ``faiss.IndexIVFFlat``
(trainer.py:549-550; nlist 100 / 40, one probed list), i.e. an approximation of this search whose result depends
on faiss's k-means (third-party, not vendored, not installed here: faiss is imported at trainer.py:12 with no
pinned version).  Parity is therefore pinned to the exact search ("parity unpinned" with respect to the IVF
approximation): every hit faiss-IVF returns is a true inner-product neighbour, and the exact list is what it
converges to as nprobe -> nlist.
"""
from __future__ import annotations

import numpy as np


def topk_inner_product(queries, corpus, k: int, query_groups=None, corpus_groups=None):
    """(scores [Q, k], indices [Q, k]) in fp64; ties broken by the lower corpus index (stable sort).  Corpus rows of the
    query's own group are skipped (trainer.py:590-607 removes hits of the query's own study)."""
    q = np.asarray(queries, dtype=np.float64)
    c = np.asarray(corpus, dtype=np.float64)
    s = q @ c.T
    if query_groups is not None:
        s = np.where(np.asarray(query_groups)[:, None] == np.asarray(corpus_groups)[None, :], -np.inf, s)
    order = np.argsort(-s, axis=1, kind="stable")[:, :k]
    val = np.take_along_axis(s, order, axis=1)
    idx = np.where(np.isfinite(val), order, -1)
    if idx.shape[1] < k:
        pad = k - idx.shape[1]
        idx = np.concatenate([idx, -np.ones((idx.shape[0], pad), dtype=idx.dtype)], axis=1)
        val = np.concatenate([val, -np.inf * np.ones((val.shape[0], pad))], axis=1)
    return val, idx
