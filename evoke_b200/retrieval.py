"""f4: exact inner-product top-k retrieval on the B200.

The reference builds its "patient-specific knowledge" by searching, for every image, the most similar TRAIN
images by inner product of the flattened [50 x output_dim] token features, with a faiss ``IndexIVFFlat``
(``METRIC_INNER_PRODUCT``, nlist 100 / 40, default nprobe) - ``PretrainTester.predict``,
modules/multiview/trainer.py:543-653.  faiss is a third-party dependency that the reference does not vendor
(``import faiss``, no pinned version in README.md:118-123) and that is absent from this image.

Here the search is EXACT (what ``faiss.IndexFlatIP`` returns; the IVF index with one probed list is an approximation
of it): the score matrix is the same dense contraction as the loss's similarity (bf16 operands, fp32 accumulate, the
tcgen05 main loop of csrc/tc_engine.cu), computed in [query chunk x corpus chunk] blocks that a streaming kernel
(csrc/k_topk.cu) folds into every query's running top-k.  ``FlatIPIndex`` mirrors the four faiss calls the
reference makes (``train``, ``add``, ``search``, ``ntotal``), so the binding is one line in trainer.py:549-550.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _lib
from . import functional as Fn


def _split_bf16(x: torch.Tensor, want_lo: bool):
    """fp32 [n, d] -> bf16 (hi, lo | None) padded to 16-byte rows (a cast, not arithmetic of the search)."""
    n, d = x.shape
    ld = Fn._round_up(d, 8)
    hi = torch.zeros((n, ld), dtype=torch.bfloat16, device=x.device)
    hi[:, :d] = x.to(torch.bfloat16)
    lo = None
    if want_lo:
        lo = torch.zeros((n, ld), dtype=torch.bfloat16, device=x.device)
        lo[:, :d] = (x - hi[:, :d].to(torch.float32)).to(torch.bfloat16)
    return hi, lo, ld


def topk_inner_product(queries: torch.Tensor, corpus: torch.Tensor, k: int, *, query_groups: Optional[torch.Tensor] = None,
                       corpus_groups: Optional[torch.Tensor] = None, precision: str = "fp32",
                       chunk_q: int = 4096, chunk_c: int = 16384):
    """(scores [Q, k] fp32, indices [Q, k] int64): the k corpus rows with the largest inner product per query, sorted
    by score descending (ties: lower index first).  query_groups / corpus_groups (int tensors): corpus rows of the
    query's own group are skipped (the reference removes hits of the query's own study, trainer.py:590-607).
    precision "fp32": 3-segment split-bf16 operands (products exact to ~2^-17; the tensor cores' fp32 accumulation
    truncates, which biases a d = 38400 score low by ~3e-4 relative - proportionally for every candidate, so the
    ranking is that of the exact search); "bf16": single-pass bf16 operands.
    Slots beyond the number of candidates hold -inf / -1."""
    if not (queries.is_cuda and corpus.is_cuda):
        raise RuntimeError("evoke_b200.retrieval runs on a CUDA (sm_100a) device only; there is no CPU path")
    if queries.dim() != 2 or corpus.dim() != 2 or queries.shape[1] != corpus.shape[1]:
        raise ValueError(f"queries / corpus must be [*, d] with equal d, got {tuple(queries.shape)} and {tuple(corpus.shape)}")
    if not 1 <= k <= 64:
        raise ValueError("k must be in 1..64")
    if precision not in ("fp32", "bf16"):
        raise ValueError("precision must be 'fp32' or 'bf16'")
    if (query_groups is None) != (corpus_groups is None):
        raise ValueError("query_groups and corpus_groups must be given together")
    n_q, d = int(queries.shape[0]), int(queries.shape[1])
    n_c = int(corpus.shape[0])
    dev = queries.device
    split = precision == "fp32"
    with torch.cuda.device(dev):
        qg = cg = None
        if query_groups is not None:
            qg = query_groups.to(device=dev, dtype=torch.int32).contiguous()
            cg = corpus_groups.to(device=dev, dtype=torch.int32).contiguous()
        best_val = torch.empty((n_q, k), dtype=torch.float32, device=dev)
        best_idx = torch.empty((n_q, k), dtype=torch.int32, device=dev)
        c_hi, c_lo, ld_c = _split_bf16(corpus.to(torch.float32), split)
        stream = torch.cuda.current_stream().cuda_stream
        for q0 in range(0, n_q, chunk_q):
            q1 = min(n_q, q0 + chunk_q)
            q_hi, q_lo, ld_q = _split_bf16(queries[q0:q1].to(torch.float32), split)
            for c0 in range(0, n_c, chunk_c):
                c1 = min(n_c, c0 + chunk_c)
                ld_s = Fn._round_up(c1 - c0, 4)
                scores = torch.empty((q1 - q0, ld_s), dtype=torch.float32, device=dev)
                _lib.call("evk_tc_gemm_nt", q_hi.data_ptr(), None if q_lo is None else q_lo.data_ptr(), ld_q,
                          c_hi[c0:c1].data_ptr(), None if c_lo is None else c_lo[c0:c1].data_ptr(), ld_c,
                          q1 - q0, c1 - c0, d, scores.data_ptr(), ld_s, stream)
                _lib.call("evk_topk_update", scores.data_ptr(), ld_s, q1 - q0, c1 - c0, c0,
                          None if qg is None else qg[q0:q1].data_ptr(), None if cg is None else cg.data_ptr(), k,
                          best_val[q0:q1].data_ptr(), best_idx[q0:q1].data_ptr(), 1 if c0 == 0 else 0, stream)
        return best_val, best_idx.to(torch.int64)


class FlatIPIndex:
    """The subset of the faiss index interface PretrainTester.predict uses (trainer.py:549-550, 566-567, 588, 619,
    636): ``train`` (no-op: the search is exact), ``add``, ``search`` -> (D, I) numpy arrays, ``ntotal``.
    Vectors are kept on the device."""

    def __init__(self, d: int, device=None, precision: str = "fp32"):
        self.d = int(d)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.precision = precision
        self._blocks = []
        self._corpus: Optional[torch.Tensor] = None
        self.groups: Optional[torch.Tensor] = None
        self.is_trained = True

    @property
    def ntotal(self) -> int:
        return sum(int(b.shape[0]) for b in self._blocks)

    def train(self, x) -> None:                        # an exact index has nothing to train
        return None

    def add(self, x) -> None:
        t = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(self.device, torch.float32)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"expected [n, {self.d}] vectors, got {tuple(t.shape)}")
        self._blocks.append(t)
        self._corpus = None

    def search(self, x, k: int, query_groups=None):
        if not self._blocks:
            raise RuntimeError("search on an empty index")
        if self._corpus is None:
            self._corpus = torch.cat(self._blocks, dim=0) if len(self._blocks) > 1 else self._blocks[0]
            self._blocks = [self._corpus]
        q = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(self.device, torch.float32)
        qg = None if query_groups is None else torch.as_tensor(np.asarray(query_groups)).to(self.device)
        val, idx = topk_inner_product(q, self._corpus, k, query_groups=qg,
                                      corpus_groups=self.groups if qg is not None else None, precision=self.precision)
        return val.cpu().numpy(), idx.cpu().numpy()
