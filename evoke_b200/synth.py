"""Seeded synthetic workloads for the contrastive hot path.

Study sizes follow the Multi-view CXR statistics the reference trains on
(results/dataset-statistics.png: ~2.2 views per study); embeddings are
``x = g[study] + 0.5*eps`` so positives are genuinely closer than negatives and
the softmax is not uniform.  Everything is numpy (PCG64) so the same arrays can be
regenerated bit-identically in the golden-vector script, the tests and bench.py.
"""
from __future__ import annotations

import numpy as np

# study-size distributions named in SURVEY.md §8(d)
SIZES_CFG2 = {1: 0.45, 2: 0.45, 3: 0.08, 4: 0.02}
SIZES_CFG3 = {1: 0.25, 2: 0.45, 3: 0.20, 4: 0.10}


def make_study_ids(n: int, sizes: dict[int, float] | None = None, seed: int = 1234,
                   shuffle: bool = True) -> np.ndarray:
    """int32 study id per row; study sizes drawn iid from ``sizes`` until n rows."""
    sizes = sizes or SIZES_CFG3
    rng = np.random.Generator(np.random.PCG64(seed))
    ks = np.array(sorted(sizes), dtype=np.int64)
    ps = np.array([sizes[k] for k in ks], dtype=np.float64)
    ps = ps / ps.sum()
    # draw more studies than can possibly be needed, then cut at n rows
    draw = rng.choice(ks, size=n, p=ps)
    ids = np.repeat(np.arange(n, dtype=np.int64), draw)[:n]
    if shuffle:
        ids = ids[rng.permutation(n)]
    return ids.astype(np.int32)


def make_patient_study_ids(n: int, seed: int = 1234) -> tuple[np.ndarray, np.ndarray]:
    """(patient_id, study_id) int32 arrays: 1-3 studies per patient, 1-4 views per study
    (cfg4).  The reference's key is the conjunction "p<subject>_s<study>"
    (modules/dataloaders_v0401.py:83)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pat, stu = [], []
    p = s = 0
    while len(pat) < n:
        for _ in range(int(rng.integers(1, 4))):
            v = int(rng.integers(1, 5))
            pat += [p] * v
            stu += [s] * v
            s += 1
        p += 1
    pat = np.asarray(pat[:n], dtype=np.int32)
    stu = np.asarray(stu[:n], dtype=np.int32)
    perm = rng.permutation(n)
    return pat[perm], stu[perm]


def make_embeddings(ids: np.ndarray, d: int, seed: int = 1234, noise: float = 0.5,
                    dtype=np.float32) -> np.ndarray:
    """x = g[study] + noise*eps with g, eps ~ N(0,1)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    _, inv = np.unique(ids, return_inverse=True)
    g = rng.standard_normal((int(inv.max()) + 1, d), dtype=np.float32)
    eps = rng.standard_normal((len(ids), d), dtype=np.float32)
    return (g[inv] + noise * eps).astype(dtype)


def ids_as_strings(ids: np.ndarray) -> np.ndarray:
    """The reference's id type: numpy unicode array (dataloaders_v0401.py:83,115)."""
    return np.array([f"p{int(i) // 3}_s{int(i)}" for i in ids])
