"""Golden-vector case table shared by oracle/make_golden.py (which RUNS THE REFERENCE in the
build container and writes tests/golden/*.npz) and by the tests (which rebuild the same
seeded inputs and compare).  Inputs are regenerated from seeds, never stored, so the
fixtures stay small; outputs are stored in full for small cases and as sampled rows +
norms for large ones.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from evoke_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL_GRAD_LIMIT = 64 * 1024      # elements; above this only SAMPLE_ROWS rows are stored
SAMPLE_ROWS = 16


@dataclass(frozen=True)
class Case:
    name: str
    kind: str                    # "G" (global_alignment_loss) | "MPC" (multi_pos_contra_images_v0401)
                                 # "AG" / "AMPC": the PretrainNewMulPos variants (:748-815, v0404 :670-708)
    n: int
    d: int
    tau: float = 0.5
    ids: str = "cfg3"            # recipe, see build_ids
    seed: int = 1234
    string_ids: bool = False     # pass numpy <U ids, as the reference's dataloader does
    extra_ids: int = 0           # G only: patient_ids longer than B (aux views), :488 truncation
    zero_row: int = -1           # row index forced to all-zero (normalize eps clamp)
    tags: tuple = field(default_factory=tuple)


CASES = [
    # (i) hand-checkable, groups of size 1/2/3
    Case("g_n8_d16", "G", 8, 16, ids="hand8"),
    Case("mpc_n8_d16", "MPC", 8, 16, ids="hand8"),
    # (ii) BASELINE cfg1: 32 two-view studies, D=768; MPC over 64 rows, G over the 32 anchors
    Case("g_cfg1", "G", 32, 768, ids="twoview", extra_ids=32, string_ids=True),
    Case("mpc_cfg1", "MPC", 64, 768, ids="twoview", string_ids=True),
    # (iii) shuffled groups of 1-4
    Case("g_n1024_d768", "G", 1024, 768, ids="cfg3"),
    Case("mpc_n1024_d768", "MPC", 1024, 768, ids="cfg3"),
    Case("g_n300_d512_t007", "G", 300, 512, tau=0.07, ids="cfg2"),       # ragged N, cold temperature
    Case("mpc_n300_d512_t007", "MPC", 300, 512, tau=0.07, ids="cfg2"),
    Case("g_n257_d2048", "G", 257, 2048, ids="cfg3", seed=7),            # reference output_dim
    # (iv) MPC with no multi-view study: shape-[1] zero leaf
    Case("mpc_all_single", "MPC", 16, 64, ids="unique"),
    # (v) ids longer than B
    Case("g_extra_ids", "G", 24, 96, ids="cfg3", extra_ids=11, seed=5),
    # (vi) zero-norm row
    Case("g_zero_row", "G", 16, 64, ids="cfg2", zero_row=3, seed=9),
    Case("mpc_zero_row", "MPC", 16, 64, ids="twoview", zero_row=3, seed=9),
    # (vii) string ids == int ids
    Case("g_strings", "G", 64, 128, ids="cfg3", string_ids=True, seed=11),
    # all rows one study (every pair positive)
    Case("g_all_same", "G", 12, 32, ids="same"),
    Case("mpc_all_same", "MPC", 12, 32, ids="same"),
]

# 'averaged positive logit' variants of PretrainNewMulPos
AVGPOS_CASES = [
    Case("ag_n8_d16", "AG", 8, 16, ids="hand8"),
    Case("ag_cfg1", "AG", 32, 768, ids="twoview", extra_ids=32, string_ids=True),   # all single-positive rows
    Case("ag_n96_d128", "AG", 96, 128, ids="cfg3", seed=21),
    Case("ag_n300_d512_t007", "AG", 300, 512, tau=0.07, ids="cfg2"),
    Case("ag_all_same", "AG", 12, 32, ids="same"),
    Case("ampc_n8_d16", "AMPC", 8, 16, ids="hand8"),
    Case("ampc_cfg1", "AMPC", 64, 768, ids="twoview", string_ids=True),
    Case("ampc_n96_d128", "AMPC", 96, 128, ids="cfg3", seed=21),
    Case("ampc_all_single", "AMPC", 16, 64, ids="unique"),
]

BY_NAME = {c.name: c for c in CASES + AVGPOS_CASES}


def build_ids(case: Case) -> np.ndarray:
    total = case.n + case.extra_ids
    if case.ids == "hand8":
        ids = np.array([0, 1, 0, 2, 1, 0, 3, 4], dtype=np.int32)
    elif case.ids == "twoview":
        half = (total + 1) // 2
        ids = np.concatenate([np.arange(half), np.arange(half)])[:total].astype(np.int32)
    elif case.ids == "unique":
        ids = np.arange(total, dtype=np.int32)
    elif case.ids == "same":
        ids = np.zeros(total, dtype=np.int32)
    elif case.ids == "cfg2":
        ids = synth.make_study_ids(total, synth.SIZES_CFG2, seed=case.seed)
    elif case.ids == "cfg3":
        ids = synth.make_study_ids(total, synth.SIZES_CFG3, seed=case.seed)
    else:
        raise ValueError(case.ids)
    return ids


def build_inputs(case: Case):
    """-> dict(ids=<np array, str or int32, length n+extra>, image=[n,d] f32, text=[n,d] f32|None)"""
    ids = build_ids(case)
    image = synth.make_embeddings(ids[: case.n], case.d, seed=case.seed + 1)
    text = None
    if case.kind in ("G", "AG"):
        text = synth.make_embeddings(ids[: case.n], case.d, seed=case.seed + 2)
    if case.zero_row >= 0:
        image[case.zero_row] = 0.0
    ids_out = synth.ids_as_strings(ids) if case.string_ids else ids
    return dict(ids=ids_out, int_ids=ids, image=image, text=text)


def sample_rows(case: Case) -> np.ndarray:
    if case.n * case.d <= FULL_GRAD_LIMIT:
        return np.arange(case.n)
    rng = np.random.Generator(np.random.PCG64(case.seed + 99))
    return np.sort(rng.choice(case.n, size=SAMPLE_ROWS, replace=False))


def golden_path(case: Case) -> str:
    return os.path.join(GOLDEN_DIR, case.name + ".npz")


def load_golden(case: Case):
    return np.load(golden_path(case))
