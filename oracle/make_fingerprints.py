"""TEST / BENCH INFRASTRUCTURE: gradient fingerprints of the benchmark workloads from the fp64 oracle.

bench.py prints a fingerprint of the gradients it produced (loss, ||dImage||, ||dText||, norms and leading values of
16 fixed rows) at every GPU count and compares it with the values recorded here, so that the driver's own bench and
scaling runs prove that the timed configuration computes the reference's gradients - on 1 GPU and on 2/4/8.

The closed form is that of oracle/evoke_oracle.py::g_loss_closed_form (reference
models/model_pretrain_finetune_v0520.py:486-504), evaluated in row blocks so that N = 32768 fits in host memory;
tests/test_oracle.py pins the blockwise form to the plain one.  Run in the build container:

    python oracle/make_fingerprints.py            # writes tests/golden/fingerprint_<cfg>.json
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from evoke_b200 import synth  # noqa: E402  (numpy-only workload generator shared with bench.py)
from oracle import evoke_oracle as orc  # noqa: E402

N_ROWS = 16


def fingerprint_rows(n: int) -> np.ndarray:
    return (np.arange(N_ROWS) * (n // N_ROWS) + 7) % n


def g_loss_blockwise(image, text, key, tau, key2=None, block=2048):
    """fp64 loss and gradients of the G loss, O(block * N) memory.  Same math as g_loss_closed_form (without the
    reference's fp32 rounding of 1/c, a 3e-8 effect)."""
    x = np.asarray(image, dtype=np.float64)
    y = np.asarray(text, dtype=np.float64)
    n = x.shape[0]
    k1 = np.asarray(key)[:n]
    k2 = None if key2 is None else np.asarray(key2)[:n]
    xh, nx = orc.l2_normalize(x)
    yh, ny = orc.l2_normalize(y)

    def mask(r0, r1):
        m = k1[r0:r1, None] == k1[None, :]
        if k2 is not None:
            m &= k2[r0:r1, None] == k2[None, :]
        return m

    lse_r = np.empty(n)
    pos = np.empty(n)
    cnt = np.empty(n)
    col_m = np.full(n, -np.inf)
    col_s = np.zeros(n)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        s = xh[r0:r1] @ yh.T / tau
        m = mask(r0, r1)
        mx = s.max(1)
        lse_r[r0:r1] = mx + np.log(np.exp(s - mx[:, None]).sum(1))
        pos[r0:r1] = np.where(m, s, 0.0).sum(1)
        cnt[r0:r1] = m.sum(1)
        new_m = np.maximum(col_m, s.max(0))
        col_s = col_s * np.exp(col_m - new_m) + np.exp(s - new_m[None, :]).sum(0)
        col_m = new_m
    lse_c = col_m + np.log(col_s)
    loss = 0.5 * ((lse_r - pos / cnt).mean() + (lse_c.mean() - (pos / cnt).mean()))
    d_xh = np.empty_like(xh)
    d_yh = np.zeros_like(yh)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        s = xh[r0:r1] @ yh.T / tau
        m = mask(r0, r1)
        ds = np.exp(s - lse_r[r0:r1, None]) + np.exp(s - lse_c[None, :])
        ds -= m / cnt[r0:r1, None]
        ds -= m / cnt[None, :]
        ds /= 2.0 * n
        d_xh[r0:r1] = ds @ yh / tau
        d_yh += ds.T @ xh[r0:r1] / tau
    return float(loss), orc.l2_normalize_bwd(x, d_xh), orc.l2_normalize_bwd(y, d_yh)


def fingerprint(loss, d_image, d_text) -> dict:
    rows = fingerprint_rows(d_image.shape[0])
    return {
        "loss": float(loss),
        "d_image_norm": float(np.linalg.norm(d_image)),
        "d_text_norm": float(np.linalg.norm(d_text)),
        "rows": [int(r) for r in rows],
        "d_image_row_norms": [float(v) for v in np.linalg.norm(d_image[rows], axis=1)],
        "d_text_row_norms": [float(v) for v in np.linalg.norm(d_text[rows], axis=1)],
        "d_image_row_head": [[float(v) for v in d_image[r, :4]] for r in rows],
        "d_text_row_head": [[float(v) for v in d_text[r, :4]] for r in rows],
    }


def workload(cfg: str):
    """(image, text, key, key2, tau) of a bench workload - the arrays bench.py builds (same seeds)."""
    if cfg == "cfg2":
        n, d, sizes = 4096, 768, synth.SIZES_CFG2
    elif cfg == "cfg3":
        n, d, sizes = 16384, 768, synth.SIZES_CFG3
    elif cfg == "cfg4":
        n, d = 32768, 512
        pat, stu = synth.make_patient_study_ids(n, seed=1234)
        return synth.make_embeddings(stu, d, seed=1235), synth.make_embeddings(stu, d, seed=1236), pat, stu, 0.5
    else:
        raise ValueError(cfg)
    ids = synth.make_study_ids(n, sizes, seed=1234)
    return synth.make_embeddings(ids, d, seed=1235), synth.make_embeddings(ids, d, seed=1236), ids, None, 0.5


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for cfg in (sys.argv[1:] or ["cfg2", "cfg3", "cfg4"]):
        xi, xt, key, key2, tau = workload(cfg)
        loss, d_i, d_t = g_loss_blockwise(xi, xt, key, tau, key2)
        fp = fingerprint(loss, d_i, d_t)
        fp["config"] = cfg
        fp["source"] = "oracle/make_fingerprints.py: fp64 closed form of reference v0520.py:486-504, blockwise numpy"
        with open(os.path.join(out_dir, f"fingerprint_{cfg}.json"), "w") as f:
            json.dump(fp, f, indent=1)
        print(cfg, fp["loss"], fp["d_image_norm"], fp["d_text_norm"], flush=True)


if __name__ == "__main__":
    main()
