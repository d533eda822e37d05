"""Generate tests/golden/*.npz by RUNNING THE REFERENCE (unmodified, via oracle/ref_shim.py)
on the seeded inputs of tests/golden_cases.py.  Container-only: needs /root/reference.

    python oracle/make_golden.py            # writes every case
    python oracle/make_golden.py g_cfg1     # one case

Stored per case (all from the reference's own PyTorch code on CPU):
  loss32, loss64        the scalar returned for fp32 / fp64 inputs (nan when the MPC path
                        returned the shape-[1] zero leaf; ``empty`` is then 1)
  rows                  row indices whose gradients are stored (all rows for small cases)
  d_image32/64[rows], d_text32/64[rows]   autograd gradients of the inputs
  d_image_norm64, d_text_norm64           Frobenius norms of the full gradients
  out_shape             shape of the returned tensor (() or (1,))
  mask_bits, counts     for n <= 1024: np.packbits(little) of the reference's label>0 matrix
                        as uint32 words, and its row sums (positives per row)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import golden_cases as gc  # noqa: E402
from oracle import ref_shim  # noqa: E402


def _run(case: gc.Case, dtype):
    inp = gc.build_inputs(case)
    image = torch.tensor(inp["image"], dtype=dtype, requires_grad=True)
    if case.kind in ("G", "AG"):
        text = torch.tensor(inp["text"], dtype=dtype, requires_grad=True)
        fn = ref_shim.global_alignment_loss if case.kind == "G" else ref_shim.avgpos_global_alignment_loss
        out = fn(image, text, inp["ids"], case.tau)
        out.sum().backward()
        return out, image.grad, text.grad
    fn = ref_shim.multi_pos_contra_images_v0401 if case.kind == "MPC" else ref_shim.avgpos_multi_pos_contra_images_v0404
    out = fn(image, inp["ids"], case.tau)
    if out.grad_fn is None:                     # the [0.0] leaf: nothing flows to the input
        return out, torch.zeros_like(image), None
    out.sum().backward()
    return out, image.grad, None


def _reference_mask(case: gc.Case):
    """The label matrix exactly as the reference builds it (numpy ==, :489 / :422-424)."""
    ids = gc.build_inputs(case)["ids"]
    if case.kind in ("G", "AG"):
        ids = ids[: case.n]
    m = ids.reshape(-1, 1) == ids.reshape(1, -1)
    if case.kind in ("MPC", "AMPC"):
        np.fill_diagonal(m, False)
    n = m.shape[0]
    words = (n + 31) // 32
    pad = np.zeros((n, words * 32), dtype=bool)
    pad[:, :n] = m
    bits = np.packbits(pad, axis=1, bitorder="little").view("<u4").reshape(n, words)
    return bits.astype(np.uint32), m.sum(1).astype(np.int32)


def make(case: gc.Case):
    torch.manual_seed(0)
    rows = gc.sample_rows(case)
    rec = {"rows": rows}
    for tag, dtype in (("32", torch.float32), ("64", torch.float64)):
        out, gi, gt = _run(case, dtype)
        empty = out.grad_fn is None
        rec["empty"] = np.int32(empty)
        rec["out_shape"] = np.array(out.shape, dtype=np.int64)
        rec["loss" + tag] = np.float64(out.detach().double().sum().item())
        rec["d_image" + tag] = gi.detach().numpy()[rows]
        if gt is not None:
            rec["d_text" + tag] = gt.detach().numpy()[rows]
        if tag == "64":
            rec["d_image_norm64"] = np.float64(gi.double().norm().item())
            if gt is not None:
                rec["d_text_norm64"] = np.float64(gt.double().norm().item())
    if case.n <= 1024:
        rec["mask_bits"], rec["counts"] = _reference_mask(case)
    os.makedirs(gc.GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(gc.golden_path(case), **rec)
    print(f"{case.name:24s} loss64={rec['loss64']:.12f} shape={tuple(rec['out_shape'])} "
          f"rows={len(rows)} -> {os.path.getsize(gc.golden_path(case))} B")


LOCAL_CASES = {"local_b3_l7_p5_d16": (3, 7, 5, 16, 0.5, 41), "local_b4_l99_p49_d64": (4, 99, 49, 64, 0.5, 42),
               "local_b2_l20_p49_d768_t02": (2, 20, 49, 768, 0.2, 43)}


def local_inputs(name):
    b, l, p, d, tau, seed = LOCAL_CASES[name]
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((b, p, d)).astype(np.float32), rng.standard_normal((b, l, d)).astype(np.float32), tau


def make_local(name):
    """f1: Pretrain.local_text_token_alignment_loss on seeded token tensors (fp64 run of the reference)."""
    v, t, tau = local_inputs(name)
    vi = torch.tensor(v, dtype=torch.float64, requires_grad=True)
    ti = torch.tensor(t, dtype=torch.float64, requires_grad=True)
    out = ref_shim.local_text_token_alignment_loss(vi, ti, tau)
    out.backward()
    path = os.path.join(gc.GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, loss64=np.float64(out.item()), d_image64=vi.grad.numpy()[:, :4], d_text64=ti.grad.numpy()[:, :4],
                        d_image_norm64=np.float64(vi.grad.norm().item()), d_text_norm64=np.float64(ti.grad.norm().item()))
    print(f"{name:28s} loss64={out.item():.12f} -> {os.path.getsize(path)} B")


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    names = sys.argv[1:] or ([c.name for c in gc.CASES + gc.AVGPOS_CASES] + list(LOCAL_CASES))
    for nm in names:
        if nm in LOCAL_CASES:
            make_local(nm)
        else:
            make(gc.BY_NAME[nm])
