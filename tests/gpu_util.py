"""Helpers for the -m gpu parity tests (all comparisons are against oracle/ or tests/golden/)."""
import numpy as np
import torch

import golden_cases as gc
from oracle import evoke_oracle as orc

DEV = "cuda"

# tolerances stated by BASELINE.json north_star
TOL = {
    "fp32": dict(loss=1e-5, grad=1e-4),
    "bf16": dict(loss=2e-3, grad=2e-2),     # loss in bf16 mode: not stated upstream; 2e-3 rel used here
}


def rel_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def run_case(case, precision, path, dtype=torch.float32):
    """Run the product on a golden case.  Returns (out tensor, d_image np, d_text np|None)."""
    import evoke_b200
    inp = gc.build_inputs(case)
    image = torch.tensor(inp["image"], device=DEV, dtype=dtype, requires_grad=True)
    if case.kind == "G":
        text = torch.tensor(inp["text"], device=DEV, dtype=dtype, requires_grad=True)
        out = evoke_b200.global_alignment(image, text, inp["ids"], case.tau, precision=precision, path=path)
        out.backward()
        return out, image.grad.float().cpu().numpy(), text.grad.float().cpu().numpy()
    out = evoke_b200.multi_pos_contra_images(image, inp["ids"], case.tau, precision=precision, path=path)
    if out.grad_fn is None:
        return out, np.zeros_like(inp["image"]), None
    out.backward()
    return out, image.grad.float().cpu().numpy(), None


def check_against_golden(case, precision, path):
    gold = gc.load_golden(case)
    out, d_i, d_t = run_case(case, precision, path)
    tol = TOL[precision]
    assert tuple(out.shape) == tuple(gold["out_shape"]), (out.shape, gold["out_shape"])
    if gold["empty"]:
        assert out.item() == 0.0 and out.requires_grad and out.grad_fn is None
        return dict(loss_rel=0.0)
    loss_rel = abs(out.item() - gold["loss64"]) / abs(gold["loss64"])
    rows = gold["rows"]
    gi = rel_max(d_i[rows], gold["d_image64"])
    res = dict(loss_rel=loss_rel, d_image_rel=gi)
    assert loss_rel <= tol["loss"], f"{case.name} {precision}/{path}: loss rel err {loss_rel:.3e}"
    assert gi <= tol["grad"], f"{case.name} {precision}/{path}: d_image rel err {gi:.3e}"
    nrm = float(np.linalg.norm(d_i.astype(np.float64)))
    assert abs(nrm - gold["d_image_norm64"]) <= 5 * tol["grad"] * gold["d_image_norm64"]
    if d_t is not None:
        gt = rel_max(d_t[rows], gold["d_text64"])
        res["d_text_rel"] = gt
        assert gt <= tol["grad"], f"{case.name} {precision}/{path}: d_text rel err {gt:.3e}"
    return res


def oracle_g(image, text, ids, tau):
    loss, d_i, d_t, _ = orc.g_loss_closed_form(image, text, ids, tau)
    return loss, d_i, d_t
