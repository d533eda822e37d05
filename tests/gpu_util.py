"""Helpers for the -m gpu parity tests (all comparisons are against oracle/ or tests/golden/)."""
import numpy as np
import torch

import golden_cases as gc
from oracle import evoke_oracle as orc

DEV = "cuda"

# tolerances stated by BASELINE.json north_star
TOL = {
    "fp32": dict(loss=1e-5, grad=1e-4),
    "bf16": dict(loss=2e-5, grad=2e-2),     # loss in bf16 mode: not stated upstream; achieved 7e-8 .. 8e-7 from N = 300 up
    "bf16_tiny": dict(loss=2e-3, grad=2e-2),   # N * D < 64 Ki (hand-checkable cases): no averaging over rows, 8e-5 seen
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
    tol = TOL["bf16_tiny" if precision == "bf16" and case.n * case.d < 65536 else precision]
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


# ------------------------------------------------------------------------------------- full-size reference
def row_rel_l2(got, want):
    """Per-row relative L2 error ||g_i - w_i|| / ||w_i|| (max over rows whose reference gradient is not tiny):
    stricter than the max-norm metric for rows with small gradients."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    nw = np.linalg.norm(want, axis=1)
    ok = nw > 1e-3 * nw.max()
    return float((np.linalg.norm(got - want, axis=1)[ok] / nw[ok]).max())


def g_loss_fp64_gpu(image, text, key, tau, key2=None, block=4096):
    """TEST INFRASTRUCTURE: the closed form of oracle.g_loss_closed_form (reference :486-504) evaluated in fp64 with
    torch on the GPU, in row blocks, for sizes the numpy oracle does not finish in seconds (N = 16384, 32768).
    tests/test_gpu_fullsize.py pins it to the numpy oracle at a small size before using it.
    image/text: numpy or torch [N, D]; key (, key2): integer ids (two-component key = patient AND study).
    Returns (loss, d_image, d_text) as float / numpy fp64."""
    dev = torch.device(DEV)
    x = torch.as_tensor(image).to(dev, torch.float64)
    y = torch.as_tensor(text).to(dev, torch.float64)
    n = x.shape[0]
    k1 = torch.as_tensor(np.asarray(key)).to(dev).long()[:n]
    k2 = None if key2 is None else torch.as_tensor(np.asarray(key2)).to(dev).long()[:n]
    nx = x.norm(dim=1, keepdim=True)
    ny = y.norm(dim=1, keepdim=True)
    xh = x / nx.clamp_min(1e-12)
    yh = y / ny.clamp_min(1e-12)

    def mask(r0, r1):
        m = k1[r0:r1, None] == k1[None, :]
        if k2 is not None:
            m &= k2[r0:r1, None] == k2[None, :]
        return m

    # pass 1: row LSE, positive sums, counts; column LSE by streaming log-sum-exp over the row blocks
    lse_r = torch.empty(n, dtype=torch.float64, device=dev)
    pos = torch.empty(n, dtype=torch.float64, device=dev)
    cnt = torch.empty(n, dtype=torch.float64, device=dev)
    col_m = torch.full((n,), -float("inf"), dtype=torch.float64, device=dev)
    col_s = torch.zeros(n, dtype=torch.float64, device=dev)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        s = xh[r0:r1] @ yh.T / tau
        m = mask(r0, r1)
        lse_r[r0:r1] = torch.logsumexp(s, dim=1)
        pos[r0:r1] = (s * m).sum(1)
        cnt[r0:r1] = m.sum(1)
        bm = s.max(dim=0).values
        new_m = torch.maximum(col_m, bm)
        col_s = col_s * torch.exp(col_m - new_m) + torch.exp(s - new_m[None, :]).sum(0)
        col_m = new_m
    lse_c = col_m + col_s.log()
    # symmetric mask: sum_i Y_ji S_ij over column j equals sum over the row-side positives of the transposed problem;
    # M_ij = 1 implies c_i = c_j, so sum_j (1/c_j) sum_i M_ij S_ij = sum_i pos_i / c_i
    loss = 0.5 * ((lse_r - pos / cnt).mean() + (lse_c.mean() - (pos / cnt).mean()))
    # pass 2: dS block -> gradients
    d_xh = torch.empty_like(xh)
    d_yh = torch.zeros_like(yh)
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        s = xh[r0:r1] @ yh.T / tau
        m = mask(r0, r1).to(torch.float64)
        ds = torch.exp(s - lse_r[r0:r1, None]) + torch.exp(s - lse_c[None, :])
        ds -= m / cnt[r0:r1, None]
        ds -= m / cnt[None, :]
        ds /= 2.0 * n
        d_xh[r0:r1] = ds @ yh / tau
        d_yh += ds.T @ xh[r0:r1] / tau

    def through_normalize(v, vh, nv, g):
        out = (g - vh * (vh * g).sum(1, keepdim=True)) / nv.clamp_min(1e-12)
        return torch.where(nv < 1e-12, g / 1e-12, out)

    d_x = through_normalize(x, xh, nx, d_xh)
    d_y = through_normalize(y, yh, ny, d_yh)
    return float(loss.item()), d_x.cpu().numpy(), d_y.cpu().numpy()
