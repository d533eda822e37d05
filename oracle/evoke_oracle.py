"""CPU oracle for EVOKE's multi-positive contrastive hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product path
(``evoke_b200``) never does: it calls the sm_100a kernels through the C-ABI
library and fails loudly when that library is missing.

What is restated here (file:line relative to the reference checkout,
models/model_pretrain_finetune_v0520.py):

* ``posmask_dense`` / ``posmask_packed``  -> label matrix build, :488-490 and :422-424
* ``l2_normalize``                         -> F.normalize(dim=-1, p=2), :495-496, :436
* ``g_loss_closed_form``                   -> Pretrain.global_alignment_loss, :486-504
* ``mpc_closed_form``                      -> Pretrain.multi_pos_contra_images_v0401, :421-446
* ``avgpos_g_loss_closed_form`` / ``avgpos_mpc_closed_form``
                                           -> PretrainNewMulPos.global_alignment_loss :748-815,
                                              .multi_pos_contra_images_v0404 :670-708
* ``*_port``                               -> the same two functions as the op sequence the
                                              reference executes in PyTorch (normalize, two mm,
                                              soft-target cross_entropy), used as the timed CPU
                                              baseline (kind "port") because the Python reference
                                              cannot travel to the GPU box.

Third-party arithmetic: the numerics of the reference live in PyTorch (pinned by the
reference at torch==2.1.2, README.md:120; this image has 2.11).  Semantics relied on:
``F.normalize`` = x / max(||x||_2, 1e-12); ``F.cross_entropy`` with float targets of the
logits' shape = mean_i( -sum_j target_ij * log_softmax(logits)_ij ).

Parity pin: the reference ships no tests or golden vectors for this path, so the pin is
the reference ITSELF executed in the build container through ``oracle/ref_shim.py``;
``oracle/make_golden.py`` records its outputs (loss, gradients, mask) as fixtures under
``tests/golden/`` and ``tests/test_oracle.py`` checks every function here against them.

The closed forms are fp64 numpy.  Notation: M_ij=[id_i==id_j], c_i=sum_j M_ij,
Y=M/c_i, S = Ihat That^T / tau.
"""
from __future__ import annotations

import math

import numpy as np

EPS = 1e-12  # F.normalize default eps


# ----------------------------------------------------------------------------- ids / mask
def factorize_ids(ids) -> np.ndarray:
    """Opaque keys (numpy <U strings or ints) -> dense int32 codes with the same equality
    structure.  The reference only ever tests ids for equality (:489, :422)."""
    ids = np.asarray(ids)
    _, inv = np.unique(ids, return_inverse=True)
    return inv.astype(np.int32).reshape(-1)


def posmask_dense(ids_row, ids_col=None, clear_diag: bool = False, row_offset: int = 0) -> np.ndarray:
    """bool [R, C]: (ids.reshape(-1,1) == ids.reshape(1,-1)), optionally with the diagonal
    (column == row + row_offset) cleared as multi_pos_contra_images_v0401 does (:424)."""
    ids_row = np.asarray(ids_row).reshape(-1, 1)
    ids_col = ids_row.reshape(1, -1) if ids_col is None else np.asarray(ids_col).reshape(1, -1)
    m = ids_row == ids_col
    if clear_diag:
        r = np.arange(m.shape[0])
        c = r + row_offset
        ok = (c >= 0) & (c < m.shape[1])
        m[r[ok], c[ok]] = False
    return m


def posmask_packed(ids_row, ids_col=None, clear_diag: bool = False, row_offset: int = 0):
    """(uint32 [R, ceil(C/32)], int32 counts[R]).  Bit k of word w is column 32*w+k, i.e.
    ``np.packbits(M, axis=1, bitorder='little')`` viewed as little-endian uint32."""
    m = posmask_dense(ids_row, ids_col, clear_diag, row_offset)
    r, c = m.shape
    words = (c + 31) // 32
    pad = np.zeros((r, words * 32), dtype=bool)
    pad[:, :c] = m
    packed = np.packbits(pad, axis=1, bitorder="little").view("<u4").reshape(r, words)
    return packed.astype(np.uint32), m.sum(1).astype(np.int32)


# ----------------------------------------------------------------------------- normalise
def l2_normalize(x: np.ndarray):
    """(xhat, norm): xhat = x / max(||x||, EPS) row-wise."""
    nrm = np.sqrt((x * x).sum(-1, keepdims=True))
    den = np.maximum(nrm, EPS)
    return x / den, nrm


def l2_normalize_bwd(x: np.ndarray, g_hat: np.ndarray) -> np.ndarray:
    """Gradient of l2_normalize: (g - xhat (xhat.g)) / ||x||, and g/EPS where the clamp is
    active (||x|| < EPS: the denominator is the constant EPS)."""
    xhat, nrm = l2_normalize(x)
    den = np.maximum(nrm, EPS)
    proj = (xhat * g_hat).sum(-1, keepdims=True)
    g = (g_hat - xhat * proj) / den
    clamp = (nrm < EPS)
    return np.where(clamp, g_hat / EPS, g)


def _lse(a: np.ndarray, axis: int) -> np.ndarray:
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


# ----------------------------------------------------------------------------- G loss
def g_loss_closed_form(image, text, ids, tau: float):
    """Pretrain.global_alignment_loss (:486-504) in closed form, fp64.

    loss = 1/2 [ mean_i( LSE_j S_ij - sum_j Y_ij S_ij ) + mean_j( LSE_i S_ij - sum_i Y_ji S_ij ) ]
    dS   = 1/(2N) [ softmax_row(S) + softmax_col(S) - Y - Y^T ]
    Returns (loss, dImage, dText, extras).  The reference's labels are fp32 (1/c rounded to
    fp32, :490-491); that rounding is reproduced so the pin holds to ~1e-12 instead of 6e-8.
    """
    image = np.asarray(image, dtype=np.float64)
    text = np.asarray(text, dtype=np.float64)
    n = image.shape[0]
    ids = np.asarray(ids)[:n]                                   # :488 truncation
    m = posmask_dense(ids).astype(np.float32)
    y = (m / m.sum(1, keepdims=True)).astype(np.float64)        # fp32 division, as :491
    ih, _ = l2_normalize(image)
    th, _ = l2_normalize(text)
    s = ih @ th.T / tau
    lse_r = _lse(s, 1)
    lse_c = _lse(s, 0)
    # torch evaluates -sum_j y_ij*log_softmax_ij, so LSE_i is weighted by sum_j y_ij, which
    # is 1 only up to the fp32 rounding of 1/c (3*fp32(1/3) = 1+3e-8).  Kept for a tight pin.
    ysum = y.sum(1)
    loss = 0.5 * ((ysum * lse_r - (y * s).sum(1)).mean() + (ysum * lse_c - (y.T * s).sum(0)).mean())
    p_r = np.exp(s - lse_r[:, None])
    p_c = np.exp(s - lse_c[None, :])
    ds = (p_r * ysum[:, None] + p_c * ysum[None, :] - y - y.T) / (2.0 * n)
    d_ih = ds @ th / tau
    d_th = ds.T @ ih / tau
    extras = dict(S=s, lse_row=lse_r, lse_col=lse_c, d_ihat=d_ih, d_that=d_th)
    return float(loss), l2_normalize_bwd(image, d_ih), l2_normalize_bwd(text, d_th), extras


# ----------------------------------------------------------------------------- MPC loss
def mpc_kept_rows(ids) -> np.ndarray:
    """Rows with at least one OTHER row of the same id (:424-426)."""
    m = posmask_dense(ids, clear_diag=True)
    return np.nonzero(m.sum(1) != 0)[0]


def mpc_closed_form(x, ids, tau: float):
    """Pretrain.multi_pos_contra_images_v0401 (:421-446), fp64 closed form.

    Rows without a second view are dropped from queries AND keys (:429); the diagonal is
    excluded from the softmax (:438, -1e9 fill => exactly zero probability and gradient).
    Returns (loss, dX) with dX zero on dropped rows, or (None, zeros) when nothing is kept
    (the reference then returns the leaf tensor([0.0]) of shape [1], :427-428).
    """
    x = np.asarray(x, dtype=np.float64)
    ids = np.asarray(ids)
    idx = mpc_kept_rows(ids)
    dx = np.zeros_like(x)
    if len(idx) == 0:
        return None, dx
    xk = x[idx]
    m = posmask_dense(ids[idx], clear_diag=True).astype(np.float64)
    y = m / m.sum(1, keepdims=True)
    xh, _ = l2_normalize(xk)
    s = xh @ xh.T / tau
    np.fill_diagonal(s, -np.inf)
    lse = _lse(s, 1)
    s0 = np.where(np.isfinite(s), s, 0.0)
    k = len(idx)
    loss = (lse - (y * s0).sum(1)).mean()
    p = np.exp(s - lse[:, None])
    ds = (p - y) / k
    d_xh = (ds + ds.T) @ xh / tau
    dx[idx] = l2_normalize_bwd(xk, d_xh)
    return float(loss), dx


# ----------------------------------------------------------------------------- a8 variants
def _avgpos_rows(s: np.ndarray, m: np.ndarray) -> np.ndarray:
    """Per-row loss of the 'averaged positive logit' rule (:783-811, :691-705):
    l_i = -pbar_i + log(exp(pbar_i) + sum_{neg} exp(S_ij)), pbar_i = mean of positive logits."""
    out = np.zeros(s.shape[0])
    for i in range(s.shape[0]):
        pos = m[i] != 0
        pbar = s[i, pos].sum() / pos.sum()
        neg = s[i, ~pos & np.isfinite(s[i])]
        z = np.concatenate([[pbar], neg])
        out[i] = -pbar + _lse(z, 0)
    return out


def _avgpos_rows_grad(s: np.ndarray, m: np.ndarray):
    """Vectorised per-row loss AND its gradient w.r.t. s (rows with no positive get 0 / 0).
    dl_i/ds_ij = e^{s_ij}/Z_i for negatives, (e^{pbar_i}/Z_i - 1)/c_i for positives (entries at -inf: 0)."""
    m = m.astype(bool)
    c = m.sum(1)
    ok = c > 0
    cs = np.where(ok, c, 1)
    fin = np.isfinite(s)
    pbar = np.where(m, s, 0.0).sum(1) / cs
    neg = (~m) & fin
    mx = np.maximum(pbar, np.where(neg, s, -np.inf).max(1, initial=-np.inf))
    e = np.where(neg, np.exp(np.where(neg, s, 0.0) - mx[:, None]), 0.0)
    u = np.exp(pbar - mx)
    z = u + e.sum(1)
    loss = np.where(ok, -pbar + mx + np.log(z), 0.0)
    g = e / z[:, None] + m * ((u / z - 1.0) / cs)[:, None]
    g[~ok] = 0.0
    return loss, g


def avgpos_g_closed_form(image, text, ids, tau: float):
    """PretrainNewMulPos.global_alignment_loss (:748-815) with gradients (fp64):
    -> (loss, d_image, d_text).  Single-positive rows reduce to plain CE (:770-777), so one formula
    covers both branches; loss = 0.5 * (sum_i l_i + sum_j l'_j) / B (:813)."""
    image = np.asarray(image, dtype=np.float64)
    text = np.asarray(text, dtype=np.float64)
    n = image.shape[0]
    m = posmask_dense(np.asarray(ids)[:n])
    ih, _ = l2_normalize(image)
    th, _ = l2_normalize(text)
    s = ih @ th.T / tau
    lr, gr = _avgpos_rows_grad(s, m)
    lc, gc_ = _avgpos_rows_grad(s.T, m)
    loss = 0.5 * (lr.sum() + lc.sum()) / n
    ds = 0.5 / n * (gr + gc_.T)
    d_ih = ds @ th / tau
    d_th = ds.T @ ih / tau
    return float(loss), l2_normalize_bwd(image, d_ih), l2_normalize_bwd(text, d_th)


def avgpos_g_loss_closed_form(image, text, ids, tau: float) -> float:
    """Forward value of avgpos_g_closed_form, by the row-by-row restatement of the reference's loop."""
    image = np.asarray(image, dtype=np.float64)
    text = np.asarray(text, dtype=np.float64)
    n = image.shape[0]
    m = posmask_dense(np.asarray(ids)[:n])
    ih, _ = l2_normalize(image)
    th, _ = l2_normalize(text)
    s = ih @ th.T / tau
    return float(0.5 * (_avgpos_rows(s, m).sum() + _avgpos_rows(s.T, m).sum()) / n)


def avgpos_mpc_grad_closed_form(x, ids, tau: float):
    """PretrainNewMulPos.multi_pos_contra_images_v0404 (:670-708) with gradient (fp64): -> (loss, dx) or
    (None, zeros).  Single-view rows are removed from the QUERIES only (:685); every row stays a key."""
    x = np.asarray(x, dtype=np.float64)
    ids = np.asarray(ids)
    idx = mpc_kept_rows(ids)
    if len(idx) == 0:
        return None, np.zeros_like(x)
    m = posmask_dense(ids, clear_diag=True)
    xh, _ = l2_normalize(x)
    s = xh @ xh.T / tau
    np.fill_diagonal(s, -np.inf)
    l, g = _avgpos_rows_grad(s, m)          # rows without positives contribute 0 / 0
    loss = l.sum() / len(idx)
    ds = g / len(idx)
    d_xh = (ds + ds.T) @ xh / tau
    return float(loss), l2_normalize_bwd(x, d_xh)


def avgpos_mpc_closed_form(x, ids, tau: float):
    """PretrainNewMulPos.multi_pos_contra_images_v0404 (:670-708): rows filtered, ALL
    columns kept as keys (:685), diagonal excluded."""
    x = np.asarray(x, dtype=np.float64)
    ids = np.asarray(ids)
    idx = mpc_kept_rows(ids)
    if len(idx) == 0:
        return None
    m = posmask_dense(ids, clear_diag=True)
    xh, _ = l2_normalize(x)
    s = xh @ xh.T / tau
    np.fill_diagonal(s, -np.inf)
    return float(_avgpos_rows(s[idx], m[idx]).sum() / len(idx))


# ----------------------------------------------------------------------------- f1 (next row)
def local_token_alignment_closed_form(local_image, local_text, tau: float):
    """Pretrain.local_text_token_alignment_loss (:506-526), fp64, with gradients: -> (loss, d_image, d_text).

    local_image [B, P, D] patch tokens, local_text [B, L, D] text tokens.  Per sample: the text tokens attend over
    the patches (softmax(T V^T / sqrt(D)) V, :509-511), both sides are L2-normalised (:514-515), and an L x L
    token-level InfoNCE with identity targets is taken in both directions over all B*L rows (:518-525).
    Not yet built in CUDA; this pins the oracle for it."""
    v = np.asarray(local_image, dtype=np.float64)
    t = np.asarray(local_text, dtype=np.float64)
    b, l, d = t.shape
    s1 = np.einsum("bld,bpd->blp", t, v) / math.sqrt(d)
    s1 = s1 - s1.max(-1, keepdims=True)
    a = np.exp(s1)
    a /= a.sum(-1, keepdims=True)
    o = np.einsum("blp,bpd->bld", a, v)
    on = np.maximum(np.linalg.norm(o, axis=-1, keepdims=True), 1e-12)
    tn = np.maximum(np.linalg.norm(t, axis=-1, keepdims=True), 1e-12)
    oh, th = o / on, t / tn
    sim = np.einsum("bld,bmd->blm", th, oh) / tau               # [b, n1 (text), n2 (attended)]
    lse_r = _lse(sim, 2)
    lse_c = _lse(sim, 1)
    diag = np.einsum("bll->bl", sim)
    n_rows = b * l
    loss = 0.5 * ((lse_r - diag).sum() + (lse_c - diag).sum()) / n_rows
    # d loss / d sim
    p_r = np.exp(sim - lse_r[:, :, None])
    p_c = np.exp(sim - lse_c[:, None, :])
    eye = np.eye(l)[None]
    dsim = 0.5 / n_rows * ((p_r - eye) + (p_c - eye))
    d_th = np.einsum("blm,bmd->bld", dsim, oh) / tau
    d_oh = np.einsum("blm,bld->bmd", dsim, th) / tau
    # through the normalisations (no clamp in these cases)
    d_o = (d_oh - oh * (oh * d_oh).sum(-1, keepdims=True)) / on
    d_t = (d_th - th * (th * d_th).sum(-1, keepdims=True)) / tn
    # through the attention
    d_a = np.einsum("bld,bpd->blp", d_o, v)
    d_v = np.einsum("blp,bld->bpd", a, d_o)
    d_s1 = a * (d_a - (d_a * a).sum(-1, keepdims=True)) / math.sqrt(d)
    d_t = d_t + np.einsum("blp,bpd->bld", d_s1, v)
    d_v = d_v + np.einsum("blp,bld->bpd", d_s1, t)
    return float(loss), d_v, d_t


# ----------------------------------------------------------------------------- torch ports
def global_alignment_loss_port(image, text, ids, tau: float):
    """The op sequence of :486-504 on whatever device/dtype ``image`` lives on (CPU fp32 in
    the timed baseline).  ids: numpy array (str or int), as the reference receives them."""
    import torch
    import torch.nn.functional as F
    ids = np.asarray(ids)[: image.shape[0]]
    same = (ids[:, None] == ids[None, :]).astype(int)
    target = torch.from_numpy(same).float().to(image.device)
    target = target / target.sum(1, keepdim=True)
    ih = F.normalize(image, p=2, dim=-1)
    th = F.normalize(text, p=2, dim=-1)
    sim_it = ih @ th.t()
    sim_ti = th @ ih.t()
    return 0.5 * (F.cross_entropy(sim_it / tau, target) + F.cross_entropy(sim_ti / tau, target))


def multi_pos_contra_images_port(x, ids, tau: float):
    """The op sequence of :421-446."""
    import torch
    import torch.nn.functional as F
    ids = np.asarray(ids)
    same = (ids[:, None] == ids[None, :]).astype(float)
    target = torch.from_numpy(same).to(x)
    target.fill_diagonal_(0.0)
    keep = torch.nonzero(target.sum(1) != 0).reshape(-1)
    if keep.numel() == 0:
        return torch.tensor([0.0], requires_grad=True, device=x.device)
    x = x[keep]
    target = target[keep][:, keep]
    target = target / target.sum(1, keepdim=True)
    xh = F.normalize(x, p=2, dim=-1)
    logits = xh @ xh.t() / tau
    logits.fill_diagonal_(-1e9)
    logits = logits - logits.max(dim=-1, keepdim=True).values.detach()
    return F.cross_entropy(logits, target)


def local_text_token_alignment_port(local_image, local_text, tau: float):
    """The op sequence of Pretrain.local_text_token_alignment_loss (:506-526) on whatever device the inputs live on
    (timed baseline of the full-step harness, tools/pretrain_step.py)."""
    import torch
    import torch.nn.functional as F
    sim = local_text @ local_image.permute(0, 2, 1)
    sco = F.softmax(sim / math.sqrt(local_image.shape[2]), dim=-1)
    out = F.normalize(torch.bmm(sco, local_image), dim=-1, p=2)
    text = F.normalize(local_text, dim=-1, p=2)
    word = torch.bmm(text, out.permute(0, 2, 1)) / tau
    b, n1, n2 = word.shape
    targets = torch.arange(n1, device=word.device).long().repeat(b)
    l1 = F.cross_entropy(word.reshape(b * n1, n2), targets)
    l2 = F.cross_entropy(word.permute(0, 2, 1).reshape(b * n2, n1), targets)
    return (l1 + l2) / 2.0
