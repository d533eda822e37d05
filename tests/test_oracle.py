"""T0: the oracle (oracle/evoke_oracle.py) against the golden vectors recorded from the
reference itself (oracle/make_golden.py), and - in the build container only - against the
live reference through the import shim."""
import numpy as np
import pytest
import torch

import golden_cases as gc
from oracle import evoke_oracle as orc
from oracle import ref_shim

G_CASES = [c for c in gc.CASES if c.kind == "G"]
MPC_CASES = [c for c in gc.CASES if c.kind == "MPC"]


def _rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.mark.parametrize("case", G_CASES, ids=lambda c: c.name)
def test_g_closed_form_matches_reference_golden(case):
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    loss, d_i, d_t, _ = orc.g_loss_closed_form(inp["image"], inp["text"], inp["ids"], case.tau)
    rows = gold["rows"]
    assert abs(loss - gold["loss64"]) <= 1e-11 * abs(gold["loss64"])
    # the zero-norm row has a 1e12-scale gradient (g/eps); compare relatively
    assert _rel(d_i[rows], gold["d_image64"]) < 1e-9
    assert _rel(d_t[rows], gold["d_text64"]) < 1e-9
    assert abs(np.linalg.norm(d_i) - gold["d_image_norm64"]) <= 1e-9 * gold["d_image_norm64"]
    assert abs(np.linalg.norm(d_t) - gold["d_text_norm64"]) <= 1e-9 * gold["d_text_norm64"]
    # the fp32 run of the reference agrees with its own fp64 run to fp32 accuracy
    assert abs(gold["loss32"] - gold["loss64"]) <= 2e-6 * abs(gold["loss64"])


@pytest.mark.parametrize("case", MPC_CASES, ids=lambda c: c.name)
def test_mpc_closed_form_matches_reference_golden(case):
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    loss, dx = orc.mpc_closed_form(inp["image"], inp["ids"], case.tau)
    if gold["empty"]:
        assert loss is None and tuple(gold["out_shape"]) == (1,) and gold["loss64"] == 0.0
        assert not dx.any()
        return
    assert tuple(gold["out_shape"]) == ()
    assert abs(loss - gold["loss64"]) <= 1e-11 * abs(gold["loss64"])
    assert _rel(dx[gold["rows"]], gold["d_image64"]) < 1e-9
    assert abs(np.linalg.norm(dx) - gold["d_image_norm64"]) <= 1e-9 * gold["d_image_norm64"]


@pytest.mark.parametrize("case", gc.AVGPOS_CASES, ids=lambda c: c.name)
def test_avgpos_closed_forms_match_reference_golden(case):
    """PretrainNewMulPos.global_alignment_loss / multi_pos_contra_images_v0404."""
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    rows = gold["rows"]
    assert tuple(gold["out_shape"]) == (1,)
    if case.kind == "AG":
        loss, d_i, d_t = orc.avgpos_g_closed_form(inp["image"], inp["text"], inp["ids"], case.tau)
        # the reference accumulates this loss in a float32 tensor (torch.tensor([0.0]) :768, :693) whatever the
        # input dtype, so its value is only fp32-accurate; its gradients are fp64
        assert abs(loss - gold["loss64"]) <= 2e-6 * max(abs(gold["loss64"]), 1.0)
        assert abs(orc.avgpos_g_loss_closed_form(inp["image"], inp["text"], inp["ids"], case.tau) - loss) <= 1e-12 * max(abs(loss), 1.0)
        scale = max(np.abs(gold["d_image64"]).max(), 1e-30)
        # (the fp32 loss tensor also rounds the 1/B factor of the backward: gradients agree to ~6e-8 relative)
        assert np.abs(d_i[rows] - gold["d_image64"]).max() <= 3e-7 * scale + 1e-15
        assert np.abs(d_t[rows] - gold["d_text64"]).max() <= 3e-7 * max(np.abs(gold["d_text64"]).max(), 1e-30) + 1e-15
        return
    loss, dx = orc.avgpos_mpc_grad_closed_form(inp["image"], inp["ids"], case.tau)
    if gold["empty"]:
        assert loss is None and gold["loss64"] == 0.0 and not dx.any()
        return
    assert abs(loss - gold["loss64"]) <= 2e-6 * abs(gold["loss64"])
    assert abs(orc.avgpos_mpc_closed_form(inp["image"], inp["ids"], case.tau) - loss) <= 1e-12 * abs(loss)
    assert _rel(dx[rows], gold["d_image64"]) < 3e-7


@pytest.mark.parametrize("case", [c for c in gc.CASES if c.n <= 1024], ids=lambda c: c.name)
def test_packed_mask_bit_exact(case):
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    ids = inp["ids"][: case.n] if case.kind == "G" else inp["ids"]
    bits, counts = orc.posmask_packed(ids, clear_diag=(case.kind == "MPC"))
    assert bits.dtype == np.uint32 and np.array_equal(bits, gold["mask_bits"])
    assert np.array_equal(counts, gold["counts"])
    # factorised int ids have the same equality structure as the string keys
    bits2, counts2 = orc.posmask_packed(orc.factorize_ids(ids), clear_diag=(case.kind == "MPC"))
    assert np.array_equal(bits2, bits) and np.array_equal(counts2, counts)


def test_packed_mask_rectangular_and_offset():
    rng = np.random.default_rng(0)
    ids = rng.integers(0, 40, size=200).astype(np.int32)
    full, _ = orc.posmask_packed(ids, clear_diag=True)
    blk, cnt = orc.posmask_packed(ids[64:128], ids, clear_diag=True, row_offset=64)
    assert np.array_equal(blk, full[64:128])
    assert np.array_equal(cnt, orc.posmask_dense(ids, clear_diag=True)[64:128].sum(1))


@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c.name)
def test_torch_port_matches_golden(case):
    """The timed CPU baseline (kind 'port') computes what the reference computes."""
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    image = torch.tensor(inp["image"], dtype=torch.float64, requires_grad=True)
    if case.kind == "G":
        text = torch.tensor(inp["text"], dtype=torch.float64, requires_grad=True)
        out = orc.global_alignment_loss_port(image, text, inp["ids"], case.tau)
    else:
        out = orc.multi_pos_contra_images_port(image, inp["ids"], case.tau)
    assert tuple(out.shape) == tuple(gold["out_shape"])
    assert abs(out.sum().item() - gold["loss64"]) <= 1e-11 * max(abs(gold["loss64"]), 1.0)
    if out.grad_fn is not None:
        out.backward()
        assert _rel(image.grad.numpy()[gold["rows"]], gold["d_image64"]) < 1e-9


def test_normalize_bwd_matches_autograd():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((9, 33))
    x[4] = 0.0
    g = rng.standard_normal((9, 33))
    xt = torch.tensor(x, requires_grad=True)
    torch.nn.functional.normalize(xt, dim=-1, p=2).backward(torch.tensor(g))
    assert _rel(orc.l2_normalize_bwd(x, g), xt.grad.numpy()) < 1e-12


# ------------------------------------------------------------------ live reference (container only)
@pytest.mark.reference
@pytest.mark.parametrize("n,d,tau", [(8, 16, 0.5), (50, 96, 0.07), (130, 64, 0.5)])
def test_closed_forms_against_live_reference(n, d, tau):
    ids = gc.synth.make_study_ids(n, seed=n)
    x = gc.synth.make_embeddings(ids, d, seed=n + 1).astype(np.float64)
    t = gc.synth.make_embeddings(ids, d, seed=n + 2).astype(np.float64)
    xi = torch.tensor(x, requires_grad=True)
    ti = torch.tensor(t, requires_grad=True)
    ref = ref_shim.global_alignment_loss(xi, ti, ids, tau)
    ref.backward()
    loss, d_i, d_t, _ = orc.g_loss_closed_form(x, t, ids, tau)
    assert abs(loss - ref.item()) < 1e-11 * abs(ref.item())
    assert _rel(d_i, xi.grad.numpy()) < 1e-9 and _rel(d_t, ti.grad.numpy()) < 1e-9
    xm = torch.tensor(x, requires_grad=True)
    refm = ref_shim.multi_pos_contra_images_v0401(xm, ids, tau)
    lossm, dxm = orc.mpc_closed_form(x, ids, tau)
    if refm.grad_fn is None:
        assert lossm is None
    else:
        refm.backward()
        assert abs(lossm - refm.item()) < 1e-11 * abs(refm.item())
        assert _rel(dxm, xm.grad.numpy()) < 1e-9


@pytest.mark.reference
def test_avgpos_variants_against_live_reference():
    ids = gc.synth.make_study_ids(40, seed=3)
    x = gc.synth.make_embeddings(ids, 48, seed=4).astype(np.float64)
    t = gc.synth.make_embeddings(ids, 48, seed=5).astype(np.float64)
    # the reference accumulates into an fp32 tensor([0.0]) (:771, :690), so 1e-7 is its own noise
    ref = ref_shim.avgpos_global_alignment_loss(torch.tensor(x), torch.tensor(t), ids, 0.5)
    assert abs(orc.avgpos_g_loss_closed_form(x, t, ids, 0.5) - ref.item()) < 2e-6
    ref2 = ref_shim.avgpos_multi_pos_contra_images_v0404(torch.tensor(x), ids, 0.5)
    assert abs(orc.avgpos_mpc_closed_form(x, ids, 0.5) - ref2.item()) < 2e-6


@pytest.mark.parametrize("name", ["local_b3_l7_p5_d16", "local_b4_l99_p49_d64", "local_b2_l20_p49_d768_t02"])
def test_local_token_alignment_closed_form_matches_reference_golden(name):
    """SURVEY.md §8 f1 (next row): the oracle of Pretrain.local_text_token_alignment_loss (:506-526), pinned to
    golden vectors recorded from the reference (inputs are regenerated from the seeds in oracle/make_golden.py)."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import make_golden as mg
    v, t, tau = mg.local_inputs(name)
    gold = np.load(os.path.join(gc.GOLDEN_DIR, name + ".npz"))
    loss, d_v, d_t = orc.local_token_alignment_closed_form(v, t, tau)
    assert abs(loss - gold["loss64"]) <= 1e-11 * abs(gold["loss64"])
    assert _rel(d_v[:, :4], gold["d_image64"]) < 1e-9 and _rel(d_t[:, :4], gold["d_text64"]) < 1e-9
    assert abs(np.linalg.norm(d_v) - gold["d_image_norm64"]) <= 1e-9 * gold["d_image_norm64"]
    assert abs(np.linalg.norm(d_t) - gold["d_text_norm64"]) <= 1e-9 * gold["d_text_norm64"]


def test_blockwise_fingerprint_form_equals_the_plain_closed_form():
    """oracle/make_fingerprints.py evaluates the G loss in row blocks (so that N = 32768 fits in memory); it must be
    the same function as g_loss_closed_form, for one-component and two-component keys."""
    import numpy as np
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    from oracle import make_fingerprints as mf
    n, d, tau = 700, 48, 0.3
    ids = synth.make_study_ids(n, seed=3)
    xi = synth.make_embeddings(ids, d, seed=4)
    xt = synth.make_embeddings(ids, d, seed=5)
    xi[5] = 0.0
    want, w_i, w_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    loss, d_i, d_t = mf.g_loss_blockwise(xi, xt, ids, tau, block=97)
    assert abs(loss - want) <= 1e-7 * abs(want)                  # the plain form reproduces the fp32 rounding of 1/c (3e-8)
    assert np.abs(d_i - w_i).max() <= 1e-6 * np.abs(w_i).max() and np.abs(d_t - w_t).max() <= 1e-6 * np.abs(w_t).max()
    pat, stu = synth.make_patient_study_ids(n, seed=6)
    skey = np.array([f"p{p}_s{s}" for p, s in zip(pat, stu)])
    want, w_i, w_t, _ = orc.g_loss_closed_form(xi, xt, skey, tau)
    loss, d_i, d_t = mf.g_loss_blockwise(xi, xt, pat, tau, key2=stu)
    assert abs(loss - want) <= 1e-7 * abs(want)
    assert np.abs(d_i - w_i).max() <= 1e-6 * np.abs(w_i).max()
    fp = mf.fingerprint(loss, d_i, d_t)
    assert len(fp["rows"]) == 16 and len(set(fp["rows"])) == 16


def test_committed_fingerprints_belong_to_the_bench_workloads():
    """The golden fingerprints bench.py compares against are those of ITS workloads (cheap check: cfg2 recomputed)."""
    import json
    import os
    from oracle import make_fingerprints as mf
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    xi, xt, key, key2, tau = mf.workload("cfg2")
    loss, d_i, d_t = mf.g_loss_blockwise(xi, xt, key, tau, key2)
    got = mf.fingerprint(loss, d_i, d_t)
    want = json.load(open(os.path.join(root, "tests", "golden", "fingerprint_cfg2.json")))
    assert abs(got["loss"] - want["loss"]) <= 1e-12 * abs(want["loss"])
    assert got["rows"] == want["rows"]
    assert max(abs(a - b) for a, b in zip(got["d_image_row_norms"], want["d_image_row_norms"])) <= 1e-12
    for cfg in ("cfg3", "cfg4"):
        assert os.path.isfile(os.path.join(root, "tests", "golden", f"fingerprint_{cfg}.json"))


def test_retrieval_oracle_known_answers():
    """f4: exact inner-product search, best first, stable ties, own-group exclusion, padded short lists."""
    import numpy as np
    from oracle import retrieval_oracle as ro
    q = np.array([[1.0, 0.0], [0.0, 1.0]])
    c = np.array([[1.0, 0.0], [0.5, 0.5], [0.0, 2.0], [1.0, 0.0]])
    val, idx = ro.topk_inner_product(q, c, 3)
    assert idx.tolist() == [[0, 3, 1], [2, 1, 0]] and val.tolist() == [[1.0, 1.0, 0.5], [2.0, 0.5, 0.0]]
    val, idx = ro.topk_inner_product(q, c, 3, [0, 1], [0, 0, 1, 1])
    assert idx.tolist() == [[3, 2, -1], [1, 0, -1]] and np.isneginf(val[:, 2]).all()
    val, idx = ro.topk_inner_product(q, c[:2], 4)
    assert idx.shape == (2, 4) and (idx[:, 2:] == -1).all()
