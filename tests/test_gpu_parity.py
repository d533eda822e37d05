"""T2/T3 (GPU): the loss + gradients through the public API against the golden vectors
recorded from the reference (tests/golden) and against the fp64 oracle on fresh inputs.

Tolerances (BASELINE.json north_star): mask bit-exact (test_gpu_kernels.py); fp32 mode loss
<= 1e-5 rel, gradients <= 1e-4 rel (max-norm); bf16 mode gradients <= 2e-2 rel.
"""
import numpy as np
import pytest
import torch

import evoke_b200
import golden_cases as gc
from evoke_b200 import functional as Fn
from evoke_b200 import synth
from gpu_util import DEV, TOL, check_against_golden, rel_max
from oracle import evoke_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c.name)
def test_small_path_fp32_against_golden(case):
    check_against_golden(case, "fp32", "small")


@pytest.mark.parametrize("a_major,b_major", [(0, 0), (0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("m,n,k,splits", [(128, 256, 64, 1), (300, 520, 200, 1), (256, 768, 1024, 3)])
def test_tc_main_loop_gemm(a_major, b_major, m, n, k, splits):
    """The tcgen05 main loop on its own: C = A B^T for every operand major-ness."""
    torch.manual_seed(m + n + k)
    a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
    b = torch.randn(n, k, device=DEV).to(torch.bfloat16)
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8)).contiguous()
    a_st = pad(a.t().contiguous()) if a_major else pad(a)
    b_st = pad(b.t().contiguous()) if b_major else pad(b)
    c = Fn.tc_gemm_probe(a_st, b_st, a_major, b_major, m, n, k, variant=0, splits=splits)
    want = a.float() @ b.float().t()
    assert rel_max(c.cpu().numpy(), want.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("case", gc.CASES, ids=lambda c: c.name)
def test_tc_path_fp32_mode_against_golden(case):
    check_against_golden(case, "fp32", "tc")


@pytest.mark.parametrize("case", [c for c in gc.CASES if c.zero_row < 0], ids=lambda c: c.name)
def test_tc_path_bf16_mode_against_golden(case):
    check_against_golden(case, "bf16", "tc")


@pytest.mark.parametrize("n,d,tau", [(2048, 768, 0.5), (1500, 512, 0.07)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_tc_path_against_oracle_mid_size(n, d, tau, precision):
    ids = synth.make_study_ids(n, seed=n)
    xi = synth.make_embeddings(ids, d, seed=1)
    xt = synth.make_embeddings(ids, d, seed=2)
    image = torch.tensor(xi, device=DEV, requires_grad=True)
    text = torch.tensor(xt, device=DEV, requires_grad=True)
    out = evoke_b200.global_alignment(image, text, ids, tau, precision=precision, path="tc")
    out.backward()
    loss, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    tol = TOL[precision]
    assert abs(out.item() - loss) <= tol["loss"] * abs(loss)
    assert rel_max(image.grad.cpu().numpy(), d_i) <= tol["grad"]
    assert rel_max(text.grad.cpu().numpy(), d_t) <= tol["grad"]
    x = torch.tensor(xi, device=DEV, requires_grad=True)
    outm = evoke_b200.multi_pos_contra_images(x, ids, tau, precision=precision, path="tc")
    outm.backward()
    lossm, dxm = orc.mpc_closed_form(xi, ids, tau)
    assert abs(outm.item() - lossm) <= tol["loss"] * abs(lossm)
    assert rel_max(x.grad.cpu().numpy(), dxm) <= tol["grad"]


def test_small_and_tc_paths_agree():
    ids = synth.make_study_ids(400, seed=1)
    xi = torch.tensor(synth.make_embeddings(ids, 256, seed=3), device=DEV)
    xt = torch.tensor(synth.make_embeddings(ids, 256, seed=4), device=DEV)
    a = evoke_b200.global_alignment(xi, xt, ids, 0.5, path="small")
    b = evoke_b200.global_alignment(xi, xt, ids, 0.5, path="tc", precision="fp32")
    assert abs(a.item() - b.item()) < 1e-5 * abs(a.item())


# ------------------------------------------------------------------------------------- drop-in (T3)
class _FakePretrain(torch.nn.Module):
    """Stands in for the reference's Pretrain: only `args` and the two method names matter."""

    def __init__(self):
        super().__init__()
        self.args = {"instance_temp": 0.5, "region_temp": 0.5}

    def global_alignment_loss(self, a, b, ids):      # to be replaced
        raise AssertionError("not patched")

    def multi_pos_contra_images_v0401(self, a, ids):
        raise AssertionError("not patched")


def test_drop_in_patch_with_reference_argument_types():
    case = gc.BY_NAME["g_cfg1"]
    inp = gc.build_inputs(case)
    model = evoke_b200.patch_pretrain(_FakePretrain())
    # strided embeddings exactly as the projection head hands them over (:484, :399)
    b, d, p1 = case.n, case.d, 50
    head_i = torch.zeros(b, d, p1, device=DEV)
    head_t = torch.zeros(b, d, p1, device=DEV)
    head_i[:, :, 0] = torch.tensor(inp["image"], device=DEV)
    head_t[:, :, 0] = torch.tensor(inp["text"], device=DEV)
    head_i.requires_grad_(True)
    head_t.requires_grad_(True)
    img = head_i.permute(0, 2, 1)[:, 0, :]
    txt = head_t.permute(0, 2, 1)[:, 0, :]
    ids = inp["ids"]                                   # numpy <U strings, longer than B
    assert ids.dtype.kind == "U" and len(ids) > b
    loss = model.global_alignment_loss(img, txt, ids)
    views = torch.tensor(gc.build_inputs(gc.BY_NAME["mpc_cfg1"])["image"], device=DEV, requires_grad=True)
    mpc = model.multi_pos_contra_images_v0401(views, gc.build_inputs(gc.BY_NAME["mpc_cfg1"])["ids"])
    total = loss + mpc                                 # summed like all_loss (:563)
    total.backward()
    g = gc.load_golden(case)
    gm = gc.load_golden(gc.BY_NAME["mpc_cfg1"])
    assert abs(loss.cpu().detach().item() - g["loss64"]) < 1e-5 * g["loss64"]      # trainer_v0401.py:264
    assert abs(mpc.item() - gm["loss64"]) < 1e-5 * gm["loss64"]
    assert rel_max(head_i.grad[:, :, 0].cpu().numpy(), g["d_image64"]) < 1e-4
    assert rel_max(views.grad.cpu().numpy(), gm["d_image64"]) < 1e-4
    with torch.no_grad():
        again = model.global_alignment_loss(img, txt, ids)
    assert not again.requires_grad and abs(again.item() - loss.item()) < 1e-6


def test_class_level_patch_and_device_int_ids():
    evoke_b200.patch_pretrain(_FakePretrain, precision="fp32")
    model = _FakePretrain()
    ids = synth.make_study_ids(96, seed=2)
    xi = torch.tensor(synth.make_embeddings(ids, 64, seed=5), device=DEV, requires_grad=True)
    xt = torch.tensor(synth.make_embeddings(ids, 64, seed=6), device=DEV, requires_grad=True)
    ids_dev = torch.from_numpy(ids.astype(np.int64)).to(DEV)         # int64 device tensor
    loss = model.global_alignment_loss(xi, xt, ids_dev)
    want, _, _, _ = orc.g_loss_closed_form(xi.detach().cpu().numpy(), xt.detach().cpu().numpy(), ids, 0.5)
    assert abs(loss.item() - want) < 1e-5 * want
    mpc = model.multi_pos_contra_images_v0401(xi, ids_dev)
    wantm, _ = orc.mpc_closed_form(xi.detach().cpu().numpy(), ids, 0.5)
    assert abs(mpc.item() - wantm) < 1e-5 * wantm


def test_upstream_gradient_scale_and_reuse():
    ids = synth.make_study_ids(50, seed=3)
    xi = torch.tensor(synth.make_embeddings(ids, 32, seed=7), device=DEV, requires_grad=True)
    xt = torch.tensor(synth.make_embeddings(ids, 32, seed=8), device=DEV, requires_grad=True)
    (3.0 * evoke_b200.global_alignment(xi, xt, ids, 0.5)).backward()
    _, d_i, _, _ = orc.g_loss_closed_form(xi.detach().cpu().numpy(), xt.detach().cpu().numpy(), ids, 0.5)
    assert rel_max(xi.grad.cpu().numpy(), 3.0 * d_i) < 1e-4


def test_temperature_out_of_range_is_an_error_on_the_tc_path():
    ids = np.arange(600) // 2
    x = torch.randn(600, 64, device=DEV)
    with pytest.raises(ValueError, match="tau"):
        evoke_b200.global_alignment(x, x, ids, 0.01, path="tc")


def test_cuda_graph_step_matches_eager_and_tracks_new_inputs():
    n, d = 1024, 256
    g = evoke_b200.GraphedGlobalAlignment(n, d, 0.5, precision="fp32", path="tc").capture()
    for seed in (1, 2):
        ids = synth.make_study_ids(n, seed=seed)
        xi = synth.make_embeddings(ids, d, seed=seed + 10)
        xt = synth.make_embeddings(ids, d, seed=seed + 20)
        g.load(torch.tensor(xi, device=DEV), torch.tensor(xt, device=DEV), torch.from_numpy(ids).to(DEV))
        loss = g.step()
        want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, 0.5)
        assert abs(loss.item() - want) <= 1e-5 * abs(want)
        assert rel_max(g.image.grad.cpu().numpy(), d_i) <= 1e-4
        assert rel_max(g.text.grad.cpu().numpy(), d_t) <= 1e-4


# ------------------------------------------------------------------------------------- drop-in graph cache
@pytest.mark.parametrize("zero_copy", [False, True], ids=["static-inputs", "zero-copy"])
@pytest.mark.parametrize("precision,path,n,d", [("bf16", "tc", 1536, 256), ("fp32", "tc", 700, 128), ("fp32", "small", 64, 768)])
def test_graph_cached_call_equals_the_eager_launch_sequence(precision, path, n, d, zero_copy, monkeypatch):
    """global_alignment(graph=True) replays two captured graphs over static buffers; the kernels are those of the eager
    call, so loss and gradients agree to fp32 rounding (not bitwise: the split-K reduce-adds land in L2 in any order, and
    a strided input is gathered into the static buffer before K1 instead of being read through K1's strided loader) -
    on fresh data of the same signature too, with strided inputs, and under no_grad."""
    from evoke_b200 import graphs
    graphs.clear_graph_cache()
    monkeypatch.setattr(graphs, "ZERO_COPY", zero_copy)      # both forms of the cache (see evoke_b200/graphs.py)
    for seed in (1, 2, 3):
        ids = synth.make_study_ids(n, seed=seed)
        xi = synth.make_embeddings(ids, d, seed=seed + 10)
        xt = synth.make_embeddings(ids, d, seed=seed + 20)
        res = []
        for graph in (False, True):
            if seed == 3:      # strided head views
                full_i = torch.zeros((n, d, 3), device=DEV)
                full_i[:, :, 0] = torch.tensor(xi, device=DEV)
                full_i.requires_grad_(True)
                image, leaf_i = full_i.permute(0, 2, 1)[:, 0, :], full_i
            else:
                image = leaf_i = torch.tensor(xi, device=DEV, requires_grad=True)
            text = torch.tensor(xt, device=DEV, requires_grad=True)
            out = evoke_b200.global_alignment(image, text, ids, 0.5, precision=precision, path=path, graph=graph)
            (out * 3.0).backward()                                  # a non-unit upstream gradient
            gi = leaf_i.grad[:, :, 0] if seed == 3 else leaf_i.grad
            res.append((out.item(), gi.clone(), text.grad.clone()))
        assert abs(res[0][0] - res[1][0]) <= 1e-6 * abs(res[0][0])
        for a, b in ((res[0][1], res[1][1]), (res[0][2], res[1][2])):
            assert rel_max(b.cpu().numpy(), a.cpu().numpy()) <= 1e-5
    assert len(graphs._CACHE) == 1                                  # contiguous and strided inputs share one signature
    with torch.no_grad():                                           # forward-only entry
        a = evoke_b200.global_alignment(image.detach(), text.detach(), ids, 0.5, precision=precision, path=path, graph=True)
    assert abs(a.item() - res[1][0]) <= 1e-6 * abs(res[1][0]) and not a.requires_grad and len(graphs._CACHE) == 2


def test_graph_cached_call_rejects_a_stale_backward():
    from evoke_b200 import graphs
    graphs.clear_graph_cache()
    n, d = 1024, 128
    ids = synth.make_study_ids(n, seed=5)
    x = torch.tensor(synth.make_embeddings(ids, d, seed=1), device=DEV, requires_grad=True)
    y = torch.tensor(synth.make_embeddings(ids, d, seed=2), device=DEV, requires_grad=True)
    first = evoke_b200.global_alignment(x, y, ids, 0.5, precision="bf16", path="tc", graph=True)
    second = evoke_b200.global_alignment(x, y, ids, 0.5, precision="bf16", path="tc", graph=True)
    with pytest.raises(RuntimeError, match="stale"):
        first.backward()
    second.backward()
    with pytest.raises(RuntimeError):
        second.backward()
