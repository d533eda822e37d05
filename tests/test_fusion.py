"""f2 / f3: MultiviewFusion and the permute-free projection head against the reference's own modules, executed
unmodified through oracle/ref_shim.py (container only: marker `reference`), and against the committed golden
fixture recorded from them (tests/golden/fusion_m9.npz, runs anywhere; on the GPU with -m gpu)."""
import os

import numpy as np
import pytest
import torch

from evoke_b200 import fusion

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "fusion_m9.npz")


def _inputs(m, p, d, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(m, d, generator=g), torch.randn(m, p, d, generator=g)


IDS = np.array(["p1_s1", "p2_s1", "p3_s1", "p4_s1", "p5_s9", "p1_s1", "p1_s1", "p3_s1", "p7_s2"])    # 5 anchors + 4 aux views


@pytest.mark.reference
@pytest.mark.parametrize("train", [False, True])
def test_multiview_fusion_equals_the_reference_module(train):
    from oracle import ref_shim
    d, d_out, p, b = 32, 16, 5, 5
    ref = ref_shim.make_fusion_self(d, d_out, seed=3)
    mine = fusion.MultiviewFusion(d, d_out, heads=8)
    missing = mine.load_state_dict(ref.state_dict(), strict=True)       # identical keys and shapes
    assert not missing.missing_keys and not missing.unexpected_keys
    for mod in (ref, mine):
        mod.train(train)
        mod.multiview_cross_attention.dropout.p = 0.0                   # dropout is random: parity needs it off
    gi, li = _inputs(len(IDS), p, d, seed=1)
    outs = []
    for mod, fn in ((ref, lambda *a: ref_shim.multiview_fusion(ref, *a)), (mine, mine)):
        g = gi.clone().requires_grad_(True)
        l = li.clone().requires_grad_(True)
        o0, o1 = fn(g, l, IDS, b)
        (o0.square().sum() + (o1 * 0.3).sum()).backward()
        outs.append((o0, o1, g.grad, l.grad, {k: v.grad.clone() for k, v in mod.named_parameters()},
                     mod.visual_head.head[1].running_mean.clone()))
    (r0, r1, rg, rl, rp, rm), (m0, m1, mg, ml, mp, mm) = outs
    assert torch.allclose(m0, r0, atol=2e-5, rtol=1e-4) and torch.allclose(m1, r1, atol=2e-5, rtol=1e-4)
    assert torch.allclose(mg, rg, atol=1e-5, rtol=1e-3) and torch.allclose(ml, rl, atol=1e-5, rtol=1e-3)
    for k in rp:
        assert torch.allclose(mp[k], rp[k], atol=1e-5, rtol=2e-3), k
    assert torch.allclose(mm, rm, atol=1e-6)
    # rows without another view pass through LayerNorm and the head only; aux-view gradients flow through K / V = none
    assert float(rg[5:].abs().max()) == 0.0 and float(mg[5:].abs().max()) == 0.0


@pytest.mark.reference
def test_permute_free_head_equals_the_reference_head():
    from oracle import ref_shim
    u = ref_shim.utils_classes()
    torch.manual_seed(0)
    ref = u.TextProjectionHeadPretrain(24, hidden_dim=12, output_dim=12)
    mine = fusion.convert_head(ref)
    x = torch.randn(6, 7, 24)
    for train in (True, False):
        ref.train(train)
        mine.train(train)
        a = ref(x)
        state = {k: v.clone() for k, v in ref.state_dict().items()}
        ref.load_state_dict(state)
        b = mine(x)
        assert torch.allclose(a, b, atol=1e-6)
    assert mine(x)[:, 0, :].stride(-1) == 1 and ref(x)[:, 0, :].stride(-1) != 1      # the point of the layout


def _run_golden(device):
    gold = np.load(GOLD)
    d, d_out, p, b = int(gold["d"]), int(gold["d_out"]), int(gold["p"]), int(gold["b"])
    mine = fusion.MultiviewFusion(d, d_out, heads=8).to(device)
    sd = {k[3:]: torch.tensor(gold[k]) for k in gold.files if k.startswith("sd.")}
    mine.load_state_dict(sd, strict=True)
    mine.train(True)
    mine.multiview_cross_attention.dropout.p = 0.0
    g = torch.tensor(gold["global"], device=device, requires_grad=True)
    l = torch.tensor(gold["local"], device=device, requires_grad=True)
    o0, o1 = mine(g, l, gold["ids"], b)
    (o0.square().sum() + (o1 * 0.3).sum()).backward()
    tol = dict(atol=3e-5, rtol=2e-4)
    assert np.allclose(o0.detach().cpu().numpy(), gold["out_global"], **tol)
    assert np.allclose(o1.detach().cpu().numpy(), gold["out_local"], **tol)
    assert np.allclose(g.grad.cpu().numpy(), gold["d_global"], atol=2e-5, rtol=2e-3)
    assert np.allclose(l.grad.cpu().numpy(), gold["d_local"], atol=2e-5, rtol=2e-3)
    assert np.allclose(mine.multiview_cross_attention.fc_k.weight.grad.cpu().numpy(), gold["d_fc_k"], atol=2e-5, rtol=5e-3)


def test_multiview_fusion_against_the_golden_fixture_cpu():
    _run_golden("cpu")


@pytest.mark.gpu
def test_multiview_fusion_against_the_golden_fixture_gpu():
    torch.backends.cuda.matmul.allow_tf32 = False
    _run_golden("cuda")


def test_partner_lists_follow_the_reference_order():
    assert fusion.partner_lists(IDS, 5) == [[5, 6], [], [7], [], []]
    assert fusion.partner_lists(torch.tensor([3, 3, 1, 3]), 2) == [[1, 3], [0, 3]]


@pytest.mark.reference
def test_patch_pretrain_fusion_rebinds_the_method_on_a_reference_style_module():
    """patch_pretrain(..., fusion=True): the reference's own sub-modules, the loop-free forward."""
    import evoke_b200
    from oracle import ref_shim
    d, d_out, p, b = 32, 16, 5, 5
    holder = ref_shim.make_fusion_self(d, d_out, seed=5)
    holder.args = {"instance_temp": 0.5, "region_temp": 0.5}
    holder.eval()
    gi, li = _inputs(len(IDS), p, d, seed=4)
    want = ref_shim.multiview_fusion(holder, gi, li, IDS, b)
    evoke_b200.patch_pretrain(holder, fusion=True)
    got = holder.multiview_fusion(gi, li, IDS, b)
    assert torch.allclose(got[0], want[0], atol=2e-5, rtol=1e-4) and torch.allclose(got[1], want[1], atol=2e-5, rtol=1e-4)
