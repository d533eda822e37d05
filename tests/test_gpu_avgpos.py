"""GPU: the 'averaged positive logit' variants of PretrainNewMulPos - global_alignment_loss
(:748-815) and multi_pos_contra_images_v0404 (:670-708) - through the public API against the golden vectors
recorded from the reference and against the fp64 oracle.  fp32 path: loss <= 1e-5 rel, gradients <= 1e-4 rel
(BASELINE.json tolerances; the reference accumulates this loss in fp32, see tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

import evoke_b200
import golden_cases as gc
from evoke_b200 import synth
from gpu_util import DEV, rel_max
from oracle import evoke_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", gc.AVGPOS_CASES, ids=lambda c: c.name)
def test_avgpos_against_reference_golden(case):
    inp = gc.build_inputs(case)
    gold = gc.load_golden(case)
    image = torch.tensor(inp["image"], device=DEV, requires_grad=True)
    if case.kind == "AG":
        text = torch.tensor(inp["text"], device=DEV, requires_grad=True)
        out = evoke_b200.global_alignment_avgpos(image, text, inp["ids"], case.tau)
    else:
        text = None
        out = evoke_b200.multi_pos_contra_images_avgpos(image, inp["ids"], case.tau)
    assert tuple(out.shape) == tuple(gold["out_shape"]) == (1,)
    if gold["empty"]:
        assert out.item() == 0.0 and out.requires_grad and out.grad_fn is None
        return
    out.sum().backward()
    rows = gold["rows"]
    # absolute floors only for the degenerate all-positives case (loss == 0 and zero gradients in exact arithmetic)
    degenerate = case.name == "ag_all_same"
    floor_l, floor_g = (2e-6, 1e-8) if degenerate else (0.0, 0.0)
    assert abs(out.item() - gold["loss64"]) <= 1e-5 * abs(gold["loss64"]) + floor_l
    scale_i = max(np.abs(gold["d_image64"]).max(), 1e-12)
    assert np.abs(image.grad.cpu().numpy()[rows] - gold["d_image64"]).max() <= 1e-4 * scale_i + floor_g
    if text is not None:
        scale_t = max(np.abs(gold["d_text64"]).max(), 1e-12)
        assert np.abs(text.grad.cpu().numpy()[rows] - gold["d_text64"]).max() <= 1e-4 * scale_t + floor_g


def test_avgpos_against_oracle_on_fresh_inputs_and_patched_methods():
    n, d, tau = 200, 96, 0.3
    ids = synth.make_study_ids(n, seed=77)
    xi = synth.make_embeddings(ids, d, seed=1)
    xt = synth.make_embeddings(ids, d, seed=2)

    class _Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.args = {"instance_temp": tau, "region_temp": tau}

    model = evoke_b200.patch_pretrain_newmulpos(_Fake())
    image = torch.tensor(xi, device=DEV, requires_grad=True)
    text = torch.tensor(xt, device=DEV, requires_grad=True)
    loss = model.global_alignment_loss(image, text, synth.ids_as_strings(ids))
    loss.backward()
    want, d_i, d_t = orc.avgpos_g_closed_form(xi, xt, ids, tau)
    assert abs(loss.item() - want) <= 1e-5 * abs(want)
    assert rel_max(image.grad.cpu().numpy(), d_i) <= 1e-4 and rel_max(text.grad.cpu().numpy(), d_t) <= 1e-4
    x = torch.tensor(xi, device=DEV, requires_grad=True)
    lm = model.multi_pos_contra_images_v0404(x, torch.from_numpy(ids).to(DEV))      # device ids
    lm.backward()
    wantm, dx = orc.avgpos_mpc_grad_closed_form(xi, ids, tau)
    assert abs(lm.item() - wantm) <= 1e-5 * abs(wantm)
    assert rel_max(x.grad.cpu().numpy(), dx) <= 1e-4
