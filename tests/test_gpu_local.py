"""GPU (the first "next" row): Pretrain.local_text_token_alignment_loss (:506-526) through the
public API against the golden vectors recorded from the reference (oracle/make_golden.py LOCAL_CASES) and against
the fp64 oracle on fresh inputs.  fp32 path: loss <= 1e-5 rel, gradients <= 1e-4 rel (BASELINE.json)."""
import os
import sys

import numpy as np
import pytest
import torch

import evoke_b200
import golden_cases as gc
from gpu_util import DEV, rel_max
from oracle import evoke_oracle as orc

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["local_b3_l7_p5_d16", "local_b4_l99_p49_d64", "local_b2_l20_p49_d768_t02"])
def test_local_token_alignment_against_reference_golden(name):
    import make_golden as mg
    v, t, tau = mg.local_inputs(name)
    gold = np.load(os.path.join(gc.GOLDEN_DIR, name + ".npz"))
    image = torch.tensor(v, device=DEV, requires_grad=True)
    text = torch.tensor(t, device=DEV, requires_grad=True)
    loss = evoke_b200.local_text_token_alignment(image, text, tau)
    assert loss.shape == ()
    loss.backward()
    assert abs(loss.item() - gold["loss64"]) <= 1e-5 * abs(gold["loss64"])
    assert rel_max(image.grad.cpu().numpy()[:, :4], gold["d_image64"]) <= 1e-4
    assert rel_max(text.grad.cpu().numpy()[:, :4], gold["d_text64"]) <= 1e-4
    assert abs(float(image.grad.double().norm()) - gold["d_image_norm64"]) <= 5e-4 * gold["d_image_norm64"]
    assert abs(float(text.grad.double().norm()) - gold["d_text_norm64"]) <= 5e-4 * gold["d_text_norm64"]


def test_local_token_alignment_reference_shapes_strided_inputs_and_patched_method():
    """B=32, L=99, P=49, D=768 as in the reference's 224-px run; inputs are permuted views as the projection heads
    hand them over (utils_v0511.py:145-148); method rebound on a stand-in module."""
    rng = np.random.default_rng(9)
    b, l, p, d, tau = 32, 99, 49, 768, 0.5
    v = rng.standard_normal((b, p, d)).astype(np.float32)
    t = rng.standard_normal((b, l, d)).astype(np.float32)

    class _Fake(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.args = {"instance_temp": tau, "region_temp": tau}

    model = evoke_b200.patch_pretrain(_Fake(), local_tokens=True)
    head_v = torch.tensor(v, device=DEV).permute(0, 2, 1).contiguous().requires_grad_(True)      # [B, D, P]
    head_t = torch.tensor(t, device=DEV).permute(0, 2, 1).contiguous().requires_grad_(True)
    loss = model.local_text_token_alignment_loss(head_v.permute(0, 2, 1), head_t.permute(0, 2, 1))
    (2.0 * loss).backward()
    want, d_v, d_t = orc.local_token_alignment_closed_form(v, t, tau)
    assert abs(loss.item() - want) <= 1e-5 * abs(want)
    assert rel_max(head_v.grad.permute(0, 2, 1).cpu().numpy(), 2.0 * d_v) <= 1e-4
    assert rel_max(head_t.grad.permute(0, 2, 1).cpu().numpy(), 2.0 * d_t) <= 1e-4


def test_local_token_alignment_batched_small_path_and_long_sequences(monkeypatch):
    """The fallback (batched small-path kernels) on the same inputs, l > 128 which always takes it, and D % 4 != 0,
    which takes the scalar-load variants of the attention / token-similarity kernels (no 128-bit loads, no cluster)."""
    from evoke_b200 import functional as Fn
    rng = np.random.default_rng(5)
    for (b, l, p, d, force) in ((3, 40, 20, 96, True), (2, 150, 49, 64, False), (2, 20, 9, 30, False), (3, 33, 7, 50, False)):
        monkeypatch.setattr(Fn, "F1_TOKEN_SIM", not force)
        v = rng.standard_normal((b, p, d)).astype(np.float32)
        t = rng.standard_normal((b, l, d)).astype(np.float32)
        image = torch.tensor(v, device=DEV, requires_grad=True)
        text = torch.tensor(t, device=DEV, requires_grad=True)
        loss = evoke_b200.local_text_token_alignment(image, text, 0.5)
        loss.backward()
        want, d_v, d_t = orc.local_token_alignment_closed_form(v, t, 0.5)
        assert abs(loss.item() - want) <= 1e-5 * abs(want)
        assert rel_max(image.grad.cpu().numpy(), d_v) <= 1e-4 and rel_max(text.grad.cpu().numpy(), d_t) <= 1e-4


def test_local_token_alignment_cuda_graph_tracks_new_inputs():
    b, p, l, d, tau = 4, 49, 30, 128, 0.5
    g = evoke_b200.GraphedLocalTokenAlign(b, p, l, d, tau).capture()
    rng = np.random.default_rng(3)
    for _ in range(2):
        v = rng.standard_normal((b, p, d)).astype(np.float32)
        t = rng.standard_normal((b, l, d)).astype(np.float32)
        g.load(torch.tensor(v, device=DEV), torch.tensor(t, device=DEV))
        loss = g.step()
        want, d_v, d_t = orc.local_token_alignment_closed_form(v, t, tau)
        assert abs(loss.item() - want) <= 1e-5 * abs(want)
        assert rel_max(g.image.grad.cpu().numpy(), d_v) <= 1e-4
        assert rel_max(g.text.grad.cpu().numpy(), d_t) <= 1e-4
