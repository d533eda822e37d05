"""Import shim that executes the reference's OWN loss methods, unmodified, in the build
container.  TEST INFRASTRUCTURE ONLY; it reads /root/reference, which does not exist on the
GPU box, so nothing on the ``-m gpu`` / smoke / bench path may import it.  It is used by
``oracle/make_golden.py`` (fixture generation) and by the container-only tests that
cross-check the oracle against the live reference.

Recipe: the reference module imports siblings that need packages this image
does not have (transformers 4.23 private APIs, matplotlib, pycocoevalcap, radgraph, nltk).
The loss methods themselves need only torch/numpy/einops, so the seven sibling modules are
replaced by MagicMocks before import and the methods are called UNBOUND with a stand-in
``self`` that carries ``args``.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace
from unittest.mock import MagicMock

REFERENCE_ROOT = os.environ.get("EVOKE_REFERENCE_ROOT", "/root/reference")

_STUBBED = [
    "models.language_encoder.language_model",
    "models.language_encoder.bert_model",
    "models.vision_encoder.vit",
    "modules.base_cmn",
    "modules.encoder_decoder",
    "modules.utils_v0511",
    "modules.visual_extractor",
]

_mod = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "model_pretrain_finetune_v0520.py"))


def load():
    """Return the reference module models.model_pretrain_finetune_v0520."""
    global _mod
    if _mod is not None:
        return _mod
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True      # the mount is read-only
    for name in _STUBBED:
        sys.modules.setdefault(name, MagicMock())
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import models.model_pretrain_finetune_v0520 as m  # noqa: E402
    _mod = m
    return m


def fake_self(instance_temp: float = 0.5, region_temp: float = 0.5):
    return SimpleNamespace(args={"instance_temp": instance_temp, "region_temp": region_temp})


def global_alignment_loss(image, text, ids, temp: float = 0.5):
    return load().Pretrain.global_alignment_loss(fake_self(instance_temp=temp), image, text, ids)


def multi_pos_contra_images_v0401(x, ids, temp: float = 0.5):
    return load().Pretrain.multi_pos_contra_images_v0401(fake_self(region_temp=temp), x, ids)


def avgpos_global_alignment_loss(image, text, ids, temp: float = 0.5):
    return load().PretrainNewMulPos.global_alignment_loss(fake_self(instance_temp=temp), image, text, ids)


def avgpos_multi_pos_contra_images_v0404(x, ids, temp: float = 0.5):
    return load().PretrainNewMulPos.multi_pos_contra_images_v0404(fake_self(region_temp=temp), x, ids)


def local_text_token_alignment_loss(local_image, local_text, temp: float = 0.5):
    """Pretrain.local_text_token_alignment_loss (:506-526): the next row of the hot path."""
    return load().Pretrain.local_text_token_alignment_loss(fake_self(region_temp=temp), local_image, local_text)


# ------------------------------------------------------------------------------------- f2 / f3 (next rows)
_utils_ns = None


def utils_classes():
    """The reference's ``ScaledDotProductAttention`` (modules/utils_v0511.py:211-279) and
    ``VisualProjectionHeadPretrain`` / ``TextProjectionHeadPretrain`` (:131-168), executed unmodified.  The module
    itself cannot be imported here (it needs matplotlib / cv2 / pycocoevalcap at import time), so the three class
    definitions are cut out of the file with ``ast`` at run time and executed in a namespace that provides what
    they use (torch, nn, np).  Nothing is copied into the repository."""
    global _utils_ns
    if _utils_ns is not None:
        return _utils_ns
    import ast
    import numpy as np
    import torch
    from torch import nn
    path = os.path.join(REFERENCE_ROOT, "modules", "utils_v0511.py")
    src = open(path).read()
    tree = ast.parse(src)
    want = {"ScaledDotProductAttention", "VisualProjectionHeadPretrain", "TextProjectionHeadPretrain"}
    ns = {"torch": torch, "nn": nn, "np": np, "F": torch.nn.functional}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in want:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    _utils_ns = SimpleNamespace(**{k: ns[k] for k in want})
    return _utils_ns


def make_fusion_self(visual_dim: int, output_dim: int, seed: int = 0):
    """Stand-in ``self`` for Pretrain.multiview_fusion (:456-484) holding the four sub-modules it touches, built from
    the reference's own classes exactly as Pretrain.__init__ does (:347-357)."""
    import torch
    from torch import nn
    u = utils_classes()
    torch.manual_seed(seed)
    holder = nn.Module()
    holder.layer_norm_1 = nn.LayerNorm(visual_dim)
    holder.layer_norm_2 = nn.LayerNorm(visual_dim)
    holder.visual_head = u.VisualProjectionHeadPretrain(visual_dim, output_dim=output_dim, hidden_dim=output_dim)
    holder.multiview_cross_attention = u.ScaledDotProductAttention(visual_dim, visual_dim, visual_dim, h=8)
    # the reference initialises every Linear with std 0.001 (utils_v0511.py:236-247): the attention branch would be
    # numerically invisible next to the residual; scale the weights up so the fixture actually exercises it
    with torch.no_grad():
        for p in holder.multiview_cross_attention.parameters():
            if p.dim() == 2:
                p.mul_(60.0)
            else:
                p.normal_(0.0, 0.05)
        for m in (holder.layer_norm_1, holder.layer_norm_2):
            m.weight.normal_(1.0, 0.1)
            m.bias.normal_(0.0, 0.1)
    return holder


def multiview_fusion(holder, global_embed, local_embed, patient_ids, batch_size: int):
    """Pretrain.multiview_fusion (:456-484) called unbound on the stand-in self."""
    return load().Pretrain.multiview_fusion(holder, global_embed, local_embed, patient_ids, batch_size)
