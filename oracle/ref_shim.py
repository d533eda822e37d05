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
