"""evoke_b200 — B200-native (sm_100a) implementation of EVOKE's multi-view, multi-positive
image-text contrastive objective, forward and backward, behind the reference's own method
signatures.  See DESIGN.md / INTEGRATION.md.

Importing the package does not load the CUDA library; the first loss call does, and fails
loudly if it is missing (there is no CPU or PyTorch fallback).
"""
from .loss import (ContrastiveObjective, global_alignment, global_alignment_avgpos, global_alignment_loss,
                   local_text_token_alignment, local_text_token_alignment_loss,
                   multi_pos_contra_images, multi_pos_contra_images_avgpos, multi_pos_contra_images_v0401,
                   multi_pos_contra_images_v0404, patch_pretrain, patch_pretrain_newmulpos)
from .lm_loss import LanguageModelCriterion, compute_lm_loss
from .graphs import GraphedGlobalAlignment, GraphedLocalTokenAlign

__all__ = [
    "ContrastiveObjective", "global_alignment", "global_alignment_loss", "multi_pos_contra_images",
    "multi_pos_contra_images_v0401", "patch_pretrain", "global_alignment_avgpos", "multi_pos_contra_images_avgpos",
    "multi_pos_contra_images_v0404", "patch_pretrain_newmulpos", "local_text_token_alignment",
    "local_text_token_alignment_loss", "GraphedGlobalAlignment", "GraphedLocalTokenAlign", "LanguageModelCriterion", "compute_lm_loss",
]
__version__ = "0.1.0"
