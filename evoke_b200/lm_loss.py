"""Boundary only: the reference's modules/loss.py holds the stage-2 language-model loss
(`LanguageModelCriterion.forward(input, target, mask)` :9-16, `compute_lm_loss(output,
reports_ids, reports_masks)` :19-22).  It is not on the contrastive hot path and nothing here
is accelerated; the two names are provided with unchanged signatures so that a tree that
imports its losses from one module keeps working after switching to evoke_b200.
"""
import torch
import torch.nn as nn


class LanguageModelCriterion(nn.Module):
    def forward(self, input, target, mask):
        steps = input.size(1)
        target, mask = target[:, :steps], mask[:, :steps]
        picked = input.gather(2, target.long().unsqueeze(2)).squeeze(2)
        return -(picked * mask).sum() / mask.sum()


def compute_lm_loss(output, reports_ids, reports_masks):
    # position 0 is the [cls]/BOS token and is not predicted
    return LanguageModelCriterion()(output, reports_ids[:, 1:], reports_masks[:, 1:]).mean()
