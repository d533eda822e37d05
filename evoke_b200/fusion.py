"""f2 / f3: the producers of the embeddings the contrastive path consumes.

f2  ``Pretrain.multiview_fusion`` (reference models/model_pretrain_finetune_v0520.py:456-484) with
    ``ScaledDotProductAttention`` (modules/utils_v0511.py:211-279): for every anchor image that has other views of
    the same study in the batch, its 1 + P tokens attend (8 heads, d_k = d_v = visual_dim per head) over the tokens of
    those views, followed by residual + LayerNorm and the visual projection head.  The reference loops over the B
    anchors in Python, concatenates the partner views per anchor and pushes each through the four
    visual_dim <-> 8*visual_dim Linear layers separately - every view's K/V projection is recomputed for each anchor
    that uses it.  Here the projections run ONCE over the distinct images (queries over the anchors that have
    partners), and the attention of all anchors is one padded, masked batched product.  Same parameters, same
    state_dict keys, same results.
f3  ``VisualProjectionHeadPretrain`` / ``TextProjectionHeadPretrain`` (utils_v0511.py:131-168): Conv1d(k=1) -> BatchNorm1d
    -> ReLU -> Conv1d(k=1) between two permutes.  A 1x1 convolution over [B, C, T] IS a Linear over the last axis of
    [B, T, C]: without the permutes the output is contiguous [B, T, C_out], so the global embedding ``[:, 0, :]`` the
    losses receive has unit feature stride and K1 reads it with 128-bit loads (the reference hands over a view whose
    feature stride is 1 + P, :484/:399).  Parameters keep the Conv1d shapes, so reference checkpoints load as they are.

These are the callers of the hot path, not the path: the dense products are plain library GEMMs (torch / cuBLAS), as
in the reference; what is B200-specific here is the data movement (no per-anchor concatenation, no recomputation, no
strided hand-over).  The module therefore runs wherever torch runs.
"""
from __future__ import annotations

import math
from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn


class ProjectionHeadPretrain(nn.Module):
    """Conv1d(k=1) -> BatchNorm1d -> ReLU -> Conv1d(k=1) over the channel axis of [B, T, C_in] (utils_v0511.py:131-168),
    without the permutes.  ``head`` has the reference's layout (indices 0, 1, 3 carry parameters)."""

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int) -> None:
        super().__init__()
        self.head = nn.Sequential(
            nn.Conv1d(input_dim, hidden_dim, kernel_size=1, stride=1, padding=0),
            nn.BatchNorm1d(hidden_dim),
            nn.ReLU(inplace=True),
            nn.Conv1d(hidden_dim, output_dim, kernel_size=1, stride=1, padding=0),
        )

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        b, t, c = x.shape
        c1, bn, _, c2 = self.head
        h = F.linear(x.reshape(b * t, c), c1.weight.squeeze(-1), c1.bias)              # [B*T, hidden]
        # BatchNorm1d over (B, T) per channel == batch_norm of the [B*T, hidden] matrix (same running-stat updates)
        if bn.training and bn.track_running_stats and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        momentum = bn.momentum if bn.momentum is not None else (
            1.0 / float(bn.num_batches_tracked) if bn.training and bn.track_running_stats else 0.0)
        h = F.batch_norm(h, bn.running_mean if (not bn.training or bn.track_running_stats) else None,
                         bn.running_var if (not bn.training or bn.track_running_stats) else None, bn.weight, bn.bias,
                         bn.training or (bn.running_mean is None and bn.running_var is None), momentum, bn.eps)
        h = F.relu(h, inplace=True)
        out = F.linear(h, c2.weight.squeeze(-1), c2.bias)
        return out.view(b, t, -1)                                                        # contiguous: [:, 0, :] has unit stride


VisualProjectionHeadPretrain = ProjectionHeadPretrain
TextProjectionHeadPretrain = ProjectionHeadPretrain


class ScaledDotProductAttention(nn.Module):
    """Parameters of the reference's attention block (utils_v0511.py:211-247; same names, same initialisation)."""

    def __init__(self, d_model: int, d_k: int, d_v: int, h: int, dropout: float = 0.1):
        super().__init__()
        self.fc_q = nn.Linear(d_model, h * d_k)
        self.fc_k = nn.Linear(d_model, h * d_k)
        self.fc_v = nn.Linear(d_model, h * d_v)
        self.fc_o = nn.Linear(h * d_v, d_model)
        self.dropout = nn.Dropout(dropout)
        self.d_model, self.d_k, self.d_v, self.h = d_model, d_k, d_v, h
        for m in (self.fc_q, self.fc_k, self.fc_v, self.fc_o):
            nn.init.normal_(m.weight, std=0.001)
            nn.init.constant_(m.bias, 0)


def partner_lists(patient_ids, batch_size: int):
    """For each of the first ``batch_size`` rows (the anchors) the indices j != i of the rows with the same id, in
    ascending order - the order in which the reference concatenates them (:470)."""
    ids = np.asarray(patient_ids.cpu() if isinstance(patient_ids, torch.Tensor) else patient_ids).reshape(-1)
    _, inv = np.unique(ids, return_inverse=True)
    inv = inv.reshape(-1)
    order = np.argsort(inv, kind="stable")
    starts = np.flatnonzero(np.r_[True, inv[order][1:] != inv[order][:-1]])
    ends = np.r_[starts[1:], len(order)]
    members = {int(inv[order[s]]): order[s:e] for s, e in zip(starts, ends)}
    return [[int(j) for j in members[int(inv[i])] if j != i] for i in range(batch_size)]


class MultiviewFusion(nn.Module):
    """layer_norm_1 / multiview_cross_attention / layer_norm_2 / visual_head of the reference's ``Pretrain``
    (:347-357) with ``forward == Pretrain.multiview_fusion`` (:456-484)."""

    def __init__(self, visual_dim: int = 2048, output_dim: int = 768, heads: int = 8, dropout: float = 0.1):
        super().__init__()
        self.layer_norm_1 = nn.LayerNorm(visual_dim)
        self.layer_norm_2 = nn.LayerNorm(visual_dim)
        self.visual_head = ProjectionHeadPretrain(visual_dim, hidden_dim=output_dim, output_dim=output_dim)
        self.multiview_cross_attention = ScaledDotProductAttention(visual_dim, visual_dim, visual_dim, h=heads, dropout=dropout)

    def forward(self, global_image_embed: torch.Tensor, local_image_embed: torch.Tensor, patient_ids, batch_size: int):
        """global [M, D], local [M, P, D], patient_ids [M] (numpy str/int or tensor) -> ([B, D_out], [B, P, D_out])."""
        att = self.multiview_cross_attention
        x = torch.cat([global_image_embed.unsqueeze(1), local_image_embed], dim=1)        # [M, T, D]
        x = self.layer_norm_1(x)
        t, d = x.shape[1], x.shape[2]
        partners = partner_lists(patient_ids, batch_size)
        with_views = [i for i, p in enumerate(partners) if p]
        fused = x[:batch_size]
        if with_views:
            dev = x.device
            h, dk, dv = att.h, att.d_k, att.d_v
            need = sorted({j for i in with_views for j in partners[i]})                   # images whose K / V are used
            pos = {j: n for n, j in enumerate(need)}
            max_p = max(len(partners[i]) for i in with_views)
            g = len(with_views)
            slot = torch.full((g, max_p), 0, dtype=torch.long)
            live = torch.zeros((g, max_p), dtype=torch.bool)
            for r, i in enumerate(with_views):
                for c, j in enumerate(partners[i]):
                    slot[r, c] = pos[j]
                    live[r, c] = True
            slot, live = slot.to(dev), live.to(dev)
            anchors = torch.tensor(with_views, dtype=torch.long, device=dev)
            xd = x.detach()[torch.tensor(need, dtype=torch.long, device=dev)]             # keys / values are detached (:473)
            q = att.fc_q(x[anchors]).view(g, t, h, dk).permute(0, 2, 1, 3)                 # [G, h, T, dk]
            k_all = att.fc_k(xd).view(len(need), t, h, dk)                                # once per distinct image
            v_all = att.fc_v(xd).view(len(need), t, h, dv)
            k = k_all[slot].view(g, max_p * t, h, dk).permute(0, 2, 3, 1)                  # [G, h, dk, nk]
            v = v_all[slot].view(g, max_p * t, h, dv).permute(0, 2, 1, 3)                  # [G, h, nk, dv]
            score = torch.matmul(q, k) / math.sqrt(dk)                                    # [G, h, T, nk]
            dead = ~live.repeat_interleave(t, dim=1)                                      # padded partner slots
            score = score.masked_fill(dead[:, None, None, :], float("-inf"))
            prob = att.dropout(torch.softmax(score, dim=-1))
            out = torch.matmul(prob, v).permute(0, 2, 1, 3).reshape(g, t, h * dv)
            out = att.fc_o(out)                                                           # [G, T, D]
            upd = self.layer_norm_2(out + x[anchors])
            fused = fused.index_copy(0, anchors, upd)
        new = self.visual_head(fused)
        return new[:, 0, :], new[:, 1:, :]


def patch_multiview_fusion(model: nn.Module) -> nn.Module:
    """Rebind ``multiview_fusion`` on a reference ``Pretrain`` instance (its own sub-modules keep their parameters;
    only the per-anchor Python loop is replaced).  The visual head is used as it is - swap it for
    ``ProjectionHeadPretrain`` (``convert_head``) to get the permute-free layout."""
    import types

    def multiview_fusion(self, global_image_embed, local_image_embed, patient_ids, batch_size):
        return MultiviewFusion.forward(self, global_image_embed, local_image_embed, patient_ids, batch_size)

    model.multiview_fusion = types.MethodType(multiview_fusion, model)
    return model


def convert_head(head: nn.Module) -> ProjectionHeadPretrain:
    """A reference projection head (Sequential ``head`` of Conv1d/BN/ReLU/Conv1d) -> the permute-free form, sharing
    the parameter tensors."""
    c1, bn, _, c2 = head.head
    new = ProjectionHeadPretrain(c1.in_channels, c1.out_channels, c2.out_channels)
    new.head[0], new.head[1], new.head[3] = c1, bn, c2
    return new
