"""Drop-in replacements for the reference's contrastive loss methods.

The reference implements the objective as methods of the ``Pretrain`` nn.Module
(models/model_pretrain_finetune_v0520.py; byte-identical bodies in
model_pretrain_finetune_v0425_ablation.py:274-294/:324-342, _v0623_large_res.py:262-282/:311-329,
_v0520_abn.py, _v0719_twoview.py, _v0425_ori.py):

    Pretrain.global_alignment_loss(self, global_image_embed, global_text_embed, patient_ids)   :486-504
    Pretrain.multi_pos_contra_images_v0401(self, global_image_embed, patient_ids)              :421-446

The two functions below keep those names, argument order and meaning, read the temperatures
from ``self.args['instance_temp']`` / ``self.args['region_temp']`` exactly as the reference does,
return a differentiable tensor on the input device (0-dim; shape [1] leaf of zeros when MPC
finds no multi-view study, :427-428), and raise ordinary Python exceptions on bad input.
``patch_pretrain`` rebinds them on a reference ``Pretrain`` class or instance.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import functional as Fn
from . import ids as idmod

DEFAULT_PRECISION = "fp32"      # the reference computes in fp32; "bf16" is the throughput mode


MAX_INV_TAU = 40.0        # EVK_MAX_INV_TAU of include/evoke_b200.h


def _inv_tau(temp: float) -> float:
    """1/temperature, inside the domain of the fixed-shift softmax: the kernels use E = exp(S - 1/tau) (|S| <= 1/tau
    for unit rows), and exp(-2/tau) must stay a normal fp32 number or a row of weakly aligned pairs could sum to
    zero.  The reference (F.cross_entropy's running max) accepts any temperature; this library raises below
    tau = 0.025 instead of returning inf/NaN.  EVOKE trains with 0.5 (config/finetune_config.yaml:73-74)."""
    temp = float(temp)
    if not temp > 0.0:
        raise ValueError(f"temperature must be positive, got {temp}")
    if 1.0 / temp > MAX_INV_TAU:
        raise ValueError(f"evoke_b200: temperature {temp} is below the supported minimum {1.0 / MAX_INV_TAU} "
                         "(fixed-shift softmax: exp(-2/tau) must stay normal in fp32)")
    return 1.0 / temp


def global_alignment(image: torch.Tensor, text: torch.Tensor, patient_ids, temp: float, *,
                     precision: str = DEFAULT_PRECISION, path: str = "auto", graph: Optional[bool] = None) -> torch.Tensor:
    """Multi-positive image<->text InfoNCE in both directions (reference :486-504).

    image, text: [B, D] projected embeddings (any strides; fp32/bf16/fp16) on a CUDA device.
    patient_ids: what the reference passes (numpy array of str/int, length >= B; only the first B
    are used, :488), a (patient, study) pair, or an integer tensor already on the device.
    graph: replay the call from the CUDA-graph cache (evoke_b200.graphs.GraphedStep: one graph for the forward,
    one for the backward, captured on first use of a (shape, dtype, precision, temperature) signature; the inputs
    are copied into static buffers and the returned gradients are static buffers, as with
    torch.cuda.make_graphed_callables.  EVOKE_B200_GRAPH_ZERO_COPY=1 keeps K1 / K1b outside the graphs instead - no
    copy, fresh gradient tensors, but 6 % slower at N = 16384 because those kernels no longer overlap their
    neighbours).  None: EVOKE_B200_GRAPHS (default on), never inside an outer capture.
    """
    Fn._require_cuda(image, "global_image_embed")
    Fn._require_cuda(text, "global_text_embed")
    if image.shape != text.shape:
        raise ValueError(f"image/text embedding shapes differ: {tuple(image.shape)} vs {tuple(text.shape)}")
    if image.dtype != text.dtype:
        raise TypeError(f"image/text embedding dtypes differ: {image.dtype} vs {text.dtype}")
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    b, d = int(image.shape[0]), int(image.shape[1])
    if b == 0:
        raise ValueError("empty batch")
    dev_ids, _ = idmod.to_device_ids(patient_ids, image.device, n=b)
    if len(dev_ids) < b:
        raise ValueError(f"patient_ids has {len(dev_ids)} entries for a batch of {b}")
    inv_tau = _inv_tau(temp)
    path = Fn.choose_path(path, b, b, d)
    with torch.cuda.device(image.device):
        from . import graphs
        use_graph = graphs.DROPIN_GRAPHS if graph is None else bool(graph)
        if use_graph and not torch.cuda.is_current_stream_capturing():
            def cfg_of(ids):
                return Fn.LossConfig(kind="G", inv_tau=inv_tau, precision=precision, path=path, row_ids=ids)

            def norm(im, tx, out):                     # K1: outside the graphs, straight from the caller's tensors
                return Fn.normalize_pair(cfg_of(dev_ids), im, tx, out)

            def fwd(pre, ids, need):
                return Fn.mpce_forward(cfg_of(ids), None, None, need, pre=pre)

            def bwd(st, g):
                return Fn.mpce_backward(st, g, finish=False)

            key = ("G", inv_tau, precision, path, Fn.E_STRIP, Fn.OVERLAP_STREAMS, Fn.MASK_FREE, graphs.ZERO_COPY)
            if graphs.ZERO_COPY:
                return graphs.graphed_call(key, fwd, bwd, image, text, dev_ids, norm=norm, finish=Fn.mpce_finish)

            def fwd_all(im, tx, ids, need):            # everything inside the graphs, inputs copied into static buffers
                return Fn.mpce_forward(cfg_of(ids), im, tx, need)
            return graphs.graphed_call(key, fwd_all, Fn.mpce_backward, image, text, dev_ids)
        cfg = Fn.LossConfig(kind="G", inv_tau=inv_tau, precision=precision, path=path, row_ids=dev_ids)
        return Fn.multi_positive_ce(cfg, image, text)


def multi_pos_contra_images(image: torch.Tensor, patient_ids, temp: float, *,
                            precision: str = DEFAULT_PRECISION, path: str = "auto") -> torch.Tensor:
    """Image<->image multi-positive contrastive loss over all views (reference :421-446):
    self-similarity excluded, views of single-view studies removed from queries and keys, one
    direction, mean over the kept rows.  Returns tensor([0.0]) (shape [1], requires_grad leaf)
    when no study has a second view (:427-428)."""
    Fn._require_cuda(image, "global_image_embed")
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    m, d = int(image.shape[0]), int(image.shape[1])
    dev_ids, codes = idmod.to_device_ids(patient_ids, image.device, n=None)
    if len(dev_ids) != m:
        raise ValueError(f"patient_ids has {len(dev_ids)} entries for {m} image embeddings")
    with torch.cuda.device(image.device):
        if codes is not None:
            keep_host = idmod.multi_view_rows(codes)                     # host ids: no device sync
            keep = torch.from_numpy(keep_host).to(image.device) if len(keep_host) else None
            n_keep = len(keep_host)
        else:
            # device ids: count same-key partners with K2, then one size-determining sync (the
            # reference syncs here too: `len(idx) == 0`, :427)
            _, cnt = Fn.posmask_build(dev_ids, dev_ids, clear_diag=True)
            keep = torch.nonzero(cnt > 0).reshape(-1).to(torch.int32)
            n_keep = int(keep.numel())
        if n_keep == 0:
            return torch.tensor([0.0], requires_grad=True, device=image.device)
        gather = None if n_keep == m else keep.contiguous()
        row_ids = dev_ids if gather is None else dev_ids.index(keep.long())
        cfg = Fn.LossConfig(kind="MPC", inv_tau=_inv_tau(temp), precision=precision,
                            path=Fn.choose_path(path, n_keep, n_keep, d), row_ids=row_ids, gather=gather)
        return Fn.multi_positive_ce(cfg, image, None)


def global_alignment_avgpos(image: torch.Tensor, text: torch.Tensor, patient_ids, temp: float) -> torch.Tensor:
    """PretrainNewMulPos.global_alignment_loss (reference :748-815): the positives of a row are averaged into ONE
    logit and a plain cross entropy against the negatives follows; both directions, x0.5, / B.  Returns shape [1]
    like the reference.  fp32 (small path)."""
    Fn._require_cuda(image, "global_image_embed")
    Fn._require_cuda(text, "global_text_embed")
    if image.shape != text.shape:
        raise ValueError(f"image/text embedding shapes differ: {tuple(image.shape)} vs {tuple(text.shape)}")
    b = int(image.shape[0])
    if b == 0:
        raise ValueError("empty batch")
    dev_ids, _ = idmod.to_device_ids(patient_ids, image.device, n=b)
    if len(dev_ids) < b:
        raise ValueError(f"patient_ids has {len(dev_ids)} entries for a batch of {b}")
    cfg = Fn.LossConfig(kind="AG", inv_tau=_inv_tau(temp), precision="fp32", path="small", row_ids=dev_ids)
    with torch.cuda.device(image.device):
        return Fn.avgpos_ce(cfg, image, text)


def multi_pos_contra_images_avgpos(image: torch.Tensor, patient_ids, temp: float) -> torch.Tensor:
    """PretrainNewMulPos.multi_pos_contra_images_v0404 (reference :670-708): like v0401 but with the averaged
    positive logit, and single-view rows are removed from the queries only - every row stays a key (:685).
    Returns shape [1]; tensor([0.0]) leaf when no study has a second view (:676-677)."""
    Fn._require_cuda(image, "global_image_embed")
    m = int(image.shape[0])
    dev_ids, codes = idmod.to_device_ids(patient_ids, image.device, n=None)
    if len(dev_ids) != m:
        raise ValueError(f"patient_ids has {len(dev_ids)} entries for {m} image embeddings")
    with torch.cuda.device(image.device):
        if codes is not None:
            n_keep = len(idmod.multi_view_rows(codes))                   # host ids: no device sync
        else:
            _, cnt = Fn.posmask_build(dev_ids, dev_ids, clear_diag=True)
            n_keep = int((cnt > 0).sum().item())                         # the reference syncs here too (:676)
        if n_keep == 0:
            return torch.tensor([0.0], requires_grad=True, device=image.device)
        cfg = Fn.LossConfig(kind="AMPC", inv_tau=_inv_tau(temp), precision="fp32", path="small", row_ids=dev_ids,
                            n_keep=n_keep)
        return Fn.avgpos_ce(cfg, image, None)


def local_text_token_alignment(local_image: torch.Tensor, local_text: torch.Tensor, temp: float) -> torch.Tensor:
    """Pretrain.local_text_token_alignment_loss (reference :506-526): local_image [B, P, D] patch tokens, local_text
    [B, L, D] text tokens (any strides / float dtype) -> 0-dim loss.  fp32 kernels; P <= 1024, D <= 4096."""
    for name, x in (("local_image_embed", local_image), ("local_text_embed", local_text)):
        if not x.is_cuda:
            raise RuntimeError(f"evoke_b200: `{name}` lives on {x.device}; this library only runs on a CUDA (sm_100a) device")
        if x.dim() != 3:
            raise ValueError(f"evoke_b200: `{name}` must be 3-D [B, tokens, D], got shape {tuple(x.shape)}")
    if local_image.shape[0] != local_text.shape[0] or local_image.shape[2] != local_text.shape[2]:
        raise ValueError(f"patch / text token shapes do not match: {tuple(local_image.shape)} vs {tuple(local_text.shape)}")
    if local_text.shape[1] > 4096 or local_image.shape[1] > 1024 or local_image.shape[2] > 4096:
        raise ValueError("local_text_token_alignment supports up to 4096 text tokens, 1024 patches and D <= 4096")
    with torch.cuda.device(local_image.device):
        return Fn.local_token_align(_inv_tau(temp), local_image, local_text)


# ---------------------------------------------------------------- methods with the reference's signatures
def global_alignment_loss(self, global_image_embed, global_text_embed, patient_ids):
    """Same signature as Pretrain.global_alignment_loss (reference :486)."""
    return global_alignment(global_image_embed, global_text_embed, patient_ids, self.args["instance_temp"],
                            precision=getattr(self, "_evoke_b200_precision", DEFAULT_PRECISION),
                            graph=getattr(self, "_evoke_b200_graphs", None))


def multi_pos_contra_images_v0401(self, global_image_embed, patient_ids):
    """Same signature as Pretrain.multi_pos_contra_images_v0401 (reference :421)."""
    return multi_pos_contra_images(global_image_embed, patient_ids, self.args["region_temp"],
                                   precision=getattr(self, "_evoke_b200_precision", DEFAULT_PRECISION))


def global_alignment_loss_newmulpos(self, global_image_embed, global_text_embed, patient_ids):
    """Same signature as PretrainNewMulPos.global_alignment_loss (reference :748)."""
    return global_alignment_avgpos(global_image_embed, global_text_embed, patient_ids, self.args["instance_temp"])


def multi_pos_contra_images_v0404(self, global_image_embed, patient_ids):
    """Same signature as PretrainNewMulPos.multi_pos_contra_images_v0404 (reference :670)."""
    return multi_pos_contra_images_avgpos(global_image_embed, patient_ids, self.args["region_temp"])


def patch_pretrain_newmulpos(target):
    """Rebind the two loss methods of a reference ``PretrainNewMulPos`` class or instance (:748, :670)."""
    if isinstance(target, type):
        target.global_alignment_loss = global_alignment_loss_newmulpos
        target.multi_pos_contra_images_v0404 = multi_pos_contra_images_v0404
    else:
        import types
        target.global_alignment_loss = types.MethodType(global_alignment_loss_newmulpos, target)
        target.multi_pos_contra_images_v0404 = types.MethodType(multi_pos_contra_images_v0404, target)
    return target


def local_text_token_alignment_loss(self, local_image_embed, local_text_embed):
    """Same signature as Pretrain.local_text_token_alignment_loss (reference :506)."""
    return local_text_token_alignment(local_image_embed, local_text_embed, self.args["region_temp"])


def patch_pretrain(target, precision: str = DEFAULT_PRECISION, local_tokens: bool = False, graphs: Optional[bool] = None,
                   fusion: bool = False):
    """Rebind the two loss methods (and, with ``local_tokens=True``, ``local_text_token_alignment_loss`` :506) on a
    reference ``Pretrain`` class (affects every instance) or on a single instance.  Works for all six model files because only the method names and
    ``self.args`` are relied upon.  ``graphs``: run ``global_alignment_loss`` from the CUDA-graph cache (None: the
    EVOKE_B200_GRAPHS default, on).  ``fusion=True`` also rebinds ``multiview_fusion`` (:456-484) to the loop-free form
    of evoke_b200.fusion (same sub-modules and parameters: ``layer_norm_1/2``, ``multiview_cross_attention``,
    ``visual_head``).  Returns ``target``."""
    if precision not in ("fp32", "bf16"):
        raise ValueError(f"precision must be 'fp32' or 'bf16', got {precision!r}")
    if isinstance(target, type):
        target.global_alignment_loss = global_alignment_loss
        target.multi_pos_contra_images_v0401 = multi_pos_contra_images_v0401
        if local_tokens:
            target.local_text_token_alignment_loss = local_text_token_alignment_loss
        target._evoke_b200_precision = precision
        target._evoke_b200_graphs = graphs
        if fusion:
            from .fusion import MultiviewFusion
            target.multiview_fusion = MultiviewFusion.forward
    else:
        import types
        target.global_alignment_loss = types.MethodType(global_alignment_loss, target)
        target.multi_pos_contra_images_v0401 = types.MethodType(multi_pos_contra_images_v0401, target)
        if local_tokens:
            target.local_text_token_alignment_loss = types.MethodType(local_text_token_alignment_loss, target)
        target._evoke_b200_precision = precision
        target._evoke_b200_graphs = graphs
        if fusion:
            from .fusion import patch_multiview_fusion
            patch_multiview_fusion(target)
    return target


class ContrastiveObjective(torch.nn.Module):
    """Stand-alone module form: holds ``args`` like the reference's Pretrain does."""

    def __init__(self, instance_temp: float = 0.5, region_temp: float = 0.5, precision: str = DEFAULT_PRECISION):
        super().__init__()
        self.args = {"instance_temp": instance_temp, "region_temp": region_temp}
        self._evoke_b200_precision = precision

    global_alignment_loss = global_alignment_loss
    multi_pos_contra_images_v0401 = multi_pos_contra_images_v0401

    def forward(self, global_image_embed, global_text_embed, patient_ids, view_embed: Optional[torch.Tensor] = None):
        """instance loss (+ multi-view loss over ``view_embed`` when given), as summed at :563."""
        loss = self.global_alignment_loss(global_image_embed, global_text_embed, patient_ids)
        if view_embed is not None:
            loss = loss + self.multi_pos_contra_images_v0401(view_embed, np.asarray(patient_ids) if not
                                                             isinstance(patient_ids, torch.Tensor) else patient_ids)
        return loss
