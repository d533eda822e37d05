#!/usr/bin/env python
"""BASELINE.json config 5: a full EVOKE pre-training step on synthetic 224x224 multi-view CXRs with the new contrastive
path, data-parallel over the GPUs of one node.  A BENCHMARK HARNESS: the encoders, heads and optimiser are the
reference's callers of the hot path (out of scope as products) rebuilt from stock parts so
that the step of reference modules/trainer_v0401.py:256-263 can be timed on a box that has no reference checkout:

    torchvision resnet101 (random init, modules/visual_extractor.py:9-24)  -> 49 patch tokens x 2048 + pooled 2048
    HF BertModel, 6 layers x 768 (config/finetune_config.yaml:21-22), random init, L = 100 tokens
    multiview_fusion + projection heads (evoke_b200.fusion = v0520.py:456-484, utils_v0511.py:131-168)
    all_loss = instance + sen_text + mul_pos (v0520.py:528-572), clip_grad_value_(0.1), Adam (trainer_v0401.py:260-263)

Two loss back ends over the SAME model and data:
    evoke_b200   the three losses from this library; with N > 1 ranks the instance loss and the multi-view loss see
                 the GLOBAL batch (global_alignment_sharded / multi_pos_contra_images_sharded)
    reference    the reference's op sequences in PyTorch CUDA eager (oracle/evoke_oracle.py ports), per-rank batch -
                 what `nn.DataParallel` / DDP around the reference module would compute

    python tools/pretrain_step.py [--steps 10] [--studies 32] [--losses evoke_b200|reference] [--amp]
    torchrun --nproc-per-node 8 tools/pretrain_step.py ...
Prints one JSON line (rank 0).  bench.py --config cfg5 calls run() and adds the bench contract's keys.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build_model(output_dim: int, text_layers: int, vocab: int, max_len: int):
    import torch
    import torchvision
    from torch import nn
    from transformers import BertConfig, BertModel

    from evoke_b200 import fusion

    class PretrainLike(nn.Module):
        def __init__(self):
            super().__init__()
            trunk = torchvision.models.resnet101()
            self.visual = nn.Sequential(*list(trunk.children())[:-2])
            self.avg = nn.AvgPool2d(kernel_size=7, stride=1, padding=0)
            self.text_encoder = BertModel(BertConfig(vocab_size=vocab, hidden_size=768, num_hidden_layers=text_layers,
                                                     num_attention_heads=12, intermediate_size=3072,
                                                     max_position_embeddings=max_len + 8), add_pooling_layer=False)
            self.fusion = fusion.MultiviewFusion(2048, output_dim, heads=8)
            self.text_head = fusion.ProjectionHeadPretrain(768, output_dim, output_dim)

        def embed(self, images, token_ids, token_mask, patient_ids, batch_size):
            feat = self.visual(images)                                              # [M, 2048, 7, 7]
            fc = self.avg(feat).reshape(feat.shape[0], feat.shape[1])               # pooled feature (raw, :531)
            att = feat.flatten(2).permute(0, 2, 1)                                  # [M, 49, 2048]
            v_fc, v_att = self.fusion(fc.float(), att.float(), patient_ids, batch_size)
            text = self.text_encoder(input_ids=token_ids, attention_mask=token_mask).last_hidden_state
            text = self.text_head(text.float())
            return fc.float(), v_fc, v_att, text[:, 0, :], text[:, 1:, :]

    return PretrainLike()


def make_batch(studies: int, rank: int, step: int, max_len: int, vocab: int, device):
    """One rank's batch as the reference's collate_fn builds it (dataloaders_v0401.py:60-116): the anchor view of each
    study first, then the auxiliary views; patient_ids in the same order; one report per study."""
    import numpy as np
    import torch
    from evoke_b200 import synth
    rng = np.random.Generator(np.random.PCG64(1000 * step + rank))
    # the SAME multiset of study sizes on every rank and step (shuffled): equal shard sizes for the sharded losses,
    # {1: 25 %, 2: 45 %, 3: 20 %, 4: 10 %} as in SURVEY.md §8d
    base = np.repeat([1, 2, 3, 4], [round(studies * 0.25), round(studies * 0.45), round(studies * 0.20), 0])
    base = np.concatenate([base, np.full(studies - len(base), 4)])[:studies]
    sizes = rng.permutation(base)
    study = np.arange(studies, dtype=np.int64) + rank * studies + step * 100003          # globally unique study ids
    ids = np.concatenate([study, np.repeat(study, sizes - 1)])
    g = torch.Generator().manual_seed(7 * step + rank)
    images = torch.randn((len(ids), 3, 224, 224), generator=g)
    tokens = torch.randint(5, vocab, (studies, max_len), generator=g)
    lens = torch.randint(max_len // 3, max_len + 1, (studies,), generator=g)
    mask = (torch.arange(max_len)[None, :] < lens[:, None]).long()
    _ = synth
    return images.to(device), tokens.to(device), mask.to(device), ids


def run(steps: int = 10, warmup: int = 3, studies: int = 32, losses: str = "evoke_b200", amp: bool = True,
        output_dim: int = 768, text_layers: int = 6, max_len: int = 100, vocab: int = 4000, tau: float = 0.5):
    import numpy as np
    import torch
    import torch.distributed as dist

    import evoke_b200
    from evoke_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                                   # identical initial weights on every rank
    model = build_model(output_dim, text_layers, vocab, max_len).to(dev)
    model.fusion.multiview_cross_attention.dropout.p = 0.0     # deterministic step (the reference uses 0.1)
    opt = torch.optim.Adam(model.parameters(), lr=5e-5, amsgrad=True)
    if losses == "reference":
        from oracle import evoke_oracle as orc          # the reference's op sequences, CUDA eager (baseline leg)

    class Embed(torch.nn.Module):                             # DDP hooks the forward of the wrapped module
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, *a):
            return self.m.embed(*a)

    fwd = torch.nn.parallel.DistributedDataParallel(Embed(model), device_ids=[local_rank], find_unused_parameters=True) \
        if world > 1 else Embed(model)
    t_loss = []

    def step_fn(step):
        images, tokens, mask, ids = make_batch(studies, rank, step, max_len, vocab, dev)
        b = studies
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
            fc, v_fc, v_att, t_fc, t_att = fwd(images, tokens, mask, ids, b)
        fc, v_fc, v_att, t_fc, t_att = (x.float() for x in (fc, v_fc, v_att, t_fc, t_att))     # fp32 hand-over, as the reference
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if losses == "evoke_b200":
            if world > 1:
                from evoke_b200.distributed import global_alignment_sharded, multi_pos_contra_images_sharded
                mul_pos = multi_pos_contra_images_sharded(fc, ids, tau, precision="bf16")     # all views of all ranks
                instance = global_alignment_sharded(v_fc, t_fc, ids[:b], tau, precision="bf16")  # global negatives
            else:
                mul_pos = evoke_b200.multi_pos_contra_images(fc, ids, tau, precision="fp32")
                instance = evoke_b200.global_alignment(v_fc, t_fc, ids, tau, precision="fp32")
            sen_text = evoke_b200.local_text_token_alignment(v_att, t_att, tau)
        else:
            mul_pos = orc.multi_pos_contra_images_port(fc, ids, tau)
            instance = orc.global_alignment_loss_port(v_fc, t_fc, ids, tau)
            sen_text = orc.local_text_token_alignment_port(v_att, t_att, tau)
        all_loss = instance + sen_text + mul_pos.reshape(())
        e1.record()
        opt.zero_grad(set_to_none=True)
        all_loss.backward()
        torch.nn.utils.clip_grad_value_(model.parameters(), 0.1)
        opt.step()
        t_loss.append((e0, e1))
        return all_loss

    for s in range(warmup):
        step_fn(s)
    t_loss.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in range(steps):
        loss = step_fn(warmup + s)
    z.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(z)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_step = ms / steps
    return {"ms_per_step": ms_step, "value": world * studies / (ms_step * 1e-3), "unit": "pairs/s", "n_gpus": world,
            "studies_per_gpu": studies, "losses": losses, "amp_bf16_encoders": bool(amp), "loss": float(loss.item()),
            "ms_loss_forward": float(np.mean([x.elapsed_time(y) for x, y in t_loss])),
            "gpu_launches": _lib.launch_count - launches0, "steps": steps, "warmup": warmup,
            "model": f"resnet101 + BertModel({text_layers}x768) + multiview_fusion + heads({output_dim}), random init, "
                     f"{studies} studies x 1-4 views per GPU, 224x224, L={max_len}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--studies", type=int, default=32)
    ap.add_argument("--losses", default="evoke_b200", choices=["evoke_b200", "reference"])
    ap.add_argument("--no-amp", action="store_true")
    args = ap.parse_args()
    out = run(args.steps, args.warmup, args.studies, args.losses, not args.no_amp)
    if int(os.environ.get("RANK", "0")) == 0:
        print(json.dumps(out), flush=True)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
