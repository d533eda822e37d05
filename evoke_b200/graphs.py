"""CUDA-graph form of the G loss step: forward + backward of `global_alignment` captured once
for a fixed (N, D) and replayed, which removes the ~20 launch gaps of the eager sequence.

The library's entry points never allocate or synchronise and only enqueue on the current
stream, so the whole autograd forward+backward is capturable.  Inputs live in static buffers
(`image`, `text`, `ids`); `load()` copies a new batch into them (device or pinned-host sources),
`step()` replays the graph and returns the static loss tensor; gradients are in
`image.grad` / `text.grad` (static as well).  Device-resident int32 ids only: host string ids
need the factorisation of evoke_b200.ids first.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import loss as _loss
from .ids import DeviceIds


class GraphedGlobalAlignment:
    def __init__(self, n: int, d: int, temp: float, *, device=None, dtype=torch.float32,
                 precision: str = "bf16", path: str = "auto", two_keys: bool = False, warmup: int = 3,
                 sharded: bool = False, group=None, shard_mode: str = "auto"):
        """n = rows held by THIS process (the whole batch, or this rank's shard when sharded=True:
        the step then is evoke_b200.distributed.global_alignment_sharded, NCCL collectives captured
        with it; every rank must capture and replay in lockstep)."""
        self.sharded, self.group, self.shard_mode = sharded, group, shard_mode
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n, self.d, self.temp, self.precision, self.path = n, d, float(temp), precision, path
        self.image = torch.zeros((n, d), device=device, dtype=dtype, requires_grad=True)
        self.text = torch.zeros((n, d), device=device, dtype=dtype, requires_grad=True)
        self.key = torch.arange(n, device=device, dtype=torch.int32)
        self.key2 = torch.zeros(n, device=device, dtype=torch.int32) if two_keys else None
        self._ids = DeviceIds(self.key, self.key2)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self._warmup = warmup
        with torch.no_grad():                      # non-degenerate contents for the warm-up passes
            self.image.normal_()
            self.text.normal_()

    def _eager(self):
        if self.sharded:
            from .distributed import global_alignment_sharded
            out = global_alignment_sharded(self.image, self.text, self._ids, self.temp, group=self.group,
                                           precision=self.precision, mode=self.shard_mode)
        else:
            out = _loss.global_alignment(self.image, self.text, self._ids, self.temp, precision=self.precision,
                                         path=self.path)
        out.backward()
        return out

    def capture(self):
        side = torch.cuda.Stream(device=self.image.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):          # also binds the context / sets kernel attributes on the
                self.image.grad = None             # autograd thread before capture starts
                self.text.grad = None
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.image.grad = None
        self.text.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        return self

    def load(self, image: torch.Tensor, text: torch.Tensor, ids=None, ids2=None, non_blocking: bool = True):
        """Copy a batch into the static buffers (enqueued on the current stream)."""
        with torch.no_grad():
            self.image.copy_(image, non_blocking=non_blocking)
            self.text.copy_(text, non_blocking=non_blocking)
            if ids is not None:
                self.key.copy_(ids, non_blocking=non_blocking)
            if ids2 is not None and self.key2 is not None:
                self.key2.copy_(ids2, non_blocking=non_blocking)

    def step(self) -> torch.Tensor:
        if self.graph is None:
            self.capture()
        if self.sharded:
            # a replayed graph never re-enters the Python forward: poll the transport's failure mirror here
            from . import peer
            for c in peer._CONTEXTS.values():
                if isinstance(c, peer.PeerContext):
                    c.raise_if_failed()
        self.graph.replay()
        return self.loss


class GraphedLocalTokenAlign:
    """CUDA-graph form of ``local_text_token_alignment`` (reference :506-526) for a fixed (B, P, L, D): at the
    reference's sizes (B=32, L=99, P=49, D=768) the step is launch-bound - about twenty small kernels - so replaying
    the captured forward + backward removes the host cost.  ``load()`` copies a batch into the static buffers,
    ``step()`` replays; gradients are in ``image.grad`` / ``text.grad``."""

    def __init__(self, b: int, p: int, l: int, d: int, temp: float, *, device=None, warmup: int = 3):
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.temp = float(temp)
        self.image = torch.zeros((b, p, d), device=device, requires_grad=True)
        self.text = torch.zeros((b, l, d), device=device, requires_grad=True)
        with torch.no_grad():
            self.image.normal_()
            self.text.normal_()
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self._warmup = warmup

    def _eager(self):
        out = _loss.local_text_token_alignment(self.image, self.text, self.temp)
        out.backward()
        return out

    def capture(self):
        side = torch.cuda.Stream(device=self.image.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(self._warmup):
                self.image.grad = None
                self.text.grad = None
                self._eager()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.image.grad = None
        self.text.grad = None
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._eager()
        return self

    def load(self, image: torch.Tensor, text: torch.Tensor, non_blocking: bool = True):
        with torch.no_grad():
            self.image.copy_(image, non_blocking=non_blocking)
            self.text.copy_(text, non_blocking=non_blocking)

    def step(self) -> torch.Tensor:
        if self.graph is None:
            self.capture()
        self.graph.replay()
        return self.loss
