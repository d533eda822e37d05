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
                                           precision=self.precision, mode=self.shard_mode, graph=False)
        else:
            out = _loss.global_alignment(self.image, self.text, self._ids, self.temp, precision=self.precision,
                                         path=self.path, graph=False)     # this object IS the graph
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


# ------------------------------------------------------------------------------------- drop-in graph cache
import os as _os
from collections import OrderedDict as _OrderedDict

DROPIN_GRAPHS = _os.environ.get("EVOKE_B200_GRAPHS", "1") == "1"     # default of the reference-signature calls
_MAX_CACHED = int(_os.environ.get("EVOKE_B200_GRAPH_CACHE", "6"))
# the single-device loss with K1 / K1b OUTSIDE the graphs (no input copy, fresh gradient tensors).  Measured slower than
# copying the inputs into static buffers (1.39 vs 1.31 ms at N = 16384; 0.25 vs 0.20 ms at N = 4096): outside the
# graphs those four kernels no longer run beside K2 / the second contraction, and every eager-kernel <-> graph-launch
# transition costs launch latency.  Off by default.
ZERO_COPY = _os.environ.get("EVOKE_B200_GRAPH_ZERO_COPY", "0") == "1"
_CACHE: "_OrderedDict[tuple, GraphedStep]" = _OrderedDict()
import threading as _threading
_CACHE_LOCK = _threading.Lock()          # nn.DataParallel calls the loss from one host thread per GPU


class GraphedStep:
    """Forward and backward of one loss call, captured as TWO CUDA graphs, so that the reference-signature call
    (``model.global_alignment_loss(image, text, ids)`` ... ``all_loss.backward()``) costs two graph launches instead
    of ~20 kernel launches: the forward graph is replayed inside the autograd Function's forward, the backward
    graph - which reads the upstream gradient from a static device scalar - inside its backward.  A backward must
    follow its own forward (a stale or repeated backward raises).  An entry keeps its buffers - in bf16 mode the
    N x N bf16 strip, 0.5 GB at N = 16384 - until it leaves the cache (EVOKE_B200_GRAPH_CACHE entries, least
    recently used first; ``clear_graph_cache()`` frees everything).

    Two forms:
      * ``norm`` / ``finish`` given (the single-device loss): the kernels that touch the CALLER's tensors stay outside
        the graphs - K1 (``norm``) reads the embeddings as they are (any strides) and writes the static normalised
        operands, the graphs hold everything in between, K1b (``finish``) turns the graph's dQhat / dKhat into fresh
        gradient tensors.  No copy of the inputs, no static outputs.
            norm(image, text, out | None) -> operands;  fwd(operands, ids, need) -> (loss, state);
            bwd(state, g) -> (dq, dk);  finish(state, image, text, dq, dk, g) -> (d_image, d_text)
      * otherwise (the sharded peer path): the inputs are copied into static buffers and the returned gradients are
        static buffers, valid until the next call with the same signature (torch.cuda.make_graphed_callables semantics).
            fwd(image, text, ids, need) -> (loss, state);  bwd(state, g) -> (d_image, d_text)"""

    def __init__(self, fwd, bwd, image: torch.Tensor, text: Optional[torch.Tensor], ids: DeviceIds, need, warmup: int = 3,
                 norm=None, finish=None):
        dev = image.device
        self.norm, self.finish = norm, finish
        self.zero_copy = norm is not None
        self.image = self.text = self.pre = None
        if not self.zero_copy:
            self.image = torch.empty(image.shape, dtype=image.dtype, device=dev)
            self.text = None if text is None else torch.empty(text.shape, dtype=text.dtype, device=dev)
        self.key = torch.empty_like(ids.key)
        self.key2 = None if ids.key2 is None else torch.empty_like(ids.key2)
        self.ids = DeviceIds(self.key, self.key2)
        self.g = torch.ones(1, dtype=torch.float32, device=dev)
        self.need = (bool(need[0]), bool(need[1]))
        self.gen = 0               # forward generation; a backward must present the generation it belongs to
        self.bwd_gen = -1
        with torch.no_grad():
            if self.zero_copy:
                self.pre = norm(image, text, None)             # allocates the static operands
            self._load(image, text, ids)
        torch.cuda.synchronize(dev)

        def run_fwd():
            return fwd(self.pre, self.ids, self.need) if self.zero_copy else fwd(self.image, self.text, self.ids, self.need)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                    # first-call work (kernel attributes, context binding) before capture
                if self.zero_copy:
                    norm(image, text, self.pre)
                loss, st = run_fwd()
                if any(self.need):
                    out = bwd(st, self.g)
                    if self.zero_copy:
                        finish(st, image, text, out[0], out[1], self.g)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph_f = torch.cuda.CUDAGraph()
        # thread_local: other host threads (a DataLoader's pin-memory thread ...) may keep calling CUDA during capture
        with torch.no_grad(), torch.cuda.graph(self.graph_f, capture_error_mode="thread_local"):
            self.loss, self.state = run_fwd()
        self.graph_b = None
        self.out_a = self.out_b = None                 # (d_image, d_text) static, or (dq, dk) in the zero-copy form
        if any(self.need):
            self.graph_b = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.graph_b, pool=self.graph_f.pool(), capture_error_mode="thread_local"):
                self.out_a, self.out_b = bwd(self.state, self.g)
            # the capture itself ran neither graph: the E strip of `state` is produced by the first replay

    def _load(self, image, text, ids: DeviceIds):
        with torch.no_grad():
            if self.zero_copy:
                self.norm(image, text, self.pre)                       # K1 straight from the caller's tensors
            else:
                self.image.copy_(image, non_blocking=True)             # also gathers strided [:,0,:] views
                if self.text is not None:
                    self.text.copy_(text, non_blocking=True)
            self.key.copy_(ids.key, non_blocking=True)
            if self.key2 is not None:
                self.key2.copy_(ids.key2, non_blocking=True)

    def run_forward(self, image, text, ids: DeviceIds) -> int:
        self._load(image, text, ids)
        self.graph_f.replay()
        self.gen += 1
        return self.gen

    def run_backward(self, gen: int, grad_out: torch.Tensor, image=None, text=None):
        if gen != self.gen:
            raise RuntimeError("evoke_b200: backward of a stale graphed loss call: another forward with the same signature "
                               "ran in between (set EVOKE_B200_GRAPHS=0 or graph=False to keep several calls alive)")
        if self.bwd_gen == gen:
            raise RuntimeError("evoke_b200: backward called twice on a graphed loss call (the E strip was consumed)")
        self.bwd_gen = gen
        with torch.no_grad():
            self.g.copy_(grad_out.reshape(1), non_blocking=True)
            self.graph_b.replay()
            if self.zero_copy:
                return self.finish(self.state, image, text, self.out_a, self.out_b, self.g)
        return self.out_a, self.out_b


class _GraphedLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gs: GraphedStep, ids: DeviceIds, image: torch.Tensor, text: Optional[torch.Tensor]):
        ctx.gs = gs
        ctx.gen = gs.run_forward(image, text, ids)
        if gs.zero_copy:                                # K1b reads the caller's embeddings again
            ctx.save_for_backward(*([image] if text is None else [image, text]))
        ctx.has_text = text is not None
        out = gs.loss.clone().reshape(())
        return out if image.dtype == torch.float32 else out.to(image.dtype)

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        image = text = None
        if ctx.gs.zero_copy:
            saved = ctx.saved_tensors
            image, text = saved[0], (saved[1] if ctx.has_text else None)
        d_image, d_text = ctx.gs.run_backward(ctx.gen, grad_out.to(torch.float32), image, text)
        return None, None, d_image if ctx.needs_input_grad[2] else None, d_text if ctx.needs_input_grad[3] else None


def graphed_call(key: tuple, fwd, bwd, image: torch.Tensor, text: Optional[torch.Tensor], ids: DeviceIds,
                 norm=None, finish=None) -> torch.Tensor:
    """Run one loss call through the graph cache (capturing on first use of this signature)."""
    need = (image.requires_grad and torch.is_grad_enabled(),
            text is not None and text.requires_grad and torch.is_grad_enabled())
    key = key + (tuple(image.shape), image.dtype, image.device.index, ids.key2 is not None, need)
    with _CACHE_LOCK:
        gs = _CACHE.get(key)
        if gs is not None:
            _CACHE.move_to_end(key)
    if gs is None:
        # (captured outside the lock: a capture takes tens of milliseconds and, on the sharded path, is collective)
        gs = GraphedStep(fwd, bwd, image.detach(), None if text is None else text.detach(), ids, need, norm=norm, finish=finish)
        with _CACHE_LOCK:
            _CACHE[key] = gs
            while len(_CACHE) > _MAX_CACHED:
                _CACHE.popitem(last=False)            # least recently used: frees its graphs and their memory pool
    return _GraphedLoss.apply(gs, ids, image, text)


def clear_graph_cache() -> None:
    _CACHE.clear()
