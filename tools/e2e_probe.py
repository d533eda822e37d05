import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch, numpy as np
import evoke_b200
from evoke_b200 import synth
N, D = 16384, 768
ids = synth.make_study_ids(N, seed=1)
h_img = torch.from_numpy(synth.make_embeddings(ids, D, seed=2)).pin_memory()
h_txt = torch.from_numpy(synth.make_embeddings(ids, D, seed=3)).pin_memory()
h_ids = torch.from_numpy(ids).pin_memory()
dev = torch.device("cuda")
gs = [evoke_b200.GraphedGlobalAlignment(N, D, 0.5, precision="bf16", path="tc").capture() for _ in range(2)]
cs = torch.cuda.Stream()
def h2d_only(k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.cuda.stream(cs):
        for i in range(k): gs[i & 1].load(h_img, h_txt, h_ids)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
def compute_only(k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k): gs[i & 1].step()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
def both_free(k):
    """copies and compute issued back-to-back with no dependencies at all (upper bound on overlap)"""
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(k):
        with torch.cuda.stream(cs): gs[(i + 1) & 1].load(h_img, h_txt, h_ids)
        gs[i & 1].step()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
def pipelined(k, read_loss):
    ready = [torch.cuda.Event(), torch.cuda.Event()]; consumed = [torch.cuda.Event(), torch.cuda.Event()]
    for e in consumed: e.record()
    def pre(s):
        with torch.cuda.stream(cs):
            cs.wait_event(consumed[s]); gs[s].load(h_img, h_txt, h_ids); ready[s].record(cs)
    torch.cuda.synchronize(); t0 = time.perf_counter(); pre(0)
    for i in range(k):
        s = i & 1
        if i + 1 < k: pre(s ^ 1)
        torch.cuda.current_stream().wait_event(ready[s]); l = gs[s].step(); consumed[s].record()
        if read_loss: l.item()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k * 1e3
for f, a in ((h2d_only, ()), (compute_only, ()), (both_free, ()), (pipelined, (False,)), (pipelined, (True,))):
    f(5, *a); print(f.__name__, a, round(f(40, *a), 3), "ms/step", flush=True)
