"""Timing of local_text_token_alignment at the reference's shape, next to the same op sequence in
PyTorch CUDA eager (the reference's own code path on a GPU).  Run on the GPU box."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import evoke_b200

b, l, p, d, tau = 32, 99, 49, 768, 0.5
torch.manual_seed(0)
v = torch.randn(b, p, d, device="cuda", requires_grad=True)
t = torch.randn(b, l, d, device="cuda", requires_grad=True)


def ref():
    att = F.softmax(t @ v.permute(0, 2, 1) / math.sqrt(d), dim=-1)
    o = F.normalize(torch.bmm(att, v), dim=-1)
    th = F.normalize(t, dim=-1)
    sim = torch.bmm(th, o.permute(0, 2, 1)) / tau
    tgt = torch.arange(l, device="cuda").repeat(b)
    return 0.5 * (F.cross_entropy(sim.reshape(b * l, l), tgt) + F.cross_entropy(sim.permute(0, 2, 1).reshape(b * l, l), tgt))


def ours():
    return evoke_b200.local_text_token_alignment(v, t, tau)


def timed(fn, reps=30):
    for _ in range(5):
        v.grad = t.grad = None
        fn().backward()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        v.grad = t.grad = None
        loss = fn()
        loss.backward()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3, loss.item()


print("torch eager fwd+bwd  us/step, loss:", timed(ref))
print("evoke_b200 fwd+bwd   us/step, loss:", timed(ours))

g = evoke_b200.GraphedLocalTokenAlign(b, p, l, d, tau).capture()
g.load(v.detach(), t.detach())
for _ in range(5):
    g.step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(100):
    loss = g.step()
e1.record()
torch.cuda.synchronize()
print("evoke_b200 CUDA graph us/step, loss:", (e0.elapsed_time(e1) / 100 * 1e3, loss.item()))
