"""Three eager steps of local_text_token_alignment at the reference's shape (for an ncu launch list)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import evoke_b200

b, l, p, d, tau = 32, 99, 49, 768, 0.5
torch.manual_seed(0)
v = torch.randn(b, p, d, device="cuda", requires_grad=True)
t = torch.randn(b, l, d, device="cuda", requires_grad=True)
for _ in range(3):
    v.grad = t.grad = None
    evoke_b200.local_text_token_alignment(v, t, tau).backward()
torch.cuda.synchronize()
