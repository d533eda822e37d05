"""Main-loop rate experiments (GPU box): probe GEMM at various shapes / layouts, output dropped."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from evoke_b200 import functional as Fn, _lib

def run(am, bm, m, n, k, variant, splits, reps=5):
    a = torch.randn(k if am else m, m if am else k, device="cuda").to(torch.bfloat16)
    b = torch.randn(k if bm else n, n if bm else k, device="cuda").to(torch.bfloat16)
    c = torch.zeros((m, n if (variant & 8) == 0 else 256), dtype=torch.float32, device="cuda")
    ldc = c.stride(0) if (variant & 8) == 0 else n   # output dropped: pointer never dereferenced
    def go():
        _lib.call("evk_tc_gemm_probe", a.data_ptr(), a.stride(0), am, b.data_ptr(), b.stride(0), bm, m, n, k,
                  c.data_ptr(), ldc, variant, splits, torch.cuda.current_stream().cuda_stream)
    for _ in range(2): go()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"A{'MN' if am else 'K '} B{'MN' if bm else 'K '} m={m:6d} n={n:6d} k={k:6d} cta2={bool(variant&2)} noepi={bool(variant&8)} "
          f"splits={splits}: {t*1e3:8.1f} us  {2.0*m*n*k/t/1e9:8.1f} TF/s", flush=True)

N, D = 16384, 768
for cta in (2, 4):
    run(0, 0, N, N, D, cta | 8, 1)       # fwd shape, no epilogue
    run(0, 1, N, N, D, cta | 8, 1)       # fwd shape with MN-major B
    run(0, 1, N, D, N, cta | 8, 0)       # dQ GEMM shape, no epilogue
    run(0, 1, N, D, N, cta, 0)           # dQ GEMM shape, with red.add epilogue
    run(1, 1, N, D, N, cta | 8, 0)       # dK GEMM shape
    run(0, 0, N, D, N, cta | 8, 0)       # K-major both, GEMM shape
# cuBLAS reference points
for (m, n, k) in ((N, N, D), (N, D, N), (8192, 8192, 8192)):
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16); b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    for _ in range(2): (a @ b.t())
    torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); c = a @ b.t(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"cuBLAS m={m} n={n} k={k}: {min(ts)*1e3:.1f} us {2.0*m*n*k/min(ts)/1e9:.1f} TF/s", flush=True)
