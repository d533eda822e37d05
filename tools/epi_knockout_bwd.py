"""Which part of the K4a epilogue slows the tensor pipe?  (GPU box; outputs are NOT valid)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import torch
from evoke_b200 import functional as Fn, _lib, synth
from evoke_b200.ids import DeviceIds
N, D = 16384, 768
ids = synth.make_study_ids(N, seed=1)
x = torch.tensor(synth.make_embeddings(ids, D, seed=2), device="cuda"); y = torch.tensor(synth.make_embeddings(ids, D, seed=3), device="cuda")
q = Fn.l2norm_fwd(x, want_f32=False, want_hi=True, want_lo=False); k = Fn.l2norm_fwd(y, want_f32=False, want_hi=True, want_lo=False)
dev = DeviceIds(torch.from_numpy(ids).cuda()); bits, counts = Fn.posmask_build(dev, dev, clear_diag=False)
a = torch.rand(N, device="cuda") + 0.5; b = torch.rand(N, device="cuda") + 0.5
w = torch.empty((N, N), dtype=torch.bfloat16, device="cuda")
def run(flags, label):
    def go():
        _lib.call("evk_mpce_bwd_w", q.hi.data_ptr(), None, q.ld, k.hi.data_ptr(), None, k.ld, N, N, D, bits.data_ptr(), bits.stride(0),
                  counts.data_ptr(), a.data_ptr(), b.data_ptr(), 2.0, flags, 0, w.data_ptr(), None, N, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): go()
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{label:52s} {min(ts)*1e3:7.1f} us (median {sorted(ts)[4]*1e3:7.1f})", flush=True)
run(0, "full epilogue")
run(0x2000, "no TMA store (staging writes kept)")
run(0x1000, "no staging writes, no TMA store")
run(0x400, "no exp/scale math")
run(0x400 | 0x1000, "no math, no staging, no store (ld + pack only)")
