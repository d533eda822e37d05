"""Which part of the K3 epilogue slows the tensor pipe?  (GPU box; results are NOT valid losses)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from evoke_b200 import functional as Fn, _lib, synth
from evoke_b200.ids import DeviceIds
N, D = 16384, 768
ids = synth.make_study_ids(N, seed=1)
x = torch.tensor(synth.make_embeddings(ids, D, seed=2), device="cuda")
y = torch.tensor(synth.make_embeddings(ids, D, seed=3), device="cuda")
q = Fn.l2norm_fwd(x, want_f32=False, want_hi=True, want_lo=False)
k = Fn.l2norm_fwd(y, want_f32=False, want_hi=True, want_lo=False)
dev = DeviceIds(torch.from_numpy(ids).cuda())
bits, counts = Fn.posmask_build(dev, dev, clear_diag=False)
n_ct, n_rt = N // 256, N // 128
rs = torch.empty((n_ct * 2, N), device="cuda"); rp = torch.empty_like(rs); cs = torch.empty((n_rt, N), device="cuda")
def run(flags, label):
    def go():
        _lib.call("evk_mpce_fwd", q.hi.data_ptr(), None, q.ld, k.hi.data_ptr(), None, k.ld, N, N, D, bits.data_ptr(), bits.stride(0),
                  2.0, flags, 0, rs.data_ptr(), rp.data_ptr(), N, cs.data_ptr(), N, torch.cuda.current_stream().cuda_stream)
    for _ in range(3): go()
    torch.cuda.synchronize(); ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{label:46s} {min(ts)*1e3:7.1f} us (median {sorted(ts)[4]*1e3:7.1f})", flush=True)
run(0, "full epilogue")
run(2, "no column sums (no butterfly)")
run(0x400, "no exp, with butterfly")
run(2 | 0x400, "no exp, no butterfly (ld + row adds)")
run(2 | 0x400 | 0x800, "no TMEM loads at all")
run(0x800, "no TMEM loads, exp + butterfly on zeros")
