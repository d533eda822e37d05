"""Can the HBM-bound K4t (E -> W transform) hide next to the tensor-bound gradient contractions?
Times, at the bench shape: the two contractions alone, K4t alone, and both at once on separate
streams (no data dependency: K4t works on a second strip).  Run on the GPU box."""
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from evoke_b200 import functional as Fn, synth
from evoke_b200.ids import to_device_ids

N, D = 16384, 768
dev = torch.device("cuda")
ids = synth.make_study_ids(N, seed=1)
xq = torch.tensor(synth.make_embeddings(ids, D, seed=2), device=dev)
xk = torch.tensor(synth.make_embeddings(ids, D, seed=3), device=dev)
q = Fn.l2norm_fwd(xq, want_f32=False, want_hi=True, want_lo=False)
k = Fn.l2norm_fwd(xk, want_f32=False, want_hi=True, want_lo=False)
rid, _ = to_device_ids(ids, dev, n=N)
bits, counts = Fn.posmask_build(rid, rid, clear_diag=False)
rs, rp, cs, e1, ld = Fn.tc_fwd_store(q, k, bits, 2.0, 0)
e2 = e1.clone()
a_row = 1.0 / Fn.reduce_partials(rs, int(rs.shape[0]), N)
b_col = 1.0 / Fn.reduce_partials(cs, int(cs.shape[0]), N)
dq = torch.zeros(N, D, device=dev)
dk = torch.zeros(N, D, device=dev)
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def gemms(two_streams=False):
    Fn.tc_bwd_gemm(e1, None, ld, N, N, False, k, 0, out=dq)
    if two_streams:
        with torch.cuda.stream(s2):
            Fn.tc_bwd_gemm(e1, None, ld, N, N, True, q, 0, out=dk)
    else:
        Fn.tc_bwd_gemm(e1, None, ld, N, N, True, q, 0, out=dk)


def xform(exact=True):
    Fn.tc_w_from_e(e2, ld, N, bits, counts, a_row, b_col, q if exact else None, k if exact else None, 2.0)


def timed(fn, reps=8):
    best = 1e9
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        main = torch.cuda.current_stream()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for s in (s1, s2, s3):
            s.wait_stream(main)
        fn()
        for s in (s1, s2, s3):
            main.wait_stream(s)
        t1.record()
        torch.cuda.synchronize()
        best = min(best, t0.elapsed_time(t1))
    return best * 1e3


def both(two_streams):
    def f():
        with torch.cuda.stream(s1):
            gemms(two_streams)
        with torch.cuda.stream(s3):
            xform()
    return f


def both_xform_first(two_streams):
    def f():
        with torch.cuda.stream(s3):
            xform()
        with torch.cuda.stream(s1):
            gemms(two_streams)
    return f


print("gemms alone (1 stream)      us", round(timed(lambda: gemms(False)), 1))
print("gemms alone (2 streams)     us", round(timed(lambda: gemms(True)), 1))
print("K4t alone (exact positives) us", round(timed(lambda: xform(True)), 1))
print("K4t alone (strip only)      us", round(timed(lambda: xform(False)), 1))
print("gemms(1 stream) || K4t      us", round(timed(both(False)), 1))
print("gemms(2 streams) || K4t     us", round(timed(both(True)), 1))
print("K4t first, gemms(1 stream)  us", round(timed(both_xform_first(False)), 1))
