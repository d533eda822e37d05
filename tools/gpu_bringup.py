"""Staged GPU bring-up: every stage runs in its own subprocess under a timeout so a hang or a
fault in one kernel does not take the rest of the call with it.  Writes gpurun_out/bringup.log.

    python tools/gpu_bringup.py            # all stages
    python tools/gpu_bringup.py probe      # one stage (in-process)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage_k12():
    import numpy as np, torch
    from evoke_b200 import functional as Fn, ids as idmod, synth
    from oracle import evoke_oracle as orc
    x = torch.randn(300, 768, device="cuda")
    out = Fn.l2norm_fwd(x, want_f32=True, want_hi=True, want_lo=True)
    ref = torch.nn.functional.normalize(x, dim=-1)
    print("K1 max abs err", (out.f32 - ref).abs().max().item())
    ids = synth.make_study_ids(1000, seed=1)
    dev = idmod.DeviceIds(torch.from_numpy(ids).cuda())
    bits, counts = Fn.posmask_build(dev, dev, clear_diag=True)
    wb, wc = orc.posmask_packed(ids, clear_diag=True)
    got = bits.cpu().numpy().view(np.uint32)
    print("K2 bits equal", np.array_equal(got[:, :wb.shape[1]], wb), "counts equal", np.array_equal(counts.cpu().numpy(), wc))


def stage_small():
    import golden_cases as gc
    from gpu_util import check_against_golden
    for c in gc.CASES:
        try:
            print(c.name, check_against_golden(c, "fp32", "small"))
        except AssertionError as e:
            print("FAIL", c.name, e)


def _probe(a_major, b_major, m, n, k, variant, splits=1):
    import torch
    from evoke_b200 import functional as Fn
    torch.manual_seed(0)
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    b = torch.randn(n, k, device="cuda").to(torch.bfloat16)
    pad = lambda t: torch.nn.functional.pad(t, (0, (-t.shape[1]) % 8)).contiguous()
    a_st = pad(a.t().contiguous()) if a_major else pad(a)
    b_st = pad(b.t().contiguous()) if b_major else pad(b)
    c = Fn.tc_gemm_probe(a_st, b_st, a_major, b_major, m, n, k, variant=variant, splits=splits)
    torch.cuda.synchronize()
    want = a.float() @ b.float().t()
    err = (c - want).abs().max().item() / want.abs().max().item()
    print(f"probe a_major={a_major} b_major={b_major} m={m} n={n} k={k} variant={variant} splits={splits}: rel err {err:.3e}",
          "OK" if err < 1e-5 else "MISMATCH", flush=True)
    if err >= 1e-5 and m <= 128:
        bad = ((c - want).abs() > 1e-3 * want.abs().max()).nonzero()
        print("   first mismatches (row, col):", bad[:8].tolist(), " count", len(bad))


def stage_probe_kk():
    _probe(0, 0, 128, 256, 64, 0)
    _probe(0, 0, 128, 256, 256, 0)
    _probe(0, 0, 300, 520, 200, 0)
    _probe(0, 0, 1024, 1024, 1024, 0, splits=3)


def stage_probe_mn0():
    for am, bm in ((0, 1), (1, 0), (1, 1)):
        _probe(am, bm, 128, 256, 64, 0)
        _probe(am, bm, 300, 520, 200, 0)


def stage_probe_cta1():
    for am, bm in ((0, 0), (0, 1), (1, 0), (1, 1)):
        _probe(am, bm, 300, 520, 200, 4)
    _probe(0, 0, 1024, 1024, 1024, 4, splits=3)


def stage_probe_cta2_small():
    _probe(0, 0, 256, 256, 64, 2)
    _probe(0, 0, 256, 256, 256, 2)


def stage_probe_cta2():
    for am, bm in ((0, 0), (0, 1), (1, 0), (1, 1)):
        _probe(am, bm, 128, 256, 64, 2)
        _probe(am, bm, 300, 520, 200, 2)
    _probe(0, 0, 1024, 1024, 1024, 2, splits=3)
    _probe(1, 1, 2048, 768, 4096, 2)


def stage_tc():
    import golden_cases as gc
    from gpu_util import check_against_golden
    for prec in ("fp32", "bf16"):
        for c in gc.CASES:
            if prec == "bf16" and c.zero_row >= 0:
                continue
            try:
                print(prec, c.name, check_against_golden(c, prec, "tc"), flush=True)
            except AssertionError as e:
                print("FAIL", prec, c.name, e, flush=True)


def stage_smoke():
    import __graft_entry__ as g
    g.smoke()


STAGES = {"k12": stage_k12, "small": stage_small, "probe_kk": stage_probe_kk, "probe_mn0": stage_probe_mn0,
          "probe_cta1": stage_probe_cta1, "probe_cta2_small": stage_probe_cta2_small, "probe_cta2": stage_probe_cta2,
          "tc": stage_tc, "smoke": stage_smoke}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        for s in sys.argv[1:]:
            STAGES[s]()
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "w")
    names = [a for a in os.environ.get("BRINGUP_STAGES", "").split(",") if a] or list(STAGES)
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], cwd=ROOT, timeout=120,
                               stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            out, rc = r.stdout, r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or ""), "TIMEOUT"
        msg = f"===== stage {name}: rc={rc} ({time.time() - t0:.1f}s)\n{out}\n"
        print(msg, flush=True)
        log.write(msg)
        log.flush()
