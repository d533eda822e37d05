#!/usr/bin/env python
"""Benchmark of the contrastive hot path (BASELINE.json metric): multi-positive image<->text
contrastive loss forward+backward, pairs/s at global batch N=16384, D=768, bf16 operands.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg3]   # this repo's CUDA path
    python bench.py --impl reference [--gpus N] [--steps K] ...           # the reference's CPU path
    torchrun --nproc-per-node N bench.py --gpus N ...                     # N>1: one rank per GPU

A step is one fwd+bwd of `global_alignment_loss` over one synthetic batch.  --config selects the workload
(BASELINE.json configs): cfg3 (default, the configuration the metric is quoted on), cfg1 (the
reference's own batch: 32 two-view studies, fp32 small path, G + MPC), cfg2 (N=4096), cfg4 (N=32768, D=512,
(patient, study) keys).  Prints ONE JSON line on rank 0:

  value          device-timed replay of the captured step, inputs resident in HBM
  e2e            through the public API from pinned HOST buffers (H2D of embeddings + ids, D2H of the loss, timed)
  dropin         the reference-signature call (patch_pretrain's method / global_alignment_sharded), as a user's
                 training loop makes it: device inputs -> loss -> .backward(); graph cache on
  roofline       the slowest tcgen05 kernel, CUDA events around its launches (eager pass of the same steps)
  sustained      the same replay repeated for >= 1 s (clocks settle under the power cap)
  fingerprint    loss, gradient norms and 16 fixed gradient rows against tests/golden/fingerprint_<cfg>.json
                 (fp64 oracle): a run whose gradients differ by more than 2e-2 FAILS - at every GPU count
  cpu_baseline   the oracle's PyTorch-CPU port of the reference on the host cores
  cuda_eager_baseline   the same port (the reference's op sequence, fp32, TF32 off) on this B200
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "contrastive loss fwd+bwd pairs/sec at N=16384,D=768"
TAU = 0.5
CONFIGS = {
    "cfg1": dict(n=32, d=768, precision="fp32", path="small", two_keys=False, sizes=None,
                 workload="cfg1: EVOKE-224 reference batch: 32 two-view studies, D=768 synthetic projected embeddings, fp32 - "
                          "multi_pos_contra_images_v0401 over the 64 views + global_alignment_loss over 32 pairs, fwd+bwd, tau=0.5"),
    "cfg2": dict(n=4096, d=768, precision="bf16", path="tc", two_keys=False, sizes="SIZES_CFG2",
                 workload="cfg2: global_alignment_loss fwd+bwd, global batch 4096 image-view/report pairs, D=768, bf16, "
                          "study sizes {1:.45,2:.45,3:.08,4:.02} shuffled, tau=0.5"),
    "cfg3": dict(n=16384, d=768, precision="bf16", path="tc", two_keys=False, sizes="SIZES_CFG3",
                 workload="cfg3: EVOKE multi-view multi-positive image<->text contrastive loss (global_alignment_loss) fwd+bwd, "
                          "global batch 16384 pairs, D=768, study sizes {1:.25,2:.45,3:.20,4:.10} shuffled, tau=0.5"),
    "cfg4": dict(n=32768, d=512, precision="bf16", path="tc", two_keys=True, sizes=None,
                 workload="cfg4: patient-specific positives (patient-id AND study-id keys, 1-3 studies per patient, 1-4 views per "
                          "study), global_alignment_loss fwd+bwd, global batch 32768, D=512, bf16, tau=0.5"),
}
FP_TOL = 2e-2           # north_star's bf16-mode gradient tolerance


def metric_name(cfg):
    c = CONFIGS[cfg]
    return METRIC if cfg == "cfg3" else f"contrastive loss fwd+bwd pairs/sec at N={c['n']},D={c['d']}"


def base_config(cfg):
    """The `config` object: identical in both arms (the driver compares them)."""
    c = CONFIGS[cfg]
    return {"workload": c["workload"], "global_batch": c["n"], "dim": c["d"],
            "l2": "no explicit flush: every step streams a bf16 E/W strip of 2*N^2/n_gpus bytes per GPU (537 MB at cfg3 on one "
                  "GPU) plus operands and gradients through the 126 MB L2, so no timed iteration starts with its inputs cached"
                  if c["path"] == "tc" else "inputs (0.4 MB) are L2-resident by nature of the workload: the step is launch-bound"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_tflops=float(p["bf16_tflops"]), bf16_sustained=float(p.get("bf16_tflops_sustained", 0.0)),
                    hbm_gbs=float(p["hbm_gbs"]), source="measured")
    except Exception:
        return dict(bf16_tflops=1590.0, bf16_sustained=1400.0, hbm_gbs=6650.0, source="fallback")


class ClockSampler:
    """Polls NVML for SM clock and clock-event reasons while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index: int, period_s: float = 0.01):
        self.samples, self.reason_bits, self.ok = [], 0, False
        self.period, self._stop = period_s, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:                                   # pragma: no cover
            self.err = repr(e)
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                try:
                    self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    self.reason_bits |= int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self.ok:
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.ok and self.thread.is_alive():
            self.thread.join(timeout=1.0)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml_unavailable"]}
        s = sorted(self.samples)
        reasons = [n for b, n in self.REASONS.items() if self.reason_bits & b]
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(s)}


def physical_gpu_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# ------------------------------------------------------------------------------------ workloads
def make_workload(cfg: str):
    """Global (unsharded) numpy arrays of a config: dict(image, text, key, key2 | None, view | None, view_ids | None).
    The same arrays oracle/make_fingerprints.py and tests/test_gpu_fullsize.py build (same seeds)."""
    import numpy as np
    from evoke_b200 import synth
    c = CONFIGS[cfg]
    n, d = c["n"], c["d"]
    if cfg == "cfg1":
        ids64 = np.concatenate([np.arange(32), np.arange(32)]).astype(np.int32)          # 32 studies x 2 views
        view = synth.make_embeddings(ids64, d, seed=1237)
        return dict(image=synth.make_embeddings(ids64[:32], d, seed=1235), text=synth.make_embeddings(ids64[:32], d, seed=1236),
                    key=ids64[:32].copy(), key2=None, view=view, view_ids=ids64)
    if c["two_keys"]:
        pat, stu = synth.make_patient_study_ids(n, seed=1234)
        return dict(image=synth.make_embeddings(stu, d, seed=1235), text=synth.make_embeddings(stu, d, seed=1236),
                    key=pat, key2=stu, view=None, view_ids=None)
    ids = synth.make_study_ids(n, getattr(synth, c["sizes"]), seed=1234)
    return dict(image=synth.make_embeddings(ids, d, seed=1235), text=synth.make_embeddings(ids, d, seed=1236),
                key=ids, key2=None, view=None, view_ids=None)


def reference_ids(w):
    """ids as the reference's loader hands them over: one key per row (two components -> the conjunction)."""
    import numpy as np
    if w["key2"] is None:
        return w["key"]
    return w["key"].astype(np.int64) * (int(w["key2"].max()) + 1) + w["key2"].astype(np.int64)


# ------------------------------------------------------------------------------------ reference arm
def port_time(cfg: str, n: int, reps: int, warmup: int, threads: int, device: str = "cpu"):
    """Seconds per fwd+bwd of the oracle's PyTorch port of the reference (global_alignment_loss; cfg1: + the multi-view
    term) on the first n rows of the workload."""
    import numpy as np
    import torch
    from oracle import evoke_oracle as orc
    if device == "cpu":
        torch.set_num_threads(threads)
    w = make_workload(cfg)
    ids = reference_ids(w)[:n]
    image = torch.tensor(w["image"][:n], device=device, requires_grad=True)
    text = torch.tensor(w["text"][:n], device=device, requires_grad=True)
    view = torch.tensor(w["view"], device=device, requires_grad=True) if w["view"] is not None else None
    times = []
    for it in range(warmup + reps):
        image.grad = text.grad = None
        if device != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        loss = orc.global_alignment_loss_port(image, text, ids, TAU)
        if view is not None:
            view.grad = None
            loss = loss + orc.multi_pos_contra_images_port(view, w["view_ids"], TAU)
        loss.backward()
        if device != "cpu":
            torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return float(np.mean(times)), float(loss.item())


def host_mem_gb() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


def cpu_model() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def cpu_sample_size(cfg: str, threads: int, total_steps: int, budget_s: float):
    """Largest power-of-two slice of the workload whose `total_steps` CPU steps fit in budget_s (and in host memory:
    the port materialises ~10 N x N fp32 temporaries)."""
    n_full = CONFIGS[cfg]["n"]
    if n_full <= 2048:
        return n_full, None
    t_probe, _ = port_time(cfg, 2048, reps=1, warmup=1, threads=threads)
    n_s = n_full
    mem = host_mem_gb()
    while n_s > 2048 and (t_probe * (n_s / 2048.0) ** 2 * total_steps > budget_s or 40.0 * n_s * n_s / 1e9 > 0.6 * mem):
        n_s //= 2
    return n_s, t_probe


def run_reference(args):
    """The reference's own CPU implementation of the path (its PyTorch op sequence, restated in
    oracle/evoke_oracle.py because the Python reference cannot travel to the GPU box), all host
    threads.  Each step is a bounded sample: the full batch when K+W such steps fit in ~3 minutes, otherwise the
    largest power-of-two N_s that does, converted to the metric's unit with the O(N^2) work ratio (stated in `sample`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = args.config
    n_full, d = CONFIGS[cfg]["n"], CONFIGS[cfg]["d"]
    threads = os.cpu_count() or 1
    n_s, _ = cpu_sample_size(cfg, threads, args.steps + args.warmup, 170.0)
    t_step, loss = port_time(cfg, n_s, reps=args.steps, warmup=args.warmup, threads=threads)
    scale = (n_full / n_s) ** 2
    t_full = t_step * scale
    value = n_full / t_full
    sample = (f"G loss fwd+bwd{' + MPC' if cfg == 'cfg1' else ''}, PyTorch CPU fp32 port of the reference op sequence, N_s={n_s}, D={d}, "
              f"{threads} threads ({cpu_model()}), {args.steps} reps after {args.warmup} warm-up")
    if n_s != n_full:
        sample += f"; time scaled by ({n_full}/{n_s})^2 = {scale:.0f}x (O(N^2) work) to the N={n_full} workload"
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_full * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(cfg),
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "loss": loss,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import evoke_b200
    from evoke_b200 import _lib, synth  # noqa: F401
    from evoke_b200.ids import DeviceIds

    cfg = args.config
    C = CONFIGS[cfg]
    n_global, dim, precision, path = C["n"], C["d"], C["precision"], C["path"]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun --nproc-per-node {args.gpus} (one rank per GPU)")
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU path)")
    if cfg == "cfg1" and world > 1:
        raise SystemExit("cfg1 is the reference's single-device batch (32 studies): run it with --gpus 1")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
        from evoke_b200.distributed import global_alignment_sharded
    if n_global % world:
        raise SystemExit("global batch must divide by the number of ranks")
    n_loc = n_global // world
    lo, hi = rank * n_loc, (rank + 1) * n_loc

    w = make_workload(cfg)
    img_np, txt_np = w["image"][lo:hi], w["text"][lo:hi]
    image = torch.tensor(img_np, device=dev, requires_grad=True)
    text = torch.tensor(txt_np, device=dev, requires_grad=True)
    key_np = np.ascontiguousarray(w["key"][lo:hi])
    key2_np = None if w["key2"] is None else np.ascontiguousarray(w["key2"][lo:hi])
    ids_dev = DeviceIds(torch.from_numpy(key_np).to(dev), None if key2_np is None else torch.from_numpy(key2_np).to(dev))
    view = torch.tensor(w["view"], device=dev, requires_grad=True) if w["view"] is not None else None
    view_ids = w["view_ids"]

    def public_call(x, y, ids, graph):
        """One fwd+bwd through the public API (the call a user's training loop makes)."""
        if world > 1:
            l = global_alignment_sharded(x, y, ids, TAU, precision=precision, mode=args.shard_mode, graph=graph)
        else:
            l = evoke_b200.global_alignment(x, y, ids, TAU, precision=precision, path=path, graph=graph)
            if view is not None:                                   # cfg1: the multi-view term over all 64 views (:536)
                view.grad = None
                l = l + evoke_b200.multi_pos_contra_images(view, view_ids, TAU, precision=precision, path=path)
        l.backward()
        return l

    def make_graphed():
        g = evoke_b200.GraphedGlobalAlignment(n_loc, dim, TAU, device=dev, precision=precision, path=path,
                                              two_keys=C["two_keys"], sharded=world > 1, shard_mode=args.shard_mode)
        g.load(image.detach(), text.detach(), ids_dev.key, ids_dev.key2)
        g._warmup = max(args.warmup, 3)
        return g.capture()

    # cfg1 has no single-graph form (its multi-view term sizes itself from the ids): its step is the public call
    use_graph = not args.no_graph and cfg != "cfg1"
    graphed = None
    launches_per_replay = 0
    if use_graph:
        before = _lib.launch_count
        graphed = make_graphed()
        # kernels of this library inside ONE replay = launches seen while capturing (warm-up excluded)
        launches_per_replay = (_lib.launch_count - before) // (max(args.warmup, 3) + 1)

    def step():
        if graphed is not None:
            return graphed.step()
        image.grad = None
        text.grad = None
        return public_call(image, text, ids_dev if cfg != "cfg1" else key_np, graph=None if cfg == "cfg1" else False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # per-kernel CUDA events around the tcgen05 launches (recorded on the launching stream)
    TC = ("evk_mpce_fwd", "evk_mpce_fwd_store", "evk_mpce_bwd_w", "evk_mpce_bwd_gemm", "evk_mpce_bwd_gemm_scatter")
    HBM_K = ("evk_mpce_w_from_e",)        # K4t: 4 B per (i, j) read+written
    ev = {}                               # every entry point of the library gets its own event pairs
    pending = {}

    def hook(name, phase):
        ev.setdefault(name, [])
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream())
        if phase == "before":
            pending[name] = e
        else:
            ev[name].append((pending.pop(name), e))

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    if not args.no_clocks:
        sampler.start()
    launches0 = _lib.launch_count
    t_beg, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_beg.record()
    for _ in range(args.steps):
        loss = step()
    t_end.record()
    barrier()
    launches = _lib.launch_count - launches0
    if graphed is not None:
        launches = launches_per_replay * args.steps
    ms_total = t_beg.elapsed_time(t_end)
    loss_val = float(loss.item())

    # ---- per-kernel timing pass (eager launches, CUDA events around every entry point, same stream), right after the
    # timed replays so that it sees the same clocks; it uses its own input tensors, not the graph's static buffers
    ms_eager_total = None
    if not args.no_kernel_events:
        def eager_step():
            image.grad = None
            text.grad = None
            return public_call(image, text, ids_dev if cfg != "cfg1" else key_np, graph=False)
        for _ in range(3):
            eager_step()
        barrier()
        _lib.call_hook = hook
        e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_beg.record()
        for _ in range(args.steps):
            eager_step()
        e_end.record()
        barrier()
        _lib.call_hook = None
        ms_eager_total = e_beg.elapsed_time(e_end)
    # ---- gradient fingerprint of the step that was just timed (graph replay: static .grad buffers)
    if graphed is not None:
        g_img, g_txt = graphed.image.grad, graphed.text.grad
    else:
        g_img, g_txt = image.grad, text.grad
    fingerprint = None
    fp_path = os.path.join(ROOT, "tests", "golden", f"fingerprint_{cfg}.json")
    if os.path.isfile(fp_path):
        with open(fp_path) as f:
            want = json.load(f)
        rows = want["rows"]
        sq = torch.stack([(g_img.double() ** 2).sum(), (g_txt.double() ** 2).sum()])
        got_rows = torch.zeros((2, len(rows), dim), dtype=torch.float64, device=dev)
        for k, r in enumerate(rows):
            if lo <= r < hi:
                got_rows[0, k] = g_img[r - lo].double()
                got_rows[1, k] = g_txt[r - lo].double()
        if world > 1:
            dist.all_reduce(sq)
            dist.all_reduce(got_rows)
        norms = sq.sqrt().tolist()
        row_norms = got_rows.norm(dim=2).cpu().numpy()
        heads = got_rows[:, :, :4].cpu().numpy()
        rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())
        errs = {"loss": abs(loss_val - want["loss"]) / abs(want["loss"]),
                "d_image_norm": abs(norms[0] - want["d_image_norm"]) / want["d_image_norm"],
                "d_text_norm": abs(norms[1] - want["d_text_norm"]) / want["d_text_norm"],
                "d_image_row_norms": rel(row_norms[0], want["d_image_row_norms"]),
                "d_text_row_norms": rel(row_norms[1], want["d_text_row_norms"]),
                "d_image_row_head": rel(heads[0], want["d_image_row_head"]),
                "d_text_row_head": rel(heads[1], want["d_text_row_head"])}
        worst = max(v for k, v in errs.items() if k != "loss")
        fingerprint = {"loss": loss_val, "d_image_norm": norms[0], "d_text_norm": norms[1],
                       "d_image_row_norms": [float(v) for v in row_norms[0]], "rel_err_vs_fp64_oracle": errs,
                       "tol": FP_TOL, "loss_tol": 1e-5, "ok": bool(worst <= FP_TOL and errs["loss"] <= 1e-5),
                       "reference": f"tests/golden/fingerprint_{cfg}.json (oracle/make_fingerprints.py, fp64)"}

    # ---- drop-in: the reference-signature call with device inputs, as a training loop makes it (graph cache on);
    # measured BEFORE the sustained loop so that it sees the clocks of `value`
    def timed_calls(fn, k):
        for _ in range(3):
            fn()
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / k

    dropin = None
    if not args.no_dropin:
        if world == 1:
            model = evoke_b200.ContrastiveObjective(instance_temp=TAU, region_temp=TAU, precision=precision)
            ids_arg = ids_dev if cfg != "cfg1" else key_np         # cfg1: the reference's host numpy ids

            def call_dropin():
                image.grad = None
                text.grad = None
                l = model.global_alignment_loss(image, text, ids_arg)
                if view is not None:
                    view.grad = None
                    l = l + model.multi_pos_contra_images_v0401(view, view_ids)
                l.backward()
        else:
            def call_dropin():
                image.grad = None
                text.grad = None
                global_alignment_sharded(image, text, ids_dev, TAU, precision=precision, mode=args.shard_mode).backward()
        ms_d = timed_calls(call_dropin, args.steps)
        dropin = {"ms_per_step": ms_d, "value": n_global / (ms_d * 1e-3), "unit": "pairs/s",
                  "api": ("ContrastiveObjective.global_alignment_loss(image, text, ids) + .backward() [= patch_pretrain's method]"
                          if world == 1 else "global_alignment_sharded(image, text, ids, tau) + .backward()"),
                  "launch_mode": "cuda-graph cache (forward graph + backward graph, evoke_b200.graphs.GraphedStep)"}
        if world == 1 and cfg != "cfg1" and not args.no_fp32:
            model32 = evoke_b200.ContrastiveObjective(instance_temp=TAU, region_temp=TAU, precision="fp32")

            def call_fp32():
                image.grad = None
                text.grad = None
                model32.global_alignment_loss(image, text, ids_dev).backward()
            ms32 = timed_calls(call_fp32, max(3, args.steps // 4))
            dropin["fp32_parity_mode"] = {"ms_per_step": ms32, "value": n_global / (ms32 * 1e-3), "unit": "pairs/s",
                                          "note": "default precision of the drop-in: 3-segment split-bf16 operands, "
                                                  "loss <= 1e-5 / gradients <= 1e-4 vs the reference; 8*3 N^2 D executed FLOP"}

    # ---- sustained: the same replay for >= 1 s
    sustained = None
    if graphed is not None and not args.no_sustained:
        # every rank must replay the SAME number of steps (the sharded step contains cross-GPU syncs): agree on it
        t_loc = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_loc, op=dist.ReduceOp.MAX)
        reps = max(args.steps, int(1.05e3 / max(float(t_loc.item()) / args.steps, 1e-3)) + 1)
        s_beg, s_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s_beg.record()
        for _ in range(reps):
            graphed.step()
        s_end.record()
        barrier()
        s_ms = s_beg.elapsed_time(s_end)
        if world > 1:
            t = torch.tensor([s_ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            s_ms = float(t.item())
        sustained = {"steps": reps, "seconds": s_ms / 1e3, "ms_per_step": s_ms / reps, "value": n_global / (s_ms / reps * 1e-3),
                     "unit": "pairs/s"}

    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = n_global / (ms_step * 1e-3)

    # roofline of the slowest tcgen05 kernel: algorithmic FLOP per launch = 2 * rows * cols * D
    peaks = load_peaks()
    flop_launch = 2.0 * n_loc * n_global * dim
    kern = {}
    for name, pairs in ev.items():
        if pairs:
            ms = [a.elapsed_time(b) for a, b in pairs]
            kern[name] = dict(launches_per_step=len(pairs) / args.steps, avg_ms=float(np.mean(ms)),
                              share_of_step=float(np.sum(ms)) / (ms_eager_total or ms_total))
            if name in HBM_K:
                kern[name]["gbs"] = (4.0 * n_loc * n_global) / (float(np.mean(ms)) * 1e-3) / 1e9
                kern[name]["frac_of_hbm_peak"] = kern[name]["gbs"] / peaks["hbm_gbs"]
            elif name in TC:
                kern[name]["tflops"] = flop_launch / (float(np.mean(ms)) * 1e-3) / 1e12
    tck = [k for k in kern if k in TC]
    dom = max(tck, key=lambda k: kern[k]["avg_ms"]) if tck else None
    traffic = None          # (N > 1: ncu cannot wrap a multi-rank command, so there is no per-launch DRAM capture)
    if world == 1 and cfg == "cfg3":
        try:
            with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
                traffic = json.load(f).get(dom)
        except Exception:
            pass
    roofline = None
    if dom:
        roofline = {"bound": "tensor", "kernel": dom, "achieved": kern[dom]["tflops"], "peak": peaks["bf16_tflops"],
                    "unit": "TFLOP/s", "frac": kern[dom]["tflops"] / peaks["bf16_tflops"], "traffic": traffic,
                    "peak_source": peaks["source"] + " (MEASURED_PEAKS.json bf16_tflops, burst)",
                    "flop_per_launch": flop_launch,
                    "timing": "CUDA events around the launch on its stream, eager pass of the same steps right after the "
                              "timed graph replays (graph nodes cannot be bracketed)"}
    alg = 6.0 * n_global * n_global * dim
    step_tflops = alg / (ms_step * 1e-3) / 1e12
    roofline_step = {"algorithmic_flop": alg, "achieved": step_tflops,
                     "peak": peaks["bf16_tflops"] * world, "unit": "TFLOP/s",
                     "frac": step_tflops / (peaks["bf16_tflops"] * world),
                     "frac_of_sustained": step_tflops / (peaks["bf16_sustained"] * world) if peaks["bf16_sustained"] else None}
    if sustained is not None and peaks["bf16_sustained"]:
        sustained["frac_of_sustained_peak"] = alg / (sustained["ms_per_step"] * 1e-3) / 1e12 / (peaks["bf16_sustained"] * world)

    # ---- e2e: public API, HOST (pinned) inputs, H2D + loss D2H inside the timed region.  The next
    # step's H2D (embeddings AND ids, all on one copy stream so nothing on the compute stream
    # queues behind it in the copy engine) overlaps the current step's compute.
    h_img = torch.from_numpy(np.ascontiguousarray(img_np)).pin_memory()
    h_txt = torch.from_numpy(np.ascontiguousarray(txt_np)).pin_memory()
    key_pinned = torch.from_numpy(key_np).pin_memory()
    key2_pinned = None if key2_np is None else torch.from_numpy(key2_np).pin_memory()
    copy_stream = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    if graphed is not None:
        # ping-pong two captured steps; H2D lands directly in each graph's static input buffers
        slots = [graphed, make_graphed()]
    else:
        bufs = [(torch.empty_like(image), torch.empty_like(text), torch.empty_like(ids_dev.key),
                 None if ids_dev.key2 is None else torch.empty_like(ids_dev.key2)) for _ in range(2)]

    def prefetch(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            if graphed is not None:
                slots[slot].load(h_img, h_txt, key_pinned, key2_pinned)
            else:
                bufs[slot][0].copy_(h_img, non_blocking=True)
                bufs[slot][1].copy_(h_txt, non_blocking=True)
                bufs[slot][2].copy_(key_pinned, non_blocking=True)
                if key2_pinned is not None:
                    bufs[slot][3].copy_(key2_pinned, non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(k):
        for s_ in range(2):
            consumed[s_].record(torch.cuda.current_stream())
        prefetch(0)
        last = None
        for i in range(k):
            slot = i & 1
            if i + 1 < k:
                prefetch(slot ^ 1)
            torch.cuda.current_stream().wait_event(ready[slot])
            if graphed is not None:
                l = slots[slot].step()
            else:
                x = bufs[slot][0].detach().requires_grad_(True)
                y = bufs[slot][1].detach().requires_grad_(True)
                l = public_call(x, y, DeviceIds(bufs[slot][2], bufs[slot][3]) if cfg != "cfg1" else key_np,
                                graph=None if cfg == "cfg1" else False)
            consumed[slot].record(torch.cuda.current_stream())
            last = l.item()                                        # D2H read of the step's result
        return last

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    h2d = int((h_img.numel() + h_txt.numel()) * 4 + key_np.nbytes + (0 if key2_np is None else key2_np.nbytes)) * world
    e2e = {"value": n_global * args.steps / e2e_s, "unit": "pairs/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": 4 * world, "ms_per_step": e2e_s / args.steps * 1e3,
           "note": f"fp32 embeddings + int32 ids from pinned host memory every step; the next step's H2D overlaps compute"
                   f"{'; PCIe-bound at one GPU (%.1f MB/step)' % (h2d / 1e6) if world == 1 and h2d > 5e7 else ''}"}

    cpu_baseline = cuda_eager = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_s, _ = cpu_sample_size(cfg, threads, 3, 30.0)              # 1 warm-up + 2 reps in ~30 s
        reps = 2 if n_s > 2048 else 20
        t_cpu, _ = port_time(cfg, n_s, reps=reps, warmup=1, threads=threads)
        scale = (n_global / n_s) ** 2
        sample = (f"{reps} fwd+bwd (after 1 warm-up) of the PyTorch-CPU fp32 port at N_s={n_s}, D={dim}, {threads} threads "
                  f"({cpu_model()})")
        if n_s != n_global:
            sample += f"; time scaled by {scale:.0f}x (O(N^2)) to N={n_global}"
        cpu_baseline = {"value": n_global / (t_cpu * scale), "unit": "pairs/s", "cores": threads, "kind": "port",
                        "sample": sample}
    if world == 1 and rank == 0 and not args.no_cuda_eager:
        # the reference's own op sequence on this GPU: fp32, TF32 off (PyTorch's default for matmul), host-built labels
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        try:
            t_eager, loss_eager = port_time(cfg, n_global, reps=5, warmup=2, threads=1, device=str(dev))
            cuda_eager = {"value": n_global / t_eager, "unit": "pairs/s", "ms_per_step": t_eager * 1e3, "loss": loss_eager,
                          "what": "oracle's PyTorch port of the reference op sequence (:486-504) in CUDA eager fp32 on this GPU, "
                                  "TF32 off, N x N label matrix built on the host and copied every step as the reference does; "
                                  "wall clock with device sync, 5 reps after 2 warm-up"}
            torch.cuda.empty_cache()
        except torch.cuda.OutOfMemoryError as e:                   # pragma: no cover
            cuda_eager = {"unavailable": repr(e)[:200]}

    transport = "none"
    if world > 1:
        from evoke_b200 import peer as _peer
        pcs = [c for c in _peer._CONTEXTS.values() if isinstance(c, _peer.PeerContext)]
        if pcs:
            transport = (f"peer memory over NVLink (this library's kernels: all-gather stores, GEMM-epilogue scatter of "
                         f"{pcs[0].exchange} partials, flag barriers; CUDA IPC)")
            for c in pcs:
                c.check()                                  # a barrier that timed out invalidates the run
        else:
            transport = "NCCL (all-gather / all-reduce / reduce-scatter)"
    ok = True
    if rank == 0:
        line = {
            "metric": metric_name(cfg), "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
            "config": base_config(cfg),
            "config_detail": {"name": cfg, "rows_per_rank": n_loc,
                              "precision": ("bf16 operands, fp32 accumulate/statistics; fp32 inputs and gradients" if precision == "bf16"
                                            else "fp32 SIMT kernels (small path)"),
                              "parallelism": (f"dp{world} row shards, shard_mode={args.shard_mode}, transport={transport}"
                                              if world > 1 else "single GPU")},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
            "roofline_step": roofline_step, "sustained": sustained, "dropin": dropin, "fingerprint": fingerprint,
            "kernels": kern, "cpu_baseline": cpu_baseline, "cuda_eager_baseline": cuda_eager, "loss": loss_val,
            "launch_mode": "cuda_graph" if graphed is not None else "public API calls (graph cache for the G loss)",
            "ms_per_step_eager": (ms_eager_total / args.steps) if ms_eager_total else None,
        }
        print(json.dumps(line), flush=True)
        if fingerprint is not None and not fingerprint["ok"]:
            ok = False
            sys.stderr.write(f"bench.py: GRADIENT FINGERPRINT MISMATCH vs the fp64 oracle: {fingerprint['rel_err_vs_fp64_oracle']}\n")
    if world > 1:
        # captured graphs hold NCCL work: drop them and drain the device before tearing the group
        # down, and do not let a slow communicator teardown keep the process alive
        sys.stdout.flush()
        try:
            slots = graphed = None                                # noqa: F841
            torch.cuda.synchronize()
            dist.barrier()
        finally:
            os._exit(0 if ok else 3)
    if not ok:
        raise SystemExit(3)


def run_cfg5(args):
    """BASELINE.json config 5: the full pre-training step (tools/pretrain_step.py), timed with this library's losses and -
    same model, same data - with the reference's loss op sequences in CUDA eager.  One JSON line (rank 0)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import pretrain_step
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"WORLD_SIZE={world} does not match --gpus {args.gpus}")
    steps, warmup = min(args.steps, 20), max(min(args.warmup, 5), 3)
    if args.impl == "reference":
        out = pretrain_step.run(steps, warmup, 32, "reference", True)
        ref = None
    else:
        out = pretrain_step.run(steps, warmup, 32, "evoke_b200", True)
        torch.cuda.empty_cache()
        ref = pretrain_step.run(max(3, steps // 2), warmup, 32, "reference", True)
    if rank == 0:
        workload = ("cfg5: full EVOKE pre-training step on synthetic 224x224 multi-view CXRs: " + out["model"] +
                    "; all_loss = instance + sen_text + mul_pos, clip_grad_value_(0.1), Adam; bf16 autocast encoders")
        line = {"metric": "pretrain step pairs/sec (32 studies per GPU)", "value": out["value"], "unit": "pairs/s",
                "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": out["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload, "global_batch": 32 * world, "dim": 768,
                           "l2": "every step runs ResNet-101 and BERT over fresh synthetic images: nothing stays cached"},
                "gpu_launches": out["gpu_launches"], "loss": out["loss"], "ms_loss_forward": out["ms_loss_forward"],
                "losses": out["losses"], "roofline": None, "cpu_baseline": None,
                "e2e": {"value": out["value"], "unit": "pairs/s", "h2d_bytes_per_step": int(69 * 3 * 224 * 224 * 4 + 32 * 100 * 16) * world,
                        "d2h_bytes_per_step": 4 * world,
                        "note": "the step's images and token ids are generated on the host and copied H2D inside the timed step"},
                "same_model_reference_losses": None if ref is None else {
                    "value": ref["value"], "ms_per_step": ref["ms_per_step"], "ms_loss_forward": ref["ms_loss_forward"],
                    "what": "identical model, data and optimiser; the three losses as the reference's PyTorch op sequence in "
                            "CUDA eager on each rank's own batch (oracle ports)"}}
        if args.impl == "reference":
            line["impl"] = "reference"
        print(json.dumps(line), flush=True)
    if world > 1:
        sys.stdout.flush()
        os._exit(0)


def run_f1(args):
    """f1: Pretrain.local_text_token_alignment_loss (:506-526) at the reference's shape - B = 32 samples,
    L - 1 = 99 text tokens, P = 49 patch tokens, D = 768 - forward + backward.  One JSON line."""
    import numpy as np
    import torch
    import evoke_b200
    from evoke_b200 import _lib
    from oracle import evoke_oracle as orc
    b, l, pt, d = 32, 99, 49, 768
    rng = np.random.default_rng(1234)
    v_np = rng.standard_normal((b, pt, d)).astype(np.float32)
    t_np = rng.standard_normal((b, l, d)).astype(np.float32)

    def port_ms(device, reps, warm):
        v = torch.tensor(v_np, device=device, requires_grad=True)
        t = torch.tensor(t_np, device=device, requires_grad=True)
        ts = []
        for it in range(warm + reps):
            v.grad = t.grad = None
            if device != "cpu":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            loss = orc.local_text_token_alignment_port(v, t, TAU)
            loss.backward()
            if device != "cpu":
                torch.cuda.synchronize()
            if it >= warm:
                ts.append(time.perf_counter() - t0)
        return float(np.mean(ts)) * 1e3, float(loss.item())

    workload = ("f1: local_text_token_alignment_loss fwd+bwd at the reference's shape: 32 samples x 99 text tokens x 49 patch "
                "tokens, D=768, fp32, tau=0.5")
    config = {"workload": workload, "global_batch": b, "dim": d,
              "l2": "inputs (7.4 MB) stay L2-resident by nature of the workload: the step is launch / latency bound"}
    if args.impl == "reference":
        torch.set_num_threads(os.cpu_count() or 1)
        ms, loss = port_ms("cpu", args.steps, args.warmup)
        line = {"impl": "reference", "metric": "local token alignment fwd+bwd samples/sec at B=32,L=99,P=49,D=768", "value": b / (ms * 1e-3),
                "unit": "samples/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": b / (ms * 1e-3), "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
                                 "sample": "the full workload, PyTorch-CPU port of the reference op sequence"},
                "e2e": {"value": b / (ms * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "loss": loss}
        print(json.dumps(line), flush=True)
        return
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    g = evoke_b200.GraphedLocalTokenAlign(b, pt, l, d, TAU, device=dev)
    g.load(torch.tensor(v_np, device=dev), torch.tensor(t_np, device=dev))
    before = _lib.launch_count
    g.capture()
    per_replay = (_lib.launch_count - before) // 4
    for _ in range(max(args.warmup, 3)):
        g.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(physical_gpu_index(0))
    sampler.start()
    a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = g.step()
    z.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(z) / args.steps
    clocks = sampler.stop()
    # drop-in (eager launches of the method) and e2e (pinned host inputs every step)
    v = torch.tensor(v_np, device=dev, requires_grad=True)
    t = torch.tensor(t_np, device=dev, requires_grad=True)
    for _ in range(3):
        evoke_b200.local_text_token_alignment(v, t, TAU).backward()
    torch.cuda.synchronize()
    a.record()
    for _ in range(args.steps):
        v.grad = t.grad = None
        evoke_b200.local_text_token_alignment(v, t, TAU).backward()
    z.record()
    torch.cuda.synchronize()
    ms_eager = a.elapsed_time(z) / args.steps
    hv, ht = torch.from_numpy(v_np).pin_memory(), torch.from_numpy(t_np).pin_memory()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.load(hv, ht)
        last = g.step().item()
    e2e_s = (time.perf_counter() - t0) / args.steps
    want = orc.local_token_alignment_closed_form(v_np, t_np, TAU)[0]
    ms_cuda, _ = port_ms(str(dev), 20, 5)
    torch.set_num_threads(os.cpu_count() or 1)
    ms_cpu, _ = port_ms("cpu", 10, 2)
    flop = 2.0 * b * (2 * l * pt * d + 2 * l * l * d) * 3           # forward products x (1 fwd + 2 bwd)
    line = {"metric": "local token alignment fwd+bwd samples/sec at B=32,L=99,P=49,D=768", "value": b / (ms * 1e-3), "unit": "samples/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config, "clocks": clocks,
            "gpu_launches": per_replay * args.steps, "loss": float(loss.item()), "loss_rel_err_vs_fp64_oracle": abs(last - want) / abs(want),
            "ms_per_step_eager": ms_eager,
            "e2e": {"value": b / e2e_s, "unit": "samples/s", "h2d_bytes_per_step": int(hv.numel() + ht.numel()) * 4, "d2h_bytes_per_step": 4},
            "roofline": {"bound": "latency", "achieved": flop / (ms * 1e-3) / 1e12, "peak": None, "unit": "TFLOP/s", "frac": None,
                         "traffic": None, "note": "~1.1 GFLOP in ~20 dependent fp32 SIMT launches: neither HBM nor the tensor pipe is the bound"},
            "cpu_baseline": {"value": b / (ms_cpu * 1e-3), "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "the full workload, 10 reps after 2 warm-up"},
            "cuda_eager_baseline": {"value": b / (ms_cuda * 1e-3), "unit": "samples/s", "ms_per_step": ms_cuda,
                                    "what": "the reference's op sequence (:506-526) in PyTorch CUDA eager fp32 on this GPU"}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="evoke_b200", choices=["evoke_b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS) + ["cfg5", "f1"],
                    help="workload (BASELINE.json configs): cfg3 = the metric's configuration (default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-eager", action="store_true", help="skip the CUDA-eager reference-port baseline")
    ap.add_argument("--no-dropin", action="store_true", help="skip the reference-signature (drop-in) timing")
    ap.add_argument("--no-fp32", action="store_true", help="skip the fp32-parity-mode drop-in line")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 1 s sustained replay")
    ap.add_argument("--no-kernel-events", action="store_true", help="debug: skip per-kernel CUDA events")
    ap.add_argument("--no-clocks", action="store_true", help="debug: skip the NVML clock sampler")
    ap.add_argument("--no-graph", action="store_true", help="run the eager launch sequence instead of the CUDA graph")
    ap.add_argument("--shard-mode", default="auto", choices=["auto", "rs", "sym", "peer"],
                    help="N>1: peer = exchanges by this library's kernels over peer-mapped memory (default when possible); "
                         "rs / sym = NCCL transports (reduce-scatter of partial dK / recomputed key-side block)")
    args = ap.parse_args()
    if args.config == "cfg5":
        run_cfg5(args)
    elif args.config == "f1":
        run_f1(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
