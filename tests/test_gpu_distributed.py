"""T4 (GPU, NCCL, world_size 2 / 4 / 8): sharded G loss == single-GPU G loss on the concatenated batch,
and both == the fp64 oracle.  Each case is skipped on boxes with fewer GPUs than ranks.  Also: the reference's
real argument types on the sharded path (strided [:,0,:] head views, bf16 inputs) and a rank that misses a
barrier (NaN loss and gradients + a host-side error, never a silently wrong step)."""
import os
import sys
import types

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _head_view(x: np.ndarray, tokens: int, dtype=torch.float32) -> torch.Tensor:
    """x [n, d] as the reference hands it over (v0520.py:484,399): the [:,0,:] slice of the permuted projection-head
    output [n, d, 1+P] -> strides (d*(1+P), 1+P); a leaf-like tensor that requires grad."""
    n, d = x.shape
    full = torch.zeros((n, d, tokens), device="cuda", dtype=dtype)
    full[:, :, 0] = torch.tensor(x, device="cuda").to(dtype)
    full.requires_grad_(True)
    return full


def _worker(rank, world, port, n_total, d, tau, precision, mode, out_dir, views=False, in_dtype="float32"):
    sys.path.insert(0, ROOT)
    if mode == "peer-gather":                        # the all-gather pushed from a side stream WHILE K3 sweeps
        os.environ["EVOKE_B200_OVERLAP_GATHER"] = "1"
        mode = "peer"
    elif mode.startswith("peer-"):                   # "peer-fp32": fp32 partials on NVLink instead of bf16
        os.environ["EVOKE_B200_PEER_EXCHANGE"] = mode.split("-")[1]
        mode = "peer"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from evoke_b200 import synth
        from evoke_b200.distributed import global_alignment_sharded
        ids = synth.make_study_ids(n_total, seed=31)
        xi = synth.make_embeddings(ids, d, seed=32)
        xt = synth.make_embeddings(ids, d, seed=33)
        n = n_total // world
        sl = slice(rank * n, (rank + 1) * n)
        dt = getattr(torch, in_dtype)
        if views:
            image_full, text_full = _head_view(xi[sl], 50, dt), _head_view(xt[sl], 7, dt)
            image, text = image_full, text_full
        else:
            image = torch.tensor(xi[sl], device="cuda", dtype=dt, requires_grad=True)
            text = torch.tensor(xt[sl], device="cuda", dtype=dt, requires_grad=True)
        reps = 3 if mode == "peer" else 1             # the symmetric buffers and barrier epochs are reused step to step
        for _ in range(reps):
            image.grad = text.grad = None
            a, b = (image.permute(0, 2, 1)[:, 0, :], text.permute(0, 2, 1)[:, 0, :]) if views else (image, text)
            loss = global_alignment_sharded(a, b, ids[sl].copy(), tau, precision=precision, mode=mode)
            loss.backward()
        torch.cuda.synchronize()
        if views:                                    # gradient of the head output: non-zero in token 0 only
            assert float(image.grad[:, :, 1:].abs().max()) == 0.0
            image = types.SimpleNamespace(grad=image.grad[:, :, 0].float())
            text = types.SimpleNamespace(grad=text.grad[:, :, 0].float())
        if mode == "peer":
            from evoke_b200 import peer
            for c in peer._CONTEXTS.values():
                if isinstance(c, peer.PeerContext):
                    c.check()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.float().item(), d_image=image.grad.float().cpu().numpy(),
                 d_text=text.grad.float().cpu().numpy())
    finally:
        dist.destroy_process_group()


def _graph_worker(rank, world, port, n_total, d, tau, out_dir):
    """Peer-memory path captured in a CUDA graph and replayed on fresh inputs."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import evoke_b200
        from evoke_b200 import synth
        n = n_total // world
        sl = slice(rank * n, (rank + 1) * n)
        g = evoke_b200.GraphedGlobalAlignment(n, d, tau, precision="bf16", path="tc", sharded=True, shard_mode="peer").capture()
        out = {}
        for seed in (41, 51):
            ids = synth.make_study_ids(n_total, seed=seed)
            xi = synth.make_embeddings(ids, d, seed=seed + 1)
            xt = synth.make_embeddings(ids, d, seed=seed + 2)
            g.load(torch.tensor(xi[sl], device="cuda"), torch.tensor(xt[sl], device="cuda"), torch.from_numpy(ids[sl].copy()).cuda())
            loss = g.step()
            torch.cuda.synchronize()
            out[f"loss{seed}"] = loss.item()
            out[f"d_image{seed}"] = g.image.grad.cpu().numpy()
            out[f"d_text{seed}"] = g.text.grad.cpu().numpy()
        np.savez(os.path.join(out_dir, f"grank{rank}.npz"), **out)
        g = None
    finally:
        dist.destroy_process_group()


def test_peer_path_in_a_cuda_graph_on_two_gpus(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 2048, 256, 0.5
    port = 29800 + (os.getpid() % 90)
    mp.spawn(_graph_worker, args=(world, port, n_total, d, tau, str(tmp_path)), nprocs=world, join=True)
    n = n_total // world
    for seed in (41, 51):
        ids = synth.make_study_ids(n_total, seed=seed)
        xi = synth.make_embeddings(ids, d, seed=seed + 1)
        xt = synth.make_embeddings(ids, d, seed=seed + 2)
        want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
        for r in range(world):
            got = np.load(tmp_path / f"grank{r}.npz")
            sl = slice(r * n, (r + 1) * n)
            assert abs(float(got[f"loss{seed}"]) - want) <= 2e-5 * abs(want)
            assert np.abs(got[f"d_image{seed}"] - d_i[sl]).max() <= 2e-2 * np.abs(d_i).max()
            assert np.abs(got[f"d_text{seed}"] - d_t[sl]).max() <= 2e-2 * np.abs(d_t).max()


def _port(salt: int) -> int:
    return 29900 + (os.getpid() % 50) * 2 + salt * 101 % 997


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", ["rs", "sym", "peer", "peer-fp32", "peer-gather"])
@pytest.mark.parametrize("precision,ltol,gtol", [("fp32", 1e-5, 1e-4), ("bf16", 2e-5, 2e-2)])
def test_sharded_equals_oracle(tmp_path, precision, ltol, gtol, mode, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if mode.startswith("peer") and precision != "bf16":
        pytest.skip("the peer-memory path is a bf16-mode path")
    if world > 2 and mode == "peer-fp32":
        pytest.skip("fp32 partials are covered at world 2")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    n_total, d, tau = 768 * world, 256, 0.5
    port = _port(world * 16 + (1 if precision == "fp32" else 0) + 2 * ["rs", "sym", "peer", "peer-fp32", "peer-gather"].index(mode))
    mp.spawn(_worker, args=(world, port, n_total, d, tau, precision, mode, str(tmp_path)), nprocs=world, join=True)
    ids = synth.make_study_ids(n_total, seed=31)
    xi = synth.make_embeddings(ids, d, seed=32)
    xt = synth.make_embeddings(ids, d, seed=33)
    want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    n = n_total // world
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        sl = slice(r * n, (r + 1) * n)
        assert abs(float(got["loss"]) - want) <= ltol * abs(want)
        assert np.abs(got["d_image"] - d_i[sl]).max() <= gtol * np.abs(d_i).max()
        assert np.abs(got["d_text"] - d_t[sl]).max() <= gtol * np.abs(d_t).max()


@pytest.mark.parametrize("mode,in_dtype", [("peer", "float32"), ("peer", "bfloat16"), ("rs", "float32"), ("sym", "bfloat16")])
def test_sharded_takes_the_reference_argument_types(tmp_path, mode, in_dtype):
    """Strided [:,0,:] head views (feature stride 1+P, v0520.py:484,399) and bf16 inputs through every transport."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 1024, 256, 0.5
    mp.spawn(_worker, args=(world, _port(40 + len(mode) + len(in_dtype)), n_total, d, tau, "bf16", mode, str(tmp_path), True, in_dtype),
             nprocs=world, join=True)
    ids = synth.make_study_ids(n_total, seed=31)
    xi = synth.make_embeddings(ids, d, seed=32)
    xt = synth.make_embeddings(ids, d, seed=33)
    if in_dtype == "bfloat16":                      # the oracle sees what the kernels see: bf16-rounded inputs
        xi = torch.tensor(xi).bfloat16().float().numpy()
        xt = torch.tensor(xt).bfloat16().float().numpy()
    want, d_i, d_t, _ = orc.g_loss_closed_form(xi, xt, ids, tau)
    n = n_total // world
    ltol = 2e-5 if in_dtype == "float32" else 4e-3          # a bf16 loss TENSOR carries 8 bits
    gtol = 2e-2 if in_dtype == "float32" else 3e-2          # + bf16 rounding of the returned gradients
    for r in range(world):
        got = np.load(tmp_path / f"rank{r}.npz")
        sl = slice(r * n, (r + 1) * n)
        assert abs(float(got["loss"]) - want) <= ltol * abs(want)
        assert np.abs(got["d_image"] - d_i[sl]).max() <= gtol * np.abs(d_i).max()
        assert np.abs(got["d_text"] - d_t[sl]).max() <= gtol * np.abs(d_t).max()


def _late_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["EVOKE_B200_PEER_TIMEOUT_MS"] = "300"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from evoke_b200 import peer, synth
        from evoke_b200.distributed import global_alignment_sharded
        n_total, d, tau = 1024, 128, 0.5
        ids = synth.make_study_ids(n_total, seed=3)
        n = n_total // world
        sl = slice(rank * n, (rank + 1) * n)
        image = torch.tensor(synth.make_embeddings(ids, d, seed=4)[sl], device="cuda", requires_grad=True)
        text = torch.tensor(synth.make_embeddings(ids, d, seed=5)[sl], device="cuda", requires_grad=True)
        loss = global_alignment_sharded(image, text, ids[sl].copy(), tau, precision="bf16", mode="peer")    # healthy step
        loss.backward()
        torch.cuda.synchronize()
        healthy = float(loss.item())
        dist.barrier()
        if rank == 1:
            torch.cuda._sleep(int(3.0e9))            # ~1.5 s of GPU time on this rank's stream: rank 0's barrier times out
        image.grad = text.grad = None
        loss = global_alignment_sharded(image, text, ids[sl].copy(), tau, precision="bf16", mode="peer")
        loss.backward()
        torch.cuda.synchronize()
        raised = False
        try:
            global_alignment_sharded(image, text, ids[sl].copy(), tau, precision="bf16", mode="peer")
        except RuntimeError as e:
            raised = "barrier" in str(e)
        np.savez(os.path.join(out_dir, f"late{rank}.npz"), healthy=healthy, loss=float(loss.item()),
                 grad_nan=bool(torch.isnan(text.grad).all().item() and torch.isnan(image.grad).all().item()), raised=raised)
    finally:
        os._exit(0)                                   # the transport is dead by design: no orderly teardown


def test_a_rank_that_misses_a_barrier_poisons_the_step(tmp_path):
    """a timed-out barrier must not yield a valid-looking loss or gradient.  Rank 1 stalls for longer
    than the barrier timeout; rank 0 must see NaN loss and gradients for that step and a RuntimeError at its next
    entry (host-side mirror of the failure flag, no device sync needed)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    mp.spawn(_late_worker, args=(2, _port(77), str(tmp_path)), nprocs=2, join=True)
    for r in range(2):                               # the late rank must not trust the step either
        got = np.load(tmp_path / f"late{r}.npz")
        assert np.isfinite(got["healthy"])
        assert np.isnan(got["loss"]) and bool(got["grad_nan"]) and bool(got["raised"]), (r, dict(got))


def _mpc_worker(rank, world, port, n_total, d, tau, precision, kind, out_dir, gather_all=True):
    sys.path.insert(0, ROOT)
    if not gather_all:                                # force the row-sharded form even for a small global batch
        os.environ["EVOKE_B200_GATHER_ALL_MAX"] = "0"
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from evoke_b200 import synth
        from evoke_b200.distributed import multi_pos_contra_images_sharded
        ids = _mpc_ids(n_total, kind)
        x = synth.make_embeddings(ids, d, seed=62)
        m = n_total // world
        sl = slice(rank * m, (rank + 1) * m)
        xs = torch.tensor(x[sl], device="cuda", requires_grad=True)
        loss = multi_pos_contra_images_sharded(xs, ids[sl].copy(), tau, precision=precision)
        if loss.grad_fn is not None:
            loss.backward()
        torch.cuda.synchronize()
        grad = xs.grad.cpu().numpy() if xs.grad is not None else np.zeros_like(x[sl])
        np.savez(os.path.join(out_dir, f"mpc{rank}.npz"), loss=loss.detach().float().cpu().numpy().reshape(-1), grad=grad)
    finally:
        dist.destroy_process_group()


def _mpc_ids(n_total, kind):
    from evoke_b200 import synth
    if kind == "mixed":
        return synth.make_study_ids(n_total, seed=61)
    ids = np.arange(n_total, dtype=np.int32)            # "lopsided": only rank 0's first rows have second views
    ids[: n_total // 8] = ids[: n_total // 8] // 2
    return ids


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("gather_all", [False, True], ids=["sharded", "gathered"])
@pytest.mark.parametrize("kind", ["mixed", "lopsided"])
@pytest.mark.parametrize("precision,ltol,gtol", [("fp32", 1e-5, 1e-4), ("bf16", 2e-5, 2e-2)])
def test_sharded_mpc_equals_oracle(tmp_path, precision, ltol, gtol, kind, gather_all, world):
    """multi_pos_contra_images_v0401 (:421-446) over the views of all ranks: cross-rank positives, rows dropped from
    queries and keys, a rank without kept rows; E strip + mask-free lists in bf16 mode on a rectangular row block."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    if gather_all and world > 2:
        pytest.skip("the gathered form (small global batches) is covered at world 2")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    n_total, d, tau = 640 * world, 256, 0.5
    port = _port(200 + world * 8 + (1 if precision == "fp32" else 0) + (2 if kind == "mixed" else 0) + (4 if gather_all else 0))
    mp.spawn(_mpc_worker, args=(world, port, n_total, d, tau, precision, kind, str(tmp_path), gather_all), nprocs=world, join=True)
    ids = _mpc_ids(n_total, kind)
    x = synth.make_embeddings(ids, d, seed=62)
    want, dx = orc.mpc_closed_form(x, ids, tau)
    m = n_total // world
    for r in range(world):
        got = np.load(tmp_path / f"mpc{r}.npz")
        assert abs(float(got["loss"][0]) - want) <= ltol * abs(want)
        assert np.abs(got["grad"] - dx[r * m:(r + 1) * m]).max() <= gtol * np.abs(dx).max()


def _small_g_worker(rank, world, port, n_total, d, tau, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from evoke_b200 import synth
        from evoke_b200.distributed import global_alignment_sharded
        ids = synth.make_study_ids(n_total, seed=71)
        n = n_total // world
        sl = slice(rank * n, (rank + 1) * n)
        image = torch.tensor(synth.make_embeddings(ids, d, seed=72)[sl], device="cuda", requires_grad=True)
        text = torch.tensor(synth.make_embeddings(ids, d, seed=73)[sl], device="cuda", requires_grad=True)
        loss = global_alignment_sharded(image, text, ids[sl].copy(), tau, precision="fp32")        # mode="auto"
        (2.0 * loss).backward()
        torch.cuda.synchronize()
        np.savez(os.path.join(out_dir, f"small{rank}.npz"), loss=loss.item(), d_image=image.grad.cpu().numpy(),
                 d_text=text.grad.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_small_global_batches_are_gathered_not_sharded(tmp_path):
    """32 pairs per rank (the reference's per-GPU batch): mode='auto' all-gathers once, evaluates the whole loss on
    every rank with the single-device (small-path) kernels and keeps its own gradient rows."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 64, 768, 0.5
    mp.spawn(_small_g_worker, args=(world, _port(321), n_total, d, tau, str(tmp_path)), nprocs=world, join=True)
    ids = synth.make_study_ids(n_total, seed=71)
    want, d_i, d_t, _ = orc.g_loss_closed_form(synth.make_embeddings(ids, d, seed=72), synth.make_embeddings(ids, d, seed=73), ids, tau)
    n = n_total // world
    for r in range(world):
        got = np.load(tmp_path / f"small{r}.npz")
        sl = slice(r * n, (r + 1) * n)
        assert abs(float(got["loss"]) - want) <= 1e-5 * abs(want)
        assert np.abs(got["d_image"] - 2.0 * d_i[sl]).max() <= 1e-4 * np.abs(d_i).max() * 2.0
        assert np.abs(got["d_text"] - 2.0 * d_t[sl]).max() <= 1e-4 * np.abs(d_t).max() * 2.0
