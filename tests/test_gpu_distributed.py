"""T4 (GPU, NCCL, world_size 2): sharded G loss == single-GPU G loss on the concatenated batch,
and both == the fp64 oracle.  Skipped on boxes with fewer than two GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_total, d, tau, precision, mode, out_dir):
    sys.path.insert(0, ROOT)
    if mode.startswith("peer-"):                     # "peer-fp32": fp32 partials on NVLink instead of bf16
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
        image = torch.tensor(xi[sl], device="cuda", requires_grad=True)
        text = torch.tensor(xt[sl], device="cuda", requires_grad=True)
        reps = 3 if mode == "peer" else 1             # the symmetric buffers and barrier epochs are reused step to step
        for _ in range(reps):
            image.grad = text.grad = None
            loss = global_alignment_sharded(image, text, ids[sl].copy(), tau, precision=precision, mode=mode)
            loss.backward()
        torch.cuda.synchronize()
        if mode == "peer":
            from evoke_b200 import peer
            for c in peer._CONTEXTS.values():
                if isinstance(c, peer.PeerContext):
                    c.check()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), loss=loss.item(), d_image=image.grad.cpu().numpy(),
                 d_text=text.grad.cpu().numpy())
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
            assert abs(float(got[f"loss{seed}"]) - want) <= 2e-3 * abs(want)
            assert np.abs(got[f"d_image{seed}"] - d_i[sl]).max() <= 2e-2 * np.abs(d_i).max()
            assert np.abs(got[f"d_text{seed}"] - d_t[sl]).max() <= 2e-2 * np.abs(d_t).max()


@pytest.mark.parametrize("mode", ["rs", "sym", "peer", "peer-fp32"])
@pytest.mark.parametrize("precision,ltol,gtol", [("fp32", 1e-5, 1e-4), ("bf16", 2e-3, 2e-2)])
def test_sharded_equals_oracle_on_two_gpus(tmp_path, precision, ltol, gtol, mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if mode.startswith("peer") and precision != "bf16":
        pytest.skip("the peer-memory path is a bf16-mode path")
    from evoke_b200 import synth
    from oracle import evoke_oracle as orc
    world, n_total, d, tau = 2, 1536, 256, 0.5
    port = 29900 + (os.getpid() % 90) + (1 if precision == "fp32" else 0) + (2 if mode == "sym" else 0) + (4 if mode == "peer" else 0) + (6 if mode == "peer-fp32" else 0)
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
