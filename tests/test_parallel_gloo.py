"""world_size-2 (and 3) gloo tests of the multi-GPU host logic: round-robin sharding and the single
all-gather of keyframe poses that precedes the pose graph."""
import importlib
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module(PKG + ".parallel")
    synth = importlib.import_module(PKG + ".synth")
    g = synth.pose_graph(0, n=n_total, n_loops=2)
    mine = par.shard_indices(n_total, rank, world)
    full = par.allgather_kf_poses(g["poses_gt"][mine], n_total)
    np.save(os.path.join(out_dir, f"full_{rank}.npy"), full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_total", [(2, 101), (3, 64), (2, 2)])
def test_allgather_reassembles_keyframe_order(tmp_path, world, n_total):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_total, str(tmp_path)), nprocs=world, join=True)
    synth = importlib.import_module(PKG + ".synth")
    want = synth.pose_graph(0, n=n_total, n_loops=2)["poses_gt"]
    for r in range(world):
        got = np.load(tmp_path / f"full_{r}.npy")
        assert got.shape == want.shape and np.array_equal(got, want), r


def test_sharding_is_a_partition():
    par = importlib.import_module(PKG + ".parallel")
    for n, world in [(4541, 8), (742, 4), (5, 8), (0, 2)]:
        owned = np.concatenate([par.shard_indices(n, r, world) for r in range(world)])
        assert np.array_equal(np.sort(owned), np.arange(n))
        assert max(len(par.shard_indices(n, r, world)) for r in range(world)) <= par.shard_capacity(n, world)


def test_single_process_path_needs_no_process_group():
    par = importlib.import_module(PKG + ".parallel")
    p = np.random.default_rng(0).normal(size=(9, 7))
    assert np.array_equal(par.allgather_kf_poses(p, 9, rank=0, world=1), p)


def _blob_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    par = importlib.import_module(PKG + ".parallel")
    mine = bytes([rank + 1]) * (0 if rank == 1 else 1000 * (rank + 1) + 7)       # ragged, one rank with nothing to send
    got = par.allgather_blobs(mine)
    ok = len(got) == world and all(got[r] == bytes([r + 1]) * (0 if r == 1 else 1000 * (r + 1) + 7) for r in range(world))
    open(os.path.join(out_dir, f"blob_{rank}.txt"), "w").write("ok" if ok else "bad")
    dist.barrier()
    dist.destroy_process_group()


def test_allgather_blobs_of_ragged_sizes(tmp_path):
    """The replay's keyframe records travel as one byte string per rank (parallel.allgather_blobs)."""
    port = _free_port()
    mp.spawn(_blob_worker, args=(3, port, str(tmp_path)), nprocs=3, join=True)
    for r in range(3):
        assert (tmp_path / f"blob_{r}.txt").read_text() == "ok", r
