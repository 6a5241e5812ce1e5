"""Multi-GPU plumbing: one process per GPU, torch.distributed for the single exchange step.

The per-frame path (extract, match, local BA in replay) shards by unit with no data-path collective: frame /
window i belongs to rank i mod world.  The only exchange is one all-gather of the keyframe poses (7 doubles per
keyframe, <= 42 KB for KITTI-00) before the pose graph, which every rank then solves redundantly (the solver is
deterministic, so all ranks hold identical results).  Works on NCCL (GPU tensors) and gloo (CPU tensors)."""
import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin ownership (BASELINE config 5): items rank, rank + world, ..."""
    return np.arange(rank, n_items, world)


def shard_capacity(n_items, world):
    return (n_items + world - 1) // world


def allgather_kf_poses(local_poses, n_total, rank=None, world=None, device=None):
    """local_poses: [n_local, 7] float64 poses of the keyframes this rank owns (round-robin ownership, in
    ascending keyframe order).  Returns the full [n_total, 7] array in keyframe order on every rank.
    One fixed-size all-gather of padded buffers; the counts follow from n_total, no second collective."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    cap = shard_capacity(n_total, world)
    local = torch.as_tensor(np.ascontiguousarray(local_poses, np.float64))
    n_local = len(shard_indices(n_total, rank, world))
    assert local.shape == (n_local, 7), (local.shape, n_local)
    buf = torch.zeros((cap, 7), dtype=torch.float64)
    buf[:n_local] = local
    if device is not None:
        buf = buf.to(device)
    if world == 1:
        gathered = buf[None]
    else:
        flat = torch.empty((world * cap, 7), dtype=torch.float64, device=buf.device)
        dist.all_gather_into_tensor(flat, buf)
        gathered = flat.view(world, cap, 7)
    # un-interleave: keyframe k lives at [k % world][k // world]
    full = gathered.permute(1, 0, 2).reshape(cap * world, 7)[:n_total]
    return full.cpu().numpy()
