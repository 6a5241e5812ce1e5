"""Multi-GPU plumbing: one process per GPU; the single exchange step goes through the C ABI.

The per-frame path (extract, match, local BA in replay) shards by unit with no data-path collective: frame /
window i belongs to rank i mod world.  The only exchange is one all-gather of the keyframe poses (7 doubles per
keyframe, <= 42 KB for KITTI-00) before the pose graph, which every rank then solves redundantly (the solver is
deterministic, so all ranks hold identical results).

On GPUs the collective is libslamb200's `sb_allgather_kf_poses` (csrc/collective.cu: one ncclAllGather on the caller's
communicator) — the same entry point a C++ host calls; torch.distributed only supplies the communicator (its NCCL
process group's, or one bootstrapped with sb_nccl_unique_id / sb_nccl_comm_init when torch does not expose it).
With the gloo backend (CPU tests of the host logic) the same fixed-size records travel through
torch.distributed.all_gather_into_tensor."""
import ctypes as C

import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin ownership (BASELINE config 5): items rank, rank + world, ..."""
    return np.arange(rank, n_items, world)


def shard_capacity(n_items, world):
    return (n_items + world - 1) // world


def uninterleave(gathered, n_total):
    """gathered [world][cap][...] -> [n_total][...] in item order: item k lives at [k % world][k // world]."""
    world, cap = gathered.shape[:2]
    return np.ascontiguousarray(np.swapaxes(gathered, 0, 1).reshape((cap * world,) + gathered.shape[2:])[:n_total])


_OWN_COMM = {}


def nccl_comm(device_index):
    """ncclComm_t (as an int) for the default process group: torch's own communicator when it exposes it, else one
    bootstrapped through the C ABI (the id travels over torch.distributed's object broadcast)."""
    import torch
    import torch.distributed as dist
    from . import capi
    torch.cuda.set_device(device_index)
    try:
        backend = dist.group.WORLD._get_backend(torch.device("cuda", device_index))
        # NCCL communicators are created lazily: one tiny collective makes sure this one exists
        dist.all_reduce(torch.zeros(1, device=torch.device("cuda", device_index)))
        ptr = int(backend._comm_ptr())
        if ptr:
            return ptr
    except Exception:
        pass
    key = (dist.get_rank(), dist.get_world_size(), device_index)
    if key not in _OWN_COMM:
        ident = [None]
        if dist.get_rank() == 0:
            buf = (C.c_uint8 * 128)()
            capi._check(capi.lib().sb_nccl_unique_id(buf))
            ident = [bytes(buf)]
        dist.broadcast_object_list(ident, src=0)
        comm = C.c_void_p()
        buf = (C.c_uint8 * 128).from_buffer_copy(ident[0])
        capi._check(capi.lib().sb_nccl_comm_init(C.byref(comm), device_index, dist.get_world_size(), dist.get_rank(), buf))
        _OWN_COMM[key] = comm.value
    return _OWN_COMM[key]


def allgather_kf_poses_capi(local_poses, cap, comm, stream=0):
    """sb_allgather_kf_poses through ctypes: returns (all [world][cap][7], counts [world])."""
    from . import capi
    import torch.distributed as dist
    world = dist.get_world_size()
    local = np.zeros((cap, 7), np.float64)
    local[:len(local_poses)] = local_poses
    out = np.zeros((world, cap, 7), np.float64)
    counts = np.zeros(world, np.int32)
    capi._check(capi.lib().sb_allgather_kf_poses(C.c_void_p(comm), C.c_void_p(stream), capi._p(local), len(local_poses),
                                                 capi._p(out), capi._p(counts), cap))
    return out, counts


def allgather_kf_poses(local_poses, n_total, rank=None, world=None, device=None):
    """local_poses: [n_local, 7] float64 poses of the keyframes this rank owns (round-robin ownership, in
    ascending keyframe order).  Returns the full [n_total, 7] array in keyframe order on every rank.
    One fixed-size all-gather of padded records."""
    import torch
    import torch.distributed as dist
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    cap = shard_capacity(n_total, world)
    local = np.ascontiguousarray(local_poses, np.float64).reshape(-1, 7)
    n_local = len(shard_indices(n_total, rank, world))
    assert local.shape == (n_local, 7), (local.shape, n_local)
    if world == 1:
        return local.copy()
    if dist.get_backend() == "nccl":
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        gathered, counts = allgather_kf_poses_capi(local, cap, nccl_comm(dev.index if dev.index is not None else torch.cuda.current_device()))
        assert counts.tolist() == [len(shard_indices(n_total, r, world)) for r in range(world)], counts
        return uninterleave(gathered, n_total)
    buf = torch.zeros((cap, 7), dtype=torch.float64)
    buf[:n_local] = torch.from_numpy(local)
    flat = torch.empty((world * cap, 7), dtype=torch.float64)
    dist.all_gather_into_tensor(flat, buf)
    return uninterleave(flat.view(world, cap, 7).numpy(), n_total)


def allgather_blobs(blob, device=None):
    """All-gather of one variable-length byte string per rank (the replay's keyframe records: harness plumbing, not the hot
    path's collective): sizes first, then ONE padded all_gather_into_tensor — NCCL on `device`, gloo on the CPU.
    torch.distributed.all_gather_object does the same through per-object pickling and took 0.6 - 1.3 s for 2 x 45 MB."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    dev = torch.device("cpu") if device is None else torch.device(device)
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8) if len(blob) else torch.zeros(0, dtype=torch.uint8)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, torch.tensor([mine.numel()], dtype=torch.int64, device=dev))
    sizes = sizes.cpu().tolist()
    cap = max(1, max(sizes))
    buf = torch.zeros(cap, dtype=torch.uint8, device=dev)
    buf[:mine.numel()] = mine.to(dev)
    out = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, buf)
    out = out.cpu().numpy().reshape(world, cap)
    return [out[r, :sizes[r]].tobytes() for r in range(world)]
