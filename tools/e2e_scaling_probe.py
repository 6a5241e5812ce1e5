"""Which host resource bounds the end-to-end number at N GPUs?  (e2e efficiency 0.59 at N = 8.)

Run under torchrun with N ranks.  Every rank moves the e2e step's payload (63.8 MB host -> device, 18.4 MB device -> host,
pinned) in a loop, alone and together, with the process (a) left where the OS put it and (b) bound to the cores of the GPU's
NUMA node before the pinned buffers are allocated (first touch puts them on that node).  Rank 0 prints per-rank and
aggregate GB/s plus the box topology, as one JSON object.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/e2e_scaling_probe.py
"""
import glob
import json
import os
import subprocess
import time

import torch
import torch.distributed as dist

H2D_BYTES, D2H_BYTES, ITERS = 63_803_648, 18_425_856, 40


def numa_of_gpu(index):
    try:
        bus = torch.cuda.get_device_properties(index).pci_bus_id
        dom = torch.cuda.get_device_properties(index).pci_domain_id
        dev = torch.cuda.get_device_properties(index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        return int(open(path).read().strip())
    except Exception:
        return -1


def cpus_of_node(node):
    try:
        txt = open(f"/sys/devices/system/node/node{node}/cpulist").read().strip()
        out = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            out += list(range(int(a), int(b or a) + 1))
        return out
    except Exception:
        return []


def measure(local):
    dev = torch.device("cuda", local)
    src = torch.empty(H2D_BYTES, dtype=torch.uint8).pin_memory()
    src.fill_(1)                                    # touch: pages land on the node this thread runs on
    dst = torch.empty(D2H_BYTES, dtype=torch.uint8).pin_memory()
    dst.fill_(0)
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device=dev)
    d_out = torch.ones(D2H_BYTES, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}
    for mode in ("h2d", "d2h", "both"):
        torch.cuda.synchronize(dev)
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(ITERS):
            if mode in ("h2d", "both"):
                with torch.cuda.stream(s1):
                    d_in.copy_(src, non_blocking=True)
            if mode in ("d2h", "both"):
                with torch.cuda.stream(s2):
                    dst.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        nbytes = (H2D_BYTES if mode != "d2h" else 0) + (D2H_BYTES if mode != "h2d" else 0)
        res[mode] = ITERS * nbytes / dt / 1e9
    return res


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    node = numa_of_gpu(local)
    out = {"world": world, "payload_bytes": {"h2d": H2D_BYTES, "d2h": D2H_BYTES}}
    for label in ("unbound", "bound_to_gpu_numa_node"):
        if label != "unbound":
            cpus = cpus_of_node(node) if node >= 0 else []
            if cpus:
                os.sched_setaffinity(0, cpus)
        r = measure(local)
        t = torch.tensor([r["h2d"], r["d2h"], r["both"]], dtype=torch.float64, device="cuda")
        g = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(g, t)
        if rank == 0:
            per = [[round(float(x), 2) for x in gi] for gi in g]
            out[label] = {"per_rank_gbs_h2d_d2h_both": per, "aggregate_gbs": {"h2d": sum(p[0] for p in per), "d2h": sum(p[1] for p in per),
                                                                             "both": sum(p[2] for p in per)}}
    nodes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(nodes, torch.tensor([node], dtype=torch.int64, device="cuda"))
    if rank == 0:
        out["gpu_numa_node"] = [int(x) for x in nodes]
        out["numa_nodes"] = {os.path.basename(p): open(p + "/cpulist").read().strip() for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}
        out["host_cores"] = os.cpu_count()
        try:
            out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout[-3000:]
        except Exception as ex:
            out["topo"] = repr(ex)
        # what the e2e step needs per rank at the single-GPU rate (33 k frames/s): 64 frames per 1.9 ms
        out["needed_per_rank_gbs_at_1gpu_rate"] = {"h2d": H2D_BYTES / 1.93e-3 / 1e9, "d2h": D2H_BYTES / 1.93e-3 / 1e9}
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
