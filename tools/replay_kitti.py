"""BASELINE config 4 / 5 from the command line: the full-pipeline replay (package module replay.py) on 1 GPU, or on N GPUs
under torchrun (one process per GPU, frames and keyframes sharded round-robin).  Rank 0 prints one JSON summary and, with
--digests PATH, writes the per-frame digests + loop decisions + pose-graph digest that tests/test_gpu_multi.py compares
between 1 rank and N ranks.

    python tools/replay_kitti.py --frames 4541 --keyframes 742
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/replay_kitti.py
    python tools/replay_kitti.py --kitti /data/kitti/sequences/00      # image_0/ image_1/ times.txt (app/run_kitti_stereo.cpp:114-144)
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=4541)
    ap.add_argument("--keyframes", type=int, default=0, help="default: frames / kf-every")
    ap.add_argument("--kf-every", type=float, default=6.12)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--kitti", default=None, help="KITTI odometry sequence directory (image_0/, image_1/, times.txt)")
    ap.add_argument("--poses", default=None, help="KITTI ground-truth poses file (12 numbers per line) for the odometry edges")
    ap.add_argument("--digests", default=None)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    replay = importlib.import_module(PKG + ".replay")
    if args.kitti:
        seq = replay.KittiSequence(args.kitti, poses_file=args.poses, max_frames=args.frames, kf_every=args.kf_every)
    else:
        n_kf = args.keyframes or max(2, int(round(args.frames / args.kf_every)))
        seq = replay.Sequence(frames=args.frames, n_kf=n_kf)
    small = seq.n_kf < 200
    ops = replay.GpuOps(device=local, batch=args.batch, kf_batch=min(32, args.batch), n_kf=seq.n_kf)
    res = replay.run(seq, ops, rank=rank, world=world, db_min_size=5 if small else 50, min_gap=5 if small else 20,
                     with_digests=args.digests is not None, log=(print if args.verbose and rank == 0 else None))
    t = torch.tensor([res["timings"]["total_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["frames_per_s"] = seq.frames / float(t[0])
    if rank == 0:
        if args.digests:
            json.dump(res, open(args.digests, "w"))
        print(json.dumps({k: v for k, v in res.items() if k != "frame_digests"}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
