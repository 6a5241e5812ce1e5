#!/usr/bin/env python
"""Benchmark of the B200 hot path: KITTI-shaped stereo frames/s through ORB extract (both views,
2000 features, 8-level pyramid) + left<->right Hamming match + one sliding-window local BA per frame
(BASELINE.json configs 2 + 3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--pairs B]

One "step" = one batch of B synthetic stereo pairs (1241x376 u8) through extract + match, plus B local-BA
windows (7 keyframes x 300 landmarks, ~1400 observations) — one window per frame, although the live system
only optimises once per keyframe (~1 frame in 6).
  value : frames/s with the inputs already resident in HBM (device entry points, CUDA events)
  e2e   : frames/s through the host-pointer C ABI (pinned host buffers, H2D + D2H inside the timed region)
  --impl reference : the reference's CPU path (its own ORBextractor.cpp from oracle/_ref + the C restatement of the OpenCV /
                     g2o parts, all host cores) on the same workload
Prints ONE JSON line on rank 0.

Scheduling defaults (each env switch reproduces the A/B it was measured with; see the comment where it is read):
  device-resident loop : 1 extract + match stream beside BA, 2 without BA (BENCH_EXT_HANDLES); 2 BA batches in flight
                         (BENCH_BA_STREAMS), BA batch i released at the start of extraction i (BENCH_BA_LOCKSTEP),
                         BA stream priority 0 (BENCH_BA_PRIORITY)
  end-to-end loop      : 2 front-end handles beside BA, 3 without (BENCH_FE_HANDLES), their kernels on one shared stream beside BA
                         (BENCH_FE_SHARED_STREAM), 2 prioritised BA streams (BENCH_BA_E2E_STREAMS, BENCH_BA_E2E_PRIORITY)
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"

W, H = 1241, 376
ORB_PARAMS = (2000, 1.2, 8, 20, 7)
KITTI_K = (718.856, 718.856, 607.1928, 185.2157)
METRIC = "KITTI stereo frames/s (extract+match+local-BA) @1/2/4/8 B200; % HBM roofline"
WORKLOAD = ("config 3: ORB extract (both views, 2000 feats) + L<->R Hamming match + sliding-window local BA "
            "(7 KFs x 300 landmarks, one window per frame), 1241x376 stereo, synthetic replay")
STAGES = ["copy_level0", "resize_pyramid", "fast_cells", "quadtree", "gauss_blur", "describe"]
BA_CAPS = dict(max_poses=7, max_points=320, max_obs=2304)


CALC_MACS = 64 * 62 * 82 * 25 + 128 * 32 * 42 * 1024 + 4 * 14 * 19 * 1152   # multiply-adds of the three convolutions per image


# dram__bytes_read.sum + dram__bytes_write.sum per launch come from the committed raw page of the `ncu --set full` capture of
# this same command (profiles/r<round>_ncu_full*.csv, newest round first) — never from numbers typed into this file.
KERNEL_STAGE = {"k_copy_level0": "copy_level0", "k_resize_quads": "resize_pyramid", "k_resize": "resize_pyramid",
                "k_pyr_levels": "resize_pyramid", "k_fast_cells": "fast_cells", "k_quadtree": "quadtree", "k_blur": "gauss_blur",
                "k_describe": "describe", "k_orient": "describe", "k_brief": "describe", "k_expand": "hamming_match",
                "k_hamming_umma": "hamming_match", "k_hamming_decode": "hamming_match", "k_ba_solve": "local_ba"}
_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def ncu_traffic():
    """-> ({stage: DRAM bytes of one step's launches of that stage}, {stage: {metric: value}}, source path) or ({}, {}, None)."""
    import csv
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full*.csv")),
                   key=lambda f: (int(re.search(r"r(\d+)_", os.path.basename(f)).group(1)), os.path.getmtime(f)))
    if not files:
        return {}, {}, None
    rows = list(csv.reader(open(files[-1])))
    head, units = rows[0], rows[1]
    col = {c: i for i, c in enumerate(head)}
    need = ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum")
    if any(c not in col for c in need):
        return {}, {}, None
    traffic, extra, seen = {}, {}, {}
    for r in rows[2:]:
        if len(r) <= col["dram__bytes_write.sum"]:
            continue
        name = r[col["Kernel Name"]].replace("void ", "").split("(")[0].split("<")[0]
        stage = KERNEL_STAGE.get(name)
        if stage is None:
            continue
        # one step of the capture = the first launch of every distinct (kernel, grid) pair; repeated steps are not added up
        key = (name, r[col["Grid Size"]] if "Grid Size" in col else "")
        if key in seen and name not in ("k_expand",):
            continue
        if name == "k_expand" and seen.get(key, 0) >= 2:
            continue
        seen[key] = seen.get(key, 0) + 1
        b = sum(float(r[col[c]]) * _UNIT.get(units[col[c]], 1.0) for c in need[1:])
        traffic[stage] = traffic.get(stage, 0.0) + b
        m = extra.setdefault(stage, {})
        for c, k in (("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
                     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
                     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
                     ("launch__registers_per_thread", "registers")):
            if c in col and r[col[c]] not in ("", "n/a"):
                try:
                    m[k] = max(m.get(k, 0.0), float(r[col[c]]))
                except ValueError:
                    pass
    return traffic, extra, os.path.relpath(files[-1], ROOT)


def pyramid_bytes():
    """Bytes of the 8 pyramid levels of one image."""
    inv = [np.float32(1.0)]
    s = np.float32(1.0)
    for _ in range(1, 8):
        s = np.float32(np.float64(s) * np.float64(np.float32(1.2)))
        inv.append(np.float32(1.0) / s)
    sizes = [(W, H)] + [(int(np.rint(np.float32(W) * i)), int(np.rint(np.float32(H) * i))) for i in inv[1:]]
    return [w * h for w, h in sizes]


def algorithmic_bytes(stage, n_images, n_kps, n_cands):
    """Compulsory HBM bytes of one launch group of `stage` over n_images images."""
    lv = pyramid_bytes()
    if stage == "copy_level0":
        return n_images * 2 * lv[0]
    if stage == "resize_pyramid":
        return n_images * sum(lv[l - 1] + lv[l] for l in range(1, 8))
    if stage == "fast_cells":
        return n_images * sum(lv) + 4 * n_cands
    if stage == "quadtree":
        return 4 * n_cands + 4 * n_kps
    if stage == "gauss_blur":
        return n_images * 2 * sum(lv)
    if stage == "describe":
        return n_kps * (749 + 512 + 28 + 32 + 4)
    if stage == "hamming_match":
        return n_kps * 32 + (n_kps // 2) * 8
    if stage == "local_ba":
        return 0   # filled by the caller: observations + points + poses of the batch
    raise KeyError(stage)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class CpuReference:
    """The CPU restatement of the reference path (oracle/: ORBextractor::DetectAndCompute on both views +
    BFMatcher Hamming), one extractor per worker thread (the C code releases the GIL)."""

    def __init__(self, cores):
        from concurrent.futures import ThreadPoolExecutor
        from oracle import oracle as O
        from oracle import ref as R
        O.build()
        self.O, self.cores = O, cores
        # the extractor is the reference's OWN source when oracle/_ref is there (src/ORBextractor.cpp compiled unmodified; its
        # OpenCV calls land in the same C primitives the port uses), else the restatement
        self.R = None
        try:
            if R.available():
                R.lib()
                self.R = R
        except Exception:
            self.R = None
        self.kind = "reference" if self.R else "port"
        self.kind_detail = ("extractor: the reference's own src/ORBextractor.cpp compiled unmodified (oracle/_ref); the OpenCV primitives it "
                            "calls, BFMatcher and the g2o LM / Schur BA: restated in C (oracle/)") if self.R else \
            "oracle/: C restatement of the reference's OpenCV / g2o based path"
        self.local = threading.local()
        self.pool = ThreadPoolExecutor(cores)

    def _one(self, job):
        pair, win = job
        if not hasattr(self.local, "ext"):
            self.local.ext = (self.R or self.O).ORBextractor(*ORB_PARAMS)
        _, dl = self.local.ext.DetectAndCompute(pair[0])
        _, dr = self.local.ext.DetectAndCompute(pair[1])
        idx, _ = self.O.hamming_match(dl, dr)
        if win is not None:   # Backend::OptimizeActiveMap's solve on one window
            self.O.ba_solve(win["poses0"], win["points0"], win["fixed"], win["obs_pose"], win["obs_point"], win["uv"], KITTI_K)
        return len(idx)

    def run(self, frames, windows=None):
        """frames [n, 2, H, W] (+ one BA window per frame) -> seconds"""
        jobs = [(frames[i], windows[i % len(windows)] if windows else None) for i in range(len(frames))]
        t0 = time.perf_counter()
        list(self.pool.map(self._one, jobs))
        return time.perf_counter() - t0


class Cv2Primitives:
    """the OpenCV primitives of the extract + match path through the in-container cv2 4.13.0 (SIMD / IPP
    builds of what the reference links): the chained cv::resize pyramid, one cv::FAST call per 30-px grid cell per level
    (src/ORBextractor.cpp:838-865, with the 20 -> 7 fallback), GaussianBlur 7x7 sigma 2 per level, BFMatcher(HAMMING).match.
    The reference's own C++ around them (quadtree, orientation, rBRIEF sampling) is NOT included, so this is a LOWER bound on
    the reference's CPU time per frame.  Each primitive is bit-compared with the port by tests/test_oracle_cv2.py."""

    def __init__(self, sizes, desc_pair):
        import cv2
        self.cv2, self.sizes, self.desc_pair = cv2, sizes, desc_pair
        self.local = threading.local()

    def _tools(self):
        if not hasattr(self.local, "f20"):
            self.local.f20 = self.cv2.FastFeatureDetector_create(20, True)
            self.local.f7 = self.cv2.FastFeatureDetector_create(7, True)
            self.local.bf = self.cv2.BFMatcher(self.cv2.NORM_HAMMING)
        return self.local.f20, self.local.f7, self.local.bf

    def image(self, img):
        cv2 = self.cv2
        f20, f7, _ = self._tools()
        levels = [img]
        for (w, h) in self.sizes[1:]:
            levels.append(cv2.resize(levels[-1], (w, h), interpolation=cv2.INTER_LINEAR))
        n = 0
        for lv in levels:
            h, w = lv.shape
            minB, maxBX, maxBY = 16, w - 16, h - 16
            width, height = float(maxBX - minB), float(maxBY - minB)
            nCols, nRows = int(width / 30), int(height / 30)
            wCell, hCell = int(np.ceil(width / nCols)), int(np.ceil(height / nRows))
            for i in range(nRows):
                iniY = minB + i * hCell
                maxY = min(iniY + hCell + 6, maxBY)
                if iniY >= maxBY - 3:
                    continue
                for j in range(nCols):
                    iniX = minB + j * wCell
                    maxX = min(iniX + wCell + 6, maxBX)
                    if iniX >= maxBX - 6:
                        continue
                    roi = lv[iniY:maxY, iniX:maxX]
                    k = f20.detect(roi)
                    if not k:
                        k = f7.detect(roi)
                    n += len(k)
            cv2.GaussianBlur(lv, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        return n

    def frame(self, pair):
        n = self.image(pair[0]) + self.image(pair[1])
        self._tools()[2].match(self.desc_pair[0], self.desc_pair[1])
        return n

    def fps(self, frames, threads):
        from concurrent.futures import ThreadPoolExecutor
        self.cv2.setNumThreads(1 if threads > 1 else 1)   # one frame per thread; cv2's own pool stays out of the way
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(self.frame, [frames[i] for i in range(min(len(frames), threads))]))   # warm
            t0 = time.perf_counter()
            list(pool.map(self.frame, [frames[i] for i in range(len(frames))]))
            return len(frames) / (time.perf_counter() - t0)


def cpu_reference_fps(frames, windows, cores):
    ref = CpuReference(cores)
    ref.run(frames[:cores], windows)   # warm: library load, per-thread extractors
    return len(frames) / ref.run(frames, windows), ref.kind, ref.kind_detail


_JSON_FD = None


def claim_stdout():
    """stdout must carry exactly one JSON line.  Native libraries (NCCL prints its version from C) write to fd 1 too, so fd 1
    is pointed at stderr for the whole run and the JSON line goes to a private copy of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(out):
    line = (json.dumps(out) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, line)


def run_reference(args, rank, world):
    if rank != 0:
        return
    synth = importlib.import_module(PKG + ".synth")
    cores = os.cpu_count() or 1
    per_step = max(cores, 8)
    frames = synth.stereo_batch(0, per_step)
    windows = None if args.no_ba else [synth.ba_window(s) for s in range(per_step)]
    ref = CpuReference(cores)
    for _ in range(args.warmup):
        ref.run(frames, windows)
    total = sum(ref.run(frames, windows) for _ in range(args.steps))
    value = per_step * args.steps / total
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": {"workload": WORKLOAD if not args.no_ba else "config 2: ORB extract (both views) + L<->R Hamming match only",
                      "frames_per_step": per_step},
           "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": ref.kind, "kind_detail": ref.kind_detail,
                            "sample": f"{per_step} synthetic stereo frames per step, one frame per thread: " + ref.kind_detail},
           "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def run_config4(args, rank, world, local_rank):
    """BASELINE config 4 (and 5 under torchrun): the full-pipeline replay of the package's replay.py — every frame through
    extract + match, one BA window and the loop-closing front end per keyframe, DeepLCD detection, geometric verification
    and pose-graph optimisation in keyframe order — host buffers in, host results out (this IS the end-to-end number)."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    replay = importlib.import_module(PKG + ".replay")
    n_kf = max(2, int(round(args.frames / 6.12)))
    small = n_kf < 200
    ops = replay.GpuOps(device=local_rank, batch=args.pairs, kf_batch=min(32, args.pairs), n_kf=n_kf)
    warm = replay.Sequence(frames=3 * args.pairs, n_kf=min(32, args.pairs))  # warm-up: three batches through every operator, a full keyframe batch
    for _ in range(3):
        replay.run(warm, ops, rank=0, world=1, db_min_size=5, min_gap=5)
        ops.reset()
    if world > 1:   # ... and once through the exchange step: NCCL creates its communicator at the first collective (0.6 - 1.7 s)
        par_ = importlib.import_module(PKG + ".parallel")
        par_.allgather_kf_poses(np.zeros((1, 7)), world, rank, world, device=torch.device("cuda", local_rank))
        par_.allgather_blobs(b"warm", device=torch.device("cuda", local_rank))
        torch.cuda.synchronize()
        dist.barrier()
    seq = replay.Sequence(frames=args.frames, n_kf=n_kf)
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = replay.run(seq, ops, rank=rank, world=world, db_min_size=5 if small else 50, min_gap=5 if small else 20)
    clocks = sampler.stop()
    t = torch.tensor([res["timings"][k] for k in ("total_s", "frames_s", "keyframes_s", "exchange_s", "loop_closing_s", "posegraph_s")],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total, fr, kf, ex, lc, pg = (float(x) for x in t)
    if rank == 0:
        value = seq.frames / total
        H2D = seq.frames * 2 * H * W + seq.n_kf * 2 * H * W
        out = {"metric": "KITTI stereo frames/s, full pipeline (extract+match+local BA+DeepLCD loop detect/verify+pose graph)",
               "value": value, "unit": "frames/s", "n_gpus": world, "steps": 1, "warmup": 3, "ms_per_step": total * 1e3,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8 / f32 (CNN) / f64 (BA, pose graph)",
               "data": "synthetic", "config": {"workload": f"config {4 if world == 1 else 5}: full pipeline replay, {seq.frames} stereo frames 1241x376, "
                                                         f"{seq.n_kf} keyframes, {len(seq.loop_pairs)} planted revisits, 2000 feats/frame; one step = the whole sequence",
                                               "sharding": "frames and keyframes round-robin by rank; keyframe poses all-gathered through sb_allgather_kf_poses, "
                                                           "loop closing redundantly on every rank" if world > 1 else "single GPU"},
               "clocks": clocks,
               "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": H2D, "d2h_bytes_per_step": None,
                       "note": "the replay runs on host buffers through the host-pointer C ABI: value IS end to end"},
               "stage_seconds": {"extract_match_ba": fr, "keyframe_front_end_and_descriptors": kf, "exchange": ex,
                                 "loop_closing_total": lc, "of_which_pose_graph": pg},
               "keyframe_stage_seconds_rank0": {k[3:-2]: round(v, 4) for k, v in res["timings"].items() if k.startswith("kf_") and k.endswith("_s")},
               "keyframe_detect_batches_ms_rank0": res["timings"].get("kf_detect_batches_ms"),
               "loops_closed": len(res["loops"]), "loops_planted": len(seq.loop_pairs), "posegraph_runs": res["posegraph_runs"],
               "mean_position_error_m": {"dead_reckoned": res["mean_position_error_dead_reckoned_m"], "final": res["mean_position_error_final_m"]},
               "keypoints": res["keypoints"], "matches": res["matches"]}
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500, help="default: >= 1 s of timed region at ~2 ms per step")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4],
                    help="BASELINE config: 2 = extract + match, 3 = + local BA (the metric's config, default), "
                         "4 = full pipeline replay incl. DeepLCD loop detection + pose graph over 4541 frames")
    ap.add_argument("--frames", type=int, default=4541, help="config 4: frames of the replayed sequence")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=64, help="stereo pairs (and BA windows) per step")
    ap.add_argument("--pool", type=int, default=192, help="distinct resident stereo pairs (> L2 in total)")
    ap.add_argument("--no-ba", action="store_true", help="config 2 only: extract + match")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.config == 2:
        args.no_ba = True
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if args.config == 4:
        run_config4(args, rank, world, local_rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    lib = pkg.lib()
    B, P = args.pairs, max(args.pool, args.pairs)
    P = (P // B) * B
    with_ba = not args.no_ba

    # ---- inputs: a pool of distinct frames, larger than L2 (126 MB) in total, resident in HBM
    uniq = min(P, 48)
    base = synth.stereo_batch(1000 * rank, uniq)
    pool_np = np.empty((P, 2, H, W), np.uint8)
    for i in range(P):   # horizontal shifts of the unique frames: distinct bytes, same statistics
        pool_np[i] = np.roll(base[i % uniq], 7 * (i // uniq), axis=-1)
    pool = torch.from_numpy(pool_np).cuda()
    n_uniq_w = 16
    windows = [synth.ba_window(100 * rank + s) for s in range(n_uniq_w)]
    windows = [windows[i % n_uniq_w] for i in range(B)]

    # Extractor + matcher on one stream, BA on a second one.  Without BA (config 2) two handle pairs alternate on two streams: the
    # latency-bound stretches of one batch (quadtree, the small pyramid levels, kernel tails) are filled by the other batch's
    # kernels (57.6 k -> 60.9 k frames/s; the timed region carries no per-stage events, the serialised pass gives the stage times).
    # Beside BA a second handle pair changes nothing (42.9 k either way): BA's CTAs already fill those stretches.
    NH = int(os.environ.get("BENCH_EXT_HANDLES", "1" if with_ba else "2"))
    sx = [torch.cuda.Stream() for _ in range(NH)]
    # (BENCH_BA_PRIORITY=-1 gives the BA stream the higher priority: measured 39.0 k -> 36.0 k frames/s — a window's CTA needs a
    #  whole SM, and holding back the extractor's CTAs until 64 SMs have drained costs more SM-time than it saves)
    s2 = torch.cuda.Stream(priority=int(os.environ.get("BENCH_BA_PRIORITY", "0")))
    exts = [pkg.ORBextractor(*ORB_PARAMS, max_w=W, max_h=H, max_batch=2 * B, device=local_rank) for _ in range(NH)]
    mats = [pkg.HammingMatcher(max_batch=B, max_rows=exts[0].cap, device=local_rank) for _ in range(NH)]
    ba = pkg.LocalBA(max_windows=B, device=local_rank, **BA_CAPS)
    NBA = int(os.environ.get("BENCH_BA_STREAMS", "2"))   # back-end batches in flight (each on its own stream and buffers)
    for k in range(NH):
        exts[k].set_stream(sx[k].cuda_stream)
        mats[k].set_stream(sx[k].cuda_stream)
    ba.set_stream(s2.cuda_stream)
    ext = exts[0]
    cap = ext.cap
    kps = [torch.zeros((2 * B, cap, 28), dtype=torch.uint8, device="cuda") for _ in range(NH)]
    desc = [torch.zeros((2 * B, cap, 32), dtype=torch.uint8, device="cuda") for _ in range(NH)]
    counts = [torch.zeros(2 * B, dtype=torch.int32, device="cuda") for _ in range(NH)]
    midx = [torch.full((B, cap), -1, dtype=torch.int32, device="cuda") for _ in range(NH)]
    mdist = [torch.full((B, cap), -1, dtype=torch.int32, device="cuda") for _ in range(NH)]
    bh = ba.pack(windows)                                  # host batch (padded slots)
    bd = {k: torch.from_numpy(v).cuda() for k, v in bh.items()}
    bd0 = {"poses": bd["poses"].clone(), "points": bd["points"].clone()}
    bd["chi2"] = torch.zeros((B, BA_CAPS["max_obs"]), dtype=torch.float64, device="cuda")
    bd["outlier"] = torch.zeros((B, BA_CAPS["max_obs"]), dtype=torch.uint8, device="cuda")
    bd["info"] = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    ba_set = [(ba, bd, s2)]
    for _ in range(1, NBA if with_ba else 1):
        b_ = pkg.LocalBA(max_windows=B, device=local_rank, **BA_CAPS)
        s_ = torch.cuda.Stream()
        b_.set_stream(s_.cuda_stream)
        ba_set.append((b_, {k: v.clone() for k, v in bd.items()}, s_))
    ba_bytes = int(bh["ne"].sum()) * (8 + 16 + 8 + 1) + int(bh["nl"].sum()) * (48 + 1) + int(bh["np"].sum()) * 112

    # BA batch i starts with the extraction of step i: the BA stream waits for an event recorded on the extract stream at the start
    # of the step.  Left alone, the BA stream runs steps ahead and overlaps with itself (128 SMs of BA, the extractor's latency-bound
    # kernels squeezed onto 20); in lockstep its 64 whole-SM CTAs are placed while the short pyramid kernels run (40.7 k -> 43.0 k frames/s)
    lockstep = os.environ.get("BENCH_BA_LOCKSTEP", "1") == "1"
    ev_step = [torch.cuda.Event() for _ in range(4)]

    def step_dev(i, ev=None):
        off = (i * B) % P
        k = i % NH
        if with_ba:
            ba_, bd_, s_ = ba_set[i % len(ba_set)]
            if lockstep:   # the BA batch of step i starts with the extraction of step i (not steps ahead of it)
                ev_step[i % 4].record(sx[k])
                s_.wait_event(ev_step[i % 4])
            with torch.cuda.stream(s_):
                bd_["poses"].copy_(bd0["poses"], non_blocking=True)      # every step starts from the same windows
                bd_["points"].copy_(bd0["points"], non_blocking=True)
                if ev:
                    ev[2].record(s_)
                ba_.solve_dev(B, bd_, KITTI_K)
                if ev:
                    ev[3].record(s_)
        exts[k].detect_and_compute_dev(2 * B, pool[off], H * W, W, H, W, kps[k], desc[k], counts[k], cap)
        if ev:
            ev[0].record(sx[k])
        mats[k].match_dev(B, desc[k], 2 * cap * 32, counts[k], 2, desc[k][0, :, :].data_ptr() + cap * 32, 2 * cap * 32,
                          counts[k].data_ptr() + 4, 2, cap, midx[k], mdist[k], cap)
        if ev:
            ev[1].record(sx[k])

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(max(args.warmup, NH)):
        step_dev(i)
    barrier()
    for e in exts:
        e.sync_status()
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    # the timed region holds the kernels only: no per-stage events (an event record between two kernels also switches off
    # their programmatic dependent launch); the stage times come from two short passes afterwards
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    e0.record(main)
    for st in sx + [b[2] for b in ba_set]:
        st.wait_stream(main)
    for i in range(args.steps):
        step_dev(args.warmup + i)
    for st in sx + [b[2] for b in ba_set]:
        main.wait_stream(st)
    e1.record(main)
    barrier()
    clocks = sampler.stop()
    for e in exts:
        e.sync_status()
    dev_ms = e0.elapsed_time(e1)
    # ---- concurrent stage pass: the same loop with the per-stage events on (both streams running, like the timed region)
    CONC_STEPS = min(args.steps, 40)
    for e in exts:
        lib.sb_orb_profile(e._h, 1)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(CONC_STEPS)]
    for i in range(CONC_STEPS):
        step_dev(args.warmup + args.steps + i, evs[i])
    barrier()
    ms = np.zeros(6, np.float32)
    launches = np.zeros(6, np.int32)
    for e in exts:
        ms_k = np.zeros(6, np.float32)
        la_k = np.zeros(6, np.int32)
        lib.sb_orb_profile_read(e._h, C.c_void_p(ms_k.ctypes.data), C.c_void_p(la_k.ctypes.data), 6)
        lib.sb_orb_profile(e._h, 0)
        ms += ms_k
        launches += la_k
    match_ms = float(sum(e[0].elapsed_time(e[1]) for e in evs))
    ba_ms = float(sum(e[2].elapsed_time(e[3]) for e in evs)) if with_ba else 0.0
    counts, mdist = counts[(args.warmup + args.steps - 1) % NH], mdist[(args.warmup + args.steps - 1) % NH]
    ext = exts[(args.warmup + args.steps - 1) % NH]
    n_kps = int(counts.sum().item())
    sample_imgs = list(range(0, 2 * B, max(1, 2 * B // 4)))
    n_cands = sum(len(ext.debug_candidates(b, level)) for b in sample_imgs for level in range(8))
    n_cands = int(n_cands * (2 * B) / len(sample_imgs))
    n_matched = int((mdist >= 0).sum().item())
    ba_info = bd["info"].cpu().numpy() if with_ba else np.zeros((B, 4), np.int32)

    # ---- serialised stage pass (after the timed region): the extract/match stream alone with the per-stage events on, then
    #      the BA stream alone — stage times without interference, reproducible run to run; used for the roofline entries.
    SER_STEPS = 6
    barrier()
    saved_ba, with_ba = with_ba, False
    for e in exts:
        lib.sb_orb_profile(e._h, 1)
    sev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(SER_STEPS)]
    for i in range(SER_STEPS):
        step_dev(args.warmup + args.steps + CONC_STEPS + i, sev[i])
        torch.cuda.synchronize()
    ser_ms = {}
    ms_s = np.zeros(6, np.float32)
    la_s = np.zeros(6, np.int32)
    for e in exts:
        ms_k = np.zeros(6, np.float32)
        la_k = np.zeros(6, np.int32)
        lib.sb_orb_profile_read(e._h, C.c_void_p(ms_k.ctypes.data), C.c_void_p(la_k.ctypes.data), 6)
        lib.sb_orb_profile(e._h, 0)
        ms_s += ms_k
    for i, sname in enumerate(STAGES):
        ser_ms[sname] = float(ms_s[i])
    ser_ms["hamming_match"] = float(sum(e[0].elapsed_time(e[1]) for e in sev))
    with_ba = saved_ba
    if with_ba:
        tot = 0.0
        for i in range(SER_STEPS):
            with torch.cuda.stream(s2):
                bd["poses"].copy_(bd0["poses"], non_blocking=True)
                bd["points"].copy_(bd0["points"], non_blocking=True)
                sev[i][2].record(s2)
                ba.solve_dev(B, bd, KITTI_K)
                sev[i][3].record(s2)
            torch.cuda.synchronize()
            tot += sev[i][2].elapsed_time(sev[i][3])
        ser_ms["local_ba"] = tot
    # algorithmic flops of one BA launch: per LM iteration every edge is linearised and accumulated
    # (~450 flops), per trial every pair of a landmark's edges costs one 6x3 x 3x6 product (216) + Hpl Dinv per edge (54) +
    # the landmark update (36 per edge) + the (6 P)^3 / 3 Cholesky
    ba_flops = 0.0
    if with_ba:
        for k in range(B):
            ne_k, np_k = int(bh["ne"][k]), int(bh["np"][k])
            per_lm = np.bincount(bh["ol"][k, :ne_k], minlength=int(bh["nl"][k]))
            free = bh["fixed"][k, :len(per_lm)] == 0
            pairs_k = float((per_lm[free] * (per_lm[free] + 1) / 2).sum())
            iters_k = float(ba_info[k, 1])
            trials_k = iters_k * 1.3                     # measured mean of LM trials per iteration on these windows
            ba_flops += iters_k * ne_k * 450 + trials_k * (pairs_k * 216 + ne_k * 90 + (6 * np_k) ** 3 / 3)

    # ---- e2e: the host-pointer C ABI with pinned host buffers (H2D + D2H inside the timed region):
    #      sb_stereo handles used in turn (copy in, kernels and copy out of consecutive batches overlap)
    #      and sb_ba_submit / sb_ba_wait on pinned host arrays.
    del ext, exts, mats
    hp = min(P, 4 * B)
    host_pool = torch.from_numpy(pool_np[:hp]).pin_memory()
    host_np = host_pool.numpy()
    # front-end handles in flight: copy in, kernels and copy out of consecutive batches overlap.  Beside BA two are enough and better
    # than three (a third batch's 60 MB copy in sits in the copy engine's queue ahead of the next BA batch's 4 MB)
    NE = int(os.environ.get("BENCH_FE_HANDLES", "2" if with_ba else "3"))
    fes = [pkg.StereoFrontend(*ORB_PARAMS, max_w=W, max_h=H, max_pairs=B, device=local_rank) for _ in range(NE)]
    outs = [fe.alloc_outputs(B, pinned=True) for fe in fes]
    # with the back end beside it, the front-end handles put their kernels on ONE stream (copies stay on their own): batches follow one
    # another like a single-stream loop and BA's whole-SM CTAs find free SMs at every kernel boundary (measured e2e 35.7 k -> 37.6 k
    # frames/s); without BA the kernels of neighbouring batches may as well overlap (48.3 k vs 46.7 k)
    fe_shared = os.environ.get("BENCH_FE_SHARED_STREAM", "1" if with_ba else "0") == "1"
    if fe_shared:
        s_fe = torch.cuda.Stream()
        for fe in fes:
            fe.set_compute_stream(s_fe.cuda_stream)
    # two back-end handles used alternately: a batch of windows is submitted without waiting for the previous one
    bas = [ba, pkg.LocalBA(max_windows=B, device=local_rank, **BA_CAPS)] if with_ba else []
    # the two back-end handles on their own streams (the copies of one batch overlap the kernel of the other), with the higher
    # priority (end to end the BA batch is what the host waits for; measured 36.0 k -> 38.4 k frames/s with the shared front-end stream)
    e2e_prio = int(os.environ.get("BENCH_BA_E2E_PRIORITY", "-1"))
    s2e = torch.cuda.Stream(priority=e2e_prio)
    e2e_ba_streams = [s2e, torch.cuda.Stream(priority=e2e_prio)] if os.environ.get("BENCH_BA_E2E_STREAMS", "2") == "2" else [s2e, s2e]
    for b_, st_ in zip(bas, e2e_ba_streams):
        b_.set_stream(st_.cuda_stream)
    hbs = [{k: torch.from_numpy(v).pin_memory() for k, v in bh.items()} for _ in range(2)]
    hb = hbs[0]
    hb0 = {"poses": hb["poses"].clone(), "points": hb["points"].clone()}
    h_outs = [(torch.zeros((B, BA_CAPS["max_obs"]), dtype=torch.float64).pin_memory(),
               torch.zeros((B, BA_CAPS["max_obs"]), dtype=torch.uint8).pin_memory(),
               torch.zeros((B, 4), dtype=torch.int32).pin_memory()) for _ in range(2)]
    h_chi2, h_outl, h_info = h_outs[0]
    Kd = np.ascontiguousarray(KITTI_K, np.float64)
    ext7 = np.array([0, 0, 0, 1, 0, 0, 0], np.float64)

    ba_pending = [False, False]

    def ba_wait(k):
        if ba_pending[k]:
            assert lib.sb_ba_wait(bas[k]._h) == 0, pkg.last_error()
            ba_pending[k] = False

    # (the same lockstep between a BA batch and the front end's latest batch, through an event on the shared stream, was measured
    #  here too: 41.7 k -> 39.0 k frames/s — the host already paces this loop; BENCH_E2E_LOCKSTEP=1 reproduces it)
    e2e_lockstep = fe_shared and os.environ.get("BENCH_E2E_LOCKSTEP", "0") == "1"
    ev_fe = [torch.cuda.Event() for _ in range(4)]
    fe_marks = []                                   # events recorded on the shared front-end stream where a batch's kernels begin

    def ba_submit(k):
        ba_wait(k)                                  # this handle's previous batch (two steps ago): its results sit in the pinned buffers
        if e2e_lockstep and fe_marks:               # start with the front end's latest batch, as in the device-resident loop
            e2e_ba_streams[k].wait_event(fe_marks[-1])
        hbk, (c2, ol_, inf) = hbs[k], h_outs[k]
        hbk["poses"].copy_(hb0["poses"])
        hbk["points"].copy_(hb0["points"])
        rc = lib.sb_ba_submit(bas[k]._h, B, C.c_void_p(hbk["np"].data_ptr()), C.c_void_p(hbk["nl"].data_ptr()),
                              C.c_void_p(hbk["ne"].data_ptr()), C.c_void_p(hbk["poses"].data_ptr()),
                              C.c_void_p(hbk["points"].data_ptr()), C.c_void_p(hbk["fixed"].data_ptr()),
                              C.c_void_p(hbk["op"].data_ptr()), C.c_void_p(hbk["ol"].data_ptr()), C.c_void_p(hbk["uv"].data_ptr()),
                              C.c_void_p(Kd.ctypes.data), C.c_void_p(ext7.ctypes.data), C.c_double(5.991), C.c_double(5.991),
                              5, 10, C.c_void_p(c2.data_ptr()), C.c_void_p(ol_.data_ptr()), C.c_void_p(inf.data_ptr()))
        assert rc == 0, pkg.last_error()
        ba_pending[k] = True

    host_t = {"ba_submit": 0.0, "fe_wait": 0.0, "fe_submit": 0.0}   # where the host thread spends the end-to-end loop

    def run_host(nsteps):
        # the reference's threading: the front end (extract + match) and the back end (local BA) run side by side;
        # here both are asynchronous submissions from one host thread, collected one step later
        pending = [False] * NE
        for i in range(nsteps):
            k = i % NE
            ta = time.perf_counter()
            if with_ba:
                ba_submit(i % 2)   # first: its 4 MB of copies must not queue behind the 60 MB of frames on the copy engine
            tb = time.perf_counter()
            if pending[k]:
                fes[k].wait()
            tc = time.perf_counter()
            off = (i * B) % hp
            if e2e_lockstep:
                ev_fe[i % 4].record(s_fe)           # fires when the previous batch's kernels are done = where this batch's begin
                fe_marks.append(ev_fe[i % 4])
                del fe_marks[:-1]
            fes[k].submit(host_np[off:off + B], outs[k])
            pending[k] = True
            td = time.perf_counter()
            host_t["ba_submit"] += tb - ta
            host_t["fe_wait"] += tc - tb
            host_t["fe_submit"] += td - tc
        for k in range(NE):
            if pending[k]:
                fes[k].wait()
        for k in range(2):
            ba_wait(k)

    run_host(4)
    barrier()
    for k_ in host_t:
        host_t[k_] = 0.0
    t0 = time.perf_counter()
    run_host(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    img_bytes = H * W
    h2d = 2 * B * img_bytes
    d2h = 2 * B * cap * (28 + 32) + 2 * B * 4 + 2 * B * cap * 4
    if with_ba:
        h2d += sum(int(hb[k].numel() * hb[k].element_size()) for k in ("np", "nl", "ne", "poses", "points", "fixed", "op", "ol", "uv"))
        d2h += int(hb["poses"].numel() * 8 + hb["points"].numel() * 8 + h_chi2.numel() * 8 + h_outl.numel() + h_info.numel() * 4)

    # ---- config 4/5 extras (outside the timed region): DeepLCD scoring of every keyframe against the database,
    #      the single all-gather of keyframe poses (round-robin ownership) and the pose graph on every rank
    par = importlib.import_module(PKG + ".parallel")
    extras = {}
    try:
        n_kf = 742
        g = synth.pose_graph(0, n=n_kf)
        mine = par.shard_indices(n_kf, rank, world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        full = par.allgather_kf_poses(g["poses0"][mine], n_kf, rank, world, device="cuda")
        torch.cuda.synchronize()
        extras["kf_pose_allgather_ms"] = (time.perf_counter() - t0) * 1e3
        assert np.array_equal(full, g["poses0"])
        pgs = pkg.PoseGraph(1024, 2048, device=local_rank)
        pgs.solve(full, g["fixed"], g["v0"], g["v1"], g["meas"])
        t0 = time.perf_counter()
        _, pinfo = pgs.solve(full, g["fixed"], g["v0"], g["v1"], g["meas"])
        extras["posegraph_ms"] = (time.perf_counter() - t0) * 1e3
        extras["posegraph"] = {"vertices": n_kf, "edges": int(len(g["v0"])), "chi2_start": pinfo["chi2_start"], "chi2": pinfo["chi2"]}
        db = synth.lcd_database(0)
        lcd = pkg.DeepLCDScorer(capacity=1024, dtype=1, max_queries=n_kf, device=local_rank)
        lcd.add_batch(np.arange(n_kf), db)
        dq = torch.from_numpy(db).cuda()
        dscore = torch.zeros((n_kf, n_kf), dtype=torch.float32, device="cuda")
        sl = torch.cuda.Stream()
        lcd.set_stream(sl.cuda_stream)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lcd.score_dev(n_kf, dq, dscore, n_kf)
        a0.record(sl)
        for _ in range(10):
            lcd.score_dev(n_kf, dq, dscore, n_kf)
        a1.record(sl)
        torch.cuda.synchronize()
        lms = a0.elapsed_time(a1) / 10
        extras["lcd_score_queries_per_s"] = n_kf / (lms * 1e-3)
        extras["lcd_score_gbs"] = n_kf * n_kf * 1088 * 2 / (lms * 1e-3) / 1e9
        extras["lcd_score_path"] = "tcgen05 kind::f16 GEMM (>= 64 queries per call; a single query runs the fp16 GEMV)"
        # the same kernel at a size that fills the GPU: 4096 queries against a 16384-row database
        nq_big, n_big = 4096, 16384
        rng = np.random.default_rng(0)
        big = rng.standard_normal((n_big, 1064), dtype=np.float32)
        big /= np.linalg.norm(big, axis=1, keepdims=True)
        lcd_big = pkg.DeepLCDScorer(capacity=n_big, dtype=1, max_queries=nq_big, device=local_rank)
        lcd_big.add_batch(np.arange(n_big), big)
        lcd_big.set_stream(sl.cuda_stream)
        dq_big = torch.from_numpy(big[:nq_big]).cuda()
        ds_big = torch.zeros((nq_big, n_big), dtype=torch.float32, device="cuda")
        lcd_big.score_dev(nq_big, dq_big, ds_big, n_big)
        a0.record(sl)
        for _ in range(5):
            lcd_big.score_dev(nq_big, dq_big, ds_big, n_big)
        a1.record(sl)
        torch.cuda.synchronize()
        gms = a0.elapsed_time(a1) / 5
        extras["lcd_gemm_4096x16384"] = {"ms": gms, "tflops_fp16": 2.0 * nq_big * n_big * 1088 / (gms * 1e-3) / 1e12,
                                         "score_write_gbs": nq_big * n_big * 4 / (gms * 1e-3) / 1e9}
        lcd_big.close()
        del dq_big, ds_big
        # DeepLCD CNN forward ("next" row 2): whole-image descriptors of 64 keyframes per call (seeded random weights of the
        # CALC architecture; the trained model is a configure-time download of the reference)
        nb = 64
        net = pkg.DeepLCD(synth.calc_weights(0), max_batch=nb, max_img_w=W, max_img_h=H, device=local_rank)
        net.set_stream(sl.cuda_stream)
        ddescr = torch.zeros((nb, net.dim), dtype=torch.float32, device="cuda")
        net.descr_original_dev(nb, pool[:, 0], 2 * H * W, W, H, W, ddescr)
        a0.record(sl)
        for r in range(10):
            net.descr_original_dev(nb, pool[(r % (P // nb)) * nb:, 0], 2 * H * W, W, H, W, ddescr)
        a1.record(sl)
        torch.cuda.synchronize()
        cms = a0.elapsed_time(a1) / 10
        extras["calc_descr_keyframes_per_s"] = nb / (cms * 1e-3)
        extras["calc_descr_tflops_fp32"] = nb * 2 * CALC_MACS / (cms * 1e-3) / 1e12   # fp32-equivalent (the tensor-core layers do 3 tf32 products per multiply-add)
        extras["calc_conv_path"] = "conv2 / conv3: tcgen05 kind::tf32, 3-product split, 4-D TMA implicit GEMM; conv1: fp32 CUDA cores"
    except Exception as ex:   # the extras never invalidate the headline measurement
        extras["error"] = repr(ex)

    # ---- aggregate over ranks (max time)
    t = torch.tensor([dev_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max = float(t[0]), float(t[1])
    frames = B * args.steps * world
    value = frames / (dev_ms_max * 1e-3)
    e2e_value = frames / (e2e_ms_max * 1e-3)

    if rank == 0:
        stage_ms = {s: float(ms[i]) * args.steps / CONC_STEPS for i, s in enumerate(STAGES)}      # scaled to the timed region's steps
        stage_ms["hamming_match"] = match_ms * args.steps / CONC_STEPS
        stage_launches = {s: int(launches[i]) * args.steps // CONC_STEPS for i, s in enumerate(STAGES)}
        stage_launches["hamming_match"] = 3 * args.steps   # k_expand (both sides), k_hamming_umma, k_hamming_decode
        if with_ba:
            stage_ms["local_ba"] = ba_ms * args.steps / CONC_STEPS
            stage_launches["local_ba"] = args.steps

        def abytes(k):
            return ba_bytes if k == "local_ba" else algorithmic_bytes(k, 2 * B, n_kps, n_cands)

        # Roofline entries.  The stage times come from the SERIALISED pass (each stream alone, see above): reproducible, no
        # interference.  `roofline` names the kernel that dominates the timed step — whichever stage is longest, BA included —
        # and `roofline_extract` the dominant kernel of the extract / match stage (the one north_star's HBM target is about).
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic, ncu_extra, ncu_src = ncu_traffic()
        ser = {k: v / SER_STEPS for k, v in ser_ms.items()}          # ms per step, serialised

        def hbm_entry(k):
            ach = abytes(k) / (ser[k] * 1e-3) / 1e9
            return {"kernel": k, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(k) if 2 * B == 128 else None, "traffic_source": ncu_src, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": abytes(k), "ms_per_launch": ser[k], **ncu_extra.get(k, {})}

        def ba_entry():
            # fp64 pipe: 64 FMA / clk / SM measured for DFMA and DMMA alike (tools/dmma_probe.cu, profiles/r2_ba_dmma_ab.md)
            fp64_peak = 148 * 64 * 2 * (clocks.get("sm_max_mhz") or 1965.0) * 1e6 / 1e12
            ach = ba_flops / (ser["local_ba"] * 1e-3) / 1e12
            return {"kernel": "local_ba", "bound": "latency (fp64 pipe; no HBM or tensor roof applies: 0.07 % of HBM)",
                    "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                    "traffic": traffic.get("local_ba") if B == 64 else None, "traffic_source": ncu_src,
                    "peak_source": "148 SMs x 64 fp64 FMA/clk (measured per-SM rate) x SM clock; no fp64 figure in MEASURED_PEAKS.json",
                    "algorithmic_flops_per_launch": ba_flops, "algorithmic_bytes_per_launch": ba_bytes,
                    "ms_per_launch": ser["local_ba"], "sm_ms_per_launch": B * ser["local_ba"], **ncu_extra.get("local_ba", {})}

        top = max(ser, key=ser.get)
        top_extract = max((k for k in ser if k != "local_ba"), key=ser.get)
        roofline = ba_entry() if top == "local_ba" else hbm_entry(top)
        roofline["stage_ms_per_step_serialised"] = ser
        roofline["stage_ms_per_step_concurrent"] = {k: v / args.steps for k, v in stage_ms.items()}
        roofline["stage_gbs_serialised"] = {k: abytes(k) / (ser[k] * 1e-3) / 1e9 for k in ser if ser[k] > 0}
        roofline["note"] = ("serialised = each stream run alone after the timed region (CUDA events on the launching stream, "
                            f"{SER_STEPS} steps); concurrent = a pass of {CONC_STEPS} steps right after the timed region with the BA stream and the extract stream overlapping as in it (the timed region itself carries no per-stage events)")
        roofline_extract = hbm_entry(top_extract)
        out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u8 (extract/match), f64 (BA)", "data": "synthetic",
               "config": {"workload": WORKLOAD if with_ba else "config 2: ORB extract (both views) + L<->R Hamming match only",
                          "stereo_pairs_per_step_per_gpu": B, "ba_windows_per_step_per_gpu": B if with_ba else 0,
                          "resident_pool_pairs": P,
                          "l2_policy": f"inputs larger than L2: {P * 2 * img_bytes / 1e6:.0f} MB pool of distinct frames cycled",
                          "keypoints_per_frame": n_kps / (2 * B), "matches_per_pair": n_matched / B,
                          "ba_obs_per_window": float(bh["ne"].mean()), "ba_lm_iterations_per_window": float(ba_info[:, 1].mean()),
                          "sharding": "frames and windows round-robin by rank, no data-path collective",
                          "streams": (f"{NH} extract + match stream(s) (batches in flight), {len(ba_set) if with_ba else 0} BA stream(s)"
                                      + (", BA batch i released at the start of extraction i" if with_ba and lockstep else ""))},
               "clocks": clocks,
               "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": e2e_ms_max / args.steps,
                       "host_thread_ms_per_step": {k_: 1e3 * v_ / args.steps for k_, v_ in host_t.items()},
                       "api": f"sb_stereo_submit/wait on {NE} handles" + (" (kernels on one shared stream, sb_stereo_set_compute_stream)" if fe_shared else "") + " and sb_ba_submit/wait on two, used in turn (host pointers, pinned)"},
               "gpu_launches": int(sum(stage_launches.values())),
               "roofline": roofline,
               "roofline_extract": roofline_extract,
               "loop_closing_extras": extras}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = max(8 * cores, 64)                         # about 10 s of CPU work (0.08 s per frame and core)
            sample = pool_np[:min(n, P)]
            fps, kind, kind_detail = cpu_reference_fps(sample, windows if with_ba else None, cores)
            out["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "kind_detail": kind_detail,
                                   "sample": f"{len(sample)} of the same synthetic stereo frames (+ one BA window each) through "
                                             "oracle/ (C restatement of the reference's OpenCV/g2o-based path), one frame per thread"}
            try:   # BASELINE.md section 4 leg: the OpenCV primitives through cv2 (SIMD builds), 1 thread and all cores
                inv = [np.float32(1.0)]
                sc = np.float32(1.0)
                for _ in range(1, 8):
                    sc = np.float32(np.float64(sc) * np.float64(np.float32(1.2)))
                    inv.append(np.float32(1.0) / sc)
                sizes = [(W, H)] + [(int(np.rint(np.float32(W) * i)), int(np.rint(np.float32(H) * i))) for i in inv[1:]]
                rng = np.random.default_rng(0)
                dpair = (rng.integers(0, 256, (2000, 32), dtype=np.uint8), rng.integers(0, 256, (2000, 32), dtype=np.uint8))
                cvp = Cv2Primitives(sizes, dpair)
                f1 = cvp.fps(sample[:max(4, len(sample) // 8)], 1)
                fa = cvp.fps(sample, cores)
                out["cpu_baseline_cv2"] = {"value": fa, "value_1_thread": f1, "unit": "frames/s", "cores": cores, "kind": "cv2 4.13 primitives only",
                                           "sample": f"{len(sample)} of the same frames: chained cv2.resize pyramid, per-cell cv2 FAST (20 -> 7 "
                                                     "fallback), GaussianBlur per level, BFMatcher 2000 x 2000 per pair; the reference's own "
                                                     "quadtree / orientation / rBRIEF code and BA are NOT included: a lower bound on its CPU time"}
            except Exception as ex:
                out["cpu_baseline_cv2"] = {"unavailable": repr(ex)}
        emit(out)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
