"""Multi-GPU (SURVEY.md §4 `tests/multi`, §8e): the exchange step through the C ABI from a plain C++ host, and an
N-rank sharded replay under torchrun that must reproduce the 1-GPU results byte for byte.  Skipped with < 2 GPUs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKGD = os.path.join(ROOT, "a-simple-stereo-slam-system-with-deep-loop-closing_b200")


def n_gpus():
    import torch
    return torch.cuda.device_count()


def test_cpp_host_allgathers_kf_poses_over_nccl(tmp_path, pkg):
    if n_gpus() < 2:
        pytest.skip("needs >= 2 GPUs")
    exe = str(tmp_path / "nccl_allgather_test")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-o", exe, os.path.join(ROOT, "tests", "nccl_allgather_test.cpp"),
                           "-L" + PKGD, "-lslamb200", "-Wl,-rpath," + PKGD])
    for world, n_kf in ((2, 742), (min(n_gpus(), 8), 101), (2, 1)):
        out = subprocess.run([exe, str(world), str(n_kf)], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0 and "\nOK:" in "\n" + out.stdout, out.stdout + out.stderr   # NCCL prints its version first


def test_sharded_replay_reproduces_the_single_gpu_result(tmp_path):
    """tools/replay_kitti.py on 1 rank and on N ranks: identical keypoint / descriptor / match digests for every frame,
    identical loop decisions and an identical optimised pose graph."""
    world = min(n_gpus(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    outs = []
    for n, name in ((1, "one.json"), (world, "many.json")):
        path = str(tmp_path / name)
        cmd = [sys.executable, os.path.join(ROOT, "tools", "replay_kitti.py"), "--frames", "96", "--kf-every", "3", "--digests", path]
        if n > 1:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                   "--master-port", "29671"] + cmd[1:]
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs.append(json.load(open(path)))
    one, many = outs
    assert many["world"] == world and one["world"] == 1
    assert one["frame_digests"] == many["frame_digests"]
    assert one["loops"] == many["loops"] and len(one["loops"]) > 0
    assert one["posegraph_digest"] == many["posegraph_digest"]
