"""BASELINE config 4 as a tested run (src/loopclosing.cpp:52-121 flow): the replay's host logic on the CPU checker, and
the CUDA replay against it — every planted revisit must be detected, verified and closed, identically on both sides."""
import importlib

import numpy as np
import pytest

PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
SMALL = dict(frames=44, n_kf=22)          # revisits planted at keyframes 13 -> 2 and 20 -> 9
RUN = dict(db_min_size=5, min_gap=5, with_digests=True)


@pytest.fixture(scope="module")
def cpu_result(synth):
    replay = importlib.import_module(PKG + ".replay")
    capi = importlib.import_module(PKG + ".capi")
    from replay_cpu_ops import CpuOps
    seq = replay.Sequence(**SMALL)
    assert seq.loop_pairs == [(13, 2), (20, 9)]
    return replay.run(seq, CpuOps(synth, capi.KP_DTYPE), **RUN)


def test_sequence_layout(synth):
    replay = importlib.import_module(PKG + ".replay")
    seq = replay.Sequence()
    assert seq.frames == 4541 and seq.n_kf == 742 and len(seq.loop_pairs) == 17
    assert len(set(seq.kf_frame.tolist())) == 742 and seq.kf_frame[0] == 0
    a, b = seq.loop_pairs[0]
    assert np.array_equal(seq.T_gt[a], seq.T_gt[b])                         # a revisit stands where its partner stood
    la, _ = seq.frame_images(seq.kf_frame[a])
    lb, _ = seq.frame_images(seq.kf_frame[b])
    assert 0 < np.abs(la.astype(int) - lb).max() <= 8                       # same scene, fresh sensor noise
    T = replay.T_from7(replay.T_to7(seq.T_gt[5]))
    assert np.abs(T - seq.T_gt[5]).max() < 1e-12
    assert abs(replay.se3_log_norm(seq.odo[5]) - np.linalg.norm(__import__("oracle.posegraph_oracle", fromlist=["x"]).se3_log(
        (seq.odo[5][:3, :3], seq.odo[5][:3, 3])))) < 1e-9


def test_cpu_replay_closes_every_planted_loop(cpu_result):
    r = cpu_result
    assert [l[:2] for l in r["loops"]] == [[13, 2], [20, 9]], r["loops"]
    assert all(l[3] >= 10 for l in r["loops"])
    assert r["keypoints"] > 44 * 2 * 1900 and len(r["frame_digests"]) == 44
    if r["posegraph_runs"]:
        assert r["mean_position_error_final_m"] < r["mean_position_error_dead_reckoned_m"]


@pytest.mark.gpu
def test_cuda_replay_equals_the_cpu_replay(cpu_result, pkg):
    replay = importlib.import_module(PKG + ".replay")
    seq = replay.Sequence(**SMALL)
    got = replay.run(seq, replay.GpuOps(batch=8, kf_batch=8, n_kf=seq.n_kf), **RUN)
    want = cpu_result
    assert got["frame_digests"] == want["frame_digests"]                    # keypoints, descriptors, matches of every frame: bit-exact
    assert [l[:3] for l in got["loops"]] == [l[:3] for l in want["loops"]]  # same loops, same "needs correction" decisions
    assert got["posegraph_runs"] == want["posegraph_runs"]
    assert abs(got["mean_position_error_final_m"] - want["mean_position_error_final_m"]) < 0.05
    assert got["kf_pose_digest"] != "" and got["keypoints"] == want["keypoints"] and got["matches"] == want["matches"]


def test_batched_octave_expansion_equals_the_per_keyframe_one(pkg):
    """replay.expand_octaves_batch packs src/loopclosing.cpp:94-105 for a whole batch of keyframes (ragged, one empty)."""
    import importlib
    replay = importlib.import_module(pkg.__name__ + ".replay")
    capi = importlib.import_module(pkg.__name__ + ".capi")
    rng = np.random.default_rng(0)
    feats = []
    for n in (300, 17, 0, 123):
        f = np.zeros(n, capi.KP_DTYPE)
        f["x"], f["y"] = rng.uniform(0, 1241, n), rng.uniform(0, 376, n)
        f["size"], f["angle"] = 7, -1
        feats.append(f)
    kin, n_in = replay.expand_octaves_batch(feats)
    assert kin.shape == (4, 2400) and n_in.tolist() == [2400, 136, 0, 984]
    for b, f in enumerate(feats):
        assert kin[b, :n_in[b]].tobytes() == replay.expand_octaves(f).tobytes()
        assert not kin[b, n_in[b]:].view(np.uint8).any()
