"""Parity of the CUDA pose-only optimiser with the oracle (src/frontend.cpp:176-276, src/loopclosing.cpp:339-433);
tolerance 1e-4 relative on the pose, identical outlier decisions."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pre_rounds", [0, 1])
def test_batch_matches_oracle(pkg, oracle, synth, pre_rounds):
    frames = [synth.pose_only_frame(s, n_points=120 + 40 * (s % 4)) for s in range(12)]
    frames.append({"pose0": frames[0]["pose0"], "points": np.zeros((0, 3)), "uv": np.zeros((0, 2))})   # a frame without matches
    opt = pkg.PoseOnlyOptimizer(max_frames=16, max_obs=512)
    res = opt.solve(frames, synth.KITTI_K, pre_rounds=pre_rounds)
    for f, (pose, outl, info) in zip(frames, res):
        wp, wo, wi = oracle.pose_only_solve(f["pose0"], f["points"], f["uv"], synth.KITTI_K, pre_rounds=pre_rounds)
        assert np.all(np.abs(pose - wp) <= 1e-4 * np.maximum(1.0, np.abs(wp))), np.abs(pose - wp).max()
        assert np.array_equal(outl, wo)
        assert info[0] == wi[0] and info[2] == wi[2]
    assert np.array_equal(res[-1][0], frames[0]["pose0"])      # nothing to optimise: the pose is returned untouched
    opt.close()
