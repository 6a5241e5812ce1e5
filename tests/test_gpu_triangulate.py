"""Batched DLT triangulation on the GPU against the numpy restatement of myslam::triangulation
(include/myslam/algorithm.h:16-33, callers src/frontend.cpp:385-417,451-488); tolerance 1e-4 relative on accepted points."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_stereo_correspondences(pkg, oracle, synth):
    rng = np.random.default_rng(0)
    K = synth.KITTI_K
    baseline = synth.KITTI_BF / synth.KITTI_FX                     # 0.537 m
    pose_l = np.array([0, 0, 0, 1, 0, 0, 0.0])
    pose_r = np.array([0, 0, 0, 1, -baseline, 0, 0.0])             # right camera: x_r = x_l - b
    n = 3000
    z = rng.uniform(3, 90, n)
    u = rng.uniform(20, 1220, n)
    v = rng.uniform(20, 356, n)
    X = np.stack([(u - K[2]) / K[0] * z, (v - K[3]) / K[1] * z, z], 1)
    ul = np.stack([u, v], 1) + rng.normal(0, 0.3, (n, 2))
    ur = np.stack([u - synth.KITTI_BF / z, v], 1) + rng.normal(0, 0.3, (n, 2))
    ur[::50] = rng.uniform(0, 376, (len(ur[::50]), 2))             # wrong matches: large vertical disparity or negative depth
    twc = synth.pose7(synth._rot_y(0.3), np.array([4.0, -1.0, 12.0]))
    for T in (None, twc):
        gp, gok = pkg.triangulate(ul, ur, K, K, pose_l, pose_r, T)
        wp, wok = oracle.triangulate(ul, ur, K, K, pose_l, pose_r, T)
        assert np.array_equal(gok, wok)
        assert gok.sum() > 0.9 * n and (~gok).sum() > 10
        sel = wok
        assert np.all(np.abs(gp[sel] - wp[sel]) <= 1e-4 * np.maximum(1.0, np.abs(wp[sel]))), np.abs(gp[sel] - wp[sel]).max()
    # the clean correspondences triangulate back to the planted points
    gp, gok = pkg.triangulate(np.stack([u, v], 1), np.stack([u - synth.KITTI_BF / z, v], 1), K, K, pose_l, pose_r)
    good = gok & (z < 40)
    assert np.abs(gp[good] - X[good]).max() < 0.05 * 40
    assert pkg.triangulate(np.zeros((0, 2)), np.zeros((0, 2)), K, K, pose_l, pose_r)[0].shape == (0, 3)
