"""Parity of the CUDA PnP-RANSAC (loop geometric verification, src/loopclosing.cpp:207-293) with cv2.solvePnPRansac
called with the reference's arguments.  OpenCV draws its samples from cv::RNG, so the comparison is by tolerance:
  * inlier sets: both are the inliers of a best MINIMAL-sample model (4-5 noisy points), so they legitimately differ on
    correspondences that are inliers of the refined pose but not of one coarse model.  Checked: no gross outlier
    (error > 3 x threshold under the checker's refined pose) is accepted, the sets differ on < 5 % of the
    correspondences, and at least 95 % as many inliers as the checker are found;
  * pose: both refine by least squares on almost the same inliers, so they agree to a fraction of the estimate's own
    standard error (0.7 px pixel noise, scene depth 4-60 m): translation within max(1e-2, 0.5 / sqrt(#inliers)) metres,
    rotation within max(1e-3, 0.05 / sqrt(#inliers)) rad (2.4e-2 m / 2.4e-3 rad at 420 inliers) — and the
    returned pose must be the least-squares optimum of its own inlier set (the checker's pose cannot fit it better).
The chain ComputeCorrectPose runs next (OptimizeCurrentPose = sb_pose_solve with pre_rounds = 1) is checked too."""
import numpy as np
import pytest

from oracle import pnp_oracle as PO

pytestmark = pytest.mark.gpu


def _rvec_to_R(rvec):
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def _check(synth, pr, got):
    ok, rvec, tvec, mask = PO.solve_pnp_ransac(pr["obj"], pr["img"], synth.KITTI_K)
    assert ok and got["found"]
    err = PO.reprojection_errors(pr["obj"], pr["img"], synth.KITTI_K, rvec, tvec)
    diff = got["inliers"] != mask
    assert not (got["inliers"] & (err > 3 * 5.991)).any()
    assert diff.mean() < 0.05, (int(diff.sum()), err[diff])
    assert got["inliers"].sum() >= 0.95 * mask.sum(), (int(got["inliers"].sum()), int(mask.sum()))
    ninl = max(int(mask.sum()), 1)
    tol_t, tol_r = max(1e-2, 0.5 / np.sqrt(ninl)), max(1e-3, 0.05 / np.sqrt(ninl))
    assert np.abs(got["tvec"] - tvec).max() < tol_t, (got["tvec"], tvec, tol_t)
    dR = _rvec_to_R(got["rvec"]) @ _rvec_to_R(rvec).T
    ang = np.arccos(np.clip((np.trace(dR) - 1) / 2, -1, 1))
    assert ang < tol_r, (ang, tol_r)
    # the returned pose is the least-squares optimum on the returned inliers: no other pose (the checker's) fits them better
    own = got["inliers"]
    mine = PO.reprojection_errors(pr["obj"], pr["img"], synth.KITTI_K, got["rvec"], got["tvec"])
    assert (mine[own] ** 2).sum() <= (err[own] ** 2).sum() * (1 + 1e-9)
    # pose7 is the same pose as (rvec, tvec)
    x, y, z, w = got["pose7"][:4]
    Rq = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                   [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                   [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    assert np.allclose(Rq, _rvec_to_R(got["rvec"]), atol=1e-9) and np.allclose(got["pose7"][4:], got["tvec"])


def test_batch_of_loop_candidates(pkg, synth):
    probs = [synth.pnp_problem(s, n_points=n, outlier_frac=f) for s, n, f in
             ((0, 400, 0.3), (1, 400, 0.1), (2, 1500, 0.4), (3, 60, 0.2), (4, 12, 0.0), (5, 800, 0.5))]
    solver = pkg.PnPRansac(max_problems=len(probs), max_points=2048)
    res = solver.solve([(p["obj"], p["img"]) for p in probs], synth.KITTI_K)
    for pr, got in zip(probs, res):
        _check(synth, pr, got)
    again = solver.solve([(p["obj"], p["img"]) for p in probs], synth.KITTI_K)      # deterministic
    for a, b in zip(res, again):
        assert np.array_equal(a["pose7"], b["pose7"]) and np.array_equal(a["inliers"], b["inliers"])
    solver.close()


def test_degenerate_inputs(pkg, synth):
    solver = pkg.PnPRansac(max_problems=3, max_points=256)
    few = synth.pnp_problem(7, n_points=3)
    rng = np.random.default_rng(0)
    junk_obj = rng.uniform(-10, 10, (100, 3)).astype(np.float32)                     # no consistent pose at all
    junk_img = np.stack([rng.uniform(0, 1241, 100), rng.uniform(0, 376, 100)], 1).astype(np.float32)
    good = synth.pnp_problem(8, n_points=100)
    res = solver.solve([(few["obj"], few["img"]), (junk_obj, junk_img), (good["obj"], good["img"])], synth.KITTI_K)
    assert not res[0]["found"] and not res[0]["inliers"].any()
    assert (not res[1]["found"]) or res[1]["inliers"].sum() < 10                    # the reference rejects < 10 inliers (:284)
    _check(synth, good, res[2])
    solver.close()


def test_compute_correct_pose_chain(pkg, synth):
    """solvePnPRansac -> OptimizeCurrentPose (src/loopclosing.cpp:263-275): the refined pose keeps >= 10 inliers and sits
    at the planted pose."""
    pr = synth.pnp_problem(11, n_points=500, outlier_frac=0.3)
    got = pkg.PnPRansac(max_problems=1, max_points=512).solve([(pr["obj"], pr["img"])], synth.KITTI_K)[0]
    assert got["found"]
    opt = pkg.PoseOnlyOptimizer(max_frames=1, max_obs=512)
    pose, outl, info = opt.solve([dict(pose0=got["pose7"], points=pr["obj"].astype(np.float64), uv=pr["img"].astype(np.float64))],
                                 synth.KITTI_K, pre_rounds=1)[0]
    assert info[0] >= 10
    assert np.abs(pose[4:] - pr["pose_gt"][4:]).max() < 0.05
    assert (outl[pr["planted"]].mean() > 0.9) and (outl[~pr["planted"]].mean() < 0.1)
