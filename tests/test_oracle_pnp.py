"""The PnP-RANSAC checker (cv2.solvePnPRansac with the reference's arguments, src/loopclosing.cpp:263-264) on the
seeded loop candidates of synth.pnp_problem: it must recover the planted pose and separate the planted outliers —
otherwise the GPU parity test would be comparing against noise."""
import numpy as np

from oracle import pnp_oracle as PO


def _rot_from_quat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_cv2_recovers_planted_pose_and_outliers(synth):
    for seed in range(4):
        pr = synth.pnp_problem(seed)
        ok, rvec, tvec, mask = PO.solve_pnp_ransac(pr["obj"], pr["img"], synth.KITTI_K)
        assert ok
        R, t = _rot_from_quat(pr["pose_gt"][:4]), pr["pose_gt"][4:]
        err = PO.reprojection_errors(pr["obj"], pr["img"], synth.KITTI_K, rvec, tvec)
        assert np.abs(tvec - t).max() < 0.05
        assert (mask[~pr["planted"]].mean() > 0.98) and (mask[pr["planted"]].mean() < 0.1)
        assert err[mask].max() <= 5.991 * 1.5          # inliers of the RANSAC model, measured under the refined pose


def test_too_few_points_is_reported(synth):
    pr = synth.pnp_problem(9, n_points=3)
    ok, _, _, mask = PO.solve_pnp_ransac(pr["obj"], pr["img"], synth.KITTI_K)
    assert not ok and not mask.any()
