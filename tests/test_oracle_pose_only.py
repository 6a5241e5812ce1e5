"""First-principles checks of the pose-only checker (oracle/ba_oracle.c: orc_pose_only_solve), the restatement of
Frontend::EstimateCurrentPose's solver (src/frontend.cpp:176-276).  "Parity unpinned" like the BA oracle."""
import numpy as np


def test_recovers_pose_and_flags_planted_outliers(oracle, synth):
    for seed in range(4):
        f = synth.pose_only_frame(seed)
        pose, outl, info = oracle.pose_only_solve(f["pose0"], f["points"], f["uv"], synth.KITTI_K)
        assert np.abs(pose - f["pose_gt"])[4:].max() < 0.05 and np.abs(pose - f["pose_gt"])[:4].max() < 2e-3
        # every grossly displaced observation (> 40 px box) that is really off is flagged, few inliers are
        off = np.linalg.norm(f["uv"] - synth.ba_window(seed, n_poses=1, n_points=250, pix_noise=0.0, outlier_frac=0.0,
                                                       pose_noise=(0, 0), point_noise=0.0, fixed_frac=1.0)["uv"], axis=1)
        assert outl[off > 6].all()
        assert outl[off < 1.5].mean() < 0.1
        assert info[0] == len(outl) - outl.sum() and info[2] == 4


def test_pre_round_variant_and_fixed_point(oracle, synth):
    f = synth.pose_only_frame(7, outlier_frac=0.0, pix_noise=0.0, pose_noise=(0, 0))
    pose, outl, info = oracle.pose_only_solve(f["pose_gt"], f["points"], f["uv"], synth.KITTI_K, pre_rounds=1)
    assert np.abs(pose - f["pose_gt"]).max() < 1e-6 and outl.sum() == 0 and info[2] == 5
