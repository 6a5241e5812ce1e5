"""CPU checker for the PnP-RANSAC stage (TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing else; the product
never loads it).

The reference calls cv::solvePnPRansac(vLoopPoints3d, vCurrentPoints2d, K, Mat(), rvec, tvec, false, 100, 5.991, 0.99)
(src/loopclosing.cpp:263-264) — OpenCV is an un-vendored third-party dependency (README: 3.4.8); the implementation
available in this image is cv2 4.13.0, which is called here directly with the reference's arguments.  Its samples come
from cv::RNG, so two valid implementations agree on well-posed problems (same inlier set up to correspondences whose
error sits at the threshold, same pose up to the refinement's termination tolerance), not bit for bit: parity for this
row is therefore a tolerance on the pose and equality of the inlier sets away from the threshold, both stated in
tests/test_gpu_pnp.py.  The reference holds no golden vectors for this path.
"""
import numpy as np


def solve_pnp_ransac(obj, img, K, iterations=100, reproj_err=5.991, confidence=0.99):
    """-> (found, rvec [3], tvec [3], inlier mask [n]) exactly as the reference's call would return them."""
    import cv2
    fx, fy, cx, cy = K
    Kmat = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    obj = np.ascontiguousarray(obj, np.float32).reshape(-1, 1, 3)
    img = np.ascontiguousarray(img, np.float32).reshape(-1, 1, 2)
    try:
        ok, rvec, tvec, inl = cv2.solvePnPRansac(obj, img, Kmat, None, None, None, False, iterations, reproj_err, confidence)
    except cv2.error:
        return False, np.zeros(3), np.zeros(3), np.zeros(len(obj), bool)
    mask = np.zeros(len(obj), bool)
    if ok and inl is not None:
        mask[inl.ravel()] = True
    return bool(ok), rvec.ravel().astype(np.float64), tvec.ravel().astype(np.float64), mask


def reprojection_errors(obj, img, K, rvec, tvec):
    """Pixel distances of the correspondences under (rvec, tvec) — numpy restatement of cv::projectPoints without distortion."""
    fx, fy, cx, cy = K
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        R = np.eye(3)
    else:
        k = rvec / th
        Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    pc = np.asarray(obj, np.float64) @ R.T + tvec
    uv = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1)
    return np.linalg.norm(uv - np.asarray(img, np.float64), axis=1)
