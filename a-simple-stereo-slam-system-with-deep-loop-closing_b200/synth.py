"""Seeded synthetic replay inputs: KITTI-shaped stereo pairs, BA windows,
DeepLCD databases and pose graphs.  numpy only; used by tests/, bench.py and smoke()."""
import numpy as np

KITTI_W, KITTI_H = 1241, 376
KITTI_FX = KITTI_FY = 718.856
KITTI_CX, KITTI_CY = 607.1928, 185.2157
KITTI_BF = 386.1448  # config/stereo/gray/KITTI00-02.yaml:35


def _blur3(img):
    """3x3 Gaussian, sigma 0.8, edge-replicated, float32 in/out."""
    k = np.exp(-np.arange(-1, 2) ** 2 / (2 * 0.8 ** 2)).astype(np.float32)
    k /= k.sum()
    p = np.pad(img, 1, mode="edge")
    t = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * t[:-2] + k[1] * t[1:-1] + k[2] * t[2:]


def stereo_pair(seed, w=KITTI_W, h=KITTI_H, n_rect=1500, noise=4, grid=(3, 8), noise_seed=None):
    """Left/right u8 images of one synthetic stereo frame.

    Left: mid-grey canvas + `n_rect` filled random rectangles (w 4-60, h 4-40, grey 0-255) + 3x3
    Gaussian (sigma 0.8).  Depth is piecewise planar: a `grid` of fronto-parallel planes with
    z in [5, 80] m; the right view samples the left one at x + round(bf / z) inside each plane.
    Independent uniform noise of +-`noise` is then added to both views.  `noise_seed` (replay: a place seen again)
    draws that noise from its own generator: same scene, different sensor noise.
    """
    rng = np.random.default_rng(seed)
    canvas = np.full((h, w), 128, np.float32)
    rw = rng.integers(4, 61, n_rect)
    rh = rng.integers(4, 41, n_rect)
    x0 = rng.integers(-20, w, n_rect)
    y0 = rng.integers(-10, h, n_rect)
    g = rng.integers(0, 256, n_rect)
    for i in range(n_rect):
        canvas[max(0, y0[i]):min(h, y0[i] + rh[i]), max(0, x0[i]):min(w, x0[i] + rw[i])] = g[i]
    clean = _blur3(canvas)
    z = rng.uniform(5.0, 80.0, grid)
    disp = np.rint(KITTI_BF / z).astype(np.int64)
    ys = np.minimum(np.arange(h) * grid[0] // h, grid[0] - 1)
    xs = np.minimum(np.arange(w) * grid[1] // w, grid[1] - 1)
    src_x = np.clip(np.arange(w)[None, :] + disp[ys][:, xs], 0, w - 1)
    right_clean = np.take_along_axis(clean, src_x, axis=1)
    out = []
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)
    for img in (clean, right_clean):
        img = img + rng.integers(-noise, noise + 1, (h, w)).astype(np.float32)
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return out[0], out[1]


def stereo_batch(seed0, batch, w=KITTI_W, h=KITTI_H):
    """[batch, 2, h, w] u8: image (b, 0) = left, (b, 1) = right of frame seed0 + b."""
    out = np.empty((batch, 2, h, w), np.uint8)
    for b in range(batch):
        out[b, 0], out[b, 1] = stereo_pair(seed0 + b, w, h)
    return out


# ---------------------------------------------------------------------------------------------------
# Local-BA windows.  Poses are T_cw as (qx, qy, qz, qw, tx, ty, tz) — Sophus'
# storage order — landmarks are world points, observations are left-camera pixels.
# ---------------------------------------------------------------------------------------------------
KITTI_K = (KITTI_FX, KITTI_FY, KITTI_CX, KITTI_CY)
IDENTITY_EXT = (0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0)  # left camera extrinsics = identity (src/system.cpp:141-142)


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _so3_exp(w):
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + W
    return np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W


def _quat_from_R(R):
    tr = np.trace(R)
    if tr > 0:
        s = np.sqrt(tr + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        q[3] = (R[k, j] - R[j, k]) / s
    if q[3] < 0:
        q = -q
    return q / np.linalg.norm(q)


def pose7(R, t):
    return np.concatenate([_quat_from_R(R), t])


def ba_window(seed, n_poses=7, n_points=300, pix_noise=1.0, outlier_frac=0.05, pose_noise=(0.02, 0.2), point_noise=0.3,
              fixed_frac=0.3, w=KITTI_W, h=KITTI_H):
    """One sliding window.  Returns a dict of numpy arrays:
    poses0 [P,7] / points0 [L,3] (perturbed start), poses_gt / points_gt, fixed [L] u8,
    obs_pose / obs_point [E] i32, uv [E,2] f64."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = KITTI_K
    centres, Rs = [], []
    z = 0.0
    for i in range(n_poses):
        centres.append(np.array([rng.normal(0, 0.05), rng.normal(0, 0.02), z]))
        Rs.append(_rot_y(rng.normal(0, 0.03)).T)      # R_cw
        z += rng.uniform(6.0, 8.0)                     # keyframe spacing read off result/trajectory.txt:1-5
    poses_gt = np.stack([pose7(R, -R @ c) for R, c in zip(Rs, centres)])
    # landmarks in the union frustum, depth 5-50 m ahead of a random keyframe
    pts = np.empty((n_points, 3))
    for j in range(n_points):
        k = rng.integers(0, n_poses)
        d = rng.uniform(5.0, 50.0)
        u, v = rng.uniform(0, w), rng.uniform(0, h)
        pc = np.array([(u - cx) / fx * d, (v - cy) / fy * d, d])
        pts[j] = Rs[k].T @ pc + centres[k]
    obs_pose, obs_point, uv = [], [], []
    for j in range(n_points):
        for i in range(n_poses):
            pc = Rs[i] @ (pts[j] - centres[i])
            if pc[2] <= 0.5:
                continue
            u, v = fx * pc[0] / pc[2] + cx, fy * pc[1] / pc[2] + cy
            if 0 <= u < w and 0 <= v < h:
                nu, nv = rng.normal(0, pix_noise, 2) if pix_noise > 0 else (0.0, 0.0)
                if rng.uniform() < outlier_frac:
                    nu, nv = rng.uniform(-50, 50, 2)
                obs_pose.append(i)
                obs_point.append(j)
                uv.append((np.float32(u + nu), np.float32(v + nv)))  # cv::KeyPoint::pt is float
    fixed = (rng.uniform(size=n_points) < fixed_frac).astype(np.uint8)
    poses0 = poses_gt.copy()
    for i in range(n_poses):
        dR = _so3_exp(rng.normal(0, pose_noise[0], 3)) if pose_noise[0] > 0 else np.eye(3)
        dt = rng.normal(0, pose_noise[1], 3) if pose_noise[1] > 0 else np.zeros(3)
        R = dR @ Rs[i]
        poses0[i] = pose7(R, dR @ (-Rs[i] @ centres[i]) + dt)
    points0 = pts + (rng.normal(0, point_noise, pts.shape) if point_noise > 0 else 0.0)
    points0[fixed == 1] = pts[fixed == 1]   # fixed landmarks were optimised by earlier windows
    return {"poses0": poses0, "points0": points0, "poses_gt": poses_gt, "points_gt": pts, "fixed": fixed,
            "obs_pose": np.array(obs_pose, np.int32), "obs_point": np.array(obs_point, np.int32),
            "uv": np.array(uv, np.float64).reshape(-1, 2)}


# ---------------------------------------------------------------------------------------------------
# DeepLCD database: unit-norm 1064-d descriptors with planted near-duplicates
# at the loop pairs of the reference's result/loopEdges.txt.
# ---------------------------------------------------------------------------------------------------
LOOP_PAIRS = [(389, 27), (395, 30), (402, 35), (408, 38), (414, 42), (421, 46), (428, 50), (582, 100), (589, 104),
              (596, 109), (603, 113), (610, 117), (690, 160), (697, 165), (704, 170), (711, 175), (718, 180)]


def lcd_database(seed=0, n=742, dim=1064, pairs=LOOP_PAIRS, cos=0.96):
    """[n, dim] float32, rows L2-normalised; row a of every (a, b) in `pairs` has cosine ~`cos` with row b."""
    rng = np.random.default_rng(seed)
    d = rng.normal(0, 1, (n, dim))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for a, b in pairs:
        if a < n and b < n:
            noise = rng.normal(0, 1, dim)
            noise -= (noise @ d[b]) * d[b]
            noise /= np.linalg.norm(noise)
            d[a] = cos * d[b] + np.sqrt(1 - cos * cos) * noise
    d = d.astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)   # descriptor /= descriptor.norm() in fp32 (src/deeplcd.cpp:88)
    return d


# ---------------------------------------------------------------------------------------------------
# Pose graph: keyframes on a closed route driven ~1.97 times, sequential edges
# from noisy odometry, loop edges where the route is revisited.  Poses are T_cw in Sophus storage order.
# ---------------------------------------------------------------------------------------------------
def pose_graph(seed=0, n=742, n_loops=17, n_active=7, odo_noise=(0.0005, 0.02), loop_noise=(0.001, 0.01), radius=300.0):
    """Returns dict: poses0 [n,7] (dead-reckoned start), poses_gt [n,7], fixed [n] u8, v0/v1 [m] i32, meas [m,7].
    Edge k connects vertex v0 (the later keyframe) to v1 like EdgePoseGraph: measurement = T_v0 * T_v1^-1."""
    rng = np.random.default_rng(seed)
    per_rev = max(8, int(round(n / 1.97)))
    Rs, ts = [], []
    for i in range(n):
        ang = 2 * np.pi * i / per_rev
        c = np.array([radius * np.sin(ang), 0.02 * i, radius * (1 - np.cos(ang))])   # slow climb: revisits are near, not equal
        Rwc = _rot_y(ang)
        Rs.append(Rwc.T)
        ts.append(-Rwc.T @ c)
    Rs, ts = np.array(Rs), np.array(ts)

    def rel(i, j, noise):   # T_i * T_j^-1 with noise on the left
        R = Rs[i] @ Rs[j].T
        t = ts[i] - R @ ts[j]
        dR = _so3_exp(rng.normal(0, noise[0], 3))
        return dR @ R, dR @ t + rng.normal(0, noise[1], 3)

    v0, v1, meas = [], [], []
    for i in range(1, n):
        R, t = rel(i, i - 1, odo_noise)
        v0.append(i); v1.append(i - 1); meas.append(pose7(R, t))
    loops = []
    if n > per_rev + 25 and n_loops > 0:
        cand = np.linspace(per_rev + 12, n - n_active - 3, n_loops).astype(int)
        for i in cand:
            j = int(i - per_rev + rng.integers(-2, 3))
            if 0 < j < i - 20:
                loops.append((int(i), j))
    for i, j in loops:
        R, t = rel(i, j, loop_noise)
        v0.append(i); v1.append(j); meas.append(pose7(R, t))
    # dead reckoning from the noisy odometry
    p0R, p0t = [Rs[0]], [ts[0]]
    for k in range(n - 1):
        q = meas[k]
        Rm = _quat_to_R_np(q[:4])
        p0R.append(Rm @ p0R[-1])
        p0t.append(Rm @ p0t[-1] + q[4:])
    # the reference corrects the current keyframe through the loop before optimising: start the fixed tail at truth
    fixed = np.zeros(n, np.uint8)
    fixed[0] = 1
    fixed[max(0, n - n_active):] = 1
    if loops:
        fixed[loops[-1][1]] = 1
    poses0 = np.stack([pose7(R, t) for R, t in zip(p0R, p0t)])
    poses_gt = np.stack([pose7(R, t) for R, t in zip(Rs, ts)])
    poses0[fixed == 1] = poses_gt[fixed == 1]
    return {"poses0": poses0, "poses_gt": poses_gt, "fixed": fixed, "v0": np.array(v0, np.int32),
            "v1": np.array(v1, np.int32), "meas": np.array(meas), "loops": loops}


def scene_depth(seed, grid=(3, 8)):
    """The plane depths z [grid] of stereo_pair(seed) (same generator, same draw order)."""
    rng = np.random.default_rng(seed)
    n_rect = 1500
    for _ in range(5):
        rng.integers(0, 2, n_rect)   # rw, rh, x0, y0, g: five draws of n_rect integers precede the depths
    return rng.uniform(5.0, 80.0, grid)


def _quat_to_R_np(q):
    x, y, z, w = q / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def pose_only_frame(seed, n_points=250, pix_noise=0.7, outlier_frac=0.1, pose_noise=(0.01, 0.15)):
    """One tracked frame for Frontend::EstimateCurrentPose: map points seen by a camera, noisy pixel observations
    (float like cv::KeyPoint::pt), a fraction of gross outliers (LK mis-tracks), a perturbed starting pose.
    Returns dict: pose0 [7], pose_gt [7], points [n,3], uv [n,2], planted [n] bool."""
    w = ba_window(seed, n_poses=1, n_points=n_points, pix_noise=pix_noise, outlier_frac=0.0, pose_noise=pose_noise,
                  point_noise=0.0, fixed_frac=1.0)
    rng = np.random.default_rng(seed + 9999)
    pts = w["points_gt"][w["obs_point"]]
    uv = w["uv"].copy()
    planted = rng.uniform(size=len(uv)) < outlier_frac
    uv[planted] += rng.uniform(-40, 40, (int(planted.sum()), 2))
    uv = uv.astype(np.float32).astype(np.float64)
    return {"pose0": w["poses0"][0], "pose_gt": w["poses_gt"][0], "points": pts, "uv": uv, "planted": planted}


def pnp_problem(seed, n_points=400, pix_noise=0.7, outlier_frac=0.3):
    """One loop candidate for LoopClosing::ComputeCorrectPose (src/loopclosing.cpp:207-268): map points of the loop
    keyframe (float32 like cv::Point3f) and their matched keypoints in the current keyframe (float32 pixels) under an
    unknown pose; `outlier_frac` of the matches are wrong (uniformly random pixels).
    Returns dict: obj [n,3] f32, img [n,2] f32, pose_gt [7], planted [n] bool."""
    rng = np.random.default_rng(50000 + seed)
    R = _so3_exp(rng.normal(size=3) * 0.3)
    t = rng.normal(size=3) * np.array([2.0, 0.5, 2.0])
    fx, fy, cx, cy = KITTI_K
    pts = []
    while len(pts) < n_points:                      # points in front of the camera that project inside the image
        pc = np.array([rng.uniform(-25, 25), rng.uniform(-6, 6), rng.uniform(4, 60)])
        u, v = fx * pc[0] / pc[2] + cx, fy * pc[1] / pc[2] + cy
        if 0 <= u < KITTI_W and 0 <= v < KITTI_H:
            pts.append(R.T @ (pc - t))
    obj = np.array(pts)
    pc = obj @ R.T + t
    img = np.stack([fx * pc[:, 0] / pc[:, 2] + cx, fy * pc[:, 1] / pc[:, 2] + cy], 1) + rng.normal(size=(n_points, 2)) * pix_noise
    planted = rng.uniform(size=n_points) < outlier_frac
    img[planted] = np.stack([rng.uniform(0, KITTI_W, int(planted.sum())), rng.uniform(0, KITTI_H, int(planted.sum()))], 1)
    return {"obj": obj.astype(np.float32), "img": img.astype(np.float32), "pose_gt": pose7(R, t), "planted": planted}


# ---------------------------------------------------------------------------------------------------
# DeepLCD CNN ("next" row 2).  The trained calc.caffemodel is a configure-time download of the reference
# (get_model.sh) and absent: tests and bench use seeded random weights of the same architecture.
CALC_CONVS = ((64, 1, 5), (128, 64, 4), (4, 128, 3))  # (Cout, Cin, k) of the three Convolution layers


def calc_weights(seed=0, convs=CALC_CONVS):
    """Flat fp32 buffer in Caffe blob order: per Convolution layer W [Cout][Cin][k][k] then bias [Cout]."""
    rng = np.random.default_rng(seed)
    out = []
    for co, ci, k in convs:
        out.append((rng.standard_normal(co * ci * k * k) * np.sqrt(2.0 / (ci * k * k))).astype(np.float32))
        out.append((rng.standard_normal(co) * 0.05).astype(np.float32))
    return np.concatenate(out)
