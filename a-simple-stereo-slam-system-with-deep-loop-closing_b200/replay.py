"""BASELINE configs 4 and 5 as one replay: the full pipeline over a KITTI-00-sized synthetic sequence.

What runs (reference flow, paths relative to the reference root):
  every frame   ORBextractor::DetectAndCompute on both views + BF-Hamming match        sb_stereo_submit / wait
  every KF      Backend::OptimizeActiveMap's solve (src/backend.cpp:126-269)            sb_ba_submit / wait
                front end at a keyframe (src/frontend.cpp:302-328, :335-379, :451-488):  sb_orb_detect, sb_lk_track,
                  Detect -> LK into the right view -> triangulation                        sb_triangulate
                LoopClosing::ProcessNewKF (src/loopclosing.cpp:81-121): whole-image       sb_calc_descr_original,
                  descriptor, 8-octave expansion, screening, ORB descriptors               sb_orb_screen_params, sb_orb_calc_descriptors
  in KF order   DetectLoop (:124-161) -> MatchFeatures (:167-203) -> ComputeCorrectPose   sb_lcd_detect_loop, sb_hamming_match,
                  (:208-335) -> LoopLocalFusion (poses, :463-476) -> PoseGraphOptimization   sb_pnp_ransac, sb_pose_solve,
                  (:537-646) -> AddToDatabase (:651-659), incl. the 5-keyframe hold-off (:671-680)   sb_posegraph_solve, sb_lcd_add

Multi-GPU (config 5): frames and keyframes are owned round-robin; the per-frame and per-keyframe stages need nothing from
other ranks.  Before the sequential loop-closing stage the keyframe poses are all-gathered through the C ABI
(sb_allgather_kf_poses) and the per-keyframe records (descriptors, features, landmarks) through torch.distributed; every rank
then runs the loop-closing stage redundantly — it is deterministic, so all ranks hold identical results.

Inputs are synthetic and seeded (synth.py): keyframe k shows scene 100000 + k; a keyframe that revisits a place shows its
partner's scene under fresh sensor noise and has its partner's true pose; the other frames cycle through a pool of scenes.
This module is harness (Python over ctypes, like capi.py): all arithmetic happens behind include/slamb200.h.
"""
import hashlib
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import capi, synth
from . import parallel as par

ORB_PARAMS = (2000, 1.2, 8, 20, 7)
_SYNC_TIMERS = bool(int(__import__("os").environ.get("REPLAY_SYNC_TIMERS", "0")))
KF_SCENE0 = 100000


# ---------------------------------------------------------------------------------------------------------------------
# SE(3) helpers on 4x4 matrices (host bookkeeping of the replay only; the solvers take / return pose7)
# ---------------------------------------------------------------------------------------------------------------------
def T_from7(p):
    T = np.eye(4)
    T[:3, :3] = synth._quat_to_R_np(np.asarray(p[:4], np.float64))
    T[:3, 3] = p[4:7]
    return T


def T_to7(T):
    return synth.pose7(T[:3, :3], T[:3, 3])


def T_to7_batch(Ts):
    """T_to7 for a stack of poses [n, 4, 4] -> [n, 7]: the trace > 0 rows are vectorised, the rest take the scalar path."""
    Ts = np.asarray(Ts)
    R, out = Ts[:, :3, :3], np.empty((len(Ts), 7))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    pos = tr > 0
    s_ = np.sqrt(np.where(pos, tr, 0.0) + 1.0) * 2
    q = np.stack([(R[:, 2, 1] - R[:, 1, 2]) / s_, (R[:, 0, 2] - R[:, 2, 0]) / s_, (R[:, 1, 0] - R[:, 0, 1]) / s_, 0.25 * s_], 1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)       # q[3] = s / 4 > 0: no sign flip
    out[:, :4] = q
    out[:, 4:] = Ts[:, :3, 3]
    for i in np.nonzero(~pos)[0]:
        out[i] = T_to7(Ts[i])
    return out


def T_from7_batch(p):
    """T_from7 for [n, 7] -> [n, 4, 4]."""
    p = np.asarray(p, np.float64)
    q = p[:, :4] / np.linalg.norm(p[:, :4], axis=1, keepdims=True)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    T = np.zeros((len(p), 4, 4))
    T[:, 0, 0] = 1 - 2 * (y * y + z * z); T[:, 0, 1] = 2 * (x * y - z * w); T[:, 0, 2] = 2 * (x * z + y * w)
    T[:, 1, 0] = 2 * (x * y + z * w); T[:, 1, 1] = 1 - 2 * (x * x + z * z); T[:, 1, 2] = 2 * (y * z - x * w)
    T[:, 2, 0] = 2 * (x * z - y * w); T[:, 2, 1] = 2 * (y * z + x * w); T[:, 2, 2] = 1 - 2 * (x * x + y * y)
    T[:, :3, 3] = p[:, 4:7]
    T[:, 3, 3] = 1.0
    return T


def T_inv(T):
    R, t = T[:3, :3], T[:3, 3]
    out = np.eye(4)
    out[:3, :3] = R.T
    out[:3, 3] = -R.T @ t
    return out


def se3_log_norm(T):
    """|log(T)| with Sophus' tangent (translation part V^-1 t, rotation part omega)."""
    R, t = T[:3, :3], T[:3, 3]
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-9:
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    else:
        w = th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    th = np.linalg.norm(w)
    W = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        Vinv = np.eye(3) - 0.5 * W
    else:
        Vinv = np.eye(3) - 0.5 * W + (1 - th * np.cos(th / 2) / (2 * np.sin(th / 2))) / (th * th) * (W @ W)
    return float(np.linalg.norm(np.concatenate([Vinv @ t, w])))


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


# ---------------------------------------------------------------------------------------------------------------------
# the synthetic sequence
# ---------------------------------------------------------------------------------------------------------------------
class Sequence:
    """frames stereo frames, n_kf keyframes (keyframe k = frame floor(k * frames / n_kf)), planted revisits."""

    def __init__(self, frames=4541, n_kf=742, loop_pairs=None, seed=0, pool=192, odo_noise=(0.0008, 0.03), radius=300.0):
        self.frames, self.n_kf, self.pool = frames, n_kf, pool
        self._pool_cache = {}
        if loop_pairs is None:
            loop_pairs = synth.LOOP_PAIRS if n_kf >= 742 else self.default_pairs(n_kf)
        self.loop_pairs = [(a, b) for a, b in loop_pairs if b < a < n_kf]
        self.partner = dict(self.loop_pairs)
        self.kf_frame = (np.arange(n_kf) * frames) // n_kf
        self.kf_of_frame = {int(f): k for k, f in enumerate(self.kf_frame)}
        rng = np.random.default_rng(seed)
        per_rev = max(8, int(round(n_kf / 1.97)))
        Ts = []
        for k in range(n_kf):
            j = self.partner.get(k, k)                      # a revisit stands exactly where its partner stood
            ang = 2 * np.pi * j / per_rev
            c = np.array([radius * np.sin(ang), 0.02 * j, radius * (1 - np.cos(ang))])
            Rwc = synth._rot_y(ang)
            T = np.eye(4)
            T[:3, :3] = Rwc.T
            T[:3, 3] = -Rwc.T @ c
            Ts.append(T)
        self.T_gt = Ts
        # odometry: T_k T_{k-1}^-1 with noise on the left (KeyFrame::mRelativePoseToLastKF)
        self.odo = [None]
        for k in range(1, n_kf):
            D = Ts[k] @ T_inv(Ts[k - 1])
            N = np.eye(4)
            N[:3, :3] = synth._so3_exp(rng.normal(0, odo_noise[0], 3))
            N[:3, 3] = rng.normal(0, odo_noise[1], 3)
            self.odo.append(N @ D)

    @staticmethod
    def default_pairs(n_kf):
        """Short sequences: revisits every 7 keyframes in the last 40 % of the run, each ~half the run back."""
        out = []
        a = int(0.6 * n_kf)
        while a < n_kf - 1:
            out.append((a, a - n_kf // 2))
            a += 7
        return out

    def frame_images(self, f):
        """(left, right) of frame f."""
        k = self.kf_of_frame.get(int(f))
        if k is None:
            s = int(f) % self.pool
            if s not in self._pool_cache:                   # the pool scenes repeat: generate each once
                self._pool_cache[s] = synth.stereo_pair(s)
            return self._pool_cache[s]
        if k in self.partner:
            return synth.stereo_pair(KF_SCENE0 + self.partner[k], noise_seed=7_000_000 + k)
        return synth.stereo_pair(KF_SCENE0 + k)

    def load(self, frame_ids, threads=None):
        """[n, 2, H, W] u8 of the given frames (generated in a thread pool; the replay's "dataset read")."""
        out = np.empty((len(frame_ids), 2, synth.KITTI_H, synth.KITTI_W), np.uint8)

        def one(i):
            out[i, 0], out[i, 1] = self.frame_images(frame_ids[i])
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(one, range(len(frame_ids))))
        return out


class KittiSequence(Sequence):
    """A real KITTI odometry sequence directory in the layout the reference's app reads (app/run_kitti_stereo.cpp:114-144):
    <dir>/image_0/%06d.png (left), <dir>/image_1/%06d.png (right), <dir>/times.txt.  Keyframes: every `kf_every`-th frame.
    Odometry edges come from `poses_file` (KITTI ground truth, 12 numbers of T_w_cam0 per line) perturbed by the same drift
    model as the synthetic route; without it the keyframe chain is the identity and only detection + verification are
    meaningful.  Images are cropped / zero-padded to 1241 x 376 (KITTI-00 needs neither)."""

    def __init__(self, path, poses_file=None, max_frames=None, kf_every=6.12, seed=0, odo_noise=(0.0008, 0.03)):
        import os
        self.path = path
        times = [l for l in open(os.path.join(path, "times.txt")).read().split("\n") if l.strip()]
        n = len(times) if not max_frames else min(len(times), max_frames)
        self.frames, self.n_kf, self.pool = n, max(2, int(round(n / kf_every))), 0
        self.loop_pairs, self.partner = [], {}
        self.kf_frame = (np.arange(self.n_kf) * self.frames) // self.n_kf
        self.kf_of_frame = {int(f): k for k, f in enumerate(self.kf_frame)}
        rng = np.random.default_rng(seed)
        Ts = [np.eye(4) for _ in range(self.n_kf)]
        if poses_file:
            rows = np.loadtxt(poses_file).reshape(-1, 3, 4)
            for k, f in enumerate(self.kf_frame):
                Twc = np.eye(4)
                Twc[:3] = rows[int(f)]
                Ts[k] = T_inv(Twc)
        self.T_gt = Ts
        self.odo = [None]
        for k in range(1, self.n_kf):
            N = np.eye(4)
            N[:3, :3] = synth._so3_exp(rng.normal(0, odo_noise[0], 3))
            N[:3, 3] = rng.normal(0, odo_noise[1], 3)
            self.odo.append(N @ (Ts[k] @ T_inv(Ts[k - 1])))

    def frame_images(self, f):
        import os
        import cv2
        out = []
        for cam in ("image_0", "image_1"):
            img = cv2.imread(os.path.join(self.path, cam, "%06d.png" % int(f)), cv2.IMREAD_GRAYSCALE)
            if img is None:
                raise FileNotFoundError(os.path.join(self.path, cam, "%06d.png" % int(f)))
            canvas = np.zeros((synth.KITTI_H, synth.KITTI_W), np.uint8)
            h, w = min(img.shape[0], synth.KITTI_H), min(img.shape[1], synth.KITTI_W)
            canvas[:h, :w] = img[:h, :w]
            out.append(canvas)
        return out[0], out[1]


# ---------------------------------------------------------------------------------------------------------------------
# the operators of the pipeline (GPU: the C ABI).  tests/ substitutes a CPU checker with the same methods.
# ---------------------------------------------------------------------------------------------------------------------
class GpuOps:
    def __init__(self, device=0, batch=64, kf_features=300, kf_batch=32, n_kf=742, lcd_dtype=0, pin=True):
        self.device, self.batch, self.kf_batch, self._lcd_dtype = device, batch, kf_batch, lcd_dtype
        self.slots = 3                                    # front-end batches in flight (copy in, kernels, copy out overlap)
        self.fe = [capi.StereoFrontend(*ORB_PARAMS, max_pairs=batch, device=device) for _ in range(self.slots)]
        self.fe_out = [f.alloc_outputs(batch, pinned=pin) for f in self.fe]
        self.ba = capi.LocalBA(max_windows=batch, max_poses=7, max_points=320, max_obs=2304, device=device)
        self.kf_ext = capi.ORBextractor(kf_features, 1.2, 8, 20, 7, max_batch=kf_batch, device=device)
        self.lk = capi.LKTracker(max_batch=kf_batch, max_pts=kf_features + 64, device=device)
        self.net = capi.DeepLCD(synth.calc_weights(0), max_batch=kf_batch, device=device)
        self.lcd = capi.DeepLCDScorer(capacity=max(n_kf, 64), dtype=lcd_dtype, device=device)
        self.matcher = capi.HammingMatcher(max_batch=1, max_rows=8 * (kf_features + 64), device=device)
        self.pnp = capi.PnPRansac(max_problems=1, max_points=4096, device=device)
        self.refine = capi.PoseOnlyOptimizer(max_frames=1, max_obs=4096, device=device)
        self.pg = capi.PoseGraph(max(n_kf, 64) + 8, 2 * max(n_kf, 64) + 64, device=device)

    def reset(self):
        """Empty the keyframe database (a new sequence on the same handles)."""
        cap, dt, dev = self.lcd.capacity, self._lcd_dtype, self.device
        self.lcd.close()
        self.lcd = capi.DeepLCDScorer(capacity=cap, dtype=dt, device=dev)

    # per-frame stage ------------------------------------------------------------------------------------------------
    def stereo_submit(self, slot, images):
        self.fe[slot].submit(images, self.fe_out[slot])

    def stereo_wait(self, slot):
        self.fe[slot].wait()
        return self.fe_out[slot]

    def ba_submit(self, windows):
        self.ba.submit(windows, synth.KITTI_K)

    def ba_wait(self):
        return self.ba.wait()

    # per-keyframe stage ---------------------------------------------------------------------------------------------
    def kf_detect(self, lefts):
        return self.kf_ext.DetectBatch(lefts)

    def lk_right(self, lefts, rights, pts):
        return self.lk.track(lefts, rights, pts, next_pts0=pts)

    def triangulate(self, ul, ur, T_wc7):
        b = synth.KITTI_BF / synth.KITTI_FX
        return capi.triangulate(ul, ur, synth.KITTI_K, synth.KITTI_K, np.array([0, 0, 0, 1, 0, 0, 0.0]),
                                np.array([0, 0, 0, 1, -b, 0, 0.0]), T_wc7, device=self.device)

    def triangulate_batch(self, uls, urs):
        """One sb_triangulate call for the correspondences of a whole batch of keyframes (same two camera poses)."""
        n = [len(u) for u in uls]
        p, ok = self.triangulate(np.concatenate(uls), np.concatenate(urs), None)
        o = np.concatenate([[0], np.cumsum(n)])
        return [(p[o[i]:o[i + 1]], ok[o[i]:o[i + 1]]) for i in range(len(n))]

    def cnn_descr(self, lefts):
        """DeepLCD::calcDescrOriginalImg: returns descriptors; `lefts` come back blurred in place (quirk Q8)."""
        return self.net.calcDescrOriginalImgBatch(lefts, in_place=True)

    def screen_and_describe(self, imgs, feats):
        """src/loopclosing.cpp:94-113 for a batch of keyframes: every feature as a keypoint on each octave, then
        ScreenAndComputeKPsParams + CalcDescriptors (sb_orb_screen_describe) -> [(kps, desc)]."""
        kin, n_in = expand_octaves_batch(feats)
        return [(kout, desc) for _, kout, desc in self.kf_ext.ScreenAndDescribeBatch(imgs, kin, n_in=n_in, copy=False)]

    # loop-closing stage ---------------------------------------------------------------------------------------------
    def lcd_add(self, kf_id, d):
        self.lcd.add(kf_id, d)

    def lcd_size(self):
        return len(self.lcd)

    def lcd_detect(self, kf_id, d, min_gap):
        return self.lcd.DetectLoop(kf_id, d, 0.94, 0.92, min_gap, 3)

    def match(self, q, t):
        return self.matcher.match(q, t)

    def pnp_ransac(self, obj, img, seed):
        return self.pnp.solve([(obj, img)], synth.KITTI_K, 100, 5.991, seed)[0]

    def pose_refine(self, pose7, pts, uv):
        return self.refine.solve([dict(pose0=pose7, points=pts, uv=uv)], synth.KITTI_K, pre_rounds=1)[0]

    def posegraph(self, poses, fixed, v0, v1, meas):
        return self.pg.solve(poses, fixed, v0, v1, meas)


def match_filter(idx, dist, loop_class_id, cur_class_id):
    """LoopClosing::MatchFeatures (src/loopclosing.cpp:175-194): keep distance <= max(2 * min, 30), map keypoints to
    features through class_id, de-duplicate; returned in std::set<pair<cur, loop>> order."""
    if len(idx) == 0:
        return []
    valid = dist >= 0
    if not valid.any():
        return []
    th = max(2.0 * float(dist[valid].min()), 30.0)
    s = set()
    for q in np.nonzero(valid & (dist <= th))[0]:
        s.add((int(cur_class_id[idx[q]]), int(loop_class_id[q])))
    return sorted(s)


# ---------------------------------------------------------------------------------------------------------------------
# the replay
# ---------------------------------------------------------------------------------------------------------------------
def expand_octaves(feats, nlevels=8):
    """src/loopclosing.cpp:94-105: every feature as a keypoint on each octave, response -1, class_id = feature index."""
    kin = np.zeros(len(feats) * nlevels, capi.KP_DTYPE)
    rep = np.repeat(np.arange(len(feats)), nlevels)
    for name in ("x", "y", "size", "angle"):
        kin[name] = feats[name][rep]
    kin["response"] = -1
    kin["octave"] = np.tile(np.arange(nlevels), len(feats))
    kin["class_id"] = rep
    return kin


def expand_octaves_batch(feats_list, nlevels=8):
    """expand_octaves for a batch of keyframes, packed as the C ABI takes it: ([B, cap_in] keypoints, [B] counts)."""
    n = np.array([len(f) for f in feats_list], np.int32) * nlevels
    kin = np.zeros((len(feats_list), max(1, int(n.max()))), capi.KP_DTYPE)
    for b, f in enumerate(feats_list):
        row = kin[b, :n[b]].reshape(len(f), nlevels)
        for name in ("x", "y", "size", "angle"):
            row[name] = f[name][:, None]
        row["response"] = -1
        row["octave"] = np.arange(nlevels)[None, :]
        row["class_id"] = np.arange(len(f))[:, None]
    return kin, n


def lap(t, key, t0):
    if _SYNC_TIMERS:   # diagnosis only: attribute asynchronous device work to the stage that enqueued it
        import torch
        torch.cuda.synchronize()
    now = time.perf_counter()
    t[key] = t.get(key, 0.0) + now - t0
    return now


def run(seq, ops, rank=0, world=1, db_min_size=50, min_gap=20, with_digests=False, comm=None, log=None):
    """Runs the whole replay on this rank's share; returns a dict (identical on every rank except timings)."""
    B = ops.batch
    t = {}
    my_frames = par.shard_indices(seq.frames, rank, world)
    my_kfs = par.shard_indices(seq.n_kf, rank, world)

    # ---- inputs ("dataset read", outside the timed region)
    t0 = time.perf_counter()
    frames_np = seq.load(my_frames)
    kf_imgs = seq.load(seq.kf_frame[my_kfs])
    windows = [synth.ba_window(int(k)) for k in my_kfs]
    pin = None
    try:   # page-locked host buffers: the H2D copies of every operator run at PCIe speed and asynchronously
        import torch
        if torch.cuda.is_available():
            pin = (torch.from_numpy(frames_np).pin_memory(), torch.from_numpy(kf_imgs).pin_memory())
            frames_np, kf_imgs = pin[0].numpy(), pin[1].numpy()
    except Exception:
        pass
    t["load_s"] = time.perf_counter() - t0

    t_start = time.perf_counter()
    # ---- stage A: every frame through extract (both views) + match; one BA window per keyframe, in the background
    frame_digest = {}
    n_kps = n_matches = 0
    nb = (len(my_frames) + B - 1) // B
    NS = getattr(ops, "slots", 2)
    ba_results, wb = [], 0

    def collect(slot, b):
        nonlocal n_kps, n_matches
        out = ops.stereo_wait(slot)
        lo = b * B
        n = min(B, len(my_frames) - lo)
        n_kps += int(out["counts"][:n].sum())
        live = np.arange(out["mdist"].shape[1])[None, :] < out["counts"][:n, 0, None]      # query rows of every frame
        n_matches += int(((out["mdist"][:n] >= 0) & live).sum())
        if with_digests:
            for i in range(n):
                cl, cr = int(out["counts"][i, 0]), int(out["counts"][i, 1])
                frame_digest[int(my_frames[lo + i])] = digest(out["kps"][i, 0, :cl], out["kps"][i, 1, :cr], out["desc"][i, 0, :cl],
                                                              out["desc"][i, 1, :cr], out["midx"][i, :cl], out["mdist"][i, :cl])

    ba_pending = False
    for b in range(nb):
        slot = b % NS
        if b >= NS:
            collect(slot, b - NS)
        if wb < len(windows):                           # the back end works beside the front end (src/backend.cpp:29-45)
            if ba_pending:
                ba_results += ops.ba_wait()
            ops.ba_submit(windows[wb:wb + B])
            wb += B
            ba_pending = True
        lo = b * B
        batch = frames_np[lo:lo + B]
        ops.stereo_submit(slot, batch)
    for b in range(max(0, nb - NS), nb):
        collect(b % NS, b)
    while True:
        if ba_pending:
            ba_results += ops.ba_wait()
            ba_pending = False
        if wb >= len(windows):
            break
        ops.ba_submit(windows[wb:wb + B])
        wb += B
        ba_pending = True
    t["frames_s"] = time.perf_counter() - t_start

    # ---- stage B: keyframe processing (front end at a keyframe + LoopClosing::ProcessNewKF) for my keyframes
    t1 = time.perf_counter()
    KB = ops.kf_batch
    rec = {}
    for lo in range(0, len(my_kfs), KB):
        ks = my_kfs[lo:lo + KB]
        lefts = [kf_imgs[lo + i, 0] for i in range(len(ks))]
        rights = [kf_imgs[lo + i, 1] for i in range(len(ks))]
        tk = time.perf_counter()
        feats = ops.kf_detect(lefts)
        pts = [np.stack([f["x"], f["y"]], 1).astype(np.float32) for f in feats]
        t.setdefault("kf_detect_batches_ms", []).append(round(1e3 * (time.perf_counter() - tk), 3))
        tk = lap(t, "kf_detect_s", tk)
        tracked = ops.lk_right(lefts, rights, pts)
        tk = lap(t, "kf_lk_s", tk)
        descr = ops.cnn_descr(lefts)                    # blurs `lefts` in place: the ORB descriptors below see the blurred image
        tk = lap(t, "kf_cnn_s", tk)
        described = ops.screen_and_describe(lefts, feats)
        tk = lap(t, "kf_screen_describe_s", tk)
        tri = ops.triangulate_batch(pts, [tr[0] for tr in tracked])
        tk = lap(t, "kf_triangulate_s", tk)
        for i, k in enumerate(ks):
            ur, status = tracked[i]
            p_cam, ok = tri[i]
            ok = ok & (status != 0)
            kout, desc = described[i]
            rec[int(k)] = dict(feats=feats[i], p_cam=p_cam.astype(np.float64), has_mp=ok, descr=descr[i].copy(), pyr=kout, orb=desc)
    t["keyframes_s"] = time.perf_counter() - t1

    # ---- exchange: keyframe poses through the C ABI collective, records through torch.distributed
    t2 = time.perf_counter()
    poses_ba = np.zeros((len(my_kfs), 7))
    for i in range(len(my_kfs)):
        poses_ba[i] = ba_results[i][0][-1]              # the newest pose of keyframe k's window after BA (replayed windows)
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(ops.device)               # the C ABI selects its handle's device per call; make torch's explicit too
        poses_all = par.allgather_kf_poses(poses_ba, seq.n_kf, rank, world, device=torch.device("cuda", ops.device))
        import pickle
        blobs = par.allgather_blobs(pickle.dumps((rec, frame_digest, n_kps, n_matches), protocol=pickle.HIGHEST_PROTOCOL),
                                    device=torch.device("cuda", ops.device) if dist.get_backend() == "nccl" else None)
        rec, frame_digest = {}, {}
        n_kps = n_matches = 0
        for blob in blobs:
            r_, fd, a, b_ = pickle.loads(blob)
            rec.update(r_)
            frame_digest.update(fd)
            n_kps += a
            n_matches += b_
    else:
        poses_all = poses_ba
    t["exchange_s"] = time.perf_counter() - t2

    # ---- stage C: loop closing over the keyframes in order (every rank, redundantly)
    t3 = time.perf_counter()
    loops, est, pg_runs = [], [None] * seq.n_kf, 0
    last_closed = None
    loop_edge = {}                                        # kf -> (loop kf, T_rel 4x4)
    t_pg = 0.0
    for k in range(seq.n_kf):
        est[k] = seq.T_gt[0].copy() if k == 0 else seq.odo[k] @ est[k - 1]
        if last_closed is not None and k - last_closed <= 5:     # InsertNewKeyFrame hold-off (:671-680)
            continue
        r = rec[k]
        confirmed = False
        if ops.lcd_size() > db_min_size:
            found, best, score, cnt = ops.lcd_detect(k, r["descr"], min_gap)
            if found:
                L = rec[int(best)]
                ok, info = verify_loop(ops, r, L, est[k], est[int(best)], seed=k)
                if log:
                    log(f"kf {k}: candidate {best} score {score:.4f} suspected {cnt} -> {info}")
                if ok:
                    confirmed = True
                    T_corr = T_from7(info["pose"])
                    loop_edge[k] = (int(best), T_corr @ T_inv(est[int(best)]))
                    last_closed = k
                    need = se3_log_norm(est[k] @ T_inv(T_corr)) > 1.0
                    loops.append([k, int(best), bool(need), int(info["inliers"])])
                    if need:
                        tp = time.perf_counter()
                        correct_and_optimise(ops, seq, est, k, int(best), T_corr, loop_edge)
                        t_pg += time.perf_counter() - tp
                        pg_runs += 1
        if not confirmed:
            ops.lcd_add(k, r["descr"])
    t["loop_closing_s"] = time.perf_counter() - t3
    t["posegraph_s"] = t_pg
    t["total_s"] = time.perf_counter() - t_start
    final = np.stack([T_to7(T) for T in est])
    err0 = float(np.mean([np.linalg.norm(T_inv(seq.T_gt[k])[:3, 3] - T_inv(dead)[:3, 3]) for k, dead in enumerate(dead_reckon(seq))]))
    err1 = float(np.mean([np.linalg.norm(T_inv(seq.T_gt[k])[:3, 3] - T_inv(est[k])[:3, 3]) for k in range(seq.n_kf)]))
    out = dict(world=world, frames=seq.frames, keyframes=seq.n_kf, planted=[list(p) for p in seq.loop_pairs], loops=loops,
               posegraph_runs=pg_runs, keypoints=int(n_kps), matches=int(n_matches), timings=t,
               posegraph_digest=digest(np.round(final, 9)), kf_pose_digest=digest(poses_all),
               kf_record_digest=digest(*[rec[k]["orb"] for k in range(seq.n_kf)], *[rec[k]["descr"] for k in range(seq.n_kf)]),
               mean_position_error_dead_reckoned_m=err0, mean_position_error_final_m=err1)
    if with_digests:
        out["frame_digests"] = [frame_digest[f] for f in range(seq.frames)]
    return out


def dead_reckon(seq):
    T = [seq.T_gt[0].copy()]
    for k in range(1, seq.n_kf):
        T.append(seq.odo[k] @ T[-1])
    return T


def verify_loop(ops, cur, loop, T_cur, T_loop, seed=0):
    """MatchFeatures + ComputeCorrectPose (src/loopclosing.cpp:167-335) -> (accepted, info)."""
    if len(loop["orb"]) == 0 or len(cur["orb"]) == 0:
        return False, dict(reason="no descriptors")
    idx, dist = ops.match(loop["orb"], cur["orb"])                  # query = LOOP keyframe, train = CURRENT (quirk Q13)
    pairs = match_filter(idx, dist, loop["pyr"]["class_id"], cur["pyr"]["class_id"])
    if len(pairs) < 10:
        return False, dict(reason="matches", n=len(pairs))
    pairs = [(c, l) for c, l in pairs if loop["has_mp"][l]]          # only loop features linked to a map point (:221-240)
    if len(pairs) < 10:
        return False, dict(reason="map points", n=len(pairs))
    ci = np.array([c for c, _ in pairs])
    li = np.array([l for _, l in pairs])
    T_wl = T_inv(T_loop)
    obj = (loop["p_cam"][li] @ T_wl[:3, :3].T + T_wl[:3, 3]).astype(np.float32)       # MapPoint::Pos() of the loop keyframe's features
    img = np.stack([cur["feats"]["x"][ci], cur["feats"]["y"][ci]], 1).astype(np.float32)
    sol = ops.pnp_ransac(obj, img, seed)
    if not sol["found"]:
        return False, dict(reason="pnp")
    pose, outl, info = ops.pose_refine(sol["pose7"], obj.astype(np.float64), img.astype(np.float64))
    inliers = int(info[0])
    if inliers < 10:
        return False, dict(reason="refine", n=inliers)
    return True, dict(pose=pose, inliers=inliers, matches=len(pairs))


def correct_and_optimise(ops, seq, est, k, loop_kf, T_corr, loop_edge, n_active=7):
    """LoopLocalFusion's pose part (:463-476) + PoseGraphOptimization (:537-646) over keyframes 0..k."""
    active = list(range(max(0, k - n_active + 1), k + 1))
    T_cur = est[k].copy()
    for a in active:
        est[a] = (est[a] @ T_inv(T_cur)) @ T_corr                  # Tac * corrected current pose
    n = k + 1
    fixed = np.zeros(n, np.uint8)
    fixed[active] = 1
    fixed[loop_kf] = 1
    fixed[0] = 1
    v0, v1, meas = [], [], []
    odo7 = getattr(seq, "_odo7", None)                             # the odometry measurements as pose7, converted once
    if odo7 is None:
        odo7 = seq._odo7 = [T_to7(T) if i > 0 else None for i, T in enumerate(seq.odo)]
    for i in range(n):                                             # allKFs in id order: sequential edge, then loop edge
        if i > 0:
            v0.append(i); v1.append(i - 1); meas.append(odo7[i])
        if i in loop_edge:
            v0.append(i); v1.append(loop_edge[i][0]); meas.append(T_to7(loop_edge[i][1]))
    poses = T_to7_batch(np.stack(est[:n]))
    new, _ = ops.posegraph(poses, fixed, np.array(v0, np.int32), np.array(v1, np.int32), np.array(meas))
    Tn = T_from7_batch(new)
    for i in range(n):
        est[i] = Tn[i]
