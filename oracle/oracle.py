"""ctypes front-end of the CPU checker (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — see the header of oracle/orb_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_i32p = C.POINTER(C.c_int32)
_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".inc"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_fast_atan2.restype = C.c_float
        L.orc_fast_atan2.argtypes = [C.c_float, C.c_float]
        for name in ("orc_detect_and_compute", "orc_detect_with_pyramid", "orc_detect", "orc_screen_params",
                     "orc_calc_descriptors", "orc_get_level", "orc_get_blurred_level", "orc_resize_linear_u8",
                     "orc_gauss7_u8", "orc_fast9_16", "orc_hamming_match", "orc_match_filter",
                     "orc_distribute_octtree", "orc_debug_candidate_count", "orc_ba_solve"):
            getattr(L, name).restype = C.c_int
    return _LIB


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    rc = lib().orc_resize_linear_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dw, dh, dw)
    assert rc == 0
    return dst


def gauss7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty_like(src)
    rc = lib().orc_gauss7_u8(_p(src), src.shape[1], src.shape[0], src.strides[0], _p(dst), dst.strides[0])
    assert rc == 0
    return dst


def fast_atan2(y, x):
    return float(lib().orc_fast_atan2(C.c_float(y), C.c_float(x)))


def fast9_16(img, threshold, nonmax=True):
    """img may be a non-contiguous 2-D view (ROI).  Returns int array [n,3] = (x, y, response), row-major."""
    assert img.dtype == np.uint8 and img.strides[1] == 1
    h, w = img.shape
    cap = max(1, w * h)
    out = np.empty((cap, 3), np.int32)
    n = lib().orc_fast9_16(C.c_void_p(img.ctypes.data), w, h, img.strides[0], int(threshold), int(nonmax), _p(out),
                           cap)
    return out[:n].copy()


def distribute_octtree(kx, ky, kr, minX, maxX, minY, maxY, N):
    kx = np.ascontiguousarray(kx, np.float32)
    ky = np.ascontiguousarray(ky, np.float32)
    kr = np.ascontiguousarray(kr, np.float32)
    out = np.empty(max(1, len(kx)), np.int32)
    n = lib().orc_distribute_octtree(_p(kx), _p(ky), _p(kr), len(kx), minX, maxX, minY, maxY, N, _p(out))
    return out[:n].copy()


def hamming_match(q, t):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty(len(q), np.int32)
    dist = np.empty(len(q), np.int32)
    lib().orc_hamming_match(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist))
    return idx, dist


def match_filter(train_idx, dist, loop_class_id, cur_class_id):
    train_idx = np.ascontiguousarray(train_idx, np.int32)
    dist = np.ascontiguousarray(dist, np.int32)
    lc = np.ascontiguousarray(loop_class_id, np.int32)
    cc = np.ascontiguousarray(cur_class_id, np.int32)
    pairs = np.empty((max(1, len(train_idx)), 2), np.int32)
    n = lib().orc_match_filter(_p(train_idx), _p(dist), len(train_idx), _p(lc), _p(cc), _p(pairs))
    return pairs[:n].copy()


class ORBextractor:
    """Mirror of myslam::ORBextractor (include/myslam/ORBextractor.h:47-138) over the C restatement."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST):
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = C.c_void_p(lib().orc_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST))
        assert self._h
        n = nlevels
        self.scale = np.zeros(n, np.float32)
        self.inv_scale = np.zeros(n, np.float32)
        self.sigma2 = np.zeros(n, np.float32)
        self.inv_sigma2 = np.zeros(n, np.float32)
        self.quota = np.zeros(n, np.int32)
        self.umax = np.zeros(16, np.int32)
        lib().orc_get_tables(self._h, _p(self.scale), _p(self.inv_scale), _p(self.sigma2), _p(self.inv_sigma2),
                             _p(self.quota), _p(self.umax))
        self.cap = nfeatures + 67 * nlevels + 600  # each level may keep max(quota + 3, 4 * nIni) keypoints

    def __del__(self):
        try:
            lib().orc_destroy(self._h)
        except Exception:
            pass

    @staticmethod
    def _img(a):
        a = np.ascontiguousarray(a, np.uint8)
        assert a.ndim == 2
        return a

    def DetectAndCompute(self, image, mask=None, with_stats=False):
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        lc = np.zeros(self.nlevels, np.int32)
        cands = np.zeros(self.nlevels, np.int32)
        n = lib().orc_detect_and_compute(self._h, _p(image), _p(mask), image.shape[1], image.shape[0],
                                         image.strides[0], mask.strides[0] if mask is not None else 0, _p(kps),
                                         _p(desc), self.cap, _p(lc), _p(cands))
        assert n <= self.cap
        if with_stats:
            return kps[:n].copy(), desc[:n].copy(), lc, cands
        return kps[:n].copy(), desc[:n].copy()

    def DetectWithPyramid(self, image, mask=None):
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        n = lib().orc_detect_with_pyramid(self._h, _p(image), _p(mask), image.shape[1], image.shape[0],
                                          image.strides[0], mask.strides[0] if mask is not None else 0, _p(kps),
                                          self.cap)
        return kps[:n].copy()

    def Detect(self, image, mask=None):
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        n = lib().orc_detect(self._h, _p(image), _p(mask), image.shape[1], image.shape[0], image.strides[0],
                             mask.strides[0] if mask is not None else 0, _p(kps), self.cap)
        return kps[:n].copy()

    def DetectWithCandidates(self, image, mask=None, cap=1 << 16):
        """Detect() plus the FAST candidate list [n,3] = (x, y, response), border-relative, in the order the
        reference's quadtree receives it (test hook)."""
        buf = np.zeros((cap, 3), np.float32)
        lib().orc_debug_arm_candidates(_p(buf), cap)
        try:
            kps = self.Detect(image, mask)
            n = lib().orc_debug_candidate_count()
        finally:
            lib().orc_debug_arm_candidates(None, 0)
        assert n <= cap
        return kps, buf[:n].copy()

    def ScreenAndComputeKPsParams(self, image, kps_in):
        """Returns (mutated input, surviving keypoints) like the reference's in/out vectors."""
        image = self._img(image)
        kin = np.ascontiguousarray(kps_in, KP_DTYPE).copy()
        out = np.zeros(max(1, len(kin)), KP_DTYPE)
        n = lib().orc_screen_params(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(kin),
                                    len(kin), _p(out))
        return kin, out[:n].copy()

    def CalcDescriptors(self, image, kps):
        image = self._img(image)
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.zeros((max(1, len(kps)), 32), np.uint8)
        n = lib().orc_calc_descriptors(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(kps),
                                       len(kps), _p(desc))
        return desc[:n].copy()

    def level(self, level, mask=False):
        w, h = C.c_int(), C.c_int()
        lib().orc_get_level_size(self._h, level, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        rc = lib().orc_get_level(self._h, level, int(mask), _p(out))
        assert rc == 0
        return out

    def blurred_level(self, level):
        w, h = C.c_int(), C.c_int()
        lib().orc_get_level_size(self._h, level, C.byref(w), C.byref(h))
        out = np.empty((h.value, w.value), np.uint8)
        rc = lib().orc_get_blurred_level(self._h, level, _p(out))
        assert rc == 0
        return out


# ---------------------------------------------------------------------------------------------------
# local bundle adjustment (oracle/ba_oracle.c)
# ---------------------------------------------------------------------------------------------------
def ba_solve(poses, points, fixed, obs_pose, obs_point, uv, K, ext7=(0, 0, 0, 1, 0, 0, 0), huber_delta=5.991,
             chi2_th=5.991, outer_max=5, inner_iters=10):
    """Backend::OptimizeActiveMap's solver part.  Returns (poses, points, chi2, outlier, info)."""
    poses = np.ascontiguousarray(poses, np.float64).copy()
    points = np.ascontiguousarray(points, np.float64).copy()
    fixed = np.ascontiguousarray(fixed, np.uint8)
    op = np.ascontiguousarray(obs_pose, np.int32)
    ol = np.ascontiguousarray(obs_point, np.int32)
    uv = np.ascontiguousarray(uv, np.float64)
    K = np.ascontiguousarray(K, np.float64)
    ext = np.ascontiguousarray(ext7, np.float64)
    chi2 = np.zeros(max(1, len(op)), np.float64)
    outl = np.zeros(max(1, len(op)), np.uint8)
    info = np.zeros(4, np.int32)
    rc = lib().orc_ba_solve(len(poses), len(points), len(op), _p(poses), _p(points), _p(fixed), _p(op), _p(ol), _p(uv),
                            _p(K), _p(ext), C.c_double(huber_delta), C.c_double(chi2_th), outer_max, inner_iters,
                            _p(chi2), _p(outl), _p(info))
    assert rc == 0
    return poses, points, chi2[:len(op)], outl[:len(op)], info


def se3_exp(d):
    d = np.ascontiguousarray(d, np.float64)
    R = np.zeros(9)
    t = np.zeros(3)
    lib().orc_se3_exp(_p(d), _p(R), _p(t))
    return R.reshape(3, 3), t


def pose_oplus(pose7, d):
    p = np.ascontiguousarray(pose7, np.float64).copy()
    d = np.ascontiguousarray(d, np.float64)
    lib().orc_pose_oplus(_p(p), _p(d))
    return p


def ba_edge(pose7, pt, uv, K, ext7=(0, 0, 0, 1, 0, 0, 0)):
    """-> (error[2], A[2,6], B[2,3]) of one EdgeProjection."""
    pose7 = np.ascontiguousarray(pose7, np.float64)
    pt = np.ascontiguousarray(pt, np.float64)
    uv = np.ascontiguousarray(uv, np.float64)
    K = np.ascontiguousarray(K, np.float64)
    ext = np.ascontiguousarray(ext7, np.float64)
    err, A, B = np.zeros(2), np.zeros(12), np.zeros(6)
    lib().orc_ba_edge_error(_p(pose7), _p(pt), _p(uv), _p(K), _p(ext), _p(err))
    lib().orc_ba_edge_jacobians(_p(pose7), _p(pt), _p(K), _p(ext), _p(A), _p(B))
    return err, A.reshape(2, 6), B.reshape(2, 3)


# ---------------------------------------------------------------------------------------------------
# DeepLCD scoring (numpy restatement of src/deeplcd.cpp:35-39 and src/loopclosing.cpp:124-161)
# ---------------------------------------------------------------------------------------------------
def lcd_score(d1, d2):
    """DeepLCD::score: float result = d1.transpose() * d2 (fp32)."""
    return np.float32(np.dot(np.asarray(d1, np.float32), np.asarray(d2, np.float32)))


def lcd_detect_loop(db_ids, db_descr, cur_id, cur_descr, thres_high=0.94, thres_low=0.92, min_gap=20, max_suspected=3):
    """LoopClosing::DetectLoop -> (found, bestId, maxScore, cntSuspected).  db_ids ascending (std::map)."""
    max_score, cnt, best = np.float32(0), 0, 0
    for i, d in zip(db_ids, db_descr):
        if (int(cur_id) - int(i)) % (1 << 64) < min_gap:   # unsigned long arithmetic, and `break`
            break
        s = lcd_score(cur_descr, d)
        if s > max_score:
            max_score, best = s, int(i)
        if s > np.float32(thres_low):
            cnt += 1
    if max_score < np.float32(thres_high) or cnt > max_suspected:
        return False, best, float(max_score), cnt
    return True, best, float(max_score), cnt


def pose_only_solve(pose0, points, uv, K, huber_delta=1.0, chi2_th=5.991, pre_rounds=0, rounds=4, inner_iters=10):
    """Frontend::EstimateCurrentPose / LoopClosing::OptimizeCurrentPose solver -> (pose [7], outlier [n], info [4])."""
    pose = np.ascontiguousarray(pose0, np.float64).copy()
    pts = np.ascontiguousarray(points, np.float64)
    uv = np.ascontiguousarray(uv, np.float64)
    K = np.ascontiguousarray(K, np.float64)
    outl = np.zeros(max(1, len(pts)), np.uint8)
    info = np.zeros(4, np.int32)
    lib().orc_pose_only_solve.restype = C.c_int
    rc = lib().orc_pose_only_solve(len(pts), _p(pose), _p(pts), _p(uv), _p(K), C.c_double(huber_delta), C.c_double(chi2_th),
                                   pre_rounds, rounds, inner_iters, _p(outl), _p(info))
    assert rc == 0
    return pose, outl[:len(pts)], info


def triangulate(uv_left, uv_right, K_left, K_right, pose_left7, pose_right7, T_wc7=None, ratio_th=1e-2):
    """numpy restatement of myslam::triangulation (include/myslam/algorithm.h:16-33: A from the two 3x4 pose matrices
    and the normalised points, SVD, V.col(3) / V(3,3)) and of the callers' test `ok && z > 0` (src/frontend.cpp:403,474)."""
    from oracle import posegraph_oracle as PG
    ul = np.asarray(uv_left, np.float32).reshape(-1, 2).astype(np.float64)
    ur = np.asarray(uv_right, np.float32).reshape(-1, 2).astype(np.float64)
    M = []
    for p in (pose_left7, pose_right7):
        R, t = PG.se3_from7(np.asarray(p, np.float64))
        M.append(np.concatenate([R, t[:, None]], 1))
    pts, ok = np.zeros((len(ul), 3)), np.zeros(len(ul), bool)
    for i in range(len(ul)):
        rows = []
        for (u, v), K, m in ((ul[i], K_left, M[0]), (ur[i], K_right, M[1])):
            x, y = (u - K[2]) / K[0], (v - K[3]) / K[1]          # Camera::pixel2camera, depth 1
            rows += [x * m[2] - m[0], y * m[2] - m[1]]
        _, s, vt = np.linalg.svd(np.array(rows))
        p = vt[3, :3] / vt[3, 3]
        ok[i] = (s[3] / s[2] < ratio_th) and p[2] > 0
        pts[i] = p
    if T_wc7 is not None:
        R, t = PG.se3_from7(np.asarray(T_wc7, np.float64))
        pts = pts @ R.T + t
    return pts, ok


# ---------------------------------------------------------------------------------------------------
# pyramidal Lucas-Kanade (oracle/lk_oracle.c)
# ---------------------------------------------------------------------------------------------------
def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_pyr_down_u8(_p(img), w, h, img.strides[0], _p(out), out.strides[0])
    return out


def scharr(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.empty((h, w, 2), np.int16)
    lib().orc_scharr_u8(_p(img), w, h, img.strides[0], _p(out))
    return out


def lk_track(prev, nxt, prev_pts, next_pts0=None, win=11, max_level=3, max_count=30, eps=0.01, min_eig_th=1e-4):
    """cv::calcOpticalFlowPyrLK as the front end calls it (src/frontend.cpp:150-153).  next_pts0 given ->
    OPTFLOW_USE_INITIAL_FLOW.  -> (next_pts [n,2] float32, status [n] uint8)"""
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    pp = np.ascontiguousarray(prev_pts, np.float32).reshape(-1, 2)
    use_init = next_pts0 is not None
    npts = np.ascontiguousarray(next_pts0, np.float32).reshape(-1, 2).copy() if use_init else pp.copy()
    status = np.zeros(max(1, len(pp)), np.uint8)
    lib().orc_lk_track.restype = C.c_int
    rc = lib().orc_lk_track(_p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], len(pp), _p(pp), _p(npts),
                            _p(status), win, max_level, max_count, C.c_double(eps), int(use_init), C.c_float(min_eig_th))
    assert rc == 0
    return npts, status[:len(pp)]
