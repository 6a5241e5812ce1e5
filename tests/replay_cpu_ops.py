"""CPU checker for the replay (TEST INFRASTRUCTURE): the operator set of replay.GpuOps over oracle/ — the same pipeline
driven by the same host logic, so a GPU replay can be compared stage by stage with the restated reference arithmetic."""
import numpy as np

from oracle import calc_oracle as CO
from oracle import oracle as O
from oracle import pnp_oracle as PO
from oracle import posegraph_oracle as PG


def _expand_octaves(feats, kp_dtype, nlevels=8):
    """Every feature as a keypoint on each octave, response -1, class_id = feature index (src/loopclosing.cpp:94-105)."""
    kin = np.zeros(len(feats) * nlevels, kp_dtype)
    rep = np.repeat(np.arange(len(feats)), nlevels)
    for name in ("x", "y", "size", "angle"):
        kin[name] = feats[name][rep]
    kin["response"] = -1
    kin["octave"] = np.tile(np.arange(nlevels), len(feats))
    kin["class_id"] = rep
    return kin


class CpuOps:
    def __init__(self, synth, capi_kp_dtype, batch=8, kf_features=300, kf_batch=8):
        self.synth, self.batch, self.kf_batch = synth, batch, kf_batch
        self.ext = O.ORBextractor(2000, 1.2, 8, 20, 7)
        self.kf_ext = O.ORBextractor(kf_features, 1.2, 8, 20, 7)
        self.cap = self.ext.cap
        self.kp_dtype = capi_kp_dtype
        self.weights = synth.calc_weights(0)
        self.db_ids, self.db = [], []
        self._slots, self._ba = {}, None

    def stereo_submit(self, slot, images):
        n = len(images)
        cap = 2400
        out = {"kps": np.zeros((n, 2, cap), self.kp_dtype), "desc": np.zeros((n, 2, cap, 32), np.uint8), "counts": np.zeros((n, 2), np.int32),
               "midx": np.full((n, cap), -1, np.int32), "mdist": np.full((n, cap), -1, np.int32)}
        for i in range(n):
            ds = []
            for v in range(2):
                k, d = self.ext.DetectAndCompute(images[i, v])
                out["kps"][i, v, :len(k)] = k
                out["desc"][i, v, :len(k)] = d
                out["counts"][i, v] = len(k)
                ds.append(d)
            idx, dist = O.hamming_match(ds[0], ds[1])
            out["midx"][i, :len(idx)] = idx
            out["mdist"][i, :len(idx)] = dist
        self._slots[slot] = out

    def stereo_wait(self, slot):
        return self._slots[slot]

    def ba_submit(self, windows):
        self._ba = [O.ba_solve(w["poses0"], w["points0"], w["fixed"], w["obs_pose"], w["obs_point"], w["uv"], self.synth.KITTI_K) for w in windows]

    def ba_wait(self):
        return self._ba

    def kf_detect(self, lefts):
        return [self.kf_ext.Detect(im) for im in lefts]

    def lk_right(self, lefts, rights, pts):
        return [O.lk_track(l, r, p, next_pts0=p) for l, r, p in zip(lefts, rights, pts)]

    def triangulate(self, ul, ur, T_wc7):
        b = self.synth.KITTI_BF / self.synth.KITTI_FX
        return O.triangulate(ul, ur, self.synth.KITTI_K, self.synth.KITTI_K, np.array([0, 0, 0, 1, 0, 0, 0.0]), np.array([0, 0, 0, 1, -b, 0, 0.0]), T_wc7)

    def triangulate_batch(self, uls, urs):
        return [self.triangulate(a, b, None) for a, b in zip(uls, urs)]

    def cnn_descr(self, lefts):
        out = []
        for im in lefts:
            d, blurred = CO.calc_descr_original(im, self.weights)
            im[...] = blurred
            out.append(np.asarray(d, np.float32).ravel())
        return np.stack(out)

    def screen_and_describe(self, imgs, feats):
        out = []
        for img, f in zip(imgs, feats):
            kin = _expand_octaves(f, self.kp_dtype)      # src/loopclosing.cpp:94-105
            _, kout = self.kf_ext.ScreenAndComputeKPsParams(img, kin)
            out.append((kout, self.kf_ext.CalcDescriptors(img, kout) if len(kout) else np.zeros((0, 32), np.uint8)))
        return out

    def lcd_add(self, kf_id, d):
        self.db_ids.append(kf_id)
        self.db.append(np.asarray(d, np.float32))

    def lcd_size(self):
        return len(self.db)

    def lcd_detect(self, kf_id, d, min_gap):
        return O.lcd_detect_loop(self.db_ids, self.db, kf_id, d, 0.94, 0.92, min_gap, 3)

    def match(self, q, t):
        return O.hamming_match(q, t)

    def pnp_ransac(self, obj, img, seed):
        ok, rvec, tvec, mask = PO.solve_pnp_ransac(obj, img, self.synth.KITTI_K)
        import cv2
        R = cv2.Rodrigues(rvec.reshape(3, 1))[0] if ok else np.eye(3)
        return dict(found=ok, pose7=self.synth.pose7(R, tvec), rvec=rvec, tvec=tvec, inliers=mask)

    def pose_refine(self, pose7, pts, uv):
        return O.pose_only_solve(pose7, pts, uv, self.synth.KITTI_K, pre_rounds=1)

    def posegraph(self, poses, fixed, v0, v1, meas):
        return PG.solve(poses, fixed, v0, v1, meas)
