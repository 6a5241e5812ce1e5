"""ctypes binding of include/slamb200.h.  Class and method names follow the reference's operator
surface (include/myslam/ORBextractor.h:47-138; cv::DescriptorMatcher::match as used at
src/loopclosing.cpp:172) so that the parity tests read like calls into the reference."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

SB_OK, SB_ERR_INVALID, SB_ERR_CUDA, SB_ERR_CAPACITY, SB_ERR_OVERFLOW = 0, -1, -2, -3, -4


class SlamB200Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"libslamb200 error {code}: {text}")
        self.code = code


def lib_path():
    # SLAMB200_LIB: development hook to load an experimental build of the same library (tools/)
    return os.environ.get("SLAMB200_LIB") or os.path.join(_HERE, "libslamb200.so")


def lib():
    """Loads the product library; raises if it has not been built (there is no fallback)."""
    global _LIB
    if _LIB is None:
        path = lib_path()
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C csrc); this package has no CPU fallback")
        L = C.CDLL(path)
        L.sb_last_error.restype = C.c_char_p
        L.sb_version.restype = C.c_char_p
        _LIB = L
    return _LIB


def last_error():
    return lib().sb_last_error().decode()


def _check(rc):
    if rc != SB_OK:
        raise SlamB200Error(rc, last_error())


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _dev_ptr(t):
    """torch CUDA tensor / int address -> c_void_p."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(int(t))


class ORBextractor:
    """myslam::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST) on `device`."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_w=1241, max_h=376, max_batch=1,
                 device=0):
        self._h = C.c_void_p()
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.max_batch = max_batch
        _check(lib().sb_orb_create(C.byref(self._h), device, nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST,
                                   minThFAST, max_w, max_h, max_batch))
        self.cap = lib().sb_orb_capacity(self._h)
        n = nlevels
        self.scale = np.zeros(n, np.float32)
        self.inv_scale = np.zeros(n, np.float32)
        self.sigma2 = np.zeros(n, np.float32)
        self.inv_sigma2 = np.zeros(n, np.float32)
        self.quota = np.zeros(n, np.int32)
        nl = C.c_int()
        _check(lib().sb_orb_get_tables(self._h, C.byref(nl), _p(self.scale), _p(self.inv_scale), _p(self.sigma2),
                                       _p(self.inv_sigma2), _p(self.quota)))

    def close(self):
        if self._h:
            lib().sb_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- getters of the reference class (ORBextractor.h:88-106)
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactors(self):
        return self.scale

    def GetInverseScaleFactors(self):
        return self.inv_scale

    def GetScaleSigmaSquares(self):
        return self.sigma2

    def GetInverseScaleSigmaSquares(self):
        return self.inv_sigma2

    def set_stream(self, stream_ptr):
        _check(lib().sb_orb_set_stream(self._h, C.c_void_p(stream_ptr)))

    def sync_status(self):
        _check(lib().sb_orb_sync_status(self._h))

    @staticmethod
    def _imgs(images):
        if isinstance(images, np.ndarray) and images.ndim == 2:
            images = [images]
        imgs = [np.ascontiguousarray(i, np.uint8) for i in images]
        h, w = imgs[0].shape
        assert all(i.shape == (h, w) for i in imgs)
        return imgs, w, h

    @staticmethod
    def _ptr_array(arrs):
        if arrs is None:
            return None
        PA = C.c_void_p * len(arrs)
        return PA(*[a.ctypes.data if a is not None else None for a in arrs])

    def _masks(self, masks, n, w, h):
        if masks is None:
            return None, None
        if isinstance(masks, np.ndarray) and masks.ndim == 2:
            masks = [masks]
        ms = [np.ascontiguousarray(m, np.uint8) if m is not None else None for m in masks]
        assert len(ms) == n and all(m is None or m.shape == (h, w) for m in ms)
        return ms, self._ptr_array(ms)

    def DetectAndComputeBatch(self, images, masks=None, descriptors=True):
        """DetectAndCompute on a list of equally sized images -> list of (keypoints, descriptors)."""
        imgs, w, h = self._imgs(images)
        n = len(imgs)
        ms, mptr = self._masks(masks, n, w, h)
        kps = np.zeros((n, self.cap), KP_DTYPE)
        desc = np.zeros((n, self.cap, 32), np.uint8) if descriptors else None
        counts = np.zeros(n, np.int32)
        _check(lib().sb_orb_detect_and_compute(self._h, n, self._ptr_array(imgs), mptr, w, h, w, w, _p(kps), _p(desc),
                                               _p(counts), self.cap))
        return [(kps[b, :counts[b]].copy(), desc[b, :counts[b]].copy() if descriptors else None) for b in range(n)]

    def DetectAndCompute(self, image, mask=None):
        return self.DetectAndComputeBatch([image], None if mask is None else [mask])[0]

    def DetectWithPyramid(self, image, mask=None):
        return self.DetectAndComputeBatch([image], None if mask is None else [mask], descriptors=False)[0][0]

    def DetectBatch(self, images, masks=None):
        imgs, w, h = self._imgs(images)
        n = len(imgs)
        ms, mptr = self._masks(masks, n, w, h)
        kps = np.zeros((n, self.cap), KP_DTYPE)
        counts = np.zeros(n, np.int32)
        _check(lib().sb_orb_detect(self._h, n, self._ptr_array(imgs), mptr, w, h, w, w, _p(kps), _p(counts), self.cap))
        return [kps[b, :counts[b]].copy() for b in range(n)]

    def Detect(self, image, mask=None):
        return self.DetectBatch([image], None if mask is None else [mask])[0]

    def ScreenAndComputeKPsParams(self, image, kps_in):
        """-> (input keypoints as the reference leaves them, surviving keypoints)."""
        imgs, w, h = self._imgs(image)
        kin = np.ascontiguousarray(kps_in, KP_DTYPE).copy()
        out = np.zeros(max(1, len(kin)), KP_DTYPE)
        n_out = C.c_int32()
        _check(lib().sb_orb_screen_params(self._h, _p(imgs[0]), w, h, w, _p(kin), len(kin), _p(out), C.byref(n_out)))
        return kin, out[:n_out.value].copy()

    def ScreenAndDescribeBatch(self, images, kps_list, n_in=None, copy=True):
        """sb_orb_screen_describe: ScreenAndComputeKPsParams + CalcDescriptors on the survivors for a batch of images.
        -> list of (mutated input, surviving keypoints, descriptors) per image.  kps_list: a list of keypoint arrays, or — with
        n_in — one packed [B, cap_in] array (no per-image packing); copy=False returns views into the batch arrays."""
        imgs, w, h = self._imgs(images)
        B = len(imgs)
        if n_in is None:
            cap_in = max(1, max(len(k) for k in kps_list))
            kin = np.zeros((B, cap_in), KP_DTYPE)
            n_in = np.zeros(B, np.int32)
            for b, k in enumerate(kps_list):
                kin[b, :len(k)] = k
                n_in[b] = len(k)
        else:
            kin = np.ascontiguousarray(kps_list, KP_DTYPE)
            n_in = np.ascontiguousarray(n_in, np.int32)
            cap_in = kin.shape[1]
        out = np.empty((B, cap_in), KP_DTYPE)          # the library writes the first n_out[b] rows
        n_out = np.zeros(B, np.int32)
        desc = np.empty((B, cap_in, 32), np.uint8)
        _check(lib().sb_orb_screen_describe(self._h, B, self._ptr_array(imgs), w, h, imgs[0].strides[0], _p(kin), _p(n_in), cap_in,
                                            _p(out), _p(n_out), _p(desc)))
        if copy:
            return [(kin[b, :n_in[b]].copy(), out[b, :n_out[b]].copy(), desc[b, :n_out[b]].copy()) for b in range(B)]
        return [(kin[b, :n_in[b]], out[b, :n_out[b]], desc[b, :n_out[b]]) for b in range(B)]

    def CalcDescriptors(self, image, kps):
        imgs, w, h = self._imgs(image)
        k = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.zeros((max(1, len(k)), 32), np.uint8)
        _check(lib().sb_orb_calc_descriptors(self._h, _p(imgs[0]), w, h, w, _p(k), len(k), _p(desc)))
        return desc[:len(k)].copy()

    # -- asynchronous device entry points (torch tensors or raw addresses)
    def detect_and_compute_dev(self, batch, d_img, img_pitch, w, h, stride, d_kps, d_desc, d_counts, cap, d_mask=None,
                               mask_pitch=0, mstride=0):
        _check(lib().sb_orb_detect_and_compute_dev(self._h, batch, _dev_ptr(d_img), C.c_int64(img_pitch),
                                                   _dev_ptr(d_mask), C.c_int64(mask_pitch), w, h, stride, mstride,
                                                   _dev_ptr(d_kps), _dev_ptr(d_desc), _dev_ptr(d_counts), cap))

    # -- inspection
    def debug_level(self, b, level, which=0):
        lw, lh = C.c_int(), C.c_int()
        _check(lib().sb_orb_debug_level(self._h, b, level, which, None, 0, C.byref(lw), C.byref(lh)))
        out = np.empty((lh.value, lw.value), np.uint8)
        _check(lib().sb_orb_debug_level(self._h, b, level, which, _p(out), out.size, C.byref(lw), C.byref(lh)))
        return out

    def debug_candidates(self, b, level, cap=1 << 15):
        out = np.zeros(cap, np.uint32)
        n = C.c_int32()
        _check(lib().sb_orb_debug_candidates(self._h, b, level, _p(out), cap, C.byref(n)))
        out = out[:n.value]
        return np.stack([out & 0xfff, (out >> 12) & 0xfff, out >> 24], 1).astype(np.int64)


class HammingMatcher:
    """cv::BFMatcher(NORM_HAMMING): match(query, train) -> (trainIdx, distance) per query row."""

    def __init__(self, max_batch=1, max_rows=4096, device=0):
        self._h = C.c_void_p()
        self.max_batch, self.max_rows = max_batch, max_rows
        _check(lib().sb_matcher_create(C.byref(self._h), device, max_batch, max_rows))

    def close(self):
        if self._h:
            lib().sb_matcher_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(lib().sb_matcher_set_stream(self._h, C.c_void_p(stream_ptr)))

    def match_batch(self, queries, trains):
        n = len(queries)
        cap = max(1, max(max(len(q) for q in queries), max(len(t) for t in trains)))
        Q = np.zeros((n, cap, 32), np.uint8)
        T = np.zeros((n, cap, 32), np.uint8)
        nq = np.array([len(q) for q in queries], np.int32)
        nt = np.array([len(t) for t in trains], np.int32)
        for b in range(n):
            Q[b, :nq[b]] = np.asarray(queries[b], np.uint8).reshape(-1, 32)
            T[b, :nt[b]] = np.asarray(trains[b], np.uint8).reshape(-1, 32)
        idx = np.zeros((n, cap), np.int32)
        dist = np.zeros((n, cap), np.int32)
        _check(lib().sb_hamming_match(self._h, n, _p(Q), _p(nq), _p(T), _p(nt), cap, _p(idx), _p(dist)))
        return [(idx[b, :nq[b]].copy(), dist[b, :nq[b]].copy()) for b in range(n)]

    def match(self, query, train):
        return self.match_batch([query], [train])[0]

    def match_dev(self, batch, d_q, q_set_stride, d_nq, nq_stride, d_t, t_set_stride, d_nt, nt_stride, max_rows, d_idx,
                  d_dist, out_stride):
        _check(lib().sb_hamming_match_dev(self._h, batch, _dev_ptr(d_q), C.c_int64(q_set_stride), _dev_ptr(d_nq),
                                          nq_stride, _dev_ptr(d_t), C.c_int64(t_set_stride), _dev_ptr(d_nt), nt_stride,
                                          max_rows, _dev_ptr(d_idx), _dev_ptr(d_dist), C.c_int64(out_stride)))


class LocalBA:
    """The g2o solve of Backend::OptimizeActiveMap (src/backend.cpp:126-269), batched over windows."""

    def __init__(self, max_windows=1, max_poses=7, max_points=1024, max_obs=8192, device=0):
        self._h = C.c_void_p()
        self.W, self.MP, self.ML, self.MO = max_windows, max_poses, max_points, max_obs
        _check(lib().sb_ba_create(C.byref(self._h), device, max_windows, max_poses, max_points, max_obs))

    def close(self):
        if self._h:
            lib().sb_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(lib().sb_ba_set_stream(self._h, C.c_void_p(stream_ptr)))

    def pack(self, windows):
        """list of dicts (poses0, points0, fixed, obs_pose, obs_point, uv) -> padded batch arrays"""
        n = len(windows)
        b = {"np": np.zeros(n, np.int32), "nl": np.zeros(n, np.int32), "ne": np.zeros(n, np.int32),
             "poses": np.zeros((n, self.MP, 7)), "points": np.zeros((n, self.ML, 3)),
             "fixed": np.zeros((n, self.ML), np.uint8), "op": np.zeros((n, self.MO), np.int32),
             "ol": np.zeros((n, self.MO), np.int32), "uv": np.zeros((n, self.MO, 2))}
        b["poses"][:, :, 3] = 1.0
        for k, w in enumerate(windows):
            p, l, e = len(w["poses0"]), len(w["points0"]), len(w["obs_pose"])
            b["np"][k], b["nl"][k], b["ne"][k] = p, l, e
            b["poses"][k, :p] = w["poses0"]
            b["points"][k, :l] = w["points0"]
            b["fixed"][k, :l] = w["fixed"]
            b["op"][k, :e] = w["obs_pose"]
            b["ol"][k, :e] = w["obs_point"]
            b["uv"][k, :e] = w["uv"]
        return b

    def solve(self, windows, K, ext7=(0, 0, 0, 1, 0, 0, 0), huber_delta=5.991, chi2_th=5.991, outer_max=5,
              inner_iters=10):
        """-> list of (poses, points, chi2, outlier, info) per window."""
        b = self.pack(windows)
        n = len(windows)
        K = np.ascontiguousarray(K, np.float64)
        ext = np.ascontiguousarray(ext7, np.float64)
        chi2 = np.zeros((n, self.MO))
        outl = np.zeros((n, self.MO), np.uint8)
        info = np.zeros((n, 4), np.int32)
        _check(lib().sb_ba_solve(self._h, n, _p(b["np"]), _p(b["nl"]), _p(b["ne"]), _p(b["poses"]), _p(b["points"]),
                                 _p(b["fixed"]), _p(b["op"]), _p(b["ol"]), _p(b["uv"]), _p(K), _p(ext),
                                 C.c_double(huber_delta), C.c_double(chi2_th), outer_max, inner_iters, _p(chi2), _p(outl),
                                 _p(info)))
        out = []
        for k in range(n):
            p, l, e = b["np"][k], b["nl"][k], b["ne"][k]
            out.append((b["poses"][k, :p].copy(), b["points"][k, :l].copy(), chi2[k, :e].copy(), outl[k, :e].copy(),
                        info[k].copy()))
        return out

    def submit(self, windows, K, ext7=(0, 0, 0, 1, 0, 0, 0), huber_delta=5.991, chi2_th=5.991, outer_max=5, inner_iters=10):
        """sb_ba_submit: enqueue the batch and return; wait() collects it (same result as solve)."""
        b = self.pack(windows)
        n = len(windows)
        st = dict(b=b, n=n, K=np.ascontiguousarray(K, np.float64), ext=np.ascontiguousarray(ext7, np.float64),
                  chi2=np.zeros((n, self.MO)), outl=np.zeros((n, self.MO), np.uint8), info=np.zeros((n, 4), np.int32))
        _check(lib().sb_ba_submit(self._h, n, _p(b["np"]), _p(b["nl"]), _p(b["ne"]), _p(b["poses"]), _p(b["points"]),
                                  _p(b["fixed"]), _p(b["op"]), _p(b["ol"]), _p(b["uv"]), _p(st["K"]), _p(st["ext"]),
                                  C.c_double(huber_delta), C.c_double(chi2_th), outer_max, inner_iters, _p(st["chi2"]),
                                  _p(st["outl"]), _p(st["info"])))
        self._inflight = st   # keeps the host arrays alive until wait()

    def wait(self):
        st, self._inflight = self._inflight, None
        _check(lib().sb_ba_wait(self._h))
        b, out = st["b"], []
        for k in range(st["n"]):
            p, l, e = b["np"][k], b["nl"][k], b["ne"][k]
            out.append((b["poses"][k, :p].copy(), b["points"][k, :l].copy(), st["chi2"][k, :e].copy(), st["outl"][k, :e].copy(),
                        st["info"][k].copy()))
        return out

    def solve_dev(self, n, d, K, ext7=(0, 0, 0, 1, 0, 0, 0), huber_delta=5.991, chi2_th=5.991, outer_max=5,
                  inner_iters=10):
        """d: dict of device tensors with the keys of pack() plus chi2, outlier, info."""
        K = np.ascontiguousarray(K, np.float64)
        ext = np.ascontiguousarray(ext7, np.float64)
        _check(lib().sb_ba_solve_dev(self._h, n, _dev_ptr(d["np"]), _dev_ptr(d["nl"]), _dev_ptr(d["ne"]),
                                     _dev_ptr(d["poses"]), _dev_ptr(d["points"]), _dev_ptr(d["fixed"]), _dev_ptr(d["op"]),
                                     _dev_ptr(d["ol"]), _dev_ptr(d["uv"]), _p(K), _p(ext), C.c_double(huber_delta),
                                     C.c_double(chi2_th), outer_max, inner_iters, _dev_ptr(d["chi2"]),
                                     _dev_ptr(d["outlier"]), _dev_ptr(d["info"])))


class DeepLCDScorer:
    """DeepLCD::score over the keyframe database + LoopClosing::DetectLoop (src/deeplcd.cpp:35-39,
    src/loopclosing.cpp:124-161)."""
    FP32, FP16 = 0, 1

    def __init__(self, capacity=1024, dtype=1, max_queries=1, device=0):
        self._h = C.c_void_p()
        self.capacity, self.max_queries = capacity, max_queries
        _check(lib().sb_lcd_create(C.byref(self._h), device, capacity, dtype, max_queries))

    def close(self):
        if self._h:
            lib().sb_lcd_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return lib().sb_lcd_size(self._h)

    def set_stream(self, stream_ptr):
        _check(lib().sb_lcd_set_stream(self._h, C.c_void_p(stream_ptr)))

    def add(self, kf_id, descr):
        d = np.ascontiguousarray(descr, np.float32).reshape(1064)
        _check(lib().sb_lcd_add(self._h, C.c_int64(kf_id), _p(d)))

    def add_batch(self, kf_ids, descr):
        ids = np.ascontiguousarray(kf_ids, np.int64)
        d = np.ascontiguousarray(descr, np.float32).reshape(len(ids), 1064)
        _check(lib().sb_lcd_add_batch(self._h, len(ids), _p(ids), _p(d)))

    def remove(self, kf_id):
        _check(lib().sb_lcd_remove(self._h, C.c_int64(kf_id)))

    def score(self, queries):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, 1064)
        out = np.zeros((len(q), len(self)), np.float32)
        _check(lib().sb_lcd_score(self._h, len(q), _p(q), _p(out)))
        return out

    def score_dev(self, nq, d_queries, d_scores, stride):
        _check(lib().sb_lcd_score_dev(self._h, nq, _dev_ptr(d_queries), _dev_ptr(d_scores), stride))

    def DetectLoop(self, cur_kf_id, query, thres_high=0.94, thres_low=0.92, min_gap=20, max_suspected=3):
        q = np.ascontiguousarray(query, np.float32).reshape(1064)
        found, cnt, best, mx = C.c_int(), C.c_int(), C.c_int64(), C.c_float()
        _check(lib().sb_lcd_detect_loop(self._h, C.c_int64(cur_kf_id), _p(q), C.c_float(thres_high), C.c_float(thres_low),
                                        min_gap, max_suspected, C.byref(found), C.byref(best), C.byref(mx), C.byref(cnt)))
        return bool(found.value), best.value, mx.value, cnt.value


class PoseGraph:
    """The g2o solve of LoopClosing::PoseGraphOptimization (src/loopclosing.cpp:537-646)."""

    def __init__(self, max_vertices=1024, max_edges=2048, device=0, max_loops=None):
        self._h = C.c_void_p()
        if max_loops is None:
            _check(lib().sb_posegraph_create(C.byref(self._h), device, max_vertices, max_edges))
        else:
            _check(lib().sb_posegraph_create_loops(C.byref(self._h), device, max_vertices, max_edges, max_loops))

    def close(self):
        if self._h:
            lib().sb_posegraph_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(lib().sb_posegraph_set_stream(self._h, C.c_void_p(stream_ptr)))

    def solve(self, poses, fixed, v0, v1, meas, iters=20):
        """-> (poses [n,7], info dict)"""
        p = np.ascontiguousarray(poses, np.float64).copy()
        fx = np.ascontiguousarray(fixed, np.uint8)
        a = np.ascontiguousarray(v0, np.int32)
        b = np.ascontiguousarray(v1, np.int32)
        z = np.ascontiguousarray(meas, np.float64)
        info = np.zeros(4, np.int32)
        stats = np.zeros(2, np.float64)
        _check(lib().sb_posegraph_solve(self._h, len(p), _p(p), _p(fx), len(a), _p(a), _p(b), _p(z), iters, _p(info),
                                        _p(stats)))
        return p, {"lm_iters": int(info[0]), "trials": int(info[1]), "free": int(info[2]), "loops": int(info[3]),
                   "chi2_start": float(stats[0]), "chi2": float(stats[1])}


class StereoFrontend:
    """DetectAndCompute on both views + Hamming match(left -> right) for a batch of frames in one call
    (sb_stereo_*): host buffers in, host buffers out; submit()/wait() for pipelining with a second handle."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_w=1241, max_h=376, max_pairs=1,
                 device=0):
        self._h = C.c_void_p()
        self.max_pairs = max_pairs
        _check(lib().sb_stereo_create(C.byref(self._h), device, nfeatures, C.c_float(scaleFactor), nlevels, iniThFAST,
                                      minThFAST, max_w, max_h, max_pairs))
        self.cap = lib().sb_stereo_capacity(self._h)

    def close(self):
        if self._h:
            lib().sb_stereo_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_compute_stream(self, stream):
        """Kernels on `stream` (an int cudaStream_t shared by several handles), copies on the handle's own stream."""
        _check(lib().sb_stereo_set_compute_stream(self._h, C.c_void_p(stream)))

    def alloc_outputs(self, pairs, pinned=False):
        """Output buffers for submit(): dict of numpy arrays (backed by pinned torch tensors if pinned)."""
        shapes = {"kps": ((pairs, 2, self.cap), KP_DTYPE), "desc": ((pairs, 2, self.cap, 32), np.uint8),
                  "counts": ((pairs, 2), np.int32), "midx": ((pairs, self.cap), np.int32),
                  "mdist": ((pairs, self.cap), np.int32)}
        out = {}
        if pinned:
            import torch
            out["_keep"] = []
            for k, (shp, dt) in shapes.items():
                nbytes = int(np.prod(shp)) * np.dtype(dt).itemsize
                t = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                out["_keep"].append(t)
                out[k] = t.numpy().view(dt).reshape(shp)
        else:
            for k, (shp, dt) in shapes.items():
                out[k] = np.zeros(shp, dt)
        return out

    def submit(self, images, out):
        """images: uint8 [pairs, 2, h, w] (C-contiguous; pinned for asynchronous copies)."""
        pairs, two, h, w = images.shape
        assert two == 2 and images.dtype == np.uint8 and images.flags["C_CONTIGUOUS"]
        _check(lib().sb_stereo_submit(self._h, pairs, C.c_void_p(images.ctypes.data), C.c_int64(2 * h * w), C.c_int64(h * w),
                                      w, h, w, _p(out["kps"]), _p(out["desc"]), _p(out["counts"]), _p(out["midx"]),
                                      _p(out["mdist"])))

    def wait(self):
        _check(lib().sb_stereo_wait(self._h))

    def extract_match(self, images):
        images = np.ascontiguousarray(images, np.uint8)
        out = self.alloc_outputs(images.shape[0])
        self.submit(images, out)
        self.wait()
        return out


class PoseOnlyOptimizer:
    """The g2o solve of Frontend::EstimateCurrentPose (src/frontend.cpp:176-276; pre_rounds=0) and
    LoopClosing::OptimizeCurrentPose (src/loopclosing.cpp:339-433; pre_rounds=1), batched over frames."""

    def __init__(self, max_frames=1, max_obs=2048, device=0):
        self._h = C.c_void_p()
        self.F, self.MO = max_frames, max_obs
        _check(lib().sb_pose_create(C.byref(self._h), device, max_frames, max_obs))

    def close(self):
        if self._h:
            lib().sb_pose_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, frames, K, huber_delta=1.0, chi2_th=5.991, pre_rounds=0, rounds=4, inner_iters=10):
        """frames: list of dicts (pose0 [7], points [n,3], uv [n,2]) -> list of (pose, outlier, info)."""
        n = len(frames)
        cnt = np.array([len(f["points"]) for f in frames], np.int32)
        poses = np.stack([np.asarray(f["pose0"], np.float64) for f in frames])
        pts = np.zeros((n, self.MO, 3))
        uv = np.zeros((n, self.MO, 2))
        for k, f in enumerate(frames):
            pts[k, :cnt[k]] = f["points"]
            uv[k, :cnt[k]] = f["uv"]
        K = np.ascontiguousarray(K, np.float64)
        outl = np.zeros((n, self.MO), np.uint8)
        info = np.zeros((n, 4), np.int32)
        _check(lib().sb_pose_solve(self._h, n, _p(cnt), _p(poses), _p(pts), _p(uv), _p(K), C.c_double(huber_delta),
                                   C.c_double(chi2_th), pre_rounds, rounds, inner_iters, _p(outl), _p(info)))
        return [(poses[k].copy(), outl[k, :cnt[k]].copy(), info[k].copy()) for k in range(n)]


def triangulate(uv_left, uv_right, K_left, K_right, pose_left7, pose_right7, T_wc7=None, ratio_th=1e-2, device=0):
    """myslam::triangulation (include/myslam/algorithm.h:16-33) + the callers' acceptance test, for n correspondences.
    -> (points [n,3] float64, ok [n] bool)"""
    ul = np.ascontiguousarray(uv_left, np.float32).reshape(-1, 2)
    ur = np.ascontiguousarray(uv_right, np.float32).reshape(-1, 2)
    n = len(ul)
    pts = np.zeros((max(n, 1), 3))
    ok = np.zeros(max(n, 1), np.uint8)
    Kl, Kr = np.ascontiguousarray(K_left, np.float64), np.ascontiguousarray(K_right, np.float64)
    pl, pr = np.ascontiguousarray(pose_left7, np.float64), np.ascontiguousarray(pose_right7, np.float64)
    tw = np.ascontiguousarray(T_wc7, np.float64) if T_wc7 is not None else None
    _check(lib().sb_triangulate(device, n, _p(ul), _p(ur), _p(Kl), _p(Kr), _p(pl), _p(pr), _p(tw), C.c_double(ratio_th),
                                _p(pts), _p(ok)))
    return pts[:n], ok[:n].astype(bool)


class PnPRansac:
    """cv::solvePnPRansac as LoopClosing::ComputeCorrectPose calls it (src/loopclosing.cpp:259-268), batched over loop
    candidates: hypotheses in parallel, inliers of the best one, Levenberg-Marquardt refinement on them."""

    def __init__(self, max_problems=1, max_points=2048, device=0):
        self._h = C.c_void_p()
        self.P, self.MP = max_problems, max_points
        _check(lib().sb_pnp_create(C.byref(self._h), device, max_problems, max_points))

    def close(self):
        if self._h:
            lib().sb_pnp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def solve(self, problems, K, iterations=100, reproj_err=5.991, seed=0):
        """problems: list of (obj [n,3], img [n,2]) -> list of dicts (found, pose7, rvec, tvec, inliers [n] bool, info [4])."""
        n = len(problems)
        cnt = np.array([len(o) for o, _ in problems], np.int32)
        obj = np.zeros((n, self.MP, 3), np.float32)
        img = np.zeros((n, self.MP, 2), np.float32)
        for k, (o, m) in enumerate(problems):
            obj[k, :cnt[k]] = o
            img[k, :cnt[k]] = m
        K = np.ascontiguousarray(K, np.float64)
        pose = np.zeros((n, 7))
        rt = np.zeros((n, 6))
        inl = np.zeros((n, self.MP), np.uint8)
        info = np.zeros((n, 4), np.int32)
        _check(lib().sb_pnp_ransac(self._h, n, _p(cnt), _p(obj), _p(img), _p(K), int(iterations), C.c_double(reproj_err),
                                   C.c_uint64(seed), _p(pose), _p(rt), _p(inl), _p(info)))
        return [dict(found=bool(info[k, 0]), pose7=pose[k].copy(), rvec=rt[k, :3].copy(), tvec=rt[k, 3:].copy(),
                     inliers=inl[k, :cnt[k]].astype(bool), info=info[k].copy()) for k in range(n)]


class LKTracker:
    """cv::calcOpticalFlowPyrLK as Frontend::TrackLastFrame / FindFeaturesInRight call it (src/frontend.cpp:150-153,
    :358-361), batched over image pairs."""

    def __init__(self, max_w=1241, max_h=376, max_batch=1, max_pts=512, max_level=3, device=0):
        self._h = C.c_void_p()
        self.max_batch, self.max_pts = max_batch, max_pts
        _check(lib().sb_lk_create(C.byref(self._h), device, max_w, max_h, max_batch, max_pts, max_level))

    def close(self):
        if self._h:
            lib().sb_lk_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def track(self, prev_imgs, next_imgs, prev_pts, next_pts0=None, win=11, max_count=30, eps=0.01, min_eig_th=1e-4):
        """prev_imgs / next_imgs: lists of equally sized uint8 images; prev_pts: list of [n,2]; next_pts0: list of [n,2] or
        None (no OPTFLOW_USE_INITIAL_FLOW).  -> list of (next_pts [n,2] float32, status [n] uint8)"""
        B = len(prev_imgs)
        prev_imgs = [np.ascontiguousarray(i, np.uint8) for i in prev_imgs]
        next_imgs = [np.ascontiguousarray(i, np.uint8) for i in next_imgs]
        h, w = prev_imgs[0].shape
        cnt = np.array([len(p) for p in prev_pts], np.int32)
        pp = np.zeros((B, self.max_pts, 2), np.float32)
        npts = np.zeros((B, self.max_pts, 2), np.float32)
        for b in range(B):
            pp[b, :cnt[b]] = prev_pts[b]
            npts[b, :cnt[b]] = next_pts0[b] if next_pts0 is not None else prev_pts[b]
        status = np.zeros((B, self.max_pts), np.uint8)
        PA = C.c_void_p * B
        _check(lib().sb_lk_track(self._h, B, PA(*[i.ctypes.data for i in prev_imgs]), PA(*[i.ctypes.data for i in next_imgs]), w, h, w,
                                 _p(cnt), _p(pp), _p(npts), _p(status), win, max_count, C.c_double(eps),
                                 int(next_pts0 is not None), C.c_float(min_eig_th)))
        return [(npts[b, :cnt[b]].copy(), status[b, :cnt[b]].copy()) for b in range(B)]


class CalcLayer(C.Structure):
    """sb_calc_layer (include/slamb200.h)."""
    _fields_ = [("type", C.c_int32), ("num_output", C.c_int32), ("kernel", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32),
                ("local_size", C.c_int32), ("alpha", C.c_float), ("beta", C.c_float), ("k", C.c_float)]


# deploy.prototxt of the reference's DeepLCD model restated as a layer list (absent from the reference tree: CALC's
# published architecture, which yields the 1064 outputs src/deeplcd.cpp:82 asserts)
CALC_LAYERS = (
    dict(type=0, num_output=64, kernel=5, stride=2, pad=4), dict(type=1),
    dict(type=2, kernel=3, stride=2, pad=0), dict(type=3, local_size=5, alpha=1e-4, beta=0.75, k=1.0),
    dict(type=0, num_output=128, kernel=4, stride=1, pad=2), dict(type=1),
    dict(type=2, kernel=3, stride=2, pad=0), dict(type=3, local_size=5, alpha=1e-4, beta=0.75, k=1.0),
    dict(type=0, num_output=4, kernel=3, stride=1, pad=0), dict(type=1),
)


def parse_caffe(prototxt, caffemodel):
    """(layers, weights, (in_h, in_w)) of a Caffe deploy.prototxt + .caffemodel pair (sb_calc_parse_caffe; no device needed)."""
    nl, nw, shape = C.c_int(), C.c_int64(), (C.c_int * 2)()
    a, b = os.fsencode(prototxt), os.fsencode(caffemodel)
    _check(lib().sb_calc_parse_caffe(a, b, None, 0, C.byref(nl), None, C.c_int64(0), C.byref(nw), shape))
    arr = (CalcLayer * nl.value)()
    w = np.empty(nw.value, np.float32)
    _check(lib().sb_calc_parse_caffe(a, b, arr, nl.value, C.byref(nl), _p(w), C.c_int64(w.size), C.byref(nw), shape))
    layers = [dict(type=L.type, num_output=L.num_output, kernel=L.kernel, stride=L.stride, pad=L.pad, local_size=L.local_size,
                   alpha=L.alpha, beta=L.beta, k=L.k) for L in arr]
    return layers, w, (shape[0], shape[1])


class DeepLCD:
    """myslam::DeepLCD (include/myslam/deeplcd.h:20-48) with the network given as (layers, weights) instead of
    (deploy.prototxt, calc.caffemodel): calcDescrOriginalImg / calcDescr / score, batched."""

    def __init__(self, weights, layers=CALC_LAYERS, in_h=120, in_w=160, max_batch=1, max_img_w=1241, max_img_h=376, device=0):
        self._h = C.c_void_p()
        arr = (CalcLayer * len(layers))()
        for i, L in enumerate(layers):
            arr[i] = CalcLayer(L["type"], L.get("num_output", 0), L.get("kernel", 0), L.get("stride", 0), L.get("pad", 0),
                               L.get("local_size", 0), L.get("alpha", 0.0), L.get("beta", 0.0), L.get("k", 0.0))
        w = np.ascontiguousarray(weights, np.float32)
        _check(lib().sb_calc_create(C.byref(self._h), device, in_h, in_w, arr, len(layers), _p(w), C.c_int64(w.size), max_batch,
                                    max_img_w, max_img_h))
        self.dim = lib().sb_calc_descr_dim(self._h)
        self.in_h, self.in_w, self.max_batch = in_h, in_w, max_batch

    @classmethod
    def from_caffe(cls, prototxt="calc_model/deploy.prototxt", caffemodel="calc_model/calc.caffemodel", gpu_id=0, **kw):
        """DeepLCD(network_definition_file, pre_trained_model_file, gpu_id) (include/myslam/deeplcd.h:35)."""
        layers, weights, (in_h, in_w) = parse_caffe(prototxt, caffemodel)
        return cls(weights, layers=layers, in_h=in_h, in_w=in_w, device=gpu_id, **kw)

    def close(self):
        if self._h:
            lib().sb_calc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, stream_ptr):
        _check(lib().sb_calc_set_stream(self._h, C.c_void_p(stream_ptr)))

    def calcDescrOriginalImgBatch(self, images, in_place=True):
        """Returns descriptors [n][dim]; with in_place the images (writable u8 arrays) come back blurred, as the
        reference leaves KeyFrame::mImageLeft (src/deeplcd.cpp:46)."""
        imgs = [np.asarray(i) for i in images]
        h, w = imgs[0].shape
        for i in imgs:
            assert i.dtype == np.uint8 and i.shape == (h, w) and i.strides[1] == 1
        stride = imgs[0].strides[0]
        assert all(i.strides[0] == stride for i in imgs)
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        descr = np.empty((len(imgs), self.dim), np.float32)
        _check(lib().sb_calc_descr_original(self._h, len(imgs), ptrs, w, h, stride, _p(descr), ptrs if in_place else None))
        return descr

    def calcDescrOriginalImg(self, image, in_place=True):
        return self.calcDescrOriginalImgBatch([image], in_place)[0]

    def calcDescrBatch(self, images):
        imgs = [np.ascontiguousarray(i, np.uint8) for i in images]
        for i in imgs:
            assert i.shape == (self.in_h, self.in_w)
        ptrs = (C.c_void_p * len(imgs))(*[i.ctypes.data for i in imgs])
        descr = np.empty((len(imgs), self.dim), np.float32)
        _check(lib().sb_calc_descr(self._h, len(imgs), ptrs, self.in_w, _p(descr)))
        return descr

    def calcDescr(self, image):
        return self.calcDescrBatch([image])[0]

    def descr_original_dev(self, batch, d_img, img_pitch, w, h, stride, d_descr, d_blurred=None):
        _check(lib().sb_calc_descr_original_dev(self._h, batch, _dev_ptr(d_img), C.c_int64(img_pitch), w, h, stride, _dev_ptr(d_descr),
                                                _dev_ptr(d_blurred)))

    @staticmethod
    def score(d1, d2):
        raise NotImplementedError("scoring runs on the keyframe database: DeepLCDScorer.score / DetectLoop")
