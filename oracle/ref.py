"""ctypes front-end of oracle/_ref/libmyslam_orb_ref.so — the reference's OWN extractor
(/root/reference/src/ORBextractor.cpp compiled unmodified against oracle/ref_shim/, recipe: oracle/Makefile `ref`).

TEST INFRASTRUCTURE ONLY.  Same call surface as oracle.ORBextractor so that tests can run
reference == restatement == CUDA on the same inputs.  /root/reference does not exist on the GPU box: there the
prebuilt library (it travels with the snapshot) is loaded as is; where neither the library nor the reference
sources exist, available() is False and the tests that need it skip.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .oracle import KP_DTYPE, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libmyslam_orb_ref.so")
_REF_SRC = "/root/reference/src/ORBextractor.cpp"
_LIB = None


def build():
    """(Re)build when the reference sources are present; otherwise use the prebuilt file if there is one."""
    if os.path.exists(_REF_SRC):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return _SO if os.path.exists(_SO) else None


def available():
    return os.path.exists(_SO) or os.path.exists(_REF_SRC)


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            raise RuntimeError("oracle/_ref is not built and /root/reference is absent")
        _LIB = C.CDLL(so)
        _LIB.ref_create.restype = C.c_void_p
        _LIB.ref_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _LIB.ref_destroy.argtypes = [C.c_void_p]
        for name in ("ref_detect_and_compute", "ref_detect_with_pyramid", "ref_detect", "ref_screen_params",
                     "ref_calc_descriptors", "ref_get_level", "ref_distribute_octtree", "ref_error_count",
                     "ref_list_node_bytes"):
            getattr(_LIB, name).restype = C.c_int
    return _LIB


class ORBextractor:
    """myslam::ORBextractor itself (include/myslam/ORBextractor.h:47-138).

    monotone_nodes=True serves the quadtree's std::list nodes from an address-monotone arena, which turns the
    reference's pointer tie-break (src/ORBextractor.cpp:731) into "later-created node first"; False leaves them to
    glibc malloc (the order a real run happens to get)."""

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, monotone_nodes=True):
        self.nfeatures, self.nlevels, self.monotone = nfeatures, nlevels, bool(monotone_nodes)
        self._h = C.c_void_p(lib().ref_create(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST))
        assert self._h
        n = nlevels
        self.scale = np.zeros(n, np.float32)
        self.inv_scale = np.zeros(n, np.float32)
        self.sigma2 = np.zeros(n, np.float32)
        self.inv_sigma2 = np.zeros(n, np.float32)
        self.quota = np.zeros(n, np.int32)
        self.umax = np.zeros(16, np.int32)
        lib().ref_get_tables(self._h, _p(self.scale), _p(self.inv_scale), _p(self.sigma2), _p(self.inv_sigma2),
                             _p(self.quota), _p(self.umax))
        self.cap = nfeatures + 67 * nlevels + 600

    def __del__(self):
        try:
            lib().ref_destroy(self._h)
        except Exception:
            pass

    def _mode(self):
        lib().ref_set_monotone_nodes(int(self.monotone))

    @staticmethod
    def _img(a):
        a = np.ascontiguousarray(a, np.uint8)
        assert a.ndim == 2
        return a

    def DetectAndCompute(self, image, mask=None):
        self._mode()
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        n = lib().ref_detect_and_compute(self._h, _p(image), _p(mask), image.shape[1], image.shape[0], image.strides[0],
                                         mask.strides[0] if mask is not None else 0, _p(kps), _p(desc), self.cap)
        assert 0 <= n <= self.cap
        return kps[:n].copy(), desc[:n].copy()

    def DetectWithPyramid(self, image, mask=None):
        self._mode()
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        n = lib().ref_detect_with_pyramid(self._h, _p(image), _p(mask), image.shape[1], image.shape[0], image.strides[0],
                                          mask.strides[0] if mask is not None else 0, _p(kps), self.cap)
        assert 0 <= n <= self.cap
        return kps[:n].copy()

    def Detect(self, image, mask=None):
        self._mode()
        image = self._img(image)
        mask = self._img(mask) if mask is not None else None
        kps = np.zeros(self.cap, KP_DTYPE)
        n = lib().ref_detect(self._h, _p(image), _p(mask), image.shape[1], image.shape[0], image.strides[0],
                             mask.strides[0] if mask is not None else 0, _p(kps), self.cap)
        assert 0 <= n <= self.cap
        return kps[:n].copy()

    def ScreenAndComputeKPsParams(self, image, kps_in):
        image = self._img(image)
        kin = np.ascontiguousarray(kps_in, KP_DTYPE).copy()
        out = np.zeros(max(1, len(kin)), KP_DTYPE)
        n = lib().ref_screen_params(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(kin),
                                    len(kin), _p(out))
        return kin, out[:n].copy()

    def CalcDescriptors(self, image, kps):
        image = self._img(image)
        kps = np.ascontiguousarray(kps, KP_DTYPE)
        desc = np.zeros((max(1, len(kps)), 32), np.uint8)
        n = lib().ref_calc_descriptors(self._h, _p(image), image.shape[1], image.shape[0], image.strides[0], _p(kps),
                                       len(kps), _p(desc))
        return desc[:n].copy()

    def level(self, level):
        w, h = C.c_int(), C.c_int()
        assert lib().ref_get_level(self._h, level, None, C.byref(w), C.byref(h)) == 0
        out = np.empty((h.value, w.value), np.uint8)
        assert lib().ref_get_level(self._h, level, _p(out), C.byref(w), C.byref(h)) == 0
        return out

    def distribute_octtree(self, kx, ky, kr, minX, maxX, minY, maxY, N):
        """DistributeOctTree on a candidate list; returns [n,3] (x, y, response) in the reference's output order."""
        self._mode()
        kx = np.ascontiguousarray(kx, np.float32)
        ky = np.ascontiguousarray(ky, np.float32)
        kr = np.ascontiguousarray(kr, np.float32)
        out = np.zeros((max(1, len(kx)), 3), np.float32)
        n = lib().ref_distribute_octtree(self._h, _p(kx), _p(ky), _p(kr), len(kx), minX, maxX, minY, maxY, N, _p(out))
        return out[:n].copy()
