"""Host build of csrc/orb_core.inl (the arithmetic the CUDA kernels execute) against the oracle."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hm") / "liborbmodel.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(HERE, "orb_host_model.cpp")])
    lib = C.CDLL(so)
    lib.hm_fast_atan2.restype = C.c_float
    lib.hm_fast_atan2.argtypes = [C.c_float, C.c_float]
    return lib


def _p(a):
    return C.c_void_p(a.ctypes.data)


@pytest.mark.parametrize("t0", [7, 20])
def test_fast_response_plane(hm, oracle, synth, t0):
    left, _ = synth.stereo_pair(6)
    roi = np.ascontiguousarray(left[40:200, 300:700])
    h, w = roi.shape
    plane = np.zeros((h, w), np.uint8)
    hm.hm_fast_plane(_p(roi), w, h, w, t0, _p(plane))
    want = np.zeros((h, w), np.uint8)
    got = oracle.fast9_16(roi, t0, nonmax=False)
    want[got[:, 1], got[:, 0]] = got[:, 2]
    assert got[:, 2].min() >= t0
    assert np.array_equal(plane, want)


def test_fast_atan2(hm, oracle):
    rng = np.random.default_rng(1)
    for y, x in rng.integers(-300000, 300000, (20000, 2)).tolist() + [[0, 0], [0, -3], [4, 0], [-4, 0]]:
        assert hm.hm_fast_atan2(float(y), float(x)) == oracle.fast_atan2(float(y), float(x))


def test_resize_and_gauss(hm, oracle, synth):
    left, _ = synth.stereo_pair(8)
    for (dw, dh) in [(1034, 313), (517, 200), (1240, 375)]:
        out = np.zeros((dh, dw), np.uint8)
        hm.hm_resize(_p(left), 1241, 376, _p(out), dw, dh)
        assert np.array_equal(out, oracle.resize_linear(left, dw, dh))
    small = np.ascontiguousarray(left[:105, :346])
    out = np.zeros_like(small)
    hm.hm_gauss(_p(small), 346, 105, _p(out))
    assert np.array_equal(out, oracle.gauss7(small))
