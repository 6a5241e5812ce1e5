"""Host build of csrc/pnp_core.inl (quartic roots, Grunert P3P) against numpy: the arithmetic of the PnP-RANSAC kernel
(loop geometric verification, reference src/loopclosing.cpp:207-293) checked on the CPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pnp") / "libpnpmodel.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so, os.path.join(HERE, "pnp_host_model.cpp")])
    lib = C.CDLL(so)
    lib.hm_quartic.argtypes = [C.c_double] * 4 + [C.c_void_p]
    lib.hm_p3p.argtypes = [C.c_void_p] * 3
    return lib


def _rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def test_quartic_roots_match_numpy(hm):
    rng = np.random.default_rng(0)
    out = np.zeros(4)
    for trial in range(2000):
        if trial % 4 == 0:      # four real roots
            r = rng.uniform(-5, 5, 4)
            co = np.poly(r)
        elif trial % 4 == 1:    # biquadratic (q = 0 branch)
            co = np.array([1.0, 0.0, rng.uniform(-5, 1), 0.0, rng.uniform(-2, 2)])
        else:
            co = np.concatenate([[1.0], rng.normal(size=4) * 3])
        n = hm.hm_quartic(co[1], co[2], co[3], co[4], out.ctypes.data)
        want = np.roots(co)
        sep = min([abs(a - b) for i, a in enumerate(want) for b in want[i + 1:]] + [1.0])
        if sep < 1e-3:
            continue            # (nearly) multiple roots are ill-conditioned for any solver
        want = np.sort(want[np.abs(want.imag) < 1e-9].real)
        got = np.sort(out[:n])
        assert len(got) == len(want), (co, got, want)
        assert np.allclose(got, want, rtol=1e-8, atol=1e-8), (co, got, want)


def test_p3p_recovers_the_pose(hm):
    rng = np.random.default_rng(1)
    Rt = np.zeros(48)
    done = 0
    for trial in range(500):
        R = _rot(rng)
        t = rng.normal(size=3) * 2 + np.array([0, 0, 8.0])
        P = rng.uniform(-3, 3, (3, 3))
        Pc = P @ R.T + t
        if (Pc[:, 2] < 0.5).any():
            continue
        j = np.ascontiguousarray(Pc / np.linalg.norm(Pc, axis=1, keepdims=True))
        P = np.ascontiguousarray(P)
        n = hm.hm_p3p(P.ctypes.data, j.ctypes.data, Rt.ctypes.data)
        assert 1 <= n <= 4
        sols = Rt[:12 * n].reshape(n, 12)
        err = [max(np.abs(s[:9].reshape(3, 3) - R).max(), np.abs(s[9:] - t).max()) for s in sols]
        assert min(err) < 1e-6, (trial, err)
        for s in sols:          # every returned solution is a rotation that reproduces the three bearings
            Rs, ts = s[:9].reshape(3, 3), s[9:]
            assert np.allclose(Rs @ Rs.T, np.eye(3), atol=1e-9) and np.linalg.det(Rs) > 0.999
            q = P @ Rs.T + ts
            assert np.allclose(q / np.linalg.norm(q, axis=1, keepdims=True), j, atol=1e-6)
        done += 1
    assert done > 400
