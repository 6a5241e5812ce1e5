"""Property test (hypothesis) of the level-synchronous quadtree (csrc/qt_core.inl, host build) against the literal
oracle on random candidate sets: any number of candidates, any quota, elongated and squat boxes."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def qt_lib():
    global _LIB
    if _LIB is None:
        import tempfile
        so = os.path.join(tempfile.mkdtemp(), "libqtmodel.so")
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                               os.path.join(HERE, "qt_host_model.cpp")])
        _LIB = C.CDLL(so)
        _LIB.qt_model.restype = C.c_int
    return _LIB


@settings(max_examples=60, deadline=None)
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 1500), N=st.integers(1, 700),
       dims=st.sampled_from([(1209, 344), (314, 73), (640, 448), (200, 200), (90, 400)]))
def test_random_candidates(seed, n, N, dims):
    from oracle import oracle as O
    O.build()
    width, height = dims
    rng = np.random.default_rng(seed)
    nCols, nRows = int(np.float32(width) / np.float32(30)), int(np.float32(height) / np.float32(30))
    wCell, hCell = int(math.ceil(np.float32(width) / np.float32(nCols))), int(math.ceil(np.float32(height) / np.float32(nRows)))
    # distinct integer positions inside the tested area of the grid, responses 7..254
    xs = rng.integers(3, min(width - 3, nCols * wCell + 3), n)
    ys = rng.integers(3, min(height - 3, nRows * hCell + 3), n)
    pos = np.unique(np.stack([ys, xs], 1), axis=0)
    ys, xs = pos[:, 0], pos[:, 1]
    resp = rng.integers(7, 255, len(xs))
    # the reference's candidate order: cell row, cell column, then row-major inside the cell
    ci, cj = (ys - 3) // hCell, (xs - 3) // wCell
    order = np.lexsort((xs, ys, cj, ci))
    xs, ys, resp = xs[order], ys[order], resp[order]
    want_idx = O.distribute_octtree(xs, ys, resp, 16, 16 + width, 16, 16 + height, N)
    want = np.stack([xs[want_idx], ys[want_idx], resp[want_idx]], 1).astype(np.int64)
    words = (xs.astype(np.uint32) | (ys.astype(np.uint32) << 12) | (resp.astype(np.uint32) << 24))
    words = np.ascontiguousarray(words[rng.permutation(len(words))])
    nIni = max(1, int(round(width / height)))
    cap = max(N + 3, 4 * nIni) + 16
    out = np.zeros(cap, np.uint32)
    m = qt_lib().qt_model(words.ctypes.data_as(C.c_void_p), len(words), width, height, N, nCols, wCell, hCell,
                          out.ctypes.data_as(C.c_void_p), cap)
    got = np.stack([out[:m] & 0xfff, (out[:m] >> 12) & 0xfff, out[:m] >> 24], 1).astype(np.int64)
    assert got.shape == want.shape and np.array_equal(got, want)
