"""The level-synchronous quadtree of csrc/qt_core.inl (host build, every cooperative loop serial)
against the literal restatement of ORBextractor::DistributeOctTree in oracle/orb_oracle.c
(reference src/ORBextractor.cpp:586-810)."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def qt(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("qt") / "libqtmodel.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-o", so,
                           os.path.join(HERE, "qt_host_model.cpp")])
    lib = C.CDLL(so)
    lib.qt_model.restype = C.c_int
    return lib


def grid_geometry(cols, rows):
    """ComputeKeyPointsOctTree grid (src/ORBextractor.cpp:826-836), float arithmetic as the reference."""
    f = np.float32
    width, height = f(cols - 32), f(rows - 32)
    nCols, nRows = int(width / f(30)), int(height / f(30))
    wCell, hCell = int(math.ceil(width / f(nCols))), int(math.ceil(height / f(nRows)))
    return int(width), int(height), nCols, nRows, wCell, hCell


def run_model(qt, cand, width, height, N, nCols, wCell, hCell):
    words = (cand[:, 0].astype(np.uint32) | (cand[:, 1].astype(np.uint32) << 12) |
             (cand[:, 2].astype(np.uint32) << 24))
    # the device appends candidates in arbitrary order: shuffle to prove order independence
    rng = np.random.default_rng(len(words))
    words = np.ascontiguousarray(words[rng.permutation(len(words))])
    cap = max(N + 3, 64) + 16
    out = np.zeros(cap, np.uint32)
    n = qt.qt_model(words.ctypes.data_as(C.c_void_p), len(words), width, height, N, nCols, wCell, hCell,
                    out.ctypes.data_as(C.c_void_p), cap)
    out = out[:n]
    return np.stack([out & 0xfff, (out >> 12) & 0xfff, out >> 24], 1).astype(np.int64)


@pytest.mark.parametrize("seed", [0, 1, 2])
@pytest.mark.parametrize("nfeatures", [2000, 300, 100, 37])
def test_quadtree_matches_oracle_on_pyramid_levels(qt, oracle, synth, seed, nfeatures):
    left, _ = synth.stereo_pair(seed)
    ext = oracle.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    ext.DetectAndCompute(left)
    for level in range(8):
        img = ext.level(level)
        N = int(ext.quota[level])
        single = oracle.ORBextractor(N, 1.2, 8, 20, 7)
        kps, cand = single.DetectWithCandidates(img)
        width, height, nCols, nRows, wCell, hCell = grid_geometry(img.shape[1], img.shape[0])
        got = run_model(qt, cand, width, height, N, nCols, wCell, hCell)
        want = np.stack([kps["x"] - 16, kps["y"] - 16, kps["response"]], 1).astype(np.int64)
        assert got.shape == want.shape, (level, got.shape, want.shape)
        assert np.array_equal(got, want), level


@pytest.mark.parametrize("N", [1, 2, 3, 4, 5, 8, 50, 500, 5000])
def test_quadtree_small_and_saturated(qt, oracle, synth, N):
    left, _ = synth.stereo_pair(7)
    single = oracle.ORBextractor(N, 1.2, 8, 20, 7)
    kps, cand = single.DetectWithCandidates(left)
    width, height, nCols, nRows, wCell, hCell = grid_geometry(left.shape[1], left.shape[0])
    got = run_model(qt, cand, width, height, N, nCols, wCell, hCell)
    want = np.stack([kps["x"] - 16, kps["y"] - 16, kps["response"]], 1).astype(np.int64)
    assert np.array_equal(got, want)


def test_quadtree_sparse_candidates(qt, oracle):
    """Few candidates: some roots empty, nodes that never split."""
    rng = np.random.default_rng(3)
    img = np.full((376, 1241), 100, np.uint8)
    for _ in range(25):
        x, y = rng.integers(30, 1200), rng.integers(30, 340)
        img[y:y + 6, x:x + 6] = 220
    single = oracle.ORBextractor(300, 1.2, 8, 20, 7)
    kps, cand = single.DetectWithCandidates(img)
    assert 0 < len(cand) < 300
    width, height, nCols, nRows, wCell, hCell = grid_geometry(1241, 376)
    got = run_model(qt, cand, width, height, 300, nCols, wCell, hCell)
    want = np.stack([kps["x"] - 16, kps["y"] - 16, kps["response"]], 1).astype(np.int64)
    assert np.array_equal(got, want)
