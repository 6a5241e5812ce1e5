"""BASELINE config 1, the correctness anchor: 200 seeded synthetic stereo frames through extract (both views) + left->right
match, against the committed digests of tests/golden/config1_digests.npz (made by tests/golden/make_golden.py from the CPU
restatement; KITTI and the reference's own binary are not available here).  CPU: the oracle still reproduces its own
committed results (a drift of the checker would otherwise go unnoticed).  GPU: all 200 frames bit-exact through the C ABI."""
import hashlib
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ORB_PARAMS = (2000, 1.2, 8, 20, 7)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


@pytest.fixture(scope="module")
def golden():
    g = np.load(os.path.join(HERE, "golden", "config1_digests.npz"))
    assert tuple(g["orb_params"]) == ORB_PARAMS and len(g["counts"]) == 200
    return g


def test_oracle_reproduces_committed_results(golden, oracle, synth):
    ext = oracle.ORBextractor(*ORB_PARAMS)
    for f in range(12):
        left, right = synth.stereo_pair(f)
        kl, dl = ext.DetectAndCompute(left)
        kr, dr = ext.DetectAndCompute(right)
        idx, dist = oracle.hamming_match(dl, dr)
        assert (len(kl), len(kr)) == tuple(golden["counts"][f])
        assert np.array_equal(digest(kl, kr), golden["kp_digest"][f])
        assert np.array_equal(digest(dl, dr), golden["desc_digest"][f])
        assert np.array_equal(digest(idx.astype(np.int32), dist.astype(np.int32)), golden["match_digest"][f])
        if f < 2:   # the full arrays of the first two frames are committed as well
            assert kl.tobytes() == golden[f"f{f}_kl"].tobytes() and np.array_equal(dl, golden[f"f{f}_dl"])
            assert np.array_equal(idx, golden[f"f{f}_idx"]) and np.array_equal(dist, golden[f"f{f}_dist"])


@pytest.mark.gpu
def test_cuda_path_matches_all_200_frames(golden, pkg, synth):
    B = 50
    ext = pkg.ORBextractor(*ORB_PARAMS, max_batch=2 * B)
    matcher = pkg.HammingMatcher(max_batch=B, max_rows=ext.cap)
    for f0 in range(0, 200, B):
        pairs = [synth.stereo_pair(f) for f in range(f0, f0 + B)]
        res = ext.DetectAndComputeBatch([im for p in pairs for im in p])
        ms = matcher.match_batch([res[2 * i][1] for i in range(B)], [res[2 * i + 1][1] for i in range(B)])
        for i in range(B):
            f = f0 + i
            (kl, dl), (kr, dr), (idx, dist) = res[2 * i], res[2 * i + 1], ms[i]
            assert (len(kl), len(kr)) == tuple(golden["counts"][f]), f
            assert np.array_equal(digest(kl, kr), golden["kp_digest"][f]), f
            assert np.array_equal(digest(dl, dr), golden["desc_digest"][f]), f
            assert np.array_equal(digest(idx.astype(np.int32), dist.astype(np.int32)), golden["match_digest"][f]), f
    matcher.close()
    ext.close()


def _ba_golden():
    return np.load(os.path.join(HERE, "golden", "config3_ba.npz"))


def test_ba_oracle_reproduces_committed_results(oracle, synth):
    g = _ba_golden()
    for seed in range(2):
        w = synth.ba_window(seed)
        p, x, chi2, outl, info = oracle.ba_solve(w["poses0"], w["points0"], w["fixed"], w["obs_pose"], w["obs_point"], w["uv"], synth.KITTI_K)
        assert np.array_equal(np.asarray(info, np.int32), g[f"w{seed}_info"])
        assert np.abs(p - g[f"w{seed}_poses"]).max() < 1e-9 and np.abs(x - g[f"w{seed}_points"]).max() < 1e-9


@pytest.mark.gpu
def test_cuda_ba_matches_committed_results(pkg, synth):
    """BASELINE tolerance: poses and landmarks within 1e-4 relative; fp64 on both sides lands far inside it."""
    g = _ba_golden()
    ba = pkg.LocalBA(max_windows=4, max_poses=7, max_points=512, max_obs=4096)
    res = ba.solve([synth.ba_window(s) for s in range(4)], synth.KITTI_K)
    for seed, (p, x, chi2, outl, info) in enumerate(res):
        assert np.array_equal(info, g[f"w{seed}_info"])
        assert np.abs(p - g[f"w{seed}_poses"]).max() < 1e-7 and np.abs(x - g[f"w{seed}_points"]).max() < 1e-6
        assert (outl != g[f"w{seed}_outlier"]).sum() <= 2      # only edges whose chi2 sits at the threshold may flip
    ba.close()
