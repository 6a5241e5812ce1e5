"""The CUDA pyramidal LK tracker against the oracle (oracle/lk_oracle.c, itself pinned to cv2 in tests/test_oracle_lk.py):
BIT-EXACT positions and status — both sides accumulate the sums of products exactly in 64-bit integers and run the
same un-contracted fp32 arithmetic.  Reference call sites: src/frontend.cpp:150-153, :358-361."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pts(oracle, img, n):
    k = oracle.ORBextractor(n, 1.2, 8, 20, 7).Detect(img)
    return np.stack([k["x"], k["y"]], 1).astype(np.float32)


def test_left_to_right_and_temporal_batches(pkg, oracle, synth):
    frames = synth.stereo_batch(40, 3)
    lk = pkg.LKTracker(max_batch=4, max_pts=400)
    prev = [frames[b, 0] for b in range(3)]
    nxt = [frames[b, 1] for b in range(3)]
    pts = [_pts(oracle, prev[b], 300 - 40 * b) for b in range(3)]
    # FindFeaturesInRight: initial guess = the left position
    res = lk.track(prev, nxt, pts, [p.copy() for p in pts])
    for b in range(3):
        want, wst = oracle.lk_track(prev[b], nxt[b], pts[b], pts[b].copy())
        got, gst = res[b]
        assert np.array_equal(gst, wst), (b, np.nonzero(gst != wst)[0][:5])
        assert got.tobytes() == want.tobytes(), (b, np.abs(got - want).max())
        assert gst.mean() > 0.6
    # TrackLastFrame: next frame = shifted view, guess off by a few pixels; and the no-initial-flow path
    shifted = [np.roll(p, (1, -6), axis=(0, 1)) for p in prev]
    guess = [p + np.float32([-4.0, 0.5]) for p in pts]
    for init in (guess, None):
        res = lk.track(prev, shifted, pts, init)
        for b in range(3):
            want, wst = oracle.lk_track(prev[b], shifted[b], pts[b], None if init is None else init[b])
            assert np.array_equal(res[b][1], wst) and res[b][0].tobytes() == want.tobytes()
    lk.close()


def test_border_points_other_window_and_small_images(pkg, oracle, synth):
    left, right = synth.stereo_pair(9)
    rng = np.random.default_rng(1)
    pts = np.stack([rng.uniform(-4, 1246, 500), rng.uniform(-4, 381, 500)], 1).astype(np.float32)
    lk = pkg.LKTracker(max_batch=1, max_pts=512)
    for win in (11, 7, 21):
        got, gst = lk.track([left], [right], [pts], [pts.copy()], win=win)[0]
        want, wst = oracle.lk_track(left, right, pts, pts.copy(), win=win)
        assert np.array_equal(gst, wst) and got.tobytes() == want.tobytes(), win
    small_l, small_r = np.ascontiguousarray(left[:90, :160]), np.ascontiguousarray(right[:90, :160])   # pyramid stops at level 2
    sp = np.stack([rng.uniform(5, 150, 60), rng.uniform(5, 85, 60)], 1).astype(np.float32)
    got, gst = lk.track([small_l], [small_r], [sp], None)[0]
    want, wst = oracle.lk_track(small_l, small_r, sp, None)
    assert np.array_equal(gst, wst) and got.tobytes() == want.tobytes()
    assert lk.track([left], [right], [np.zeros((0, 2), np.float32)], None)[0][0].shape == (0, 2)
    lk.close()
