"""Pins the LK checker (oracle/lk_oracle.c) against cv2 4.13.0: pyrDown and the Scharr derivative bit-exactly, the
tracker itself within the tolerance its header explains (exact integer accumulation vs OpenCV's SIMD-ordered fp32)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

LK = dict(winSize=(11, 11), maxLevel=3, criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01))


def test_pyr_down_bit_exact(oracle, synth):
    left, _ = synth.stereo_pair(1)
    img = left
    for _ in range(3):
        want = cv2.pyrDown(img)
        got = oracle.pyr_down(img)
        assert np.array_equal(got, want)
        img = want
    odd = np.ascontiguousarray(left[:101, :77])
    assert np.array_equal(oracle.pyr_down(odd), cv2.pyrDown(odd))


def test_scharr_bit_exact(oracle, synth):
    left, _ = synth.stereo_pair(2)
    img = np.ascontiguousarray(left[:120, :333])
    gx = cv2.Scharr(img, cv2.CV_16S, 1, 0, borderType=cv2.BORDER_REFLECT_101)
    gy = cv2.Scharr(img, cv2.CV_16S, 0, 1, borderType=cv2.BORDER_REFLECT_101)
    d = oracle.scharr(img)
    assert np.array_equal(d[..., 0], gx) and np.array_equal(d[..., 1], gy)


def _points(synth, oracle, img, n, seed):
    kps = oracle.ORBextractor(n, 1.2, 8, 20, 7).Detect(img)
    return np.stack([kps["x"], kps["y"]], 1).astype(np.float32)


@pytest.mark.parametrize("seed", [0, 3])
def test_left_to_right_tracking_like_find_features_in_right(oracle, synth, seed):
    left, right = synth.stereo_pair(seed)
    pts = _points(synth, oracle, left, 300, seed)
    init = pts.copy()                                         # "use same pixel position in left image" (:352)
    want, wst, _ = cv2.calcOpticalFlowPyrLK(left, right, pts, init.copy(), flags=cv2.OPTFLOW_USE_INITIAL_FLOW, **LK)
    got, gst = oracle.lk_track(left, right, pts, init)
    wst = wst.ravel()
    assert (gst != wst).mean() < 0.01                         # borderline minEig / bounds decisions only
    both = (gst == 1) & (wst == 1)
    d = np.abs(got[both] - want[both]).max(1)
    assert both.sum() > 100
    assert (d < 2e-3).mean() > 0.995 and d.max() < 0.05, (np.sort(d)[-5:],)


def test_temporal_tracking_with_projected_initial_guess(oracle, synth):
    left, _ = synth.stereo_pair(5)
    nxt = np.roll(left, (2, -7), axis=(0, 1))                 # the "next frame": a shifted view
    pts = _points(synth, oracle, left, 200, 5)
    init = pts + np.float32([-5.5, 1.25])                     # an imperfect projection-based guess (:137-139)
    want, wst, _ = cv2.calcOpticalFlowPyrLK(left, nxt, pts, init.copy(), flags=cv2.OPTFLOW_USE_INITIAL_FLOW, **LK)
    got, gst = oracle.lk_track(left, nxt, pts, init)
    both = (gst == 1) & (wst.ravel() == 1)
    assert (gst != wst.ravel()).mean() < 0.01 and both.sum() > 100
    d = np.abs(got[both] - want[both]).max(1)
    assert (d < 2e-3).mean() > 0.995 and d.max() < 0.05
    flow = np.median(got[both] - pts[both], 0)
    assert np.allclose(flow, [-7, 2], atol=0.05)


def test_without_initial_flow_and_points_near_the_border(oracle, synth):
    left, right = synth.stereo_pair(6)
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-3, 1245, 400), rng.uniform(-3, 380, 400)], 1).astype(np.float32)
    want, wst, _ = cv2.calcOpticalFlowPyrLK(left, right, pts, None, **LK)
    got, gst = oracle.lk_track(left, right, pts, None)
    wst = wst.ravel()
    assert (gst != wst).mean() < 0.02
    both = (gst == 1) & (wst == 1)
    d = np.abs(got[both] - want[both]).max(1)
    assert (d < 2e-3).mean() > 0.99 and d.max() < 0.1
