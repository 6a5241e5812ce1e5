"""Pins the CPU checker (oracle/orb_oracle.c) against the in-container OpenCV (cv2 4.13.0), the only
implementation of the reference's un-vendored OpenCV primitives available here.
Bit-exact throughout."""
import math

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


def test_resize_linear_chained_kitti_levels(oracle, synth):
    left, _ = synth.stereo_pair(11)
    ext = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    prev = left
    for level in range(1, 8):
        s = ext.inv_scale[level]
        dw, dh = int(np.rint(np.float32(1241) * s)), int(np.rint(np.float32(376) * s))
        want = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = oracle.resize_linear(prev, dw, dh)
        assert np.array_equal(got, want), level
        prev = want


@pytest.mark.parametrize("shape,dst", [((37, 53), (44, 31)), ((100, 7), (6, 83)), ((64, 64), (64, 64)),
                                       ((9, 200), (167, 8))])
def test_resize_linear_odd_shapes(oracle, shape, dst):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    dw, dh = dst
    if dw > shape[1] or dh > shape[0]:
        pytest.skip("the reference only shrinks")
    assert np.array_equal(oracle.resize_linear(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))


@pytest.mark.parametrize("shape", [(376, 1241), (105, 346), (8, 9), (7, 300), (31, 4)])
def test_gaussian_7x7_sigma2(oracle, shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    want = cv2.GaussianBlur(img.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(oracle.gauss7(img), want)


def test_fast_atan2(oracle):
    rng = np.random.default_rng(0)
    ys = rng.integers(-200000, 200000, 20000)
    xs = rng.integers(-200000, 200000, 20000)
    for y, x in zip(ys.tolist() + [0, 0, 5, -5, 0], xs.tolist() + [0, 7, 0, 0, -7]):
        assert oracle.fast_atan2(float(y), float(x)) == cv2.fastAtan2(float(y), float(x)), (y, x)


@pytest.mark.parametrize("threshold", [1, 7, 20, 60])
@pytest.mark.parametrize("nonmax", [True, False])
def test_fast_9_16_whole_image_and_rois(oracle, synth, threshold, nonmax):
    left, _ = synth.stereo_pair(3)
    det = cv2.FastFeatureDetector_create(threshold=threshold, nonmaxSuppression=nonmax,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    rois = [left, left[16:54, 16:53], left[100:140, 500:538], left[5:12, 5:12], left[0:7, 0:50], left[0:6, 0:50]]
    for roi in rois:
        kps = det.detect(np.ascontiguousarray(roi))
        want = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        got = oracle.fast9_16(roi, threshold, nonmax)
        if not nonmax:  # cv2 leaves response = 0 when it does not need the score
            got, want = got[:, :2], want[:, :2]
        assert np.array_equal(got, want), roi.shape
        for k in kps[:3]:
            assert k.size == 7 and k.angle == -1 and k.octave == 0 and k.class_id == -1


def test_hamming_match_vs_bfmatcher(oracle):
    rng = np.random.default_rng(9)
    q = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    t[100] = t[50]          # exact duplicates: ties must resolve to the lowest trainIdx
    q[3] = t[50]
    q[4] = 0
    t[10] = 0
    t[20] = 0
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
    idx, dist = oracle.hamming_match(q, t)
    assert len(m) == len(q)
    for d in m:
        assert idx[d.queryIdx] == d.trainIdx and dist[d.queryIdx] == int(d.distance)
    assert idx[3] == 50 and dist[3] == 0 and idx[4] == 10


def _grid_fast_cv2(img, ini_th, min_th):
    """ComputeKeyPointsOctTree's cell loop (src/ORBextractor.cpp:826-883) with cv2.FAST per cell."""
    f = np.float32
    rows, cols = img.shape
    minB, maxBX, maxBY = 16, cols - 16, rows - 16
    width, height = f(maxBX - minB), f(maxBY - minB)
    nCols, nRows = int(width / f(30)), int(height / f(30))
    wCell, hCell = int(math.ceil(width / f(nCols))), int(math.ceil(height / f(nRows)))
    out = []
    d20 = cv2.FastFeatureDetector_create(ini_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    d7 = cv2.FastFeatureDetector_create(min_th, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    for i in range(nRows):
        iniY = minB + i * hCell
        maxY = min(iniY + hCell + 6, maxBY)
        if iniY >= maxBY - 3:
            continue
        for j in range(nCols):
            iniX = minB + j * wCell
            maxX = min(iniX + wCell + 6, maxBX)
            if iniX >= maxBX - 6:
                continue
            roi = np.ascontiguousarray(img[iniY:maxY, iniX:maxX])
            kps = d20.detect(roi) or d7.detect(roi)
            out += [(k.pt[0] + j * wCell, k.pt[1] + i * hCell, k.response) for k in kps]
    return np.array(out, np.float32).reshape(-1, 3)


@pytest.mark.parametrize("seed", [0, 5])
def test_grid_fast_candidates_vs_cv2_cells(oracle, synth, seed):
    left, _ = synth.stereo_pair(seed)
    ext = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    ext.DetectAndCompute(left)
    for level in (0, 3, 7):
        img = ext.level(level)
        _, cand = oracle.ORBextractor(50, 1.2, 8, 20, 7).DetectWithCandidates(img)
        assert np.array_equal(cand, _grid_fast_cv2(img, 20, 7)), level


def test_pyramid_and_blur_levels_vs_cv2(oracle, synth):
    left, _ = synth.stereo_pair(2)
    ext = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    ext.DetectAndCompute(left)
    prev = left
    for level in range(8):
        if level:
            s = ext.inv_scale[level]
            prev = cv2.resize(prev, (int(np.rint(np.float32(1241) * s)), int(np.rint(np.float32(376) * s))),
                              interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(ext.level(level), prev)
        assert np.array_equal(ext.blurred_level(level),
                              cv2.GaussianBlur(prev.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101))


def test_descriptor_pattern_sanity_vs_cv2_orb(oracle, synth):
    """cv2.ORB.compute blurs a sub-matrix (float Gaussian path) so it is NOT bit-exact (SURVEY §8c vi);
    it still shares the 256-pair pattern and the steering, so almost all bits agree."""
    left, _ = synth.stereo_pair(4)
    ext = oracle.ORBextractor(500, 1.2, 1, 20, 7)
    kps, desc = ext.DetectAndCompute(left)
    cvk = [cv2.KeyPoint(float(k["x"]), float(k["y"]), 31.0, float(k["angle"]), float(k["response"]), 0) for k in kps]
    orb = cv2.ORB_create(nfeatures=len(cvk), nlevels=1, edgeThreshold=19, patchSize=31)
    cvk2, d2 = orb.compute(left, cvk)
    assert len(cvk2) == len(cvk)
    bits = np.unpackbits(desc ^ d2).sum()
    assert bits / (desc.size * 8) < 0.01
