"""The two-tier formulation of k_fast_cells (csrc/orb.cu) against the reference's per-cell rule (src/ORBextractor.cpp:838-883:
cv::FAST(cell, iniThFAST, nms); if nothing was found, cv::FAST(cell, minThFAST, nms)) — on the CPU, through the oracle's FAST.

The kernel never calls FAST twice on a textured cell.  Tier 0 computes the threshold-independent response only where it can
reach iniTh and suppresses non-maxima among THOSE responses; tier 1 redoes a cell at minTh only when tier 0 kept nothing
there.  This test states that model in numpy (response plane from the oracle, 3 x 3 strict maxima) and holds it to the
two-call rule on cells of every kind: textured, flat, low contrast (fallback), and with equal neighbouring responses."""
import numpy as np
import pytest


def _maxima(plane):
    """strict 3 x 3 maxima of a response plane (0 = not a corner), as cv::FAST's non-maximum suppression keeps them"""
    h, w = plane.shape
    pad = np.zeros((h + 2, w + 2), plane.dtype)
    pad[1:-1, 1:-1] = plane
    nb = np.zeros_like(plane)
    for dy in (-1, 0, 1):
        for dx in (-1, 0, 1):
            if dy or dx:
                nb = np.maximum(nb, pad[1 + dy:1 + dy + h, 1 + dx:1 + dx + w])
    ys, xs = np.nonzero((plane > 0) & (plane > nb))
    return sorted(zip(xs.tolist(), ys.tolist(), plane[ys, xs].tolist()))


def _plane(oracle, roi, t):
    """responses >= t of the ROI (cv::FAST without suppression reports every corner with its response)"""
    p = np.zeros(roi.shape, np.int32)
    for x, y, r in oracle.fast9_16(roi, t, nonmax=False):
        p[y, x] = r
    return p


def two_tier(oracle, roi, ini, mn):
    tier0 = _maxima(_plane(oracle, roi, ini))          # responses below iniTh neither count nor suppress
    if tier0 or mn >= ini:
        return tier0
    return _maxima(_plane(oracle, roi, mn))            # the cell found nothing: the whole cell again at minTh


def reference_rule(oracle, roi, ini, mn):
    k = oracle.fast9_16(roi, ini, nonmax=True)
    if len(k) == 0:
        k = oracle.fast9_16(roi, mn, nonmax=True)
    return sorted(map(tuple, k.tolist()))


@pytest.mark.parametrize("ini,mn", [(20, 7), (12, 12), (7, 20), (40, 5)])
def test_two_tier_equals_the_two_fast_calls(oracle, ini, mn):
    rng = np.random.default_rng(ini * 100 + mn)
    kinds = 0
    for trial in range(120):
        kind = trial % 4
        if kind == 0:      # textured
            roi = rng.integers(0, 256, (36, 36), dtype=np.uint8)
        elif kind == 1:    # low contrast: corners exist at minTh only
            roi = (128 + rng.integers(-9, 10, (36, 36))).astype(np.uint8)
        elif kind == 2:    # blocks: plateaus of equal responses (strict maxima drop both of an equal pair)
            roi = (np.kron(rng.integers(0, 2, (9, 9)), np.ones((4, 4))) * rng.integers(15, 120) + 60).astype(np.uint8)
        else:              # flat with a few spikes
            roi = np.full((36, 36), 100, np.uint8)
            for _ in range(rng.integers(0, 4)):
                roi[rng.integers(4, 32), rng.integers(4, 32)] = rng.integers(101, 200)
        got, want = two_tier(oracle, roi, ini, mn), reference_rule(oracle, roi, ini, mn)
        assert got == want, (ini, mn, trial, kind, len(got), len(want))
        kinds |= 1 << (0 if not want else 1 if any(r < ini for _, _, r in want) else 2)
    if ini > mn:
        assert kinds & 2, "no trial exercised the minTh fallback"
