"""Seeded synthetic replay inputs: KITTI-shaped stereo pairs, BA windows,
DeepLCD databases and pose graphs.  numpy only; used by tests/, bench.py and smoke()."""
import numpy as np

KITTI_W, KITTI_H = 1241, 376
KITTI_FX = KITTI_FY = 718.856
KITTI_CX, KITTI_CY = 607.1928, 185.2157
KITTI_BF = 386.1448  # config/stereo/gray/KITTI00-02.yaml:35


def _blur3(img):
    """3x3 Gaussian, sigma 0.8, edge-replicated, float32 in/out."""
    k = np.exp(-np.arange(-1, 2) ** 2 / (2 * 0.8 ** 2)).astype(np.float32)
    k /= k.sum()
    p = np.pad(img, 1, mode="edge")
    t = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * t[:-2] + k[1] * t[1:-1] + k[2] * t[2:]


def stereo_pair(seed, w=KITTI_W, h=KITTI_H, n_rect=1500, noise=4, grid=(3, 8)):
    """Left/right u8 images of one synthetic stereo frame.

    Left: mid-grey canvas + `n_rect` filled random rectangles (w 4-60, h 4-40, grey 0-255) + 3x3
    Gaussian (sigma 0.8).  Depth is piecewise planar: a `grid` of fronto-parallel planes with
    z in [5, 80] m; the right view samples the left one at x + round(bf / z) inside each plane.
    Independent uniform noise of +-`noise` is then added to both views.
    """
    rng = np.random.default_rng(seed)
    canvas = np.full((h, w), 128, np.float32)
    rw = rng.integers(4, 61, n_rect)
    rh = rng.integers(4, 41, n_rect)
    x0 = rng.integers(-20, w, n_rect)
    y0 = rng.integers(-10, h, n_rect)
    g = rng.integers(0, 256, n_rect)
    for i in range(n_rect):
        canvas[max(0, y0[i]):min(h, y0[i] + rh[i]), max(0, x0[i]):min(w, x0[i] + rw[i])] = g[i]
    clean = _blur3(canvas)
    z = rng.uniform(5.0, 80.0, grid)
    disp = np.rint(KITTI_BF / z).astype(np.int64)
    ys = np.minimum(np.arange(h) * grid[0] // h, grid[0] - 1)
    xs = np.minimum(np.arange(w) * grid[1] // w, grid[1] - 1)
    src_x = np.clip(np.arange(w)[None, :] + disp[ys][:, xs], 0, w - 1)
    right_clean = np.take_along_axis(clean, src_x, axis=1)
    out = []
    for img in (clean, right_clean):
        img = img + rng.integers(-noise, noise + 1, (h, w)).astype(np.float32)
        out.append(np.clip(np.rint(img), 0, 255).astype(np.uint8))
    return out[0], out[1]


def stereo_batch(seed0, batch, w=KITTI_W, h=KITTI_H):
    """[batch, 2, h, w] u8: image (b, 0) = left, (b, 1) = right of frame seed0 + b."""
    out = np.empty((batch, 2, h, w), np.uint8)
    for b in range(batch):
        out[b, 0], out[b, 1] = stereo_pair(seed0 + b, w, h)
    return out
