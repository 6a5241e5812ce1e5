"""Pins of oracle/calc_oracle.py (DeepLCD CNN forward, "next" row 2): OpenCV pre-processing against cv2 4.13.0
bit-exactly, Caffe layer arithmetic against torch's CPU fp32 operators."""
import importlib

import numpy as np
import pytest

from oracle import calc_oracle as CO

synth = importlib.import_module("a-simple-stereo-slam-system-with-deep-loop-closing_b200.synth")
cv2 = pytest.importorskip("cv2")


@pytest.mark.parametrize("shape", [(376, 1241), (120, 160), (97, 203), (480, 752)])
def test_preprocess_matches_cv2(shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    want_blur = cv2.GaussianBlur(img.copy(), (7, 7), 0)
    want_small = cv2.resize(want_blur, (160, 120))
    blur, small = CO.preprocess(img)
    assert np.array_equal(blur, want_blur)
    assert np.array_equal(small, want_small)


def test_sigma0_kernel_is_the_builtin_table():
    assert np.array_equal(cv2.getGaussianKernel(7, 0).ravel() * 256, CO.GAUSS7_SIGMA0)


def test_shapes_give_1064():
    s = CO.shapes()
    assert s[0] == (64, 62, 82) and s[2] == (64, 31, 41) and s[4] == (128, 32, 42) and s[6] == (128, 16, 21)
    assert s[-1] == (4, 14, 19) and int(np.prod(s[-1])) == 1064  # assert(p == 1064), src/deeplcd.cpp:82
    assert CO.n_weights() == synth.calc_weights(0).size


def test_layers_match_torch():
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    torch.set_num_threads(1)
    rng = np.random.default_rng(1)
    w = synth.calc_weights(3)
    (W1, b1), (W2, b2), (W3, b3) = CO.split_weights(w)
    x = rng.random((1, 120, 160), dtype=np.float32)
    t = torch.from_numpy(x)[None]
    # layer by layer so that a mismatch names its layer
    y = CO.conv(x, W1, b1, 2, 4)
    ty = F.conv2d(t, torch.from_numpy(W1), torch.from_numpy(b1), stride=2, padding=4)
    assert y.shape == tuple(ty.shape[1:]) and np.abs(y - ty[0].numpy()).max() < 1e-5
    y, ty = np.maximum(y, 0), F.relu(ty)
    y2 = CO.pool_max(y, 3, 2)
    ty2 = F.max_pool2d(ty, 3, 2, ceil_mode=True)
    assert y2.shape == tuple(ty2.shape[1:]) and np.array_equal(y2, F.max_pool2d(torch.from_numpy(y)[None], 3, 2, ceil_mode=True)[0].numpy())
    y3 = CO.lrn(y2, 5, 1e-4, 0.75, 1.0)
    ty3 = F.local_response_norm(torch.from_numpy(y2)[None], 5, alpha=1e-4, beta=0.75, k=1.0)
    assert np.abs(y3 - ty3[0].numpy()).max() < 1e-6
    # and the whole net
    d = CO.forward(x, w)
    t = F.local_response_norm(F.max_pool2d(F.relu(F.conv2d(t, torch.from_numpy(W1), torch.from_numpy(b1), stride=2, padding=4)), 3, 2, ceil_mode=True), 5, 1e-4, 0.75, 1.0)
    t = F.local_response_norm(F.max_pool2d(F.relu(F.conv2d(t, torch.from_numpy(W2), torch.from_numpy(b2), padding=2)), 3, 2, ceil_mode=True), 5, 1e-4, 0.75, 1.0)
    t = F.relu(F.conv2d(t, torch.from_numpy(W3), torch.from_numpy(b3)))
    want = t.reshape(-1).numpy()
    assert d.shape == (1064,) and np.abs(d - want).max() < 2e-5 * max(1.0, np.abs(want).max())


def test_pool_padded_and_lrn_wide():
    """Caffe's pooling rules with pad > 0 (last window must start inside in + pad) and an LRN wider than C."""
    torch = pytest.importorskip("torch")
    F = torch.nn.functional
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3, 13, 10)).astype(np.float32)
    for k, s, p in ((3, 2, 1), (2, 2, 1), (3, 3, 1), (3, 1, 0)):
        want = F.max_pool2d(torch.from_numpy(x)[None], k, s, padding=p, ceil_mode=True)[0].numpy()
        got = CO.pool_max(x, k, s, p)
        assert got.shape == want.shape and np.array_equal(got, want), (k, s, p)
    want = F.local_response_norm(torch.from_numpy(x)[None], 5, alpha=0.3, beta=0.6, k=2.0)[0].numpy()
    assert np.abs(CO.lrn(x, 5, 0.3, 0.6, 2.0) - want).max() < 1e-6


def test_descriptor_is_unit_norm_and_discriminative():
    w = synth.calc_weights(0)
    a, _ = synth.stereo_pair(5)
    b, _ = synth.stereo_pair(6)
    da, blurred = CO.calc_descr_original(a, w)
    db, _ = CO.calc_descr_original(b, w)
    noisy = np.clip(a.astype(np.int32) + np.random.default_rng(0).integers(-3, 4, a.shape), 0, 255).astype(np.uint8)
    dn, _ = CO.calc_descr_original(noisy, w)
    assert abs(float(np.linalg.norm(da)) - 1) < 1e-6 and blurred.shape == a.shape
    assert float(da @ dn) > float(da @ db)
