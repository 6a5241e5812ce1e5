"""Parity of the CUDA ORB extractor (through the C ABI) with the CPU oracle — bit-exact keypoints
(x, y, size, angle, response, octave, class_id) and descriptors.  Reference: src/ORBextractor.cpp."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PARAMS = (2000, 1.2, 8, 20, 7)


@pytest.fixture(scope="module")
def gpu_ext(pkg):
    ext = pkg.ORBextractor(*PARAMS, max_w=1241, max_h=376, max_batch=6)
    yield ext
    ext.close()


@pytest.fixture(scope="module")
def cpu_ext(oracle):
    return oracle.ORBextractor(*PARAMS)


def assert_kps_equal(got, want, ctx=""):
    assert got.shape == want.shape, (ctx, got.shape, want.shape)
    for f in want.dtype.names:
        bad = np.nonzero(got[f].view(np.uint32) != want[f].view(np.uint32))[0] if got[f].dtype.kind == "f" else \
            np.nonzero(got[f] != want[f])[0]
        assert len(bad) == 0, (ctx, f, bad[:5], got[bad[:5]], want[bad[:5]])


def test_pyramid_blur_and_candidates_per_level(gpu_ext, cpu_ext, oracle, synth):
    left, right = synth.stereo_pair(0)
    gpu_ext.DetectAndComputeBatch([left, right])
    for b, img in enumerate((left, right)):
        cpu_ext.DetectAndCompute(img)
        for level in range(8):
            assert np.array_equal(gpu_ext.debug_level(b, level, 0), cpu_ext.level(level)), ("pyramid", b, level)
            assert np.array_equal(gpu_ext.debug_level(b, level, 1), cpu_ext.blurred_level(level)), ("blur", b, level)
            _, cand = oracle.ORBextractor(10, 1.2, 8, 20, 7).DetectWithCandidates(cpu_ext.level(level))
            got = gpu_ext.debug_candidates(b, level)
            want = cand.astype(np.int64)
            got = got[np.lexsort((got[:, 0], got[:, 1]))]
            want = want[np.lexsort((want[:, 0], want[:, 1]))]
            assert np.array_equal(got, want), ("candidates", b, level, len(got), len(want))


@pytest.mark.parametrize("seed0", [0, 100])
def test_detect_and_compute_batch_bit_exact(gpu_ext, cpu_ext, synth, seed0):
    imgs = []
    for s in range(seed0, seed0 + 3):
        imgs += list(synth.stereo_pair(s))
    res = gpu_ext.DetectAndComputeBatch(imgs)
    for b, img in enumerate(imgs):
        wk, wd = cpu_ext.DetectAndCompute(img)
        gk, gd = res[b]
        assert_kps_equal(gk, wk, ("kps", seed0, b))
        assert np.array_equal(gd, wd), ("desc", seed0, b)
        assert len(gk) >= 1900


def test_handle_is_reusable_and_deterministic(gpu_ext, synth):
    left, _ = synth.stereo_pair(42)
    a = gpu_ext.DetectAndCompute(left)
    b = gpu_ext.DetectAndCompute(left)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


@pytest.mark.parametrize("nfeatures", [300, 100, 1])
def test_other_feature_budgets(pkg, oracle, synth, nfeatures):
    left, _ = synth.stereo_pair(5)
    g = pkg.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    c = oracle.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    assert np.array_equal(g.quota, c.quota) and np.array_equal(g.scale, c.scale)
    assert np.array_equal(g.inv_sigma2, c.inv_sigma2)
    gk, gd = g.DetectAndCompute(left)
    wk, wd = c.DetectAndCompute(left)
    assert_kps_equal(gk, wk)
    assert np.array_equal(gd, wd)
    g.close()


def _frontend_mask(shape, kps):
    """Frontend::DetectFeatures (src/frontend.cpp:305-309): 255 with 41x41 zero squares around existing features."""
    mask = np.full(shape, 255, np.uint8)
    for k in kps:
        x, y = int(k["x"]), int(k["y"])
        mask[max(0, y - 20):y + 21, max(0, x - 20):x + 21] = 0
    return mask


def test_masked_detect_and_compute(gpu_ext, cpu_ext, synth):
    left, right = synth.stereo_pair(9)
    kps, _ = cpu_ext.DetectAndCompute(left)
    mask = _frontend_mask(left.shape, kps[::7])
    res = gpu_ext.DetectAndComputeBatch([left, right], [mask, None])
    for img, m, (gk, gd) in zip((left, right), (mask, None), res):
        wk, wd = cpu_ext.DetectAndCompute(img, m if m is not None else np.full(img.shape, 255, np.uint8))
        assert_kps_equal(gk, wk)
        assert np.array_equal(gd, wd)
    assert len(res[0][0]) < len(res[1][0]) + 50


@pytest.mark.parametrize("nfeatures", [300, 100])
def test_detect_level0_with_frontend_mask(pkg, oracle, synth, nfeatures):
    """ORBextractor::Detect as the live front-end calls it (src/frontend.cpp:302-328)."""
    g = pkg.ORBextractor(nfeatures, 1.2, 8, 20, 7, max_batch=2)
    c = oracle.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    left, right = synth.stereo_pair(13)
    first = c.Detect(left)
    mask = _frontend_mask(left.shape, first[::2])
    got = g.DetectBatch([left, right], [mask, None])
    assert_kps_equal(got[0], c.Detect(left, mask))
    assert_kps_equal(got[1], c.Detect(right))
    assert (got[0]["size"] == 7).all() and (got[0]["angle"] == -1).all() and (got[0]["octave"] == 0).all()
    g.close()


def test_detect_with_pyramid(gpu_ext, cpu_ext, synth):
    left, _ = synth.stereo_pair(21)
    assert_kps_equal(gpu_ext.DetectWithPyramid(left), cpu_ext.DetectWithPyramid(left))


def test_screen_params_and_calc_descriptors(pkg, oracle, synth):
    """The loop-closing path: each feature expanded to 8 octaves (src/loopclosing.cpp:94-105), screened, described."""
    g = pkg.ORBextractor(100, 1.2, 8, 20, 7)
    c = oracle.ORBextractor(100, 1.2, 8, 20, 7)
    left, _ = synth.stereo_pair(17)
    feats = oracle.ORBextractor(300, 1.2, 8, 20, 7).Detect(left)
    kin = np.zeros(len(feats) * 8, feats.dtype)
    for i, f in enumerate(feats):
        for level in range(8):
            k = kin[i * 8 + level]
            k["x"], k["y"], k["size"], k["angle"], k["response"] = f["x"], f["y"], 7, -1, f["response"]
            k["octave"], k["class_id"] = level, i
    g_in, g_out = g.ScreenAndComputeKPsParams(left, kin)
    c_in, c_out = c.ScreenAndComputeKPsParams(left, kin)
    assert_kps_equal(g_in, c_in, "mutated input")
    assert_kps_equal(g_out, c_out, "survivors")
    assert 0 < len(g_out) < len(kin)
    assert np.array_equal(g.CalcDescriptors(left, c_out), c.CalcDescriptors(left, c_out))
    g.close()


def test_batched_screen_and_describe_equals_the_two_single_image_calls(pkg, oracle, synth):
    """sb_orb_screen_describe (LoopClosing::ProcessNewKF for a batch of keyframes) against the oracle's two calls per image."""
    g = pkg.ORBextractor(100, 1.2, 8, 20, 7, max_batch=3)
    c = oracle.ORBextractor(100, 1.2, 8, 20, 7)
    imgs, kins = [], []
    for seed, nf in ((17, 300), (18, 120), (19, 0)):
        left, _ = synth.stereo_pair(seed)
        feats = oracle.ORBextractor(max(nf, 1), 1.2, 8, 20, 7).Detect(left)[:nf]
        kin = np.zeros(len(feats) * 8, feats.dtype)
        for i, f in enumerate(feats):
            for level in range(8):
                k = kin[i * 8 + level]
                k["x"], k["y"], k["size"], k["angle"], k["response"], k["octave"], k["class_id"] = f["x"], f["y"], 7, -1, -1, level, i
        imgs.append(left)
        kins.append(kin)
    got = g.ScreenAndDescribeBatch(imgs, kins)
    for b in range(3):
        if len(kins[b]) == 0:
            assert len(got[b][1]) == 0 and len(got[b][2]) == 0
            continue
        c_in, c_out = c.ScreenAndComputeKPsParams(imgs[b], kins[b])
        assert_kps_equal(got[b][0], c_in, ("mutated input", b))
        assert_kps_equal(got[b][1], c_out, ("survivors", b))
        assert np.array_equal(got[b][2], c.CalcDescriptors(imgs[b], c_out)), b
        assert 0 < len(c_out) < len(kins[b])
    g.close()


# (64, 70): one FAST cell per level-0 grid row; (62, 400): a single grid row of cells; (300, 1000): runs of 5 cells + a rest;
# (97, 131): cells wider than 40 px (short runs)
@pytest.mark.parametrize("shape,nlevels", [((480, 640), 8), ((200, 333), 4), ((376, 1241), 1), ((64, 70), 1), ((62, 400), 1),
                                           ((300, 1000), 3), ((97, 131), 2)])
def test_other_image_sizes(pkg, oracle, shape, nlevels):
    rng = np.random.default_rng(shape[0])
    img = np.full(shape, 120, np.uint8)
    for _ in range(400):
        x, y = rng.integers(0, shape[1] - 10), rng.integers(0, shape[0] - 10)
        img[y:y + rng.integers(3, 30), x:x + rng.integers(3, 30)] = rng.integers(0, 256)
    img = (img.astype(np.int32) + rng.integers(-3, 4, shape)).clip(0, 255).astype(np.uint8)
    g = pkg.ORBextractor(500, 1.2, nlevels, 20, 7, max_w=shape[1], max_h=shape[0])
    c = oracle.ORBextractor(500, 1.2, nlevels, 20, 7)
    gk, gd = g.DetectAndCompute(img)
    wk, wd = c.DetectAndCompute(img)
    assert_kps_equal(gk, wk)
    assert np.array_equal(gd, wd)
    g.close()


def test_flat_image_gives_no_keypoints(gpu_ext):
    img = np.full((376, 1241), 77, np.uint8)
    k, d = gpu_ext.DetectAndCompute(img)
    assert len(k) == 0 and len(d) == 0


def test_errors_are_reported_not_thrown_away(pkg):
    with pytest.raises(pkg.SlamB200Error) as e:
        pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_w=100, max_h=100)   # too small for 8 levels
    assert e.value.code == -1
    g = pkg.ORBextractor(100, 1.2, 2, 20, 7, max_w=320, max_h=240)
    with pytest.raises(pkg.SlamB200Error):
        g.DetectAndCompute(np.zeros((376, 1241), np.uint8))            # larger than max_w x max_h
    g.close()


# scale factor 2.5: four destination pixels span more than 8 source bytes, so the pyramid falls back from k_resize_quads to
# the generic k_resize
@pytest.mark.parametrize("params", [(1200, 1.5, 5, 20, 7), (800, 2.0, 3, 30, 10), (1500, 1.1, 6, 12, 5), (600, 1.2, 8, 40, 40),
                                    (500, 2.5, 2, 20, 7)])
def test_other_pyramid_and_threshold_parameters(pkg, oracle, synth, params):
    left, right = synth.stereo_pair(33)
    g = pkg.ORBextractor(*params, max_batch=2)
    c = oracle.ORBextractor(*params)
    assert np.array_equal(g.quota, c.quota) and np.array_equal(g.scale, c.scale) and np.array_equal(g.inv_scale, c.inv_scale)
    for img, (gk, gd) in zip((left, right), g.DetectAndComputeBatch([left, right])):
        wk, wd = c.DetectAndCompute(img)
        assert_kps_equal(gk, wk, params)
        assert np.array_equal(gd, wd)
    g.close()


def test_low_texture_images_exercise_the_threshold_fallback(pkg, oracle):
    """Soft blobs only: most cells find nothing at iniThFAST = 20 and fall back to minThFAST = 7 (src/ORBextractor.cpp:858-865)."""
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:376, 0:1241]
    img = np.full((376, 1241), 110.0)
    for _ in range(300):
        cx, cy, s, a = rng.uniform(0, 1241), rng.uniform(0, 376), rng.uniform(2, 5), rng.uniform(-35, 35)
        img += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    img = np.clip(np.rint(img + rng.integers(-1, 2, img.shape)), 0, 255).astype(np.uint8)
    g = pkg.ORBextractor(1000, 1.2, 8, 20, 7)
    c = oracle.ORBextractor(1000, 1.2, 8, 20, 7)
    gk, gd = g.DetectAndCompute(img)
    wk, wd = c.DetectAndCompute(img)
    assert_kps_equal(gk, wk)
    assert np.array_equal(gd, wd)
    assert 0 < len(gk) and (gk["response"] < 20).any() and (gk["response"] >= 20).any()
    g.close()


def test_crowded_levels_keep_their_candidate_state_in_global_memory(pkg, oracle):
    """A level with more candidates than the quadtree kernel keeps in shared memory (QT_CAND_SMEM = 6144, hard cap 16384):
    levels 0 and 1 of a frame that is one third dense noise have about 12 000 and 7 000 — the spill path must give the
    same keypoints as the oracle, next to ordinary levels in the same launch and next to an ordinary frame in the batch."""
    rng = np.random.default_rng(5)
    img = np.full((376, 1241), 120, np.uint8)
    img[:, :434] = rng.integers(60, 200, (376, 434), dtype=np.uint8)
    plain = np.ascontiguousarray(img[:, ::-1] // 2 + 60)
    ext = pkg.ORBextractor(*PARAMS, max_w=1241, max_h=376, max_batch=2)
    cpu = oracle.ORBextractor(*PARAMS)
    got = ext.DetectAndComputeBatch([img, plain])
    assert len(ext.debug_candidates(0, 0)) > 6144 and len(ext.debug_candidates(0, 1)) > 6144
    assert len(ext.debug_candidates(0, 2)) < 6144
    for (gk, gd), im in zip(got, (img, plain)):
        wk, wd = cpu.DetectAndCompute(im)
        assert_kps_equal(gk, wk)
        assert np.array_equal(gd, wd)
    ext.close()
