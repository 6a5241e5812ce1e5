"""The parity pin that comes from the reference itself.

oracle/_ref/libmyslam_orb_ref.so is /root/reference/src/ORBextractor.cpp compiled UNMODIFIED (recipe: oracle/Makefile
`ref`; OpenCV is replaced by the shim headers in oracle/ref_shim/, whose cv::FAST / resize / GaussianBlur / fastAtan2
are the C primitives that tests/test_oracle_cv2.py pins bit-exactly to cv2 4.13.0).  Held here, bit for bit:

    reference source (_ref)  ==  restatement (oracle/orb_oracle.c)  ==  committed goldens   [CPU, this file]
    CUDA path (C ABI)        ==  reference source (_ref)                                     [-m gpu, this file]

for every public operator of include/myslam/ORBextractor.h:47-138.

The one place where the reference is undefined is the quadtree's (size, ExtractorNode*) sort, src/ORBextractor.cpp:731:
equal sizes are ordered by heap address.  _ref is therefore run in two modes —
  * monotone: list nodes come from an address-monotone arena, i.e. "later-created node first", the documented rule of
    the restatement and of the CUDA path: must equal the oracle on every frame;
  * glibc: plain malloc.  The run records the address of every list node; the oracle re-run with those addresses as tie
    keys must reproduce it exactly — so the heap address is the ONLY degree of freedom between any real run of the
    reference and the restatement.
"""
import ctypes as C
import hashlib
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

HERE = os.path.dirname(os.path.abspath(__file__))
ORB_PARAMS = (2000, 1.2, 8, 20, 7)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def assert_same(got, want, ctx=""):
    gk, gd = got
    wk, wd = want
    assert gk.shape == wk.shape, (ctx, gk.shape, wk.shape)
    assert gk.tobytes() == wk.tobytes(), (ctx, "keypoints")
    assert np.array_equal(gd, wd), (ctx, "descriptors")


def _frontend_mask(shape, kps):
    """Frontend::DetectFeatures (src/frontend.cpp:305-309)."""
    mask = np.full(shape, 255, np.uint8)
    for k in kps:
        x, y = int(k["x"]), int(k["y"])
        mask[max(0, y - 20):y + 21, max(0, x - 20):x + 21] = 0
    return mask


def _expand_octaves(feats):
    """src/loopclosing.cpp:94-105: every feature as a keypoint on each of the 8 octaves."""
    kin = np.zeros(len(feats) * 8, feats.dtype)
    for i, f in enumerate(feats):
        for level in range(8):
            k = kin[i * 8 + level]
            k["x"], k["y"], k["size"], k["angle"], k["response"] = f["x"], f["y"], 7, -1, f["response"]
            k["octave"], k["class_id"] = level, i
    return kin


def _other_image(shape, seed):
    rng = np.random.default_rng(seed)
    img = np.full(shape, 120, np.uint8)
    for _ in range(400):
        x, y = rng.integers(0, shape[1] - 10), rng.integers(0, shape[0] - 10)
        img[y:y + rng.integers(3, 30), x:x + rng.integers(3, 30)] = rng.integers(0, 256)
    return (img.astype(np.int32) + rng.integers(-3, 4, shape)).clip(0, 255).astype(np.uint8)


def _low_texture_image():
    rng = np.random.default_rng(4)
    yy, xx = np.mgrid[0:376, 0:1241]
    img = np.full((376, 1241), 110.0)
    for _ in range(300):
        cx, cy, s, a = rng.uniform(0, 1241), rng.uniform(0, 376), rng.uniform(2, 5), rng.uniform(-35, 35)
        img += a * np.exp(-((xx - cx) ** 2 + (yy - cy) ** 2) / (2 * s * s))
    return np.clip(np.rint(img + rng.integers(-1, 2, img.shape)), 0, 255).astype(np.uint8)


# ------------------------------------------------------------------------------------------------------------------
# CPU: reference source == restatement == goldens
# ------------------------------------------------------------------------------------------------------------------
def test_ref_is_the_unmodified_reference_source(ref):
    """The recipe compiles the file where it lies; the hash list written next to the library names what went in."""
    path = os.path.join(os.path.dirname(ref._SO), "SOURCES.sha256")
    assert os.path.exists(path)
    lines = open(path).read().split("\n")
    assert any(l.endswith("src/ORBextractor.cpp") for l in lines) and any(l.endswith("myslam/ORBextractor.h") for l in lines)
    for l in lines:
        if l.strip() and os.path.exists(l.split()[1]):           # only where /root/reference exists (this container)
            assert hashlib.sha256(open(l.split()[1], "rb").read()).hexdigest() == l.split()[0]
    assert ref.lib().ref_list_node_bytes() > 0


@pytest.mark.parametrize("params", [ORB_PARAMS, (300, 1.2, 8, 20, 7), (100, 1.2, 8, 20, 7), (1200, 1.5, 5, 20, 7), (500, 2.5, 2, 20, 7)])
def test_constructor_tables(ref, oracle, params):
    """src/ORBextractor.cpp:384-445 through the reference's own getters."""
    r, o = ref.ORBextractor(*params), oracle.ORBextractor(*params)
    for name in ("scale", "inv_scale", "sigma2", "inv_sigma2", "quota", "umax"):
        assert getattr(r, name).tobytes() == getattr(o, name).tobytes(), name


def test_all_200_golden_frames_reference_equals_restatement_equals_golden(ref, oracle, synth):
    """BASELINE config 1: both views of all 200 frames through the reference's DetectAndCompute and the restatement's;
    keypoint and descriptor digests of both must be the committed ones (tests/golden/config1_digests.npz)."""
    with np.load(os.path.join(HERE, "golden", "config1_digests.npz")) as z:
        g = {k: z[k] for k in ("counts", "kp_digest", "desc_digest")}     # NpzFile is not thread-safe

    def one(f):
        r, o = ref.ORBextractor(*ORB_PARAMS), oracle.ORBextractor(*ORB_PARAMS)
        left, right = synth.stereo_pair(f)
        rl, rr, ol, orr = r.DetectAndCompute(left), r.DetectAndCompute(right), o.DetectAndCompute(left), o.DetectAndCompute(right)
        ok = rl[0].tobytes() == ol[0].tobytes() and rr[0].tobytes() == orr[0].tobytes() and np.array_equal(rl[1], ol[1]) \
            and np.array_equal(rr[1], orr[1])
        ok_g = (len(rl[0]), len(rr[0])) == tuple(g["counts"][f]) and np.array_equal(digest(rl[0], rr[0]), g["kp_digest"][f]) \
            and np.array_equal(digest(rl[1], rr[1]), g["desc_digest"][f])
        return ok, ok_g

    with ThreadPoolExecutor(os.cpu_count() or 1) as pool:
        res = list(pool.map(one, range(200)))
    assert [f for f, (a, _) in enumerate(res) if not a] == [], "reference source != restatement"
    assert [f for f, (_, b) in enumerate(res) if not b] == [], "reference source != committed golden digests"
    assert ref.lib().ref_error_count() == 0


def test_pyramid_levels(ref, oracle, synth):
    """public member mvImagePyramid (ORBextractor.h:106) after ComputePyramid (:1229-1265)."""
    r, o = ref.ORBextractor(*ORB_PARAMS), oracle.ORBextractor(*ORB_PARAMS)
    left, _ = synth.stereo_pair(3)
    r.DetectAndCompute(left)
    o.DetectAndCompute(left)
    for level in range(8):
        assert np.array_equal(r.level(level), o.level(level)), level


def test_masked_detect_and_compute_and_detect_with_pyramid(ref, oracle, synth):
    r, o = ref.ORBextractor(*ORB_PARAMS), oracle.ORBextractor(*ORB_PARAMS)
    left, _ = synth.stereo_pair(9)
    kps, _ = o.DetectAndCompute(left)
    mask = _frontend_mask(left.shape, kps[::7])
    assert_same(r.DetectAndCompute(left, mask), o.DetectAndCompute(left, mask), "masked")
    assert r.DetectWithPyramid(left, mask).tobytes() == o.DetectWithPyramid(left, mask).tobytes()
    assert r.DetectWithPyramid(left).tobytes() == o.DetectWithPyramid(left).tobytes()


@pytest.mark.parametrize("nfeatures", [300, 100, 1])
def test_detect_level0_with_frontend_mask(ref, oracle, synth, nfeatures):
    """ORBextractor::Detect (:989-1074) as the live front-end calls it (src/frontend.cpp:302-328)."""
    r, o = ref.ORBextractor(nfeatures, 1.2, 8, 20, 7), oracle.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    left, right = synth.stereo_pair(13)
    first = o.Detect(left)
    mask = _frontend_mask(left.shape, first[::2])
    assert r.Detect(left, mask).tobytes() == o.Detect(left, mask).tobytes()
    assert r.Detect(right).tobytes() == o.Detect(right).tobytes()
    assert r.Detect(left).tobytes() == first.tobytes()


def test_screen_params_and_calc_descriptors(ref, oracle, synth):
    """The loop-closing operators (:1083-1129, :1180-1226), incl. the mutated input vector (quirk Q5) and Q6's in-place rows."""
    r, o = ref.ORBextractor(100, 1.2, 8, 20, 7), oracle.ORBextractor(100, 1.2, 8, 20, 7)
    for seed in (17, 18):
        left, _ = synth.stereo_pair(seed)
        kin = _expand_octaves(oracle.ORBextractor(300, 1.2, 8, 20, 7).Detect(left))
        r_in, r_out = r.ScreenAndComputeKPsParams(left, kin)
        o_in, o_out = o.ScreenAndComputeKPsParams(left, kin)
        assert r_in.tobytes() == o_in.tobytes() and r_out.tobytes() == o_out.tobytes()
        assert 0 < len(r_out) < len(kin)
        assert np.array_equal(r.CalcDescriptors(left, r_out), o.CalcDescriptors(left, o_out))


@pytest.mark.parametrize("shape,nlevels", [((480, 640), 8), ((200, 333), 4), ((64, 70), 1), ((62, 400), 1), ((300, 1000), 3), ((97, 131), 2)])
def test_other_image_sizes(ref, oracle, shape, nlevels):
    img = _other_image(shape, shape[0])
    assert_same(ref.ORBextractor(500, 1.2, nlevels, 20, 7).DetectAndCompute(img), oracle.ORBextractor(500, 1.2, nlevels, 20, 7).DetectAndCompute(img), shape)


@pytest.mark.parametrize("params", [(1200, 1.5, 5, 20, 7), (800, 2.0, 3, 30, 10), (1500, 1.1, 6, 12, 5), (600, 1.2, 8, 40, 40), (500, 2.5, 2, 20, 7)])
def test_other_pyramid_and_threshold_parameters(ref, oracle, synth, params):
    left, _ = synth.stereo_pair(33)
    assert_same(ref.ORBextractor(*params).DetectAndCompute(left), oracle.ORBextractor(*params).DetectAndCompute(left), params)


def test_low_texture_and_flat_images(ref, oracle):
    img = _low_texture_image()
    got = ref.ORBextractor(1000, 1.2, 8, 20, 7).DetectAndCompute(img)
    assert_same(got, oracle.ORBextractor(1000, 1.2, 8, 20, 7).DetectAndCompute(img))
    assert (got[0]["response"] < 20).any() and (got[0]["response"] >= 20).any()      # the 20 -> 7 fallback ran (:858-865)
    k, d = ref.ORBextractor(*ORB_PARAMS).DetectAndCompute(np.full((376, 1241), 77, np.uint8))
    assert len(k) == 0 and len(d) == 0


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 900), st.integers(1, 400), st.sampled_from([(1209, 344), (400, 400), (300, 90), (64, 200)]))
def test_distribute_octtree_on_random_candidate_lists(ref, oracle, seed, n, N, extent):
    """DistributeOctTree (:586-810) directly: integer coordinates (FAST positions), few distinct responses (ties)."""
    rng = np.random.default_rng(seed)
    w, h = extent
    if round(np.float32(w) / np.float32(h)) < 1:
        return                                                  # nIni = 0: the reference divides by zero
    xy = np.unique(np.stack([rng.integers(0, w, n), rng.integers(0, h, n)], 1), axis=0)
    xy = xy[np.lexsort((xy[:, 0], xy[:, 1]))]                    # row-major, like the grid scan
    kx, ky = xy[:, 0].astype(np.float32), xy[:, 1].astype(np.float32)
    kr = rng.integers(7, 40, len(xy)).astype(np.float32)
    r = ref.ORBextractor(*ORB_PARAMS)
    got = r.distribute_octtree(kx, ky, kr, 16, 16 + w, 16, 16 + h, N)
    idx = oracle.distribute_octtree(kx, ky, kr, 16, 16 + w, 16, 16 + h, N)
    want = np.stack([kx[idx], ky[idx], kr[idx]], 1) if len(idx) else np.zeros((0, 3), np.float32)
    assert np.array_equal(got, want)


def test_glibc_heap_order_is_the_only_freedom(ref, oracle, synth):
    """_ref with plain malloc: record each list node's address, replay them as the restatement's tie keys -> identical.
    Also measures how often the heap's order differs from "later-created first" (it does on every frame: equal node
    sizes are everywhere at 2000 features), and that the monotone build is what the restatement's default rule gives."""
    g, m, o = ref.ORBextractor(*ORB_PARAMS, monotone_nodes=False), ref.ORBextractor(*ORB_PARAMS), oracle.ORBextractor(*ORB_PARAMS)
    RL, OL = ref.lib(), oracle.lib()
    buf = np.zeros(1 << 18, np.uint64)
    differs, sym_diff = 0, []
    for f in range(6):
        for img in synth.stereo_pair(f):
            RL.ref_record_node_addresses(buf.ctypes.data_as(C.c_void_p), len(buf))
            got = g.DetectAndCompute(img)
            n = RL.ref_recorded_count()
            RL.ref_record_node_addresses(None, 0)
            assert 0 < n <= len(buf)
            OL.orc_debug_set_tie_keys(buf.ctypes.data_as(C.c_void_p), n)
            try:
                want = o.DetectAndCompute(img)
                assert OL.orc_debug_tie_keys_used() == n          # same number of list nodes, in the same order
            finally:
                OL.orc_debug_set_tie_keys(None, 0)
            assert_same(got, want, ("glibc replay", f))
            default = o.DetectAndCompute(img)
            assert_same(m.DetectAndCompute(img), default, ("monotone", f))
            differs += got[0].tobytes() != default[0].tobytes()
            sym_diff.append(len(set(map(bytes, got[0])) ^ set(map(bytes, default[0]))))
    print(f"glibc heap order != later-created-first on {differs}/12 images; keypoints that differ per image: "
          f"min {min(sym_diff)}, max {max(sym_diff)} of ~2000")
    assert max(sym_diff) < 400                                    # same quadtree, different choice inside ties only


# ------------------------------------------------------------------------------------------------------------------
# GPU: CUDA path (through the C ABI) == reference source
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_cuda_detect_and_compute_equals_reference_source(pkg, ref, synth):
    B = 12
    ext = pkg.ORBextractor(*ORB_PARAMS, max_batch=2 * B)
    r = ref.ORBextractor(*ORB_PARAMS)
    imgs = [im for f in range(40, 40 + B) for im in synth.stereo_pair(f)]
    for b, (img, got) in enumerate(zip(imgs, ext.DetectAndComputeBatch(imgs))):
        assert_same(got, r.DetectAndCompute(img), b)
    left = imgs[0]
    mask = _frontend_mask(left.shape, r.DetectAndCompute(left)[0][::7])
    assert_same(ext.DetectAndComputeBatch([left], [mask])[0], r.DetectAndCompute(left, mask), "masked")
    assert ext.DetectWithPyramid(left).tobytes() == r.DetectWithPyramid(left).tobytes()
    ext.close()


@pytest.mark.gpu
@pytest.mark.parametrize("nfeatures", [300, 100])
def test_cuda_detect_screen_calc_equal_reference_source(pkg, ref, synth, nfeatures):
    g = pkg.ORBextractor(nfeatures, 1.2, 8, 20, 7, max_batch=2)
    r = ref.ORBextractor(nfeatures, 1.2, 8, 20, 7)
    left, right = synth.stereo_pair(13)
    first = r.Detect(left)
    mask = _frontend_mask(left.shape, first[::2])
    got = g.DetectBatch([left, right], [mask, None])
    assert got[0].tobytes() == r.Detect(left, mask).tobytes() and got[1].tobytes() == r.Detect(right).tobytes()
    kin = _expand_octaves(ref.ORBextractor(300, 1.2, 8, 20, 7).Detect(left))
    g_in, g_out = g.ScreenAndComputeKPsParams(left, kin)
    r_in, r_out = r.ScreenAndComputeKPsParams(left, kin)
    assert g_in.tobytes() == r_in.tobytes() and g_out.tobytes() == r_out.tobytes()
    assert np.array_equal(g.CalcDescriptors(left, r_out), r.CalcDescriptors(left, r_out))
    g.close()


@pytest.mark.gpu
def test_cuda_other_sizes_and_low_texture_equal_reference_source(pkg, ref):
    for shape, nlevels in [((480, 640), 8), ((200, 333), 4), ((97, 131), 2)]:
        img = _other_image(shape, shape[0])
        g = pkg.ORBextractor(500, 1.2, nlevels, 20, 7, max_w=shape[1], max_h=shape[0])
        assert_same(g.DetectAndCompute(img), ref.ORBextractor(500, 1.2, nlevels, 20, 7).DetectAndCompute(img), shape)
        g.close()
    img = _low_texture_image()
    g = pkg.ORBextractor(1000, 1.2, 8, 20, 7)
    assert_same(g.DetectAndCompute(img), ref.ORBextractor(1000, 1.2, 8, 20, 7).DetectAndCompute(img), "low texture")
    g.close()
