"""Parity of the CUDA brute-force Hamming matcher with the oracle (cv::BFMatcher semantics,
src/loopclosing.cpp:172): nearest train row per query row, ties -> lowest trainIdx."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def matcher(pkg):
    m = pkg.HammingMatcher(max_batch=8, max_rows=4096)
    yield m
    m.close()


def test_random_sets_with_ties(matcher, oracle):
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (2007, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (1999, 32), dtype=np.uint8)
    t[1500] = t[3]
    t[700] = t[3]
    q[5] = t[3]
    q[6] = 0
    t[1000] = 0
    t[10] = 0
    idx, dist = matcher.match(q, t)
    widx, wdist = oracle.hamming_match(q, t)
    assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
    assert idx[5] == 3 and dist[5] == 0 and idx[6] == 10


def test_ragged_batch_and_empty_sets(matcher, oracle):
    rng = np.random.default_rng(3)
    sizes = [(1, 1), (257, 3), (3, 513), (600, 0), (0, 50), (2000, 2000)]
    qs = [rng.integers(0, 256, (a, 32), dtype=np.uint8) for a, _ in sizes]
    ts = [rng.integers(0, 256, (b, 32), dtype=np.uint8) for _, b in sizes]
    res = matcher.match_batch(qs, ts)
    for (a, b), q, t, (idx, dist) in zip(sizes, qs, ts, res):
        assert len(idx) == a
        if b == 0:
            assert (idx == -1).all() and (dist == -1).all()
        else:
            widx, wdist = oracle.hamming_match(q, t)
            assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)


def test_left_right_descriptors_of_a_stereo_pair(pkg, matcher, oracle, synth):
    """BASELINE config 2: DetectAndCompute on both views, match(query=left, train=right)."""
    left, right = synth.stereo_pair(1)
    ext = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_batch=2)
    (kl, dl), (kr, dr) = ext.DetectAndComputeBatch([left, right])
    idx, dist = matcher.match(dl, dr)
    widx, wdist = oracle.hamming_match(dl, dr)
    assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
    # the synthetic right view is the left one shifted by a per-plane disparity: good matches are on the same row
    good = dist <= 30
    assert good.sum() > 500
    dy = np.abs(kl["y"][good] - kr["y"][idx[good]])
    assert np.median(dy) <= 1.0
    # LoopClosing::MatchFeatures filter (src/loopclosing.cpp:175-194) on top of the GPU matches
    pairs = oracle.match_filter(idx, dist, np.arange(len(kl), dtype=np.int32), np.arange(len(kr), dtype=np.int32))
    assert len(pairs) > 500
    ext.close()


def test_long_train_set_and_ties_across_tiles(pkg, oracle):
    """More train rows than one CTA slice / many 256-row tiles: the mbarrier phases of the TMA ring and of the two TMEM
    accumulators wrap several times; equal distances in different tiles and slices must still pick the lowest index."""
    rng = np.random.default_rng(7)
    m = pkg.HammingMatcher(max_batch=2, max_rows=6144)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (6000, 32), dtype=np.uint8)
    for k in (5, 300, 1023, 1024, 2999, 5999):      # copies of query rows spread over tiles and slices
        t[k] = q[17]
    t[4000:4100] = q[200]
    idx, dist = m.match(q, t)
    widx, wdist = oracle.hamming_match(q, t)
    assert np.array_equal(idx, widx) and np.array_equal(dist, wdist)
    assert idx[17] == 5 and dist[17] == 0 and idx[200] == 4000
    same = np.tile(rng.integers(0, 256, (1, 32), dtype=np.uint8), (5000, 1))   # every train row identical
    idx, dist = m.match(q, same)
    assert (idx == 0).all()
    res = m.match_batch([q, q[:129]], [t[:257], t])                            # two problems of different shape in one launch
    for (qq, tt), (i2, d2) in zip([(q, t[:257]), (q[:129], t)], res):
        wi, wd = oracle.hamming_match(qq, tt)
        assert np.array_equal(i2, wi) and np.array_equal(d2, wd)
    m.close()
