"""The one-call stereo front end (sb_stereo_*: DetectAndCompute on both views + Hamming match) against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_extract_match_batch_bit_exact(pkg, oracle, synth):
    frames = synth.stereo_batch(300, 3)
    fe = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=4)
    out = fe.extract_match(frames)
    cpu = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    for p in range(3):
        descs = []
        for v in range(2):
            wk, wd = cpu.DetectAndCompute(frames[p, v])
            n = out["counts"][p, v]
            assert n == len(wk)
            assert out["kps"][p, v, :n].tobytes() == wk.tobytes()
            assert np.array_equal(out["desc"][p, v, :n], wd)
            descs.append(wd)
        widx, wdist = oracle.hamming_match(descs[0], descs[1])
        n = len(widx)
        assert np.array_equal(out["midx"][p, :n], widx) and np.array_equal(out["mdist"][p, :n], wdist)
    # two handles used alternately (the pipelined mode of bench.py's e2e leg) give the same answers
    fe2 = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=4)
    o1, o2 = fe.alloc_outputs(3, pinned=True), fe2.alloc_outputs(3, pinned=True)
    fe.submit(frames, o1)
    fe2.submit(frames[::-1].copy(), o2)
    fe.wait()
    fe2.wait()
    assert np.array_equal(o1["midx"], out["midx"]) and np.array_equal(o2["midx"][::-1], out["midx"])
    assert np.array_equal(o1["desc"], out["desc"])
    # ... and so do three handles whose kernels share ONE stream while their copies stay on their own
    # (sb_stereo_set_compute_stream), several batches each, results collected one round later
    import torch
    shared = torch.cuda.Stream()
    fe3 = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=4)
    hs = [fe, fe2, fe3]
    for h in hs:
        h.set_compute_stream(shared.cuda_stream)
    pinned = torch.from_numpy(np.stack([frames, frames[::-1].copy()])).pin_memory().numpy()
    os_ = [h.alloc_outputs(3, pinned=True) for h in hs]
    def same(got, flipped):                                  # rows past a frame's count are not defined
        for p in range(3):
            q = 2 - p if flipped else p
            n0, n1 = out["counts"][q]
            assert tuple(got["counts"][p]) == (n0, n1)
            assert np.array_equal(got["midx"][p, :n0], out["midx"][q, :n0]) and np.array_equal(got["mdist"][p, :n0], out["mdist"][q, :n0])
            assert np.array_equal(got["desc"][p, 0, :n0], out["desc"][q, 0, :n0]) and np.array_equal(got["desc"][p, 1, :n1], out["desc"][q, 1, :n1])
            assert got["kps"][p, 1, :n1].tobytes() == out["kps"][q, 1, :n1].tobytes()

    for rnd in range(3):
        for k, h in enumerate(hs):
            if rnd:
                h.wait()
                same(os_[k], (rnd - 1 + k) % 2 == 1)
            h.submit(pinned[(rnd + k) % 2], os_[k])
    for k, h in enumerate(hs):
        h.wait()
        same(os_[k], (2 + k) % 2 == 1)
    fe.set_compute_stream(None)                              # back to the default mode
    fe.submit(frames, o1)
    fe.wait()
    same(o1, False)
    fe.close()
    fe2.close()
    fe3.close()


def test_full_size_batch_properties(pkg, oracle, synth):
    """BASELINE-sized call (64 stereo pairs = 128 images, 2000 features): the oracle would need minutes, so the
    batch is checked through properties that do not depend on it: (a) a frame's result does not depend on its
    slot or its neighbours (the same frame planted in several slots gives identical bytes, and a sampled slot is
    bit-exact against the oracle), (b) every reported match distance is the popcount of the xor of the two rows
    and no train row is closer, ties resolved to the lowest index, (c) keypoints are level-major and inside the
    19-px border of their level."""
    base = synth.stereo_batch(700, 6)
    idx = np.array([i % 6 for i in range(64)])
    frames = np.ascontiguousarray(base[idx])
    fe = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=64)
    out = fe.extract_match(frames)
    for p in range(6, 64):                                   # (a) slot independence
        q = p % 6
        assert np.array_equal(out["counts"][p], out["counts"][q])
        n0, n1 = out["counts"][p]
        assert out["kps"][p, 0, :n0].tobytes() == out["kps"][q, 0, :n0].tobytes()
        assert np.array_equal(out["desc"][p, 1, :n1], out["desc"][q, 1, :n1])
        assert np.array_equal(out["midx"][p, :n0], out["midx"][q, :n0])
    cpu = oracle.ORBextractor(2000, 1.2, 8, 20, 7)
    wk, wd = cpu.DetectAndCompute(frames[63, 0])
    assert out["kps"][63, 0, :len(wk)].tobytes() == wk.tobytes() and np.array_equal(out["desc"][63, 0, :len(wk)], wd)
    pop = np.array([bin(i).count("1") for i in range(256)], np.int32)
    for p in (0, 17, 63):                                    # (b) matches are true nearest neighbours
        n0, n1 = out["counts"][p]
        dl, dr = out["desc"][p, 0, :n0], out["desc"][p, 1, :n1]
        for qi in range(0, n0, 97):
            d = pop[dl[qi][None, :] ^ dr].sum(1)
            assert out["mdist"][p, qi] == d.min() and out["midx"][p, qi] == int(np.argmin(d))
    scale = np.float32(1.2) ** np.arange(8)
    for p in (5, 40):                                        # (c) ordering and border
        k = out["kps"][p, 0, :out["counts"][p, 0]]
        assert np.all(np.diff(k["octave"]) >= 0)
        lvl_x, lvl_y = k["x"] / scale[k["octave"]], k["y"] / scale[k["octave"]]
        assert lvl_x.min() >= 18.9 and lvl_y.min() >= 18.9
        assert np.all(k["size"] == np.floor(31 * np.cumprod(np.r_[1.0, np.full(7, 1.2)]))[k["octave"]].astype(np.float32))
    fe.close()
