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
    fe.close()
    fe2.close()
