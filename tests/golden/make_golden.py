"""Generates tests/golden/config1_digests.npz — the committed correctness anchor of BASELINE config 1 — and
tests/golden/config3_ba.npz (local BA results of four seeded windows).

Config 1 is "KITTI-00 first 200 frames through the CPU reference path".  KITTI is not available, so the anchor is 200 seeded
synthetic stereo pairs (synth.stereo_pair(seed), seed = frame index).  Keypoints and descriptors come from THE REFERENCE'S
OWN EXTRACTOR — oracle/_ref, i.e. /root/reference/src/ORBextractor.cpp compiled unmodified (oracle/Makefile `ref`; list nodes
from the address-monotone arena, see oracle/ref_shim/ref_capi.cpp) — and this script refuses to write unless the restatement
in oracle/orb_oracle.c gives the same bytes.  The match result comes from the restatement of cv::BFMatcher (OpenCV is
un-vendored; pinned against cv2 4.13.0 by tests/test_oracle_cv2.py).  Per frame: the SHA-256 of the keypoint records, of the
descriptors and of the left->right match result, plus the full arrays of frames 0 and 1.  Needs /root/reference (this
container).  Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import importlib
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
N_FRAMES = 200
ORB_PARAMS = (2000, 1.2, 8, 20, 7)


def digest(*arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8)


def frame_record(O, R, synth, seed):
    ext = R.ORBextractor(*ORB_PARAMS)           # the reference source itself
    chk = O.ORBextractor(*ORB_PARAMS)           # the restatement must agree before anything is written
    left, right = synth.stereo_pair(seed)
    kl, dl = ext.DetectAndCompute(left)
    kr, dr = ext.DetectAndCompute(right)
    for img, k, d in ((left, kl, dl), (right, kr, dr)):
        ck, cd = chk.DetectAndCompute(img)
        assert ck.tobytes() == k.tobytes() and np.array_equal(cd, d), f"frame {seed}: restatement != reference source"
    idx, dist = O.hamming_match(dl, dr)
    return dict(kl=kl, dl=dl, kr=kr, dr=dr, idx=idx, dist=dist)


def main():
    synth = importlib.import_module(PKG + ".synth")
    from oracle import oracle as O, ref as R
    O.build()
    R.lib()
    with ThreadPoolExecutor(os.cpu_count() or 1) as pool:
        recs = list(pool.map(lambda s: frame_record(O, R, synth, s), range(N_FRAMES)))
    out = {
        "orb_params": np.array(ORB_PARAMS, np.float64),
        "counts": np.array([[len(r["kl"]), len(r["kr"])] for r in recs], np.int32),
        "kp_digest": np.stack([digest(r["kl"], r["kr"]) for r in recs]),
        "desc_digest": np.stack([digest(r["dl"], r["dr"]) for r in recs]),
        "match_digest": np.stack([digest(r["idx"].astype(np.int32), r["dist"].astype(np.int32)) for r in recs]),
    }
    for f in (0, 1):
        for k, v in recs[f].items():
            out[f"f{f}_{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "config1_digests.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", int(out["counts"].sum()), "keypoints in", N_FRAMES, "stereo frames")
    # config 3: local BA of four seeded windows (synth.ba_window) through the g2o-faithful restatement (fp64)
    ba = {}
    for seed in range(4):
        w = synth.ba_window(seed)
        p, x, chi2, outl, info = O.ba_solve(w["poses0"], w["points0"], w["fixed"], w["obs_pose"], w["obs_point"], w["uv"], synth.KITTI_K)
        ba[f"w{seed}_poses"], ba[f"w{seed}_points"], ba[f"w{seed}_info"] = p, x, np.asarray(info, np.int32)
        ba[f"w{seed}_outlier"] = np.asarray(outl, np.uint8)
    path = os.path.join(ROOT, "tests", "golden", "config3_ba.npz")
    np.savez_compressed(path, **ba)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
