"""Development check of the tcgen05 Hamming kernel against the numpy oracle on random descriptors (prints mismatch statistics)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
pkg = importlib.import_module(PKG)
from oracle import oracle as O
rng = np.random.default_rng(0)
for nq, nt in ((128, 256), (300, 700), (2000, 2000)):
    q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    t[: min(nq, nt) // 2] = q[: min(nq, nt) // 2]        # exact matches for half of the queries
    m = pkg.HammingMatcher(max_batch=1, max_rows=max(nq, nt))
    idx, dist = m.match(q, t)
    widx, wdist = O.hamming_match(q, t)
    bad = np.flatnonzero((idx != widx) | (dist != wdist))
    print(f"nq {nq} nt {nt}: mismatches {len(bad)}")
    for i in bad[:8]:
        print("   row", i, "got", idx[i], dist[i], "want", widx[i], wdist[i])
