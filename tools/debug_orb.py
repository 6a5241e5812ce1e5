"""Runs one DetectAndCompute on cuda:0 (for compute-sanitizer / ncu sessions)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
left, right = synth.stereo_pair(0)
ext = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_batch=2)
res = ext.DetectAndComputeBatch([left, right])
print([len(k) for k, d in res])
