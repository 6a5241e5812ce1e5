"""B200-native hot path of the stereo SLAM system (ORB extract, Hamming match, local BA, DeepLCD
scoring, pose graph) — Python harness over the C-ABI product library ``libslamb200.so``.

The product is the shared library built from ``csrc/`` (C ABI declared in ``include/slamb200.h``) and
the C++ adaptor classes in ``host/``.  This package only binds the C ABI with ctypes for the tests,
``bench.py`` and ``__graft_entry__``; it contains no compute of its own and no CPU fallback.
"""
from .capi import (KP_DTYPE, SlamB200Error, ORBextractor, HammingMatcher, LocalBA, DeepLCDScorer, PoseGraph, StereoFrontend, PoseOnlyOptimizer, PnPRansac, LKTracker, DeepLCD, CALC_LAYERS, parse_caffe, triangulate, lib, lib_path, last_error)  # noqa: F401
