"""k_describe (csrc/orb.cu) stages every keypoint's neighbourhood with two TMA boxes whose innermost coordinate must be a
multiple of 16 bytes: 48 x 31 for the orientation disc (radius 15), 64 x 37 for the rBRIEF sampling window.  This test holds
the geometry those boxes rely on: how far the rotated test pattern (src/ORBextractor.cpp:101-359, :59-98) can reach, and that
an aligned box of that size always contains the window of a keypoint that passed the reference's 19-pixel border (:895-896)."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "a-simple-stereo-slam-system-with-deep-loop-closing_b200", "csrc")


def _defines():
    src = open(os.path.join(CSRC, "orb.cu")).read()
    return {k: int(v) for k, v in re.findall(r"#define (DESC_[A-Z]+)\s+(\d+)\b", src)}


def _pattern():
    txt = re.sub(r"/\*.*?\*/", "", open(os.path.join(CSRC, "orb_pattern.inc")).read(), flags=re.S)
    v = np.array([int(t) for t in re.findall(r"-?\d+", txt)], np.float32)
    assert len(v) == 1024
    return v.reshape(512, 2)


def test_rotated_pattern_stays_within_18_pixels():
    p = _pattern()
    reach = 0
    for deg in np.arange(0, 360, 0.25, dtype=np.float32):          # cvRound(x * b + y * a), cvRound(x * a - y * b) in fp32
        ang = np.float32(deg) * np.float32(np.pi / 180.0)
        a, b = np.float32(np.cos(np.float64(ang))), np.float32(np.sin(np.float64(ang)))
        dy = np.rint(p[:, 0] * b + p[:, 1] * a)
        dx = np.rint(p[:, 0] * a - p[:, 1] * b)
        reach = max(reach, float(np.abs(dx).max()), float(np.abs(dy).max()))
    assert reach == 18          # (-13, -13) at 45 degrees; the box must cover +-18, not the nominal 15-pixel patch radius


def test_aligned_boxes_contain_the_windows():
    d = _defines()
    aw, ah, cw, ch = d["DESC_AW"], d["DESC_AH"], d["DESC_CW"], d["DESC_CH"]
    assert aw % 16 == 0 and cw % 16 == 0 and ah == 31 and ch == 37
    for x in range(19, 4096):                                       # level x of a keypoint: >= EDGE_THRESHOLD = 19
        x0 = (x - 15) & ~15                                         # orientation box
        assert 0 <= x0 <= x - 15 and x + 15 < x0 + aw
        x0 = (x - 18) & ~15                                         # sampling box
        assert 0 <= x0 <= x - 18 and x + 18 < x0 + cw
    assert ah == 2 * 15 + 1 and ch == 2 * 18 + 1                    # rows y - 15 .. y + 15 / y - 18 .. y + 18, no alignment needed
