import importlib, os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
left, right = synth.stereo_pair(0)
ext = pkg.ORBextractor(2000, 1.2, 8, 20, 7, max_batch=2)
L = pkg.lib()
cell = int(sys.argv[1]) if len(sys.argv) > 1 else 1
tb = C.c_int()
assert L.sb_orb_debug_fast_cell(ext._h, cell, None, 0, C.byref(tb)) == 0
res = ext.DetectAndComputeBatch([left, right])
buf = np.zeros(2 * tb.value + 64, np.uint8)
assert L.sb_orb_debug_fast_cell(ext._h, cell, C.c_void_p(buf.ctypes.data), buf.size, C.byref(tb)) == 0
info = buf[2 * tb.value:].view(np.int32)[:10]
BW, BH, xo, x0, y0, rw, rh, s_n, s_any, level = info
print("info", info)
tile = buf[:BW * BH].reshape(BH, BW)
sc = buf[tb.value:tb.value + BW * BH].reshape(BH, BW)
want = left[y0:y0 + BH, x0 - xo:x0 - xo + BW]
print("tile matches image:", np.array_equal(tile[:, :want.shape[1]], want))
if not np.array_equal(tile[:, :want.shape[1]], want):
    bad = np.argwhere(tile[:, :want.shape[1]] != want)
    print("first mismatches", bad[:10], tile[bad[0][0], :20], want[bad[0][0], :20])
from oracle import oracle as O
roi = np.ascontiguousarray(left[y0:y0 + rh, x0:x0 + rw])
o = O.fast9_16(roi, 7, False)
plane = np.zeros((rh, rw), np.int32)
plane[o[:, 1], o[:, 0]] = o[:, 2]
got = sc[:rh, xo:xo + rw].astype(np.int32)
print("score plane matches:", np.array_equal(got, plane), "nonzero got/want", (got > 0).sum(), (plane > 0).sum())
if not np.array_equal(got, plane):
    bad = np.argwhere(got != plane)
    for y, x in bad[:10]:
        print((x, y), got[y, x], plane[y, x])
