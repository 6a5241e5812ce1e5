import ctypes as C, importlib, os, sys, numpy as np, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
capi = importlib.import_module(PKG + ".capi")
capi.lib_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("SB_PROF_LIB", "libslamb200_prof.so"))
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
ba = pkg.LocalBA(max_windows=64, max_poses=7, max_points=320, max_obs=2304)
ws = [synth.ba_window(s) for s in range(8)]
ba.solve(ws, synth.KITTI_K)
prof = (C.c_longlong * 16)()
pkg.lib().sb_ba_debug_profile(prof)
res = ba.solve(ws[:1], synth.KITTI_K); print("info", res[0][4])
pkg.lib().sb_ba_debug_profile(prof)
names = ["loop/accept", "errors", "build (landmark pass)", "lambda+push+Dinv", "schur", "cholesky+solve (rest)", "xl+update", "errors(trial)",
         "  cholesky: diagonal block + panel", "  cholesky: trailing update", "  substitutions", "  build: pose pass"]
tot = sum(prof[:12])
for n, v in zip(names, prof[:12]):
    print(f"{n:18s} {v:10d} cyc  {100*v/tot:5.1f}%")
print("total cycles", tot, "=", tot / 1.965e3, "us at 1965 MHz")
