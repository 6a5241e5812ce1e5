"""Per-phase clock64 profile of k_posegraph (build the library with -DPG_PROFILE into tools/libslamb200_prof.so)."""
import ctypes as C, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
capi = importlib.import_module(PKG + ".capi")
capi.lib_path = lambda: os.path.join(os.path.dirname(os.path.abspath(__file__)), os.environ.get("SB_PROF_LIB", "libslamb200_prof.so"))
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
g = synth.pose_graph(0, n=742)
pg = pkg.PoseGraph(1024, 2048)
pg.solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"])
prof = (C.c_longlong * 16)()
pkg.lib().sb_posegraph_debug_profile(prof)
_, info = pg.solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"])
pkg.lib().sb_posegraph_debug_profile(prof)
print("info", info)
names = ["errors", "numeric Jacobians", "assembly", "push + system + right-hand sides", "cyclic reduction, forward", "cyclic reduction, backward",
         "capacitance + Cholesky + x", "update / trial errors"]
names += ["  CR forward: phase a (rhs of survivors, 6x6 inverses)", "  CR forward: phase b (y = G r, Schur update)", "  capacitance: build", "  capacitance: Cholesky + substitutions"]
tot = sum(prof[:12])
for n, v in zip(names, prof[:12]):
    print(f"{n:30s} {v:12d} cyc  {100 * v / tot:5.1f}%")
print("total cycles", tot, "=", tot / 1.965e6, "ms at 1965 MHz")
