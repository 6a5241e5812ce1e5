"""One small invocation of every kernel family (for compute-sanitizer memcheck / racecheck / initcheck)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "orb"):
    frames = synth.stereo_batch(0, 1)
    fe = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=1)
    out = fe.extract_match(frames)
    print("stereo", out["counts"].tolist(), int((out["mdist"][0, :out["counts"][0, 0]] <= 30).sum()))
    # low contrast: cells whose maxima stay below iniTh -> the second FAST tier (minTh) with its TMA re-fetch of the tile
    flat = (frames // 6 + 100).astype(np.uint8)
    out = fe.extract_match(flat)
    print("stereo, low contrast", out["counts"].tolist())
    # two handles with their kernels on one shared stream (sb_stereo_set_compute_stream), copies on their own
    import torch
    shared = torch.cuda.Stream()
    fe2 = pkg.StereoFrontend(2000, 1.2, 8, 20, 7, max_pairs=1)
    o1, o2 = fe.alloc_outputs(1, pinned=True), fe2.alloc_outputs(1, pinned=True)
    pin = torch.from_numpy(np.stack([frames, flat])).pin_memory().numpy()
    for h_ in (fe, fe2):
        h_.set_compute_stream(shared.cuda_stream)
    fe.submit(pin[0], o1); fe2.submit(pin[1], o2); fe.wait(); fe2.wait()
    print("stereo, shared kernel stream", o1["counts"].tolist(), o2["counts"].tolist())
    ext = pkg.ORBextractor(300, 1.2, 8, 20, 7)
    mask = np.full(frames[0, 0].shape, 255, np.uint8); mask[100:200, 300:600] = 0
    k = ext.Detect(frames[0, 0], mask)
    kin = np.repeat(k, 8); kin["octave"] = np.tile(np.arange(8), len(k))
    _, ks = ext.ScreenAndComputeKPsParams(frames[0, 0], kin)
    d = ext.CalcDescriptors(frames[0, 0], ks)
    print("detect/screen/calc", len(k), len(ks), d.shape)
    got = ext.ScreenAndDescribeBatch([frames[0, 0]], [kin])   # sb_orb_screen_describe
    print("screen_describe", len(got[0][1]), got[0][2].shape)
if which in ("all", "ba"):
    w = [synth.ba_window(s, n_points=60) for s in range(2)]
    for k_ in ("obs_pose", "obs_point", "uv"):   # a duplicated observation: the chain path of the kernel
        w[1][k_] = np.concatenate([w[1][k_], w[1][k_][:3]])
    r = pkg.LocalBA(max_windows=2, max_poses=7, max_points=64, max_obs=512).solve(w, synth.KITTI_K)
    print("ba", r[0][4].tolist())
if which in ("all", "lcd"):
    db = synth.lcd_database(0, n=100, pairs=[(90, 5)])
    lcd = pkg.DeepLCDScorer(capacity=128, dtype=1, max_queries=4)
    lcd.add_batch(np.arange(90), db[:90])
    print("lcd", lcd.DetectLoop(90, db[90]), lcd.score(db[:3]).shape)
if which in ("all", "pg"):
    g = synth.pose_graph(1, n=50, n_loops=2)
    p, info = pkg.PoseGraph(64, 128).solve(g["poses0"], g["fixed"], g["v0"], g["v1"], g["meas"], iters=5)
    print("pg", info)
    big = pkg.PoseGraph(64, 256, max_loops=40)   # capacitance matrix in global memory (more loop edges than fit shared memory)
    v0 = np.concatenate([g["v0"], np.arange(20, 48, dtype=np.int32)]); v1 = np.concatenate([g["v1"], np.arange(2, 30, dtype=np.int32)])
    meas = np.concatenate([g["meas"], np.tile(np.array([[0, 0, 0, 1, 0, 0, 0.0]]), (28, 1))])
    print("pg loops", big.solve(g["poses0"], g["fixed"], v0, v1, meas, iters=2)[1])
if which in ("all", "pnp"):
    pr = [synth.pnp_problem(s, n_points=200) for s in range(2)]
    r = pkg.PnPRansac(max_problems=2, max_points=256).solve([(p["obj"], p["img"]) for p in pr], synth.KITTI_K)
    print("pnp", [x["info"].tolist() for x in r])
