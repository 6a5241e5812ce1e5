"""Per-operator wall time of the keyframe stage of the replay (config 4), with a device synchronise around every call."""
import importlib, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
replay = importlib.import_module(PKG + ".replay")
seq = replay.Sequence(frames=64, n_kf=64)
ops = replay.GpuOps(device=0, batch=64, kf_batch=32, n_kf=64)
imgs = torch.from_numpy(seq.load(seq.kf_frame[np.arange(32)])).pin_memory().numpy()
lefts0 = [imgs[i, 0] for i in range(32)]
rights = [imgs[i, 1] for i in range(32)]
def timed(name, f, n=6):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(f"{name:22s} best {1e3 * min(ts):8.3f} ms   median {1e3 * sorted(ts)[len(ts) // 2]:8.3f} ms")
    return r
feats = timed("kf_detect", lambda: ops.kf_detect(lefts0))
pts = [np.stack([f["x"], f["y"]], 1).astype(np.float32) for f in feats]
tracked = timed("lk_right", lambda: ops.lk_right(lefts0, rights, pts))
pinned = torch.from_numpy(np.stack(lefts0)).pin_memory().numpy()      # the replay keeps its keyframe images page-locked
lefts = [pinned[i] for i in range(32)]
timed("cnn_descr", lambda: ops.cnn_descr(lefts))
timed("expand_octaves_batch (host)", lambda: replay.expand_octaves_batch(feats))
timed("screen_and_describe", lambda: ops.screen_and_describe(lefts, feats))
timed("triangulate_batch", lambda: ops.triangulate_batch(pts, [tr[0] for tr in tracked]))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    ops.kf_detect(lefts0); ops.screen_and_describe(lefts, feats)
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
