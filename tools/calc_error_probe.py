"""Max |descriptor - oracle| of the DeepLCD CNN forward for the library in SLAMB200_LIB (A/B of the tensor-core convolution),
and its device time per batch of 64 keyframes."""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
PKG = "a-simple-stereo-slam-system-with-deep-loop-closing_b200"
pkg = importlib.import_module(PKG)
synth = importlib.import_module(PKG + ".synth")
from oracle import calc_oracle as CO
import torch
w = synth.calc_weights(0)
net = pkg.DeepLCD(w, max_batch=64, max_img_w=1241, max_img_h=376)
imgs = [synth.stereo_pair(s)[0] for s in range(8)]
want = np.stack([CO.calc_descr_original(i, w)[0] for i in imgs])
got = net.calcDescrOriginalImgBatch([i.copy() for i in imgs], in_place=False)
err = np.abs(got - want)
print(f"max |descriptor - oracle| {err.max():.3e}   mean {err.mean():.3e}   max score error {np.abs(got @ got.T - want @ want.T).max():.3e}")
pool = torch.from_numpy(np.stack([synth.stereo_pair(s)[0] for s in range(64)])).cuda()
out = torch.zeros((64, net.dim), dtype=torch.float32, device="cuda")
st = torch.cuda.Stream()
net.set_stream(st.cuda_stream)
a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
H, W = pool.shape[1:]
net.descr_original_dev(64, pool, H * W, W, H, W, out)
a0.record(st)
for _ in range(10):
    net.descr_original_dev(64, pool, H * W, W, H, W, out)
a1.record(st)
torch.cuda.synchronize()
ms = a0.elapsed_time(a1) / 10
print(f"{ms:.3f} ms per 64 keyframes = {64 / ms * 1e3:.0f} keyframes/s")
