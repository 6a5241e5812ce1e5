"""DeepLCD whole-image descriptor on the GPU (SURVEY §8f "next" row 2) against oracle/calc_oracle.py, the
restatement of DeepLCD::calcDescrOriginalImg / calcDescr (reference src/deeplcd.cpp:43-91).

Bars: the pre-processing (7x7 sigma-0 blur handed back in place, 160x120 resize) is u8 fixed point -> BIT-EXACT;
the network is fp32 with a different summation order than the oracle's matrix product (Caffe's own BLAS order is
unspecified as well) -> |descriptor difference| <= 1e-5 absolute on unit-norm descriptors (fp32 sums over K = 1024
products), |score difference| <= 1e-5 (the two loop thresholds are 0.02 apart).  Observed: 2e-6 with every convolution
on the CUDA cores (SLAMB200_CALC_NO_TC=1), 5e-6 with conv2 / conv3 on the tensor cores (tf32 x 3 split, partial
accumulators: the tensor core's fp32 accumulation truncates) — tools/calc_error_probe.py prints both."""
import numpy as np
import pytest

from oracle import calc_oracle as CO

pytestmark = pytest.mark.gpu


def test_descriptor_and_in_place_blur_kitti_size(pkg, synth):
    w = synth.calc_weights(0)
    net = pkg.DeepLCD(w, max_batch=4)
    assert net.dim == 1064
    imgs = [synth.stereo_pair(s)[0] for s in range(4)]
    want = [CO.calc_descr_original(i, w) for i in imgs]
    work = [i.copy() for i in imgs]
    got = net.calcDescrOriginalImgBatch(work, in_place=True)
    for b in range(4):
        assert np.array_equal(work[b], want[b][1]), "blurred image differs"   # bit-exact, quirk: in place
        assert np.abs(got[b] - want[b][0]).max() <= 1e-5
        assert abs(float(np.linalg.norm(got[b])) - 1.0) < 1e-6
    # scores between descriptors agree
    G = got @ got.T
    W = np.stack([x[0] for x in want]) @ np.stack([x[0] for x in want]).T
    assert np.abs(G - W).max() <= 1e-5
    # not in place: the caller's image is untouched and the result is the same
    keep = imgs[0].copy()
    d = net.calcDescrOriginalImg(keep, in_place=False)
    assert np.array_equal(keep, imgs[0]) and np.array_equal(d, got[0])


@pytest.mark.parametrize("shape", [(120, 160), (97, 203), (480, 752), (131, 160)])
def test_other_image_sizes(pkg, synth, shape):
    rng = np.random.default_rng(shape[0])
    w = synth.calc_weights(1)
    net = pkg.DeepLCD(w, max_batch=2, max_img_w=800, max_img_h=480)
    imgs = [rng.integers(0, 256, shape, dtype=np.uint8) for _ in range(2)]
    work = [i.copy() for i in imgs]
    got = net.calcDescrOriginalImgBatch(work)
    for b in range(2):
        want, blurred = CO.calc_descr_original(imgs[b], w)
        assert np.array_equal(work[b], blurred)
        assert np.abs(got[b] - want).max() <= 1e-5


def test_calc_descr_on_resized_input(pkg, synth):
    w = synth.calc_weights(2)
    net = pkg.DeepLCD(w, max_batch=3)
    rng = np.random.default_rng(3)
    imgs = [rng.integers(0, 256, (120, 160), dtype=np.uint8) for _ in range(3)]
    got = net.calcDescrBatch(imgs)
    for b in range(3):
        assert np.abs(got[b] - CO.calc_descr(imgs[b], w)).max() <= 1e-5


def test_other_layer_lists(pkg):
    """The network is data: padded pooling, a convolution without ReLU, channel counts that are not multiples of the tile."""
    layers = [dict(type=0, num_output=10, kernel=3, stride=1, pad=1), dict(type=1),
              dict(type=2, kernel=3, stride=2, pad=1), dict(type=3, local_size=3, alpha=0.01, beta=0.5, k=2.0),
              dict(type=0, num_output=70, kernel=2, stride=2, pad=0),
              dict(type=2, kernel=2, stride=2, pad=0)]
    rng = np.random.default_rng(0)
    w = (rng.standard_normal(CO.n_weights(layers)) * 0.3).astype(np.float32)
    net = pkg.DeepLCD(w, layers=layers, in_h=33, in_w=47, max_batch=2, max_img_w=64, max_img_h=64)
    assert net.dim == int(np.prod(CO.shapes(layers, (1, 33, 47))[-1]))
    imgs = [rng.integers(0, 256, (33, 47), dtype=np.uint8) for _ in range(2)]
    got = net.calcDescrBatch(imgs)
    for b in range(2):
        want = CO.calc_descr(imgs[b], w, layers)
        assert np.abs(got[b] - want).max() <= 1e-5


def test_loop_detection_through_the_cnn(pkg, synth):
    """calcDescrOriginalImg -> database -> DetectLoop: a revisited place (the same scene + sensor noise) is found."""
    w = synth.calc_weights(0)
    net = pkg.DeepLCD(w, max_batch=8)
    lcd = pkg.DeepLCDScorer(capacity=64, dtype=0)
    scenes = [synth.stereo_pair(100 + s)[0] for s in range(8)]
    d = net.calcDescrOriginalImgBatch([s.copy() for s in scenes], in_place=False)
    for k in range(8):
        lcd.add(k, d[k])
    revisit = np.clip(scenes[2].astype(np.int32) + np.random.default_rng(1).integers(-2, 3, scenes[2].shape), 0, 255).astype(np.uint8)
    q = net.calcDescrOriginalImg(revisit, in_place=False)
    s = lcd.score(q[None])[0]
    assert int(np.argmax(s)) == 2 and s[2] > 0.94


def test_bad_arguments(pkg, synth):
    w = synth.calc_weights(0)
    with pytest.raises(pkg.SlamB200Error):
        pkg.DeepLCD(w[:-1])
    with pytest.raises(pkg.SlamB200Error):
        pkg.DeepLCD(np.concatenate([w, w[:1]]))
    with pytest.raises(pkg.SlamB200Error):
        pkg.DeepLCD(w, layers=[dict(type=1)])
    net = pkg.DeepLCD(w, max_batch=1)
    with pytest.raises(pkg.SlamB200Error):
        net.calcDescrOriginalImgBatch([np.zeros((376, 1241), np.uint8)] * 2)


def test_from_caffe_files(pkg, synth, tmp_path):
    """DeepLCD(deploy.prototxt, calc.caffemodel, gpu_id): the files are read by csrc/caffe_io.cu (no Caffe, no protobuf)."""
    import os
    from test_caffe_io import PROTOTXT, write_caffemodel
    w = synth.calc_weights(4)
    (W1, b1), (W2, b2), (W3, b3) = CO.split_weights(w)
    path = str(tmp_path / "calc.caffemodel")
    write_caffemodel(path, [("conv1", [W1, b1]), ("conv2", [W2, b2]), ("conv3", [W3, b3])])
    net = pkg.DeepLCD.from_caffe(PROTOTXT, path, gpu_id=0)
    img = synth.stereo_pair(9)[0]
    want, blurred = CO.calc_descr_original(img, w)
    got = net.calcDescrOriginalImg(img)
    assert np.array_equal(img, blurred) and np.abs(got - want).max() <= 1e-5
