"""CPU restatement of DeepLCD::calcDescrOriginalImg / calcDescr (reference src/deeplcd.cpp:43-91).

TEST INFRASTRUCTURE ONLY — only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this.

What is restated (paths relative to /root/reference):
  * calcDescrOriginalImg (src/deeplcd.cpp:43-52): cv::GaussianBlur(img, img, Size(7, 7), 0) IN PLACE on the
    caller's image (the `const cv::Mat&` is only a const header: the keyframe's mImageLeft comes back blurred
    and src/loopclosing.cpp:106-112 then runs ScreenAndComputeKPsParams / CalcDescriptors on that blurred
    image — quirk Q13), cv::resize to 160 x 120 (INTER_LINEAR), then calcDescr;
  * calcDescr (:55-91): u8 -> float * (1 / 255), Net::Forward, copy the 1064 outputs, descriptor /= norm.
  * OpenCV (third party, absent; author used 3.4.8, pinned here against cv2 4.13.0 by
    tests/test_oracle_calc.py): sigma = 0 with ksize 7 selects OpenCV's built-in table
    [0.03125, 0.109375, 0.21875, 0.28125, ...] = [8, 28, 56, 72, 56, 28, 8] / 256, run by the u8 fixed-point
    path (8.8 rows, 16.16 columns, (acc + 32768) >> 16, BORDER_REFLECT_101); the resize is the fixed-point
    bilinear of SURVEY A.1 (oracle/orb_oracle.c: orc_resize_linear_u8, already pinned).
  * Caffe (third party, absent) layer semantics, restated from its published layer definitions:
      Convolution: cross-correlation, out = floor((in + 2 pad - k) / stride) + 1, bias added;
      Pooling MAX: out = ceil((in + 2 pad - k) / stride) + 1, minus one if the last window would start
                   beyond in + pad; windows are clipped to the image (no padding value takes part);
      LRN ACROSS_CHANNELS: y_c = x_c * (k + alpha / n * sum_{|c' - c| <= n / 2} x_c'^2) ^ (-beta);
      ReLU; Flatten in (C, H, W) order.
  * The network itself: calc_model/deploy.prototxt and calc.caffemodel are a configure-time download
    (get_model.sh:3-16) and absent.  CALC_LAYERS below is the CALC architecture (rpng/calc) as published:
    1 x 120 x 160 -> conv 64 x 5 x 5 / 2 pad 4 -> ReLU -> max 3 / 2 -> LRN 5 -> conv 128 x 4 x 4 pad 2 -> ReLU
    -> max 3 / 2 -> LRN 5 -> conv 4 x 3 x 3 -> ReLU -> flatten; it reproduces the one number the reference
    pins (`assert(p == 1064)`, src/deeplcd.cpp:82: 4 x 14 x 19 = 1064).  The layer list is data, not code:
    whatever the real prototxt says can be passed instead.

PARITY UNPINNED for the network: neither Caffe nor the trained weights exist here.  The layer arithmetic is
pinned against torch's CPU fp32 conv2d / max_pool2d(ceil_mode) / local_response_norm
(tests/test_oracle_calc.py); weights in tests are seeded random (synth.calc_weights).
"""
import numpy as np

CONV, RELU, POOL_MAX, LRN = 0, 1, 2, 3

# (type, num_output, kernel, stride, pad, local_size, alpha, beta, k)
CALC_LAYERS = [
    dict(type=CONV, num_output=64, kernel=5, stride=2, pad=4),
    dict(type=RELU),
    dict(type=POOL_MAX, kernel=3, stride=2, pad=0),
    dict(type=LRN, local_size=5, alpha=1e-4, beta=0.75, k=1.0),
    dict(type=CONV, num_output=128, kernel=4, stride=1, pad=2),
    dict(type=RELU),
    dict(type=POOL_MAX, kernel=3, stride=2, pad=0),
    dict(type=LRN, local_size=5, alpha=1e-4, beta=0.75, k=1.0),
    dict(type=CONV, num_output=4, kernel=3, stride=1, pad=0),
    dict(type=RELU),
]
IN_H, IN_W = 120, 160

GAUSS7_SIGMA0 = np.array([8, 28, 56, 72, 56, 28, 8], np.int64)  # cv::getGaussianKernel(7, 0) * 256, exact


def gauss7_sigma0(img):
    """cv::GaussianBlur(img, img, Size(7, 7), 0) on a u8 image (src/deeplcd.cpp:46)."""
    img = np.asarray(img, np.uint8)
    h, w = img.shape
    p = np.pad(img.astype(np.int64), ((3, 3), (3, 3)), mode="reflect")
    rows = sum(GAUSS7_SIGMA0[i] * p[:, i:i + w] for i in range(7))
    cols = sum(GAUSS7_SIGMA0[i] * rows[i:i + h, :] for i in range(7))
    return ((cols + 32768) >> 16).astype(np.uint8)


def preprocess(img):
    """calcDescrOriginalImg up to the net input: returns (blurred original, 160 x 120 u8)."""
    from . import oracle as O
    blurred = gauss7_sigma0(img)
    return blurred, O.resize_linear(blurred, IN_W, IN_H)


def shapes(layers=CALC_LAYERS, in_shape=(1, IN_H, IN_W)):
    """Blob shape after every layer, Caffe's rules."""
    c, h, w = in_shape
    out = []
    for L in layers:
        if L["type"] == CONV:
            k, s, p = L["kernel"], L["stride"], L["pad"]
            c, h, w = L["num_output"], (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
        elif L["type"] == POOL_MAX:
            k, s, p = L["kernel"], L["stride"], L.get("pad", 0)
            oh, ow = -(-(h + 2 * p - k) // s) + 1, -(-(w + 2 * p - k) // s) + 1
            if p > 0:
                if (oh - 1) * s >= h + p:
                    oh -= 1
                if (ow - 1) * s >= w + p:
                    ow -= 1
            h, w = oh, ow
        out.append((c, h, w))
    return out


def n_weights(layers=CALC_LAYERS, in_channels=1):
    n, c = 0, in_channels
    for L in layers:
        if L["type"] == CONV:
            n += L["num_output"] * c * L["kernel"] ** 2 + L["num_output"]
            c = L["num_output"]
    return n


def split_weights(weights, layers=CALC_LAYERS, in_channels=1):
    """Flat fp32 buffer -> [(W [Cout][Cin][k][k], bias [Cout])] in layer order (Caffe's blob order)."""
    weights = np.asarray(weights, np.float32)
    out, o, c = [], 0, in_channels
    for L in layers:
        if L["type"] == CONV:
            co, k = L["num_output"], L["kernel"]
            W = weights[o:o + co * c * k * k].reshape(co, c, k, k)
            o += W.size
            b = weights[o:o + co]
            o += co
            out.append((W, b))
            c = co
    assert o == weights.size
    return out


def conv(x, W, b, stride, pad):
    """Caffe Convolution on one [C][H][W] fp32 blob: im2col + one fp32 matrix product, like Caffe itself."""
    c, h, w = x.shape
    co, _, k, _ = W.shape
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    xp = np.zeros((c, h + 2 * pad, w + 2 * pad), np.float32)
    xp[:, pad:pad + h, pad:pad + w] = x
    col = np.empty((c, k, k, oh, ow), np.float32)
    for ky in range(k):
        for kx in range(k):
            col[:, ky, kx] = xp[:, ky:ky + (oh - 1) * stride + 1:stride, kx:kx + (ow - 1) * stride + 1:stride]
    y = W.reshape(co, -1).astype(np.float32) @ col.reshape(c * k * k, oh * ow)
    return (y + b.astype(np.float32)[:, None]).reshape(co, oh, ow).astype(np.float32)


def pool_max(x, k, stride, pad=0):
    c, h, w = x.shape
    (_, oh, ow), = shapes([dict(type=POOL_MAX, kernel=k, stride=stride, pad=pad)], (c, h, w))
    y = np.empty((c, oh, ow), np.float32)
    for i in range(oh):
        y0, y1 = max(i * stride - pad, 0), min(i * stride - pad + k, h)
        for j in range(ow):
            x0, x1 = max(j * stride - pad, 0), min(j * stride - pad + k, w)
            y[:, i, j] = x[:, y0:y1, x0:x1].max(axis=(1, 2))
    return y


def lrn(x, n, alpha, beta, k):
    c = x.shape[0]
    sq = np.zeros((c + n - 1,) + x.shape[1:], np.float32)
    sq[n // 2:n // 2 + c] = x * x
    acc = np.zeros_like(x)
    for d in range(n):
        acc = acc + sq[d:d + c]
    scale = np.float32(k) + np.float32(alpha / n) * acc
    return (x * np.power(scale, np.float32(-beta))).astype(np.float32)


def forward(x, weights, layers=CALC_LAYERS):
    """Net::Forward on one fp32 [C][H][W] input; returns the flattened output blob."""
    x = np.asarray(x, np.float32)
    wb = iter(split_weights(weights, layers, x.shape[0]))
    for L in layers:
        if L["type"] == CONV:
            W, b = next(wb)
            x = conv(x, W, b, L["stride"], L["pad"])
        elif L["type"] == RELU:
            x = np.maximum(x, np.float32(0))
        elif L["type"] == POOL_MAX:
            x = pool_max(x, L["kernel"], L["stride"], L.get("pad", 0))
        elif L["type"] == LRN:
            x = lrn(x, L["local_size"], L["alpha"], L["beta"], L["k"])
        else:
            raise ValueError(L)
    return x.reshape(-1)


def calc_descr(img_u8, weights, layers=CALC_LAYERS):
    """DeepLCD::calcDescr (src/deeplcd.cpp:55-91) on an already resized u8 image."""
    x = np.asarray(img_u8, np.uint8).astype(np.float32) * np.float32(1.0 / 255.0)
    d = forward(x[None], weights, layers)
    return (d / np.sqrt(np.sum(d * d, dtype=np.float32))).astype(np.float32)


def calc_descr_original(img_u8, weights, layers=CALC_LAYERS):
    """DeepLCD::calcDescrOriginalImg (:43-52): returns (descriptor, the blurred image the caller is left with)."""
    blurred, small = preprocess(img_u8)
    return calc_descr(small, weights, layers), blurred
