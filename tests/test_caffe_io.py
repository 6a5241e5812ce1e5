"""Host logic of the DeepLCD model loader (csrc/caffe_io.cu; DeepLCD::DeepLCD, reference src/deeplcd.cpp:10-31):
deploy.prototxt (protobuf text) + .caffemodel (protobuf wire format, written here by a minimal encoder that
follows the published caffe.proto field numbers) -> (layer list, flat weights).  Needs no GPU."""
import os
import struct

import numpy as np
import pytest

from oracle import calc_oracle as CO

HERE = os.path.dirname(os.path.abspath(__file__))
PROTOTXT = os.path.join(HERE, "golden", "calc_deploy.prototxt")


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(field, payload):
    return _varint(field << 3 | 2) + _varint(len(payload)) + payload


def _blob(arr, packed=True, legacy_shape=False):
    arr = np.asarray(arr, np.float32)
    if packed:
        data = _ld(5, arr.tobytes())
    else:
        data = b"".join(_varint(5 << 3 | 5) + struct.pack("<f", float(x)) for x in arr.ravel())
    if legacy_shape:
        dims = (list(arr.shape) + [1, 1, 1, 1])[:4] if arr.ndim == 4 else [1, 1, 1, arr.size]
        shape = b"".join(_varint((i + 1) << 3) + _varint(d) for i, d in enumerate(dims))
    else:
        shape = _ld(7, _ld(1, b"".join(_varint(d) for d in arr.shape)))
    return shape + data


def _layer(name, blobs, v1=False, **kw):
    body = _ld(4 if v1 else 1, name.encode()) + (b"" if v1 else _ld(2, b"Convolution"))
    body += b"".join(_ld(6 if v1 else 7, _blob(b, **kw)) for b in blobs)
    return _ld(2 if v1 else 100, body)


def write_caffemodel(path, named, v1=False, **kw):
    with open(path, "wb") as f:
        f.write(_ld(1, b"calc"))
        for name, blobs in named:
            f.write(_layer(name, blobs, v1, **kw))


@pytest.mark.parametrize("variant", ["packed", "unpacked", "v1_legacy"])
def test_round_trip(pkg, synth, tmp_path, variant):
    w = synth.calc_weights(7)
    (W1, b1), (W2, b2), (W3, b3) = CO.split_weights(w)
    rng = np.random.default_rng(0)
    named = [("conv1", [W1, b1]), ("deconv_not_in_deploy", [rng.standard_normal((3, 5)).astype(np.float32)]),
             ("conv3", [W3, b3]), ("conv2", [W2, b2])]           # file order differs from net order: matched by name
    path = str(tmp_path / "calc.caffemodel")
    write_caffemodel(path, named, v1=variant == "v1_legacy", packed=variant != "unpacked", legacy_shape=variant == "v1_legacy")
    layers, weights, shape = pkg.parse_caffe(PROTOTXT, path)
    assert shape == (120, 160)
    assert np.array_equal(weights, w)
    assert len(layers) == len(CO.CALC_LAYERS)
    for got, want in zip(layers, CO.CALC_LAYERS):
        assert got["type"] == want["type"]
        for key in ("num_output", "kernel", "stride", "pad", "local_size"):
            assert got[key] == want.get(key, 0), (key, got, want)
        if want["type"] == CO.LRN:
            assert got["alpha"] == np.float32(want["alpha"]) and got["beta"] == np.float32(want["beta"]) and got["k"] == 1.0


def test_errors(pkg, synth, tmp_path):
    w = synth.calc_weights(7)
    (W1, b1), (W2, b2), (W3, b3) = CO.split_weights(w)
    path = str(tmp_path / "m.caffemodel")
    write_caffemodel(path, [("conv1", [W1, b1]), ("conv2", [W2, b2])])
    with pytest.raises(pkg.SlamB200Error, match="conv3"):
        pkg.parse_caffe(PROTOTXT, path)
    write_caffemodel(path, [("conv1", [W1, b1]), ("conv2", [W2[:, :, :3], b2]), ("conv3", [W3, b3])])
    with pytest.raises(pkg.SlamB200Error, match="conv2"):
        pkg.parse_caffe(PROTOTXT, path)
    with pytest.raises(pkg.SlamB200Error, match="cannot read"):
        pkg.parse_caffe(PROTOTXT, str(tmp_path / "missing"))
    open(path, "wb").write(b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff")
    with pytest.raises(pkg.SlamB200Error, match="not a caffemodel"):
        pkg.parse_caffe(PROTOTXT, path)
    bad = tmp_path / "bad.prototxt"
    bad.write_text('input_shape { dim: 1 dim: 1 dim: 8 dim: 8 }\nlayer { name: "x" type: "InnerProduct" }\n')
    write_caffemodel(path, [])
    with pytest.raises(pkg.SlamB200Error, match="InnerProduct"):
        pkg.parse_caffe(str(bad), path)
    bad.write_text('input_shape { dim: 1 dim: 1 dim: 8 dim: 8 \nlayer { name: "x" type: "ReLU" }\n')
    with pytest.raises(pkg.SlamB200Error, match="missing '}'"):
        pkg.parse_caffe(str(bad), path)


def test_other_prototxt_spellings(pkg, tmp_path):
    """input_dim, an Input layer, kernel_h / kernel_w, comments, 'key: { }' and caffe.proto's LRN defaults."""
    txt = tmp_path / "n.prototxt"
    txt.write_text('''
# comment
name: "n"
layer { name: "data" type: "Input" top: "data" input_param { shape: { dim: 1 dim: 1 dim: 20 dim: 30 } } }
layer { name: "c" type: "Convolution" convolution_param { num_output: 2 kernel_h: 3 kernel_w: 3 } }
layer { name: "n" type: "LRN" }
layer { name: "p" type: "Pooling" pooling_param { kernel_size: 2 stride: 2 pad: 1 } }
''')
    rng = np.random.default_rng(1)
    W, b = rng.standard_normal((2, 1, 3, 3)).astype(np.float32), rng.standard_normal(2).astype(np.float32)
    path = str(tmp_path / "n.caffemodel")
    write_caffemodel(path, [("c", [W, b])])
    layers, weights, shape = pkg.parse_caffe(str(txt), path)
    assert shape == (20, 30) and [L["type"] for L in layers] == [0, 3, 2]
    assert layers[0]["stride"] == 1 and layers[0]["pad"] == 0 and layers[0]["kernel"] == 3
    assert layers[1]["local_size"] == 5 and layers[1]["alpha"] == 1.0 and layers[1]["beta"] == 0.75
    assert layers[2]["pad"] == 1
    assert np.array_equal(weights, np.concatenate([W.ravel(), b]))
