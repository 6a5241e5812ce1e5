// caffe_io.cu — reads the two files DeepLCD::DeepLCD loads (reference src/deeplcd.cpp:10-31:
// `new caffe::Net<float>(deploy.prototxt, TEST)` + `CopyTrainedLayersFrom(calc.caffemodel)`) into the
// (layer list, flat weight buffer) form sb_calc_create takes.  Host code only, no Caffe and no protobuf
// library: deploy.prototxt is protobuf TEXT format (parsed by a small recursive-descent reader),
// calc.caffemodel is protobuf WIRE format of caffe.proto's NetParameter, of which only these fields are
// needed (numbers from the published caffe.proto):
//   NetParameter   : layer = 100 (LayerParameter), layers = 2 (V1LayerParameter, pre-2015 files)
//   LayerParameter : name = 1, blobs = 7          V1LayerParameter: name = 4, blobs = 6
//   BlobProto      : data = 5 (packed or unpacked float), shape = 7 { dim = 1 }, num/channels/height/width = 1..4
// Weights are matched to layers BY NAME, like CopyTrainedLayersFrom: layers of the model file that the deploy
// net does not have (the auto-encoder's decoder) are skipped.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------------
// protobuf text format
// ---------------------------------------------------------------------------------------------------
struct Node {
    std::vector<std::pair<std::string, std::string>> scalars;  // key: value (quotes removed)
    std::vector<std::pair<std::string, Node>> children;        // key { ... }
    const std::string *get(const char *key) const {
        for (auto &s : scalars)
            if (s.first == key) return &s.second;
        return nullptr;
    }
    const Node *child(const char *key) const {
        for (auto &c : children)
            if (c.first == key) return &c.second;
        return nullptr;
    }
    std::vector<std::string> all(const char *key) const {
        std::vector<std::string> out;
        for (auto &s : scalars)
            if (s.first == key) out.push_back(s.second);
        return out;
    }
};

struct TextParser {
    const char *p, *end;
    std::string err;
    void skip() {
        for (;;) {
            while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == ',' || *p == ';')) p++;
            if (p < end && *p == '#') {
                while (p < end && *p != '\n') p++;
                continue;
            }
            return;
        }
    }
    bool token(std::string &t) {
        skip();
        t.clear();
        if (p >= end) return false;
        if (*p == '"' || *p == '\'') {
            const char q = *p++;
            while (p < end && *p != q) {
                if (*p == '\\' && p + 1 < end) p++;
                t.push_back(*p++);
            }
            if (p >= end) { err = "unterminated string"; return false; }
            p++;
            return true;
        }
        while (p < end && !strchr(" \t\r\n:{}#,;\"'", *p)) t.push_back(*p++);
        if (t.empty()) err = std::string("unexpected character '") + *p + "'";
        return !t.empty();
    }
    bool parse(Node &n, bool top) {
        for (;;) {
            skip();
            if (p >= end) {
                if (!top) err = "missing '}'";
                return top;
            }
            if (*p == '}') {
                if (top) { err = "unbalanced '}'"; return false; }
                p++;
                return true;
            }
            std::string key;
            if (!token(key)) return false;
            skip();
            if (p < end && *p == ':') {
                p++;
                skip();
                if (p < end && *p == '{') {  // "key: { ... }" is legal text format
                    p++;
                    n.children.emplace_back(key, Node());
                    if (!parse(n.children.back().second, false)) return false;
                    continue;
                }
                std::string val;
                if (!token(val)) { if (err.empty()) err = "missing value after '" + key + ":'"; return false; }
                n.scalars.emplace_back(key, val);
            } else if (p < end && *p == '{') {
                p++;
                n.children.emplace_back(key, Node());
                if (!parse(n.children.back().second, false)) return false;
            } else {
                err = "expected ':' or '{' after '" + key + "'";
                return false;
            }
        }
    }
};

int geti(const Node *n, const char *key, int dflt) {
    const std::string *s = n ? n->get(key) : nullptr;
    return s ? atoi(s->c_str()) : dflt;
}
float getf(const Node *n, const char *key, float dflt) {
    const std::string *s = n ? n->get(key) : nullptr;
    return s ? strtof(s->c_str(), nullptr) : dflt;
}

// kernel_size / kernel_h + kernel_w (must be square here), likewise stride and pad
bool square_param(const Node *n, const char *base, const char *hname, const char *wname, int dflt, int *out) {
    const int v = geti(n, base, -1), vh = geti(n, hname, -1), vw = geti(n, wname, -1);
    if (vh >= 0 || vw >= 0) {
        if (vh != vw) return false;
        *out = vh;
    } else {
        *out = v >= 0 ? v : dflt;
    }
    return true;
}

struct ParsedNet {
    int in_c = 0, in_h = 0, in_w = 0;
    std::vector<sb_calc_layer> layers;
    std::vector<std::string> conv_names;  // one per SB_CALC_CONV layer, in order
    std::vector<int> conv_in_channels;
};

int parse_prototxt(const std::string &text, ParsedNet &net) {
    Node root;
    TextParser tp{text.data(), text.data() + text.size(), ""};
    if (!tp.parse(root, true)) {
        sb_set_error("deploy.prototxt: %s", tp.err.c_str());
        return SB_ERR_INVALID;
    }
    std::vector<int> dims;
    if (const Node *s = root.child("input_shape"))
        for (auto &d : s->all("dim")) dims.push_back(atoi(d.c_str()));
    if (dims.empty())
        for (auto &d : root.all("input_dim")) dims.push_back(atoi(d.c_str()));
    int c = 0, hh = 0, ww = 0;
    for (auto &ch : root.children) {
        if (ch.first != "layer" && ch.first != "layers") continue;
        const Node &L = ch.second;
        const std::string *type = L.get("type"), *name = L.get("name");
        if (!type) { sb_set_error("deploy.prototxt: a layer has no type"); return SB_ERR_INVALID; }
        std::string t = *type;
        for (auto &chr : t) chr = (char)toupper((unsigned char)chr);
        if (t == "INPUT" || t == "DATA") {
            const Node *ip = L.child("input_param");
            const Node *s = ip ? ip->child("shape") : nullptr;
            if (s)
                for (auto &d : s->all("dim")) dims.push_back(atoi(d.c_str()));
            continue;
        }
        if (c == 0) {
            if (dims.size() != 4 || dims[1] < 1 || dims[2] < 1 || dims[3] < 1) { sb_set_error("deploy.prototxt: no 4-d input shape before the first layer"); return SB_ERR_INVALID; }
            net.in_c = c = dims[1]; net.in_h = hh = dims[2]; net.in_w = ww = dims[3];
        }
        sb_calc_layer S;
        memset(&S, 0, sizeof(S));
        if (t == "CONVOLUTION") {
            const Node *cp = L.child("convolution_param");
            S.type = SB_CALC_CONV;
            S.num_output = geti(cp, "num_output", 0);
            int k, s, p;
            if (!square_param(cp, "kernel_size", "kernel_h", "kernel_w", 0, &k) || !square_param(cp, "stride", "stride_h", "stride_w", 1, &s) ||
                !square_param(cp, "pad", "pad_h", "pad_w", 0, &p) || k < 1 || S.num_output < 1) {
                sb_set_error("deploy.prototxt: layer '%s': only square kernels / strides / pads are supported", name ? name->c_str() : "?");
                return SB_ERR_INVALID;
            }
            if (geti(cp, "group", 1) != 1 || geti(cp, "dilation", 1) != 1 || (cp && cp->get("bias_term") && *cp->get("bias_term") == "false")) {
                sb_set_error("deploy.prototxt: layer '%s': group / dilation / bias_term: false are not supported", name ? name->c_str() : "?");
                return SB_ERR_INVALID;
            }
            S.kernel = k; S.stride = s; S.pad = p;
            net.conv_names.push_back(name ? *name : std::string());
            net.conv_in_channels.push_back(c);
            c = S.num_output;
        } else if (t == "RELU") {
            S.type = SB_CALC_RELU;
            if (getf(L.child("relu_param"), "negative_slope", 0.f) != 0.f) { sb_set_error("deploy.prototxt: leaky ReLU is not supported"); return SB_ERR_INVALID; }
        } else if (t == "POOLING") {
            const Node *pp = L.child("pooling_param");
            const std::string *pool = pp ? pp->get("pool") : nullptr;
            if (pool && *pool != "MAX" && *pool != "0") { sb_set_error("deploy.prototxt: only MAX pooling is supported"); return SB_ERR_INVALID; }
            S.type = SB_CALC_POOL_MAX;
            int k, s, p;
            if (!square_param(pp, "kernel_size", "kernel_h", "kernel_w", 0, &k) || !square_param(pp, "stride", "stride_h", "stride_w", 1, &s) ||
                !square_param(pp, "pad", "pad_h", "pad_w", 0, &p) || k < 1) {
                sb_set_error("deploy.prototxt: layer '%s': only square pooling windows are supported", name ? name->c_str() : "?");
                return SB_ERR_INVALID;
            }
            S.kernel = k; S.stride = s; S.pad = p;
        } else if (t == "LRN") {
            const Node *lp = L.child("lrn_param");
            const std::string *region = lp ? lp->get("norm_region") : nullptr;
            if (region && *region != "ACROSS_CHANNELS" && *region != "0") { sb_set_error("deploy.prototxt: only ACROSS_CHANNELS LRN is supported"); return SB_ERR_INVALID; }
            S.type = SB_CALC_LRN;
            S.local_size = geti(lp, "local_size", 5);  // caffe.proto defaults
            S.alpha = getf(lp, "alpha", 1.f);
            S.beta = getf(lp, "beta", 0.75f);
            S.k = getf(lp, "k", 1.f);
        } else if (t == "FLATTEN" || t == "DROPOUT") {
            continue;  // Flatten is implicit; Dropout is the identity at test time
        } else {
            sb_set_error("deploy.prototxt: layer type '%s' is not supported", type->c_str());
            return SB_ERR_INVALID;
        }
        net.layers.push_back(S);
    }
    (void)hh; (void)ww;
    if (net.layers.empty()) { sb_set_error("deploy.prototxt: no layers"); return SB_ERR_INVALID; }
    if (net.in_c != 1) { sb_set_error("deploy.prototxt: the net must take a 1-channel image (src/deeplcd.cpp:60-68)"); return SB_ERR_INVALID; }
    return SB_OK;
}

// ---------------------------------------------------------------------------------------------------
// protobuf wire format
// ---------------------------------------------------------------------------------------------------
struct Wire {
    const uint8_t *p, *end;
    bool ok = true;
    uint64_t varint() {
        uint64_t v = 0;
        for (int shift = 0; shift < 64; shift += 7) {
            if (p >= end) { ok = false; return 0; }
            const uint8_t b = *p++;
            v |= (uint64_t)(b & 0x7f) << shift;
            if (!(b & 0x80)) return v;
        }
        ok = false;
        return 0;
    }
    // next field: number, wire type, and for length-delimited fields the payload
    bool next(int &field, int &wt, Wire &sub, uint64_t &val) {
        if (p >= end || !ok) return false;
        const uint64_t key = varint();
        if (!ok) return false;
        field = (int)(key >> 3);
        wt = (int)(key & 7);
        val = 0;
        if (wt == 0) val = varint();
        else if (wt == 1) { if (end - p < 8) { ok = false; return false; } memcpy(&val, p, 8); p += 8; }
        else if (wt == 5) { if (end - p < 4) { ok = false; return false; } uint32_t v32; memcpy(&v32, p, 4); val = v32; p += 4; }
        else if (wt == 2) {
            const uint64_t len = varint();
            if (!ok || len > (uint64_t)(end - p)) { ok = false; return false; }
            sub.p = p; sub.end = p + len; sub.ok = true;
            p += len;
        } else { ok = false; return false; }
        return ok;
    }
};

struct Blob {
    std::vector<float> data;
    std::vector<int64_t> shape;
};

bool read_blob(Wire w, Blob &b) {
    int f, wt;
    Wire sub{nullptr, nullptr};
    uint64_t v;
    int64_t legacy[4] = {-1, -1, -1, -1};
    while (w.next(f, wt, sub, v)) {
        if (f == 5 && wt == 2) {  // packed floats
            const size_t n = (size_t)(sub.end - sub.p) / 4;
            const size_t o = b.data.size();
            b.data.resize(o + n);
            memcpy(b.data.data() + o, sub.p, n * 4);
        } else if (f == 5 && wt == 5) {
            float x;
            const uint32_t u = (uint32_t)v;
            memcpy(&x, &u, 4);
            b.data.push_back(x);
        } else if (f == 7 && wt == 2) {  // BlobShape
            Wire s = sub, ss{nullptr, nullptr};
            int f2, wt2;
            uint64_t v2;
            while (s.next(f2, wt2, ss, v2)) {
                if (f2 == 1 && wt2 == 0) b.shape.push_back((int64_t)v2);
                else if (f2 == 1 && wt2 == 2) {
                    Wire pk = ss;
                    while (pk.p < pk.end && pk.ok) b.shape.push_back((int64_t)pk.varint());
                }
            }
            if (!s.ok) return false;
        } else if (f >= 1 && f <= 4 && wt == 0) {
            legacy[f - 1] = (int64_t)v;
        }
    }
    if (b.shape.empty() && legacy[0] >= 0)
        for (int i = 0; i < 4; i++) b.shape.push_back(legacy[i] < 0 ? 1 : legacy[i]);
    return w.ok;
}

// name -> blobs of every layer of the model file
bool read_model(const std::vector<uint8_t> &buf, std::map<std::string, std::vector<Blob>> &out) {
    Wire w{buf.data(), buf.data() + buf.size()};
    int f, wt;
    Wire sub{nullptr, nullptr};
    uint64_t v;
    while (w.next(f, wt, sub, v)) {
        if (wt != 2 || (f != 100 && f != 2)) continue;
        const int name_field = f == 100 ? 1 : 4, blob_field = f == 100 ? 7 : 6;
        Wire L = sub, s2{nullptr, nullptr};
        int f2, wt2;
        uint64_t v2;
        std::string name;
        std::vector<Blob> blobs;
        while (L.next(f2, wt2, s2, v2)) {
            if (f2 == name_field && wt2 == 2) name.assign((const char *)s2.p, (size_t)(s2.end - s2.p));
            else if (f2 == blob_field && wt2 == 2) {
                blobs.emplace_back();
                if (!read_blob(s2, blobs.back())) return false;
            }
        }
        if (!L.ok) return false;
        if (!blobs.empty()) out[name] = std::move(blobs);
    }
    return w.ok;
}

bool read_file(const char *path, std::vector<uint8_t> &buf) {
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(n > 0 ? (size_t)n : 0);
    const size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    fclose(f);
    return got == buf.size();
}

int load(const char *prototxt, const char *caffemodel, ParsedNet &net, std::vector<float> &weights) {
    SB_REQUIRE(prototxt && caffemodel, "null path");
    std::vector<uint8_t> txt, bin;
    if (!read_file(prototxt, txt)) { sb_set_error("cannot read %s", prototxt); return SB_ERR_INVALID; }
    if (!read_file(caffemodel, bin)) { sb_set_error("cannot read %s", caffemodel); return SB_ERR_INVALID; }
    SB_TRY(parse_prototxt(std::string(txt.begin(), txt.end()), net));
    std::map<std::string, std::vector<Blob>> model;
    if (!read_model(bin, model)) { sb_set_error("%s: not a caffemodel (protobuf wire format error)", caffemodel); return SB_ERR_INVALID; }
    size_t ci = 0;
    for (const sb_calc_layer &L : net.layers) {
        if (L.type != SB_CALC_CONV) continue;
        const std::string &name = net.conv_names[ci];
        const size_t nw = (size_t)L.num_output * net.conv_in_channels[ci] * L.kernel * L.kernel;
        ci++;
        auto it = model.find(name);
        if (it == model.end() || it->second.size() < 2) { sb_set_error("%s: no weight + bias blobs for layer '%s'", caffemodel, name.c_str()); return SB_ERR_INVALID; }
        const Blob &W = it->second[0], &B = it->second[1];
        if (W.data.size() != nw || B.data.size() != (size_t)L.num_output) {
            sb_set_error("%s: layer '%s' holds %zu + %zu values, the deploy net needs %zu + %d", caffemodel, name.c_str(), W.data.size(), B.data.size(), nw,
                         L.num_output);
            return SB_ERR_INVALID;
        }
        weights.insert(weights.end(), W.data.begin(), W.data.end());
        weights.insert(weights.end(), B.data.begin(), B.data.end());
    }
    return SB_OK;
}

}  // namespace

// Parses the two files without touching a device.  layers [cap_layers] / weights [cap_weights] may be null to query the
// sizes; shape [2] = net input (height, width).
extern "C" int sb_calc_parse_caffe(const char *prototxt_path, const char *caffemodel_path, sb_calc_layer *layers, int cap_layers, int *n_layers,
                                   float *weights, int64_t cap_weights, int64_t *n_weights, int *shape) {
    SB_NVTX_FN();
    sb_clear_error();
    ParsedNet net;
    std::vector<float> w;
    SB_TRY(load(prototxt_path, caffemodel_path, net, w));
    if (n_layers) *n_layers = (int)net.layers.size();
    if (n_weights) *n_weights = (int64_t)w.size();
    if (shape) { shape[0] = net.in_h; shape[1] = net.in_w; }
    if (layers) {
        SB_REQUIRE(cap_layers >= (int)net.layers.size(), "layer array too small");
        memcpy(layers, net.layers.data(), net.layers.size() * sizeof(sb_calc_layer));
    }
    if (weights) {
        SB_REQUIRE(cap_weights >= (int64_t)w.size(), "weight array too small");
        memcpy(weights, w.data(), w.size() * sizeof(float));
    }
    return SB_OK;
}

// DeepLCD::DeepLCD(network_definition_file, pre_trained_model_file, gpu_id) (src/deeplcd.cpp:10-31).
extern "C" int sb_calc_create_from_caffe(sb_calc_t **h, int device, const char *prototxt_path, const char *caffemodel_path, int max_batch,
                                         int max_img_w, int max_img_h) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle pointer");
    *h = nullptr;
    ParsedNet net;
    std::vector<float> w;
    SB_TRY(load(prototxt_path, caffemodel_path, net, w));
    return sb_calc_create(h, device, net.in_h, net.in_w, net.layers.data(), (int)net.layers.size(), w.data(), (int64_t)w.size(), max_batch,
                          max_img_w > net.in_w ? max_img_w : net.in_w, max_img_h > net.in_h ? max_img_h : net.in_h);
}
