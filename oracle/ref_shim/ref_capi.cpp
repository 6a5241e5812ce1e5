// ref_capi.cpp — extern "C" door into the UNMODIFIED reference extractor (/root/reference/src/ORBextractor.cpp),
// compiled against the shim headers in this directory.  TEST INFRASTRUCTURE ONLY (oracle/): the product never
// loads oracle/_ref/libmyslam_orb_ref.so; tests use it to pin oracle/orb_oracle.c and the CUDA path to the
// reference's own source.
//
// Nothing here restates extractor logic: every entry point marshals plain buffers into cv::Mat /
// std::vector<cv::KeyPoint> and calls the reference's public methods (include/myslam/ORBextractor.h:47-138).
// Protected members are reached through a derived class, which needs no change to the reference.
//
// Quadtree tie-break (src/ORBextractor.cpp:731): the reference sorts (size, ExtractorNode*) pairs, so equal
// sizes are ordered by HEAP ADDRESS.  ref_set_monotone_nodes(1) makes that defined: std::list nodes of
// ExtractorNode are then served from a bump arena whose addresses only grow, so "higher address" == "created
// later", the rule oracle/orb_oracle.c documents.  With 0 the nodes come from glibc malloc
// and the order is whatever the heap gives — tests measure how often that differs.
#include <new>
#include <list>
#include <vector>
#include <sys/mman.h>

#include "myslam/ORBextractor.h"

namespace shim_glog { int error_count = 0; }

// ------------------------------------------------------------------------------------------------------------
// list-node arena (see header comment).  Only allocations of exactly sizeof(list node of ExtractorNode) are
// served from it, and only while the monotone mode is on; everything else is malloc/free.
// ------------------------------------------------------------------------------------------------------------
namespace {
const size_t kNodeSize = sizeof(std::_List_node<myslam::ExtractorNode>);
const size_t kArenaBytes = size_t(1) << 30;  // virtual reservation, touched pages only
struct Arena {
    char *base = nullptr;
    size_t used = 0;
    long live = 0;
};
thread_local Arena t_arena;
thread_local int t_monotone = 0;
// recording of the addresses handed out for list nodes, in allocation order (either mode)
thread_local uint64_t *t_rec = nullptr;
thread_local int t_rec_cap = 0, t_rec_n = 0;
inline void record(void *p) { if (t_rec) { if (t_rec_n < t_rec_cap) t_rec[t_rec_n] = (uint64_t)(uintptr_t)p; ++t_rec_n; } }

inline bool in_arena(const void *p) {
    const Arena &a = t_arena;
    return a.base && (const char *)p >= a.base && (const char *)p < a.base + kArenaBytes;
}
void arena_rewind_if_idle() {
    Arena &a = t_arena;
    if (a.live == 0) a.used = 0;
}
}  // namespace

void *operator new(size_t n) {
    if (t_monotone && n == kNodeSize) {
        Arena &a = t_arena;
        if (!a.base) {
            void *m = mmap(nullptr, kArenaBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (m == MAP_FAILED) throw std::bad_alloc();
            a.base = (char *)m;
        }
        const size_t sz = (n + 15) & ~size_t(15);
        if (a.used + sz > kArenaBytes) throw std::bad_alloc();
        void *p = a.base + a.used;
        a.used += sz;
        ++a.live;
        record(p);
        return p;
    }
    void *p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    if (n == kNodeSize) record(p);
    return p;
}
void operator delete(void *p) noexcept {
    if (!p) return;
    if (in_arena(p)) { --t_arena.live; return; }
    std::free(p);
}
void operator delete(void *p, size_t) noexcept { operator delete(p); }
void *operator new[](size_t n) { return operator new(n); }
void operator delete[](void *p) noexcept { operator delete(p); }
void operator delete[](void *p, size_t) noexcept { operator delete(p); }

// ------------------------------------------------------------------------------------------------------------
namespace {
struct RefExtractor : public myslam::ORBextractor {
    using myslam::ORBextractor::ORBextractor;
    const std::vector<int> &quotas() const { return mnFeaturesPerLevel; }
    const std::vector<int> &umaxTable() const { return umax; }
    std::vector<cv::KeyPoint> distribute(const std::vector<cv::KeyPoint> &k, int minX, int maxX, int minY, int maxY, int N, int level) {
        return DistributeOctTree(k, minX, maxX, minY, maxY, N, level);
    }
};

struct ref_keypoint {  // field-for-field cv::KeyPoint, 28 bytes (same as orc_keypoint / sb_keypoint)
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

cv::Mat wrap(const uint8_t *p, int w, int h, int stride) { return cv::Mat(h, w, CV_8UC1, (void *)p, (size_t)stride); }
cv::Mat full_mask(int w, int h) {
    cv::Mat m(h, w, CV_8UC1);
    std::memset(m.data, 255, (size_t)w * h);
    return m;
}
int put_kps(const std::vector<cv::KeyPoint> &v, ref_keypoint *out, int cap) {
    const int n = (int)v.size();
    for (int i = 0; i < n && i < cap; ++i) {
        const cv::KeyPoint &k = v[i];
        ref_keypoint r = {k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id};
        out[i] = r;
    }
    return n;
}
std::vector<cv::KeyPoint> get_kps(const ref_keypoint *in, int n) {
    std::vector<cv::KeyPoint> v((size_t)n);
    for (int i = 0; i < n; ++i) {
        v[i] = cv::KeyPoint(in[i].x, in[i].y, in[i].size, in[i].angle, in[i].response, in[i].octave, in[i].class_id);
    }
    return v;
}
}  // namespace

extern "C" {

void ref_set_monotone_nodes(int on) { t_monotone = on ? 1 : 0; }
int ref_error_count(void) { return shim_glog::error_count; }
// arm (buf != NULL) / disarm the recording of list-node addresses; the count restarts at 0 when armed
void ref_record_node_addresses(uint64_t *buf, int cap) { t_rec = buf; t_rec_cap = buf ? cap : 0; t_rec_n = 0; }
int ref_recorded_count(void) { return t_rec_n; }
int ref_list_node_bytes(void) { return (int)kNodeSize; }

void *ref_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    return new RefExtractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_destroy(void *h) { delete (RefExtractor *)h; }

// tables of the constructor (src/ORBextractor.cpp:384-445) through the reference's own getters
void ref_get_tables(void *h, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2, int *quota, int *umax) {
    RefExtractor *e = (RefExtractor *)h;
    const int n = e->GetLevels();
    std::vector<float> a = e->GetScaleFactors(), b = e->GetInverseScaleFactors(), c = e->GetScaleSigmaSquares(), d = e->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; quota[i] = e->quotas()[i]; }
    for (int i = 0; i < 16; ++i) umax[i] = e->umaxTable()[i];
}

// ORBextractor::DetectAndCompute (:922-985).  mask == NULL means "all 255" (the reference always gets a mask).
int ref_detect_and_compute(void *h, const uint8_t *image, const uint8_t *mask, int w, int hgt, int stride, int mstride,
                           ref_keypoint *kps, uint8_t *desc, int cap) {
    arena_rewind_if_idle();
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img = wrap(image, w, hgt, stride);
    cv::Mat msk = mask ? wrap(mask, w, hgt, mstride) : full_mask(w, hgt);
    std::vector<cv::KeyPoint> v;
    cv::Mat d;
    e->DetectAndCompute(img, msk, v, d);
    const int n = put_kps(v, kps, cap);
    for (int i = 0; i < n && i < cap && i < d.rows; ++i) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    return n;
}
// ORBextractor::DetectWithPyramid (:1132-1175)
int ref_detect_with_pyramid(void *h, const uint8_t *image, const uint8_t *mask, int w, int hgt, int stride, int mstride,
                            ref_keypoint *kps, int cap) {
    arena_rewind_if_idle();
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img = wrap(image, w, hgt, stride);
    cv::Mat msk = mask ? wrap(mask, w, hgt, mstride) : full_mask(w, hgt);
    std::vector<cv::KeyPoint> v;
    e->DetectWithPyramid(img, msk, v);
    return put_kps(v, kps, cap);
}
// ORBextractor::Detect (:989-1074)
int ref_detect(void *h, const uint8_t *image, const uint8_t *mask, int w, int hgt, int stride, int mstride, ref_keypoint *kps, int cap) {
    arena_rewind_if_idle();
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img = wrap(image, w, hgt, stride);
    cv::Mat msk = mask ? wrap(mask, w, hgt, mstride) : full_mask(w, hgt);
    std::vector<cv::KeyPoint> v;
    e->Detect(img, msk, v);
    return put_kps(v, kps, cap);
}
// ORBextractor::ScreenAndComputeKPsParams (:1083-1129): `in` is mutated in place like the reference's vector.
int ref_screen_params(void *h, const uint8_t *image, int w, int hgt, int stride, ref_keypoint *in, int n_in, ref_keypoint *out) {
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img = wrap(image, w, hgt, stride);
    std::vector<cv::KeyPoint> vin = get_kps(in, n_in), vout;
    e->ScreenAndComputeKPsParams(img, vin, vout);
    put_kps(vin, in, n_in);
    return put_kps(vout, out, n_in);
}
// ORBextractor::CalcDescriptors (:1180-1226)
int ref_calc_descriptors(void *h, const uint8_t *image, int w, int hgt, int stride, const ref_keypoint *kps, int n, uint8_t *desc) {
    RefExtractor *e = (RefExtractor *)h;
    cv::Mat img = wrap(image, w, hgt, stride);
    std::vector<cv::KeyPoint> v = get_kps(kps, n);
    cv::Mat d;
    e->CalcDescriptors(img, v, d);
    for (int i = 0; i < d.rows && i < n; ++i) std::memcpy(desc + (size_t)i * 32, d.ptr(i), 32);
    return d.rows;
}
// public member mvImagePyramid (ORBextractor.h:106) after any call that builds it
int ref_get_level(void *h, int level, uint8_t *out, int *w, int *hgt) {
    RefExtractor *e = (RefExtractor *)h;
    if (level < 0 || level >= (int)e->mvImagePyramid.size()) return -1;
    const cv::Mat &m = e->mvImagePyramid[level];
    *w = m.cols; *hgt = m.rows;
    if (out) for (int r = 0; r < m.rows; ++r) std::memcpy(out + (size_t)r * m.cols, m.ptr(r), (size_t)m.cols);
    return 0;
}
// ORBextractor::DistributeOctTree (:586-810) on a caller-supplied candidate list; returns the kept candidates
// as (x, y, response) triples in the reference's output order.
int ref_distribute_octtree(void *h, const float *kx, const float *ky, const float *kr, int n, int minX, int maxX, int minY, int maxY,
                           int N, float *out_xyr) {
    arena_rewind_if_idle();
    RefExtractor *e = (RefExtractor *)h;
    std::vector<cv::KeyPoint> v((size_t)n);
    for (int i = 0; i < n; ++i) v[i] = cv::KeyPoint(kx[i], ky[i], 7.f, -1, kr[i]);
    std::vector<cv::KeyPoint> r = e->distribute(v, minX, maxX, minY, maxY, N, 0);
    for (size_t i = 0; i < r.size(); ++i) { out_xyr[3 * i] = r[i].pt.x; out_xyr[3 * i + 1] = r[i].pt.y; out_xyr[3 * i + 2] = r[i].response; }
    return (int)r.size();
}

}  // extern "C"
