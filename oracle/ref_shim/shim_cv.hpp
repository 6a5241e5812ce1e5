// shim_cv.hpp — the smallest OpenCV-compatible surface that lets the UNMODIFIED reference translation unit
// /root/reference/src/ORBextractor.cpp (+ include/myslam/ORBextractor.h, common_include.h) compile here,
// where OpenCV itself is absent.
//
// TEST INFRASTRUCTURE ONLY (part of oracle/): it exists to build oracle/_ref/libmyslam_orb_ref.so, the
// reference's own extractor logic, which pins oracle/orb_oracle.c and the CUDA path.  Original code, not
// OpenCV source: only the behaviour the reference relies on is provided —
//   * cv::Mat for CV_8UC1 (ref-counted buffer, rowRange/colRange views, clone, ptr/at/step/step1, create,
//     Mat::zeros + the `Mat = MatExpr` rule that an already-sized header is written IN PLACE, quirk Q6);
//   * Point_/Size_/KeyPoint, InputArray/OutputArray as thin Mat proxies;
//   * cv::FAST, cv::resize(INTER_LINEAR), cv::GaussianBlur(7x7, sigma 2, REFLECT_101), cv::fastAtan2 —
//     forwarded to the C primitives of oracle/orb_oracle.c, each of which is pinned bit-exactly against
//     cv2 4.13.0 by tests/test_oracle_cv2.py;  cvRound/cvFloor/cvCeil with OpenCV's SSE2 semantics
//     (round half to even under the default rounding mode).
// Anything else aborts loudly (unsupported type / kernel size) rather than guessing.
#ifndef SLAMB200_ORACLE_SHIM_CV_HPP
#define SLAMB200_ORACLE_SHIM_CV_HPP

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_Assert(expr) do { if (!(expr)) { std::fprintf(stderr, "CV_Assert failed: %s\n", #expr); std::abort(); } } while (0)

typedef unsigned char uchar;

extern "C" {
// oracle/orb_oracle.c (cv2-pinned restatements of the OpenCV primitives)
int orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh, int dstride);
int orc_gauss7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride);
float orc_fast_atan2(float y, float x);
int orc_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nonmax, int *out, int cap);
}

// cvRound & co. live in the global namespace in OpenCV (core/fast_math.hpp)
static inline int cvRound(double v) { return (int)std::lrint(v); }
static inline int cvRound(float v) { return (int)std::lrintf(v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
static inline int cvFloor(int v) { return v; }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }
static inline int cvCeil(int v) { return v; }

namespace cv {

typedef ::uchar uchar;

enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T _x, T _y) : x(_x), y(_y) {}
};
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;

// OpenCV: a.x = saturate_cast<T>(a.x * b) for b of type int / float / double
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, int b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator*=(Point_<T> &a, double b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, int b) { a.x = (T)(a.x / b); a.y = (T)(a.y / b); return a; }
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, float b) { a.x = (T)(a.x / b); a.y = (T)(a.y / b); return a; }
template <typename T> static inline Point_<T> &operator/=(Point_<T> &a, double b) { a.x = (T)(a.x / b); a.y = (T)(a.y / b); return a; }

template <typename T> struct Size_ {
    T width, height;
    Size_() : width(0), height(0) {}
    Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size2i;
typedef Size2i Size;

struct KeyPoint {
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
        : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};

struct MatExpr {  // only Mat::zeros produces one
    int rows, cols, type;
};

class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;

    Mat() : rows(0), cols(0), data(nullptr), step(0) {}
    Mat(int r, int c, int t) : rows(0), cols(0), data(nullptr), step(0) { create(r, c, t); }
    // wrap caller-owned pixels (no copy, no ownership), as cv::Mat(rows, cols, type, data, step)
    Mat(int r, int c, int t, void *d, size_t s = 0) : rows(r), cols(c), data((uchar *)d), step(s ? s : (size_t)c) { check_type(t); }
    Mat(const MatExpr &e) : rows(0), cols(0), data(nullptr), step(0) { *this = e; }

    static MatExpr zeros(int r, int c, int t) { check_type(t); MatExpr e = {r, c, t}; return e; }

    // OpenCV: MatOp_Initializer::assign -> m.create(size, type) (a no-op when the header already has that size
    // and type, so a view keeps pointing into its parent) followed by m = Scalar(0).
    Mat &operator=(const MatExpr &e) {
        create(e.rows, e.cols, e.type);
        for (int r = 0; r < rows; ++r) std::memset(data + (size_t)r * step, 0, (size_t)cols);
        return *this;
    }

    void create(int r, int c, int t) {
        check_type(t);
        if (data && rows == r && cols == c) return;
        release();
        if (r <= 0 || c <= 0) return;
        buf_ = std::shared_ptr<uchar>((uchar *)std::malloc((size_t)r * c), std::free);
        CV_Assert(buf_.get() != nullptr);
        rows = r; cols = c; step = (size_t)c; data = buf_.get();
    }
    void release() { buf_.reset(); rows = cols = 0; data = nullptr; step = 0; }

    Mat clone() const {
        Mat m;
        m.create(rows, cols, CV_8UC1);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols);
        return m;
    }
    Mat rowRange(int r0, int r1) const {
        CV_Assert(0 <= r0 && r0 <= r1 && r1 <= rows);
        Mat m(*this);
        m.rows = r1 - r0; m.data = data + (size_t)r0 * step;
        return m;
    }
    Mat colRange(int c0, int c1) const {
        CV_Assert(0 <= c0 && c0 <= c1 && c1 <= cols);
        Mat m(*this);
        m.cols = c1 - c0; m.data = data + c0;
        return m;
    }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    size_t step1() const { return step; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols || rows == 1; }

    uchar *ptr(int r = 0) { return data + (size_t)r * step; }
    const uchar *ptr(int r = 0) const { return data + (size_t)r * step; }
    template <typename T> T *ptr(int r = 0) { static_assert(sizeof(T) == 1, "CV_8UC1 only"); return (T *)(data + (size_t)r * step); }
    template <typename T> const T *ptr(int r = 0) const { static_assert(sizeof(T) == 1, "CV_8UC1 only"); return (const T *)(data + (size_t)r * step); }
    template <typename T> T &at(int r, int c) { static_assert(sizeof(T) == 1, "CV_8UC1 only"); return *(T *)(data + (size_t)r * step + c); }
    template <typename T> const T &at(int r, int c) const { static_assert(sizeof(T) == 1, "CV_8UC1 only"); return *(const T *)(data + (size_t)r * step + c); }

private:
    static void check_type(int t) { if (t != CV_8UC1) { std::fprintf(stderr, "shim cv::Mat: only CV_8UC1\n"); std::abort(); } }
    std::shared_ptr<uchar> buf_;
};

// InputArray / OutputArray: thin proxies over a Mat
class _InputArray {
public:
    _InputArray() {}
    _InputArray(const Mat &m) : m_(m) {}
    Mat getMat() const { return m_; }
    bool empty() const { return m_.empty(); }
private:
    Mat m_;
};
class _OutputArray {
public:
    _OutputArray(Mat &m) : p_(&m) {}
    Mat getMat() const { return *p_; }
    Mat &getMatRef() const { return *p_; }
    bool empty() const { return p_->empty(); }
    void release() const { p_->release(); }
    void create(int r, int c, int t) const { p_->create(r, c, t); }
    void create(Size s, int t) const { p_->create(s.height, s.width, t); }
private:
    Mat *p_;
};
typedef const _InputArray &InputArray;
typedef const _OutputArray &OutputArray;
typedef const _OutputArray &InputOutputArray;

static inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

// cv::FAST(image, keypoints, threshold, nonmaxSuppression): FAST-9/16, row-major output,
// KeyPoint(x, y, 7.f, -1, score) — oracle/orb_oracle.c:orc_fast9_16 (pinned to cv2.FastFeatureDetector).
static inline void FAST(InputArray _img, std::vector<KeyPoint> &keypoints, int threshold, bool nonmaxSuppression = true) {
    Mat img = _img.getMat();
    keypoints.clear();
    if (img.empty()) return;
    const int cap = img.rows * img.cols;
    std::vector<int> out((size_t)cap * 3);
    const int n = orc_fast9_16(img.data, img.cols, img.rows, (int)img.step, threshold, nonmaxSuppression ? 1 : 0, out.data(), cap);
    CV_Assert(n >= 0 && n <= cap);
    keypoints.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
        keypoints.push_back(KeyPoint((float)out[3 * i], (float)out[3 * i + 1], 7.f, -1, (float)out[3 * i + 2]));
}

// cv::resize(src, dst, dsize, 0, 0, INTER_LINEAR): dst.create(dsize) then the fixed-point bilinear kernel.
static inline void resize(InputArray _src, OutputArray _dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR) {
    CV_Assert(interpolation == INTER_LINEAR && fx == 0 && fy == 0 && dsize.width > 0 && dsize.height > 0);
    Mat src = _src.getMat();
    _dst.create(dsize, CV_8UC1);
    Mat dst = _dst.getMat();
    CV_Assert(src.data != dst.data);
    int rc = orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
    CV_Assert(rc == 0);
}

// cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) on a whole (non-sub-) matrix; in place allowed.
static inline void GaussianBlur(InputArray _src, OutputArray _dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT) {
    CV_Assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
    Mat src = _src.getMat();
    _dst.create(src.rows, src.cols, CV_8UC1);
    Mat dst = _dst.getMat();
    Mat tmp = (src.data == dst.data) ? src.clone() : src;
    int rc = orc_gauss7_u8(tmp.data, tmp.cols, tmp.rows, (int)tmp.step, dst.data, (int)dst.step);
    CV_Assert(rc == 0);
}

}  // namespace cv
#endif
