// slamb200_adaptors.hpp — C++ host side above the C ABI (include/slamb200.h): drop-in replacements for the
// operator classes of the reference's libmyslam.so, keeping their names, signatures, argument meaning and
// error behaviour (silent return on empty input, no exceptions, no status codes) so that the three thread
// classes (Frontend, Backend, LoopClosing) compile against them unchanged.
//
//   myslam::ORBextractor               include/myslam/ORBextractor.h:47-138  (src/ORBextractor.cpp)
//   myslam::HammingMatcher::match      cv::DescriptorMatcher::match as used at src/loopclosing.cpp:33,172
//   myslam::LocalBASolver              the g2o block of Backend::OptimizeActiveMap, src/backend.cpp:126-269
//   myslam::DeepLCDScorer              DeepLCD::score + LoopClosing::DetectLoop, src/deeplcd.cpp:35-39,
//                                      src/loopclosing.cpp:124-161
//   myslam::DeepLCD                    include/myslam/deeplcd.h:20-48 (src/deeplcd.cpp:10-91): the CALC CNN forward
//   myslam::PoseGraphSolver            the g2o block of LoopClosing::PoseGraphOptimization, :537-646
//
// With OpenCV present (the reference's build) the classes take cv::InputArray / cv::KeyPoint / cv::Mat.
// Without it (this repository's CI, which has no OpenCV C++ headers) a minimal stand-in with the same
// field layout is used so that the header still compiles and is exercised by tests/host_adaptor_test.cpp.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "../../include/slamb200.h"

#if defined(SLAMB200_WITH_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core.hpp>) && !defined(SLAMB200_NO_OPENCV))
#include <opencv2/core.hpp>
#define SLAMB200_CV 1
#else
#define SLAMB200_CV 0
namespace cv {  // field-for-field stand-ins, only what the adaptors touch
struct Point2f { float x = 0, y = 0; };
struct KeyPoint {
    Point2f pt;
    float size = 0, angle = -1, response = 0;
    int octave = 0, class_id = -1;
};
struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
struct Mat {  // CV_8UC1 view or owner
    int rows = 0, cols = 0;
    size_t step = 0;
    uint8_t *data = nullptr;
    std::shared_ptr<std::vector<uint8_t>> own;
    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    void create(int r, int c) { own = std::make_shared<std::vector<uint8_t>>((size_t)r * c); data = own->data(); rows = r; cols = c; step = (size_t)c; }
    void release() { own.reset(); data = nullptr; rows = cols = 0; step = 0; }
    uint8_t *ptr(int r) { return data + (size_t)r * step; }
    const uint8_t *ptr(int r) const { return data + (size_t)r * step; }
};
typedef const Mat &InputArray;
typedef Mat &OutputArray;
}  // namespace cv
#endif

namespace myslam {

namespace detail {
inline const cv::Mat &as_mat(const cv::Mat &m) { return m; }
#if SLAMB200_CV
inline cv::Mat as_mat(cv::InputArray a) { return a.getMat(); }
#endif
static_assert(sizeof(sb_keypoint) == 28, "sb_keypoint must mirror cv::KeyPoint");
inline sb_keypoint to_sb(const cv::KeyPoint &k) { return sb_keypoint{k.pt.x, k.pt.y, k.size, k.angle, k.response, k.octave, k.class_id}; }
inline cv::KeyPoint from_sb(const sb_keypoint &k) {
    cv::KeyPoint o;
    o.pt.x = k.x; o.pt.y = k.y; o.size = k.size; o.angle = k.angle; o.response = k.response; o.octave = k.octave; o.class_id = k.class_id;
    return o;
}
// The reference logs with glog and carries on; the adaptors keep the last C-ABI status readable instead.
inline int &last_status() { static thread_local int s = SB_OK; return s; }
// Page-locked staging array (sb_host_alloc): what the asynchronous entry points need for their copies to be asynchronous.
template <typename T> class Pinned {
public:
    Pinned() = default;
    explicit Pinned(size_t n) { reset(n); }
    ~Pinned() { sb_host_free(p_); }
    Pinned(const Pinned &) = delete;
    Pinned &operator=(const Pinned &) = delete;
    void reset(size_t n) {
        sb_host_free(p_);
        p_ = nullptr; n_ = 0;
        void *q = nullptr;
        if (sb_host_alloc(&q, n * sizeof(T)) != SB_OK) throw std::runtime_error(std::string("sb_host_alloc: ") + sb_last_error());
        p_ = static_cast<T *>(q); n_ = n;
    }
    T *data() { return p_; }
    size_t size() const { return n_; }
    void fill(const T &v) { for (size_t i = 0; i < n_; i++) p_[i] = v; }
private:
    T *p_ = nullptr;
    size_t n_ = 0;
};
}  // namespace detail

// -----------------------------------------------------------------------------------------------------
class ORBextractor {
public:
    typedef std::shared_ptr<ORBextractor> Ptr;
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    // max_w / max_h bound the images this instance will see (device memory is reserved once); the
    // reference has no such limit, so the default covers KITTI and EuRoC.
    ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int max_w = 1280, int max_h = 1024,
                 int device = 0)
        : nfeatures(nfeatures), scaleFactor(scaleFactor), nlevels(nlevels), iniThFAST(iniThFAST), minThFAST(minThFAST) {
        // the reference shares one ORBextractor between the front-end thread (Detect) and the loop-closing
        // thread (ScreenAndComputeKPsParams / CalcDescriptors), src/system.cpp:54,66: one handle per caller
        for (sb_orb_t **h : {&h_detect_, &h_pyramid_}) {
            int rc = sb_orb_create(h, device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_w, max_h, 1);
            if (rc != SB_OK) throw std::runtime_error(std::string("sb_orb_create: ") + sb_last_error());  // a constructor cannot return
        }
        int n = 0;
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
        mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
        sb_orb_get_tables(h_pyramid_, &n, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                          mvInvLevelSigma2.data(), mnFeaturesPerLevel.data());
        cap_ = sb_orb_capacity(h_pyramid_);
    }
    ~ORBextractor() { sb_orb_destroy(h_detect_); sb_orb_destroy(h_pyramid_); }
    ORBextractor(const ORBextractor &) = delete;
    ORBextractor &operator=(const ORBextractor &) = delete;

    // src/ORBextractor.cpp:922-985
    void DetectAndCompute(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &keypoints, cv::OutputArray _descriptors) {
        const cv::Mat image = detail::as_mat(_image), mask = detail::as_mat(_mask);
        if (image.empty()) return;
        run_pyramid(image, mask, keypoints, &_descriptors);
    }
    // :1135-1176
    void DetectWithPyramid(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &keypoints) {
        const cv::Mat image = detail::as_mat(_image), mask = detail::as_mat(_mask);
        if (image.empty()) return;
        run_pyramid(image, mask, keypoints, nullptr);
    }
    // :989-1074 — re-entrant with respect to the two methods below (own handle)
    void Detect(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint> &keypoints) {
        const cv::Mat image = detail::as_mat(_image), mask = detail::as_mat(_mask);
        if (image.empty() || mask.empty()) return;
        std::vector<sb_keypoint> k((size_t)cap_);
        int32_t n = 0;
        const uint8_t *ip = image.data, *mp = mask.data;
        detail::last_status() = sb_orb_detect(h_detect_, 1, &ip, &mp, image.cols, image.rows, (int)image.step, (int)mask.step, k.data(), &n, cap_);
        keypoints.clear();
        if (detail::last_status() != SB_OK) return;
        for (int i = 0; i < n; i++) keypoints.push_back(detail::from_sb(k[i]));
    }
    // :1083-1129 — mutates _keypoints exactly like the reference
    void ScreenAndComputeKPsParams(cv::InputArray _image, std::vector<cv::KeyPoint> &_keypoints, std::vector<cv::KeyPoint> &out_keypoints) {
        const cv::Mat image = detail::as_mat(_image);
        if (image.empty() || _keypoints.empty()) return;
        std::vector<sb_keypoint> in(_keypoints.size()), out(_keypoints.size());
        for (size_t i = 0; i < in.size(); i++) in[i] = detail::to_sb(_keypoints[i]);
        int32_t n = 0;
        detail::last_status() = sb_orb_screen_params(h_pyramid_, image.data, image.cols, image.rows, (int)image.step, in.data(), (int)in.size(), out.data(), &n);
        out_keypoints.clear();
        if (detail::last_status() != SB_OK) return;
        for (size_t i = 0; i < in.size(); i++) _keypoints[i] = detail::from_sb(in[i]);
        for (int i = 0; i < n; i++) out_keypoints.push_back(detail::from_sb(out[i]));
    }
    // :1180-1226
    void CalcDescriptors(cv::InputArray _image, const std::vector<cv::KeyPoint> &_keypoints, cv::OutputArray _descriptors) {
        const cv::Mat image = detail::as_mat(_image);
        if (image.empty() || _keypoints.empty()) return;
        std::vector<sb_keypoint> k(_keypoints.size());
        for (size_t i = 0; i < k.size(); i++) k[i] = detail::to_sb(_keypoints[i]);
        std::vector<uint8_t> d(k.size() * 32);
        detail::last_status() = sb_orb_calc_descriptors(h_pyramid_, image.data, image.cols, image.rows, (int)image.step, k.data(), (int)k.size(), d.data());
        if (detail::last_status() != SB_OK) return;
        store_descriptors(_descriptors, d.data(), (int)k.size());
    }

    // Both calls of LoopClosing::ProcessNewKF (src/loopclosing.cpp:107-112) in one pass over ONE pyramid (the reference builds
    // it twice, quirk Q7): _keypoints mutated like ScreenAndComputeKPsParams does, survivors + their descriptors out.
    void ScreenAndDescribe(cv::InputArray _image, std::vector<cv::KeyPoint> &_keypoints, std::vector<cv::KeyPoint> &out_keypoints,
                           cv::OutputArray _descriptors) {
        const cv::Mat image = detail::as_mat(_image);
        if (image.empty() || _keypoints.empty()) return;
        const int32_t n_in = (int32_t)_keypoints.size();
        std::vector<sb_keypoint> in((size_t)n_in), out((size_t)n_in);
        for (size_t i = 0; i < in.size(); i++) in[i] = detail::to_sb(_keypoints[i]);
        std::vector<uint8_t> d((size_t)n_in * 32);
        int32_t n = 0;
        const uint8_t *ip = image.data;
        detail::last_status() = sb_orb_screen_describe(h_pyramid_, 1, &ip, image.cols, image.rows, (int)image.step, in.data(), &n_in, n_in,
                                                       out.data(), &n, d.data());
        out_keypoints.clear();
        if (detail::last_status() != SB_OK) return;
        for (size_t i = 0; i < in.size(); i++) _keypoints[i] = detail::from_sb(in[i]);
        for (int i = 0; i < n; i++) out_keypoints.push_back(detail::from_sb(out[i]));
        if (n > 0) store_descriptors(_descriptors, d.data(), n);
    }

    int GetLevels() { return nlevels; }
    float GetScaleFactor() { return (float)scaleFactor; }
    std::vector<float> GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

    // The reference exposes mvImagePyramid / mvMaskPyramid as public members; here the levels live on the
    // device and are copied out on demand (which: 0 image, 1 blurred, 2 mask).
    bool GetPyramidLevel(int level, int which, cv::Mat &out) {
        int lw = 0, lh = 0;
        if (sb_orb_debug_level(h_pyramid_, 0, level, which, nullptr, 0, &lw, &lh) != SB_OK) return false;
        mat_create(out, lh, lw);
        return sb_orb_debug_level(h_pyramid_, 0, level, which, out.data, lw * lh, &lw, &lh) == SB_OK;
    }

protected:
    static void mat_create(cv::Mat &m, int rows, int cols) {
#if SLAMB200_CV
        m.create(rows, cols, CV_8UC1);
#else
        m.create(rows, cols);
#endif
    }
    typedef std::remove_reference<cv::OutputArray>::type OutArrT;
    static void store_descriptors(OutArrT &dst, const uint8_t *d, int n) {
#if SLAMB200_CV
        dst.create(n, 32, CV_8U);
        cv::Mat m = dst.getMat();
        for (int i = 0; i < n; i++) std::memcpy(m.ptr(i), d + (size_t)i * 32, 32);
#else
        dst.create(n, 32);
        for (int i = 0; i < n; i++) std::memcpy(dst.ptr(i), d + (size_t)i * 32, 32);
#endif
    }
    void run_pyramid(const cv::Mat &image, const cv::Mat &mask, std::vector<cv::KeyPoint> &keypoints, OutArrT *descriptors) {
        std::vector<sb_keypoint> k((size_t)cap_);
        std::vector<uint8_t> d(descriptors ? (size_t)cap_ * 32 : 0);
        int32_t n = 0;
        const uint8_t *ip = image.data, *mp = mask.empty() ? nullptr : mask.data;
        detail::last_status() = sb_orb_detect_and_compute(h_pyramid_, 1, &ip, mp ? &mp : nullptr, image.cols, image.rows, (int)image.step,
                                                          mp ? (int)mask.step : 0, k.data(), descriptors ? d.data() : nullptr, &n, cap_);
        keypoints.clear();
        if (detail::last_status() != SB_OK) return;
        keypoints.reserve(n);
        for (int i = 0; i < n; i++) keypoints.push_back(detail::from_sb(k[i]));
        if (descriptors) {
            if (n == 0) descriptors->release();
            else store_descriptors(*descriptors, d.data(), n);
        }
    }

    int nfeatures;
    double scaleFactor;  // a double holding the float argument, like the reference (ORBextractor.h:125)
    int nlevels, iniThFAST, minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    sb_orb_t *h_detect_ = nullptr, *h_pyramid_ = nullptr;
    int cap_ = 0;
};

// -----------------------------------------------------------------------------------------------------
// cv::DescriptorMatcher::create("BruteForce-Hamming")->match(query, train, matches)  (src/loopclosing.cpp:33,172)
class HammingMatcher {
public:
    explicit HammingMatcher(int max_rows = 8192, int device = 0) : max_rows_(max_rows) {
        if (sb_matcher_create(&h_, device, 1, max_rows) != SB_OK) throw std::runtime_error(std::string("sb_matcher_create: ") + sb_last_error());
    }
    ~HammingMatcher() { sb_matcher_destroy(h_); }
    void match(cv::InputArray _query, cv::InputArray _train, std::vector<cv::DMatch> &matches) {
        const cv::Mat q = detail::as_mat(_query), t = detail::as_mat(_train);
        matches.clear();
        if (q.empty() || t.empty()) return;
        const int cap = q.rows > t.rows ? q.rows : t.rows;
        if (cap > max_rows_) { detail::last_status() = SB_ERR_CAPACITY; return; }   // more descriptors than the handle was created for
        if (q.cols != 32 || t.cols != 32) { detail::last_status() = SB_ERR_INVALID; return; }
        std::vector<uint8_t> qb((size_t)cap * 32), tb((size_t)cap * 32);
        for (int i = 0; i < q.rows; i++) std::memcpy(&qb[(size_t)i * 32], q.ptr(i), 32);
        for (int i = 0; i < t.rows; i++) std::memcpy(&tb[(size_t)i * 32], t.ptr(i), 32);
        std::vector<int32_t> idx(cap), dist(cap);
        const int32_t nq = q.rows, nt = t.rows;
        detail::last_status() = sb_hamming_match(h_, 1, qb.data(), &nq, tb.data(), &nt, cap, idx.data(), dist.data());
        if (detail::last_status() != SB_OK) return;
        for (int i = 0; i < nq; i++) {
            cv::DMatch m;
            m.queryIdx = i; m.trainIdx = idx[i]; m.imgIdx = 0; m.distance = (float)dist[i];
            matches.push_back(m);
        }
    }
private:
    sb_matcher_t *h_ = nullptr;
    int max_rows_;
};

// -----------------------------------------------------------------------------------------------------
// The solver of Backend::OptimizeActiveMap.  The Backend keeps its graph-building loops (src/backend.cpp:135-206)
// but fills these flat arrays instead of g2o vertices/edges, calls Optimize(), and reads poses, points and the
// per-edge chi2 back for its outlier handling and write-back (:236-266).
struct LocalBAProblem {
    std::vector<double> poses;       // [n_poses][7] qx qy qz qw tx ty tz  (Sophus::SE3d::data() order)
    std::vector<double> points;      // [n_points][3]
    std::vector<uint8_t> fixed;      // [n_points]
    std::vector<int32_t> obs_pose, obs_point;
    std::vector<double> uv;          // [n_obs][2]
    double K[4] = {0, 0, 0, 0};      // fx fy cx cy
    double cam_ext[7] = {0, 0, 0, 1, 0, 0, 0};
    std::vector<double> chi2;        // out
    std::vector<uint8_t> outlier;    // out
    int info[4] = {0, 0, 0, 0};      // out: outer rounds, LM iterations, inliers, outliers
};

class LocalBASolver {
public:
    LocalBASolver(int max_poses = 7, int max_points = 4096, int max_obs = 32768, int device = 0)
        : mp_(max_poses), ml_(max_points), mo_(max_obs) {
        if (sb_ba_create(&h_, device, 1, max_poses, max_points, max_obs) != SB_OK) throw std::runtime_error(std::string("sb_ba_create: ") + sb_last_error());
        // page-locked staging, allocated once: Submit's copies are then truly asynchronous (with pageable vectors
        // cudaMemcpyAsync blocks and Submit would only return when the window is solved)
        poses_.reset((size_t)mp_ * 7); points_.reset((size_t)ml_ * 3); uv_.reset((size_t)mo_ * 2); chi2_.reset((size_t)mo_);
        fixed_.reset((size_t)ml_); outl_.reset((size_t)mo_); op_.reset((size_t)mo_); ol_.reset((size_t)mo_);
        cnt_.reset(3); info_.reset(4);
    }
    ~LocalBASolver() { sb_ba_destroy(h_); }
    // chi2_th 5.991, Huber delta 5.991, up to 5 rounds of optimize(10): src/backend.cpp:155,198-200,212-232
    bool Optimize(LocalBAProblem &p, double huber_delta = 5.991, double chi2_th = 5.991, int outer_max = 5, int inner_iters = 10) {
        return Submit(p, huber_delta, chi2_th, outer_max, inner_iters) && Wait();
    }
    // The asynchronous halves (the reference's Backend is its own thread, src/backend.cpp:29-45): Submit enqueues the window and
    // returns, Wait blocks until it is solved and writes the results back into the problem given to Submit, which must stay alive.
    bool Submit(LocalBAProblem &p, double huber_delta = 5.991, double chi2_th = 5.991, int outer_max = 5, int inner_iters = 10) {
        const size_t np = p.poses.size() / 7, nl = p.points.size() / 3, ne = p.obs_pose.size();
        if (pending_) { detail::last_status() = SB_ERR_INVALID; return false; }
        if (np > (size_t)mp_ || nl > (size_t)ml_ || ne > (size_t)mo_) { detail::last_status() = SB_ERR_CAPACITY; return false; }
        // the flat arrays must agree with each other: an oversized vector would overrun the staging, a short one feed zeros
        if (p.poses.size() != np * 7 || p.points.size() != nl * 3 || p.fixed.size() != nl || p.obs_point.size() != ne || p.uv.size() != 2 * ne) {
            detail::last_status() = SB_ERR_INVALID;
            return false;
        }
        poses_.fill(0.0); points_.fill(0.0); uv_.fill(0.0); fixed_.fill(0); op_.fill(0); ol_.fill(0);
        std::copy(p.poses.begin(), p.poses.end(), poses_.data());
        std::copy(p.points.begin(), p.points.end(), points_.data());
        std::copy(p.fixed.begin(), p.fixed.end(), fixed_.data());
        std::copy(p.obs_pose.begin(), p.obs_pose.end(), op_.data());
        std::copy(p.obs_point.begin(), p.obs_point.end(), ol_.data());
        std::copy(p.uv.begin(), p.uv.end(), uv_.data());
        cnt_.data()[0] = (int32_t)np; cnt_.data()[1] = (int32_t)nl; cnt_.data()[2] = (int32_t)ne;
        detail::last_status() = sb_ba_submit(h_, 1, cnt_.data(), cnt_.data() + 1, cnt_.data() + 2, poses_.data(), points_.data(), fixed_.data(),
                                             op_.data(), ol_.data(), uv_.data(), p.K, p.cam_ext, huber_delta, chi2_th, outer_max, inner_iters,
                                             chi2_.data(), outl_.data(), info_.data());
        pending_ = detail::last_status() == SB_OK ? &p : nullptr;
        return pending_ != nullptr;
    }
    bool Wait() {
        if (!pending_) { detail::last_status() = SB_ERR_INVALID; return false; }
        LocalBAProblem &p = *pending_;
        pending_ = nullptr;
        detail::last_status() = sb_ba_wait(h_);
        if (detail::last_status() != SB_OK) return false;
        const size_t np = (size_t)cnt_.data()[0], nl = (size_t)cnt_.data()[1], ne = (size_t)cnt_.data()[2];
        std::copy(poses_.data(), poses_.data() + np * 7, p.poses.begin());
        std::copy(points_.data(), points_.data() + nl * 3, p.points.begin());
        p.chi2.assign(chi2_.data(), chi2_.data() + ne);
        p.outlier.assign(outl_.data(), outl_.data() + ne);
        for (int k = 0; k < 4; k++) p.info[k] = info_.data()[k];
        return true;
    }
private:
    sb_ba_t *h_ = nullptr;
    int mp_, ml_, mo_;
    LocalBAProblem *pending_ = nullptr;
    detail::Pinned<double> poses_, points_, uv_, chi2_;
    detail::Pinned<uint8_t> fixed_, outl_;
    detail::Pinned<int32_t> op_, ol_, cnt_, info_;
};

// -----------------------------------------------------------------------------------------------------
// DeepLCD::score over the database + the scan of LoopClosing::DetectLoop (src/loopclosing.cpp:124-161).
class DeepLCDScorer {
public:
    explicit DeepLCDScorer(int capacity = 4096, bool fp16_database = true, int device = 0) {
        if (sb_lcd_create(&h_, device, capacity, fp16_database ? SB_LCD_FP16 : SB_LCD_FP32, 1) != SB_OK)
            throw std::runtime_error(std::string("sb_lcd_create: ") + sb_last_error());
    }
    ~DeepLCDScorer() { sb_lcd_destroy(h_); }
    void AddToDatabase(unsigned long kfId, const float *descr1064) { detail::last_status() = sb_lcd_add(h_, (int64_t)kfId, descr1064); }
    void Erase(unsigned long kfId) { detail::last_status() = sb_lcd_remove(h_, (int64_t)kfId); }
    // returns true and the candidate id when the reference's DetectLoop would (thresholds: yaml LCD.similarityScoreThreshold.*)
    bool DetectLoop(unsigned long curKFId, const float *descr1064, float thresHigh, float thresLow, unsigned long &bestId, float &maxScore) {
        int found = 0, cnt = 0;
        int64_t best = 0;
        detail::last_status() = sb_lcd_detect_loop(h_, (int64_t)curKFId, descr1064, thresHigh, thresLow, 20, 3, &found, &best, &maxScore, &cnt);
        bestId = (unsigned long)best;
        return detail::last_status() == SB_OK && found != 0;
    }
private:
    sb_lcd_t *h_ = nullptr;
};

// -----------------------------------------------------------------------------------------------------
// myslam::DeepLCD (include/myslam/deeplcd.h:20-48, src/deeplcd.cpp): same constructor arguments and method names.
// DescrVector is Eigen::Matrix<float, 1064, 1> in the reference; here any contiguous float container of
// descrDim() entries (std::vector<float> by default; Eigen::Map it where Eigen is present).  gpu_id = -1 ("CPU" in
// the reference) selects device 0: there is no CPU path.
class DeepLCD {
public:
    typedef std::vector<float> DescrVector;
    typedef std::shared_ptr<DeepLCD> Ptr;
    DeepLCD(const std::string &network_definition_file = "calc_model/deploy.prototxt",
            const std::string &pre_trained_model_file = "calc_model/calc.caffemodel", int gpu_id = -1, int max_img_w = 2048, int max_img_h = 1024) {
        if (sb_calc_create_from_caffe(&h_, gpu_id < 0 ? 0 : gpu_id, network_definition_file.c_str(), pre_trained_model_file.c_str(), 1, max_img_w,
                                      max_img_h) != SB_OK)
            throw std::runtime_error(std::string("sb_calc_create_from_caffe: ") + sb_last_error());
    }
    // the network as data (what the two files hold)
    DeepLCD(const std::vector<sb_calc_layer> &layers, const std::vector<float> &weights, int in_h, int in_w, int gpu_id = 0, int max_img_w = 2048,
            int max_img_h = 1024) {
        if (sb_calc_create(&h_, gpu_id < 0 ? 0 : gpu_id, in_h, in_w, layers.data(), (int)layers.size(), weights.data(), (int64_t)weights.size(), 1,
                           max_img_w, max_img_h) != SB_OK)
            throw std::runtime_error(std::string("sb_calc_create: ") + sb_last_error());
    }
    ~DeepLCD() { sb_calc_destroy(h_); }
    DeepLCD(const DeepLCD &) = delete;
    DeepLCD &operator=(const DeepLCD &) = delete;
    int descrDim() const { return sb_calc_descr_dim(h_); }
    // src/deeplcd.cpp:35-39
    float score(const DescrVector &d1, const DescrVector &d2) const {
        float r = 0.f;
        for (size_t i = 0; i < d1.size() && i < d2.size(); i++) r += d1[i] * d2[i];
        return r;
    }
    // :43-52 — like the reference, the caller's image comes back blurred (cv::GaussianBlur(originalImg, originalImg, ...))
    DescrVector calcDescrOriginalImg(const cv::Mat &originalImg) {
        DescrVector d((size_t)descrDim(), 0.f);
        if (originalImg.empty()) { detail::last_status() = SB_ERR_INVALID; return d; }
        const uint8_t *in = originalImg.data;
        uint8_t *out = originalImg.data;
        detail::last_status() = sb_calc_descr_original(h_, 1, &in, originalImg.cols, originalImg.rows, (int)originalImg.step, d.data(), &out);
        return d;
    }
    // :55-91 — the image must already have the net's input size
    DescrVector calcDescr(const cv::Mat &im) {
        DescrVector d((size_t)descrDim(), 0.f);
        if (im.empty()) { detail::last_status() = SB_ERR_INVALID; return d; }
        int in_h = 0, in_w = 0;
        sb_calc_input_size(h_, &in_h, &in_w);
        if (im.rows != in_h || im.cols != in_w) { detail::last_status() = SB_ERR_INVALID; return d; }   // the reference resizes before calcDescr
        const uint8_t *in = im.data;
        detail::last_status() = sb_calc_descr(h_, 1, &in, (int)im.step, d.data());
        return d;
    }
private:
    sb_calc_t *h_ = nullptr;
};

// -----------------------------------------------------------------------------------------------------
// The solver of LoopClosing::PoseGraphOptimization (src/loopclosing.cpp:537-646): vertices in ascending keyframe id.
class PoseGraphSolver {
public:
    // max_loops bounds the loop edges of one optimisation; the reference re-adds every historical loop edge on each call
    // (src/loopclosing.cpp:585-599), so it must cover the whole run (KITTI-00 ends with 17)
    PoseGraphSolver(int max_vertices = 8192, int max_edges = 16384, int device = 0, int max_loops = 256) {
        if (sb_posegraph_create_loops(&h_, device, max_vertices, max_edges, max_loops) != SB_OK) throw std::runtime_error(std::string("sb_posegraph_create: ") + sb_last_error());
    }
    ~PoseGraphSolver() { sb_posegraph_destroy(h_); }
    bool Optimize(std::vector<double> &poses7, const std::vector<uint8_t> &fixed, const std::vector<int32_t> &v0, const std::vector<int32_t> &v1,
                  const std::vector<double> &meas7, int iters = 20) {
        int32_t info[4];
        double stats[2];
        detail::last_status() = sb_posegraph_solve(h_, (int)(poses7.size() / 7), poses7.data(), fixed.data(), (int)v0.size(), v0.data(), v1.data(),
                                                   meas7.data(), iters, info, stats);
        return detail::last_status() == SB_OK;
    }
private:
    sb_posegraph_t *h_ = nullptr;
};

// -----------------------------------------------------------------------------------------------------
// The solver of Frontend::EstimateCurrentPose (src/frontend.cpp:176-276, preRounds = 0) and of
// LoopClosing::OptimizeCurrentPose (src/loopclosing.cpp:339-433, preRounds = 1).
class PoseOnlySolver {
public:
    explicit PoseOnlySolver(int max_obs = 4096, int device = 0) : mo_(max_obs) {
        if (sb_pose_create(&h_, device, 1, max_obs) != SB_OK) throw std::runtime_error(std::string("sb_pose_create: ") + sb_last_error());
    }
    ~PoseOnlySolver() { sb_pose_destroy(h_); }
    // pose7: qx qy qz qw tx ty tz (in/out); points3 / uv2: one row per matched map point; returns the number of
    // inliers (what the reference's functions return) and fills `outlier`.
    int Optimize(double pose7[7], const std::vector<double> &points3, const std::vector<double> &uv2, const double K[4],
                 std::vector<uint8_t> &outlier, int preRounds = 0, double chi2_th = 5.991) {
        const int32_t n = (int32_t)(uv2.size() / 2);
        if (n > mo_) { detail::last_status() = SB_ERR_CAPACITY; return -1; }
        std::vector<double> p((size_t)mo_ * 3, 0.0), u((size_t)mo_ * 2, 0.0);
        std::copy(points3.begin(), points3.end(), p.begin());
        std::copy(uv2.begin(), uv2.end(), u.begin());
        std::vector<uint8_t> o((size_t)mo_);
        int32_t info[4];
        detail::last_status() = sb_pose_solve(h_, 1, &n, pose7, p.data(), u.data(), K, 1.0, chi2_th, preRounds, 4, 10, o.data(), info);
        if (detail::last_status() != SB_OK) return -1;
        outlier.assign(o.begin(), o.begin() + n);
        return info[0];
    }
private:
    sb_pose_t *h_ = nullptr;
    int mo_;
};

// -----------------------------------------------------------------------------------------------------
// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(11,11), 3, (COUNT+EPS, 30, 0.01),
// OPTFLOW_USE_INITIAL_FLOW) as called at src/frontend.cpp:150-153 and :358-361.  nextPts carries the initial guess in
// and the tracked position out; status[i] = 1 where the track succeeded.
class LKTracker {
public:
    LKTracker(int max_w = 1280, int max_h = 1024, int max_pts = 4096, int device = 0) : mp_(max_pts) {
        if (sb_lk_create(&h_, device, max_w, max_h, 1, max_pts, 3) != SB_OK) throw std::runtime_error(std::string("sb_lk_create: ") + sb_last_error());
    }
    ~LKTracker() { sb_lk_destroy(h_); }
    bool calcOpticalFlowPyrLK(cv::InputArray _prev, cv::InputArray _next, const std::vector<float> &prevPts2, std::vector<float> &nextPts2,
                              std::vector<uint8_t> &status) {
        const cv::Mat prev = detail::as_mat(_prev), next = detail::as_mat(_next);
        const int32_t n = (int32_t)(prevPts2.size() / 2);
        if (prev.empty() || next.empty() || n > mp_) { detail::last_status() = SB_ERR_INVALID; return false; }
        std::vector<float> pp((size_t)mp_ * 2, 0.f), np((size_t)mp_ * 2, 0.f);
        std::copy(prevPts2.begin(), prevPts2.end(), pp.begin());
        std::copy(nextPts2.begin(), nextPts2.begin() + 2 * n, np.begin());
        std::vector<uint8_t> st((size_t)mp_);
        const uint8_t *pi = prev.data, *ni = next.data;
        detail::last_status() = sb_lk_track(h_, 1, &pi, &ni, prev.cols, prev.rows, (int)prev.step, &n, pp.data(), np.data(), st.data(), 11, 30, 0.01, 1, 1e-4f);
        if (detail::last_status() != SB_OK) return false;
        nextPts2.assign(np.begin(), np.begin() + 2 * n);
        status.assign(st.begin(), st.begin() + n);
        return true;
    }
private:
    sb_lk_t *h_ = nullptr;
    int mp_;
};

// -----------------------------------------------------------------------------------------------------
// myslam::triangulation (include/myslam/algorithm.h:16-33) for all left/right correspondences of a frame at once,
// including the callers' `&& p[2] > 0` (src/frontend.cpp:403,474); Twc7 != nullptr maps accepted points to the world.
inline bool triangulation_batch(const std::vector<float> &uvLeft, const std::vector<float> &uvRight, const double Kl[4], const double Kr[4],
                                const double poseLeft7[7], const double poseRight7[7], const double *Twc7,
                                std::vector<double> &points3, std::vector<uint8_t> &ok, int device = 0) {
    const int n = (int)(uvLeft.size() / 2);
    points3.assign((size_t)n * 3, 0.0);
    ok.assign((size_t)n, 0);
    detail::last_status() = sb_triangulate(device, n, uvLeft.data(), uvRight.data(), Kl, Kr, poseLeft7, poseRight7, Twc7, 1e-2,
                                           points3.data(), ok.data());
    return detail::last_status() == SB_OK;
}

// -----------------------------------------------------------------------------------------------------
// SE(3) helpers on the host for the two decisions below (Sophus storage order qx qy qz qw tx ty tz; T = [R | t]).
namespace detail {
struct Rt { double R[9], t[3]; };
inline Rt rt_from_pose7(const double p[7]) {
    const double n = std::sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    const double x = p[0] / n, y = p[1] / n, z = p[2] / n, w = p[3] / n;
    Rt o;
    o.R[0] = 1 - 2 * (y * y + z * z); o.R[1] = 2 * (x * y - z * w);     o.R[2] = 2 * (x * z + y * w);
    o.R[3] = 2 * (x * y + z * w);     o.R[4] = 1 - 2 * (x * x + z * z); o.R[5] = 2 * (y * z - x * w);
    o.R[6] = 2 * (x * z - y * w);     o.R[7] = 2 * (y * z + x * w);     o.R[8] = 1 - 2 * (x * x + y * y);
    o.t[0] = p[4]; o.t[1] = p[5]; o.t[2] = p[6];
    return o;
}
inline Rt rt_mul(const Rt &A, const Rt &B) {
    Rt C;
    for (int i = 0; i < 3; i++) {
        for (int j = 0; j < 3; j++) C.R[3 * i + j] = A.R[3 * i] * B.R[j] + A.R[3 * i + 1] * B.R[3 + j] + A.R[3 * i + 2] * B.R[6 + j];
        C.t[i] = A.R[3 * i] * B.t[0] + A.R[3 * i + 1] * B.t[1] + A.R[3 * i + 2] * B.t[2] + A.t[i];
    }
    return C;
}
inline Rt rt_inv(const Rt &A) {
    Rt C;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C.R[3 * i + j] = A.R[3 * j + i];
    for (int i = 0; i < 3; i++) C.t[i] = -(C.R[3 * i] * A.t[0] + C.R[3 * i + 1] * A.t[1] + C.R[3 * i + 2] * A.t[2]);
    return C;
}
// |SE3::log(T)| (the 6-vector [upsilon, omega], Sophus' formulas)
inline double rt_log_norm(const Rt &T) {
    const double c = std::min(1.0, std::max(-1.0, 0.5 * (T.R[0] + T.R[4] + T.R[8] - 1.0)));
    const double th = std::acos(c);
    double w[3] = {T.R[7] - T.R[5], T.R[2] - T.R[6], T.R[3] - T.R[1]};
    if (th < 1e-10) { for (double &v : w) v *= 0.5; }
    else if (3.14159265358979323846 - th < 1e-6) {  // angle ~ pi: axis from the diagonal
        double ax[3] = {std::sqrt(std::max(0.0, 0.5 * (T.R[0] + 1))), std::sqrt(std::max(0.0, 0.5 * (T.R[4] + 1))), std::sqrt(std::max(0.0, 0.5 * (T.R[8] + 1)))};
        if (T.R[1] + T.R[3] < 0) ax[1] = -ax[1];
        if (T.R[2] + T.R[6] < 0) ax[2] = -ax[2];
        for (int i = 0; i < 3; i++) w[i] = th * ax[i];
    } else { const double k = th / (2.0 * std::sin(th)); for (double &v : w) v *= k; }
    const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
    double coef;
    if (th2 < 1e-20) coef = 1.0 / 12.0;
    else { const double t1 = std::sqrt(th2), half = 0.5 * t1; coef = (1.0 - t1 * std::cos(half) / (2.0 * std::sin(half))) / th2; }
    const double W[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
    double u[3];
    for (int i = 0; i < 3; i++) {
        u[i] = 0;
        for (int j = 0; j < 3; j++) {
            const double w2 = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
            u[i] += ((i == j ? 1.0 : 0.0) - 0.5 * W[3 * i + j] + coef * w2) * T.t[j];
        }
    }
    return std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2] + th2);
}
}  // namespace detail

// -----------------------------------------------------------------------------------------------------
// LoopClosing::ComputeCorrectPose (src/loopclosing.cpp:207-293) without the map walking: the caller collects, for the
// valid feature matches whose loop feature has a map point, the map point positions (cv::Point3f) and the current
// keyframe's keypoints (cv::Point2f).  solvePnPRansac (:263-264) -> OptimizeCurrentPose (:275, :339-433) ->
// "fewer than 10 inliers: reject" (:279-281) -> need-correct test |log(T_current * T_corrected^-1)| > 1 (:284-289).
struct LoopVerification {
    bool accepted = false;      // ComputeCorrectPose's return value
    bool needCorrect = false;   // _mbNeedCorrect
    int inliers = 0;            // OptimizeCurrentPose's return value
    double correctedPose[7] = {0, 0, 0, 1, 0, 0, 0};  // _mseCorrectedCurrentPose (T_cw)
    std::vector<uint8_t> outlier;                      // matches OptimizeCurrentPose erases from _msetValidFeatureMatches
};
class LoopVerifier {
public:
    explicit LoopVerifier(int max_points = 4096, int device = 0) : mp_(max_points), refine_(max_points, device) {
        if (sb_pnp_create(&h_, device, 1, max_points) != SB_OK) throw std::runtime_error(std::string("sb_pnp_create: ") + sb_last_error());
    }
    ~LoopVerifier() { sb_pnp_destroy(h_); }
    LoopVerification ComputeCorrectPose(const std::vector<float> &loopPoints3, const std::vector<float> &currentPoints2, const double K[4],
                                        const double currentPose7[7], uint64_t seed = 0) {
        LoopVerification out;
        const int32_t n = (int32_t)(currentPoints2.size() / 2);
        if (n < 10) return out;                                                        // :255-256
        if (n > mp_) { detail::last_status() = SB_ERR_CAPACITY; return out; }
        std::vector<float> obj((size_t)mp_ * 3, 0.f), img((size_t)mp_ * 2, 0.f);
        std::copy(loopPoints3.begin(), loopPoints3.begin() + 3 * (size_t)n, obj.begin());
        std::copy(currentPoints2.begin(), currentPoints2.begin() + 2 * (size_t)n, img.begin());
        std::vector<uint8_t> inl((size_t)mp_);
        double rt[6];
        int32_t info[4];
        detail::last_status() = sb_pnp_ransac(h_, 1, &n, obj.data(), img.data(), K, 100, 5.991, seed, out.correctedPose, rt, inl.data(), info);
        if (detail::last_status() != SB_OK || !info[0]) return out;                    // the reference's try / catch (:262-267)
        std::vector<double> p3((size_t)n * 3), uv((size_t)n * 2);
        for (size_t i = 0; i < p3.size(); i++) p3[i] = loopPoints3[i];
        for (size_t i = 0; i < uv.size(); i++) uv[i] = currentPoints2[i];
        out.inliers = refine_.Optimize(out.correctedPose, p3, uv, K, out.outlier, /*preRounds=*/1);
        if (out.inliers < 10) return out;                                              // :279-281
        const detail::Rt d = detail::rt_mul(detail::rt_from_pose7(currentPose7), detail::rt_inv(detail::rt_from_pose7(out.correctedPose)));
        out.needCorrect = detail::rt_log_norm(d) > 1.0;                                // :284-289
        out.accepted = true;
        return out;
    }
private:
    sb_pnp_t *h_ = nullptr;
    int mp_;
    PoseOnlySolver refine_;
};

// -----------------------------------------------------------------------------------------------------
// Map::RemoveOldActiveKeyframe (src/map.cpp:78-120), the choice only: which active keyframe leaves the sliding window.
// poses: (keyframe id, T_cw as pose7) of the active keyframes other than the current one, in the container's order;
// returns the id to remove (the reference's quirks kept: "else if" means a keyframe that raises the maximum is never
// considered for the minimum, and the ids start out as 0).
inline unsigned long SelectActiveKeyframeToRemove(const std::vector<std::pair<unsigned long, const double *>> &poses, const double currentPose7[7]) {
    double maxDis = 0, minDis = 9999;
    unsigned long maxKFId = 0, minKFId = 0;
    const detail::Rt Twc = detail::rt_inv(detail::rt_from_pose7(currentPose7));
    for (const auto &kf : poses) {
        const double dis = detail::rt_log_norm(detail::rt_mul(detail::rt_from_pose7(kf.second), Twc));
        if (dis > maxDis) { maxDis = dis; maxKFId = kf.first; }
        else if (dis < minDis) { minDis = dis; minKFId = kf.first; }
    }
    const double minDisTh = 0.2;
    return minDis < minDisTh ? minKFId : maxKFId;
}

}  // namespace myslam
