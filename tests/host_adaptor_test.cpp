// Compiles the C++ adaptors (host/slamb200_adaptors.hpp) without OpenCV and checks their behaviour where no GPU
// is needed: constructors surface the "no CUDA device, no CPU fallback" error; with a device (argv[1] = "gpu")
// the ORB adaptor runs the reference's call sequence on a synthetic image.
#define SLAMB200_NO_OPENCV
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/host/slamb200_adaptors.hpp"

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
    {   // host-only logic: Map::RemoveOldActiveKeyframe's choice (src/map.cpp:78-120)
        const double cur[7] = {0, 0, 0, 1, 0, 0, -10.0};           // T_cw of a camera at z = 10
        const double near_[7] = {0, 0, 0, 1, 0, 0, -9.9}, mid[7] = {0, 0, 0, 1, 0, 0, -6.0}, far_[7] = {0, 0, 0, 1, 0, 0, 0.0};
        std::vector<std::pair<unsigned long, const double *>> kfs = {{3, far_}, {5, mid}, {7, near_}};
        if (myslam::SelectActiveKeyframeToRemove(kfs, cur) != 7) return 20;   // a keyframe closer than 0.2 goes first
        kfs.pop_back();
        if (myslam::SelectActiveKeyframeToRemove(kfs, cur) != 3) return 21;   // otherwise the farthest one
        const double rot[7] = {0, std::sin(0.25), 0, std::cos(0.25), 0, 0, 0};  // 0.5 rad about y
        const double nrm = myslam::detail::rt_log_norm(myslam::detail::rt_from_pose7(rot));
        if (std::fabs(nrm - 0.5) > 1e-12) return 22;
    }
    if (!gpu) {
        try {
            myslam::ORBextractor e(300, 1.2f, 8, 20, 7);
            std::puts("constructed (a CUDA device is present)");
        } catch (const std::exception &ex) {
            std::printf("expected failure without a device: %s\n", ex.what());
            if (!std::strstr(ex.what(), "no CPU fallback")) return 1;
        }
        return 0;
    }
    cv::Mat img, mask;
    img.create(376, 1241);
    mask.create(376, 1241);
    std::memset(mask.data, 255, (size_t)376 * 1241);
    unsigned s = 12345;
    for (int i = 0; i < 376 * 1241; i++) img.data[i] = 120;
    for (int r = 0; r < 600; r++) {
        s = s * 1664525u + 1013904223u; int x = (s >> 8) % 1200;
        s = s * 1664525u + 1013904223u; int y = (s >> 8) % 350;
        s = s * 1664525u + 1013904223u; int v = (s >> 8) % 256;
        for (int yy = y; yy < y + 14; yy++) for (int xx = x; xx < x + 20; xx++) img.data[yy * 1241 + xx] = (uint8_t)v;
    }
    myslam::ORBextractor ext(300, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> kps, screened;
    ext.Detect(img, mask, kps);                                   // Frontend::DetectFeatures
    std::printf("Detect: %zu keypoints, size %.0f angle %.0f\n", kps.size(), kps.empty() ? 0.f : kps[0].size, kps.empty() ? 0.f : kps[0].angle);
    if (kps.empty() || kps[0].size != 7.f || kps[0].angle != -1.f) return 2;
    std::vector<cv::KeyPoint> pyr;                                // LoopClosing::ProcessNewKF: 8 octaves per feature
    for (size_t i = 0; i < kps.size(); i++)
        for (int l = 0; l < 8; l++) { cv::KeyPoint k = kps[i]; k.octave = l; k.class_id = (int)i; pyr.push_back(k); }
    ext.ScreenAndComputeKPsParams(img, pyr, screened);
    cv::Mat desc;
    ext.CalcDescriptors(img, screened, desc);
    std::printf("Screen: %zu -> %zu, descriptors %d x %d\n", pyr.size(), screened.size(), desc.rows, desc.cols);
    if (screened.empty() || desc.rows != (int)screened.size() || desc.cols != 32) return 3;
    myslam::HammingMatcher matcher;
    std::vector<cv::DMatch> m;
    matcher.match(desc, desc, m);
    for (size_t i = 0; i < m.size(); i++) if (m[i].distance != 0.f) return 4;     // every row matches itself (or an identical twin)
    std::printf("match: %zu matches, all distance 0\n", m.size());
    std::vector<cv::KeyPoint> all; cv::Mat d2;
    ext.DetectAndCompute(img, mask, all, d2);
    std::printf("DetectAndCompute: %zu keypoints\n", all.size());
    if (all.size() <= 200) return 5;
    // LoopClosing::ProcessNewKF: the whole-image descriptor; the keyframe image comes back blurred (src/deeplcd.cpp:46)
    std::vector<sb_calc_layer> layers = {{SB_CALC_CONV, 8, 5, 2, 4, 0, 0, 0, 0}, {SB_CALC_RELU, 0, 0, 0, 0, 0, 0, 0, 0},
                                         {SB_CALC_POOL_MAX, 0, 3, 2, 0, 0, 0, 0, 0}, {SB_CALC_LRN, 0, 0, 0, 0, 5, 1e-4f, 0.75f, 1.f}};
    std::vector<float> weights(8 * 25 + 8);
    for (size_t i = 0; i < weights.size(); i++) { s = s * 1664525u + 1013904223u; weights[i] = ((int)((s >> 8) % 2001) - 1000) * 1e-4f; }
    myslam::DeepLCD lcd(layers, weights, 120, 160);
    cv::Mat kf;
    kf.create(376, 1241);
    std::memcpy(kf.data, img.data, (size_t)376 * 1241);
    myslam::DeepLCD::DescrVector d1 = lcd.calcDescrOriginalImg(kf), dsame = lcd.calcDescrOriginalImg(img);
    const float self = lcd.score(d1, d1);
    std::printf("DeepLCD: dim %d, self score %.6f, image blurred in place: %s\n", lcd.descrDim(), self,
                std::memcmp(kf.data, img.data, (size_t)376 * 1241) == 0 ? "both" : "?");
    if (lcd.descrDim() != 8 * 31 * 41 || self < 0.9999f || self > 1.0001f) return 6;
    {   // LoopClosing::ComputeCorrectPose: 300 map points seen under a pose 2 m away from the drifted current pose
        const double K[4] = {718.856, 718.856, 607.1928, 185.2157};
        const double truth[7] = {0, std::sin(0.05), 0, std::cos(0.05), 0.5, 0.0, 2.0}, drifted[7] = {0, 0, 0, 1, 0, 0, 0};
        const myslam::detail::Rt T = myslam::detail::rt_from_pose7(truth);
        std::vector<float> p3, p2;
        for (int i = 0; i < 300; i++) {
            s = s * 1664525u + 1013904223u; const double xc = ((s >> 8) % 4000) / 100.0 - 20.0;
            s = s * 1664525u + 1013904223u; const double yc = ((s >> 8) % 800) / 100.0 - 4.0;
            s = s * 1664525u + 1013904223u; const double zc = ((s >> 8) % 4000) / 100.0 + 6.0;
            const double c[3] = {xc - T.t[0], yc - T.t[1], zc - T.t[2]};             // world = R^T (cam - t)
            for (int k = 0; k < 3; k++) p3.push_back((float)(T.R[k] * c[0] + T.R[3 + k] * c[1] + T.R[6 + k] * c[2]));
            const bool wrong = i % 4 == 0;                                            // 25 % wrong matches
            s = s * 1664525u + 1013904223u;
            p2.push_back(wrong ? (float)((s >> 8) % 1241) : (float)(K[0] * xc / zc + K[2]));
            s = s * 1664525u + 1013904223u;
            p2.push_back(wrong ? (float)((s >> 8) % 376) : (float)(K[1] * yc / zc + K[3]));
        }
        myslam::LoopVerifier verifier;
        const myslam::LoopVerification v = verifier.ComputeCorrectPose(p3, p2, K, drifted);
        std::printf("ComputeCorrectPose: accepted %d, inliers %d, needCorrect %d, t = %.3f %.3f %.3f\n", (int)v.accepted, v.inliers,
                    (int)v.needCorrect, v.correctedPose[4], v.correctedPose[5], v.correctedPose[6]);
        if (!v.accepted || v.inliers < 200 || !v.needCorrect) return 7;
        if (std::fabs(v.correctedPose[4] - 0.5) > 0.02 || std::fabs(v.correctedPose[6] - 2.0) > 0.02) return 8;
    }
    {   // Backend::OptimizeActiveMap through the adaptor: Submit / Wait on page-locked staging, argument checks, and the
        // one-pass ProcessNewKF call
        const double K[4] = {718.856, 718.856, 607.1928, 185.2157};
        myslam::LocalBAProblem p;
        for (int i = 0; i < 4; i++) { const double z = 1.5 * i; const double q[7] = {0, 0, 0, 1, 0, 0, -z}; p.poses.insert(p.poses.end(), q, q + 7); }
        for (int j = 0; j < 60; j++) {
            s = s * 1664525u + 1013904223u; const double x = ((s >> 8) % 2000) / 100.0 - 10.0;
            s = s * 1664525u + 1013904223u; const double y = ((s >> 8) % 600) / 100.0 - 3.0;
            s = s * 1664525u + 1013904223u; const double z = ((s >> 8) % 3000) / 100.0 + 12.0;
            p.points.push_back(x + 0.05); p.points.push_back(y - 0.05); p.points.push_back(z + 0.1);   // perturbed start
            p.fixed.push_back(j % 3 == 0);
            for (int i = 0; i < 4; i++) {
                const double zc = z - 1.5 * i;
                p.obs_pose.push_back(i); p.obs_point.push_back(j);
                p.uv.push_back(K[0] * x / zc + K[2]); p.uv.push_back(K[1] * y / zc + K[3]);
            }
        }
        for (int k = 0; k < 4; k++) p.K[k] = K[k];
        myslam::LocalBASolver ba(7, 256, 2048);
        myslam::LocalBAProblem bad = p;
        bad.fixed.pop_back();                                            // arrays that disagree are refused, not padded with zeros
        if (ba.Submit(bad) || myslam::detail::last_status() != SB_ERR_INVALID) return 9;
        bad = p;
        bad.uv.push_back(0.0);
        if (ba.Optimize(bad) || myslam::detail::last_status() != SB_ERR_INVALID) return 10;
        if (!ba.Submit(p)) return 11;
        if (ba.Submit(p)) return 12;                                     // one window in flight per solver
        if (!ba.Wait()) return 13;
        std::printf("LocalBA: rounds %d, LM iterations %d, inliers %d, outliers %d\n", p.info[0], p.info[1], p.info[2], p.info[3]);
        if (p.info[1] < 1 || p.info[2] < 200 || p.chi2.size() != 240) return 14;
        std::vector<cv::KeyPoint> pyr2, kept;
        for (size_t i = 0; i < kps.size(); i++)
            for (int l = 0; l < 8; l++) { cv::KeyPoint k = kps[i]; k.octave = l; k.class_id = (int)i; pyr2.push_back(k); }
        cv::Mat d3, d4;
        std::vector<cv::KeyPoint> pyr3 = pyr2, kept2;                    // img was blurred in place by calcDescrOriginalImg above:
        ext.ScreenAndComputeKPsParams(img, pyr3, kept2);                 // the two-call sequence again, on the image as it is now
        ext.CalcDescriptors(img, kept2, d4);
        ext.ScreenAndDescribe(img, pyr2, kept, d3);                      // both ProcessNewKF calls over one pyramid
        if (kept.empty() || kept.size() != kept2.size() || d3.rows != d4.rows || std::memcmp(d3.data, d4.data, (size_t)d4.rows * 32) != 0) return 15;
        for (size_t i = 0; i < kept.size(); i++)
            if (kept[i].pt.x != kept2[i].pt.x || kept[i].angle != kept2[i].angle || kept[i].class_id != kept2[i].class_id) return 17;
        std::printf("ScreenAndDescribe: %zu survivors, descriptors identical to the two-call sequence\n", kept.size());
        myslam::HammingMatcher tiny(8);
        std::vector<cv::DMatch> mm;
        tiny.match(desc, desc, mm);                                      // more rows than the handle holds: reported, not silently empty
        if (!mm.empty() || myslam::detail::last_status() != SB_ERR_CAPACITY) return 16;
    }
    return 0;
}
