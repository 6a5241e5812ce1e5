// Compiles the C++ adaptors (host/slamb200_adaptors.hpp) without OpenCV and checks their behaviour where no GPU
// is needed: constructors surface the "no CUDA device, no CPU fallback" error; with a device (argv[1] = "gpu")
// the ORB adaptor runs the reference's call sequence on a synthetic image.
#define SLAMB200_NO_OPENCV
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/host/slamb200_adaptors.hpp"

int main(int argc, char **argv) {
    const bool gpu = argc > 1 && !std::strcmp(argv[1], "gpu");
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
    return 0;
}
