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
    return all.size() > 200 ? 0 : 5;
}
