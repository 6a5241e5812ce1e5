// Host build of csrc/orb_core.inl: the per-pixel arithmetic the device kernels use, exposed so the CPU
// test-suite can compare it with the oracle before anything runs on a GPU.  Test infrastructure only.
#define SB_HOST_MODEL
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/csrc/orb_core.inl"

extern "C" {
// response plane of one cv::FAST ROI as k_fast_cells builds it: score if >= t0 and the quick test passes, else 0
void hm_fast_plane(const uint8_t *img, int w, int h, int stride, int t0, uint8_t *plane) {
    memset(plane, 0, (size_t)w * h);
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = img + (size_t)y * stride + x;
            if (sb_fast_maybe(p, stride, t0)) {
                int s = sb_fast_score(p, stride);
                if (s >= t0) plane[(size_t)y * w + x] = (uint8_t)s;
            }
        }
}
int hm_fast_score(const uint8_t *p, int stride) { return sb_fast_score(p, stride); }
float hm_fast_atan2(float y, float x) { return sb_fast_atan2(y, x); }
void hm_resize(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
    for (int y = 0; y < dh; y++) {
        SbLinCoef cy = sb_lin_coef(y, dh, sh, false);
        int sy0 = cy.s < 0 ? 0 : (cy.s < sh ? cy.s : sh - 1), sy1 = cy.s + 1 < 0 ? 0 : (cy.s + 1 < sh ? cy.s + 1 : sh - 1);
        for (int x = 0; x < dw; x++) {
            SbLinCoef cx = sb_lin_coef(x, dw, sw, true);
            int sx1 = cx.s + 1 < sw ? cx.s + 1 : sw - 1;
            int h0 = src[sy0 * sw + cx.s] * cx.c0 + src[sy0 * sw + sx1] * cx.c1;
            int h1 = src[sy1 * sw + cx.s] * cx.c0 + src[sy1 * sw + sx1] * cx.c1;
            dst[y * dw + x] = sb_lin_vert(h0, h1, cy.c0, cy.c1);
        }
    }
}
void hm_gauss(const uint8_t *src, int w, int h, uint8_t *dst) {
    std::vector<unsigned> t((size_t)w * h);
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            unsigned v[7];
            for (int k = 0; k < 7; k++) v[k] = src[(size_t)y * w + sb_reflect101(x + k - 3, w)];
            t[(size_t)y * w + x] = sb_gauss_row(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
        }
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            unsigned v[7];
            for (int k = 0; k < 7; k++) v[k] = t[(size_t)sb_reflect101(y + k - 3, h) * w + x];
            dst[(size_t)y * w + x] = sb_gauss_col(v[0], v[1], v[2], v[3], v[4], v[5], v[6]);
        }
}
int hm_brief_sample(const uint8_t *center, int pitch, float a, float b, float x, float y) {
    return sb_brief_sample(center, pitch, a, b, x, y);
}
}
