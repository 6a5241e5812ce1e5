/*
 * lk_oracle.c — CPU restatement of the pyramidal Lucas-Kanade tracker the front end calls every frame.
 *
 * TEST INFRASTRUCTURE ONLY (see orb_oracle.c).
 *
 * Reference call sites (paths relative to /root/reference): Frontend::TrackLastFrame src/frontend.cpp:150-153
 * and Frontend::FindFeaturesInRight :358-361 —
 *   cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(11,11), 3,
 *                            TermCriteria(COUNT+EPS, 30, 0.01), OPTFLOW_USE_INITIAL_FLOW)
 * The arithmetic lives in OpenCV (third party, un-vendored; the author used 3.4.8, this container has cv2
 * 4.13.0).  Its published algorithm (imgproc pyrDown, video/lkpyramid.cpp: calcScharrDeriv + LKTrackerInvoker)
 * is restated here: 5x5 [1 4 6 4 1] pyramid with BORDER_REFLECT_101, Scharr derivatives as int16, bilinear
 * patch extraction with 14-bit integer weights and CV_DESCALE rounding, 2x2 normal equations in fp32, at most
 * 30 iterations with the epsilon / oscillation stops, windows may reach 11 px outside the image (reflected
 * image, zero derivatives).
 * ONE DELIBERATE DIFFERENCE: OpenCV accumulates the sums of products (A11, A12, A22, b1, b2) in fp32 in an order
 * that depends on its SIMD width; here they are accumulated exactly in int64 and rounded to fp32 once.  Both are
 * roundings of the same exact integers (relative difference ~1e-7), so tracks agree to ~1e-4 px, not to the bit.
 * Pin: tests/test_oracle_lk.py compares with cv2.calcOpticalFlowPyrLK (status flags equal, positions within
 * 2e-3 px on > 99.5 % of the points, all within 0.05 px).  The GPU kernel uses the same exact accumulation and
 * is compared BIT-EXACTLY with this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define W_BITS 14
#define DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

/* cv::pyrDown for CV_8UC1: dst = ((w+1)/2, (h+1)/2), kernel [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101 */
void orc_pyr_down_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dstride) {
    const int dw = (sw + 1) / 2, dh = (sh + 1) / 2;
    int *row = (int *)malloc(sizeof(int) * (size_t)dw * 5);
    for (int y = 0; y < dh; y++) {
        for (int k = 0; k < 5; k++) {
            const uint8_t *S = src + (size_t)reflect101(2 * y + k - 2, sh) * sstride;
            int *R = row + (size_t)k * dw;
            for (int x = 0; x < dw; x++) {
                const int x0 = reflect101(2 * x - 2, sw), x1 = reflect101(2 * x - 1, sw), x2 = 2 * x < sw ? 2 * x : reflect101(2 * x, sw);
                const int x3 = reflect101(2 * x + 1, sw), x4 = reflect101(2 * x + 2, sw);
                R[x] = S[x0] + S[x4] + 4 * (S[x1] + S[x3]) + 6 * S[x2];
            }
        }
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < dw; x++)
            D[x] = (uint8_t)((row[x] + row[4 * dw + x] + 4 * (row[dw + x] + row[3 * dw + x]) + 6 * row[2 * dw + x] + 128) >> 8);
    }
    free(row);
}

/* calcScharrDeriv: d[y][x] = (Ix, Iy) int16, borders by reflection of the image (lkpyramid.cpp) */
void orc_scharr_u8(const uint8_t *src, int w, int h, int sstride, int16_t *d /* [h][w][2] */) {
    for (int y = 0; y < h; y++) {
        const uint8_t *r0 = src + (size_t)reflect101(y - 1, h) * sstride, *r1 = src + (size_t)y * sstride;
        const uint8_t *r2 = src + (size_t)reflect101(y + 1, h) * sstride;
        for (int x = 0; x < w; x++) {
            const int xm = reflect101(x - 1, w), xp = reflect101(x + 1, w);
            const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
            const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
            d[((size_t)y * w + x) * 2] = (int16_t)(t0p - t0m);
            d[((size_t)y * w + x) * 2 + 1] = (int16_t)((t1p + t1m) * 3 + t1c * 10);
        }
    }
}

typedef struct { const uint8_t *img; int w, h, stride; } plane;
static inline int px(const plane *p, int x, int y) { return p->img[(size_t)reflect101(y, p->h) * p->stride + reflect101(x, p->w)]; }
static inline int dv(const int16_t *d, int w, int h, int x, int y, int c) {
    return (x < 0 || y < 0 || x >= w || y >= h) ? 0 : d[((size_t)y * w + x) * 2 + c];
}

/* LKTrackerInvoker for one pyramid level; next/status are updated in place. */
static void lk_level(const plane *I, const int16_t *dI, const plane *J, int n, const float *prev, float *next, uint8_t *status,
                     int level, int max_level, int win, int max_count, double eps2, float min_eig_th, int use_initial) {
    const float half = (win - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    int *Iw = (int *)malloc(sizeof(int) * (size_t)win * win * 3);
    for (int i = 0; i < n; i++) {
        const float sc = (float)(1. / (1 << level));
        float ppx = prev[2 * i] * sc, ppy = prev[2 * i + 1] * sc, nx, ny;
        if (level == max_level) {
            if (use_initial) { nx = next[2 * i] * sc; ny = next[2 * i + 1] * sc; }
            else { nx = ppx; ny = ppy; }
        } else { nx = next[2 * i] * 2.f; ny = next[2 * i + 1] * 2.f; }
        next[2 * i] = nx; next[2 * i + 1] = ny;
        ppx -= half; ppy -= half;
        const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
        if (ipx < -win || ipx >= I->w || ipy < -win || ipy >= I->h) {
            if (level == 0) status[i] = 0;
            continue;
        }
        float a = ppx - ipx, b = ppy - ipy;
        int iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS)), iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
        int iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS)), iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
        int64_t sA11 = 0, sA12 = 0, sA22 = 0;
        for (int y = 0; y < win; y++)
            for (int x = 0; x < win; x++) {
                const int X = ipx + x, Y = ipy + y;
                const int ival = DESCALE(px(I, X, Y) * iw00 + px(I, X + 1, Y) * iw01 + px(I, X, Y + 1) * iw10 + px(I, X + 1, Y + 1) * iw11, W_BITS - 5);
                const int ix = DESCALE(dv(dI, I->w, I->h, X, Y, 0) * iw00 + dv(dI, I->w, I->h, X + 1, Y, 0) * iw01 +
                                       dv(dI, I->w, I->h, X, Y + 1, 0) * iw10 + dv(dI, I->w, I->h, X + 1, Y + 1, 0) * iw11, W_BITS);
                const int iy = DESCALE(dv(dI, I->w, I->h, X, Y, 1) * iw00 + dv(dI, I->w, I->h, X + 1, Y, 1) * iw01 +
                                       dv(dI, I->w, I->h, X, Y + 1, 1) * iw10 + dv(dI, I->w, I->h, X + 1, Y + 1, 1) * iw11, W_BITS);
                int *o = Iw + ((size_t)y * win + x) * 3;
                o[0] = (int16_t)ival; o[1] = (int16_t)ix; o[2] = (int16_t)iy;
                sA11 += (int64_t)o[1] * o[1]; sA12 += (int64_t)o[1] * o[2]; sA22 += (int64_t)o[2] * o[2];
            }
        const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
        float D = A11 * A22 - A12 * A12;
        const float minEig = (A22 + A11 - sqrtf((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (float)(2 * win * win);
        if (minEig < min_eig_th || D < 1.1920929e-07f) {
            if (level == 0) status[i] = 0;
            continue;
        }
        D = 1.f / D;
        nx -= half; ny -= half;
        float pdx = 0, pdy = 0;
        for (int j = 0; j < max_count; j++) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -win || inx >= J->w || iny < -win || iny >= J->h) {
                if (level == 0) status[i] = 0;
                break;
            }
            a = nx - inx; b = ny - iny;
            iw00 = (int)lrintf((1.f - a) * (1.f - b) * (1 << W_BITS)); iw01 = (int)lrintf(a * (1.f - b) * (1 << W_BITS));
            iw10 = (int)lrintf((1.f - a) * b * (1 << W_BITS)); iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
            int64_t sb1 = 0, sb2 = 0;
            for (int y = 0; y < win; y++)
                for (int x = 0; x < win; x++) {
                    const int X = inx + x, Y = iny + y;
                    const int *o = Iw + ((size_t)y * win + x) * 3;
                    const int diff = DESCALE(px(J, X, Y) * iw00 + px(J, X + 1, Y) * iw01 + px(J, X, Y + 1) * iw10 + px(J, X + 1, Y + 1) * iw11, W_BITS - 5) - o[0];
                    sb1 += (int64_t)diff * o[1]; sb2 += (int64_t)diff * o[2];
                }
            const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
            const float dx = (A12 * b2 - A22 * b1) * D, dy = (A12 * b1 - A11 * b2) * D;
            nx += dx; ny += dy;
            next[2 * i] = nx + half; next[2 * i + 1] = ny + half;
            if ((double)dx * dx + (double)dy * dy <= eps2) break;
            if (j > 0 && fabsf(dx + pdx) < 0.01 && fabsf(dy + pdy) < 0.01) {
                next[2 * i] -= dx * 0.5f; next[2 * i + 1] -= dy * 0.5f;
                break;
            }
            pdx = dx; pdy = dy;
        }
    }
    free(Iw);
}

/* cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, ., Size(win,win), max_level, (COUNT+EPS, max_count, eps),
 * use_initial ? OPTFLOW_USE_INITIAL_FLOW : 0, min_eig_th).  next_pts is in/out when use_initial. */
int orc_lk_track(const uint8_t *prev, const uint8_t *next, int w, int h, int stride, int n, const float *prev_pts, float *next_pts,
                 uint8_t *status, int win, int max_level, int max_count, double eps, int use_initial, float min_eig_th) {
    if (w <= win || h <= win || max_level < 0 || max_level > 8) return -1;
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    const double eps2 = eps * eps;
    uint8_t *P[9], *N[9];
    int W[9], H[9];
    int levels = 0;
    for (int l = 0; l <= max_level; l++) {  /* buildOpticalFlowPyramid stops when a level is not larger than the window */
        const int lw = l == 0 ? w : (W[l - 1] + 1) / 2, lh = l == 0 ? h : (H[l - 1] + 1) / 2;
        if (l > 0 && (lw <= win || lh <= win)) break;
        W[l] = lw; H[l] = lh;
        P[l] = (uint8_t *)malloc((size_t)lw * lh);
        N[l] = (uint8_t *)malloc((size_t)lw * lh);
        if (l == 0) {
            for (int y = 0; y < h; y++) { memcpy(P[0] + (size_t)y * w, prev + (size_t)y * stride, (size_t)w); memcpy(N[0] + (size_t)y * w, next + (size_t)y * stride, (size_t)w); }
        } else {
            orc_pyr_down_u8(P[l - 1], W[l - 1], H[l - 1], W[l - 1], P[l], lw);
            orc_pyr_down_u8(N[l - 1], W[l - 1], H[l - 1], W[l - 1], N[l], lw);
        }
        levels++;
    }
    max_level = levels - 1;
    for (int i = 0; i < n; i++) status[i] = 1;
    int16_t *d = (int16_t *)malloc(sizeof(int16_t) * 2 * (size_t)w * h);
    for (int l = max_level; l >= 0; l--) {
        orc_scharr_u8(P[l], W[l], H[l], W[l], d);
        plane I = {P[l], W[l], H[l], W[l]}, J = {N[l], W[l], H[l], W[l]};
        lk_level(&I, d, &J, n, prev_pts, next_pts, status, l, max_level, win, max_count, eps2, min_eig_th, use_initial);
    }
    free(d);
    for (int l = 0; l < levels; l++) { free(P[l]); free(N[l]); }
    return 0;
}
