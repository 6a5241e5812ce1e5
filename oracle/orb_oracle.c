/*
 * orb_oracle.c — CPU restatement of the reference ORB extractor + BF-Hamming matcher.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the CPU baseline.
 *
 * Parity status: PINNED TO THE REFERENCE'S OWN SOURCE for everything the reference itself implements
 * (ORBextractor: tables, pyramid driver, grid FAST driver + fallback + mask quirk, quadtree, IC_Angle, rBRIEF,
 * DetectAndCompute / Detect / DetectWithPyramid / ScreenAndComputeKPsParams / CalcDescriptors):
 * oracle/_ref/libmyslam_orb_ref.so is /root/reference/src/ORBextractor.cpp compiled unmodified (oracle/Makefile
 * `ref`) and tests/test_ref_pin.py holds this file to it bit for bit — all 200 golden frames, every operator,
 * and, for the one undefined spot (heap-address tie-break at :731), under the addresses of a real glibc run.
 * The OpenCV primitives the reference calls (resize, GaussianBlur, FAST, fastAtan2, BFMatcher) are un-vendored
 * third party; they are restated below from the published algorithms and each is pinned bit-exactly against the
 * in-container cv2 4.13.0 (tests/test_oracle_cv2.py).  Every function cites the lines it restates (paths
 * relative to /root/reference).
 *
 * Plain C99, no dependencies.  Scalar float arithmetic is kept un-contracted
 * (build with -ffp-contract=off) because the reference is compiled for baseline
 * x86-64 (no FMA) and OpenCV's cvRound is round-half-to-even.
 */
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_LEVELS 32
#define PATCH_SIZE 31       /* src/ORBextractor.cpp:23 */
#define HALF_PATCH_SIZE 15  /* :24 */
#define EDGE_THRESHOLD 19   /* :25 */

typedef struct {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} orc_keypoint; /* field-for-field cv::KeyPoint */

static const int8_t k_pattern[1024] = {
#include "orb_pattern.inc"
};

/* cvRound: round half to even (SSE cvtss2si / lrint under the default rounding mode). */
static inline int cv_round_f(float v) { return (int)lrintf(v); }
static inline int cv_round_d(double v) { return (int)lrint(v); }
static inline int cv_floor_f(float v) { return (int)floorf(v); }

/* ------------------------------------------------------------------------------------------
 * cv::resize(..., INTER_LINEAR) for CV_8UC1 (OpenCV imgproc resize.cpp: fixed-point
 * HResizeLinear/VResizeLinear, INTER_RESIZE_COEF_BITS = 11).  Called by the reference at
 * src/ORBextractor.cpp:1243-1244 and :1262.  OpenCV is un-vendored; this restates its
 * published algorithm and is pinned against cv2 4.13.0 in tests/test_oracle_cv2.py.
 * ------------------------------------------------------------------------------------------ */
int orc_resize_linear_u8(const uint8_t *src, int sw, int sh, int sstride, uint8_t *dst, int dw, int dh,
                         int dstride) {
    if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
    const double scale_x = 1.0 / ((double)dw / sw);
    const double scale_y = 1.0 / ((double)dh / sh);
    int *xofs = (int *)malloc(sizeof(int) * (size_t)dw);
    short *ialpha = (short *)malloc(sizeof(short) * 2 * (size_t)dw);
    int *hbuf0 = (int *)malloc(sizeof(int) * (size_t)dw);
    int *hbuf1 = (int *)malloc(sizeof(int) * (size_t)dw);
    for (int dx = 0; dx < dw; dx++) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = cv_floor_f(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        ialpha[2 * dx] = (short)cv_round_f((1.f - fx) * 2048.f);
        ialpha[2 * dx + 1] = (short)cv_round_f(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; dy++) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = cv_floor_f(fy);
        fy -= sy;
        short b0 = (short)cv_round_f((1.f - fy) * 2048.f);
        short b1 = (short)cv_round_f(fy * 2048.f);
        int sy0 = sy < 0 ? 0 : (sy < sh ? sy : sh - 1);
        int sy1 = (sy + 1) < 0 ? 0 : ((sy + 1) < sh ? sy + 1 : sh - 1);
        const uint8_t *S0 = src + (size_t)sy0 * sstride, *S1 = src + (size_t)sy1 * sstride;
        for (int dx = 0; dx < dw; dx++) {
            int sx = xofs[dx], sx1 = sx + 1 < sw ? sx + 1 : sw - 1;
            int a0 = ialpha[2 * dx], a1 = ialpha[2 * dx + 1];
            hbuf0[dx] = S0[sx] * a0 + S0[sx1] * a1;
            hbuf1[dx] = S1[sx] * a0 + S1[sx1] * a1;
        }
        uint8_t *D = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; dx++)
            D[dx] = (uint8_t)((((b0 * (hbuf0[dx] >> 4)) >> 16) + ((b1 * (hbuf1[dx] >> 4)) >> 16) + 2) >> 2);
    }
    free(xofs); free(ialpha); free(hbuf0); free(hbuf1);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * cv::GaussianBlur(src, dst, Size(7,7), 2, 2, BORDER_REFLECT_101) for a whole CV_8UC1 Mat
 * (reference call sites src/ORBextractor.cpp:965-966, :1194-1199).  OpenCV's u8 path is the
 * fixed-point separable filter: 8.8 kernel [18,34,48,56,48,34,18], 8.8 row intermediates,
 * 16.16 column accumulation, round by (+32768)>>16.  Pinned against cv2 4.13.0.
 * ------------------------------------------------------------------------------------------ */
static const int k_gauss7[7] = {18, 34, 48, 56, 48, 34, 18};

static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

int orc_gauss7_u8(const uint8_t *src, int w, int h, int sstride, uint8_t *dst, int dstride) {
    if (w <= 0 || h <= 0) return -1;
    uint16_t *tmp = (uint16_t *)malloc(sizeof(uint16_t) * (size_t)w * h);
    for (int y = 0; y < h; y++) {
        const uint8_t *S = src + (size_t)y * sstride;
        uint16_t *T = tmp + (size_t)y * w;
        for (int x = 0; x < w; x++) {
            unsigned acc = 0;
            for (int k = 0; k < 7; k++) acc += (unsigned)k_gauss7[k] * S[reflect101(x + k - 3, w)];
            T[x] = (uint16_t)acc; /* <= 255*256 */
        }
    }
    for (int y = 0; y < h; y++) {
        const uint16_t *R[7];
        for (int k = 0; k < 7; k++) R[k] = tmp + (size_t)reflect101(y + k - 3, h) * w;
        uint8_t *D = dst + (size_t)y * dstride;
        for (int x = 0; x < w; x++) {
            uint32_t acc = 0;
            for (int k = 0; k < 7; k++) acc += (uint32_t)k_gauss7[k] * R[k][x];
            D[x] = (uint8_t)((acc + 32768u) >> 16);
        }
    }
    free(tmp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * cv::fastAtan2(y, x) scalar fp32 (OpenCV core mathfuncs_core: atan2 polynomial, degrees).
 * Called by the reference at src/ORBextractor.cpp:54.  Pinned against cv2.fastAtan2.
 * ------------------------------------------------------------------------------------------ */
float orc_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    float ax = fabsf(x), ay = fabsf(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

/* ------------------------------------------------------------------------------------------
 * cv::FAST(image, keypoints, threshold, nonmaxSuppression=true), TYPE_9_16 (OpenCV features2d
 * fast.cpp / fast_score.cpp).  Called per grid cell by the reference at
 * src/ORBextractor.cpp:858-865 and :1035-1042.
 *   - ring offsets as the reference re-declares them at :365-369;
 *   - a pixel (3 <= x < w-3, 3 <= y < h-3 of THIS image/ROI) is a corner at threshold t iff 9
 *     contiguous ring pixels are all > v+t or all < v-t;
 *   - response = cornerScore<16> = largest t' for which it is still a corner
 *              = max over the 16 arcs of min over the arc of (v - ring) resp. (ring - v), minus 1;
 *   - NMS: keep iff response > each of the 8 neighbours' responses (non-corners and untested
 *     border pixels count as 0); output row-major; pt integer, size 7, angle -1, octave 0,
 *     class_id -1.
 * Pinned against cv2.FastFeatureDetector in tests/test_oracle_cv2.py.
 * ------------------------------------------------------------------------------------------ */
static const int k_ring[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                  {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static int fast_score_at(const uint8_t *p, const int *pix) {
    int d[25];
    const int v = p[0];
    for (int k = 0; k < 16; k++) d[k] = v - p[pix[k]];
    for (int k = 16; k < 25; k++) d[k] = d[k - 16];
    int best = -256;
    for (int s = 0; s < 16; s++) {
        int mn = d[s], mx = d[s];
        for (int k = 1; k < 9; k++) {
            if (d[s + k] < mn) mn = d[s + k];
            if (d[s + k] > mx) mx = d[s + k];
        }
        if (mn > best) best = mn;   /* all darker:   min(v - ring)  */
        if (-mx > best) best = -mx; /* all brighter: min(ring - v)  */
    }
    return best - 1;
}

/* out: packed triples (x, y, response) as int; returns count (may exceed cap: truncated). */
int orc_fast9_16(const uint8_t *img, int w, int h, int stride, int threshold, int nonmax, int *out, int cap) {
    int n = 0;
    if (w < 7 || h < 7) return 0;
    threshold = threshold < 0 ? 0 : (threshold > 255 ? 255 : threshold);
    int pix[16];
    for (int k = 0; k < 16; k++) pix[k] = k_ring[k][0] + k_ring[k][1] * stride;
    int *score = (int *)calloc((size_t)w * h, sizeof(int));
    uint8_t *is_corner = (uint8_t *)calloc((size_t)w * h, 1);
    for (int y = 3; y < h - 3; y++) {
        const uint8_t *row = img + (size_t)y * stride;
        for (int x = 3; x < w - 3; x++) {
            const uint8_t *p = row + x;
            const int v = p[0];
            /* every 9-arc contains ring[0] or ring[8], and ring[4] or ring[12]: cheap reject */
            int d0 = v - p[pix[0]], d8 = v - p[pix[8]];
            if (abs(d0) <= threshold && abs(d8) <= threshold) continue;
            int d4 = v - p[pix[4]], d12 = v - p[pix[12]];
            if (abs(d4) <= threshold && abs(d12) <= threshold) continue;
            int s = fast_score_at(p, pix);
            if (s >= threshold) { /* corner@t  <=>  score >= t */
                score[(size_t)y * w + x] = s;
                is_corner[(size_t)y * w + x] = 1;
            }
        }
    }
    for (int y = 3; y < h - 3; y++)
        for (int x = 3; x < w - 3; x++) {
            if (!is_corner[(size_t)y * w + x]) continue;
            const int *c = score + (size_t)y * w + x;
            const int s = c[0];
            if (nonmax && !(s > c[-1] && s > c[1] && s > c[-w - 1] && s > c[-w] && s > c[-w + 1] &&
                            s > c[w - 1] && s > c[w] && s > c[w + 1]))
                continue;
            if (n < cap) {
                out[3 * n] = x;
                out[3 * n + 1] = y;
                out[3 * n + 2] = s;
            }
            n++;
        }
    free(is_corner);
    free(score);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Extractor object: tables of ORBextractor::ORBextractor (src/ORBextractor.cpp:384-445).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int nfeatures;
    double scaleFactor; /* the member is a double holding the float argument (ORBextractor.h:125) */
    int nlevels, iniThFAST, minThFAST;
    float scale[ORC_MAX_LEVELS], inv_scale[ORC_MAX_LEVELS], sigma2[ORC_MAX_LEVELS], inv_sigma2[ORC_MAX_LEVELS];
    int quota[ORC_MAX_LEVELS];
    int umax[HALF_PATCH_SIZE + 1];
    /* mvImagePyramid / mvMaskPyramid (tight rows: stride == width) */
    uint8_t *img[ORC_MAX_LEVELS], *mask[ORC_MAX_LEVELS];
    int w[ORC_MAX_LEVELS], h[ORC_MAX_LEVELS];
} orc_extractor;

orc_extractor *orc_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST) {
    if (nlevels < 1 || nlevels > ORC_MAX_LEVELS) return NULL;
    orc_extractor *e = (orc_extractor *)calloc(1, sizeof(orc_extractor));
    e->nfeatures = nfeatures;
    e->scaleFactor = scaleFactor;
    e->nlevels = nlevels;
    e->iniThFAST = iniThFAST;
    e->minThFAST = minThFAST;
    /* :388-404 */
    e->scale[0] = 1.0f;
    e->sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) {
        e->scale[i] = (float)(e->scale[i - 1] * e->scaleFactor);
        e->sigma2[i] = e->scale[i] * e->scale[i];
    }
    for (int i = 0; i < nlevels; i++) {
        e->inv_scale[i] = 1.0f / e->scale[i];
        e->inv_sigma2[i] = 1.0f / e->sigma2[i];
    }
    /* :409-421 */
    float factor = (float)(1.0f / e->scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int level = 0; level < nlevels - 1; level++) {
        e->quota[level] = cv_round_f(nDesired);
        sum += e->quota[level];
        nDesired *= factor;
    }
    e->quota[nlevels - 1] = nfeatures - sum > 0 ? nfeatures - sum : 0;
    /* :427-444 */
    int v, v0;
    int vmax = cv_floor_f(HALF_PATCH_SIZE * sqrtf(2.f) / 2 + 1);
    int vmin = (int)ceilf(HALF_PATCH_SIZE * sqrtf(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) e->umax[v] = cv_round_d(sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
        while (e->umax[v0] == e->umax[v0 + 1]) ++v0;
        e->umax[v] = v0;
        ++v0;
    }
    return e;
}

static void free_pyramid(orc_extractor *e) {
    for (int l = 0; l < ORC_MAX_LEVELS; l++) {
        free(e->img[l]); e->img[l] = NULL;
        free(e->mask[l]); e->mask[l] = NULL;
    }
}

void orc_destroy(orc_extractor *e) {
    if (!e) return;
    free_pyramid(e);
    free(e);
}

/* tables out, for tests: scale[n], inv_scale[n], sigma2[n], inv_sigma2[n], quota[n], umax[16] */
void orc_get_tables(const orc_extractor *e, float *scale, float *inv_scale, float *sigma2, float *inv_sigma2,
                    int *quota, int *umax) {
    for (int i = 0; i < e->nlevels; i++) {
        scale[i] = e->scale[i]; inv_scale[i] = e->inv_scale[i];
        sigma2[i] = e->sigma2[i]; inv_sigma2[i] = e->inv_sigma2[i];
        quota[i] = e->quota[i];
    }
    for (int i = 0; i <= HALF_PATCH_SIZE; i++) umax[i] = e->umax[i];
}

/* ORBextractor::ComputePyramid (src/ORBextractor.cpp:1229-1265): chained bilinear, sizes from level 0. */
static void compute_pyramid(orc_extractor *e, const uint8_t *image, const uint8_t *mask, int w, int h, int stride,
                            int mstride) {
    free_pyramid(e);
    for (int level = 0; level < e->nlevels; level++) {
        float s = e->inv_scale[level];
        int lw = cv_round_f((float)w * s), lh = cv_round_f((float)h * s);
        if (level == 0) { lw = w; lh = h; }
        e->w[level] = lw;
        e->h[level] = lh;
        e->img[level] = (uint8_t *)malloc((size_t)(lw > 0 ? lw : 1) * (lh > 0 ? lh : 1));
        if (mask) e->mask[level] = (uint8_t *)malloc((size_t)(lw > 0 ? lw : 1) * (lh > 0 ? lh : 1));
        if (level == 0) {
            for (int y = 0; y < h; y++) {
                memcpy(e->img[0] + (size_t)y * w, image + (size_t)y * stride, (size_t)w);
                if (mask) memcpy(e->mask[0] + (size_t)y * w, mask + (size_t)y * mstride, (size_t)w);
            }
        } else {
            orc_resize_linear_u8(e->img[level - 1], e->w[level - 1], e->h[level - 1], e->w[level - 1], e->img[level],
                                 lw, lh, lw);
            if (mask)
                orc_resize_linear_u8(e->mask[level - 1], e->w[level - 1], e->h[level - 1], e->w[level - 1],
                                     e->mask[level], lw, lh, lw);
        }
    }
}

int orc_get_level(const orc_extractor *e, int level, int which, uint8_t *out) {
    const uint8_t *p = which ? e->mask[level] : e->img[level];
    if (!p) return -1;
    memcpy(out, p, (size_t)e->w[level] * e->h[level]);
    return 0;
}
void orc_get_level_size(const orc_extractor *e, int level, int *w, int *h) { *w = e->w[level]; *h = e->h[level]; }

/* IC_Angle (src/ORBextractor.cpp:27-55) */
static float ic_angle(const uint8_t *img, int stride, float ptx, float pty, const int *umax) {
    int m_01 = 0, m_10 = 0;
    const uint8_t *center = img + (ptrdiff_t)cv_round_f(pty) * stride + cv_round_f(ptx);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int val_plus = center[u + v * stride], val_minus = center[u - v * stride];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return orc_fast_atan2((float)m_01, (float)m_10);
}

/* computeOrbDescriptor (src/ORBextractor.cpp:58-98).  cos/sin: the reference's `cos(float)`
 * resolves to glibc cosf (<= 1 ulp, not guaranteed correctly rounded); the oracle and the GPU
 * both use the correctly-rounded float of the double cosine. */
static void orb_descriptor(float angle_deg, float ptx, float pty, const uint8_t *img, int stride, uint8_t *desc) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float angle = angle_deg * factorPI;
    float a = (float)cos((double)angle), b = (float)sin((double)angle);
    const uint8_t *center = img + (ptrdiff_t)cv_round_f(pty) * stride + cv_round_f(ptx);
    const int8_t *pat = k_pattern;
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int j = 0; j < 8; j++) {
            const int8_t *q = pat + 4 * j;
            float x0 = q[0], y0 = q[1], x1 = q[2], y1 = q[3];
            int t0 = center[cv_round_f(x0 * b + y0 * a) * stride + cv_round_f(x0 * a - y0 * b)];
            int t1 = center[cv_round_f(x1 * b + y1 * a) * stride + cv_round_f(x1 * a - y1 * b)];
            val |= (t0 < t1) << j;
        }
        desc[i] = (uint8_t)val;
    }
}

/* ------------------------------------------------------------------------------------------
 * ORBextractor::DistributeOctTree + ExtractorNode::DivideNode (src/ORBextractor.cpp:526-810).
 * The std::list is an index-linked list over a node pool; pool index == creation order.
 * Tie-break of the (size, pointer) sort at :731 is by heap address in the reference (undefined
 * across runs); here: equal sizes -> later-created node is expanded first.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    int ulx, uly, brx, bry; /* UL and BR; UR = (brx, uly), BL = (ulx, bry) */
    int *keys;              /* indices into the candidate array, input order preserved */
    int nkeys;
    int no_more;
    int prev, next; /* list links, -1 = none */
    uint64_t tie;   /* what stands in for the node's heap address in the (size, pointer) sort at :731 */
} qnode;

typedef struct {
    qnode *pool;
    int npool, cap;
    int head, tail, size;
} qlist;

/* Test hook: the reference orders equal-size nodes by HEAP ADDRESS (:731).  By default the k-th list node
 * created gets tie key k ("later-created first").  orc_debug_set_tie_keys(keys, n) replaces that by the
 * caller's keys — e.g. the addresses a real run of the reference (oracle/_ref, glibc malloc) handed out, in
 * allocation order — so that tests can show the restatement equals the reference under ANY heap order.
 * The counter runs on across the per-level calls of one extraction; re-arm before every extraction. */
static __thread const uint64_t *g_tie_keys = NULL;
static __thread int g_tie_n = 0;
static __thread int64_t g_tie_pos = 0;
void orc_debug_set_tie_keys(const uint64_t *keys, int n) { g_tie_keys = keys; g_tie_n = keys ? n : 0; g_tie_pos = 0; }
int orc_debug_tie_keys_used(void) { return (int)g_tie_pos; }
static uint64_t next_tie_key(void) {
    const int64_t k = g_tie_pos++;
    if (g_tie_keys) return k < g_tie_n ? g_tie_keys[k] : ~(uint64_t)0;
    return (uint64_t)k;
}

static int ql_new(qlist *L) {
    if (L->npool == L->cap) {
        L->cap = L->cap ? L->cap * 2 : 256;
        L->pool = (qnode *)realloc(L->pool, sizeof(qnode) * (size_t)L->cap);
    }
    memset(&L->pool[L->npool], 0, sizeof(qnode));
    L->pool[L->npool].prev = L->pool[L->npool].next = -1;
    return L->npool++;
}
static void ql_push_back(qlist *L, int id) { /* std::list::push_back: one heap node */
    L->pool[id].tie = next_tie_key();
    L->pool[id].prev = L->tail; L->pool[id].next = -1;
    if (L->tail >= 0) L->pool[L->tail].next = id; else L->head = id;
    L->tail = id; L->size++;
}
static void ql_push_front(qlist *L, int id) { /* std::list::push_front: one heap node */
    L->pool[id].tie = next_tie_key();
    L->pool[id].next = L->head; L->pool[id].prev = -1;
    if (L->head >= 0) L->pool[L->head].prev = id; else L->tail = id;
    L->head = id; L->size++;
}
static int ql_erase(qlist *L, int id) { /* returns next */
    int p = L->pool[id].prev, n = L->pool[id].next;
    if (p >= 0) L->pool[p].next = n; else L->head = n;
    if (n >= 0) L->pool[n].prev = p; else L->tail = p;
    L->size--;
    return n;
}

/* DivideNode :526-582.  Children are created in the pool (ids returned), not yet listed. */
static void divide_node(qlist *L, int id, const float *kx, const float *ky, int child[4]) {
    for (int c = 0; c < 4; c++) child[c] = ql_new(L); /* may realloc: take pointers after */
    qnode *n = &L->pool[id];
    const int halfX = (int)ceil((float)(n->brx - n->ulx) / 2);
    const int halfY = (int)ceil((float)(n->bry - n->uly) / 2);
    qnode *n1 = &L->pool[child[0]], *n2 = &L->pool[child[1]], *n3 = &L->pool[child[2]], *n4 = &L->pool[child[3]];
    n1->ulx = n->ulx; n1->uly = n->uly; n1->brx = n->ulx + halfX; n1->bry = n->uly + halfY;
    n2->ulx = n1->brx; n2->uly = n->uly; n2->brx = n->brx; n2->bry = n->uly + halfY;
    n3->ulx = n->ulx; n3->uly = n1->bry; n3->brx = n1->brx; n3->bry = n->bry;
    n4->ulx = n3->brx; n4->uly = n2->bry; n4->brx = n->brx; n4->bry = n->bry;
    for (int c = 0; c < 4; c++) L->pool[child[c]].keys = (int *)malloc(sizeof(int) * (size_t)(n->nkeys ? n->nkeys : 1));
    for (int i = 0; i < n->nkeys; i++) {
        int k = n->keys[i];
        qnode *t;
        if (kx[k] < n1->brx) t = (ky[k] < n1->bry) ? n1 : n3;
        else t = (ky[k] < n1->bry) ? n2 : n4;
        t->keys[t->nkeys++] = k;
    }
    for (int c = 0; c < 4; c++)
        if (L->pool[child[c]].nkeys == 1) L->pool[child[c]].no_more = 1;
}

typedef struct { int size, id; uint64_t tie; } size_id;
static int cmp_size_id(const void *a, const void *b) {
    const size_id *x = (const size_id *)a, *y = (const size_id *)b;
    if (x->size != y->size) return x->size < y->size ? -1 : 1;
    return x->tie < y->tie ? -1 : (x->tie > y->tie ? 1 : 0);
}

/* kx, ky, kr: candidate coordinates (border-relative) and responses, n of them.
 * out_idx: indices of the retained candidates in list order; returns their count. */
static int distribute_octtree(const float *kx, const float *ky, const float *kr, int n, int minX, int maxX, int minY,
                              int maxY, int N, int *out_idx) {
    /* :590-592 */
    const int nIni = (int)roundf((float)(maxX - minX) / (maxY - minY));
    if (nIni < 1) return 0; /* the reference would divide by 0 / index an empty vector */
    if (n == 0) { /* the reference still allocates its nIni root nodes (:599-610) and erases them (:628-629) */
        for (int i = 0; i < nIni; i++) (void)next_tie_key();
        return 0;
    }
    const float hX = (float)(maxX - minX) / nIni;
    qlist L; memset(&L, 0, sizeof(L)); L.head = L.tail = -1;
    int *ini = (int *)malloc(sizeof(int) * (size_t)nIni);
    for (int i = 0; i < nIni; i++) { /* :599-610 */
        int id = ql_new(&L);
        qnode *q = &L.pool[id];
        q->ulx = (int)(hX * (float)i); q->uly = 0;
        q->brx = (int)(hX * (float)(i + 1)); q->bry = maxY - minY;
        q->keys = (int *)malloc(sizeof(int) * (size_t)n);
        ql_push_back(&L, id);
        ini[i] = id;
    }
    for (int i = 0; i < n; i++) { /* :613-617 */
        int r = (int)(kx[i] / hX);
        if (r >= nIni) r = nIni - 1; /* cannot happen for in-range points; guards the oracle only */
        qnode *q = &L.pool[ini[r]];
        q->keys[q->nkeys++] = i;
    }
    for (int it = L.head; it >= 0;) { /* :619-632 */
        qnode *q = &L.pool[it];
        if (q->nkeys == 1) { q->no_more = 1; it = q->next; }
        else if (q->nkeys == 0) it = ql_erase(&L, it);
        else it = q->next;
    }
    int finish = 0;
    size_id *vsz = NULL; int nvsz = 0, capvsz = 0;
    size_id *vprev = NULL;
    while (!finish) { /* :641-781 */
        int prevSize = L.size;
        int nToExpand = 0;
        nvsz = 0;
        for (int it = L.head; it >= 0;) {
            if (L.pool[it].no_more) { it = L.pool[it].next; continue; }
            int child[4];
            divide_node(&L, it, kx, ky, child);
            for (int c = 0; c < 4; c++) {
                if (L.pool[child[c]].nkeys > 0) {
                    ql_push_front(&L, child[c]);
                    if (L.pool[child[c]].nkeys > 1) {
                        nToExpand++;
                        if (nvsz == capvsz) { capvsz = capvsz ? capvsz * 2 : 256; vsz = (size_id *)realloc(vsz, sizeof(size_id) * (size_t)capvsz); }
                        vsz[nvsz].size = L.pool[child[c]].nkeys; vsz[nvsz].id = child[c]; vsz[nvsz].tie = L.pool[child[c]].tie; nvsz++;
                    }
                }
            }
            it = ql_erase(&L, it);
        }
        if (L.size >= N || L.size == prevSize) {
            finish = 1;
        } else if (L.size + nToExpand * 3 > N) { /* :719-780 */
            while (!finish) {
                prevSize = L.size;
                int nprev = nvsz;
                vprev = (size_id *)realloc(vprev, sizeof(size_id) * (size_t)(nprev ? nprev : 1));
                memcpy(vprev, vsz, sizeof(size_id) * (size_t)nprev);
                nvsz = 0;
                qsort(vprev, (size_t)nprev, sizeof(size_id), cmp_size_id);
                for (int j = nprev - 1; j >= 0; j--) {
                    int child[4];
                    divide_node(&L, vprev[j].id, kx, ky, child);
                    for (int c = 0; c < 4; c++) {
                        if (L.pool[child[c]].nkeys > 0) {
                            ql_push_front(&L, child[c]);
                            if (L.pool[child[c]].nkeys > 1) {
                                if (nvsz == capvsz) { capvsz = capvsz ? capvsz * 2 : 256; vsz = (size_id *)realloc(vsz, sizeof(size_id) * (size_t)capvsz); }
                                vsz[nvsz].size = L.pool[child[c]].nkeys; vsz[nvsz].id = child[c]; vsz[nvsz].tie = L.pool[child[c]].tie; nvsz++;
                            }
                        }
                    }
                    ql_erase(&L, vprev[j].id);
                    if (L.size >= N) break;
                }
                if (L.size >= N || L.size == prevSize) finish = 1;
            }
        }
    }
    /* :783-807 — best response per node, first wins ties, list order */
    int nout = 0;
    for (int it = L.head; it >= 0; it = L.pool[it].next) {
        qnode *q = &L.pool[it];
        int best = q->keys[0];
        float maxResponse = kr[best];
        for (int k = 1; k < q->nkeys; k++)
            if (kr[q->keys[k]] > maxResponse) { best = q->keys[k]; maxResponse = kr[best]; }
        out_idx[nout++] = best;
    }
    for (int i = 0; i < L.npool; i++) free(L.pool[i].keys);
    free(L.pool); free(ini); free(vsz); free(vprev);
    return nout;
}

/* exported for direct quadtree tests */
int orc_distribute_octtree(const float *kx, const float *ky, const float *kr, int n, int minX, int maxX, int minY,
                           int maxY, int N, int *out_idx) {
    return distribute_octtree(kx, ky, kr, n, minX, maxX, minY, maxY, N, out_idx);
}

/* ------------------------------------------------------------------------------------------
 * Grid FAST over one image + mask test + quadtree: the body shared by
 * ComputeKeyPointsOctTree (src/ORBextractor.cpp:814-903, per level) and Detect (:998-1073).
 * Returns keypoints with border added; octave/size untouched (FAST's 0 / 7).
 * ------------------------------------------------------------------------------------------ */
typedef struct { float *x, *y, *r; int n, cap; } cand_list;
static void cand_push(cand_list *c, float x, float y, float r) {
    if (c->n == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 4096;
        c->x = (float *)realloc(c->x, sizeof(float) * (size_t)c->cap);
        c->y = (float *)realloc(c->y, sizeof(float) * (size_t)c->cap);
        c->r = (float *)realloc(c->r, sizeof(float) * (size_t)c->cap);
    }
    c->x[c->n] = x; c->y[c->n] = y; c->r[c->n] = r; c->n++;
}

/* test hook (not thread-safe): when armed, grid_fast_distribute copies its candidate list here */
static float *g_dbg_xyr = NULL;
static int g_dbg_cap = 0, g_dbg_n = 0;
void orc_debug_arm_candidates(float *xyr, int cap) { g_dbg_xyr = xyr; g_dbg_cap = xyr ? cap : 0; g_dbg_n = 0; }
int orc_debug_candidate_count(void) { return g_dbg_n; }

static int grid_fast_distribute(const orc_extractor *e, const uint8_t *img, const uint8_t *mask, int cols, int rows,
                                int stride, int mstride, int N, orc_keypoint *out, int cap, int *n_candidates) {
    const float W = 30;
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (n_candidates) *n_candidates = 0;
    if (nCols <= 0 || nRows <= 0) return 0; /* reference: no cell is visited -> no keypoints */
    const int wCell = (int)ceil(width / nCols), hCell = (int)ceil(height / nRows);
    cand_list cl; memset(&cl, 0, sizeof(cl));
    int fcap = (wCell + 6) * (hCell + 6);
    int *fbuf = (int *)malloc(sizeof(int) * 3 * (size_t)fcap);
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            const uint8_t *roi = img + (size_t)y0 * stride + x0;
            int nk = orc_fast9_16(roi, x1 - x0, y1 - y0, stride, e->iniThFAST, 1, fbuf, fcap);
            if (nk == 0) nk = orc_fast9_16(roi, x1 - x0, y1 - y0, stride, e->minThFAST, 1, fbuf, fcap);
            for (int k = 0; k < nk; k++) {
                float px = (float)fbuf[3 * k] + j * wCell, py = (float)fbuf[3 * k + 1] + i * hCell;
                /* mask looked up at border-relative coordinates (quirk Q1, :871-877) */
                if (mask && mask[(size_t)cv_round_f(py) * mstride + cv_round_f(px)] == 0) continue;
                cand_push(&cl, px, py, (float)fbuf[3 * k + 2]);
            }
        }
    }
    free(fbuf);
    if (n_candidates) *n_candidates = cl.n;
    if (g_dbg_cap > 0) { /* test hook: expose the candidate list in the reference's order */
        g_dbg_n = cl.n;
        for (int k = 0; k < cl.n && k < g_dbg_cap; k++) {
            g_dbg_xyr[3 * k] = cl.x[k]; g_dbg_xyr[3 * k + 1] = cl.y[k]; g_dbg_xyr[3 * k + 2] = cl.r[k];
        }
    }
    int *sel = (int *)malloc(sizeof(int) * (size_t)(cl.n ? cl.n : 1));
    int ns = distribute_octtree(cl.x, cl.y, cl.r, cl.n, minBorderX, maxBorderX, minBorderY, maxBorderY, N, sel);
    int nout = 0;
    for (int k = 0; k < ns; k++) {
        if (nout < cap) {
            orc_keypoint *kp = &out[nout];
            kp->x = cl.x[sel[k]] + minBorderX;
            kp->y = cl.y[sel[k]] + minBorderY;
            kp->size = 7.f; kp->angle = -1.f; kp->response = cl.r[sel[k]];
            kp->octave = 0; kp->class_id = -1;
        }
        nout++;
    }
    free(sel); free(cl.x); free(cl.y); free(cl.r);
    return nout;
}

/* ComputeKeyPointsOctTree :814-907; per-level results appended level-major into out;
 * level_counts[nlevels] receives the per-level counts; pt in LEVEL coordinates. */
static int compute_keypoints_octtree(orc_extractor *e, int use_mask, orc_keypoint *out, int cap, int *level_counts,
                                     int *level_cands) {
    int total = 0;
    for (int level = 0; level < e->nlevels; level++) {
        int ncand = 0;
        int room = cap - total > 0 ? cap - total : 0;
        int n = grid_fast_distribute(e, e->img[level], use_mask ? e->mask[level] : NULL, e->w[level], e->h[level],
                                     e->w[level], e->w[level], e->quota[level], out + total, room, &ncand);
        if (n > room) n = room;
        const int scaledPatchSize = (int)(PATCH_SIZE * e->scale[level]);
        for (int i = 0; i < n; i++) {
            out[total + i].octave = level;
            out[total + i].size = (float)scaledPatchSize;
        }
        level_counts[level] = n;
        if (level_cands) level_cands[level] = ncand;
        total += n;
    }
    /* :905-906 orientation on the un-blurred level images */
    int off = 0;
    for (int level = 0; level < e->nlevels; level++) {
        for (int i = 0; i < level_counts[level]; i++) {
            orc_keypoint *kp = &out[off + i];
            kp->angle = ic_angle(e->img[level], e->w[level], kp->x, kp->y, e->umax);
        }
        off += level_counts[level];
    }
    return total;
}

/* ORBextractor::DetectAndCompute :922-985.  mask may be NULL (== all 255; the reference
 * requires a non-empty mask).  desc: [cap][32].  Returns the number of keypoints.
 * level_cands (nullable) receives the number of quadtree candidates per level. */
int orc_detect_and_compute(orc_extractor *e, const uint8_t *image, const uint8_t *mask, int w, int h, int stride,
                           int mstride, orc_keypoint *kps, uint8_t *desc, int cap, int *level_counts_out,
                           int *level_cands) {
    if (!image || w <= 0 || h <= 0) return 0;
    compute_pyramid(e, image, mask, w, h, stride, mstride);
    int level_counts[ORC_MAX_LEVELS];
    int n = compute_keypoints_octtree(e, mask != NULL, kps, cap, level_counts, level_cands);
    int offset = 0;
    for (int level = 0; level < e->nlevels; level++) {
        int nl = level_counts[level];
        if (level_counts_out) level_counts_out[level] = nl;
        if (nl == 0) continue;
        if (desc) {
            uint8_t *work = (uint8_t *)malloc((size_t)e->w[level] * e->h[level]);
            orc_gauss7_u8(e->img[level], e->w[level], e->h[level], e->w[level], work, e->w[level]);
            for (int i = 0; i < nl; i++) {
                orc_keypoint *kp = &kps[offset + i];
                orb_descriptor(kp->angle, kp->x, kp->y, work, e->w[level], desc + (size_t)(offset + i) * 32);
            }
            free(work);
        }
        if (level != 0) {
            float scale = e->scale[level];
            for (int i = 0; i < nl; i++) { kps[offset + i].x *= scale; kps[offset + i].y *= scale; }
        }
        offset += nl;
    }
    return n;
}

/* ORBextractor::DetectWithPyramid :1135-1176 (DetectAndCompute without descriptors). */
int orc_detect_with_pyramid(orc_extractor *e, const uint8_t *image, const uint8_t *mask, int w, int h, int stride,
                            int mstride, orc_keypoint *kps, int cap) {
    return orc_detect_and_compute(e, image, mask, w, h, stride, mstride, kps, NULL, cap, NULL, NULL);
}

/* ORBextractor::Detect :989-1074: level 0 only, N = nfeatures, FAST's size/angle/octave kept (Q4). */
int orc_detect(orc_extractor *e, const uint8_t *image, const uint8_t *mask, int w, int h, int stride, int mstride,
               orc_keypoint *kps, int cap) {
    if (!image || w <= 0 || h <= 0) return 0;
    int n = grid_fast_distribute(e, image, mask, w, h, stride, mstride, e->nfeatures, kps, cap, NULL);
    return n > cap ? cap : n;
}

/* isFastCorner :449-511 — single-pixel FAST-9/16 at `threshold`. */
static int is_fast_corner(const uint8_t *img, int stride, float ptx, float pty, int threshold) {
    int pix[25];
    for (int k = 0; k < 16; k++) pix[k] = k_ring[k][0] + k_ring[k][1] * stride;
    for (int k = 16; k < 25; k++) pix[k] = pix[k - 16];
    threshold = threshold < 0 ? 0 : (threshold > 255 ? 255 : threshold);
    const uint8_t *ptr = img + (ptrdiff_t)cv_round_f(pty) * stride + cv_round_f(ptx);
    const int v = ptr[0];
    int vt = v - threshold, count = 0;
    for (int k = 0; k < 25; k++) {
        if (ptr[pix[k]] < vt) { if (++count > 8) return 1; }
        else count = 0;
    }
    vt = v + threshold; count = 0;
    for (int k = 0; k < 25; k++) {
        if (ptr[pix[k]] > vt) { if (++count > 8) return 1; }
        else count = 0;
    }
    return 0;
}

/* ORBextractor::ScreenAndComputeKPsParams :1083-1129.  `in` is mutated like the reference
 * mutates its input vector (pt /= scale; pt *= scale round trip, Q5).  Returns n_out. */
int orc_screen_params(orc_extractor *e, const uint8_t *image, int w, int h, int stride, orc_keypoint *in, int n_in,
                      orc_keypoint *out) {
    if (!image || w <= 0 || h <= 0 || n_in <= 0) return 0;
    compute_pyramid(e, image, NULL, w, h, stride, 0);
    int n_out = 0;
    for (int i = 0; i < n_in; i++) {
        orc_keypoint *kp = &in[i];
        int level = kp->octave;
        float scale = e->scale[level];
        kp->x /= scale; kp->y /= scale;
        if (!(kp->y - EDGE_THRESHOLD >= 0 && kp->y + EDGE_THRESHOLD < e->h[level] && kp->x - EDGE_THRESHOLD >= 0 &&
              kp->x + EDGE_THRESHOLD < e->w[level])) {
            kp->x *= scale; kp->y *= scale;
            continue;
        }
        if (!is_fast_corner(e->img[level], e->w[level], kp->x, kp->y, e->minThFAST)) {
            kp->x *= scale; kp->y *= scale;
            continue;
        }
        kp->angle = ic_angle(e->img[level], e->w[level], kp->x, kp->y, e->umax);
        kp->size = PATCH_SIZE * e->scale[level];
        kp->x *= scale; kp->y *= scale;
        out[n_out++] = *kp;
    }
    return n_out;
}

/* ORBextractor::CalcDescriptors :1180-1226 (row i of the output = descriptor of keypoint i, Q6). */
int orc_calc_descriptors(orc_extractor *e, const uint8_t *image, int w, int h, int stride, const orc_keypoint *kps,
                         int n, uint8_t *desc) {
    if (!image || w <= 0 || h <= 0 || n <= 0) return 0;
    compute_pyramid(e, image, NULL, w, h, stride, 0);
    uint8_t *work[ORC_MAX_LEVELS];
    for (int level = 0; level < e->nlevels; level++) {
        work[level] = (uint8_t *)malloc((size_t)e->w[level] * e->h[level]);
        orc_gauss7_u8(e->img[level], e->w[level], e->h[level], e->w[level], work[level], e->w[level]);
    }
    for (int i = 0; i < n; i++) {
        orc_keypoint kp = kps[i];
        int level = kp.octave;
        float scale = e->scale[level];
        kp.x /= scale; kp.y /= scale;
        orb_descriptor(kp.angle, kp.x, kp.y, work[level], e->w[level], desc + (size_t)i * 32);
    }
    for (int level = 0; level < e->nlevels; level++) free(work[level]);
    return n;
}

/* blurred level, for stage-by-stage GPU comparisons */
int orc_get_blurred_level(const orc_extractor *e, int level, uint8_t *out) {
    if (!e->img[level]) return -1;
    return orc_gauss7_u8(e->img[level], e->w[level], e->h[level], e->w[level], out, e->w[level]);
}

/* ------------------------------------------------------------------------------------------
 * cv::BFMatcher(NORM_HAMMING)::match(query, train) as used at src/loopclosing.cpp:33,172:
 * one match per query row = nearest train row, ties -> lowest trainIdx.
 * ------------------------------------------------------------------------------------------ */
int orc_hamming_match(const uint8_t *q, int nq, const uint8_t *t, int nt, int32_t *train_idx, int32_t *dist) {
    for (int i = 0; i < nq; i++) {
        const uint64_t *a = (const uint64_t *)(q + (size_t)i * 32);
        int best = -1, bd = 1 << 30;
        for (int j = 0; j < nt; j++) {
            const uint64_t *b = (const uint64_t *)(t + (size_t)j * 32);
            int d = __builtin_popcountll(a[0] ^ b[0]) + __builtin_popcountll(a[1] ^ b[1]) +
                    __builtin_popcountll(a[2] ^ b[2]) + __builtin_popcountll(a[3] ^ b[3]);
            if (d < bd) { bd = d; best = j; }
        }
        train_idx[i] = best;
        dist[i] = best < 0 ? -1 : bd;
    }
    return 0;
}

/* LoopClosing::MatchFeatures filter, src/loopclosing.cpp:175-194: keep distance <= max(2*min, 30),
 * map keypoint -> feature by class_id, de-duplicate (current, loop) pairs (a std::set: output is
 * sorted by (current, loop)).  Returns the number of valid pairs. */
static int cmp_pair(const void *a, const void *b) {
    const int32_t *x = (const int32_t *)a, *y = (const int32_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}
int orc_match_filter(const int32_t *train_idx, const int32_t *dist, int nq, const int32_t *loop_class_id,
                     const int32_t *cur_class_id, int32_t *pairs /*[nq][2]*/) {
    if (nq <= 0) return 0;
    double min_dist = 1e30;
    for (int i = 0; i < nq; i++)
        if (train_idx[i] >= 0 && (float)dist[i] < min_dist) min_dist = (float)dist[i];
    double th = 2 * min_dist > 30.0 ? 2 * min_dist : 30.0;
    int n = 0;
    for (int i = 0; i < nq; i++) {
        if (train_idx[i] < 0) continue;
        if ((float)dist[i] <= th) {
            pairs[2 * n] = cur_class_id[train_idx[i]];
            pairs[2 * n + 1] = loop_class_id[i];
            n++;
        }
    }
    qsort(pairs, (size_t)n, 2 * sizeof(int32_t), cmp_pair);
    int m = 0;
    for (int i = 0; i < n; i++)
        if (m == 0 || pairs[2 * i] != pairs[2 * (m - 1)] || pairs[2 * i + 1] != pairs[2 * (m - 1) + 1]) {
            pairs[2 * m] = pairs[2 * i]; pairs[2 * m + 1] = pairs[2 * i + 1]; m++;
        }
    return m;
}
