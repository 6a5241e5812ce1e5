// orb_core.inl — per-pixel / per-keypoint arithmetic of the ORB extractor, written once for the
// device kernels (orb.cu) and, with SB_HOST_MODEL defined, for a host build that the CPU test-suite
// compares with the oracle (tests/orb_host_model.cpp).  Everything here is integer arithmetic or
// explicitly un-contracted fp32 so that the results are bit-identical to the reference's
// OpenCV-based CPU path (reference src/ORBextractor.cpp; OpenCV primitives per SURVEY.md app. A).
#ifndef ORB_CORE_INL
#define ORB_CORE_INL

#include <stdint.h>

#ifdef SB_HOST_MODEL
#include <math.h>
#define SB_HD static inline
SB_HD float sb_fmul(float a, float b) { volatile float r = a * b; return r; }
SB_HD float sb_fadd(float a, float b) { volatile float r = a + b; return r; }
SB_HD float sb_fsub(float a, float b) { volatile float r = a - b; return r; }
SB_HD float sb_fdiv(float a, float b) { volatile float r = a / b; return r; }
SB_HD int sb_rint(float a) { return (int)lrintf(a); }
SB_HD int sb_min(int a, int b) { return a < b ? a : b; }
SB_HD int sb_max(int a, int b) { return a > b ? a : b; }
// host emulation of the packed 2 x int16 instructions the device build uses (VIADD.16x2, VIMNMX.S16x2, PRMT)
SB_HD uint32_t sb_pack2(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
SB_HD int sb_lo2(uint32_t a) { return (int)(int16_t)(a & 0xffffu); }
SB_HD int sb_hi2(uint32_t a) { return (int)(int16_t)(a >> 16); }
SB_HD uint32_t sb_vadd2(uint32_t a, uint32_t b) { return sb_pack2(sb_lo2(a) + sb_lo2(b), sb_hi2(a) + sb_hi2(b)); }
SB_HD uint32_t sb_vmin2(uint32_t a, uint32_t b) { return sb_pack2(sb_min(sb_lo2(a), sb_lo2(b)), sb_min(sb_hi2(a), sb_hi2(b))); }
SB_HD uint32_t sb_vmax2(uint32_t a, uint32_t b) { return sb_pack2(sb_max(sb_lo2(a), sb_lo2(b)), sb_max(sb_hi2(a), sb_hi2(b))); }
SB_HD uint32_t sb_vmin3_2(uint32_t a, uint32_t b, uint32_t c) { return sb_vmin2(sb_vmin2(a, b), c); }
SB_HD uint32_t sb_vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return sb_vmax2(sb_vmax2(a, b), c); }
#else
#define SB_HD static __device__ __forceinline__
SB_HD float sb_fmul(float a, float b) { return __fmul_rn(a, b); }
SB_HD float sb_fadd(float a, float b) { return __fadd_rn(a, b); }
SB_HD float sb_fsub(float a, float b) { return __fsub_rn(a, b); }
SB_HD float sb_fdiv(float a, float b) { return __fdiv_rn(a, b); }
SB_HD int sb_rint(float a) { return __float2int_rn(a); }  // cvRound: round half to even
SB_HD int sb_min(int a, int b) { return min(a, b); }
SB_HD int sb_max(int a, int b) { return max(a, b); }
SB_HD uint32_t sb_pack2(int lo, int hi) { return __byte_perm((uint32_t)lo, (uint32_t)hi, 0x5410); }
SB_HD int sb_lo2(uint32_t a) { return (int)(short)(a & 0xffffu); }
SB_HD int sb_hi2(uint32_t a) { return (int)a >> 16; }
SB_HD uint32_t sb_vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }                             // VIADD.16x2
SB_HD uint32_t sb_vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }                            // VIMNMX.S16x2
SB_HD uint32_t sb_vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
SB_HD uint32_t sb_vmin3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }     // VIMNMX3.S16x2
SB_HD uint32_t sb_vmax3_2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
#endif

#define SB_HALF_PATCH 15
#define SB_EDGE 19

// ---- FAST-9/16 (cv::FAST as called at src/ORBextractor.cpp:858-865) ------------------------------
// Bresenham ring of radius 3, clockwise from (0,+3) — the order the reference re-declares at :365-369.
#define SB_RING_DX(k) ((k) == 0 ? 0 : (k) == 1 ? 1 : (k) == 2 ? 2 : (k) == 3 ? 3 : (k) == 4 ? 3 : (k) == 5 ? 3 : (k) == 6 ? 2 : \
                       (k) == 7 ? 1 : (k) == 8 ? 0 : (k) == 9 ? -1 : (k) == 10 ? -2 : (k) == 11 ? -3 : (k) == 12 ? -3 :         \
                       (k) == 13 ? -3 : (k) == 14 ? -2 : -1)
#define SB_RING_DY(k) ((k) == 0 ? 3 : (k) == 1 ? 3 : (k) == 2 ? 2 : (k) == 3 ? 1 : (k) == 4 ? 0 : (k) == 5 ? -1 : (k) == 6 ? -2 : \
                       (k) == 7 ? -3 : (k) == 8 ? -3 : (k) == 9 ? -3 : (k) == 10 ? -2 : (k) == 11 ? -1 : (k) == 12 ? 0 :        \
                       (k) == 13 ? 1 : (k) == 14 ? 2 : 3)

// Necessary condition for "corner at threshold t": any 9 contiguous ring positions contain at least one of the
// opposite positions (0, 8) and at least one of (4, 12), and all of them lie on the same side.  So the brighter of
// the two pair minima must be darker than v - t, or the darker of the two pair maxima brighter than v + t
// (28 % of the synthetic pyramid's pixels pass; "two of the four compass pixels" passes 35 %).
SB_HD bool sb_fast_maybe(const uint8_t *p, int pitch, int t) {
    const int v = p[0];
    const int r0 = p[3 * pitch], r4 = p[3], r8 = p[-3 * pitch], r12 = p[-3];
    const int dk = sb_max(sb_min(r0, r8), sb_min(r4, r12)), br = sb_min(sb_max(r0, r8), sb_max(r4, r12));
    return dk < v - t || br > v + t;
}

// Corner response = the largest threshold for which the pixel is still a FAST-9/16 corner:
//   max over the 16 arcs of 9 contiguous ring pixels of  min(v - ring)  resp.  min(ring - v),  minus 1.
// Both polarities ride in one register as 2 x int16: high half ring - v, low half v - ring + 256 (the bias
// keeps the low half positive, so ONE 32-bit multiply-add  ring * 65535 + c(v)  builds the pair without a
// borrow into the high half).  A window of 9 is min3 of three windows of 3 (VIMNMX3.S16x2): 16 + 16
// packed instructions for all 16 arcs of both polarities.  No value is ever negated after a min/max
// (ptxas 12.9 mis-folds max(a, -max3(...)) into VIMNMX3 on sm_100a — measured, tools/fast_probe.cu).
SB_HD int sb_fast_score(const uint8_t *p, int pitch) {
    const uint32_t v = p[0];
    uint32_t cv = (0u - (v << 16)) + v + 256u;  // ((-v) << 16) + (v + 256)
#ifdef __CUDA_ARCH__
    asm volatile("" : "+r"(cv));  // opaque: otherwise nvcc rewrites r * 65535 + cv as (r - v) * 65535 + 256, two instructions per ring pixel
#endif
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t r = p[SB_RING_DY(k) * pitch + SB_RING_DX(k)];
        d[k] = r * 65535u + cv;  // (r << 16) - r + cv = ((r - v) << 16) + (v - r + 256), low half in [1, 511]
    }
    uint32_t m3[16], m9[16];
#pragma unroll
    for (int k = 0; k < 16; k++) m3[k] = sb_vmin3_2(d[k], d[(k + 1) & 15], d[(k + 2) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) m9[k] = sb_vmin3_2(m3[k], m3[(k + 3) & 15], m3[(k + 6) & 15]);
    uint32_t b0 = sb_vmax3_2(m9[0], m9[1], m9[2]), b1 = sb_vmax3_2(m9[3], m9[4], m9[5]);
    uint32_t b2 = sb_vmax3_2(m9[6], m9[7], m9[8]), b3 = sb_vmax3_2(m9[9], m9[10], m9[11]);
    uint32_t b4 = sb_vmax3_2(m9[12], m9[13], m9[14]);
    b0 = sb_vmax3_2(b0, b1, b2);
    b3 = sb_vmax3_2(b3, b4, m9[15]);
    b0 = sb_vmax2(b0, b3);
    return sb_max(sb_lo2(b0) - 256, sb_hi2(b0)) - 1;
}

// ---- cv::resize INTER_LINEAR u8 (src/ORBextractor.cpp:1243-1244,1262) ----------------
// Source index and the two 11-bit coefficients of destination index d along one axis.
// HOST ONLY (tables are built once per image size): the double expression must not be contracted
// into an FMA, which nvcc would do in device code.
struct SbLinCoef { int s; short c0, c1; };
static inline SbLinCoef sb_lin_coef(int d, int dst_len, int src_len, bool clamp_like_x) {
    const double scale = 1.0 / ((double)dst_len / (double)src_len);
    volatile double prod = ((double)d + 0.5) * scale;
    float f = (float)(prod - 0.5);
    int s = (int)floorf(f);
    f = f - (float)s;
    if (clamp_like_x) {  // x axis: coefficient collapses at both borders
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= src_len - 1) { f = 0.f; s = src_len - 1; }
    }
    SbLinCoef c;
    c.s = s;
    c.c0 = (short)lrintf((1.f - f) * 2048.f);
    c.c1 = (short)lrintf(f * 2048.f);
    return c;
}
// vertical combine of two horizontally interpolated rows
SB_HD uint8_t sb_lin_vert(int h0, int h1, int b0, int b1) {
    return (uint8_t)((((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2);
}

// ---- cv::GaussianBlur 7x7 sigma 2, u8 fixed point (src/ORBextractor.cpp:965-966) -----
#define SB_G0 18
#define SB_G1 34
#define SB_G2 48
#define SB_G3 56
SB_HD int sb_reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}
SB_HD unsigned sb_gauss_row(unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g) {
    return SB_G0 * (a + g) + SB_G1 * (b + f) + SB_G2 * (c + e) + SB_G3 * d;  // 8.8, <= 255 * 256
}
SB_HD uint8_t sb_gauss_col(unsigned a, unsigned b, unsigned c, unsigned d, unsigned e, unsigned f, unsigned g) {
    return (uint8_t)((SB_G0 * (a + g) + SB_G1 * (b + f) + SB_G2 * (c + e) + SB_G3 * d + 32768u) >> 16);
}

// ---- cv::fastAtan2 scalar fp32 (src/ORBextractor.cpp:54) -------------------------------
SB_HD float sb_fast_atan2(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale;
    const float p5 = 0.1555786518463281f * scale, p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    const float eps = 2.220446049250313e-16f;  // (float)DBL_EPSILON
    float a, c, c2;
    if (ax >= ay) {
        c = sb_fdiv(ay, sb_fadd(ax, eps));
        c2 = sb_fmul(c, c);
        a = sb_fmul(sb_fadd(sb_fmul(sb_fadd(sb_fmul(sb_fadd(sb_fmul(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = sb_fdiv(ax, sb_fadd(ay, eps));
        c2 = sb_fmul(c, c);
        a = sb_fsub(90.f, sb_fmul(sb_fadd(sb_fmul(sb_fadd(sb_fmul(sb_fadd(sb_fmul(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = sb_fsub(180.f, a);
    if (y < 0) a = sb_fsub(360.f, a);
    return a;
}

// ---- one rBRIEF test (computeOrbDescriptor, src/ORBextractor.cpp:59-98) ---------------
// (x, y) pattern point, a = cos, b = sin; returns the sampled pixel.
SB_HD int sb_brief_sample(const uint8_t *center, int pitch, float a, float b, float x, float y) {
    const int dy = sb_rint(sb_fadd(sb_fmul(x, b), sb_fmul(y, a)));
    const int dx = sb_rint(sb_fsub(sb_fmul(x, a), sb_fmul(y, b)));
    return center[dy * pitch + dx];
}

#endif  // ORB_CORE_INL
