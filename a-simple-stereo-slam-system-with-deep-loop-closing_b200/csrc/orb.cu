// orb.cu — B200 (sm_100a) ORB extractor behind the C ABI of include/slamb200.h.
//
// Replaces the arithmetic of myslam::ORBextractor (reference src/ORBextractor.cpp:384-1265) for a
// BATCH of images per call.  Stage map (reference lines -> kernel):
//   ComputePyramid            :1229-1265 -> k_copy_level0 + k_resize (one launch per level, chained)
//   cv::FAST per 30-px cell   :838-883   -> k_fast_cells   (one CTA per cell; tile staged by TMA)
//   DistributeOctTree         :586-810   -> k_quadtree     (one CTA per (image, level); qt_core.inl)
//   GaussianBlur per level    :965-966   -> k_blur         (one launch for all levels; TMA halo tiles)
// TMA note (measured on B200, tools/tma_probe.cu): the innermost box coordinate must be a multiple
// of 16 bytes — negative / out-of-range coordinates are fine and zero filled — so tiles start at the
// 16-byte boundary at or left of the wanted column.
//   IC_Angle + rBRIEF + output:27-98,:975-983 -> k_describe (one warp per keypoint)
// Data layout in HBM: every image of the batch owns one "slab" holding its pyramid levels as
// pitched u8 planes (pitch multiple of 128 B so that each level is a legal TMA tensor
// [w, h, batch] with strides [pitch, slab]); a second slab array holds the blurred levels and an
// optional third one the mask pyramid.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "orb_core.inl"
#include "qt_core.inl"

#define SB_MAX_LEVELS 12
#define SB_CAND_CAP 16384       // FAST candidates one (image, level) can hand to the quadtree
#define SB_FAST_KMAX 8          // FAST cells one CTA handles (a horizontal run of one grid row)
#define SB_FAST_SPAN 224        // ... as many as fit this many tested columns
#define SB_FAST_BW 256          // TMA box width of every FAST tile (compile time: ring offsets fold into the LDS immediates)
#define FAST_THREADS 256
#define QT_THREADS 256
#define BLUR_TW 128
#define BLUR_TH 32
#define BLUR_HX 16   // left halo: the innermost TMA coordinate must be a multiple of 16 bytes (measured: tools/tma_probe.cu)
#define BLUR_BW 160  // BLUR_HX + BLUR_TW + 3, rounded up to the 16-byte TMA granule
#define BLUR_BH 38
#define DESC_WARPS 8
#define SB_PROF_MAX 8192

__device__ __align__(16) int8_t d_pattern[1024] = {
#include "orb_pattern.inc"
};

enum { SB_STAGE_COPY = 0, SB_STAGE_RESIZE, SB_STAGE_FAST, SB_STAGE_QUADTREE, SB_STAGE_BLUR, SB_STAGE_DESCRIBE, SB_STAGE_COUNT };

struct LevelGeom {
    int w, h, pitch;
    int fast_bw, fast_bh;      // TMA box of one FAST cell tile
    long long off;             // byte offset of the level inside an image slab
    int bw, bh;                // maxBorder - minBorder (the quadtree's box)
    int nCols, nRows, wCell, hCell;
    int quota;                 // mnFeaturesPerLevel[level]
    int xtab, ytab;            // offsets into the resize tables
    int qtab;                  // offset into the quad table, -1: this level needs the generic k_resize
    float scale, patch;        // mvScaleFactor[level], (float)(int)(31 * scale)
};

struct Geom {
    LevelGeom lv[SB_MAX_LEVELS];
    int nlevels;
    long long slab;  // bytes per image
    int umax[16];
};

struct CellGroup {  // up to SB_FAST_KMAX horizontally adjacent cv::FAST calls of the reference's grid loop (one CTA)
    short level, x0, y0;   // first cell's ROI origin (iniX, iniY) in level pixels
    short rw, rh;          // ROI span of the whole group: last cell's maxX - first cell's iniX, maxY - iniY
    short offx, offy;      // j0 * wCell, i * hCell: what the reference adds to the ROI coordinates
    short ncells, wCell;
    short G;               // 4-pixel items per row
    int rcpG, rcpW;        // ceil(2^20 / G), ceil(2^20 / wCell): exact division of the small indices used
};

struct TmaMaps {
    CUtensorMap m[SB_MAX_LEVELS];
};

// ================================================================================================
// kernels
// ================================================================================================

// Level 0 = clone of the input (ComputePyramid :1240-1241, :1259-1260), re-pitched.  A thread moves 16 bytes: the source
// row starts at an arbitrary byte (KITTI rows are 1241 bytes), so it reads the five aligned words that cover its
// 16 bytes and funnel-shifts them into place; bytes past the row end are zeroed (they land in the pitch padding).
__global__ void __launch_bounds__(256) k_copy_level0(const uint8_t *__restrict__ src, long long src_img_pitch, int stride, int w, int h,
                                                    uint8_t *__restrict__ dst, long long slab, int pitch, int batch) {
    sb_pdl_enter();
    const int p16 = pitch >> 4;
    const long long total = (long long)batch * h * p16;
    const uint8_t *src_end = src + (long long)(batch - 1) * src_img_pitch + (long long)(h - 1) * stride + w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x16 = (int)(i % p16);
        const long long r = i / p16;
        const int y = (int)(r % h), b = (int)(r / h);
        const int x = x16 * 16;
        uint4 out = make_uint4(0u, 0u, 0u, 0u);
        if (x < w) {
            const uint8_t *s = src + b * src_img_pitch + (long long)y * stride + x;
            const uint32_t sh = (uint32_t)((uintptr_t)s & 3u) * 8u;
            const uint32_t *sw = reinterpret_cast<const uint32_t *>(s - ((uintptr_t)s & 3u));
            // the image ends at src + (batch - 1) * pitch + (h - 1) * stride + w: do not read words wholly past this row's end
            const int nbytes = min(16, w - x);                       // valid bytes of this chunk
            const int nwords = (int)((sh >> 3) + nbytes + 3) >> 2;   // aligned words that hold them
            uint32_t v[5];
#pragma unroll
            for (int k = 0; k < 5; k++) {
                v[k] = 0u;
                if (k < nwords) {
                    const uint8_t *a = reinterpret_cast<const uint8_t *>(sw + k);
                    if (a >= src && a + 4 <= src_end) {
                        v[k] = __ldg(sw + k);
                    } else {  // first / last word of the whole buffer: only the bytes that exist
                        for (int q = 0; q < 4; q++)
                            if (a + q >= src && a + q < src_end) v[k] |= (uint32_t)a[q] << (8 * q);
                    }
                }
            }
            out.x = __funnelshift_r(v[0], v[1], sh);
            out.y = __funnelshift_r(v[1], v[2], sh);
            out.z = __funnelshift_r(v[2], v[3], sh);
            out.w = __funnelshift_r(v[3], v[4], sh);
            if (nbytes < 16) {  // row tail: zero the bytes that belong to the next source row
                uint32_t *o = &out.x;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int keep = nbytes - 4 * k;  // bytes of word k that are inside the row
                    if (keep <= 0) o[k] = 0u;
                    else if (keep < 4) o[k] &= (1u << (8 * keep)) - 1u;
                }
            }
        }
        *reinterpret_cast<uint4 *>(dst + b * slab + (long long)y * pitch + x) = out;
    }
}

// A "null mask" level 0: all 255.
__global__ void k_fill_level0(uint8_t *dst, long long slab, int pitch, int h, int batch) {
    const int p4 = pitch >> 2;
    const long long total = (long long)batch * h * p4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x4 = (int)(i % p4);
        const long long r = i / p4;
        const int y = (int)(r % h), b = (int)(r / h);
        *reinterpret_cast<uint32_t *>(dst + b * slab + (long long)y * pitch + x4 * 4) = 0xffffffffu;
    }
}

// cv::resize(level-1 -> level, INTER_LINEAR), fixed point.  4 destination pixels per thread.
// xtab[dx] = (source column, c0 | c1 << 16).  The two source rows are read as aligned 32-bit words; a
// funnel shift brings (S[sx], S[sx+1]) into the low bytes and one DP2A forms S[sx] * c0 + S[sx+1] * c1.
__global__ void __launch_bounds__(256) k_resize(uint8_t *__restrict__ pyr, long long slab, LevelGeom src, LevelGeom dst,
                                               const int2 *__restrict__ xtab, const int2 *__restrict__ ytab) {
    sb_pdl_enter();
    const int x0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int y = blockIdx.y;
    if (x0 >= dst.w) return;
    uint8_t *base = pyr + (long long)blockIdx.z * slab;
    const int2 yt = ytab[dst.ytab + y];
    const int sy = yt.x, b0 = yt.y & 0xffff, b1 = (unsigned)yt.y >> 16;
    const int sy0 = min(max(sy, 0), src.h - 1), sy1 = min(max(sy + 1, 0), src.h - 1);
    const uint8_t *S0 = base + src.off + (long long)sy0 * src.pitch;
    const uint8_t *S1 = base + src.off + (long long)sy1 * src.pitch;
    int2 xt[4];
#pragma unroll
    for (int k = 0; k < 4; k++) xt[k] = xtab[dst.xtab + min(x0 + k, dst.w - 1)];
    const int wb = xt[0].x >> 2;  // first source word
    uint32_t out = 0;
    if (xt[3].x + 1 - 4 * wb < 12) {  // the 4 pixels read at most 12 consecutive source bytes (any scale factor <= 2)
        const uint32_t *W0 = reinterpret_cast<const uint32_t *>(S0) + wb, *W1 = reinterpret_cast<const uint32_t *>(S1) + wb;
        const uint32_t a0 = W0[0], a1 = W0[1], a2 = W0[2], c0 = W1[0], c1 = W1[1], c2 = W1[2];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int o = xt[k].x - 4 * wb;  // 0 .. 10
            const uint32_t p0 = o < 4 ? __funnelshift_r(a0, a1, 8 * o) : o < 8 ? __funnelshift_r(a1, a2, 8 * (o - 4)) : a2 >> (8 * (o - 8));
            const uint32_t p1 = o < 4 ? __funnelshift_r(c0, c1, 8 * o) : o < 8 ? __funnelshift_r(c1, c2, 8 * (o - 4)) : c2 >> (8 * (o - 8));
            const int h0 = (int)__dp2a_lo((unsigned)xt[k].y, p0, 0u);
            const int h1 = (int)__dp2a_lo((unsigned)xt[k].y, p1, 0u);
            out |= (uint32_t)sb_lin_vert(h0, h1, b0, b1) << (8 * k);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int sx = xt[k].x, sx1 = min(sx + 1, src.w - 1);
            const int ca = xt[k].y & 0xffff, cb = (unsigned)xt[k].y >> 16;
            out |= (uint32_t)sb_lin_vert(S0[sx] * ca + S0[sx1] * cb, S1[sx] * ca + S1[sx1] * cb, b0, b1) << (8 * k);
        }
    }
    *reinterpret_cast<uint32_t *>(base + dst.off + (long long)y * dst.pitch + x0) = out;
}

// The same resize for scale factors <= 2: the 4 destination pixels of a thread read at most 8 consecutive source
// bytes starting at S[sx of the first pixel]; three aligned words per source row and two funnel shifts form that
// window, and one PRMT with a precomputed selector per pixel brings (S[sx], S[sx + 1]) into place for the DP2A.
// A thread keeps its column table in registers and walks RESIZE_ROWS rows.
struct __align__(16) ResizeQuad {
    int wb;           // first source word of the quad
    uint32_t sels;    // 4 x 8-bit PRMT selectors: bytes (o, o + 1) of the 8-byte window
    uint32_t c[4];    // c0 | c1 << 16 per pixel
    int sh;           // 8 * (first source byte & 3): the window starts this many bits into word wb
    int pad;
};
#define RESIZE_ROWS 4
__global__ void __launch_bounds__(256) k_resize_quads(uint8_t *__restrict__ pyr, long long slab, LevelGeom src, LevelGeom dst,
                                                     const ResizeQuad *__restrict__ qtab, const int2 *__restrict__ ytab, int nquads) {
    sb_pdl_enter();
    const int q = blockIdx.x * 32 + threadIdx.x;
    if (q >= nquads) return;
    const int4 t0 = __ldg(reinterpret_cast<const int4 *>(qtab + q));
    const int4 t1 = __ldg(reinterpret_cast<const int4 *>(qtab + q) + 1);
    const int sh = t1.z;
    const uint32_t sels = (uint32_t)t0.y, cA = (uint32_t)t0.z, cB = (uint32_t)t0.w, cC = (uint32_t)t1.x, cD = (uint32_t)t1.y;
    uint8_t *base = pyr + (long long)blockIdx.z * slab;
    const uint32_t *sp = reinterpret_cast<const uint32_t *>(base + src.off) + t0.x;
    uint8_t *dp = base + dst.off + 4 * q;
    const int spw = src.pitch >> 2;
#pragma unroll
    for (int r = 0; r < RESIZE_ROWS; r++) {
        const int y = blockIdx.y * (8 * RESIZE_ROWS) + r * 8 + threadIdx.y;
        if (y >= dst.h) break;
        const int2 yt = __ldg(ytab + dst.ytab + y);
        const int b0 = yt.y & 0xffff, b1 = (unsigned)yt.y >> 16;
        const int sy0 = min(max(yt.x, 0), src.h - 1), sy1 = min(max(yt.x + 1, 0), src.h - 1);
        const uint32_t *W0 = sp + sy0 * spw, *W1 = sp + sy1 * spw;
        const uint32_t u0 = W0[0], u1 = W0[1], u2 = W0[2], v0 = W1[0], v1 = W1[1], v2 = W1[2];
        const uint32_t a0 = __funnelshift_r(u0, u1, sh), a1 = __funnelshift_r(u1, u2, sh);
        const uint32_t e0 = __funnelshift_r(v0, v1, sh), e1 = __funnelshift_r(v1, v2, sh);
        uint32_t out;
        {
            const int h0 = (int)__dp2a_lo(cA, __byte_perm(a0, a1, sels), 0u), h1 = (int)__dp2a_lo(cA, __byte_perm(e0, e1, sels), 0u);
            out = (uint32_t)sb_lin_vert(h0, h1, b0, b1);
        }
        {
            const int h0 = (int)__dp2a_lo(cB, __byte_perm(a0, a1, sels >> 8), 0u), h1 = (int)__dp2a_lo(cB, __byte_perm(e0, e1, sels >> 8), 0u);
            out |= (uint32_t)sb_lin_vert(h0, h1, b0, b1) << 8;
        }
        {
            const int h0 = (int)__dp2a_lo(cC, __byte_perm(a0, a1, sels >> 16), 0u), h1 = (int)__dp2a_lo(cC, __byte_perm(e0, e1, sels >> 16), 0u);
            out |= (uint32_t)sb_lin_vert(h0, h1, b0, b1) << 16;
        }
        {
            const int h0 = (int)__dp2a_lo(cD, __byte_perm(a0, a1, sels >> 24), 0u), h1 = (int)__dp2a_lo(cD, __byte_perm(e0, e1, sels >> 24), 0u);
            out |= (uint32_t)sb_lin_vert(h0, h1, b0, b1) << 24;
        }
        *reinterpret_cast<uint32_t *>(dp + (long long)y * dst.pitch) = out;
    }
}

// cv::FAST(cell ROI, iniTh, nms) with the minTh fallback (ComputeKeyPointsOctTree :838-883 /
// Detect :1015-1060).  One CTA per GROUP of up to SB_FAST_KMAX horizontally adjacent cells of one grid
// row: the tested pixels of adjacent cells tile the row without gaps (cell j tests ROI columns
// [3 + j * wCell, 3 + (j + 1) * wCell)), so one TMA box and one pass over ~5 000 pixels serve all of them.
// The corner response does not depend on the threshold and "corner at t" <=> response >= t, so one
// response plane serves both thresholds; 3x3 non-maximum suppression is threshold independent for the
// survivors (a neighbour below t is also below the survivor).  Per-cell semantics of the reference are
// kept exactly: a neighbour that belongs to the next cell counts as 0 in the suppression, and the
// 20 -> 7 fallback is decided per cell from "did any maximum of this cell reach iniTh".
struct FastArgs {
    const CellGroup *groups;
    const uint8_t *mask_pyr;  // null: no mask
    uint32_t *cand;           // [batch][nlevels][SB_CAND_CAP]
    int *cand_cnt;            // [batch][nlevels]
    int *flags;               // [0] overflow
    long long slab;
    int nlevels, iniTh, minTh, tile_bytes, list_cap, seg;
    uint8_t *dbg;  // inspection: tile + response plane of group `dbg_cell` of image 0 (null in production calls)
    int dbg_cell;
};

__global__ void __launch_bounds__(FAST_THREADS) k_fast_cells(const __grid_constant__ TmaMaps maps,
                                                            const __grid_constant__ Geom g, FastArgs a) {
    sb_pdl_enter();
    // dynamic shared memory: [tile | response plane | maxima list | candidate-pixel list]; the array keeps its
    // shared address space through plain pointer arithmetic (an integer round trip would demote every
    // access to generic LD/ST with 64-bit address math)
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int s_n, s_any[SB_FAST_KMAX];
    uint8_t *tile = smem;  // TMA destination, 128-byte aligned
    uint8_t *sc = smem + a.tile_bytes;
    uint32_t *list = reinterpret_cast<uint32_t *>(smem);  // maxima list: reuses the tile, which is dead after phase 2
    uint16_t *plist = reinterpret_cast<uint16_t *>(smem + 2 * a.tile_bytes);
    if (threadIdx.x == 0 && (sb_smem_u32(tile) & 127u)) __trap();
    const CellGroup c = a.groups[blockIdx.x];
    const int img = blockIdx.y;
    const LevelGeom &L = g.lv[c.level];
    constexpr int BW = SB_FAST_BW;
    const int BH = L.fast_bh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xo = c.x0 & 15;  // the innermost TMA coordinate must be 16-byte aligned: ROI column x is tile column xo + x

    if (tid == 0) {
        s_n = 0;
        sb_mbar_init(&bar, 1);
        sb_mbar_expect_tx(&bar, (uint32_t)(BW * BH));
        sb_tma_load_3d(tile, &maps.m[c.level], c.x0 - xo, c.y0, img, &bar);
    }
    if (tid < SB_FAST_KMAX) s_any[tid] = 0;
    for (int i = tid; i < (BW * BH) >> 2; i += FAST_THREADS) reinterpret_cast<uint32_t *>(sc)[i] = 0u;
    __syncthreads();
    sb_mbar_wait(&bar, 0);

    // Two tiers, as the reference's two cv::FAST calls: tier 0 works at iniTh over the whole group (a pixel below iniTh
    // can neither be reported nor suppress a pixel at or above it, so neither its response nor its place in the list is
    // needed); tier 1 repeats the passes at minTh for the cells in which no maximum reached iniTh (rare on textured
    // images, cheap on flat ones) after the tier-0 maxima have been flushed and the tile re-fetched.
    const int t_hi = a.iniTh, t_lo = min(a.iniTh, a.minTh);
    const int wC = c.wCell;
    const int slot = img * a.nlevels + c.level;
    const unsigned lt = (1u << lane) - 1u;
    uint16_t *mine = plist + warp * a.seg;
    unsigned empty = 0;  // tier 1: bit j set = cell j of the group found nothing at iniTh
    for (int tier = 0;; tier++) {
        const int t0 = tier ? t_lo : t_hi;
        // Phase 1: pixels that pass the cheap necessary test (sb_fast_maybe's rule on the four compass ring pixels: one
        // of each opposite pair darker than v - t, or one of each pair brighter than v + t) are compacted into a list, so that the response
        // (about 80 instructions) is later computed by full warps instead of a few lanes of every warp.
        // A work item is 4 horizontally adjacent pixels: five aligned 32-bit loads, two funnel shifts, then the
        // order statistics for two pixels at a time in packed 2 x int16 arithmetic.  Every warp appends to its own
        // segment of the list with a warp-uniform running count: no atomics.  Pixels of the first / last item of a
        // row that fall outside the tested columns [3, rw - 3) are listed too and dropped in phase 2.
        // (the list is written through a running 32-bit shared-memory address: with a plain pointer nvcc rebuilds
        //  warp * seg from the kernel parameters for every store — 65 instructions per item instead of 32)
        const uint32_t mine_a = sb_smem_u32(mine);
        uint32_t wp = mine_a;
        asm volatile("" : "+r"(wp));
        {
            // only the 4-pixel groups that overlap the tested ROI columns [3, rw - 3)
            const int g0 = (xo + 3) >> 2, G = c.G, rows = c.rh - 6;
            const int items = (rows > 0 && c.rw > 6) ? rows * G : 0;
            // per 16-bit half: s2 + t + 512 - v in [257, 1022], "< 512" <=> bit 9 clear; no borrow crosses the halves,
            // so one 32-bit IADD3 does both pixels (there is no packed 16-bit integer add on sm_100a: __vadd2 is 5 ops)
            const uint32_t Tb = (uint32_t)(t0 + 512) * 0x00010001u;
            constexpr int bw4 = BW >> 2;
            for (int it0 = warp * 32; it0 < items; it0 += FAST_THREADS) {
                const int it = it0 + lane;
                uint32_t neg[2] = {0u, 0u};  // bit 9 / bit 25 of neg[h]: pixel 2h / 2h + 1 of the item passes
                int e0 = 0;
                bool live = it < items;
                int y = 0, col = 0;
                if (live) {
                    const int yy = (int)(((uint32_t)it * (uint32_t)c.rcpG) >> 20);  // it / G
                    y = 3 + yy;
                    col = (g0 + it - yy * G) << 2;  // tile column of pixel 0
                    e0 = (y << 8) + col - xo;
                    if (tier) {  // only items that touch a cell without a tier-0 maximum (an item spans at most two cells)
                        const int xa = min(max(col - xo, 3), c.rw - 4), xb = min(max(col - xo + 3, 3), c.rw - 4);
                        const int ja = (int)(((uint32_t)(xa - 3) * (uint32_t)c.rcpW) >> 20);
                        const int jb = (int)(((uint32_t)(xb - 3) * (uint32_t)c.rcpW) >> 20);
                        live = ((empty >> ja) | (empty >> jb)) & 1u;
                    }
                }
                if (live) {
                    const uint32_t *pw = reinterpret_cast<const uint32_t *>(tile + y * BW + col);
                    const uint32_t C = pw[0], U = pw[-3 * bw4], D = pw[3 * bw4];
                    const uint32_t Lw = __funnelshift_r(pw[-1], C, 8), R = __funnelshift_r(C, pw[1], 24);
#pragma unroll
                    for (int hpair = 0; hpair < 2; hpair++) {
                        const uint32_t sel = hpair ? 0x4342u : 0x4140u;  // bytes (2,3) or (0,1) -> two 16-bit halves
                        const uint32_t v = __byte_perm(C, 0u, sel), r0 = __byte_perm(D, 0u, sel), r8 = __byte_perm(U, 0u, sel);
                        const uint32_t r4 = __byte_perm(R, 0u, sel), r12 = __byte_perm(Lw, 0u, sel);
                        // a 9-arc holds at least one pixel of each opposite pair (0, 8) and (4, 12): the brighter of the two
                        // pair minima must be darker than v - t, or the darker of the two pair maxima brighter than v + t
                        const uint32_t s2 = __vmaxs2(__vmins2(r0, r8), __vmins2(r4, r12));
                        const uint32_t s3 = __vmins2(__vmaxs2(r0, r8), __vmaxs2(r4, r12));
                        // darker: s2 < v - t  <=>  s2 + t - v < 0 ;  brighter: s3 > v + t  <=>  v + t - s3 < 0
                        const uint32_t dk = s2 + Tb - v, br = v + Tb - s3;
                        neg[hpair] = ~(dk & br);  // bit 9 of a half clear in either
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // one predicate serves the ballot and the store (as C++ nvcc derives it twice); the entry is (y << 8) | ROI column
                    asm volatile(
                        "{\n"
                        ".reg .pred p;\n"
                        ".reg .b32 t, b, n, at;\n"
                        "and.b32 t, %1, %2;\n"
                        "setp.ne.u32 p, t, 0;\n"
                        "vote.sync.ballot.b32 b, p, 0xffffffff;\n"
                        "and.b32 t, b, %3;\n"
                        "popc.b32 n, t;\n"
                        "mad.lo.u32 at, n, 2, %0;\n"
                        "@p st.shared.u16 [at], %4;\n"
                        "popc.b32 n, b;\n"
                        "mad.lo.u32 %0, n, 2, %0;\n"
                        "}"
                        : "+r"(wp)
                        : "r"(neg[j >> 1]), "r"((j & 1) ? 0x02000000u : 0x00000200u), "r"(lt), "h"((uint16_t)(e0 + j))
                        : "memory");
                }
            }
        }
        const int cnt = (int)((wp - mine_a) >> 1);
        __syncwarp();
        // Phase 2: responses of the listed pixels; every warp walks its own segment.  Only ROI columns [3, rw - 3)
        // are tested by cv::FAST: the response plane stays 0 elsewhere.  (Tier 1 recomputes a few tier-0 responses of
        // the cells it revisits: same values.)
        for (int i = lane; i < cnt; i += 32) {
            const int e = mine[i];
            const int y = e >> 8, x = e & 255;
            if (x < 3 || x >= c.rw - 3) continue;
            if (tier && !((empty >> (int)(((uint32_t)(x - 3) * (uint32_t)c.rcpW) >> 20)) & 1u)) continue;
            const int s = sb_fast_score(tile + y * BW + xo + x, BW);
            if (s >= t0) sc[y * BW + xo + x] = (uint8_t)s;
        }
        __syncthreads();
        if (a.dbg && img == 0 && (int)blockIdx.x == a.dbg_cell && tier == 0) {
            for (int i = tid; i < a.tile_bytes; i += FAST_THREADS) a.dbg[i] = tile[i];
            __syncthreads();
        }
        // Phase 3: 3x3 non-maximum suppression (only listed pixels can be maxima).  Tested columns only; a neighbour
        // in the adjacent cell counts as 0, exactly as if each cell had been given to cv::FAST on its own.
        // (measured: appending the maxima of a pass with one warp-aggregated atomic costs 108 instead of 71 instructions per pass)
        for (int i = lane; i < cnt; i += 32) {
            const int e = mine[i];
            const int y = e >> 8, x = e & 255;
            const uint8_t *q = sc + y * BW + xo + x;
            const int s = q[0];
            if (s == 0) continue;  // listed by the pre-test but not a corner
            const int jl = (int)(((uint32_t)(x - 3) * (uint32_t)c.rcpW) >> 20);  // cell of the group
            if (tier && !((empty >> jl) & 1u)) continue;  // that cell was served by tier 0
            const int xr = x - 3 - jl * wC;
            int nb = max((int)q[-BW], (int)q[BW]);
            if (xr != 0) nb = max(nb, max(max((int)q[-1], (int)q[-BW - 1]), (int)q[BW - 1]));
            if (xr != wC - 1) nb = max(nb, max(max((int)q[1], (int)q[-BW + 1]), (int)q[BW + 1]));
            if (s > nb) {  // s > 0 follows: neighbours are >= 0
                const int k = atomicAdd(&s_n, 1);
                if (k < a.list_cap) list[k] = (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
                if (s >= a.iniTh) s_any[jl] = 1;  // tier 0: every maximum; tier 1: none (its cells have no maximum >= iniTh)
            }
        }
        __syncthreads();
        if (a.dbg && img == 0 && (int)blockIdx.x == a.dbg_cell && tier == 0) {
            for (int i = tid; i < a.tile_bytes; i += FAST_THREADS) a.dbg[a.tile_bytes + i] = sc[i];
            if (tid == 0) {
                int *info = reinterpret_cast<int *>(a.dbg + 2 * a.tile_bytes);
                info[0] = BW; info[1] = BH; info[2] = xo; info[3] = c.x0; info[4] = c.y0; info[5] = c.rw; info[6] = c.rh;
                info[7] = s_n; info[8] = s_any[0]; info[9] = c.level; info[10] = c.ncells; info[11] = c.wCell;
            }
        }
        // the maxima of this tier go to the level's candidate list: one global atomic per warp
        const int n = min(s_n, a.list_cap);
        for (int base = 0; base < n; base += FAST_THREADS) {
            const int i = base + tid;
            bool ok = false;
            uint32_t out = 0;
            if (i < n) {
                const uint32_t w = list[i];
                const int s = (int)(w >> 24), xl = (int)(w & 0xfff);
                const int jl = (int)(((uint32_t)(xl - 3) * (uint32_t)c.rcpW) >> 20);
                // the reference's second cv::FAST call (minTh) only if the first (iniTh) found nothing in this cell
                ok = s >= (s_any[jl] ? a.iniTh : a.minTh);
                const int x = xl + c.offx, y = (int)((w >> 12) & 0xfff) + c.offy;  // border-relative
                // quirk Q1 (:871-877): the mask is read at the border-relative coordinates
                if (ok && a.mask_pyr && a.mask_pyr[(long long)img * a.slab + L.off + (long long)y * L.pitch + x] == 0) ok = false;
                out = (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (bal) {
                int pos = 0;
                if (lane == 0) pos = atomicAdd(&a.cand_cnt[slot], __popc(bal));
                pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(bal & lt);
                if (ok) {
                    if (pos < SB_CAND_CAP)
                        a.cand[(long long)slot * SB_CAND_CAP + pos] = out;
                    else
                        a.flags[0] = 1;
                }
            }
        }
        if (tier || a.minTh >= a.iniTh) break;  // a stricter second threshold finds a subset of nothing
        for (int j = 0; j < c.ncells; j++)
            if (!s_any[j]) empty |= 1u << j;
        if (!empty) break;
        // the maxima list lives in the tile: fetch the tile again for the second tier
        __syncthreads();
        if (tid == 0) {
            s_n = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            sb_mbar_expect_tx(&bar, (uint32_t)(BW * BH));
            sb_tma_load_3d(tile, &maps.m[c.level], c.x0 - xo, c.y0, img, &bar);
        }
        __syncthreads();
        sb_mbar_wait(&bar, 1);
    }
}

// DistributeOctTree: one CTA per (image, level).
struct QtArgs {
    const uint32_t *cand;
    const int *cand_cnt;
    uint32_t *sel;   // [batch][nlevels][selcap]
    int *sel_cnt;    // [batch][nlevels]
    int nlevels, selcap, ncap, candcap_smem;
    int N_override;  // > 0: Detect() — level 0 with N = nfeatures
    uint8_t *spill;  // [batch][nlevels][SB_CAND_CAP * 3]: candidate state of the slots with more than QT_CAND_SMEM candidates
};

// Candidates whose per-candidate state (node position, quadrant: 3 bytes) lives in shared memory.  A level of a KITTI-sized
// frame has 1 000 - 4 000 candidates; sizing the arrays for the hard cap of 16 384 cost 48 KB per CTA — three quadtree CTAs
// then fill an SM and starve the blur that runs beside them (blur + quadtree window 0.255 -> 0.235 ms).  A (image, level)
// with more candidates keeps that state in a global scratch area instead (same code, generic pointers).
#ifndef QT_CAND_SMEM
#define QT_CAND_SMEM 6144
#endif
static size_t qt_smem_bytes(int ncap) {
    int P = 1;
    while (P < ncap) P <<= 1;
    size_t b = 0;
    b += (size_t)ncap * 16;          // childcnt (also the 64-bit "best" array)
    b += (size_t)P * 4;              // keys
    b += 64 * 4;                     // scratch
    b += 5 * (size_t)ncap * 4;       // ord cpre spre cbase ubase
    b += 2 * (size_t)ncap * 8;       // boxes
    b += 2 * (size_t)ncap * 2;       // counts
    b += (size_t)QT_CAND_SMEM * 2;    // cnode
    b += (size_t)QT_CAND_SMEM;        // cq
    return b + 64;
}

__global__ void __launch_bounds__(QT_THREADS, 5) k_quadtree(const __grid_constant__ Geom g, QtArgs a) {
    sb_pdl_enter();
    extern __shared__ __align__(16) uint8_t smem[];
    const int level = blockIdx.x, img = blockIdx.y;
    const LevelGeom &L = g.lv[level];
    const int slot = img * a.nlevels + level;
    const int ncap = a.ncap;
    int P = 1;
    while (P < ncap) P <<= 1;

    QtCtx c;
    uint8_t *p = smem;
    c.childcnt = reinterpret_cast<int *>(p); p += (size_t)ncap * 16;
    c.keys = reinterpret_cast<uint32_t *>(p); p += (size_t)P * 4;
    c.scratch = reinterpret_cast<int *>(p); p += 64 * 4;
    c.ord = reinterpret_cast<int *>(p); p += (size_t)ncap * 4;
    c.cpre = reinterpret_cast<int *>(p); p += (size_t)ncap * 4;
    c.spre = reinterpret_cast<int *>(p); p += (size_t)ncap * 4;
    c.cbase = reinterpret_cast<int *>(p); p += (size_t)ncap * 4;
    c.ubase = reinterpret_cast<int *>(p); p += (size_t)ncap * 4;
    c.box[0] = reinterpret_cast<int16_t *>(p); p += (size_t)ncap * 8;
    c.box[1] = reinterpret_cast<int16_t *>(p); p += (size_t)ncap * 8;
    c.cnt[0] = reinterpret_cast<uint16_t *>(p); p += (size_t)ncap * 2;
    c.cnt[1] = reinterpret_cast<uint16_t *>(p); p += (size_t)ncap * 2;
    c.cnode = reinterpret_cast<uint16_t *>(p); p += (size_t)QT_CAND_SMEM * 2;
    c.cq = p;
    c.cand = a.cand + (long long)slot * SB_CAND_CAP;
    c.n = min(a.cand_cnt[slot], SB_CAND_CAP);
    c.ncap = ncap;
    c.width = L.bw;
    c.height = L.bh;
    c.N = a.N_override > 0 ? a.N_override : L.quota;
    c.nCols = L.nCols;
    c.wCell = L.wCell;
    c.hCell = L.hCell;
    int S;
    if (c.n > QT_CAND_SMEM) {  // rare: this CTA's candidate state goes to its slot of the global scratch area.  Two inlined copies
                               // of the pass code, so that the common one keeps its shared-memory addressing (one copy on generic
                               // pointers: 48 -> 60 registers and the blur + quadtree window back at 0.259 ms)
        uint8_t *gs = a.spill + (size_t)slot * ((size_t)SB_CAND_CAP * 3);
        c.cnode = reinterpret_cast<uint16_t *>(gs);
        c.cq = gs + (size_t)SB_CAND_CAP * 2;
        S = qt_distribute(c, a.sel + (long long)slot * a.selcap, a.selcap);
    } else {
        S = qt_distribute(c, a.sel + (long long)slot * a.selcap, a.selcap);
    }
    if (threadIdx.x == 0) a.sel_cnt[slot] = min(S, a.selcap);
}

// GaussianBlur 7x7 sigma 2 on every level (DetectAndCompute :965-966, CalcDescriptors :1194-1199).
// One CTA per 128 x 32 output tile; the 134 x 38 input window arrives as one TMA box (zero filled
// outside the image); BORDER_REFLECT_101 is resolved by index inside the tile.
struct BlurTile {
    short level, tx, ty, pad;
};
struct BlurArgs {
    const BlurTile *tiles;
    uint8_t *blur;
    long long slab;
};

__global__ void __launch_bounds__(256) k_blur(const __grid_constant__ TmaMaps maps, const __grid_constant__ Geom g,
                                             BlurArgs a) {
    __shared__ __align__(128) uint8_t tile[BLUR_BH * BLUR_BW];
    __shared__ __align__(16) uint16_t hrow[BLUR_BH * BLUR_TW];
    __shared__ __align__(8) uint64_t bar;
    const BlurTile t = a.tiles[blockIdx.x];
    const int img = blockIdx.y;
    const LevelGeom &L = g.lv[t.level];
    const int x0 = t.tx * BLUR_TW, y0 = t.ty * BLUR_TH;
    const int tid = threadIdx.x;
    if (tid == 0) {
        sb_mbar_init(&bar, 1);
        sb_mbar_expect_tx(&bar, BLUR_BH * BLUR_BW);
        sb_tma_load_3d(tile, &maps.m[t.level], x0 - BLUR_HX, y0 - 3, img, &bar);
    }
    __syncthreads();
    sb_mbar_wait(&bar, 0);

    // BORDER_REFLECT_101: tiles on the image border complete their halo inside shared memory
    // (columns first, then whole rows, so the corners come out right); interior tiles skip this.
    const bool left = x0 == 0, right = x0 + BLUR_TW + 3 > L.w, top = y0 == 0, bottom = y0 + BLUR_TH + 3 > L.h;
    if (left || right) {
        for (int i = tid; i < BLUR_BH * 3; i += 256) {
            const int r = i / 3, k = i % 3 + 1;
            uint8_t *row = tile + r * BLUR_BW;
            if (left) row[BLUR_HX - k] = row[BLUR_HX + k];
            if (right) {
                const int col = L.w - 1 + k - x0 + BLUR_HX, src = L.w - 1 - k - x0 + BLUR_HX;
                if (col < BLUR_BW) row[col] = row[src];
            }
        }
        __syncthreads();
    }
    if (top || bottom) {
        for (int i = tid; i < 3 * BLUR_BW; i += 256) {
            const int k = i / BLUR_BW + 1, c = i % BLUR_BW;
            if (top) tile[(3 - k) * BLUR_BW + c] = tile[(3 + k) * BLUR_BW + c];
            if (bottom) {
                const int r = L.h - 1 + k - (y0 - 3), src = L.h - 1 - k - (y0 - 3);
                if (r < BLUR_BH) tile[r * BLUR_BW + c] = tile[src * BLUR_BW + c];
            }
        }
        __syncthreads();
    }

    // horizontal pass: 4 outputs per item from 12 tile bytes; each output is two 4-tap dot products (DP4A)
    const uint32_t W0 = SB_G0 | (SB_G1 << 8) | (SB_G2 << 16) | (SB_G3 << 24), W1 = SB_G2 | (SB_G1 << 8) | (SB_G0 << 16);
    for (int i = tid; i < BLUR_BH * (BLUR_TW / 4); i += 256) {
        const int r = i / (BLUR_TW / 4), gq = i % (BLUR_TW / 4);
        // output column cx reads tile columns (BLUR_HX - 3) + cx + 0..6; BLUR_HX - 3 = 13 = 12 + 1
        const uint32_t *wp = reinterpret_cast<const uint32_t *>(tile + r * BLUR_BW + (BLUR_HX - 4) + 4 * gq);
        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
        const uint32_t o0 = __dp4a(__funnelshift_r(w0, w1, 8), W0, __dp4a(__funnelshift_r(w1, w2, 8), W1, 0u));
        const uint32_t o1 = __dp4a(__funnelshift_r(w0, w1, 16), W0, __dp4a(__funnelshift_r(w1, w2, 16), W1, 0u));
        const uint32_t o2 = __dp4a(__funnelshift_r(w0, w1, 24), W0, __dp4a(__funnelshift_r(w1, w2, 24), W1, 0u));
        const uint32_t o3 = __dp4a(w1, W0, __dp4a(w2, W1, 0u));
        *reinterpret_cast<uint2 *>(hrow + r * BLUR_TW + 4 * gq) = make_uint2(o0 | (o1 << 16), o2 | (o3 << 16));  // each <= 255 * 256
    }
    __syncthreads();
    // vertical pass: a thread owns 4 columns x 4 output rows (10 input rows).  Vertically adjacent 16-bit row sums of one
    // column are paired in a word (PRMT) so that a 7-tap column is four DP2As against the byte pairs (18, 34), (48, 56),
    // (48, 34), (18, 0); output rows 0 / 2 use the pairs that start on even input rows, rows 1 / 3 those on odd rows.
    {
        const int gq = tid & 31, seg = tid >> 5;
        const int ry0 = seg * 4;
        if (y0 + ry0 < L.h && x0 + 4 * gq < L.w) {
            uint2 u[10];
#pragma unroll
            for (int r = 0; r < 10; r++) u[r] = *reinterpret_cast<const uint2 *>(hrow + (ry0 + r) * BLUR_TW + 4 * gq);
            // pr[k][j]: rows (k, k + 1) of column j as (lower row | upper row << 16), k = 0 .. 8
            uint32_t pr[9][4];
#pragma unroll
            for (int k = 0; k < 9; k++) {
                pr[k][0] = __byte_perm(u[k].x, u[k + 1].x, 0x5410);
                pr[k][1] = __byte_perm(u[k].x, u[k + 1].x, 0x7632);
                pr[k][2] = __byte_perm(u[k].y, u[k + 1].y, 0x5410);
                pr[k][3] = __byte_perm(u[k].y, u[k + 1].y, 0x7632);
            }
            const uint32_t WA = SB_G0 | (SB_G1 << 8), WB = SB_G2 | (SB_G3 << 8), WC = SB_G2 | (SB_G1 << 8), WD = SB_G0, WE = SB_G0 << 8;
            uint8_t *out = a.blur + (long long)img * a.slab + L.off + x0 + 4 * gq;
#pragma unroll
            for (int ry = 0; ry < 4; ry++) {
                const int y = y0 + ry0 + ry;
                if (y >= L.h) break;
                uint32_t o = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    // rows ry .. ry + 6: pairs (ry, ry+1), (ry+2, ry+3), (ry+4, ry+5) and row ry + 6 alone
                    uint32_t s = __dp2a_lo(pr[ry][j], WA, 32768u);
                    s = __dp2a_lo(pr[ry + 2][j], WB, s);
                    s = __dp2a_lo(pr[ry + 4][j], WC, s);
                    s = ry < 3 ? __dp2a_lo(pr[ry + 6][j], WD, s) : __dp2a_lo(pr[8][j], WE, s);  // row 9 is the upper half of pair 8
                    o |= (s >> 16) << (8 * j);
                }
                *reinterpret_cast<uint32_t *>(out + (long long)y * L.pitch) = o;
            }
        }
    }
}

// IC_Angle (:27-55): moments over the radius-15 disc, one warp, lane = column u.  Column u holds the rows
// |v| <= nv(u) = #{v >= 1 : |u| <= umax[v]} (ic_rows, computed once per thread); the row pair (+v, -v) is read
// through two running pointers, and m10 = u * (sum of the column) is formed once at the end.
static __device__ __forceinline__ int ic_rows(const int *umax, int lane) {
    if (lane >= 31) return -1;
    const int au = abs(lane - SB_HALF_PATCH);
    int nv = 0;
#pragma unroll
    for (int v = 1; v <= SB_HALF_PATCH; v++) nv += au <= umax[v];
    return nv;
}
// SMEM: the patch sits in shared memory at a compile-time pitch — fully unrolled, every row offset an LDS immediate.
template <bool SMEM = false>
static __device__ __forceinline__ float warp_ic_angle(const uint8_t *center, int pitch, int nv, int lane) {
    int m10 = 0, m01 = 0;
    if (nv >= 0) {
        const int u = lane - SB_HALF_PATCH;
        const uint8_t *pu = center + u, *pd = pu;
        int col = pu[0];
        if (SMEM) {
#pragma unroll
            for (int v = 1; v <= SB_HALF_PATCH; v++) {
                if (v <= nv) {
                    const int vp = pu[v * pitch], vm = pu[-v * pitch];
                    col += vp + vm;
                    m01 += v * (vp - vm);
                }
            }
        } else {
#pragma unroll 5
            for (int v = 1; v <= SB_HALF_PATCH; v++) {
                pu += pitch;
                pd -= pitch;
                if (v <= nv) {
                    const int vp = *pu, vm = *pd;
                    col += vp + vm;
                    m01 += v * (vp - vm);
                }
            }
        }
        m10 = u * col;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    return sb_fast_atan2((float)m01, (float)m10);
}

// The 256 test pairs as floats in shared memory, transposed so that lane i's j-th pair (pair 8i + j) sits
// at [j * 32 + i]: conflict-free 16-byte loads, no int8 -> float conversions in the loop.
struct PatternF4 { float v[1024]; };
static constexpr PatternF4 make_pattern_f4() {  // the int8 table as floats, transposed: pair 8 i + j at [j * 32 + i]
    constexpr int p[1024] = {
#include "orb_pattern.inc"
    };
    PatternF4 t{};
    for (int i = 0; i < 256; i++)
        for (int k = 0; k < 4; k++) t.v[((i & 7) * 32 + (i >> 3)) * 4 + k] = (float)p[4 * i + k];
    return t;
}
__device__ __align__(16) const PatternF4 d_pattern_f4 = make_pattern_f4();
static __device__ __forceinline__ void load_pattern(float4 *pat) {  // one 16-byte load + store per thread
    for (int i = threadIdx.x; i < 256; i += blockDim.x) pat[i] = reinterpret_cast<const float4 *>(d_pattern_f4.v)[i];
}

// computeOrbDescriptor (:59-98): lane i produces byte i (pairs 8i .. 8i+7).
static __device__ __forceinline__ uint32_t warp_brief_byte(const uint8_t *center, int pitch, float angle_deg,
                                                           const float4 *pat, int lane) {
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = sb_fmul(angle_deg, factorPI);
    double sn, cs;
    sincos((double)ang, &sn, &cs);
    const float a = (float)cs, b = (float)sn;
    uint32_t val = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const float4 q = pat[j * 32 + lane];
        const int t0 = sb_brief_sample(center, pitch, a, b, q.x, q.y);
        const int t1 = sb_brief_sample(center, pitch, a, b, q.z, q.w);
        val |= (uint32_t)(t0 < t1) << j;
    }
    return val;
}

static __device__ __forceinline__ void warp_store_keypoint(sb_keypoint *dst, float x, float y, float size, float angle,
                                                           float response, int octave, int class_id, int lane) {
    uint32_t w;
    switch (lane) {
        case 0: w = __float_as_uint(x); break;
        case 1: w = __float_as_uint(y); break;
        case 2: w = __float_as_uint(size); break;
        case 3: w = __float_as_uint(angle); break;
        case 4: w = __float_as_uint(response); break;
        case 5: w = (uint32_t)octave; break;
        default: w = (uint32_t)class_id; break;
    }
    if (lane < 7) reinterpret_cast<uint32_t *>(dst)[lane] = w;
}

// Orientation, descriptor and the level-major output of DetectAndCompute (:905-906, :970-983).
struct DescArgs {
    const uint8_t *pyr, *blur;
    const uint32_t *sel;
    const int *sel_cnt;
    sb_keypoint *kps;   // [batch][cap]
    uint8_t *desc;      // [batch][cap][32] or null
    int32_t *counts;    // [batch]
    int *flags;         // [1] capacity
    long long slab;
    int nlevels, selcap, cap;
};

// k_describe stages every keypoint's patch in shared memory with ONE TMA box per phase (no staging instructions): the
// orientation disc (radius 15) from the pyramid, then the rBRIEF sampling window (the pattern reaches 18 px when rotated)
// from the blurred pyramid into the same buffer.  The innermost TMA coordinate must be a multiple of 16 bytes, so a box
// starts at the 16-byte boundary at or left of the wanted column.  With the patch at a compile-time pitch the row
// offsets become LDS immediates (the global-memory version spent half its instructions on 64-bit addresses) and the
// byte gathers hit shared-memory banks instead of 20+ L1 sectors per warp load.
#ifndef DESC_KPB
#define DESC_KPB 16      // keypoints per CTA: 2 per warp, both patches in flight
#endif
#define DESC_AW 48       // orientation box: 48 x 31 bytes (x - 15 .. x + 15 after alignment)
#define DESC_AH 31
#define DESC_CW 64       // sampling box: 64 x 37 bytes (x - 18 .. x + 18 after alignment)
#define DESC_CH 37
#ifndef DESC_DUAL
#define DESC_DUAL 0       // 1: separate buffers, both boxes requested up front (one memory round trip per CTA instead of two)
#endif
#define DESC_CBYTES 2432  // 64 * 37 = 2368 rounded up to 128 (TMA destination alignment)
#define DESC_ABYTES 1536  // 48 * 31 = 1488 rounded up
#define DESC_PBYTES (DESC_DUAL ? DESC_CBYTES + DESC_ABYTES : DESC_CBYTES)
__global__ void __launch_bounds__(DESC_WARPS * 32) k_describe(const __grid_constant__ TmaMaps pmaps,
                                                             const __grid_constant__ TmaMaps bmaps,
                                                             const __grid_constant__ Geom g, DescArgs a) {
    sb_pdl_enter();
    extern __shared__ __align__(128) uint8_t patches[];  // [DESC_KPB][DESC_PBYTES]
    __shared__ float4 pat[256];
    __shared__ float s_ang[DESC_KPB], s_cos[DESC_KPB], s_sin[DESC_KPB];
    __shared__ __align__(8) uint64_t bars[DESC_KPB], bars2[DESC_KPB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int level = blockIdx.y, img = blockIdx.z;
    const int *cnt = a.sel_cnt + img * a.nlevels;
    const int k0 = blockIdx.x * DESC_KPB;
    if (cnt[level] <= k0) {  // most CTAs of the sparse upper levels: nothing to do (the level-0 CTA 0 of an image never takes this exit)
        if (!(level == 0 && blockIdx.x == 0)) return;
    }
    int offset = 0, total = 0;
    for (int l = 0; l < a.nlevels; l++) {
        const int c = cnt[l];
        if (l < level) offset += c;
        total += c;
    }
    if (level == 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        a.counts[img] = min(total, a.cap);
        if (total > a.cap) a.flags[1] = 1;
    }
    const int nk = min(min(cnt[level], a.cap - offset) - k0, DESC_KPB);  // keypoints of this CTA
    if (nk <= 0) return;
    const LevelGeom &L = g.lv[level];
    const uint32_t *sel = a.sel + ((long long)img * a.nlevels + level) * a.selcap + k0;
    // every warp owns the barriers and buffers of its keypoints (i = warp, warp + 8, ...): no block-wide hand-over; lane 0
    // initialises a barrier and requests the box in one go, the other lanes first touch it after the __syncwarp below
    for (int i = warp; i < nk; i += DESC_WARPS) {
        const uint32_t w = sel[i];
        const int x = (int)(w & 0xfff) + SB_EDGE - 3, y = (int)((w >> 12) & 0xfff) + SB_EDGE - 3;  // :895-896
        if (lane == 0) {
            sb_mbar_init(&bars[i], 1);
            if (DESC_DUAL) sb_mbar_init(&bars2[i], 1);
            sb_mbar_expect_tx(&bars[i], DESC_AW * DESC_AH);
            sb_tma_load_3d(patches + i * DESC_PBYTES + (DESC_DUAL ? DESC_CBYTES : 0), &pmaps.m[level], (x - SB_HALF_PATCH) & ~15, y - SB_HALF_PATCH, img, &bars[i]);
            if (DESC_DUAL && a.desc) {
                sb_mbar_expect_tx(&bars2[i], DESC_CW * DESC_CH);
                sb_tma_load_3d(patches + i * DESC_PBYTES, &bmaps.m[level], (x - 18) & ~15, y - 18, img, &bars2[i]);
            }
        }
    }
    __syncwarp();
    load_pattern(pat);
    // phase A: orientation, one warp per keypoint; the buffer is then refilled with the blurred sampling window
    const int nv = ic_rows(g.umax, lane);
    for (int i = warp; i < nk; i += DESC_WARPS) {
        const uint32_t w = sel[i];
        const int x = (int)(w & 0xfff) + SB_EDGE - 3, y = (int)((w >> 12) & 0xfff) + SB_EDGE - 3;
        uint8_t *buf = patches + i * DESC_PBYTES;
        sb_mbar_wait(&bars[i], 0);
        const float angle = warp_ic_angle<true>(buf + (DESC_DUAL ? DESC_CBYTES : 0) + SB_HALF_PATCH * DESC_AW + SB_HALF_PATCH + ((x - SB_HALF_PATCH) & 15), DESC_AW, nv, lane);
        if (lane == 0) s_ang[i] = angle;
        if (!DESC_DUAL && a.desc) {
            __syncwarp();  // every lane has read the disc
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                sb_mbar_expect_tx(&bars[i], DESC_CW * DESC_CH);
                sb_tma_load_3d(buf, &bmaps.m[level], (x - 18) & ~15, y - 18, img, &bars[i]);
            }
        }
    }
    __syncthreads();
    // phase B: cos / sin in double, ONE THREAD per keypoint (as a warp-wide computation it would cost 32x)
    if (a.desc && threadIdx.x < nk) {
        const float factorPI = (float)(3.14159265358979323846 / 180.f);
        double sn, cs;
        sincos((double)sb_fmul(s_ang[threadIdx.x], factorPI), &sn, &cs);
        s_cos[threadIdx.x] = (float)cs;
        s_sin[threadIdx.x] = (float)sn;
    }
    __syncthreads();
    // phase C: descriptor (lane i -> byte i) and the keypoint record
    for (int i = warp; i < nk; i += DESC_WARPS) {
        const uint32_t w = sel[i];
        const int x = (int)(w & 0xfff) + SB_EDGE - 3, y = (int)((w >> 12) & 0xfff) + SB_EDGE - 3;
        const long long row = (long long)img * a.cap + offset + k0 + i;
        if (a.desc) {
            if (DESC_DUAL) sb_mbar_wait(&bars2[i], 0); else sb_mbar_wait(&bars[i], 1);
            const uint8_t *center = patches + i * DESC_PBYTES + 18 * DESC_CW + 18 + ((x - 18) & 15);
            const float ca = s_cos[i], sb = s_sin[i];
            uint32_t val = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const float4 q = pat[j * 32 + lane];
                const int t0 = sb_brief_sample(center, DESC_CW, ca, sb, q.x, q.y);
                const int t1 = sb_brief_sample(center, DESC_CW, ca, sb, q.z, q.w);
                val |= (uint32_t)(t0 < t1) << j;
            }
            a.desc[row * 32 + lane] = (uint8_t)val;
        }
        float fx = (float)x, fy = (float)y;
        if (level != 0) { fx = sb_fmul(fx, L.scale); fy = sb_fmul(fy, L.scale); }  // :975-981
        warp_store_keypoint(a.kps + row, fx, fy, L.patch, s_ang[i], (float)(w >> 24), level, -1, lane);
    }
}

// Output of Detect (:1062-1073): FAST's size 7 / angle -1 / octave 0 are kept (quirk Q4).
__global__ void k_emit_detect(const uint32_t *sel, const int *sel_cnt, int nlevels, int selcap, sb_keypoint *kps,
                              int32_t *counts, int cap, int *flags) {
    const int img = blockIdx.y;
    const int n = sel_cnt[img * nlevels];
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        counts[img] = min(n, cap);
        if (n > cap) flags[1] = 1;
    }
    if (k >= n || k >= cap) return;
    const uint32_t w = sel[(long long)img * nlevels * selcap + k];
    sb_keypoint kp;
    kp.x = (float)((int)(w & 0xfff) + SB_EDGE - 3);
    kp.y = (float)((int)((w >> 12) & 0xfff) + SB_EDGE - 3);
    kp.size = 7.f;
    kp.angle = -1.f;
    kp.response = (float)(w >> 24);
    kp.octave = 0;
    kp.class_id = -1;
    kps[(long long)img * cap + k] = kp;
}

// ScreenAndComputeKPsParams (:1083-1129): one warp per input keypoint; `kps` is mutated like the
// reference mutates its input vector; keep[i] = 1 for survivors.
static __device__ __forceinline__ bool is_fast_corner_at(const uint8_t *p, int pitch, int threshold) {
    // isFastCorner (:449-511): more than 8 contiguous ring pixels darker than v - t or brighter than v + t
    const int v = p[0];
    uint32_t dark = 0, bright = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const int r = p[SB_RING_DY(k) * pitch + SB_RING_DX(k)];
        dark |= (uint32_t)(r < v - threshold) << k;
        bright |= (uint32_t)(r > v + threshold) << k;
    }
    dark |= dark << 16;      // the reference walks 25 entries = the ring plus its first 9 again
    bright |= bright << 16;
    bool hit = false;
#pragma unroll
    for (int s = 0; s < 16; s++) {
        hit |= ((dark >> s) & 0x1ffu) == 0x1ffu;
        hit |= ((bright >> s) & 0x1ffu) == 0x1ffu;
    }
    return hit;
}

// per_img > 0: a batch of images, keypoint i belongs to image i / per_img and only the first counts[image] slots are live
__global__ void __launch_bounds__(DESC_WARPS * 32) k_screen(const __grid_constant__ Geom g, const uint8_t *pyr,
                                                           sb_keypoint *kps, int n, uint8_t *keep, int minTh, int per_img,
                                                           const int32_t *counts) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    if (per_img > 0) {
        const int b = i / per_img;
        if (i - b * per_img >= counts[b]) return;
        pyr += (long long)b * g.slab;
    }
    sb_keypoint kp = kps[i];
    const int level = kp.octave;
    bool ok = level >= 0 && level < g.nlevels;
    float x = kp.x, y = kp.y, angle = kp.angle, size = kp.size;
    if (ok) {
        const LevelGeom &L = g.lv[level];
        x = sb_fdiv(x, L.scale);
        y = sb_fdiv(y, L.scale);
        ok = sb_fsub(y, (float)SB_EDGE) >= 0.f && sb_fadd(y, (float)SB_EDGE) < (float)L.h &&
             sb_fsub(x, (float)SB_EDGE) >= 0.f && sb_fadd(x, (float)SB_EDGE) < (float)L.w;
        if (ok) {
            const uint8_t *c = pyr + L.off + (long long)sb_rint(y) * L.pitch + sb_rint(x);
            ok = is_fast_corner_at(c, L.pitch, minTh);
            if (ok) {
                angle = warp_ic_angle(c, L.pitch, ic_rows(g.umax, lane), lane);
                size = sb_fmul(31.f, L.scale);
            }
        }
        x = sb_fmul(x, L.scale);  // quirk Q5: the round trip happens for rejected keypoints too
        y = sb_fmul(y, L.scale);
    }
    warp_store_keypoint(kps + i, x, y, size, angle, kp.response, kp.octave, kp.class_id, lane);
    if (lane == 0) keep[i] = ok ? 1 : 0;
}

// CalcDescriptors (:1180-1226): row i = descriptor of keypoint i on the blurred level `octave`.
__global__ void __launch_bounds__(DESC_WARPS * 32) k_calc_desc(const __grid_constant__ Geom g, const uint8_t *blur,
                                                              const sb_keypoint *kps, int n, uint8_t *desc, int per_img,
                                                              const uint8_t *keep) {
    __shared__ float4 pat[256];
    load_pattern(pat);
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * DESC_WARPS + (threadIdx.x >> 5);
    if (i >= n) return;
    if (per_img > 0) {   // batch form: only the keypoints the screening kept, on their own image's pyramid
        if (!keep[i]) return;
        blur += (long long)(i / per_img) * g.slab;
    }
    const sb_keypoint kp = kps[i];
    const LevelGeom &L = g.lv[kp.octave];
    const float x = sb_fdiv(kp.x, L.scale), y = sb_fdiv(kp.y, L.scale);
    const uint8_t *c = blur + L.off + (long long)sb_rint(y) * L.pitch + sb_rint(x);
    desc[(long long)i * 32 + lane] = (uint8_t)warp_brief_byte(c, L.pitch, kp.angle, pat, lane);
}

// ================================================================================================
// host side
// ================================================================================================
struct sb_orb {
    int device;
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;  // the reference keeps the float argument in a double member (ORBextractor.h:125)
    int max_w, max_h, max_batch;
    float scale[SB_MAX_LEVELS], inv_scale[SB_MAX_LEVELS], sigma2[SB_MAX_LEVELS], inv_sigma2[SB_MAX_LEVELS];
    int quota[SB_MAX_LEVELS];
    int umax[16];
    cudaStream_t stream, own_stream, side_stream;
    cudaEvent_t ev_fork, ev_join;
    // geometry of the current image size
    int cur_w, cur_h;
    Geom geom;
    TmaMaps fast_maps, blur_maps, blur_maps_mask_unused;
    TmaMaps desc_pyr_maps, desc_blur_maps;  // k_describe's patch boxes over the pyramid / the blurred pyramid
    int n_cells, n_cells_l0, n_blur_tiles;  // n_cells: FAST cell groups (CTAs)
    int fast_tile_bytes, fast_list_cap, fast_seg;
    int selcap, ncap_pyr, ncap_detect;
    long long slab_cap;  // bytes reserved per image
    int kp_cap;          // sb_orb_capacity()
    // device memory
    uint8_t *d_pyr, *d_blur, *d_mask;
    CellGroup *d_cells;
    BlurTile *d_tiles;
    int2 *d_xtab, *d_ytab;  // resize tables: (source index, c0 | c1 << 16)
    ResizeQuad *d_qtab;
    uint32_t *d_cand, *d_sel;
    int *d_cand_cnt, *d_sel_cnt, *d_flags;
    uint8_t *d_qt_spill;  // quadtree candidate state of crowded (image, level) slots
    int tab_cap, cell_cap, tile_cap;
    // staging for the host-pointer entry points
    uint8_t *d_in, *d_in_mask, *d_desc_out;
    sb_keypoint *d_kps_out;
    int32_t *d_counts_out;
    uint8_t *d_keep;
    size_t in_cap;
    int stage_cap, stage_batch;
    int *h_flags;  // pinned
    uint8_t *h_stage;     // pinned staging of sb_orb_screen_describe's results (grow-only)
    size_t h_stage_cap;
    uint8_t *dbg_buf;  // inspection only (sb_orb_debug_fast_cell)
    int dbg_cell;
    // per-stage CUDA-event timing (sb_orb_profile): event pairs recorded on the launching stream
    int prof_on, prof_n;
    cudaEvent_t prof_ev[SB_PROF_MAX][2];
    int prof_stage[SB_PROF_MAX], prof_launches[SB_PROF_MAX];
};

static const char *const kStageName[SB_STAGE_COUNT] = {"copy_level0", "resize_pyramid", "fast_cells", "quadtree", "gauss_blur", "describe"};
static void prof_begin(sb_orb *h, int stage, int launches, cudaStream_t s) {
    nvtxRangePushA(kStageName[stage]);  // closed by prof_end: the stage's launches sit inside the range
    if (!h->prof_on || h->prof_n >= SB_PROF_MAX) return;
    const int i = h->prof_n;
    if (!h->prof_ev[i][0]) {
        cudaEventCreate(&h->prof_ev[i][0]);
        cudaEventCreate(&h->prof_ev[i][1]);
    }
    h->prof_stage[i] = stage;
    h->prof_launches[i] = launches;
    cudaEventRecord(h->prof_ev[i][0], s);
}
static void prof_end(sb_orb *h, cudaStream_t s) {
    nvtxRangePop();
    if (!h->prof_on || h->prof_n >= SB_PROF_MAX) return;
    cudaEventRecord(h->prof_ev[h->prof_n][1], s);
    h->prof_n++;
}

static void free_orb(sb_orb *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_pyr,  h->d_blur, h->d_mask,     h->d_cells,   h->d_tiles,   h->d_xtab,    h->d_ytab,    h->d_qtab,
                    h->d_cand,     h->d_sel,     h->d_cand_cnt, h->d_sel_cnt, h->d_flags, h->d_qt_spill,
                    h->d_in,   h->d_in_mask, h->d_desc_out, h->d_kps_out, h->d_counts_out, h->d_keep};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < SB_PROF_MAX; i++) {
        if (h->prof_ev[i][0]) cudaEventDestroy(h->prof_ev[i][0]);
        if (h->prof_ev[i][1]) cudaEventDestroy(h->prof_ev[i][1]);
    }
    delete h;
}

// ORBextractor::ORBextractor (:384-445): scale tables, per-level quotas, umax.
static void build_tables(sb_orb *h) {
    const int n = h->nlevels;
    h->scale[0] = 1.0f;
    h->sigma2[0] = 1.0f;
    for (int i = 1; i < n; i++) {
        h->scale[i] = (float)(h->scale[i - 1] * h->scaleFactor);
        h->sigma2[i] = h->scale[i] * h->scale[i];
    }
    for (int i = 0; i < n; i++) {
        h->inv_scale[i] = 1.0f / h->scale[i];
        h->inv_sigma2[i] = 1.0f / h->sigma2[i];
    }
    const float factor = (float)(1.0f / h->scaleFactor);
    float want = h->nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)n));
    int sum = 0;
    for (int l = 0; l < n - 1; l++) {
        h->quota[l] = (int)lrintf(want);
        sum += h->quota[l];
        want *= factor;
    }
    h->quota[n - 1] = h->nfeatures - sum > 0 ? h->nfeatures - sum : 0;
    // umax: quarter circle of radius 15, made symmetric about the diagonal
    const int vmax = (int)floorf(SB_HALF_PATCH * sqrtf(2.f) / 2 + 1);
    const int vmin = (int)ceilf(SB_HALF_PATCH * sqrtf(2.f) / 2);
    const double r2 = (double)SB_HALF_PATCH * SB_HALF_PATCH;
    memset(h->umax, 0, sizeof(h->umax));
    for (int v = 0; v <= vmax; v++) h->umax[v] = (int)lrint(sqrt(r2 - (double)v * v));
    for (int v = SB_HALF_PATCH, v0 = 0; v >= vmin; --v) {
        while (h->umax[v0] == h->umax[v0 + 1]) ++v0;
        h->umax[v] = v0;
        ++v0;
    }
}

static void level_size(const sb_orb *h, int w, int hgt, int level, int *lw, int *lh) {
    if (level == 0) { *lw = w; *lh = hgt; return; }
    *lw = (int)lrintf((float)w * h->inv_scale[level]);   // Size(cvRound(cols * scale), cvRound(rows * scale)) :1237-1238
    *lh = (int)lrintf((float)hgt * h->inv_scale[level]);
}

static long long slab_bytes(const sb_orb *h, int w, int hgt) {
    long long off = 0;
    for (int l = 0; l < h->nlevels; l++) {
        int lw, lh;
        level_size(h, w, hgt, l, &lw, &lh);
        if (lw < 1 || lh < 1) return -1;
        off += (long long)sb_align_up((size_t)sb_align_up(lw, 128) * lh, 256);
    }
    return off;
}

static size_t fast_smem_bytes(const sb_orb *h) {
    return 2 * (size_t)h->fast_tile_bytes + (size_t)(FAST_THREADS / 32) * h->fast_seg * 2 + 128;
}

static int quadtree_ncap(int N, int bw, int bh) {
    const int nIni = (int)roundf((float)bw / (float)bh);
    const int a = 4 * (nIni > 0 ? nIni : 1), b = N + 3;
    return (a > b ? a : b) + 1;
}

// Everything that depends on the image size: level geometry, resize tables, FAST cells, blur tiles,
// tensor maps.  Cached for the last (w, h).
static int configure(sb_orb *h, int w, int hgt) {
    if (h->cur_w == w && h->cur_h == hgt) return SB_OK;
    SB_REQUIRE(w <= h->max_w && hgt <= h->max_h, "image larger than max_w x max_h given at create time");
    Geom &g = h->geom;
    memset(&g, 0, sizeof(g));
    g.nlevels = h->nlevels;
    memcpy(g.umax, h->umax, sizeof(g.umax));
    std::vector<int2> xtab, ytab;
    std::vector<ResizeQuad> qtab;
    std::vector<CellGroup> cells;
    std::vector<BlurTile> tiles;
    long long off = 0;
    int fast_tile = 0, fast_list = 0, fast_items = 0, selcap = 0, ncap_pyr = 0, kp_cap = 0;
    for (int l = 0; l < h->nlevels; l++) {
        LevelGeom &L = g.lv[l];
        level_size(h, w, hgt, l, &L.w, &L.h);
        // every level must hold at least one 30-px FAST cell inside the 16-px border (:826-836)
        SB_REQUIRE(L.w - 32 >= 30 && L.h - 32 >= 30, "image too small for this number of pyramid levels");
        L.pitch = (int)sb_align_up(L.w, 128);
        L.off = off;
        off += (long long)sb_align_up((size_t)L.pitch * L.h, 256);
        L.scale = h->scale[l];
        L.patch = (float)(int)(31 * h->scale[l]);
        L.quota = h->quota[l];
        // FAST grid (:826-836), float arithmetic as in the reference
        const float width = (float)(L.w - 32), height = (float)(L.h - 32), W = 30.f;
        L.bw = L.w - 32;
        L.bh = L.h - 32;
        L.nCols = (int)(width / W);
        L.nRows = (int)(height / W);
        L.wCell = (int)ceilf(width / L.nCols);
        L.hCell = (int)ceilf(height / L.nRows);
        SB_REQUIRE(L.wCell + 6 <= 72 && L.hCell + 6 <= 72, "FAST cell larger than supported");
        // cells per CTA: runs of K adjacent cells of a grid row, K * wCell <= SB_FAST_SPAN, rows split evenly
        int K = SB_FAST_SPAN / L.wCell;
        K = K < 1 ? 1 : K > SB_FAST_KMAX ? SB_FAST_KMAX : K;
        const int runs = sb_div_up(L.nCols, K);
        K = sb_div_up(L.nCols, runs);
        SB_REQUIRE(K * L.wCell + 6 + 15 <= SB_FAST_BW, "internal: FAST tile width");
        L.fast_bw = SB_FAST_BW;
        L.fast_bh = L.hCell + 6;
        if (L.fast_bw * L.fast_bh > fast_tile) fast_tile = L.fast_bw * L.fast_bh;
        if (K * L.wCell * L.hCell / 4 + 8 > fast_list) fast_list = K * L.wCell * L.hCell / 4 + 8;  // maxima: at most 1 per 2 x 2
        const int minB = SB_EDGE - 3, maxBX = L.w - SB_EDGE + 3, maxBY = L.h - SB_EDGE + 3;
        for (int i = 0; i < L.nRows; i++) {
            const int iniY = minB + i * L.hCell;
            int maxY = iniY + L.hCell + 6;
            if (iniY >= maxBY - 3) continue;
            if (maxY > maxBY) maxY = maxBY;
            for (int j0 = 0; j0 < L.nCols; j0 += K) {
                CellGroup c;
                memset(&c, 0, sizeof(c));
                int endX = 0;
                for (int j = j0; j < L.nCols && j < j0 + K; j++) {  // the reference's column loop (:849-855)
                    const int iniX = minB + j * L.wCell;
                    int maxX = iniX + L.wCell + 6;
                    if (iniX >= maxBX - 6) continue;
                    if (maxX > maxBX) maxX = maxBX;
                    c.ncells++;
                    endX = maxX;
                }
                if (c.ncells == 0) continue;
                const int iniX = minB + j0 * L.wCell;
                c.level = (short)l; c.x0 = (short)iniX; c.y0 = (short)iniY;
                c.rw = (short)(endX - iniX); c.rh = (short)(maxY - iniY);
                c.offx = (short)(j0 * L.wCell); c.offy = (short)(i * L.hCell);
                c.wCell = (short)L.wCell;
                const int xo = iniX & 15;
                const int G = c.rw > 6 ? ((xo + c.rw - 4) >> 2) - ((xo + 3) >> 2) + 1 : 1;
                c.G = (short)G;
                c.rcpG = ((1 << 20) + G - 1) / G;
                c.rcpW = ((1 << 20) + L.wCell - 1) / L.wCell;
                const int items = c.rh > 6 ? (c.rh - 6) * G : 0;
                if (items > fast_items) fast_items = items;
                cells.push_back(c);
            }
        }
        if (l == 0) h->n_cells_l0 = (int)cells.size();
        // resize tables for level l from level l-1 (unused for l == 0)
        L.xtab = (int)xtab.size();
        L.ytab = (int)ytab.size();
        if (l > 0) {
            const LevelGeom &S = g.lv[l - 1];
            for (int d = 0; d < L.w; d++) {
                SbLinCoef c = sb_lin_coef(d, L.w, S.w, true);
                xtab.push_back(make_int2(c.s, (int)((uint32_t)(uint16_t)c.c0 | ((uint32_t)(uint16_t)c.c1 << 16))));
            }
            for (int d = 0; d < L.h; d++) {
                SbLinCoef c = sb_lin_coef(d, L.h, S.h, false);
                ytab.push_back(make_int2(c.s, (int)((uint32_t)(uint16_t)c.c0 | ((uint32_t)(uint16_t)c.c1 << 16))));
            }
            // quad table (k_resize_quads): usable when every group of 4 destination pixels reads <= 8 consecutive source bytes
            const size_t q_start = qtab.size();
            L.qtab = (int)q_start;
            const int nq = sb_div_up(L.w, 4);
            for (int q = 0; q < nq && L.qtab >= 0; q++) {
                ResizeQuad e;
                memset(&e, 0, sizeof(e));
                const int2 *xt = &xtab[L.xtab];
                e.wb = xt[4 * q].x >> 2;
                e.sh = 8 * (xt[4 * q].x & 3);
                for (int k = 0; k < 4; k++) {
                    const int2 t = xt[4 * q + k < L.w ? 4 * q + k : L.w - 1];
                    const int o = t.x - xt[4 * q].x;
                    if (o < 0 || o + 1 > 7) { L.qtab = -1; break; }  // scale factor > 2: generic kernel
                    e.sels |= (uint32_t)(o | ((o + 1) << 4)) << (8 * k);
                    e.c[k] = (uint32_t)t.y;
                }
                qtab.push_back(e);
            }
            if (L.qtab < 0) qtab.resize(q_start);
        }
        for (int ty = 0; ty < sb_div_up(L.h, BLUR_TH); ty++)
            for (int tx = 0; tx < sb_div_up(L.w, BLUR_TW); tx++) {
                BlurTile t = {(short)l, (short)tx, (short)ty, 0};
                tiles.push_back(t);
            }
        const int nc = quadtree_ncap(L.quota, L.bw, L.bh);
        if (nc > ncap_pyr) ncap_pyr = nc;
        kp_cap += nc - 1;
    }
    g.slab = off;
    SB_REQUIRE(off <= h->slab_cap, "internal: slab larger than reserved");
    h->ncap_pyr = ncap_pyr;
    h->ncap_detect = quadtree_ncap(h->nfeatures, g.lv[0].bw, g.lv[0].bh);
    selcap = (ncap_pyr > h->ncap_detect ? ncap_pyr : h->ncap_detect);
    SB_REQUIRE(selcap <= h->selcap, "internal: selection capacity");
    h->kp_cap = kp_cap > h->ncap_detect - 1 ? kp_cap : h->ncap_detect - 1;
    SB_REQUIRE(qt_smem_bytes(ncap_pyr > h->ncap_detect ? ncap_pyr : h->ncap_detect) <= 220 * 1024,
               "nfeatures too large for the on-chip quadtree");
    h->fast_tile_bytes = (int)sb_align_up(fast_tile, 128);
    h->fast_list_cap = (int)sb_align_up(fast_list, 32);
    SB_REQUIRE(h->fast_list_cap * 4 <= h->fast_tile_bytes, "internal: FAST maxima list does not fit the tile");
    h->fast_seg = sb_div_up(fast_items, FAST_THREADS) * 32 * 4;  // per-warp segment of the candidate-pixel list (phase 1)
    SB_REQUIRE(fast_smem_bytes(h) <= 200 * 1024, "internal: FAST shared memory");
    h->n_cells = (int)cells.size();
    h->n_blur_tiles = (int)tiles.size();
    SB_REQUIRE((int)xtab.size() <= h->tab_cap && (int)ytab.size() <= h->tab_cap, "internal: table capacity");
    SB_REQUIRE(h->n_cells <= h->cell_cap && h->n_blur_tiles <= h->tile_cap, "internal: cell/tile capacity");
    cudaStream_t s = h->stream;
    // the previous geometry may still be in use by queued kernels
    SB_CUDA(cudaStreamSynchronize(s));
    if (!qtab.empty()) {
        SB_REQUIRE((int)qtab.size() <= h->tab_cap, "internal: quad table capacity");
        SB_CUDA(cudaMemcpyAsync(h->d_qtab, qtab.data(), qtab.size() * sizeof(ResizeQuad), cudaMemcpyHostToDevice, s));
    }
    if (!xtab.empty()) {
        SB_CUDA(cudaMemcpyAsync(h->d_xtab, xtab.data(), xtab.size() * 8, cudaMemcpyHostToDevice, s));
        SB_CUDA(cudaMemcpyAsync(h->d_ytab, ytab.data(), ytab.size() * 8, cudaMemcpyHostToDevice, s));
    }
    SB_CUDA(cudaMemcpyAsync(h->d_cells, cells.data(), cells.size() * sizeof(CellGroup), cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_tiles, tiles.data(), tiles.size() * sizeof(BlurTile), cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaStreamSynchronize(s));
    for (int l = 0; l < h->nlevels; l++) {
        const LevelGeom &L = g.lv[l];
        const uint64_t dims[3] = {(uint64_t)L.w, (uint64_t)L.h, (uint64_t)h->max_batch};
        const uint64_t strides[2] = {(uint64_t)L.pitch, (uint64_t)g.slab};
        const uint32_t fbox[3] = {(uint32_t)L.fast_bw, (uint32_t)L.fast_bh, 1};
        const uint32_t bbox[3] = {BLUR_BW, BLUR_BH, 1};
        SB_TRY(sb_make_tensor_map_u8(&h->fast_maps.m[l], h->d_pyr + L.off, 3, dims, strides, fbox));
        SB_TRY(sb_make_tensor_map_u8(&h->blur_maps.m[l], h->d_pyr + L.off, 3, dims, strides, bbox));
        const uint32_t dabox[3] = {DESC_AW, DESC_AH, 1}, dcbox[3] = {DESC_CW, DESC_CH, 1};
        SB_TRY(sb_make_tensor_map_u8(&h->desc_pyr_maps.m[l], h->d_pyr + L.off, 3, dims, strides, dabox));
        SB_TRY(sb_make_tensor_map_u8(&h->desc_blur_maps.m[l], h->d_blur + L.off, 3, dims, strides, dcbox));
    }
    h->cur_w = w;
    h->cur_h = hgt;
    return SB_OK;
}

extern "C" int sb_orb_create(sb_orb_t **out, int device, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                             int minThFAST, int max_w, int max_h, int max_batch) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(nfeatures >= 1 && nfeatures <= 20000, "nfeatures out of range [1, 20000]");
    SB_REQUIRE(nlevels >= 1 && nlevels <= SB_MAX_LEVELS, "nlevels out of range [1, 12]");
    SB_REQUIRE(scaleFactor > 1.0f || nlevels == 1, "scaleFactor must be > 1");
    SB_REQUIRE(iniThFAST >= 1 && iniThFAST <= 255 && minThFAST >= 1 && minThFAST <= 255, "FAST thresholds out of range [1, 255]");
    SB_REQUIRE(max_w >= 62 && max_w <= 4096 && max_h >= 62 && max_h <= 4096, "max_w / max_h out of range [62, 4096]");
    SB_REQUIRE(max_batch >= 1 && max_batch <= 4096, "max_batch out of range [1, 4096]");
    SB_TRY(sb_use_device(device));
    sb_orb *h = new sb_orb();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->nfeatures = nfeatures;
    h->scaleFactor = scaleFactor;
    h->nlevels = nlevels;
    h->iniTh = iniThFAST;
    h->minTh = minThFAST;
    h->max_w = max_w;
    h->max_h = max_h;
    h->max_batch = max_batch;
    build_tables(h);
    h->slab_cap = slab_bytes(h, max_w, max_h);
    if (h->slab_cap < 0) {
        delete h;
        sb_set_error("max_w x max_h too small for %d pyramid levels", nlevels);
        return SB_ERR_INVALID;
    }
    h->slab_cap += 4096;
    // selection capacity: the largest quadtree node table any call can need
    {
        int ncap = 0;
        for (int l = 0; l < nlevels; l++) {
            int n = (h->quota[l] > 64 ? h->quota[l] : 64) + 3 + 1;
            if (n > ncap) ncap = n;
        }
        int nd = (nfeatures > 64 ? nfeatures : 64) + 3 + 1;
        h->selcap = (ncap > nd ? ncap : nd) + 600;  // 4 * nIni roots for very elongated images
    }
    h->tab_cap = 64;
    h->cell_cap = 0;
    h->tile_cap = 0;
    for (int l = 0; l < nlevels; l++) {
        int lw, lh;
        level_size(h, max_w, max_h, l, &lw, &lh);
        h->tab_cap += lw > lh ? lw : lh;
        h->cell_cap += (lw / 30 + 1) * (lh / 30 + 1);
        h->tile_cap += sb_div_up(lw, BLUR_TW) * sb_div_up(lh, BLUR_TH);
    }
#define SB_ALLOC(ptr, bytes)                                                        \
    do {                                                                            \
        cudaError_t e__ = cudaMalloc((void **)&(ptr), (bytes));                     \
        if (e__ != cudaSuccess) {                                                   \
            sb_set_error("cudaMalloc(%zu bytes) for %s -> %s", (size_t)(bytes), #ptr, cudaGetErrorString(e__)); \
            free_orb(h);                                                            \
            return SB_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)
    const size_t B = (size_t)max_batch;
    SB_ALLOC(h->d_pyr, B * h->slab_cap);
    SB_ALLOC(h->d_blur, B * h->slab_cap);
    SB_ALLOC(h->d_cells, (size_t)h->cell_cap * sizeof(CellGroup));
    SB_ALLOC(h->d_tiles, (size_t)h->tile_cap * sizeof(BlurTile));
    SB_ALLOC(h->d_xtab, (size_t)h->tab_cap * 8);
    SB_ALLOC(h->d_ytab, (size_t)h->tab_cap * 8);
    SB_ALLOC(h->d_qtab, (size_t)h->tab_cap * sizeof(ResizeQuad));
    SB_ALLOC(h->d_cand, B * nlevels * SB_CAND_CAP * 4);
    SB_ALLOC(h->d_qt_spill, B * nlevels * SB_CAND_CAP * 3);
    SB_ALLOC(h->d_sel, B * nlevels * h->selcap * 4);
    SB_ALLOC(h->d_cand_cnt, B * nlevels * 4);
    SB_ALLOC(h->d_sel_cnt, B * nlevels * 4);
    SB_ALLOC(h->d_flags, 16);
    cudaError_t e = cudaMemset(h->d_flags, 0, 16);
    if (e == cudaSuccess) e = cudaMemset(h->d_sel_cnt, 0, B * nlevels * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_flags, 16);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fast_cells, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_describe, cudaFuncAttributeMaxDynamicSharedMemorySize, DESC_KPB * DESC_PBYTES);
    if (e != cudaSuccess) {
        sb_set_error("sb_orb_create: %s", cudaGetErrorString(e));
        free_orb(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    h->cur_w = h->cur_h = -1;
    int rc = configure(h, max_w, max_h);  // validates the size now and makes sb_orb_capacity() meaningful
    if (rc != SB_OK) {
        free_orb(h);
        return rc;
    }
    *out = h;
    return SB_OK;
}

static int finish_and_check(sb_orb *h);

extern "C" int sb_orb_destroy(sb_orb_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_orb(h);
    }
    return SB_OK;
}

extern "C" int sb_orb_set_stream(sb_orb_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_orb_sync_status(sb_orb_t *h) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_TRY(sb_use_device(h->device));
    return finish_and_check(h);
}

extern "C" int sb_orb_status_async(sb_orb_t *h, int32_t *host_flags) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && host_flags, "null pointer");
    SB_TRY(sb_use_device(h->device));
    SB_CUDA(cudaMemcpyAsync(host_flags, h->d_flags, 16, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemsetAsync(h->d_flags, 0, 16, h->stream));
    return SB_OK;
}

extern "C" int sb_orb_status_decode(const int32_t *host_flags) {
    SB_REQUIRE(host_flags, "null pointer");
    if (host_flags[0]) {
        sb_set_error("more than %d FAST candidates on one pyramid level", SB_CAND_CAP);
        return SB_ERR_OVERFLOW;
    }
    if (host_flags[1]) {
        sb_set_error("keypoint capacity `cap` too small; use sb_orb_capacity()");
        return SB_ERR_CAPACITY;
    }
    return SB_OK;
}

extern "C" int sb_orb_capacity(const sb_orb_t *h) { return h ? h->kp_cap : SB_ERR_INVALID; }

extern "C" int sb_orb_get_tables(const sb_orb_t *h, int *nlevels, float *scale, float *inv_scale, float *sigma2,
                                 float *inv_sigma2, int *features_per_level) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    if (nlevels) *nlevels = h->nlevels;
    for (int i = 0; i < h->nlevels; i++) {
        if (scale) scale[i] = h->scale[i];
        if (inv_scale) inv_scale[i] = h->inv_scale[i];
        if (sigma2) sigma2[i] = h->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = h->inv_sigma2[i];
        if (features_per_level) features_per_level[i] = h->quota[i];
    }
    return SB_OK;
}

// ---- pipeline pieces (all asynchronous on h->stream) ------------------------------------------------
static int ensure_mask_buffer(sb_orb *h) {
    if (h->d_mask) return SB_OK;
    SB_CUDA(cudaMalloc((void **)&h->d_mask, (size_t)h->max_batch * h->slab_cap));
    return SB_OK;
}

static int launch_pyramid(sb_orb *h, uint8_t *pyr, const uint8_t *d_img, long long img_pitch, int stride, int batch,
                          int nlevels_to_build) {
    const Geom &g = h->geom;
    const LevelGeom &L0 = g.lv[0];
    const long long total = (long long)batch * L0.h * (L0.pitch / (d_img ? 16 : 4));
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    prof_begin(h, SB_STAGE_COPY, 1, h->stream);
    if (d_img)
        SB_CUDA(sb_launch_pdl(k_copy_level0, dim3(blocks), dim3(256), 0, h->stream, d_img, img_pitch, stride, L0.w, L0.h, pyr, g.slab, L0.pitch, batch));
    else
        k_fill_level0<<<blocks, 256, 0, h->stream>>>(pyr, g.slab, L0.pitch, L0.h, batch);
    prof_end(h, h->stream);
    if (nlevels_to_build > 1) prof_begin(h, SB_STAGE_RESIZE, nlevels_to_build - 1, h->stream);
    for (int l = 1; l < nlevels_to_build; l++) {
        const LevelGeom &D = g.lv[l];
        if (D.qtab >= 0) {
            const int nq = sb_div_up(D.w, 4);
            dim3 grid(sb_div_up(nq, 32), sb_div_up(D.h, 8 * RESIZE_ROWS), batch);
            SB_CUDA(sb_launch_pdl(k_resize_quads, grid, dim3(32, 8), 0, h->stream, pyr, g.slab, g.lv[l - 1], D, h->d_qtab + D.qtab, h->d_ytab, nq));
        } else {
            dim3 grid(sb_div_up(sb_div_up(D.w, 4), 256), D.h, batch);
            SB_CUDA(sb_launch_pdl(k_resize, grid, dim3(256), 0, h->stream, pyr, g.slab, g.lv[l - 1], D, h->d_xtab, h->d_ytab));
        }
    }
    if (nlevels_to_build > 1) prof_end(h, h->stream);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

static int launch_blur(sb_orb *h, int batch, cudaStream_t s) {
    BlurArgs a = {h->d_tiles, h->d_blur, h->geom.slab};
    prof_begin(h, SB_STAGE_BLUR, 1, s);
    k_blur<<<dim3(h->n_blur_tiles, batch), 256, 0, s>>>(h->blur_maps, h->geom, a);
    prof_end(h, s);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

static int launch_blur(sb_orb *h, int batch, cudaStream_t s);

// blur_beside_quadtree: the Gaussian blur of the same batch is forked onto the side stream right after the FAST
// kernel, so that it fills the issue slots the latency-bound quadtree leaves idle (FAST itself is issue bound).
static int launch_fast_and_quadtree(sb_orb *h, int batch, bool use_mask, bool detect_only, bool blur_beside_quadtree = false) {
    const int nl = h->nlevels;
    FastArgs fa;
    fa.groups = h->d_cells;
    fa.mask_pyr = use_mask ? h->d_mask : nullptr;
    fa.cand = h->d_cand;
    fa.cand_cnt = h->d_cand_cnt;
    fa.flags = h->d_flags;
    fa.slab = h->geom.slab;
    fa.nlevels = nl;
    fa.iniTh = h->iniTh;
    fa.minTh = h->minTh;
    fa.tile_bytes = h->fast_tile_bytes;
    fa.list_cap = h->fast_list_cap;
    fa.seg = h->fast_seg;
    fa.dbg = h->dbg_buf;
    fa.dbg_cell = h->dbg_cell;
    const int ncells = detect_only ? h->n_cells_l0 : h->n_cells;
    const size_t fsmem = fast_smem_bytes(h);
    prof_begin(h, SB_STAGE_FAST, 1, h->stream);
    SB_CUDA(sb_launch_pdl(k_fast_cells, dim3(ncells, batch), dim3(FAST_THREADS), fsmem, h->stream, h->fast_maps, h->geom, fa));
    prof_end(h, h->stream);
    if (blur_beside_quadtree) {
        SB_CUDA(cudaEventRecord(h->ev_fork, h->stream));
        SB_CUDA(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        SB_TRY(launch_blur(h, batch, h->side_stream));
        SB_CUDA(cudaEventRecord(h->ev_join, h->side_stream));
    }
    QtArgs qa;
    qa.cand = h->d_cand;
    qa.cand_cnt = h->d_cand_cnt;
    qa.sel = h->d_sel;
    qa.sel_cnt = h->d_sel_cnt;
    qa.nlevels = nl;
    qa.selcap = h->selcap;
    qa.ncap = detect_only ? h->ncap_detect : h->ncap_pyr;
    qa.candcap_smem = QT_CAND_SMEM;
    qa.spill = h->d_qt_spill;
    qa.N_override = detect_only ? h->nfeatures : 0;
    prof_begin(h, SB_STAGE_QUADTREE, 1, h->stream);
    SB_CUDA(sb_launch_pdl(k_quadtree, dim3(detect_only ? 1 : nl, batch), dim3(QT_THREADS), qt_smem_bytes(qa.ncap), h->stream, h->geom, qa));
    prof_end(h, h->stream);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

static int check_shapes(sb_orb *h, int batch, int w, int hgt, int stride, int cap) {
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch out of range [1, max_batch]");
    SB_REQUIRE(w >= 62 && hgt >= 62 && stride >= w, "bad image size / stride");
    SB_REQUIRE(cap >= 1, "cap must be positive");
    SB_TRY(sb_use_device(h->device));
    return configure(h, w, hgt);
}

// Reads the device flags after the stream drained (host-pointer entry points only).
static int finish_and_check(sb_orb *h) {
    SB_CUDA(cudaMemcpyAsync(h->h_flags, h->d_flags, 16, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemsetAsync(h->d_flags, 0, 16, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    if (h->h_flags[0]) {
        sb_set_error("more than %d FAST candidates on one pyramid level", SB_CAND_CAP);
        return SB_ERR_OVERFLOW;
    }
    if (h->h_flags[1]) {
        sb_set_error("keypoint capacity `cap` too small; use sb_orb_capacity()");
        return SB_ERR_CAPACITY;
    }
    return SB_OK;
}

extern "C" int sb_orb_detect_and_compute_dev(sb_orb_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes,
                                             const uint8_t *d_mask, int64_t mask_pitch_bytes, int w, int hgt,
                                             int stride, int mstride, sb_keypoint *d_kps, uint8_t *d_desc,
                                             int32_t *d_counts, int cap) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_TRY(check_shapes(h, batch, w, hgt, stride, cap));
    SB_REQUIRE(d_img && d_kps && d_counts, "null device pointer");
    SB_REQUIRE(!d_mask || mstride >= w, "bad mask stride");
    // (the candidate counters are reset here, not between the pyramid and FAST: the kernels then follow one another without
    //  a memset node in between, which programmatic dependent launch needs)
    SB_CUDA(cudaMemsetAsync(h->d_cand_cnt, 0, (size_t)batch * h->nlevels * 4, h->stream));
    SB_TRY(launch_pyramid(h, h->d_pyr, d_img, img_pitch_bytes, stride, batch, h->nlevels));
    if (d_mask) {
        SB_TRY(ensure_mask_buffer(h));
        SB_TRY(launch_pyramid(h, h->d_mask, d_mask, mask_pitch_bytes, mstride, batch, h->nlevels));
    }
    // the blur only depends on the pyramid: it runs on the side stream beside the quadtree
    SB_TRY(launch_fast_and_quadtree(h, batch, d_mask != nullptr, false, d_desc != nullptr));
    if (d_desc) SB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    DescArgs da;
    da.pyr = h->d_pyr;
    da.blur = h->d_blur;
    da.sel = h->d_sel;
    da.sel_cnt = h->d_sel_cnt;
    da.kps = d_kps;
    da.desc = d_desc;
    da.counts = d_counts;
    da.flags = h->d_flags;
    da.slab = h->geom.slab;
    da.nlevels = h->nlevels;
    da.selcap = h->selcap;
    da.cap = cap;
    prof_begin(h, SB_STAGE_DESCRIBE, 1, h->stream);
    SB_CUDA(sb_launch_pdl(k_describe, dim3(sb_div_up(h->ncap_pyr, DESC_KPB), h->nlevels, batch), dim3(DESC_WARPS * 32), DESC_KPB * DESC_PBYTES,
                          h->stream, h->desc_pyr_maps, h->desc_blur_maps, h->geom, da));
    prof_end(h, h->stream);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_orb_detect_dev(sb_orb_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes,
                                 const uint8_t *d_mask, int64_t mask_pitch_bytes, int w, int hgt, int stride,
                                 int mstride, sb_keypoint *d_kps, int32_t *d_counts, int cap) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_TRY(check_shapes(h, batch, w, hgt, stride, cap));
    SB_REQUIRE(d_img && d_kps && d_counts, "null device pointer");
    SB_REQUIRE(!d_mask || mstride >= w, "bad mask stride");
    SB_CUDA(cudaMemsetAsync(h->d_cand_cnt, 0, (size_t)batch * h->nlevels * 4, h->stream));
    SB_TRY(launch_pyramid(h, h->d_pyr, d_img, img_pitch_bytes, stride, batch, 1));
    if (d_mask) {
        SB_TRY(ensure_mask_buffer(h));
        SB_TRY(launch_pyramid(h, h->d_mask, d_mask, mask_pitch_bytes, mstride, batch, 1));
    }
    SB_TRY(launch_fast_and_quadtree(h, batch, d_mask != nullptr, true));
    k_emit_detect<<<dim3(sb_div_up(h->ncap_detect, 256), batch), 256, 0, h->stream>>>(h->d_sel, h->d_sel_cnt, h->nlevels,
                                                                                   h->selcap, d_kps, d_counts, cap,
                                                                                   h->d_flags);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// ---- host-pointer entry points: stage, run, copy back, synchronise ---------------------------------
static int ensure_staging(sb_orb *h, int batch, size_t img_bytes, int cap, bool need_mask) {
    const size_t need = (size_t)batch * img_bytes;
    if (need > h->in_cap) {
        SB_CUDA(cudaStreamSynchronize(h->stream));
        if (h->d_in) cudaFree(h->d_in);
        if (h->d_in_mask) cudaFree(h->d_in_mask);
        h->d_in = h->d_in_mask = nullptr;
        h->in_cap = 0;
        SB_CUDA(cudaMalloc((void **)&h->d_in, need));
        h->in_cap = need;
    }
    if (need_mask && !h->d_in_mask) SB_CUDA(cudaMalloc((void **)&h->d_in_mask, h->in_cap));
    if (cap > h->stage_cap || batch > h->stage_batch) {
        SB_CUDA(cudaStreamSynchronize(h->stream));
        if (h->d_kps_out) cudaFree(h->d_kps_out);
        if (h->d_desc_out) cudaFree(h->d_desc_out);
        if (h->d_counts_out) cudaFree(h->d_counts_out);
        if (h->d_keep) cudaFree(h->d_keep);
        h->d_kps_out = nullptr; h->d_desc_out = nullptr; h->d_counts_out = nullptr; h->d_keep = nullptr;
        const int c = cap > h->stage_cap ? cap : h->stage_cap, b = batch > h->stage_batch ? batch : h->stage_batch;
        h->stage_cap = h->stage_batch = 0;
        SB_CUDA(cudaMalloc((void **)&h->d_kps_out, (size_t)b * c * sizeof(sb_keypoint)));
        SB_CUDA(cudaMalloc((void **)&h->d_desc_out, (size_t)b * c * 32));
        SB_CUDA(cudaMalloc((void **)&h->d_counts_out, (size_t)b * 4));
        SB_CUDA(cudaMalloc((void **)&h->d_keep, (size_t)b * c));
        h->stage_cap = c;
        h->stage_batch = b;
    }
    return SB_OK;
}

// Host images -> h->d_in (one slot of align16(w) * hgt bytes per image).  A tightly packed image (stride == w) goes up as ONE
// contiguous copy and keeps its row length on the device — k_copy_level0 re-pitches unaligned rows anyway — because a 2-D
// copy of 1241-byte rows runs at a fraction of the link rate; other strides take the 2-D copy into rows of align16(w).
// *dstride / *dmstride: the row length the device-side entry points have to be given.
static int stage_images(sb_orb *h, int batch, const uint8_t *const *img, const uint8_t *const *mask, int w, int hgt,
                        int stride, int mstride, bool *any_mask, int *dstride, int *dmstride = nullptr) {
    const size_t img_bytes = (size_t)sb_align_up((size_t)w, 16) * hgt;
    *any_mask = false;
    if (mask)
        for (int b = 0; b < batch; b++)
            if (mask[b]) *any_mask = true;
    for (int b = 0; b < batch; b++) SB_REQUIRE(img[b], "null image pointer");
    const size_t row = sb_align_up((size_t)w, 16);
    const bool flat = stride == w, mflat = mstride == w;
    *dstride = flat ? w : (int)row;
    if (dmstride) *dmstride = mflat ? w : (int)row;
    for (int b = 0; b < batch; b++) {
        if (flat)
            SB_CUDA(cudaMemcpyAsync(h->d_in + b * img_bytes, img[b], (size_t)w * hgt, cudaMemcpyHostToDevice, h->stream));
        else
            SB_CUDA(cudaMemcpy2DAsync(h->d_in + b * img_bytes, row, img[b], (size_t)stride, (size_t)w, (size_t)hgt,
                                      cudaMemcpyHostToDevice, h->stream));
        if (*any_mask) {
            if (mask[b] && mflat)
                SB_CUDA(cudaMemcpyAsync(h->d_in_mask + b * img_bytes, mask[b], (size_t)w * hgt, cudaMemcpyHostToDevice, h->stream));
            else if (mask[b])
                SB_CUDA(cudaMemcpy2DAsync(h->d_in_mask + b * img_bytes, row, mask[b], (size_t)mstride, (size_t)w,
                                          (size_t)hgt, cudaMemcpyHostToDevice, h->stream));
            else
                SB_CUDA(cudaMemsetAsync(h->d_in_mask + b * img_bytes, 255, img_bytes, h->stream));
        }
    }
    return SB_OK;
}

extern "C" int sb_orb_detect_and_compute(sb_orb_t *h, int batch, const uint8_t *const *img, const uint8_t *const *mask,
                                         int w, int hgt, int stride, int mstride, sb_keypoint *kps, uint8_t *desc,
                                         int32_t *counts, int cap) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_TRY(check_shapes(h, batch, w, hgt, stride, cap));
    SB_REQUIRE(img && kps && counts, "null pointer");
    const int row = (int)sb_align_up((size_t)w, 16);
    const size_t img_bytes = (size_t)row * hgt;
    bool any_mask = false;
    if (mask)
        for (int b = 0; b < batch; b++)
            if (mask[b]) any_mask = true;
    SB_REQUIRE(!any_mask || mstride >= w, "bad mask stride");
    SB_TRY(ensure_staging(h, batch, img_bytes, cap, any_mask));
    int ds = row, dms = row;
    SB_TRY(stage_images(h, batch, img, mask, w, hgt, stride, mstride, &any_mask, &ds, &dms));
    SB_TRY(sb_orb_detect_and_compute_dev(h, batch, h->d_in, (int64_t)img_bytes, any_mask ? h->d_in_mask : nullptr,
                                         (int64_t)img_bytes, w, hgt, ds, dms, h->d_kps_out, desc ? h->d_desc_out : nullptr,
                                         h->d_counts_out, cap));
    SB_CUDA(cudaMemcpyAsync(kps, h->d_kps_out, (size_t)batch * cap * sizeof(sb_keypoint), cudaMemcpyDeviceToHost, h->stream));
    if (desc) SB_CUDA(cudaMemcpyAsync(desc, h->d_desc_out, (size_t)batch * cap * 32, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(counts, h->d_counts_out, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    return finish_and_check(h);
}

extern "C" int sb_orb_detect(sb_orb_t *h, int batch, const uint8_t *const *img, const uint8_t *const *mask, int w,
                             int hgt, int stride, int mstride, sb_keypoint *kps, int32_t *counts, int cap) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_TRY(check_shapes(h, batch, w, hgt, stride, cap));
    SB_REQUIRE(img && kps && counts, "null pointer");
    const int row = (int)sb_align_up((size_t)w, 16);
    const size_t img_bytes = (size_t)row * hgt;
    bool any_mask = false;
    if (mask)
        for (int b = 0; b < batch; b++)
            if (mask[b]) any_mask = true;
    SB_REQUIRE(!any_mask || mstride >= w, "bad mask stride");
    SB_TRY(ensure_staging(h, batch, img_bytes, cap, any_mask));
    int ds = row, dms = row;
    SB_TRY(stage_images(h, batch, img, mask, w, hgt, stride, mstride, &any_mask, &ds, &dms));
    SB_TRY(sb_orb_detect_dev(h, batch, h->d_in, (int64_t)img_bytes, any_mask ? h->d_in_mask : nullptr, (int64_t)img_bytes, w,
                             hgt, ds, dms, h->d_kps_out, h->d_counts_out, cap));
    SB_CUDA(cudaMemcpyAsync(kps, h->d_kps_out, (size_t)batch * cap * sizeof(sb_keypoint), cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(counts, h->d_counts_out, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    return finish_and_check(h);
}

extern "C" int sb_orb_screen_params(sb_orb_t *h, const uint8_t *img, int w, int hgt, int stride, sb_keypoint *in,
                                    int n_in, sb_keypoint *out, int32_t *n_out) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(n_out, "null n_out");
    *n_out = 0;
    if (n_in <= 0) return SB_OK;
    SB_TRY(check_shapes(h, 1, w, hgt, stride, n_in));
    SB_REQUIRE(img && in && out, "null pointer");
    const int row = (int)sb_align_up((size_t)w, 16);
    const size_t img_bytes = (size_t)row * hgt;
    SB_TRY(ensure_staging(h, 1, img_bytes, n_in, false));
    const uint8_t *imgs[1] = {img};
    bool any = false;
    int ds = row;
    SB_TRY(stage_images(h, 1, imgs, nullptr, w, hgt, stride, 0, &any, &ds));
    SB_TRY(launch_pyramid(h, h->d_pyr, h->d_in, (long long)img_bytes, ds, 1, h->nlevels));
    SB_CUDA(cudaMemcpyAsync(h->d_kps_out, in, (size_t)n_in * sizeof(sb_keypoint), cudaMemcpyHostToDevice, h->stream));
    k_screen<<<sb_div_up(n_in, DESC_WARPS), DESC_WARPS * 32, 0, h->stream>>>(h->geom, h->d_pyr, h->d_kps_out, n_in, h->d_keep,
                                                                           h->minTh, 0, nullptr);
    SB_CUDA(cudaGetLastError());
    std::vector<uint8_t> keep((size_t)n_in);
    SB_CUDA(cudaMemcpyAsync(in, h->d_kps_out, (size_t)n_in * sizeof(sb_keypoint), cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(keep.data(), h->d_keep, (size_t)n_in, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    int m = 0;
    for (int i = 0; i < n_in; i++)
        if (keep[i]) out[m++] = in[i];  // survivors in input order (:1125)
    *n_out = m;
    return SB_OK;
}

extern "C" int sb_orb_calc_descriptors(sb_orb_t *h, const uint8_t *img, int w, int hgt, int stride,
                                       const sb_keypoint *kps, int n, uint8_t *desc) {
    SB_NVTX_FN();
    sb_clear_error();
    if (n <= 0) return SB_OK;
    SB_TRY(check_shapes(h, 1, w, hgt, stride, n));
    SB_REQUIRE(img && kps && desc, "null pointer");
    for (int i = 0; i < n; i++) {
        SB_REQUIRE(kps[i].octave >= 0 && kps[i].octave < h->nlevels, "keypoint octave out of range");
        // the reference reads out of bounds for keypoints closer than the rBRIEF reach to the border
        const LevelGeom &L = h->geom.lv[kps[i].octave];
        const float x = kps[i].x / L.scale, y = kps[i].y / L.scale;
        SB_REQUIRE(x >= 19.f && y >= 19.f && x < (float)(L.w - 19) && y < (float)(L.h - 19),
                   "keypoint closer than 19 px to the border of its pyramid level");
    }
    const int row = (int)sb_align_up((size_t)w, 16);
    const size_t img_bytes = (size_t)row * hgt;
    SB_TRY(ensure_staging(h, 1, img_bytes, n, false));
    const uint8_t *imgs[1] = {img};
    bool any = false;
    int ds = row;
    SB_TRY(stage_images(h, 1, imgs, nullptr, w, hgt, stride, 0, &any, &ds));
    SB_TRY(launch_pyramid(h, h->d_pyr, h->d_in, (long long)img_bytes, ds, 1, h->nlevels));
    SB_TRY(launch_blur(h, 1, h->stream));
    SB_CUDA(cudaMemcpyAsync(h->d_kps_out, kps, (size_t)n * sizeof(sb_keypoint), cudaMemcpyHostToDevice, h->stream));
    k_calc_desc<<<sb_div_up(n, DESC_WARPS), DESC_WARPS * 32, 0, h->stream>>>(h->geom, h->d_blur, h->d_kps_out, n, h->d_desc_out, 0, nullptr);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaMemcpyAsync(desc, h->d_desc_out, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    return SB_OK;
}

// LoopClosing::ProcessNewKF's two extractor calls (src/loopclosing.cpp:107-112) for a BATCH of keyframe images in one pass:
// ScreenAndComputeKPsParams (:1083-1129) and then CalcDescriptors (:1180-1226) on its survivors — one staging copy, one pyramid
// and one blur per image instead of the reference's two pyramids (quirk Q7), one screening and one descriptor launch per batch.
//   in  [batch][cap_in] (first n_in[b] live; mutated like the reference's vector, quirk Q5)
//   out [batch][cap_in] survivors in input order, n_out [batch], desc [batch][cap_in][32] (row k <-> out[k])
extern "C" int sb_orb_screen_describe(sb_orb_t *h, int batch, const uint8_t *const *img, int w, int hgt, int stride, sb_keypoint *in,
                                      const int32_t *n_in, int cap_in, sb_keypoint *out, int32_t *n_out, uint8_t *desc) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(img && in && n_in && out && n_out && desc, "null pointer");
    SB_TRY(check_shapes(h, batch, w, hgt, stride, cap_in));
    int total = 0;
    for (int b = 0; b < batch; b++) {
        SB_REQUIRE(n_in[b] >= 0 && n_in[b] <= cap_in, "n_in out of range [0, cap_in]");
        n_out[b] = 0;
        total += n_in[b];
    }
    if (total == 0) return SB_OK;
    const int row = (int)sb_align_up((size_t)w, 16);
    const size_t img_bytes = (size_t)row * hgt;
    SB_TRY(ensure_staging(h, batch, img_bytes, cap_in, false));
    bool any = false;
    int ds = row;
    SB_TRY(stage_images(h, batch, img, nullptr, w, hgt, stride, 0, &any, &ds));
    SB_TRY(launch_pyramid(h, h->d_pyr, h->d_in, (long long)img_bytes, ds, batch, h->nlevels));
    SB_TRY(launch_blur(h, batch, h->stream));
    const int n = batch * cap_in;
    SB_CUDA(cudaMemcpyAsync(h->d_kps_out, in, (size_t)n * sizeof(sb_keypoint), cudaMemcpyHostToDevice, h->stream));
    SB_CUDA(cudaMemcpyAsync(h->d_counts_out, n_in, (size_t)batch * 4, cudaMemcpyHostToDevice, h->stream));
    SB_CUDA(cudaMemsetAsync(h->d_keep, 0, (size_t)n, h->stream));
    k_screen<<<sb_div_up(n, DESC_WARPS), DESC_WARPS * 32, 0, h->stream>>>(h->geom, h->d_pyr, h->d_kps_out, n, h->d_keep, h->minTh, cap_in,
                                                                           h->d_counts_out);
    SB_CUDA(cudaGetLastError());
    k_calc_desc<<<sb_div_up(n, DESC_WARPS), DESC_WARPS * 32, 0, h->stream>>>(h->geom, h->d_blur, h->d_kps_out, n, h->d_desc_out, cap_in,
                                                                              h->d_keep);
    SB_CUDA(cudaGetLastError());
    // results come back through page-locked staging (a copy into pageable memory runs at a fraction of the link rate and
    // blocks): [keypoints n x 28][keep flags n][descriptors n x 32]
    const size_t off_keep = sb_align_up((size_t)n * sizeof(sb_keypoint), 256), off_desc = off_keep + sb_align_up((size_t)n, 256);
    const size_t need = off_desc + (size_t)n * 32;
    if (need > h->h_stage_cap) {
        SB_CUDA(cudaStreamSynchronize(h->stream));
        if (h->h_stage) cudaFreeHost(h->h_stage);
        h->h_stage = nullptr;
        h->h_stage_cap = 0;
        SB_CUDA(cudaHostAlloc((void **)&h->h_stage, need, cudaHostAllocPortable));
        h->h_stage_cap = need;
    }
    const uint8_t *keep = h->h_stage + off_keep, *dtmp = h->h_stage + off_desc;
    SB_CUDA(cudaMemcpyAsync(h->h_stage, h->d_kps_out, (size_t)n * sizeof(sb_keypoint), cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(h->h_stage + off_keep, h->d_keep, (size_t)n, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(h->h_stage + off_desc, h->d_desc_out, (size_t)n * 32, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    memcpy(in, h->h_stage, (size_t)n * sizeof(sb_keypoint));   // the reference mutates its input keypoints (:1100-1121)
    for (int b = 0; b < batch; b++) {
        int m = 0;
        for (int i = 0; i < n_in[b]; i++) {
            const size_t k = (size_t)b * cap_in + i;
            if (!keep[k]) continue;
            out[(size_t)b * cap_in + m] = in[k];   // survivors in input order (:1125)
            memcpy(desc + ((size_t)b * cap_in + m) * 32, dtmp + k * 32, 32);
            m++;
        }
        n_out[b] = m;
    }
    return SB_OK;
}

// ---- inspection --------------------------------------------------------------------------------------
extern "C" int sb_orb_debug_level(sb_orb_t *h, int b, int level, int which, uint8_t *out, int out_bytes, int *lw,
                                  int *lh) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && h->cur_w > 0, "no call has been made on this handle yet");
    SB_REQUIRE(b >= 0 && b < h->max_batch && level >= 0 && level < h->nlevels && which >= 0 && which <= 2, "bad index");
    SB_TRY(sb_use_device(h->device));
    const LevelGeom &L = h->geom.lv[level];
    if (lw) *lw = L.w;
    if (lh) *lh = L.h;
    if (!out) return SB_OK;
    SB_REQUIRE(out_bytes >= L.w * L.h, "output buffer too small");
    const uint8_t *src = which == 0 ? h->d_pyr : which == 1 ? h->d_blur : h->d_mask;
    SB_REQUIRE(src, "that pyramid has not been built");
    SB_CUDA(cudaStreamSynchronize(h->stream));
    SB_CUDA(cudaMemcpy2D(out, (size_t)L.w, src + (long long)b * h->geom.slab + L.off, (size_t)L.pitch, (size_t)L.w,
                         (size_t)L.h, cudaMemcpyDeviceToHost));
    return SB_OK;
}

extern "C" int sb_orb_debug_candidates(sb_orb_t *h, int b, int level, uint32_t *out, int cap, int32_t *n) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && out && n, "null pointer");
    SB_REQUIRE(b >= 0 && b < h->max_batch && level >= 0 && level < h->nlevels, "bad index");
    SB_TRY(sb_use_device(h->device));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    int cnt = 0;
    SB_CUDA(cudaMemcpy(&cnt, h->d_cand_cnt + b * h->nlevels + level, 4, cudaMemcpyDeviceToHost));
    if (cnt > SB_CAND_CAP) cnt = SB_CAND_CAP;
    *n = cnt;
    const int m = cnt < cap ? cnt : cap;
    if (m > 0)
        SB_CUDA(cudaMemcpy(out, h->d_cand + ((long long)b * h->nlevels + level) * SB_CAND_CAP, (size_t)m * 4,
                           cudaMemcpyDeviceToHost));
    return SB_OK;
}

// Inspection: arm (buf != null) or disarm the capture of one FAST cell of image 0; after the next call
// `out` (2 * tile_bytes + 64 bytes) holds the TMA tile, the response plane and 10 ints of geometry.
extern "C" int sb_orb_debug_fast_cell(sb_orb_t *h, int cell, uint8_t *out, int out_bytes, int *tile_bytes) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_TRY(sb_use_device(h->device));
    if (tile_bytes) *tile_bytes = h->fast_tile_bytes;
    const int need = 2 * h->fast_tile_bytes + 64;
    if (!out) {  // arm
        if (!h->dbg_buf) SB_CUDA(cudaMalloc((void **)&h->dbg_buf, (size_t)need));
        SB_CUDA(cudaMemset(h->dbg_buf, 0, (size_t)need));
        h->dbg_cell = cell;
        return SB_OK;
    }
    SB_REQUIRE(h->dbg_buf && out_bytes >= need, "not armed or buffer too small");
    SB_CUDA(cudaStreamSynchronize(h->stream));
    SB_CUDA(cudaMemcpy(out, h->dbg_buf, (size_t)need, cudaMemcpyDeviceToHost));
    cudaFree(h->dbg_buf);
    h->dbg_buf = nullptr;
    return SB_OK;
}

// Per-stage timing with CUDA events on the launching streams.  enable != 0 starts a fresh recording.
extern "C" int sb_orb_profile(sb_orb_t *h, int enable) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_TRY(sb_use_device(h->device));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    SB_CUDA(cudaStreamSynchronize(h->side_stream));
    h->prof_on = enable != 0;
    if (enable) h->prof_n = 0;
    return SB_OK;
}

// Sums the recorded intervals per stage (SB_ORB_STAGE_*): ms[nstages], launches[nstages].
extern "C" int sb_orb_profile_read(sb_orb_t *h, float *ms, int32_t *launches, int nstages) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && ms && launches && nstages >= SB_STAGE_COUNT, "bad arguments");
    SB_TRY(sb_use_device(h->device));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    SB_CUDA(cudaStreamSynchronize(h->side_stream));
    for (int i = 0; i < nstages; i++) { ms[i] = 0.f; launches[i] = 0; }
    for (int i = 0; i < h->prof_n; i++) {
        float t = 0.f;
        SB_CUDA(cudaEventElapsedTime(&t, h->prof_ev[i][0], h->prof_ev[i][1]));
        ms[h->prof_stage[i]] += t;
        launches[h->prof_stage[i]] += h->prof_launches[i];
    }
    return SB_OK;
}
