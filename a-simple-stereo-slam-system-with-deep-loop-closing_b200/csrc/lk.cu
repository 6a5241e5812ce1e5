// lk.cu — batched pyramidal Lucas-Kanade tracking on B200 (sm_100a).
//
// SURVEY §8(f) "next" row 1a — what the live front end spends every frame on: replaces
// cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(11,11), 3, (COUNT+EPS, 30, 0.01),
// OPTFLOW_USE_INITIAL_FLOW) as called by Frontend::TrackLastFrame (reference src/frontend.cpp:150-153) and
// Frontend::FindFeaturesInRight (:358-361).  OpenCV's algorithm (pyrDown, calcScharrDeriv, LKTrackerInvoker) is
// followed step by step; see oracle/lk_oracle.c for the one deliberate difference (the five sums of products are
// accumulated exactly in 64-bit integers instead of in SIMD-ordered fp32) and for the cv2 pin.
// One warp tracks one point: the 121 window pixels are spread over the lanes, the integer patch (I, Ix, Iy) stays in
// shared memory for all iterations, sums are warp shuffles of int64.  All fp32 arithmetic is un-contracted
// (__fmul_rn / __fadd_rn), so the result is bit-identical to the oracle.
#include <math.h>
#include <string.h>

#include "common.cuh"

#define LK_MAX_LEVELS 8
#define LK_WARPS 4
#define LK_WBITS 14
#define LK_DESCALE(x, n) (((x) + (1 << ((n)-1))) >> (n))

struct LkLevel { int w, h, pitch; long long off, doff; };  // image plane offset (bytes), derivative plane offset (short2 elements)

struct LkGeom {
    LkLevel lv[LK_MAX_LEVELS + 1];
    int levels;          // number of pyramid levels actually built (maxLevel + 1)
    long long slab, dslab;
};

struct sb_lk {
    int device, max_w, max_h, max_batch, max_pts, max_level;
    cudaStream_t stream, own_stream;
    LkGeom geom;
    int cur_w, cur_h;
    uint8_t *d_pyr;      // [2 * batch][slab]: prev images then next images
    short2 *d_deriv;     // [batch][dslab]
    uint8_t *d_in;       // staging for host images [2 * batch][h][row]
    float *d_prev, *d_next;
    uint8_t *d_status;
    int32_t *d_n;
    long long slab_cap, dslab_cap;
};

static __device__ __forceinline__ int lk_reflect(int p, int len) {
    if (p < 0) p = -p;
    if (p >= len) p = 2 * (len - 1) - p;
    return p;
}

// level-0 copy with re-pitching (any source stride)
__global__ void k_lk_copy0(const uint8_t *__restrict__ src, long long src_img_pitch, int stride, int w, int h, uint8_t *__restrict__ dst,
                           long long slab, int pitch, int n_img) {
    const long long total = (long long)n_img * h * pitch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % pitch);
        const long long r = i / pitch;
        const int y = (int)(r % h), b = (int)(r / h);
        dst[b * slab + (long long)y * pitch + x] = x < w ? src[b * src_img_pitch + (long long)y * stride + x] : 0;
    }
}

// cv::pyrDown: [1 4 6 4 1]^2 / 256, BORDER_REFLECT_101, one thread per destination pixel
__global__ void k_lk_pyrdown(uint8_t *__restrict__ pyr, long long slab, LkLevel S, LkLevel D) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= D.w) return;
    const uint8_t *src = pyr + (long long)blockIdx.z * slab + S.off;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; k++) xs[k] = lk_reflect(2 * x + k - 2, S.w);
    int acc = 0;
    const int wk[5] = {1, 4, 6, 4, 1};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint8_t *r = src + (long long)lk_reflect(2 * y + k - 2, S.h) * S.pitch;
        acc += wk[k] * (r[xs[0]] + r[xs[4]] + 4 * (r[xs[1]] + r[xs[3]]) + 6 * r[xs[2]]);
    }
    pyr[(long long)blockIdx.z * slab + D.off + (long long)y * D.pitch + x] = (uint8_t)((acc + 128) >> 8);
}

// calcScharrDeriv of the prev image at one level
__global__ void k_lk_scharr(const uint8_t *__restrict__ pyr, long long slab, short2 *__restrict__ deriv, long long dslab, LkLevel L) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= L.w) return;
    const uint8_t *img = pyr + (long long)blockIdx.z * slab + L.off;
    const uint8_t *r0 = img + (long long)lk_reflect(y - 1, L.h) * L.pitch, *r1 = img + (long long)y * L.pitch;
    const uint8_t *r2 = img + (long long)lk_reflect(y + 1, L.h) * L.pitch;
    const int xm = lk_reflect(x - 1, L.w), xp = lk_reflect(x + 1, L.w);
    const int t0m = (r0[xm] + r2[xm]) * 3 + r1[xm] * 10, t0p = (r0[xp] + r2[xp]) * 3 + r1[xp] * 10;
    const int t1m = r2[xm] - r0[xm], t1c = r2[x] - r0[x], t1p = r2[xp] - r0[xp];
    deriv[(long long)blockIdx.z * dslab + L.doff + (long long)y * L.w + x] = make_short2((short)(t0p - t0m), (short)((t1p + t1m) * 3 + t1c * 10));
}

struct LkArgs {
    const uint8_t *pyr;   // prev images at [b], next images at [batch + b]
    const short2 *deriv;
    long long slab, dslab;
    LkLevel L;
    int batch, max_pts, level, max_level, win, max_count, use_initial;
    float min_eig_th;
    double eps2;
    const int32_t *n;
    const float *prev;    // [batch][max_pts][2]
    float *next;          // [batch][max_pts][2] in/out
    uint8_t *status;      // [batch][max_pts]
};

static __device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

static __device__ __forceinline__ void lk_weights(float a, float b, int &w00, int &w01, int &w10, int &w11) {
    const float oa = __fsub_rn(1.f, a), ob = __fsub_rn(1.f, b), s = (float)(1 << LK_WBITS);
    w00 = __float2int_rn(__fmul_rn(__fmul_rn(oa, ob), s));
    w01 = __float2int_rn(__fmul_rn(__fmul_rn(a, ob), s));
    w10 = __float2int_rn(__fmul_rn(__fmul_rn(oa, b), s));
    w11 = (1 << LK_WBITS) - w00 - w01 - w10;
}

// LKTrackerInvoker for one level: one warp per point
__global__ void __launch_bounds__(LK_WARPS * 32) k_lk_level(const __grid_constant__ LkArgs a) {
    extern __shared__ short s_win[];  // [LK_WARPS][win * win][3]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y, i = blockIdx.x * LK_WARPS + warp;
    if (i >= min(a.n[b], a.max_pts)) return;
    const int win = a.win, nwin = win * win;
    short *Iw = s_win + (size_t)warp * nwin * 3;
    const LkLevel &L = a.L;
    const uint8_t *I = a.pyr + (long long)b * a.slab + L.off, *J = a.pyr + (long long)(a.batch + b) * a.slab + L.off;
    const short2 *dI = a.deriv + (long long)b * a.dslab + L.doff;
    const long long pi = (long long)b * a.max_pts + i;
    const float half = __fmul_rn((float)(win - 1), 0.5f), FLT_SCALE = 1.f / (1 << 20);
    const float sc = (float)(1. / (1 << a.level));
    float ppx = __fmul_rn(a.prev[2 * pi], sc), ppy = __fmul_rn(a.prev[2 * pi + 1], sc), nx, ny;
    if (a.level == a.max_level) {
        if (a.use_initial) { nx = __fmul_rn(a.next[2 * pi], sc); ny = __fmul_rn(a.next[2 * pi + 1], sc); }
        else { nx = ppx; ny = ppy; }
        if (lane == 0) a.status[pi] = 1;
    } else {
        nx = __fmul_rn(a.next[2 * pi], 2.f); ny = __fmul_rn(a.next[2 * pi + 1], 2.f);
    }
    __syncwarp();
    if (lane == 0) { a.next[2 * pi] = nx; a.next[2 * pi + 1] = ny; }
    ppx = __fsub_rn(ppx, half); ppy = __fsub_rn(ppy, half);
    const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
    if (ipx < -win || ipx >= L.w || ipy < -win || ipy >= L.h) {
        if (a.level == 0 && lane == 0) a.status[pi] = 0;
        return;
    }
    int w00, w01, w10, w11;
    lk_weights(__fsub_rn(ppx, (float)ipx), __fsub_rn(ppy, (float)ipy), w00, w01, w10, w11);
    long long sA11 = 0, sA12 = 0, sA22 = 0;
    for (int k = lane; k < nwin; k += 32) {
        const int y = k / win, x = k - y * win;
        const int X = ipx + x, Y = ipy + y;
        const int x0 = lk_reflect(X, L.w), x1 = lk_reflect(X + 1, L.w);
        const uint8_t *r0 = I + (long long)lk_reflect(Y, L.h) * L.pitch, *r1 = I + (long long)lk_reflect(Y + 1, L.h) * L.pitch;
        const int ival = LK_DESCALE(r0[x0] * w00 + r0[x1] * w01 + r1[x0] * w10 + r1[x1] * w11, LK_WBITS - 5);
        // derivatives are zero outside the image (BORDER_CONSTANT)
        const bool xin0 = X >= 0 && X < L.w, xin1 = X + 1 >= 0 && X + 1 < L.w, yin0 = Y >= 0 && Y < L.h, yin1 = Y + 1 >= 0 && Y + 1 < L.h;
        const short2 z = make_short2(0, 0);
        const short2 d00 = xin0 && yin0 ? dI[(long long)Y * L.w + X] : z, d01 = xin1 && yin0 ? dI[(long long)Y * L.w + X + 1] : z;
        const short2 d10 = xin0 && yin1 ? dI[(long long)(Y + 1) * L.w + X] : z, d11 = xin1 && yin1 ? dI[(long long)(Y + 1) * L.w + X + 1] : z;
        const int ix = (short)LK_DESCALE(d00.x * w00 + d01.x * w01 + d10.x * w10 + d11.x * w11, LK_WBITS);
        const int iy = (short)LK_DESCALE(d00.y * w00 + d01.y * w01 + d10.y * w10 + d11.y * w11, LK_WBITS);
        Iw[3 * k] = (short)ival; Iw[3 * k + 1] = (short)ix; Iw[3 * k + 2] = (short)iy;
        sA11 += (long long)ix * ix; sA12 += (long long)ix * iy; sA22 += (long long)iy * iy;
    }
    __syncwarp();
    const float A11 = __fmul_rn(__ll2float_rn(warp_sum_ll(sA11)), FLT_SCALE), A12 = __fmul_rn(__ll2float_rn(warp_sum_ll(sA12)), FLT_SCALE);
    const float A22 = __fmul_rn(__ll2float_rn(warp_sum_ll(sA22)), FLT_SCALE);
    float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
    const float dA = __fsub_rn(A11, A22);
    const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), sqrtf(__fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
                                   (float)(2 * win * win));
    if (minEig < a.min_eig_th || D < 1.1920929e-07f) {
        if (a.level == 0 && lane == 0) a.status[pi] = 0;
        return;
    }
    D = __fdiv_rn(1.f, D);
    nx = __fsub_rn(nx, half); ny = __fsub_rn(ny, half);
    float pdx = 0.f, pdy = 0.f;
    for (int j = 0; j < a.max_count; j++) {
        const int inx = (int)floorf(nx), iny = (int)floorf(ny);
        if (inx < -win || inx >= L.w || iny < -win || iny >= L.h) {
            if (a.level == 0 && lane == 0) a.status[pi] = 0;
            break;
        }
        lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
        long long sb1 = 0, sb2 = 0;
        for (int k = lane; k < nwin; k += 32) {
            const int y = k / win, x = k - y * win;
            const int X = inx + x, Y = iny + y;
            const int x0 = lk_reflect(X, L.w), x1 = lk_reflect(X + 1, L.w);
            const uint8_t *r0 = J + (long long)lk_reflect(Y, L.h) * L.pitch, *r1 = J + (long long)lk_reflect(Y + 1, L.h) * L.pitch;
            const int diff = LK_DESCALE(r0[x0] * w00 + r0[x1] * w01 + r1[x0] * w10 + r1[x1] * w11, LK_WBITS - 5) - Iw[3 * k];
            sb1 += (long long)diff * Iw[3 * k + 1];
            sb2 += (long long)diff * Iw[3 * k + 2];
        }
        const float b1 = __fmul_rn(__ll2float_rn(warp_sum_ll(sb1)), FLT_SCALE), b2 = __fmul_rn(__ll2float_rn(warp_sum_ll(sb2)), FLT_SCALE);
        const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
        const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
        nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
        float ox = __fadd_rn(nx, half), oy = __fadd_rn(ny, half);
        bool stop = __dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= a.eps2;
        if (!stop && j > 0 && fabsf(__fadd_rn(dx, pdx)) < 0.01 && fabsf(__fadd_rn(dy, pdy)) < 0.01) {
            ox = __fsub_rn(ox, __fmul_rn(dx, 0.5f)); oy = __fsub_rn(oy, __fmul_rn(dy, 0.5f));
            stop = true;
        }
        if (lane == 0) { a.next[2 * pi] = ox; a.next[2 * pi + 1] = oy; }
        if (stop) break;
        pdx = dx; pdy = dy;
    }
}

// ================================================================================================
// host side
// ================================================================================================
static void free_lk(sb_lk *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_pyr, h->d_deriv, h->d_in, h->d_prev, h->d_next, h->d_status, h->d_n};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

static void lk_geometry(int w, int hgt, int max_level, int win, LkGeom *g) {
    memset(g, 0, sizeof(*g));
    long long off = 0, doff = 0;
    int lw = w, lh = hgt;
    for (int l = 0; l <= max_level; l++) {
        if (l > 0) {
            lw = (lw + 1) / 2; lh = (lh + 1) / 2;
            if (lw <= win || lh <= win) break;  // buildOpticalFlowPyramid stops here
        }
        LkLevel &L = g->lv[l];
        L.w = lw; L.h = lh; L.pitch = (int)sb_align_up(lw, 16); L.off = off; L.doff = doff;
        off += (long long)sb_align_up((size_t)L.pitch * lh, 256);
        doff += (long long)lw * lh;
        g->levels = l + 1;
    }
    g->slab = off;
    g->dslab = doff;
}

extern "C" int sb_lk_create(sb_lk_t **out, int device, int max_w, int max_h, int max_batch, int max_pts, int max_level) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_w >= 16 && max_w <= 8192 && max_h >= 16 && max_h <= 8192, "max_w / max_h out of range [16, 8192]");
    SB_REQUIRE(max_batch >= 1 && max_batch <= 4096 && max_pts >= 1 && max_pts <= (1 << 20), "max_batch / max_pts out of range");
    SB_REQUIRE(max_level >= 0 && max_level <= LK_MAX_LEVELS, "max_level out of range [0, 8]");
    SB_TRY(sb_use_device(device));
    sb_lk *h = new sb_lk();
    memset(h, 0, sizeof(*h));
    h->device = device; h->max_w = max_w; h->max_h = max_h; h->max_batch = max_batch; h->max_pts = max_pts; h->max_level = max_level;
    LkGeom g;
    lk_geometry(max_w, max_h, max_level, 1, &g);
    h->slab_cap = g.slab + 4096;
    h->dslab_cap = g.dslab + 64;
    const size_t B = max_batch;
    cudaError_t e = cudaMalloc((void **)&h->d_pyr, 2 * B * h->slab_cap);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_deriv, B * h->dslab_cap * sizeof(short2));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_in, 2 * B * sb_align_up((size_t)max_w, 16) * max_h);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_prev, B * max_pts * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_next, B * max_pts * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_status, B * max_pts);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_n, B * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        sb_set_error("sb_lk_create: %s", cudaGetErrorString(e));
        free_lk(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    h->cur_w = h->cur_h = -1;
    *out = h;
    return SB_OK;
}

extern "C" int sb_lk_destroy(sb_lk_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_lk(h);
    }
    return SB_OK;
}

extern "C" int sb_lk_set_stream(sb_lk_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

// d_prev_img / d_next_img: batch images each, image b at base + b * img_pitch_bytes, rows `stride` bytes.
extern "C" int sb_lk_track_dev(sb_lk_t *h, int batch, const uint8_t *d_prev_img, const uint8_t *d_next_img, int64_t img_pitch_bytes,
                               int w, int hgt, int stride, const int32_t *d_n_pts, const float *d_prev_pts, float *d_next_pts,
                               uint8_t *d_status, int win, int max_count, double eps, int use_initial_flow, float min_eig_th) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && d_prev_img && d_next_img && d_n_pts && d_prev_pts && d_next_pts && d_status, "null pointer");
    SB_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch out of range [1, max_batch]");
    SB_REQUIRE(w <= h->max_w && hgt <= h->max_h && stride >= w, "image larger than max_w x max_h or bad stride");
    SB_REQUIRE(win >= 3 && win <= 21 && (win & 1) && w > win && hgt > win, "window must be odd, 3..21, and smaller than the image");
    SB_TRY(sb_use_device(h->device));
    if (max_count < 0) max_count = 0;
    if (max_count > 100) max_count = 100;
    if (eps < 0) eps = 0;
    if (eps > 10) eps = 10;
    LkGeom g;
    lk_geometry(w, hgt, h->max_level, win, &g);
    cudaStream_t s = h->stream;
    const LkLevel &L0 = g.lv[0];
    const long long total = (long long)batch * L0.h * L0.pitch;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    k_lk_copy0<<<blocks, 256, 0, s>>>(d_prev_img, img_pitch_bytes, stride, w, hgt, h->d_pyr, g.slab, L0.pitch, batch);
    k_lk_copy0<<<blocks, 256, 0, s>>>(d_next_img, img_pitch_bytes, stride, w, hgt, h->d_pyr + (long long)batch * g.slab, g.slab, L0.pitch, batch);
    for (int l = 1; l < g.levels; l++)
        k_lk_pyrdown<<<dim3(sb_div_up(g.lv[l].w, 128), g.lv[l].h, 2 * batch), 128, 0, s>>>(h->d_pyr, g.slab, g.lv[l - 1], g.lv[l]);
    for (int l = 0; l < g.levels; l++)
        k_lk_scharr<<<dim3(sb_div_up(g.lv[l].w, 128), g.lv[l].h, batch), 128, 0, s>>>(h->d_pyr, g.slab, h->d_deriv, g.dslab, g.lv[l]);
    LkArgs a;
    a.pyr = h->d_pyr; a.deriv = h->d_deriv; a.slab = g.slab; a.dslab = g.dslab;
    a.batch = batch; a.max_pts = h->max_pts; a.max_level = g.levels - 1; a.win = win; a.max_count = max_count;
    a.use_initial = use_initial_flow; a.min_eig_th = min_eig_th; a.eps2 = eps * eps;
    a.n = d_n_pts; a.prev = d_prev_pts; a.next = d_next_pts; a.status = d_status;
    const size_t smem = (size_t)LK_WARPS * win * win * 3 * sizeof(short);
    for (int l = g.levels - 1; l >= 0; l--) {
        a.L = g.lv[l];
        a.level = l;
        k_lk_level<<<dim3(sb_div_up(h->max_pts, LK_WARPS), batch), LK_WARPS * 32, smem, s>>>(a);
    }
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// Host-pointer form: prev / next: `batch` pointers to images; prev_pts / next_pts [batch][max_pts][2], status [batch][max_pts].
extern "C" int sb_lk_track(sb_lk_t *h, int batch, const uint8_t *const *prev, const uint8_t *const *next, int w, int hgt, int stride,
                           const int32_t *n_pts, const float *prev_pts, float *next_pts, uint8_t *status, int win, int max_count,
                           double eps, int use_initial_flow, float min_eig_th) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && prev && next && n_pts && prev_pts && next_pts && status, "null pointer");
    SB_REQUIRE(batch >= 1 && batch <= h->max_batch, "batch out of range [1, max_batch]");
    SB_REQUIRE(w <= h->max_w && hgt <= h->max_h && stride >= w, "image larger than max_w x max_h or bad stride");
    for (int b = 0; b < batch; b++) SB_REQUIRE(prev[b] && next[b] && n_pts[b] >= 0 && n_pts[b] <= h->max_pts, "null image or n_pts out of range");
    SB_TRY(sb_use_device(h->device));
    cudaStream_t s = h->stream;
    const size_t row = sb_align_up((size_t)w, 16), plane = row * hgt, P = (size_t)batch * h->max_pts;
    // a tightly packed image goes up as one contiguous copy and keeps its row length on the device (a 2-D copy of odd-length
    // rows runs at a fraction of the link rate); other strides are re-pitched to align16(w) by the copy
    const bool flat = stride == w;
    const int dstride = flat ? w : (int)row;
    for (int b = 0; b < batch; b++) {
        if (flat) {
            SB_CUDA(cudaMemcpyAsync(h->d_in + b * plane, prev[b], (size_t)w * hgt, cudaMemcpyHostToDevice, s));
            SB_CUDA(cudaMemcpyAsync(h->d_in + ((size_t)batch + b) * plane, next[b], (size_t)w * hgt, cudaMemcpyHostToDevice, s));
        } else {
            SB_CUDA(cudaMemcpy2DAsync(h->d_in + b * plane, row, prev[b], (size_t)stride, (size_t)w, (size_t)hgt, cudaMemcpyHostToDevice, s));
            SB_CUDA(cudaMemcpy2DAsync(h->d_in + ((size_t)batch + b) * plane, row, next[b], (size_t)stride, (size_t)w, (size_t)hgt, cudaMemcpyHostToDevice, s));
        }
    }
    SB_CUDA(cudaMemcpyAsync(h->d_n, n_pts, (size_t)batch * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_prev, prev_pts, P * 8, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_next, next_pts, P * 8, cudaMemcpyHostToDevice, s));
    SB_TRY(sb_lk_track_dev(h, batch, h->d_in, h->d_in + (size_t)batch * plane, (int64_t)plane, w, hgt, dstride, h->d_n, h->d_prev, h->d_next,
                           h->d_status, win, max_count, eps, use_initial_flow, min_eig_th));
    SB_CUDA(cudaMemcpyAsync(next_pts, h->d_next, P * 8, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(status, h->d_status, P, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}
