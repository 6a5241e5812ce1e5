// pnp.cu — PnP with RANSAC for the loop geometric verification (SURVEY §8f "next" row 4), sm_100a.
//
// Replaces cv::solvePnPRansac as LoopClosing::ComputeCorrectPose calls it (reference src/loopclosing.cpp:259-268:
// objectPoints = map points of the loop keyframe (cv::Point3f), imagePoints = matched keypoints of the current
// keyframe (cv::Point2f), K, no distortion, useExtrinsicGuess = false, iterationsCount = 100,
// reprojectionError = 5.991, confidence = 0.99, SOLVEPNP_ITERATIVE).  The result feeds OptimizeCurrentPose
// (pose-only LM, pose.cu).
//
// OpenCV's implementation is sequential (hypothesis -> score -> maybe stop) and its samples come from cv::RNG; the
// reference has no golden vectors for it and its outcome is not a deterministic function of the inputs alone.
// What is reproduced is the estimator: hypotheses from minimal samples, inliers = squared reprojection error
// <= reprojectionError^2, best = most inliers, then a least-squares refinement on the inliers of the best model
// (OpenCV: solvePnP(inliers, SOLVEPNP_ITERATIVE) = Levenberg-Marquardt on the reprojection error).  On B200 the
// hypotheses are evaluated in parallel:
//   one CTA per problem (loop candidate);
//   phase 1  one thread per hypothesis: 4 distinct points from a counter-based generator, Grunert P3P on three of
//            them (pnp_core.inl), the fourth picks among the up-to-4 solutions;
//   phase 2  one warp per hypothesis, lanes over the correspondences: inlier count, best = max count (ties: lowest
//            hypothesis index) through one packed atomicMax in shared memory;
//   phase 3  inlier mask of the best hypothesis;
//   phase 4  Levenberg-Marquardt on the inliers (left-multiplied se(3) increments, the Jacobian of
//            EdgeProjectionPoseOnly, include/myslam/g2o_types.h:63-102), block-wide deterministic reductions.
// All `iterations` hypotheses are evaluated (OpenCV stops early once `confidence` is reached: a subset of these).
#include <string.h>

#include "common.cuh"
#include "pnp_core.inl"
#include "se3.cuh"

#define PNP_THREADS 256
#define PNP_MAX_HYP 1024

struct sb_pnp {
    int device, max_problems, max_points;
    cudaStream_t stream, own_stream;
    float *d_obj, *d_img;
    int32_t *d_n, *d_info;
    double *d_pose, *d_rt;
    uint8_t *d_inlier;
};

struct PnpArgs {
    const int32_t *n;     // [P]
    const float *obj;     // [P][MP][3]
    const float *img;     // [P][MP][2]
    double *pose7;        // [P][7] qx qy qz qw tx ty tz
    double *rvec_tvec;    // [P][6]
    uint8_t *inlier;      // [P][MP]
    int32_t *info;        // [P][4]: found, inliers, hypotheses with a solution, refinement iterations
    int MP, iterations;
    double fx, fy, cx, cy, thr2;
    unsigned long long seed;
};

static __device__ __forceinline__ unsigned long long pnp_mix(unsigned long long z) {  // splitmix64 finaliser
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

static __device__ __forceinline__ bool pnp_project(const double *Rt, const float *X, double fx, double fy, double cx, double cy,
                                                   double &u, double &v) {
    const double x = X[0], y = X[1], z = X[2];
    const double xc = Rt[0] * x + Rt[1] * y + Rt[2] * z + Rt[9];
    const double yc = Rt[3] * x + Rt[4] * y + Rt[5] * z + Rt[10];
    const double zc = Rt[6] * x + Rt[7] * y + Rt[8] * z + Rt[11];
    if (!(zc > 1e-9)) return false;
    u = fx * xc / zc + cx;
    v = fy * yc / zc + cy;
    return true;
}

static __device__ double pnp_block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int k = 0; k < PNP_THREADS / 32; k++) s += red[k];
    return s;
}

__global__ void __launch_bounds__(PNP_THREADS) k_pnp_ransac(const __grid_constant__ PnpArgs a) {
    extern __shared__ __align__(16) double hyp[];  // [iterations][12]
    __shared__ unsigned long long s_best;
    __shared__ int s_valid;
    __shared__ double s_rt[12], s_rtb[12], s_H[36], s_g[6], s_dx[6], red[PNP_THREADS / 32];
    __shared__ int s_flag;
    const int p = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = min(a.n[p], a.MP);
    const float *obj = a.obj + (size_t)p * a.MP * 3, *img = a.img + (size_t)p * a.MP * 2;
    uint8_t *inl = a.inlier + (size_t)p * a.MP;
    int32_t *info = a.info + 4 * p;
    if (tid == 0) { s_best = 0ull; s_valid = 0; }
    for (int i = tid; i < n; i += PNP_THREADS) inl[i] = 0;
    __syncthreads();
    if (n < 4) {
        if (tid == 0) { info[0] = 0; info[1] = 0; info[2] = 0; info[3] = 0; }
        return;
    }
    // ---- phase 1: hypotheses
    for (int h = tid; h < a.iterations; h += PNP_THREADS) {
        double *out = hyp + 12 * h;
        out[0] = nan("");
        int idx[4];
        bool ok = true;
        unsigned long long st = pnp_mix(a.seed ^ ((unsigned long long)p << 40) ^ (unsigned long long)h);
        for (int k = 0; k < 4 && ok; k++) {
            int tries = 0;
            for (;;) {
                st = pnp_mix(st);
                const int c = (int)((st >> 11) % (unsigned long long)n);
                bool dup = false;
                for (int q = 0; q < k; q++) dup |= idx[q] == c;
                if (!dup) { idx[k] = c; break; }
                if (++tries > 32) { ok = false; break; }
            }
        }
        if (!ok) continue;
        double P[9], J[9];
        for (int k = 0; k < 3; k++) {
            const float *X = obj + 3 * idx[k], *m = img + 2 * idx[k];
            P[3 * k] = X[0]; P[3 * k + 1] = X[1]; P[3 * k + 2] = X[2];
            double b[3] = {((double)m[0] - a.cx) / a.fx, ((double)m[1] - a.cy) / a.fy, 1.0};
            pnp_normalize(b);
            J[3 * k] = b[0]; J[3 * k + 1] = b[1]; J[3 * k + 2] = b[2];
        }
        double sol[48];
        const int ns = pnp_p3p(P, J, sol);
        double best_e = 1e300;
        int best_s = -1;
        for (int s = 0; s < ns; s++) {  // the fourth point selects the solution
            double u, v;
            if (!isfinite(sol[12 * s]) || !pnp_project(sol + 12 * s, obj + 3 * idx[3], a.fx, a.fy, a.cx, a.cy, u, v)) continue;
            const double du = u - (double)img[2 * idx[3]], dv = v - (double)img[2 * idx[3] + 1];
            const double e = du * du + dv * dv;
            if (e < best_e) { best_e = e; best_s = s; }
        }
        if (best_s >= 0) {
            for (int k = 0; k < 12; k++) out[k] = sol[12 * best_s + k];
            atomicAdd(&s_valid, 1);
        }
    }
    __syncthreads();
    // ---- phase 2: inlier counts, one warp per hypothesis
    for (int h = warp; h < a.iterations; h += PNP_THREADS / 32) {
        const double *Rt = hyp + 12 * h;
        if (!isfinite(Rt[0])) continue;
        int cnt = 0;
        for (int i = lane; i < n; i += 32) {
            double u, v;
            if (pnp_project(Rt, obj + 3 * i, a.fx, a.fy, a.cx, a.cy, u, v)) {
                const double du = u - (double)img[2 * i], dv = v - (double)img[2 * i + 1];
                cnt += du * du + dv * dv <= a.thr2;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0) atomicMax(&s_best, ((unsigned long long)cnt << 32) | (unsigned long long)(0xffffffffu - (unsigned)h));
    }
    __syncthreads();
    const int best_cnt = (int)(s_best >> 32), best_h = (int)(0xffffffffu - (unsigned)(s_best & 0xffffffffull));
    if (best_cnt < 4) {
        if (tid == 0) { info[0] = 0; info[1] = 0; info[2] = s_valid; info[3] = 0; }
        return;
    }
    // ---- phase 3: inlier mask of the best hypothesis
    if (tid < 12) s_rt[tid] = hyp[12 * best_h + tid];
    __syncthreads();
    for (int i = tid; i < n; i += PNP_THREADS) {
        double u, v;
        bool in = false;
        if (pnp_project(s_rt, obj + 3 * i, a.fx, a.fy, a.cx, a.cy, u, v)) {
            const double du = u - (double)img[2 * i], dv = v - (double)img[2 * i + 1];
            in = du * du + dv * dv <= a.thr2;
        }
        inl[i] = in;
    }
    __syncthreads();
    // ---- phase 4: Levenberg-Marquardt on the inliers
    double lambda = -1.0, chi_cur = 0;
    int iters = 0;
    for (int it = 0; it < 50; it++) {
        double H[21], g[6], chi = 0;
#pragma unroll
        for (int k = 0; k < 21; k++) H[k] = 0;
#pragma unroll
        for (int k = 0; k < 6; k++) g[k] = 0;
        for (int i = tid; i < n; i += PNP_THREADS) {
            if (!inl[i]) continue;
            const double x = obj[3 * i], y = obj[3 * i + 1], z = obj[3 * i + 2];
            const double X = s_rt[0] * x + s_rt[1] * y + s_rt[2] * z + s_rt[9];
            const double Y = s_rt[3] * x + s_rt[4] * y + s_rt[5] * z + s_rt[10];
            const double Z = s_rt[6] * x + s_rt[7] * y + s_rt[8] * z + s_rt[11];
            const double Zi = 1.0 / Z, Zi2 = Zi * Zi;
            const double r0 = (double)img[2 * i] - (a.fx * X * Zi + a.cx), r1 = (double)img[2 * i + 1] - (a.fy * Y * Zi + a.cy);
            const double A[12] = {-a.fx * Zi, 0, a.fx * X * Zi2, a.fx * X * Y * Zi2, -a.fx - a.fx * X * X * Zi2, a.fx * Y * Zi,
                                  0, -a.fy * Zi, a.fy * Y * Zi2, a.fy + a.fy * Y * Y * Zi2, -a.fy * X * Y * Zi2, -a.fy * X * Zi};
            chi += r0 * r0 + r1 * r1;
            int k = 0;
#pragma unroll
            for (int q = 0; q < 6; q++) {
                g[q] -= A[q] * r0 + A[6 + q] * r1;
#pragma unroll
                for (int s = q; s < 6; s++) H[k++] += A[q] * A[s] + A[6 + q] * A[6 + s];
            }
        }
        {
            int k = 0;
            for (int q = 0; q < 6; q++) {
                const double gq = pnp_block_sum(g[q], red);
                if (tid == 0) s_g[q] = gq;
                for (int s = q; s < 6; s++) {
                    const double hv = pnp_block_sum(H[k++], red);
                    if (tid == 0) { s_H[6 * q + s] = hv; s_H[6 * s + q] = hv; }
                }
            }
        }
        chi = pnp_block_sum(chi, red);
        __syncthreads();
        if (it == 0) {
            chi_cur = chi;
            double mx = 0;
            for (int q = 0; q < 6; q++) mx = fmax(mx, s_H[7 * q]);
            lambda = 1e-4 * mx;
        }
        // damped step, retried with a larger lambda until the reprojection error decreases
        bool accepted = false, converged = false;
        for (int tr = 0; tr < 12 && !accepted && !converged; tr++) {
            if (tid == 0) {
                double L[36];
                for (int k = 0; k < 36; k++) L[k] = s_H[k];
                for (int q = 0; q < 6; q++) L[7 * q] += lambda;
                bool ok = true;
                for (int j = 0; j < 6 && ok; j++) {  // Cholesky
                    double d = L[7 * j];
                    for (int k = 0; k < j; k++) d -= L[6 * j + k] * L[6 * j + k];
                    if (!(d > 0)) { ok = false; break; }
                    d = sqrt(d);
                    L[7 * j] = d;
                    for (int i = j + 1; i < 6; i++) {
                        double s = L[6 * i + j];
                        for (int k = 0; k < j; k++) s -= L[6 * i + k] * L[6 * j + k];
                        L[6 * i + j] = s / d;
                    }
                }
                if (ok) {
                    double yv[6];
                    for (int i = 0; i < 6; i++) {
                        double s = s_g[i];
                        for (int k = 0; k < i; k++) s -= L[6 * i + k] * yv[k];
                        yv[i] = s / L[7 * i];
                    }
                    for (int i = 5; i >= 0; i--) {
                        double s = yv[i];
                        for (int k = i + 1; k < 6; k++) s -= L[6 * k + i] * s_dx[k];
                        s_dx[i] = s / L[7 * i];
                    }
                    for (int k = 0; k < 12; k++) s_rtb[k] = s_rt[k];
                    pose_oplus(s_rt, s_dx);
                }
                s_flag = ok;
            }
            __syncthreads();
            const bool ok = s_flag;
            double chi_new = 0;
            bool front = true;
            if (ok) {
                for (int i = tid; i < n; i += PNP_THREADS) {
                    if (!inl[i]) continue;
                    double u, v;
                    if (!pnp_project(s_rt, obj + 3 * i, a.fx, a.fy, a.cx, a.cy, u, v)) { front = false; continue; }
                    const double du = u - (double)img[2 * i], dv = v - (double)img[2 * i + 1];
                    chi_new += du * du + dv * dv;
                }
            }
            chi_new = pnp_block_sum(chi_new, red);
            const int behind = __syncthreads_or(!front);
            double step = 0;
            for (int q = 0; q < 6; q++) step = fmax(step, fabs(s_dx[q]));
            if (ok && !behind && chi_new <= chi_cur) {
                accepted = true;
                converged = step < 1e-12 || chi_cur - chi_new <= 1e-14 * (chi_cur + 1e-30);
                chi_cur = chi_new;
                lambda *= 1.0 / 3.0;
            } else {
                __syncthreads();
                if (ok && tid < 12) s_rt[tid] = s_rtb[tid];  // undo
                lambda *= 5.0;
                __syncthreads();
            }
        }
        iters = it + 1;
        if (!accepted || converged) break;
    }
    __syncthreads();
    if (tid == 0) {
        double q[4], lg[6];
        R_to_quat(s_rt, q);
        double *o = a.pose7 + 7 * p;
        o[0] = q[0]; o[1] = q[1]; o[2] = q[2]; o[3] = q[3]; o[4] = s_rt[9]; o[5] = s_rt[10]; o[6] = s_rt[11];
        se3_log(s_rt, lg);
        double *rv = a.rvec_tvec + 6 * p;
        rv[0] = lg[3]; rv[1] = lg[4]; rv[2] = lg[5]; rv[3] = s_rt[9]; rv[4] = s_rt[10]; rv[5] = s_rt[11];
        info[0] = 1; info[1] = best_cnt; info[2] = s_valid; info[3] = iters;
    }
}

// ================================================================================================
// host side
// ================================================================================================
static void free_pnp(sb_pnp *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_obj, h->d_img, h->d_n, h->d_info, h->d_pose, h->d_rt, h->d_inlier};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int sb_pnp_create(sb_pnp_t **out, int device, int max_problems, int max_points) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_problems >= 1 && max_problems <= 65535, "max_problems out of range [1, 65535]");
    SB_REQUIRE(max_points >= 4 && max_points <= (1 << 20), "max_points out of range [4, 2^20]");
    SB_TRY(sb_use_device(device));
    sb_pnp *h = new sb_pnp();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_problems = max_problems;
    h->max_points = max_points;
    const size_t P = max_problems, MP = max_points;
    cudaError_t e = cudaMalloc((void **)&h->d_obj, P * MP * 12);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_img, P * MP * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_n, P * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_info, P * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_pose, P * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_rt, P * 48);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_inlier, P * MP);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pnp_ransac, cudaFuncAttributeMaxDynamicSharedMemorySize, PNP_MAX_HYP * 96);
    if (e != cudaSuccess) {
        sb_set_error("sb_pnp_create: %s", cudaGetErrorString(e));
        free_pnp(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_pnp_destroy(sb_pnp_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_pnp(h);
    }
    return SB_OK;
}

extern "C" int sb_pnp_set_stream(sb_pnp_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_pnp_ransac_dev(sb_pnp_t *h, int n_problems, const int32_t *d_n_points, const float *d_obj, const float *d_img,
                                 int max_points, const double *K, int iterations, double reproj_err, uint64_t seed,
                                 double *d_pose7, double *d_rvec_tvec, uint8_t *d_inlier, int32_t *d_info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(n_problems >= 1 && n_problems <= 65535, "n_problems out of range");
    SB_REQUIRE(d_n_points && d_obj && d_img && K && d_pose7 && d_rvec_tvec && d_inlier && d_info, "null pointer");
    SB_REQUIRE(max_points >= 4, "max_points must be >= 4");
    SB_REQUIRE(iterations >= 1 && iterations <= PNP_MAX_HYP, "iterations out of range [1, 1024]");
    SB_REQUIRE(reproj_err > 0, "reprojection error must be positive");
    SB_TRY(sb_use_device(h->device));
    PnpArgs a;
    a.n = d_n_points; a.obj = d_obj; a.img = d_img; a.pose7 = d_pose7; a.rvec_tvec = d_rvec_tvec; a.inlier = d_inlier; a.info = d_info;
    a.MP = max_points; a.iterations = iterations;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
    a.thr2 = reproj_err * reproj_err;
    a.seed = seed;
    k_pnp_ransac<<<n_problems, PNP_THREADS, (size_t)iterations * 96, h->stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_pnp_ransac(sb_pnp_t *h, int n_problems, const int32_t *n_points, const float *obj, const float *img,
                             const double *K, int iterations, double reproj_err, uint64_t seed, double *pose7,
                             double *rvec_tvec, uint8_t *inlier, int32_t *info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(n_problems >= 1 && n_problems <= h->max_problems, "n_problems out of range [1, max_problems]");
    SB_REQUIRE(n_points && obj && img && K && pose7 && rvec_tvec && inlier && info, "null pointer");
    for (int p = 0; p < n_problems; p++) SB_REQUIRE(n_points[p] >= 0 && n_points[p] <= h->max_points, "n_points out of range [0, max_points]");
    SB_TRY(sb_use_device(h->device));
    const size_t P = n_problems, MP = h->max_points;
    cudaStream_t s = h->stream;
    SB_CUDA(cudaMemcpyAsync(h->d_n, n_points, P * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_obj, obj, P * MP * 12, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_img, img, P * MP * 8, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemsetAsync(h->d_pose, 0, P * 56, s));
    SB_CUDA(cudaMemsetAsync(h->d_rt, 0, P * 48, s));
    SB_TRY(sb_pnp_ransac_dev(h, n_problems, h->d_n, h->d_obj, h->d_img, h->max_points, K, iterations, reproj_err, seed, h->d_pose,
                             h->d_rt, h->d_inlier, h->d_info));
    SB_CUDA(cudaMemcpyAsync(pose7, h->d_pose, P * 56, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(rvec_tvec, h->d_rt, P * 48, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(inlier, h->d_inlier, P * MP, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(info, h->d_info, P * 16, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}
