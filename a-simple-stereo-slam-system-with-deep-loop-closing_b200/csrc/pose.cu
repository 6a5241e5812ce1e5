// pose.cu — batched pose-only optimisation on B200 (sm_100a), double precision.
//
// SURVEY §8(f) "next" row 1 (the every-frame solver of the live front end): replaces the g2o solve inside
// Frontend::EstimateCurrentPose (reference src/frontend.cpp:176-276) and LoopClosing::OptimizeCurrentPose
// (src/loopclosing.cpp:339-433): one VertexPose, one EdgeProjectionPoseOnly per matched map point
// (include/myslam/g2o_types.h:63-102), Huber kernel with g2o's default delta 1.0, Levenberg-Marquardt over a
// dense 6x6 system, `rounds` rounds of optimize(inner) with chi2 classification in between (outliers leave
// the next round, may come back; the robust kernels are dropped after round rounds-2).
// One CTA per frame, everything in one launch; sums are block reductions in a fixed order.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "se3.cuh"

#define PO_THREADS 128

struct sb_pose {
    int device, max_frames, max_obs;
    cudaStream_t stream, own_stream;
    int32_t *d_n, *d_info;
    double *d_poses, *d_points, *d_uv, *d_err;
    uint8_t *d_outlier, *d_level;
};

struct PoArgs {
    const int32_t *n;
    double *poses;          // [F][7] in/out
    const double *points;   // [F][MO][3]
    const double *uv;       // [F][MO][2]
    uint8_t *outlier;       // [F][MO] out
    int32_t *info;          // [F][4]: inliers, LM iterations, rounds, 0
    double *err;            // [F][MO][2] workspace
    uint8_t *level;         // [F][MO] workspace
    int MO;
    double fx, fy, cx, cy, delta, chi2_th;
    int pre_rounds, rounds, inner_iters;
};

static __device__ __forceinline__ void po_huber(double e2, double delta, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho0 = e2; rho1 = 1.0; }
    else { const double s = sqrt(e2); rho0 = 2 * s * delta - dsqr; rho1 = delta / s; }
}

// sums `nv` values per thread over the CTA; result in out[0..nv) (shared), fixed order
template <int NV>
static __device__ void po_block_sum(double (&v)[NV], double *red, double *out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) red[wid * NV + k] = v[k];
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0;
        for (int w = 0; w < PO_THREADS / 32; w++) s += red[w * NV + threadIdx.x];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// EdgeProjectionPoseOnly::computeError (g2o_types.h:70-76)
static __device__ __forceinline__ void po_edge(const PoArgs &a, const double *Rt, const double *p, const double *uv, double *r,
                                               double *pc) {
#pragma unroll
    for (int i = 0; i < 3; i++) pc[i] = Rt[3 * i] * p[0] + Rt[3 * i + 1] * p[1] + Rt[3 * i + 2] * p[2] + Rt[9 + i];
    const double px = a.fx * pc[0] + a.cx * pc[2], py = a.fy * pc[1] + a.cy * pc[2];
    r[0] = uv[0] - px / pc[2];
    r[1] = uv[1] - py / pc[2];
}

// computeActiveErrors + activeRobustChi2 over the level-0 edges
static __device__ double po_active_errors(const PoArgs &a, const double *Rt, int n, const double *pts, const double *uv,
                                          const uint8_t *level, double *err, bool robust, double *red, double *out) {
    double chi[1] = {0};
    for (int e = threadIdx.x; e < n; e += PO_THREADS) {
        if (level[e]) continue;
        double r[2], pc[3];
        po_edge(a, Rt, pts + 3 * e, uv + 2 * e, r, pc);
        err[2 * e] = r[0];
        err[2 * e + 1] = r[1];
        const double e2 = r[0] * r[0] + r[1] * r[1];
        if (robust) { double r0, r1; po_huber(e2, a.delta, r0, r1); chi[0] += r0; }
        else chi[0] += e2;
    }
    po_block_sum<1>(chi, red, out);
    const double res = out[0];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(PO_THREADS) k_pose_only(const __grid_constant__ PoArgs a) {
    __shared__ double Rt[12], Rtb[12], Hs[27], xs[6], red[(PO_THREADS / 32) * 27], out1[1];
    __shared__ int s_ok;
    const int f = blockIdx.x, tid = threadIdx.x;
    const int n = min(max(a.n[f], 0), a.MO);
    const double *pts = a.points + (size_t)f * a.MO * 3, *uv = a.uv + (size_t)f * a.MO * 2;
    double *err = a.err + (size_t)f * a.MO * 2;
    uint8_t *level = a.level + (size_t)f * a.MO, *is_out = a.outlier + (size_t)f * a.MO;
    if (tid == 0) {
        quat_to_R(a.poses + 7 * f, Rt);
        Rt[9] = a.poses[7 * f + 4]; Rt[10] = a.poses[7 * f + 5]; Rt[11] = a.poses[7 * f + 6];
    }
    for (int e = tid; e < n; e += PO_THREADS) { level[e] = 0; is_out[e] = 0; err[2 * e] = 0; err[2 * e + 1] = 0; }
    __syncthreads();

    bool robust = true;
    int lm = 0, n_out = 0;
    const int total_rounds = a.pre_rounds + a.rounds;
    for (int round = 0; round < total_rounds; round++) {
        int cnt[1] = {0};
        {
            double c[1] = {0};
            for (int e = tid; e < n; e += PO_THREADS) c[0] += level[e] ? 0.0 : 1.0;
            po_block_sum<1>(c, red, out1);
            cnt[0] = (int)(out1[0] + 0.5);
            __syncthreads();
        }
        if (cnt[0] > 0) {  // optimizer.optimize(inner_iters): Levenberg-Marquardt on the active edges
            double lambda = 0, ni = 2;
            bool terminated = false;
            for (int it = 0; it < a.inner_iters && !terminated; it++) {
                double currentChi = po_active_errors(a, Rt, n, pts, uv, level, err, robust, red, out1);
                double h[27];  // 21 upper-triangle entries of H, then b
#pragma unroll
                for (int k = 0; k < 27; k++) h[k] = 0;
                for (int e = tid; e < n; e += PO_THREADS) {
                    if (level[e]) continue;
                    const double *p = pts + 3 * e;
                    double pc[3];
#pragma unroll
                    for (int i = 0; i < 3; i++) pc[i] = Rt[3 * i] * p[0] + Rt[3 * i + 1] * p[1] + Rt[3 * i + 2] * p[2] + Rt[9 + i];
                    const double X = pc[0], Y = pc[1], Z = pc[2];
                    const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
                    const double A[12] = {-a.fx * Zinv, 0, a.fx * X * Zinv2, a.fx * X * Y * Zinv2, -a.fx - a.fx * X * X * Zinv2, a.fx * Y * Zinv,
                                          0, -a.fy * Zinv, a.fy * Y * Zinv2, a.fy + a.fy * Y * Y * Zinv2, -a.fy * X * Y * Zinv2, -a.fy * X * Zinv};
                    const double r0 = err[2 * e], r1 = err[2 * e + 1];
                    double w = 1.0;
                    if (robust) { double q0; po_huber(r0 * r0 + r1 * r1, a.delta, q0, w); }
                    int k = 0;
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        h[21 + i] += -w * (A[i] * r0 + A[6 + i] * r1);
#pragma unroll
                        for (int j = i; j < 6; j++) h[k++] += w * (A[i] * A[j] + A[6 + i] * A[6 + j]);
                    }
                }
                po_block_sum<27>(h, red, Hs);
                if (it == 0) {  // computeLambdaInit
                    const double mx = fmax(fmax(fmax(fabs(Hs[0]), fabs(Hs[6])), fmax(fabs(Hs[11]), fabs(Hs[15]))), fmax(fabs(Hs[18]), fabs(Hs[20])));
                    lambda = 1e-5 * mx;
                    ni = 2;
                }
                double rho = 0;
                int qmax = 0;
                do {
                    if (tid < 12) Rtb[tid] = Rt[tid];
                    if (tid == 0) {  // (H + lambda I) x = b by Cholesky (LinearSolverDense)
                        double S[36], x[6];
                        int k = 0;
                        for (int i = 0; i < 6; i++)
                            for (int j = i; j < 6; j++) { S[6 * i + j] = Hs[k]; S[6 * j + i] = Hs[k]; k++; }
                        for (int i = 0; i < 6; i++) { S[7 * i] += lambda; x[i] = Hs[21 + i]; }
                        int ok = 1;
                        for (int j = 0; j < 6 && ok; j++) {
                            double d = S[7 * j];
                            for (int q = 0; q < j; q++) d -= S[6 * j + q] * S[6 * j + q];
                            if (!(d > 0)) { ok = 0; break; }
                            d = sqrt(d);
                            S[7 * j] = d;
                            for (int i = j + 1; i < 6; i++) {
                                double v = S[6 * i + j];
                                for (int q = 0; q < j; q++) v -= S[6 * i + q] * S[6 * j + q];
                                S[6 * i + j] = v / d;
                            }
                        }
                        if (ok) {
                            for (int i = 0; i < 6; i++) { double v = x[i]; for (int q = 0; q < i; q++) v -= S[6 * i + q] * x[q]; x[i] = v / S[7 * i]; }
                            for (int i = 5; i >= 0; i--) { double v = x[i]; for (int q = i + 1; q < 6; q++) v -= S[6 * q + i] * x[q]; x[i] = v / S[7 * i]; }
                            for (int i = 0; i < 6; i++) xs[i] = x[i];
                        }
                        s_ok = ok;
                    }
                    __syncthreads();
                    const int ok = s_ok;
                    double scale = 1e-3;
                    if (ok) {
#pragma unroll
                        for (int i = 0; i < 6; i++) scale += xs[i] * (lambda * xs[i] + Hs[21 + i]);
                        __syncthreads();
                        if (tid == 0) pose_oplus(Rt, xs);
                        __syncthreads();
                    }
                    double tempChi = po_active_errors(a, Rt, n, pts, uv, level, err, robust, red, out1);
                    if (!ok) tempChi = 1.7976931348623157e308;
                    rho = (currentChi - tempChi) / scale;
                    if (rho > 0 && isfinite(tempChi)) {
                        double alpha = 1. - pow(2 * rho - 1, 3);
                        alpha = fmin(alpha, 2. / 3.);
                        lambda *= fmax(1. / 3., alpha);
                        ni = 2;
                        currentChi = tempChi;
                    } else {
                        lambda *= ni;
                        ni *= 2;
                        __syncthreads();
                        if (tid < 12) Rt[tid] = Rtb[tid];
                        __syncthreads();
                    }
                    qmax++;
                } while (rho < 0 && qmax < 10);
                lm++;
                if (qmax == 10 || rho == 0) terminated = true;
            }
        }
        if (round >= a.pre_rounds) {  // classification (src/frontend.cpp:229-247)
            double c[1] = {0};
            for (int e = tid; e < n; e += PO_THREADS) {
                double r0 = err[2 * e], r1 = err[2 * e + 1];
                if (is_out[e]) {  // not part of the last optimisation: evaluate at the new pose
                    double r[2], pc[3];
                    po_edge(a, Rt, pts + 3 * e, uv + 2 * e, r, pc);
                    r0 = r[0]; r1 = r[1];
                    err[2 * e] = r0; err[2 * e + 1] = r1;
                }
                const bool o = r0 * r0 + r1 * r1 > a.chi2_th;
                is_out[e] = o;
                level[e] = o;
                c[0] += o ? 1.0 : 0.0;
            }
            po_block_sum<1>(c, red, out1);
            n_out = (int)(out1[0] + 0.5);
            __syncthreads();
            if (round - a.pre_rounds == a.rounds - 2) robust = false;
        }
    }
    if (tid == 0) {
        R_to_quat(Rt, a.poses + 7 * f);
        a.poses[7 * f + 4] = Rt[9]; a.poses[7 * f + 5] = Rt[10]; a.poses[7 * f + 6] = Rt[11];
        a.info[4 * f] = n - n_out; a.info[4 * f + 1] = lm; a.info[4 * f + 2] = total_rounds; a.info[4 * f + 3] = 0;
    }
}

static void free_pose(sb_pose *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_n, h->d_info, h->d_poses, h->d_points, h->d_uv, h->d_err, h->d_outlier, h->d_level};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int sb_pose_create(sb_pose_t **out, int device, int max_frames, int max_obs) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_frames >= 1 && max_frames <= 65535, "max_frames out of range [1, 65535]");
    SB_REQUIRE(max_obs >= 1 && max_obs <= (1 << 22), "max_obs out of range");
    SB_TRY(sb_use_device(device));
    sb_pose *h = new sb_pose();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_frames = max_frames;
    h->max_obs = max_obs;
    const size_t F = max_frames, MO = max_obs;
    cudaError_t e = cudaMalloc((void **)&h->d_n, F * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_info, F * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_poses, F * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_points, F * MO * 24);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_uv, F * MO * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_err, F * MO * 16);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_outlier, F * MO);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_level, F * MO);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        sb_set_error("sb_pose_create: %s", cudaGetErrorString(e));
        free_pose(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_pose_destroy(sb_pose_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_pose(h);
    }
    return SB_OK;
}

extern "C" int sb_pose_set_stream(sb_pose_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_pose_solve_dev(sb_pose_t *h, int n_frames, const int32_t *d_n_obs, double *d_poses, const double *d_points,
                                 const double *d_uv, const double *K, double huber_delta, double chi2_th, int pre_rounds,
                                 int rounds, int inner_iters, uint8_t *d_outlier, int32_t *d_info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && d_n_obs && d_poses && d_points && d_uv && K && d_outlier && d_info, "null pointer");
    SB_REQUIRE(n_frames >= 1 && n_frames <= h->max_frames, "n_frames out of range [1, max_frames]");
    SB_REQUIRE(huber_delta > 0 && pre_rounds >= 0 && rounds >= 1 && inner_iters >= 1, "bad solver parameters");
    SB_TRY(sb_use_device(h->device));
    PoArgs a;
    a.n = d_n_obs; a.poses = d_poses; a.points = d_points; a.uv = d_uv; a.outlier = d_outlier; a.info = d_info;
    a.err = h->d_err; a.level = h->d_level; a.MO = h->max_obs;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3]; a.delta = huber_delta; a.chi2_th = chi2_th;
    a.pre_rounds = pre_rounds; a.rounds = rounds; a.inner_iters = inner_iters;
    k_pose_only<<<n_frames, PO_THREADS, 0, h->stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_pose_solve(sb_pose_t *h, int n_frames, const int32_t *n_obs, double *poses, const double *points,
                             const double *uv, const double *K, double huber_delta, double chi2_th, int pre_rounds, int rounds,
                             int inner_iters, uint8_t *outlier, int32_t *info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && n_obs && poses && points && uv && outlier && info, "null pointer");
    SB_REQUIRE(n_frames >= 1 && n_frames <= h->max_frames, "n_frames out of range [1, max_frames]");
    for (int f = 0; f < n_frames; f++) SB_REQUIRE(n_obs[f] >= 0 && n_obs[f] <= h->max_obs, "n_obs out of range [0, max_obs]");
    SB_TRY(sb_use_device(h->device));
    const size_t F = n_frames, MO = h->max_obs;
    cudaStream_t s = h->stream;
    SB_CUDA(cudaMemcpyAsync(h->d_n, n_obs, F * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_poses, poses, F * 56, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_points, points, F * MO * 24, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_uv, uv, F * MO * 16, cudaMemcpyHostToDevice, s));
    SB_TRY(sb_pose_solve_dev(h, n_frames, h->d_n, h->d_poses, h->d_points, h->d_uv, K, huber_delta, chi2_th, pre_rounds, rounds,
                             inner_iters, h->d_outlier, h->d_info));
    SB_CUDA(cudaMemcpyAsync(poses, h->d_poses, F * 56, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(outlier, h->d_outlier, F * MO, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(info, h->d_info, F * 16, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}
