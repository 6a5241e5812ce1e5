// triangulate.cu — batched linear (DLT) stereo triangulation on B200 (sm_100a), double precision.
//
// SURVEY §8(f) "next" row 3: replaces myslam::triangulation (reference include/myslam/algorithm.h:16-33) as the
// front end calls it for every left/right correspondence (src/frontend.cpp:385-417 BuildInitMap, :451-488
// TriangulateNewPoints): two camera poses shared by all points, A = [x m.row(2) - m.row(0); y m.row(2) - m.row(1)]
// per view (4 x 4), the right singular vector of the smallest singular value de-homogenised, accepted iff
// sigma_4 / sigma_3 < 1e-2 and (caller's test) z > 0; optionally mapped by T_wc (currentPoseTwc * pcamera).
// One thread per correspondence: the singular pairs of the 4 x 4 matrix come from a cyclic Jacobi
// eigen-decomposition of A^T A in double precision.
#include <math.h>
#include <string.h>

#include <mutex>

#include "common.cuh"
#include "se3.cuh"

struct TriArgs {
    int n;
    const float *uvl, *uvr;  // [n][2] pixels (cv::KeyPoint::pt is float)
    double *pts;             // [n][3]
    uint8_t *ok;             // [n]
    double Kl[4], Kr[4];     // fx fy cx cy
    double Ml[12], Mr[12];   // 3x4 pose matrices, row major
    double Twc[12];          // R | t applied to accepted points
    int has_twc;
    double ratio_th;
};

__global__ void __launch_bounds__(128) k_triangulate(const __grid_constant__ TriArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    // Camera::pixel2camera with depth 1 (src/camera.cpp:25-29)
    const double xl = ((double)a.uvl[2 * i] - a.Kl[2]) / a.Kl[0], yl = ((double)a.uvl[2 * i + 1] - a.Kl[3]) / a.Kl[1];
    const double xr = ((double)a.uvr[2 * i] - a.Kr[2]) / a.Kr[0], yr = ((double)a.uvr[2 * i + 1] - a.Kr[3]) / a.Kr[1];
    double A[16];
#pragma unroll
    for (int c = 0; c < 4; c++) {
        A[c] = xl * a.Ml[8 + c] - a.Ml[c];
        A[4 + c] = yl * a.Ml[8 + c] - a.Ml[4 + c];
        A[8 + c] = xr * a.Mr[8 + c] - a.Mr[c];
        A[12 + c] = yr * a.Mr[8 + c] - a.Mr[4 + c];
    }
    double M[16], V[16];  // M = A^T A, V accumulates the rotations
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            M[4 * r + c] = A[r] * A[c] + A[4 + r] * A[4 + c] + A[8 + r] * A[8 + c] + A[12 + r] * A[12 + c];
            V[4 * r + c] = r == c ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 12; sweep++) {
        double off = 0;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) off += M[4 * p + q] * M[4 * p + q];
        if (off < 1e-300) break;
#pragma unroll
        for (int p = 0; p < 3; p++)
#pragma unroll
            for (int q = p + 1; q < 4; q++) {
                const double apq = M[4 * p + q];
                if (apq == 0.0) continue;
                const double theta = (M[5 * q] - M[5 * p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
                for (int k = 0; k < 4; k++) {  // M <- M J
                    const double mkp = M[4 * k + p], mkq = M[4 * k + q];
                    M[4 * k + p] = cs * mkp - sn * mkq;
                    M[4 * k + q] = sn * mkp + cs * mkq;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {  // M <- J^T M
                    const double mpk = M[4 * p + k], mqk = M[4 * q + k];
                    M[4 * p + k] = cs * mpk - sn * mqk;
                    M[4 * q + k] = sn * mpk + cs * mqk;
                }
#pragma unroll
                for (int k = 0; k < 4; k++) {  // V <- V J
                    const double vkp = V[4 * k + p], vkq = V[4 * k + q];
                    V[4 * k + p] = cs * vkp - sn * vkq;
                    V[4 * k + q] = sn * vkp + cs * vkq;
                }
            }
    }
    // smallest and second smallest eigenvalue (= squared singular values sigma_4, sigma_3)
    int i4 = 0;
#pragma unroll
    for (int k = 1; k < 4; k++)
        if (M[5 * k] < M[5 * i4]) i4 = k;
    double l3 = 1e300;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if (k != i4 && M[5 * k] < l3) l3 = M[5 * k];
    const double l4 = fmax(M[5 * i4], 0.0);
    const double w = V[12 + i4];
    double p[3] = {V[i4] / w, V[4 + i4] / w, V[8 + i4] / w};
    const bool good = sqrt(l4) / sqrt(fmax(l3, 0.0)) < a.ratio_th && p[2] > 0;
    if (a.has_twc) {
        const double x = p[0], y = p[1], z = p[2];
#pragma unroll
        for (int r = 0; r < 3; r++) p[r] = a.Twc[3 * r] * x + a.Twc[3 * r + 1] * y + a.Twc[3 * r + 2] * z + a.Twc[9 + r];
    }
    a.pts[3 * i] = p[0]; a.pts[3 * i + 1] = p[1]; a.pts[3 * i + 2] = p[2];
    a.ok[i] = good;
}

static void pose7_to_m34(const double *p, double *M) {  // rows of [R | t]
    const double n = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    const double x = p[0] / n, y = p[1] / n, z = p[2] / n, w = p[3] / n;
    const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
    for (int r = 0; r < 3; r++) { M[4 * r] = R[3 * r]; M[4 * r + 1] = R[3 * r + 1]; M[4 * r + 2] = R[3 * r + 2]; M[4 * r + 3] = p[4 + r]; }
}

extern "C" int sb_triangulate_dev(int device, void *stream, int n, const float *d_uv_left, const float *d_uv_right,
                                  const double *K_left, const double *K_right, const double *pose_left7, const double *pose_right7,
                                  const double *T_wc7, double ratio_th, double *d_points, uint8_t *d_ok) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(n >= 0, "negative count");
    if (n == 0) return SB_OK;
    SB_REQUIRE(d_uv_left && d_uv_right && K_left && K_right && pose_left7 && pose_right7 && d_points && d_ok, "null pointer");
    SB_TRY(sb_use_device(device));
    TriArgs a;
    a.n = n; a.uvl = d_uv_left; a.uvr = d_uv_right; a.pts = d_points; a.ok = d_ok;
    memcpy(a.Kl, K_left, sizeof(a.Kl));
    memcpy(a.Kr, K_right, sizeof(a.Kr));
    pose7_to_m34(pose_left7, a.Ml);
    pose7_to_m34(pose_right7, a.Mr);
    a.has_twc = T_wc7 != nullptr;
    if (T_wc7) {
        double M[12];
        pose7_to_m34(T_wc7, M);
        for (int r = 0; r < 3; r++) { a.Twc[3 * r] = M[4 * r]; a.Twc[3 * r + 1] = M[4 * r + 1]; a.Twc[3 * r + 2] = M[4 * r + 2]; a.Twc[9 + r] = M[4 * r + 3]; }
    }
    a.ratio_th = ratio_th;
    k_triangulate<<<sb_div_up(n, 128), 128, 0, (cudaStream_t)stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_triangulate(int device, int n, const float *uv_left, const float *uv_right, const double *K_left,
                              const double *K_right, const double *pose_left7, const double *pose_right7, const double *T_wc7,
                              double ratio_th, double *points, uint8_t *ok) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(n >= 0, "negative count");
    if (n == 0) return SB_OK;
    SB_REQUIRE(uv_left && uv_right && points && ok, "null pointer");
    SB_TRY(sb_use_device(device));
    // Device scratch of the host-pointer entry point: one grow-only buffer per device, kept between calls (a cudaMalloc /
    // cudaFree pair per call costs milliseconds: cudaFree synchronises the device), handed out under a lock.
    enum { SB_MAX_DEVICES = 64 };
    static std::mutex mu;
    static void *scratch[SB_MAX_DEVICES] = {nullptr};
    static size_t scratch_bytes[SB_MAX_DEVICES] = {0};
    SB_REQUIRE(device >= 0 && device < SB_MAX_DEVICES, "device index out of range");
    std::lock_guard<std::mutex> lock(mu);
    const size_t off_p = sb_align_up((size_t)n * 16, 256), off_ok = off_p + sb_align_up((size_t)n * 24, 256);
    const size_t need = off_ok + sb_align_up((size_t)n, 256);
    cudaError_t e = cudaSuccess;
    if (scratch_bytes[device] < need) {
        if (scratch[device]) cudaFree(scratch[device]);
        scratch[device] = nullptr;
        scratch_bytes[device] = 0;
        const size_t want = need < ((size_t)1 << 20) ? ((size_t)1 << 20) : 2 * need;
        e = cudaMalloc(&scratch[device], want);
        if (e == cudaSuccess) scratch_bytes[device] = want;
    }
    float *d_uv = reinterpret_cast<float *>(scratch[device]);
    double *d_p = reinterpret_cast<double *>(reinterpret_cast<uint8_t *>(scratch[device]) + off_p);
    uint8_t *d_ok = reinterpret_cast<uint8_t *>(scratch[device]) + off_ok;
    int rc = SB_OK;
    if (e == cudaSuccess) e = cudaMemcpy(d_uv, uv_left, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d_uv + 2 * (size_t)n, uv_right, (size_t)n * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) rc = sb_triangulate_dev(device, nullptr, n, d_uv, d_uv + 2 * (size_t)n, K_left, K_right, pose_left7, pose_right7, T_wc7, ratio_th, d_p, d_ok);
    if (e == cudaSuccess && rc == SB_OK) e = cudaMemcpy(points, d_p, (size_t)n * 24, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && rc == SB_OK) e = cudaMemcpy(ok, d_ok, (size_t)n, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { sb_set_error("sb_triangulate: %s", cudaGetErrorString(e)); return SB_ERR_CUDA; }
    return rc;
}
