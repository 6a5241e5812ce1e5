// se3.cuh — SE(3) helpers shared by the BA and pose-graph kernels: poses are kept as 12 doubles
// (row-major R, then t); the C ABI exchanges Sophus' storage order (qx qy qz qw tx ty tz).
// exp / log follow Sophus' published formulas (SE3d::exp, SE3d::log, SO3d::log via the unit quaternion).
#pragma once
#include <math.h>

static __device__ __forceinline__ void quat_to_R(const double *q, double *R) {  // q = (x, y, z, w)
    const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

static __device__ void R_to_quat(const double *R, double *q) {  // -> (x, y, z, w), w >= 0
    const double tr = R[0] + R[4] + R[8];
    double x, y, z, w;
    if (tr > 0) {
        const double s = sqrt(tr + 1.0) * 2; w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        const double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
    } else {
        const double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
    }
    if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
    const double n = sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

// Sophus SE3d::exp([upsilon, omega]) applied on the left: Rt <- exp(d) * Rt  (VertexPose::oplusImpl)
static __device__ void pose_oplus(double *Rt, const double *d) {
    const double wx = d[3], wy = d[4], wz = d[5];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double a, b, c;
    if (th < 1e-10) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; c = 1.0 / 6.0 - th2 / 120.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double W2[9], E[9], V[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) W2[3 * i + j] = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        E[i] = I + a * W[i] + b * W2[i];
        V[i] = I + b * W[i] + c * W2[i];
    }
    double Et[3], Rn[9], tn[3];
#pragma unroll
    for (int i = 0; i < 3; i++) Et[i] = V[3 * i] * d[0] + V[3 * i + 1] * d[1] + V[3 * i + 2] * d[2];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) Rn[3 * i + j] = E[3 * i] * Rt[j] + E[3 * i + 1] * Rt[3 + j] + E[3 * i + 2] * Rt[6 + j];
        tn[i] = E[3 * i] * Rt[9] + E[3 * i + 1] * Rt[10] + E[3 * i + 2] * Rt[11] + Et[i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) Rt[i] = Rn[i];
    Rt[9] = tn[0]; Rt[10] = tn[1]; Rt[11] = tn[2];
}


// Rt_out = A * B
static __device__ __forceinline__ void se3_mul(const double *A, const double *B, double *C) {
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
#pragma unroll
        for (int j = 0; j < 3; j++) R[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
        t[i] = A[3 * i] * B[9] + A[3 * i + 1] * B[10] + A[3 * i + 2] * B[11] + A[9 + i];
    }
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = R[i];
    C[9] = t[0]; C[10] = t[1]; C[11] = t[2];
}

// C = A^-1
static __device__ __forceinline__ void se3_inv(const double *A, double *C) {
    double R[9], t[3];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) R[3 * i + j] = A[3 * j + i];
#pragma unroll
    for (int i = 0; i < 3; i++) t[i] = -(R[3 * i] * A[9] + R[3 * i + 1] * A[10] + R[3 * i + 2] * A[11]);
#pragma unroll
    for (int i = 0; i < 9; i++) C[i] = R[i];
    C[9] = t[0]; C[10] = t[1]; C[11] = t[2];
}

// Sophus SE3d::log -> [upsilon, omega]
static __device__ void se3_log(const double *A, double *out) {
    double q[4];
    R_to_quat(A, q);
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2], w = q[3];
    double two;
    if (n2 < 1e-20) {
        two = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
    } else {
        const double n = sqrt(n2);
        two = fabs(w) < 1e-10 ? 3.14159265358979323846 / n : 2.0 * atan(n / w) / n;
    }
    const double wx = two * q[0], wy = two * q[1], wz = two * q[2];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double coef;
    if (th < 1e-10) coef = 1.0 / 12.0;
    else { const double half = 0.5 * th; coef = (1.0 - th * cos(half) / (2.0 * sin(half))) / th2; }
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double Vi[9];
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) {
            const double w2 = W[3 * i] * W[j] + W[3 * i + 1] * W[3 + j] + W[3 * i + 2] * W[6 + j];
            Vi[3 * i + j] = (i == j ? 1.0 : 0.0) - 0.5 * W[3 * i + j] + coef * w2;
        }
#pragma unroll
    for (int i = 0; i < 3; i++) out[i] = Vi[3 * i] * A[9] + Vi[3 * i + 1] * A[10] + Vi[3 * i + 2] * A[11];
    out[3] = wx; out[4] = wy; out[5] = wz;
}
