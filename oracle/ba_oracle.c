/*
 * ba_oracle.c — CPU restatement of the sliding-window local bundle adjustment.
 *
 * TEST INFRASTRUCTURE ONLY (see orb_oracle.c): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this, and only as the checker / CPU baseline.
 *
 * What it restates (paths relative to /root/reference):
 *   - the problem Backend::OptimizeActiveMap builds and the outer loop around it
 *     (src/backend.cpp:126-269): one 6-DoF vertex per active keyframe (none fixed), one marginalised
 *     3-D vertex per landmark (fixed when its first observer is outside the window, :175-177), one
 *     EdgeProjection per active observation with information I2 and a Huber kernel of delta 5.991
 *     (:198-200), up to 5 x { initializeOptimization(); optimize(10) } until the inlier ratio
 *     (chi2 <= 5.991) exceeds 0.5 (:212-232);
 *   - the edge / vertex arithmetic of include/myslam/g2o_types.h: VertexPose::oplusImpl (:32-37,
 *     T <- exp(d) T with Sophus' [translation, rotation] tangent order), VertexXYZ::oplusImpl
 *     (:50-54), EdgeProjection::computeError (:115-122) and linearizeOplus (:124-144);
 *   - g2o itself is a THIRD-PARTY dependency that is absent from /root/reference and not installed
 *     (un-pinned "g2o master", README.md:39-40; the API used — g2o::make_unique — dates it to
 *     2018-2020).  Its published algorithm is restated here: OptimizationAlgorithmLevenberg::solve
 *     (lambda_0 = 1e-5 * max diag(H), rho = (chi - chi') / (dx.(lambda dx + b) + 1e-3), good step:
 *     lambda *= max(1/3, min(1 - (2 rho - 1)^3, 2/3)), ni = 2; bad step: lambda *= ni, ni *= 2, at most
 *     10 trials, terminate when the trials are used up or rho == 0), RobustKernelHuber (rho' = 1 or
 *     delta / sqrt(e2); second-order term dropped as in g2o's robustInformation), BlockSolver_6_3 with
 *     Schur complement on the landmarks and a Cholesky solve of the reduced pose system, and
 *     SparseOptimizer::optimize, which does NOT recompute the errors after its last iteration —
 *     chi2() read afterwards is that of the last trial state, accepted or not.
 *     Sophus (also un-vendored) SE3d::exp and the group product are restated likewise.
 *
 * PARITY UNPINNED: the reference has no tests, no golden vectors and cannot be built here, and no
 * g2o / Sophus exists in this container, so this file is pinned only by first principles:
 * tests/test_oracle_ba.py checks the Jacobians against central differences, exp against scipy's
 * matrix exponential, recovery of noise-free ground truth and agreement of the reached minimum with
 * scipy.optimize.least_squares on the same robust objective.
 *
 * Plain C99, double precision throughout (g2o and the reference use double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---- SE3 as rotation matrix (row major) + translation; Sophus stores a unit quaternion (x,y,z,w) ---- */
typedef struct { double R[9], t[3]; } se3;

static void quat_to_R(const double q[4], double R[9]) { /* q = (x, y, z, w) */
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    double x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

static void R_to_quat(const double R[9], double q[4]) { /* -> (x, y, z, w), w >= 0 */
    double tr = R[0] + R[4] + R[8], x, y, z, w;
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2; w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
    }
    if (w < 0) { x = -x; y = -y; z = -z; w = -w; }
    double n = sqrt(x * x + y * y + z * z + w * w);
    q[0] = x / n; q[1] = y / n; q[2] = z / n; q[3] = w / n;
}

static void se3_from7(const double p[7], se3 *T) { quat_to_R(p, T->R); T->t[0] = p[4]; T->t[1] = p[5]; T->t[2] = p[6]; }
static void se3_to7(const se3 *T, double p[7]) { R_to_quat(T->R, p); p[4] = T->t[0]; p[5] = T->t[1]; p[6] = T->t[2]; }

static void mat3_mul(const double A[9], const double B[9], double C[9]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}
static void se3_mul(const se3 *A, const se3 *B, se3 *C) { /* C = A * B */
    se3 r;
    mat3_mul(A->R, B->R, r.R);
    for (int i = 0; i < 3; i++) r.t[i] = A->R[3 * i] * B->t[0] + A->R[3 * i + 1] * B->t[1] + A->R[3 * i + 2] * B->t[2] + A->t[i];
    *C = r;
}
static void se3_apply(const se3 *T, const double p[3], double out[3]) {
    for (int i = 0; i < 3; i++) out[i] = T->R[3 * i] * p[0] + T->R[3 * i + 1] * p[1] + T->R[3 * i + 2] * p[2] + T->t[i];
}

/* Sophus SE3d::exp([upsilon, omega]): R = exp(hat(omega)), t = V(omega) upsilon. */
void orc_se3_exp(const double d[6], double R[9], double t[3]) {
    const double wx = d[3], wy = d[4], wz = d[5];
    const double th2 = wx * wx + wy * wy + wz * wz, th = sqrt(th2);
    double a, b, c; /* R = I + a W + b W^2,  V = I + b W + c W^2 */
    if (th < 1e-10) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; c = 1.0 / 6.0 - th2 / 120.0; }
    else { a = sin(th) / th; b = (1.0 - cos(th)) / th2; c = (th - sin(th)) / (th2 * th); }
    const double W[9] = {0, -wz, wy, wz, 0, -wx, -wy, wx, 0};
    double W2[9];
    mat3_mul(W, W, W2);
    double V[9];
    for (int i = 0; i < 9; i++) {
        const double I = (i % 4 == 0) ? 1.0 : 0.0;
        R[i] = I + a * W[i] + b * W2[i];
        V[i] = I + b * W[i] + c * W2[i];
    }
    for (int i = 0; i < 3; i++) t[i] = V[3 * i] * d[0] + V[3 * i + 1] * d[1] + V[3 * i + 2] * d[2];
}

/* ---- problem ----------------------------------------------------------------------------------- */
typedef struct {
    int n_poses, n_points, n_obs;
    se3 *pose;            /* [n_poses] */
    double *pt;           /* [n_points][3] */
    const uint8_t *fixed; /* [n_points] */
    const int32_t *op, *ol; /* observation -> pose, landmark */
    const double *uv;     /* [n_obs][2] */
    double fx, fy, cx, cy;
    se3 ext;
    double delta;         /* Huber delta */
    double *err;          /* [n_obs][2] error of the last computeActiveErrors */
} ba_problem;

/* EdgeProjection::computeError (g2o_types.h:115-122) */
static void edge_error(const ba_problem *P, const se3 *pose, const double *pt, int e, double err[2]) {
    double pc[3], pe[3];
    se3_apply(&pose[P->op[e]], pt + 3 * P->ol[e], pc);
    se3_apply(&P->ext, pc, pe);
    const double px = P->fx * pe[0] + P->cx * pe[2], py = P->fy * pe[1] + P->cy * pe[2], pz = pe[2];
    err[0] = P->uv[2 * e] - px / pz;
    err[1] = P->uv[2 * e + 1] - py / pz;
}

/* EdgeProjection::linearizeOplus (:124-144): A = d err / d pose (2x6), B = d err / d point (2x3) */
static void edge_jacobians(const ba_problem *P, int e, double A[12], double B[6]) {
    const se3 *T = &P->pose[P->op[e]];
    double pc[3], pe[3];
    se3_apply(T, P->pt + 3 * P->ol[e], pc);
    se3_apply(&P->ext, pc, pe);
    const double X = pe[0], Y = pe[1], Z = pe[2], fx = P->fx, fy = P->fy;
    const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
    A[0] = -fx * Zinv; A[1] = 0; A[2] = fx * X * Zinv2; A[3] = fx * X * Y * Zinv2; A[4] = -fx - fx * X * X * Zinv2; A[5] = fx * Y * Zinv;
    A[6] = 0; A[7] = -fy * Zinv; A[8] = fy * Y * Zinv2; A[9] = fy + fy * Y * Y * Zinv2; A[10] = -fy * X * Y * Zinv2; A[11] = -fy * X * Zinv;
    double RR[9];
    mat3_mul(P->ext.R, T->R, RR);
    for (int r = 0; r < 2; r++)
        for (int c = 0; c < 3; c++) B[3 * r + c] = A[6 * r] * RR[c] + A[6 * r + 1] * RR[3 + c] + A[6 * r + 2] * RR[6 + c];
}
void orc_ba_edge_jacobians(const double pose7[7], const double pt[3], const double K[4], const double ext7[7], double A[12],
                           double B[6]) { /* test hook */
    ba_problem P;
    memset(&P, 0, sizeof(P));
    se3 T;
    se3_from7(pose7, &T);
    se3_from7(ext7, &P.ext);
    int32_t z = 0;
    double pp[3] = {pt[0], pt[1], pt[2]};
    P.pose = &T; P.pt = pp; P.op = &z; P.ol = &z; P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    edge_jacobians(&P, 0, A, B);
}
void orc_ba_edge_error(const double pose7[7], const double pt[3], const double uv[2], const double K[4],
                       const double ext7[7], double err[2]) { /* test hook */
    ba_problem P;
    memset(&P, 0, sizeof(P));
    se3 T;
    se3_from7(pose7, &T);
    se3_from7(ext7, &P.ext);
    int32_t z = 0;
    P.op = &z; P.ol = &z; P.uv = uv; P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    edge_error(&P, &T, pt, 0, err);
}
/* VertexPose::oplusImpl test hook: pose7 <- exp(d) * pose7 */
void orc_pose_oplus(double pose7[7], const double d[6]) {
    se3 T, E, N;
    se3_from7(pose7, &T);
    orc_se3_exp(d, E.R, E.t);
    se3_mul(&E, &T, &N);
    se3_to7(&N, pose7);
}

/* RobustKernelHuber::robustify */
static void huber(double e2, double delta, double rho[2]) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.0; }
    else { const double s = sqrt(e2); rho[0] = 2 * s * delta - dsqr; rho[1] = delta / s; }
}

/* computeActiveErrors + activeRobustChi2 */
static double compute_errors(ba_problem *P, const se3 *pose, const double *pt) {
    double chi = 0;
    for (int e = 0; e < P->n_obs; e++) {
        edge_error(P, pose, pt, e, P->err + 2 * e);
        double rho[2];
        huber(P->err[2 * e] * P->err[2 * e] + P->err[2 * e + 1] * P->err[2 * e + 1], P->delta, rho);
        chi += rho[0];
    }
    return chi;
}

/* dense Cholesky solve of the reduced system, in place; returns 0 if not positive definite */
static int chol_solve(double *S, double *b, int n) {
    for (int j = 0; j < n; j++) {
        double d = S[j * n + j];
        for (int k = 0; k < j; k++) d -= S[j * n + k] * S[j * n + k];
        if (!(d > 0)) return 0;
        d = sqrt(d);
        S[j * n + j] = d;
        for (int i = j + 1; i < n; i++) {
            double v = S[i * n + j];
            for (int k = 0; k < j; k++) v -= S[i * n + k] * S[j * n + k];
            S[i * n + j] = v / d;
        }
    }
    for (int i = 0; i < n; i++) { double v = b[i]; for (int k = 0; k < i; k++) v -= S[i * n + k] * b[k]; b[i] = v / S[i * n + i]; }
    for (int i = n - 1; i >= 0; i--) { double v = b[i]; for (int k = i + 1; k < n; k++) v -= S[k * n + i] * b[k]; b[i] = v / S[i * n + i]; }
    return 1;
}

static int inv3_sym(const double H[9], double Dinv[9]) {
    const double a = H[0], b = H[1], c = H[2], d = H[4], e = H[5], f = H[8];
    const double c00 = d * f - e * e, c01 = c * e - b * f, c02 = b * e - c * d;
    const double det = a * c00 + b * c01 + c * c02;
    if (det == 0 || det != det) return 0;
    const double id = 1.0 / det;
    Dinv[0] = c00 * id; Dinv[1] = c01 * id; Dinv[2] = c02 * id;
    Dinv[3] = Dinv[1]; Dinv[4] = (a * f - c * c) * id; Dinv[5] = (b * c - a * e) * id;
    Dinv[6] = Dinv[2]; Dinv[7] = Dinv[5]; Dinv[8] = (a * d - b * b) * id;
    return 1;
}

/* One optimizer.optimize(iters) call with OptimizationAlgorithmLevenberg + BlockSolver_6_3 (Schur). */
static int lm_optimize(ba_problem *P, int iters, int *lm_iters_done) {
    const int np = P->n_poses, nl = P->n_points, ne = P->n_obs, n6 = 6 * np;
    double *Hpp = (double *)calloc((size_t)n6 * n6, sizeof(double)); /* block diagonal, stored dense */
    double *bp = (double *)calloc((size_t)n6, sizeof(double));
    double *Hll = (double *)calloc((size_t)nl * 9, sizeof(double));
    double *bl = (double *)calloc((size_t)nl * 3, sizeof(double));
    double *Hpl = (double *)calloc((size_t)ne * 18, sizeof(double)); /* per edge 6x3 */
    double *S = (double *)malloc(sizeof(double) * (size_t)n6 * n6);
    double *xs = (double *)malloc(sizeof(double) * (size_t)n6);
    double *xl = (double *)malloc(sizeof(double) * (size_t)nl * 3);
    double *Dinv = (double *)malloc(sizeof(double) * (size_t)nl * 9);
    se3 *pose_bak = (se3 *)malloc(sizeof(se3) * (size_t)np);
    double *pt_bak = (double *)malloc(sizeof(double) * (size_t)nl * 3);
    double lambda = 0, ni = 2;
    int done = 0, terminated = 0;
    for (int it = 0; it < iters && !terminated; it++) {
        double currentChi = compute_errors(P, P->pose, P->pt);
        /* buildSystem: linearizeOplus + constructQuadraticForm on every edge */
        memset(Hpp, 0, sizeof(double) * (size_t)n6 * n6); memset(bp, 0, sizeof(double) * (size_t)n6);
        memset(Hll, 0, sizeof(double) * (size_t)nl * 9); memset(bl, 0, sizeof(double) * (size_t)nl * 3);
        for (int e = 0; e < ne; e++) {
            double A[12], B[6], rho[2];
            edge_jacobians(P, e, A, B);
            const double *r = P->err + 2 * e;
            huber(r[0] * r[0] + r[1] * r[1], P->delta, rho);
            const double w = rho[1];
            const int i = P->op[e], j = P->ol[e];
            for (int a = 0; a < 6; a++) {
                bp[6 * i + a] += -w * (A[a] * r[0] + A[6 + a] * r[1]);
                for (int c = 0; c < 6; c++) Hpp[(6 * i + a) * n6 + 6 * i + c] += w * (A[a] * A[c] + A[6 + a] * A[6 + c]);
            }
            if (!P->fixed[j]) {
                for (int a = 0; a < 3; a++) {
                    bl[3 * j + a] += -w * (B[a] * r[0] + B[3 + a] * r[1]);
                    for (int c = 0; c < 3; c++) Hll[9 * j + 3 * a + c] += w * (B[a] * B[c] + B[3 + a] * B[3 + c]);
                }
                for (int a = 0; a < 6; a++)
                    for (int c = 0; c < 3; c++) Hpl[18 * e + 3 * a + c] = w * (A[a] * B[c] + A[6 + a] * B[3 + c]);
            }
        }
        if (it == 0) { /* computeLambdaInit: tau * max diagonal entry over all free vertices */
            double mx = 0;
            for (int a = 0; a < n6; a++) if (fabs(Hpp[a * n6 + a]) > mx) mx = fabs(Hpp[a * n6 + a]);
            for (int j = 0; j < nl; j++) if (!P->fixed[j]) for (int a = 0; a < 3; a++) if (fabs(Hll[9 * j + 4 * a]) > mx) mx = fabs(Hll[9 * j + 4 * a]);
            lambda = 1e-5 * mx;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            memcpy(pose_bak, P->pose, sizeof(se3) * (size_t)np);   /* push() */
            memcpy(pt_bak, P->pt, sizeof(double) * (size_t)nl * 3);
            /* solve with lambda on every diagonal entry: Schur complement on the landmarks */
            memcpy(S, Hpp, sizeof(double) * (size_t)n6 * n6);
            for (int a = 0; a < n6; a++) { S[a * n6 + a] += lambda; xs[a] = bp[a]; }
            int ok = 1;
            for (int j = 0; j < nl; j++) {
                if (P->fixed[j]) continue;
                double D[9];
                memcpy(D, Hll + 9 * j, sizeof(D));
                D[0] += lambda; D[4] += lambda; D[8] += lambda;
                if (!inv3_sym(D, Dinv + 9 * j)) ok = 0;
            }
            /* observations grouped by landmark would be faster; the oracle just scans */
            for (int e1 = 0; e1 < ne && ok; e1++) {
                const int j = P->ol[e1], i1 = P->op[e1];
                if (P->fixed[j]) continue;
                double BD[18]; /* Hpl_e1 * Dinv_j */
                for (int a = 0; a < 6; a++)
                    for (int c = 0; c < 3; c++)
                        BD[3 * a + c] = Hpl[18 * e1 + 3 * a] * Dinv[9 * j + c] + Hpl[18 * e1 + 3 * a + 1] * Dinv[9 * j + 3 + c] + Hpl[18 * e1 + 3 * a + 2] * Dinv[9 * j + 6 + c];
                for (int a = 0; a < 6; a++) xs[6 * i1 + a] -= BD[3 * a] * bl[3 * j] + BD[3 * a + 1] * bl[3 * j + 1] + BD[3 * a + 2] * bl[3 * j + 2];
                for (int e2 = 0; e2 < ne; e2++) {
                    if (P->ol[e2] != j) continue;
                    const int i2 = P->op[e2];
                    for (int a = 0; a < 6; a++)
                        for (int c = 0; c < 6; c++)
                            S[(6 * i1 + a) * n6 + 6 * i2 + c] -= BD[3 * a] * Hpl[18 * e2 + 3 * c] + BD[3 * a + 1] * Hpl[18 * e2 + 3 * c + 1] + BD[3 * a + 2] * Hpl[18 * e2 + 3 * c + 2];
                }
            }
            if (ok) ok = chol_solve(S, xs, n6);
            if (ok) {
                for (int j = 0; j < nl; j++) {
                    xl[3 * j] = xl[3 * j + 1] = xl[3 * j + 2] = 0;
                    if (P->fixed[j]) continue;
                    xl[3 * j] = bl[3 * j]; xl[3 * j + 1] = bl[3 * j + 1]; xl[3 * j + 2] = bl[3 * j + 2];
                }
                for (int e = 0; e < ne; e++) {
                    const int j = P->ol[e], i = P->op[e];
                    if (P->fixed[j]) continue;
                    for (int c = 0; c < 3; c++)
                        for (int a = 0; a < 6; a++) xl[3 * j + c] -= Hpl[18 * e + 3 * a + c] * xs[6 * i + a];
                }
                for (int j = 0; j < nl; j++) {
                    if (P->fixed[j]) continue;
                    double v[3] = {xl[3 * j], xl[3 * j + 1], xl[3 * j + 2]};
                    for (int a = 0; a < 3; a++) xl[3 * j + a] = Dinv[9 * j + 3 * a] * v[0] + Dinv[9 * j + 3 * a + 1] * v[1] + Dinv[9 * j + 3 * a + 2] * v[2];
                }
                /* update(): oplus on every free vertex */
                for (int i = 0; i < np; i++) {
                    se3 E, N;
                    orc_se3_exp(xs + 6 * i, E.R, E.t);
                    se3_mul(&E, &P->pose[i], &N);
                    P->pose[i] = N;
                }
                for (int j = 0; j < nl; j++)
                    if (!P->fixed[j]) for (int a = 0; a < 3; a++) P->pt[3 * j + a] += xl[3 * j + a];
            }
            double tempChi = compute_errors(P, P->pose, P->pt);
            if (!ok) tempChi = 1.7976931348623157e308;
            rho = currentChi - tempChi;
            double scale = 1e-3; /* computeScale() + 1e-3 */
            if (ok) {
                for (int a = 0; a < n6; a++) scale += xs[a] * (lambda * xs[a] + bp[a]);
                for (int j = 0; j < nl; j++)
                    if (!P->fixed[j]) for (int a = 0; a < 3; a++) scale += xl[3 * j + a] * (lambda * xl[3 * j + a] + bl[3 * j + a]);
            }
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                if (alpha > 2. / 3.) alpha = 2. / 3.;
                double sf = alpha > 1. / 3. ? alpha : 1. / 3.;
                lambda *= sf;
                ni = 2;
                currentChi = tempChi; /* discardTop() */
            } else {
                lambda *= ni;
                ni *= 2;
                memcpy(P->pose, pose_bak, sizeof(se3) * (size_t)np); /* pop(); the edge errors stay those of the trial */
                memcpy(P->pt, pt_bak, sizeof(double) * (size_t)nl * 3);
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        done++;
        if (qmax == 10 || rho == 0) terminated = 1;
    }
    if (lm_iters_done) *lm_iters_done += done;
    free(Hpp); free(bp); free(Hll); free(bl); free(Hpl); free(S); free(xs); free(xl); free(Dinv); free(pose_bak); free(pt_bak);
    return done;
}

/* Backend::OptimizeActiveMap's solver part on one window.
 *   poses  [n_poses][7] (qx qy qz qw tx ty tz) in/out     points [n_points][3] in/out
 *   fixed  [n_points]   obs_pose/obs_point [n_obs]   uv [n_obs][2]   K = fx fy cx cy   ext[7]
 *   chi2_out [n_obs] (e'e of the last error evaluation), outlier_out [n_obs] (chi2 > chi2_th)
 *   info[4] = outer rounds run, LM iterations run, inliers, outliers
 * Returns 0, or -1 on bad input. */
int orc_ba_solve(int n_poses, int n_points, int n_obs, double *poses, double *points, const uint8_t *fixed,
                 const int32_t *obs_pose, const int32_t *obs_point, const double *uv, const double *K, const double *ext7,
                 double huber_delta, double chi2_th, int outer_max, int inner_iters, double *chi2_out, uint8_t *outlier_out,
                 int32_t *info) {
    if (n_poses < 1 || n_points < 0 || n_obs < 0) return -1;
    for (int e = 0; e < n_obs; e++)
        if (obs_pose[e] < 0 || obs_pose[e] >= n_poses || obs_point[e] < 0 || obs_point[e] >= n_points) return -1;
    ba_problem P;
    memset(&P, 0, sizeof(P));
    P.n_poses = n_poses; P.n_points = n_points; P.n_obs = n_obs;
    P.pose = (se3 *)malloc(sizeof(se3) * (size_t)n_poses);
    for (int i = 0; i < n_poses; i++) se3_from7(poses + 7 * i, &P.pose[i]);
    P.pt = points; P.fixed = fixed; P.op = obs_pose; P.ol = obs_point; P.uv = uv;
    P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
    se3_from7(ext7, &P.ext);
    P.delta = huber_delta;
    P.err = (double *)calloc((size_t)(n_obs > 0 ? n_obs : 1) * 2, sizeof(double));
    int rounds = 0, lm = 0, inl = 0, outl = 0, iteration = 0;
    while (iteration < outer_max) { /* src/backend.cpp:212-232 */
        lm_optimize(&P, inner_iters, &lm);
        rounds++;
        inl = outl = 0;
        for (int e = 0; e < n_obs; e++) {
            const double c = P.err[2 * e] * P.err[2 * e] + P.err[2 * e + 1] * P.err[2 * e + 1];
            if (c > chi2_th) outl++; else inl++;
        }
        const double ratio = inl / (double)(inl + outl);
        if (ratio > 0.5) break;
        iteration++;
    }
    for (int e = 0; e < n_obs; e++) {
        const double c = P.err[2 * e] * P.err[2 * e] + P.err[2 * e + 1] * P.err[2 * e + 1];
        if (chi2_out) chi2_out[e] = c;
        if (outlier_out) outlier_out[e] = c > chi2_th;
    }
    for (int i = 0; i < n_poses; i++) se3_to7(&P.pose[i], poses + 7 * i);
    if (info) { info[0] = rounds; info[1] = lm; info[2] = inl; info[3] = outl; }
    free(P.pose); free(P.err);
    return 0;
}

/* ==========================================================================================================
 * Pose-only optimisation: Frontend::EstimateCurrentPose (src/frontend.cpp:176-276) and
 * LoopClosing::OptimizeCurrentPose (src/loopclosing.cpp:339-433) — SURVEY §8(f) "next" row 1.
 * One VertexPose, one EdgeProjectionPoseOnly per matched map point (include/myslam/g2o_types.h:63-102,
 * no camera extrinsics), information I2, RobustKernelHuber with g2o's default delta 1.0, LinearSolverDense.
 * `pre_rounds` un-classified optimize(inner) calls first (loop closing: 1, front end: 0), then `rounds`
 * (4) rounds of { initializeOptimization(level 0); optimize(inner); classify }: an edge whose chi2 exceeds
 * chi2_th becomes an outlier (level 1, left out of the next round), otherwise an inlier; outliers of the
 * previous round are re-evaluated at the new pose first; after round rounds-2 the robust kernels are removed.
 * chi2() of an active edge is that of g2o's last error evaluation (the last LM trial).
 * ========================================================================================================== */
static void po_error(const se3 *T, const double *pt, const double *uv, const double *K, double err[2]) {
    double pc[3];
    se3_apply(T, pt, pc);
    const double px = K[0] * pc[0] + K[2] * pc[2], py = K[1] * pc[1] + K[3] * pc[2];
    err[0] = uv[0] - px / pc[2];
    err[1] = uv[1] - py / pc[2];
}

static double po_active_errors(const se3 *T, int n, const double *pts, const double *uv, const double *K,
                               const uint8_t *level, int robust, double delta, double *err) {
    double chi = 0;
    for (int e = 0; e < n; e++) {
        if (level[e]) continue;
        po_error(T, pts + 3 * e, uv + 2 * e, K, err + 2 * e);
        const double e2 = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1];
        if (robust) { double rho[2]; huber(e2, delta, rho); chi += rho[0]; }
        else chi += e2;
    }
    return chi;
}

static void po_optimize(se3 *T, int n, const double *pts, const double *uv, const double *K, const uint8_t *level,
                        int robust, double delta, int iters, double *err, int *lm_done) {
    int n_active = 0;
    for (int e = 0; e < n; e++) n_active += !level[e];
    if (n_active == 0) return; /* g2o: "0 vertices to optimize" -> optimize() returns without touching anything */
    double lambda = 0, ni = 2;
    int terminated = 0;
    for (int it = 0; it < iters && !terminated; it++) {
        double currentChi = po_active_errors(T, n, pts, uv, K, level, robust, delta, err);
        double H[36], b[6];
        memset(H, 0, sizeof(H)); memset(b, 0, sizeof(b));
        for (int e = 0; e < n; e++) {
            if (level[e]) continue;
            double pc[3];
            se3_apply(T, pts + 3 * e, pc);
            const double X = pc[0], Y = pc[1], Z = pc[2], fx = K[0], fy = K[1];
            const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
            const double A[12] = {-fx * Zinv, 0, fx * X * Zinv2, fx * X * Y * Zinv2, -fx - fx * X * X * Zinv2, fx * Y * Zinv,
                                  0, -fy * Zinv, fy * Y * Zinv2, fy + fy * Y * Y * Zinv2, -fy * X * Y * Zinv2, -fy * X * Zinv};
            const double *r = err + 2 * e;
            double w = 1.0;
            if (robust) { double rho[2]; huber(r[0] * r[0] + r[1] * r[1], delta, rho); w = rho[1]; }
            for (int a = 0; a < 6; a++) {
                b[a] += -w * (A[a] * r[0] + A[6 + a] * r[1]);
                for (int c = 0; c < 6; c++) H[6 * a + c] += w * (A[a] * A[c] + A[6 + a] * A[6 + c]);
            }
        }
        if (it == 0) {
            double mx = 0;
            for (int a = 0; a < 6; a++) if (fabs(H[7 * a]) > mx) mx = fabs(H[7 * a]);
            lambda = 1e-5 * mx;
            ni = 2;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const se3 bak = *T;
            double S[36], x[6];
            memcpy(S, H, sizeof(S)); memcpy(x, b, sizeof(x));
            for (int a = 0; a < 6; a++) S[7 * a] += lambda;
            int ok = chol_solve(S, x, 6);
            if (ok) {
                se3 E, N;
                orc_se3_exp(x, E.R, E.t);
                se3_mul(&E, T, &N);
                *T = N;
            }
            double tempChi = po_active_errors(T, n, pts, uv, K, level, robust, delta, err);
            if (!ok) tempChi = 1.7976931348623157e308;
            double scale = 1e-3;
            if (ok) for (int a = 0; a < 6; a++) scale += x[a] * (lambda * x[a] + b[a]);
            rho = (currentChi - tempChi) / scale;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                if (alpha > 2. / 3.) alpha = 2. / 3.;
                lambda *= alpha > 1. / 3. ? alpha : 1. / 3.;
                ni = 2;
                currentChi = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                *T = bak;
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        (*lm_done)++;
        if (qmax == 10 || rho == 0) terminated = 1;
    }
}

/* pose7 in/out; points [n][3] (world), uv [n][2]; outlier_out [n]; info[4] = inliers, LM iterations, rounds, 0 */
int orc_pose_only_solve(int n, double *pose7, const double *points, const double *uv, const double *K, double huber_delta,
                        double chi2_th, int pre_rounds, int rounds, int inner_iters, uint8_t *outlier_out, int32_t *info) {
    if (n < 0) return -1;
    se3 T;
    se3_from7(pose7, &T);
    uint8_t *level = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    uint8_t *is_out = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    double *err = (double *)calloc((size_t)(n > 0 ? n : 1) * 2, sizeof(double));
    int robust = 1, lm = 0, n_out = 0;
    for (int r = 0; r < pre_rounds; r++) po_optimize(&T, n, points, uv, K, level, robust, huber_delta, inner_iters, err, &lm);
    for (int round = 0; round < rounds; round++) {
        po_optimize(&T, n, points, uv, K, level, robust, huber_delta, inner_iters, err, &lm);
        n_out = 0;
        for (int e = 0; e < n; e++) {
            if (is_out[e]) po_error(&T, points + 3 * e, uv + 2 * e, K, err + 2 * e);
            const double c = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1];
            if (c > chi2_th) { is_out[e] = 1; level[e] = 1; n_out++; }
            else { is_out[e] = 0; level[e] = 0; }
        }
        if (round == rounds - 2) robust = 0;
    }
    se3_to7(&T, pose7);
    for (int e = 0; e < n; e++) outlier_out[e] = is_out[e];
    if (info) { info[0] = n - n_out; info[1] = lm; info[2] = pre_rounds + rounds; info[3] = 0; }
    free(level); free(is_out); free(err);
    return 0;
}
