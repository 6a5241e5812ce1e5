// ba.cu — batched sliding-window local bundle adjustment on B200 (sm_100a), double precision.
//
// Replaces the solver inside Backend::OptimizeActiveMap (reference src/backend.cpp:126-269): g2o's
// Levenberg-Marquardt over BlockSolver_6_3 with the landmarks marginalised (Schur complement), the
// reference's EdgeProjection / VertexPose / VertexXYZ arithmetic (include/myslam/g2o_types.h:25-59,
// 106-153), Huber kernel (delta 5.991) and the outer "up to 5 x optimize(10)" loop (:212-232).
//
// One CTA solves one window from start to finish (all LM iterations and trials inside ONE launch, no
// host round trip); a launch solves a batch of independent windows.  Nothing is stored per edge: the
// 2x6 / 2x3 Jacobians are ~60 flops and are recomputed wherever they are needed, so the working set
// per window is the 6x6 pose blocks and the reduced (6P)^2 system in shared memory plus 18 doubles
// per landmark in L2.  Every sum runs in a fixed order (no floating-point atomics): results are
// bit-reproducible run to run.  The only index structure is edge_of[landmark][pose] (+ a chain through next_dup for the rare
// window in which one keyframe observes a landmark more than once, as LoopLocalFusion can produce: src/loopclosing.cpp:478-505).
//
// All arithmetic is fp64 like g2o's: the window has no fixed pose, its gauge is held only by the LM
// damping and the fixed landmarks, and fp32 normal equations do not keep the 1e-4 parity bound there.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "se3.cuh"

#ifndef BA_THREADS
#define BA_THREADS 256
#endif
#ifndef BA_MIN_BLOCKS
#define BA_MIN_BLOCKS 1   // CTAs per SM the register budget is held to (launch bounds)
#endif
#define BA_MAX_POSES 16
#define LM_STRIDE 20   // doubles per landmark record: 160 bytes, so that bl and Dinv start on 16-byte boundaries (double2 loads)
#ifdef BA_PROFILE
__device__ long long g_ba_prof[16];
#define BA_T(i) do { __syncthreads(); if (threadIdx.x == 0 && blockIdx.x == 0) { long long t_ = clock64(); g_ba_prof[i] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define BA_T(i) ((void)0)
#endif

struct sb_ba {
    int device, max_windows, max_poses, max_points, max_obs;
    cudaStream_t stream, own_stream;
    const int32_t *pending_info;  // host `info` of the batch in flight (sb_ba_submit .. sb_ba_wait), else null
    int pending_windows;
    cudaEvent_t done;             // recorded behind the batch's last copy: sb_ba_wait waits for THIS batch only, not for
                                  // whatever else shares the stream
    // device copies of the batch (host-pointer entry point) and per-window workspace
    int32_t *d_np, *d_nl, *d_ne, *d_info;
    double *d_poses, *d_points, *d_uv, *d_chi2;
    uint8_t *d_fixed, *d_outlier;
    int32_t *d_op, *d_ol;
    int32_t *d_edge_of;  // [W][ML][MP]: first edge (lowest index) of (landmark, pose), -1 if none
    int32_t *d_next;     // [W][MO]: next edge with the same (landmark, pose), ascending, -1 at the end
    double *d_lm;        // [W][ML][18]: Hll(6) bl(3) Dinv(6) xl(3)
    double *d_ptbak;     // [W][ML][3]
    double *d_err;       // [W][2][MO][2]
    double *d_hpl;       // [W][MO][18]: w A^T B of every edge to a free landmark (the Hpl block), rebuilt each iteration
    double *d_ybd;       // [W][MO][10]: per edge the landmark's share w B^T B (6), -w B^T r (3) of the linearised system
    int2 *d_plist;       // [W][MP][ML]: per pose the compact list of (first edge, landmark) it observes — the lanes of the build pass
    int4 *d_pairs;       // [W][MP (MP + 1) / 2][ML]: per pose-block pair (i1 <= i2) the (edge to i1, edge to i2, landmark) of every free
                         //                           landmark both observe — the structure of the Schur complement, built once
};

struct BaArgs {
    const int32_t *np, *nl, *ne;
    double *poses;           // [W][MP][7] in/out
    double *points;          // [W][ML][3] in/out
    const uint8_t *fixed;    // [W][ML]
    const int32_t *op, *ol;  // [W][MO]
    const double *uv;        // [W][MO][2]
    double *chi2;            // [W][MO] out
    uint8_t *outlier;        // [W][MO] out
    int32_t *info;           // [W][4] out: outer rounds, LM iterations, inliers, outliers (or -1: bad input)
    int32_t *edge_of, *next_dup;
    double *lm, *ptbak, *err, *hpl, *ybd;
    int4 *pairs;
    int2 *plist;             // [W][MP][ML]: per pose the (edge, landmark) of every landmark it observes, ascending landmark
    int MP, ML, MO;
    double fx, fy, cx, cy;
    double extR[9], extT[3];
    double delta, chi2_th;
    int outer_max, inner_iters;
};

// RobustKernelHuber::robustify -> rho(e2) and rho'(e2)
static __device__ __forceinline__ void huber(double e2, double delta, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho0 = e2; rho1 = 1.0; }
    else { const double s = sqrt(e2); rho0 = 2 * s * delta - dsqr; rho1 = delta / s; }
}

struct EdgeCtx {
    const BaArgs *a;
    const double *Rt;   // shared: [MP][12]
    const double *pts;  // global: [ML][3]
    const double *uv;   // global: [MO][2]
    const int32_t *op, *ol;
};

// camera-frame point of edge e:  ext * (T * p)
static __device__ __forceinline__ void edge_cam(const EdgeCtx &c, int i, int j, double *pe, double *RR) {
    const double *T = c.Rt + 12 * i;
    const double *p = c.pts + 3 * j;
    const double px = p[0], py = p[1], pz = p[2];
    double pc[3];
#pragma unroll
    for (int r = 0; r < 3; r++) pc[r] = T[3 * r] * px + T[3 * r + 1] * py + T[3 * r + 2] * pz + T[9 + r];
    const BaArgs &a = *c.a;
#pragma unroll
    for (int r = 0; r < 3; r++) pe[r] = a.extR[3 * r] * pc[0] + a.extR[3 * r + 1] * pc[1] + a.extR[3 * r + 2] * pc[2] + a.extT[r];
    if (RR) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int k = 0; k < 3; k++) RR[3 * r + k] = a.extR[3 * r] * T[k] + a.extR[3 * r + 1] * T[3 + k] + a.extR[3 * r + 2] * T[6 + k];
    }
}

// EdgeProjection::computeError (g2o_types.h:115-122)
static __device__ __forceinline__ void edge_error(const EdgeCtx &c, int e, double *r) {
    double pe[3];
    edge_cam(c, c.op[e], c.ol[e], pe, nullptr);
    const BaArgs &a = *c.a;
    const double px = a.fx * pe[0] + a.cx * pe[2], py = a.fy * pe[1] + a.cy * pe[2];
    r[0] = c.uv[2 * e] - px / pe[2];
    r[1] = c.uv[2 * e + 1] - py / pe[2];
}

// EdgeProjection::linearizeOplus (:124-144) + Huber weight of the stored error
// (i, j) = the edge's pose and landmark (the caller knows them: no op[e] / ol[e] load level); RR = extR * R_i, the same for every
// edge of pose i (the caller forms it once per pose with pose_RR)
static __device__ __forceinline__ void pose_RR(const EdgeCtx &c, int i, double *RR) {
    const double *T = c.Rt + 12 * i;
    const BaArgs &a = *c.a;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int k = 0; k < 3; k++) RR[3 * r + k] = a.extR[3 * r] * T[k] + a.extR[3 * r + 1] * T[3 + k] + a.extR[3 * r + 2] * T[6 + k];
}
static __device__ __forceinline__ void edge_lin(const EdgeCtx &c, int e, int i, int j, const double *RR, const double *err, double *A, double *B,
                                                double *r, double &w) {
    double pe[3];
    edge_cam(c, i, j, pe, nullptr);
    const BaArgs &a = *c.a;
    const double X = pe[0], Y = pe[1], Z = pe[2];
    const double Zinv = 1.0 / (Z + 1e-18), Zinv2 = Zinv * Zinv;
    A[0] = -a.fx * Zinv; A[1] = 0; A[2] = a.fx * X * Zinv2; A[3] = a.fx * X * Y * Zinv2; A[4] = -a.fx - a.fx * X * X * Zinv2; A[5] = a.fx * Y * Zinv;
    A[6] = 0; A[7] = -a.fy * Zinv; A[8] = a.fy * Y * Zinv2; A[9] = a.fy + a.fy * Y * Y * Zinv2; A[10] = -a.fy * X * Y * Zinv2; A[11] = -a.fy * X * Zinv;
#pragma unroll
    for (int rr = 0; rr < 2; rr++)
#pragma unroll
        for (int k = 0; k < 3; k++) B[3 * rr + k] = A[6 * rr] * RR[k] + A[6 * rr + 1] * RR[3 + k] + A[6 * rr + 2] * RR[6 + k];
    r[0] = err[2 * e];
    r[1] = err[2 * e + 1];
    double rho0;
    huber(r[0] * r[0] + r[1] * r[1], a.delta, rho0, w);
}

// ---- warp reduction of 32 values at once: after the call lane l holds the warp-wide sum of v[l].  A butterfly that
//      halves the number of live values per stage: 31 shuffles + adds instead of 32 x 5, in a fixed order.
static __device__ __forceinline__ double warp_reduce32(double (&v)[32], int lane) {
#pragma unroll
    for (int n = 16; n >= 1; n >>= 1) {
        const bool up = lane & n;
#pragma unroll
        for (int k = 0; k < n; k++) {
            const double keep = up ? v[k + n] : v[k], send = up ? v[k] : v[k + n];
            v[k] = keep + __shfl_xor_sync(0xffffffffu, send, n);
        }
    }
    return v[0];
}

// ---- block-wide deterministic reductions ----------------------------------------------------------------
static __device__ double block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double s = 0;
    for (int k = 0; k < BA_THREADS / 32; k++) s += red[k];
    return s;
}
static __device__ double block_max(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    double s = red[0];
    for (int k = 1; k < BA_THREADS / 32; k++) s = fmax(s, red[k]);
    return s;
}

// computeActiveErrors + activeRobustChi2
// `err` always receives the errors; `err_lin` (nullable) a second copy that stays untouched by the LM
// trials: the Huber weights of the linearised system belong to the state it was linearised at.
static __device__ double compute_errors(const EdgeCtx &c, int ne, double *err, double *err_lin, double *red) {
    double chi = 0;
    for (int e = threadIdx.x; e < ne; e += BA_THREADS) {
        double r[2], rho0, rho1;
        edge_error(c, e, r);
        err[2 * e] = r[0];
        err[2 * e + 1] = r[1];
        if (err_lin) { err_lin[2 * e] = r[0]; err_lin[2 * e + 1] = r[1]; }
        huber(r[0] * r[0] + r[1] * r[1], c.a->delta, rho0, rho1);
        chi += rho0;
    }
    return block_sum(chi, red);
}

__global__ void __launch_bounds__(BA_THREADS, BA_MIN_BLOCKS) k_ba_solve(const __grid_constant__ BaArgs a) {
    extern __shared__ __align__(16) double sm[];
    const int w = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int np = a.np[w], nl = a.nl[w], ne = a.ne[w];
    const int MP = a.MP;
    const int n6 = 6 * np;
    double *Rt = sm;                    // [MP][12]
    double *Rtb = Rt + 12 * MP;         // backup (push / pop)
    double *Hpp = Rtb + 12 * MP;        // [MP][36]
    double *bp = Hpp + 36 * MP;         // [6 MP]
    double *xs = bp + 6 * MP;           // [6 MP]
    double *S = xs + 6 * MP;            // [(6 MP)^2]
    double *red = S + 36 * MP * MP;     // [16]
    double *dinv = red + 16;            // [6 MP] reciprocal pivots of the Cholesky factor
    double *linv = dinv + 6 * MP;       // [MP][36] inverses of the factor's 6 x 6 diagonal blocks (lower triangular), for the substitutions
    int *pcnt = reinterpret_cast<int *>(linv + 36 * MP);  // [MP (MP + 1) / 2] entries of each pair list
    __shared__ int s_bad, s_next, s_dup;
    __shared__ int plcnt[BA_MAX_POSES];   // entries of each pose list
    double *poses = a.poses + (size_t)w * MP * 7;
    double *pts = a.points + (size_t)w * a.ML * 3;
    const uint8_t *fixed = a.fixed + (size_t)w * a.ML;
    const int32_t *op = a.op + (size_t)w * a.MO, *ol = a.ol + (size_t)w * a.MO;
    int32_t *edge_of = a.edge_of + (size_t)w * a.ML * MP;
    int32_t *next_dup = a.next_dup + (size_t)w * a.MO;
    double *lm = a.lm + (size_t)w * a.ML * LM_STRIDE;   // per landmark: Hll(0..5) bl(6..8) - Dinv(10..15) xl(16..18) -
    double *ptb = a.ptbak + (size_t)w * a.ML * 3;
    double *hpl = a.hpl + (size_t)w * a.MO * 18;
    double *lmc = a.ybd + (size_t)w * a.MO * 10;   // per edge: the landmark's share of Hll and bl (10 doubles)
    int4 *pairs = a.pairs + (size_t)w * (MP * (MP + 1) / 2) * a.ML;
    int2 *plist = a.plist + (size_t)w * MP * a.ML;
    double *err = a.err + (size_t)w * a.MO * 4;   // errors of the last evaluation
    double *elin = err + (size_t)a.MO * 2;         // errors at the linearisation state
    int32_t *info = a.info + 4 * w;

    // ---- input checks and the edge_of[landmark][pose] table
    if (tid == 0) { s_bad = (np < 1 || np > MP || nl < 0 || nl > a.ML || ne < 0 || ne > a.MO) ? 1 : 0; s_dup = 0; }
    __syncthreads();
    if (!s_bad) {
        for (int k = tid; k < nl * MP; k += BA_THREADS) edge_of[k] = -1;
        __syncthreads();
        for (int e = tid; e < ne; e += BA_THREADS) {
            const int i = op[e], j = ol[e];
            if (i < 0 || i >= np || j < 0 || j >= nl) s_bad = 1;
            else if (atomicCAS(&edge_of[j * MP + i], -1, e) != -1) s_dup = 1;  // one keyframe observes a landmark more than once
        }
    }
    __syncthreads();
    if (!s_bad && s_dup) {
        // Rare (a window right after a loop fusion): g2o simply has several edges between the same two vertices.  Their
        // blocks add up, so the (landmark, pose) slot keeps its FIRST edge and the others hang off it in ascending edge
        // order — rebuilt serially so that the chains do not depend on which thread won the race above.
        for (int k = tid; k < nl * MP; k += BA_THREADS) edge_of[k] = -1;
        for (int e = tid; e < ne; e += BA_THREADS) next_dup[e] = -1;
        __syncthreads();
        if (tid == 0)
            for (int e = 0; e < ne; e++) {
                int *slot = &edge_of[ol[e] * MP + op[e]];
                if (*slot < 0) { *slot = e; continue; }
                int t = *slot;
                while (next_dup[t] >= 0) t = next_dup[t];
                next_dup[t] = e;
            }
        __syncthreads();
    }
    const bool has_dup = s_dup != 0;
    if (s_bad) {
        if (tid == 0) { info[0] = -1; info[1] = -s_bad; info[2] = info[3] = 0; }
        return;
    }
    for (int i = tid; i < np; i += BA_THREADS) {
        quat_to_R(poses + 7 * i, Rt + 12 * i);
        Rt[12 * i + 9] = poses[7 * i + 4]; Rt[12 * i + 10] = poses[7 * i + 5]; Rt[12 * i + 11] = poses[7 * i + 6];
    }
    // ---- structure of the Schur complement: for every pose-block pair the landmarks that couple them, compacted in
    //      ascending landmark order (one warp per pair; ballot compaction keeps it deterministic)
    const int nblk = np * (np + 1) / 2;
    for (int blk = wid; blk < nblk; blk += BA_THREADS / 32) {
        int i1 = 0, rem = blk;
        while (rem >= np - i1) { rem -= np - i1; i1++; }
        const int i2 = i1 + rem;
        int4 *pl = pairs + (size_t)blk * a.ML;
        int cnt = 0;
        for (int j0 = 0; j0 < nl; j0 += 32) {
            const int j = j0 + lane;
            int e1 = -1, e2 = -1;
            if (j < nl && !fixed[j]) { e1 = edge_of[j * MP + i1]; e2 = edge_of[j * MP + i2]; }
            const bool ok = e1 >= 0 && e2 >= 0;
            const unsigned bal = __ballot_sync(0xffffffffu, ok);
            if (ok) pl[cnt + __popc(bal & ((1u << lane) - 1u))] = make_int4(e1, e2, j, 0);
            cnt += __popc(bal);
        }
        if (lane == 0) pcnt[blk] = cnt;
    }
    // ... and per pose the landmarks it observes (fixed ones too), compacted the same way: the build pass walks these lists, so
    // every lane of a round holds an edge (by landmark index only 2 of 3 lanes would: a pose sees ~200 of the 300 landmarks)
    for (int i = wid; i < np; i += BA_THREADS / 32) {
        int2 *pl = plist + (size_t)i * a.ML;
        int cnt = 0;
        for (int j0 = 0; j0 < nl; j0 += 32) {
            const int j = j0 + lane;
            const int e = j < nl ? edge_of[j * MP + i] : -1;
            const unsigned bal = __ballot_sync(0xffffffffu, e >= 0);
            if (e >= 0) pl[cnt + __popc(bal & ((1u << lane) - 1u))] = make_int2(e, j);
            cnt += __popc(bal);
        }
        if (lane == 0) plcnt[i] = cnt;
    }
    __syncthreads();

    EdgeCtx c;
    c.a = &a; c.Rt = Rt; c.pts = pts; c.uv = a.uv + (size_t)w * a.MO * 2; c.op = op; c.ol = ol;

#ifdef BA_PROFILE
    long long t_prev = clock64();
#endif
    int rounds = 0, lm_total = 0, inl = 0, outl = 0;
    for (int outer = 0; outer < a.outer_max;) {  // src/backend.cpp:212-232
        // ================= optimizer.optimize(inner_iters): Levenberg-Marquardt =================
        double lambda = 0, ni = 2;
        bool terminated = false;
        for (int it = 0; it < a.inner_iters && !terminated; it++) {
            BA_T(0);
            double currentChi = compute_errors(c, ne, err, elin, red);
            BA_T(1);
            // ---- buildSystem.  Every edge is linearised ONCE, in the pose pass (one warp per pose, lanes over the landmarks it
            //      observes): pose block and gradient by warp reduction, and per edge to a free landmark the Hpl block w A^T B and
            //      the landmark's share (w B^T B, -w B^T r) for the landmark pass below.
            for (int i = wid; i < np; i += BA_THREADS / 32) {
                double h[21], g[6];
#pragma unroll
                for (int k = 0; k < 21; k++) h[k] = 0;
#pragma unroll
                for (int k = 0; k < 6; k++) g[k] = 0;
                const int2 *pl = plist + (size_t)i * a.ML;
                const int pn = plcnt[i];
                double RR[9];
                pose_RR(c, i, RR);
                int2 nxt = lane < pn ? pl[lane] : make_int2(0, 0);   // the list entry of the next round is fetched a round ahead
                for (int t = lane; t < pn; t += 32) {
                    const int2 ej = nxt;
                    if (t + 32 < pn) nxt = pl[t + 32];
                    const int e = ej.x, j = ej.y;
                    const bool fr = !fixed[j];
                    double hv[18], cn[9];   // Hpl block w A^T B (6x3) and the landmark's share (w B^T B: 00 01 02 11 12 22, -w B^T r)
#pragma unroll
                    for (int k2 = 0; k2 < 18; k2++) hv[k2] = 0;
#pragma unroll
                    for (int k2 = 0; k2 < 9; k2++) cn[k2] = 0;
                    for (int ee = e; ee >= 0; ee = has_dup ? next_dup[ee] : -1) {   // one edge, unless the keyframe sees the landmark twice
                        double A[12], B[6], r[2], wgt;
                        edge_lin(c, ee, i, j, RR, elin, A, B, r, wgt);
                        int k = 0;
#pragma unroll
                        for (int p = 0; p < 6; p++) {
                            g[p] += -wgt * (A[p] * r[0] + A[6 + p] * r[1]);
#pragma unroll
                            for (int q = p; q < 6; q++) h[k++] += wgt * (A[p] * A[q] + A[6 + p] * A[6 + q]);
                        }
                        if (fr) {
#pragma unroll
                            for (int p = 0; p < 6; p++)
#pragma unroll
                                for (int q = 0; q < 3; q++) hv[3 * p + q] += wgt * (A[p] * B[q] + A[6 + p] * B[3 + q]);
                            cn[0] += wgt * (B[0] * B[0] + B[3] * B[3]); cn[1] += wgt * (B[0] * B[1] + B[3] * B[4]);
                            cn[2] += wgt * (B[0] * B[2] + B[3] * B[5]); cn[3] += wgt * (B[1] * B[1] + B[4] * B[4]);
                            cn[4] += wgt * (B[1] * B[2] + B[4] * B[5]); cn[5] += wgt * (B[2] * B[2] + B[5] * B[5]);
                            cn[6] += -wgt * (B[0] * r[0] + B[3] * r[1]); cn[7] += -wgt * (B[1] * r[0] + B[4] * r[1]);
                            cn[8] += -wgt * (B[2] * r[0] + B[5] * r[1]);
                        }
                    }
                    if (fr) {   // stored in the slot of the first edge: 9 + 5 double2 stores
                        double2 *P = reinterpret_cast<double2 *>(hpl + 18 * e);
#pragma unroll
                        for (int k2 = 0; k2 < 9; k2++) P[k2] = make_double2(hv[2 * k2], hv[2 * k2 + 1]);
                        double2 *Cn = reinterpret_cast<double2 *>(lmc + 10 * e);
                        Cn[0] = make_double2(cn[0], cn[1]); Cn[1] = make_double2(cn[2], cn[3]); Cn[2] = make_double2(cn[4], cn[5]);
                        Cn[3] = make_double2(cn[6], cn[7]); Cn[4] = make_double2(cn[8], 0.0);
                    }
                }
#pragma unroll
                for (int k = 0; k < 21; k++)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) h[k] += __shfl_down_sync(0xffffffffu, h[k], o);
#pragma unroll
                for (int k = 0; k < 6; k++)
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) g[k] += __shfl_down_sync(0xffffffffu, g[k], o);
                if (lane == 0) {
                    int k = 0;
                    for (int p = 0; p < 6; p++) {
                        bp[6 * i + p] = g[p];
                        for (int q = p; q < 6; q++) { Hpp[36 * i + 6 * p + q] = h[k]; Hpp[36 * i + 6 * q + p] = h[k]; k++; }
                    }
                }
            }
            __syncthreads();
            BA_T(11);
            // Landmark blocks: one thread per free landmark sums the shares of its edges in ascending pose order (all edge
            // indices are fetched before the shares: two load levels per landmark instead of two per edge).
            // lm[j] = Hll(6: 00 01 02 11 12 22) bl(3) - Dinv(6) xl(3) -
            for (int j = tid; j < nl; j += BA_THREADS) {
                if (fixed[j]) continue;
                int es[BA_MAX_POSES];
#pragma unroll
                for (int i = 0; i < BA_MAX_POSES; i++) es[i] = i < np ? edge_of[j * MP + i] : -1;
                double H[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
#pragma unroll
                for (int i = 0; i < BA_MAX_POSES; i++) {
                    if (es[i] < 0) continue;
                    const double2 *Cn = reinterpret_cast<const double2 *>(lmc + 10 * es[i]);
                    const double2 c0 = Cn[0], c1 = Cn[1], c2 = Cn[2], c3 = Cn[3], c4 = Cn[4];
                    H[0] += c0.x; H[1] += c0.y; H[2] += c1.x; H[3] += c1.y; H[4] += c2.x; H[5] += c2.y;
                    g[0] += c3.x; g[1] += c3.y; g[2] += c4.x;
                }
                double *L = lm + LM_STRIDE * j;
#pragma unroll
                for (int k = 0; k < 6; k++) L[k] = H[k];
                L[6] = g[0]; L[7] = g[1]; L[8] = g[2];
            }
            __syncthreads();
            BA_T(2);
            if (it == 0) {  // computeLambdaInit: 1e-5 * largest diagonal entry of H over the free vertices
                double mx = 0;
                for (int k = tid; k < n6; k += BA_THREADS) mx = fmax(mx, fabs(Hpp[36 * (k / 6) + 7 * (k % 6)]));
                for (int j = tid; j < nl; j += BA_THREADS)
                    if (!fixed[j]) mx = fmax(mx, fmax(fabs(lm[LM_STRIDE * j]), fmax(fabs(lm[LM_STRIDE * j + 3]), fabs(lm[LM_STRIDE * j + 5]))));
                lambda = 1e-5 * block_max(mx, red);
                ni = 2;
            }
            double rho = 0;
            int qmax = 0;
            do {
                // ---- push(): backup of the state
                for (int k = tid; k < 12 * np; k += BA_THREADS) Rtb[k] = Rt[k];
                for (int k = tid; k < 3 * nl; k += BA_THREADS) ptb[k] = pts[k];
                // ---- (Hll + lambda I)^-1 per landmark
                int bad = 0;
                for (int j = tid; j < nl; j += BA_THREADS) {
                    if (fixed[j]) continue;
                    double *L = lm + LM_STRIDE * j;
                    const double A0 = L[0] + lambda, b = L[1], cc = L[2], d = L[3] + lambda, e = L[4], f = L[5] + lambda;
                    const double c00 = d * f - e * e, c01 = cc * e - b * f, c02 = b * e - cc * d;
                    const double det = A0 * c00 + b * c01 + cc * c02;
                    if (det == 0 || det != det) bad = 1;
                    const double id = 1.0 / det;
                    L[10] = c00 * id; L[11] = c01 * id; L[12] = c02 * id;
                    L[13] = (A0 * f - cc * cc) * id; L[14] = (b * cc - A0 * e) * id; L[15] = (A0 * d - b * b) * id;
                }
                int ok = !__syncthreads_or(bad);
                if (tid == 0) s_next = BA_THREADS / 32;
                __syncthreads();
                BA_T(3);
                // ---- Schur complement: S(i1,i2) = Hpp'(i1,i2) - sum_j Y(i1,j) Hpl(i2,j)^T over the pair's list, one warp per
                //      block (upper triangle, taken from a shared counter: the blocks differ in size), lanes over the list;
                //      the diagonal pass also reduces b.
                for (int blk = wid; blk < nblk && ok;) {
                    int i1 = 0, rem = blk;
                    while (rem >= np - i1) { rem -= np - i1; i1++; }
                    const int i2 = i1 + rem;
                    const int4 *pl = pairs + (size_t)blk * a.ML;
                    const int cnt = pcnt[blk];
                    // Y = Hpl(i1, j) Dinv_j is formed on the fly from the Hpl block and the landmark's Dinv (16-byte vector loads), so
                    // no Y array is written or read; a diagonal block (i1 == i2: the longest lists) is symmetric and needs only its
                    // 21 upper entries.  The branch is uniform over the block.
                    double acc[36], gb[6];
#pragma unroll
                    for (int k = 0; k < 36; k++) acc[k] = 0;
#pragma unroll
                    for (int k = 0; k < 6; k++) gb[k] = 0;
                    const bool diag = i1 == i2;
                    // the list entry of the NEXT round is fetched before this round's blocks are used: the entry -> blocks chain is
                    // two dependent global loads long, and the entry carries the landmark so that no third level (ol[e]) is needed
                    int4 nxt = lane < cnt ? pl[lane] : make_int4(0, 0, 0, 0);
                    for (int t = lane; t < cnt; t += 32) {
                        const int4 ee = nxt;
                        if (t + 32 < cnt) nxt = pl[t + 32];
                        const double *L = lm + LM_STRIDE * ee.z;
                        const double2 *P1 = reinterpret_cast<const double2 *>(hpl + 18 * ee.x);
                        const double2 d01 = *reinterpret_cast<const double2 *>(L + 10), d23 = *reinterpret_cast<const double2 *>(L + 12),
                                      d45 = *reinterpret_cast<const double2 *>(L + 14);
                        const double D0 = d01.x, D1 = d01.y, D2 = d23.x, D4 = d23.y, D5 = d45.x, D8 = d45.y;
                        double Pa[18], Y[18];
#pragma unroll
                        for (int k = 0; k < 9; k++) { const double2 v2 = P1[k]; Pa[2 * k] = v2.x; Pa[2 * k + 1] = v2.y; }
#pragma unroll
                        for (int p = 0; p < 6; p++) {
                            const double h0 = Pa[3 * p], h1 = Pa[3 * p + 1], h2 = Pa[3 * p + 2];
                            Y[3 * p] = h0 * D0 + h1 * D1 + h2 * D2;
                            Y[3 * p + 1] = h0 * D1 + h1 * D4 + h2 * D5;
                            Y[3 * p + 2] = h0 * D2 + h1 * D5 + h2 * D8;
                        }
                        if (diag) {   // Y Hpl^T with the same block: upper triangle only; and b_schur(i1) -= Y bl_j
                            const double2 b01 = *reinterpret_cast<const double2 *>(L + 6);
                            const double b0 = b01.x, b1 = b01.y, b2 = L[8];
#pragma unroll
                            for (int p = 0; p < 6; p++) {
                                gb[p] += Y[3 * p] * b0 + Y[3 * p + 1] * b1 + Y[3 * p + 2] * b2;
#pragma unroll
                                for (int q = p; q < 6; q++) acc[6 * p + q] += Y[3 * p] * Pa[3 * q] + Y[3 * p + 1] * Pa[3 * q + 1] + Y[3 * p + 2] * Pa[3 * q + 2];
                            }
                        } else {
                            const double2 *P2 = reinterpret_cast<const double2 *>(hpl + 18 * ee.y);
#pragma unroll
                            for (int m = 0; m < 3; m++) {   // block rows 2m and 2m + 1 of Hpl(i2, j)
                                const double2 ga = P2[3 * m], gb2 = P2[3 * m + 1], gc = P2[3 * m + 2];
#pragma unroll
                                for (int p = 0; p < 6; p++) {
                                    acc[6 * p + 2 * m] += Y[3 * p] * ga.x + Y[3 * p + 1] * ga.y + Y[3 * p + 2] * gb2.x;
                                    acc[6 * p + 2 * m + 1] += Y[3 * p] * gb2.y + Y[3 * p + 1] * gc.x + Y[3 * p + 2] * gc.y;
                                }
                            }
                        }
                    }
                    if (diag) {   // mirror the upper triangle (before the reduction: every lane holds its own partial block)
#pragma unroll
                        for (int p = 1; p < 6; p++)
#pragma unroll
                            for (int q = 0; q < p; q++) acc[6 * p + q] = acc[6 * q + p];
                    }
                    // entries 0..31 by the butterfly (lane l ends with entry l), 32..35 and b by plain trees
                    double head[32];
#pragma unroll
                    for (int k = 0; k < 32; k++) head[k] = acc[k];
                    const double mine = warp_reduce32(head, lane);
#pragma unroll
                    for (int k = 32; k < 36; k++)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) acc[k] += __shfl_xor_sync(0xffffffffu, acc[k], o);
                    {
                        const int p = lane / 6, q = lane - 6 * p;
                        double v = -mine;
                        if (i1 == i2) v += Hpp[36 * i1 + 6 * p + q] + (p == q ? lambda : 0.0);
                        S[(6 * i1 + p) * n6 + 6 * i2 + q] = v;
                        if (i1 != i2) S[(6 * i2 + q) * n6 + 6 * i1 + p] = v;  // a diagonal block holds both (p,q) and (q,p) itself
                    }
                    if (lane < 4) {
                        const int q = 2 + lane;  // entries 32..35 = row 5, columns 2..5
                        double v = -(lane == 0 ? acc[32] : lane == 1 ? acc[33] : lane == 2 ? acc[34] : acc[35]);
                        if (i1 == i2) v += Hpp[36 * i1 + 30 + q] + (q == 5 ? lambda : 0.0);
                        S[(6 * i1 + 5) * n6 + 6 * i2 + q] = v;
                        if (i1 != i2) S[(6 * i2 + q) * n6 + 6 * i1 + 5] = v;
                    }
                    if (i1 == i2) {
#pragma unroll
                        for (int k = 0; k < 6; k++)
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) gb[k] += __shfl_xor_sync(0xffffffffu, gb[k], o);
                        if (lane < 6) {
                            const double g = lane == 0 ? gb[0] : lane == 1 ? gb[1] : lane == 2 ? gb[2] : lane == 3 ? gb[3] : lane == 4 ? gb[4] : gb[5];
                            xs[6 * i1 + lane] = bp[6 * i1 + lane] - g;
                        }
                    }
                    blk = 0;
                    if (lane == 0) blk = atomicAdd(&s_next, 1);
                    blk = __shfl_sync(0xffffffffu, blk, 0);
                }
                __syncthreads();
                BA_T(4);
                // ---- Cholesky of the reduced system, right-looking by 6 x 6 pose blocks (two barriers per block column instead
                //      of one per column): every thread of the first three warps factors the diagonal block in registers (the
                //      same 21 values: no broadcast needed), thread t then solves row t of the panel below it, and the whole
                //      CTA applies the rank-6 update to the trailing lower triangle.  Column j of L goes to ROW j of the
                //      upper triangle (S[j][i], i > j: contiguous for the substitutions), the pivots are kept as
                //      reciprocals.  Per entry the subtractions happen in the same order as in the left-looking form.
                if (ok) {
                    int bad_pivot = 0;
                    if (tid == 0) s_bad = 0;
                    __syncthreads();
                    for (int J = 0; J < n6; J += 6) {
                        const int below = n6 - J - 6;  // rows under the diagonal block
                        if (tid < 96) {  // n6 - 6 <= 90 panel rows (16 poses)
                            double Lb[21], inv[6];  // lower triangle of the block, row-major packed: (r, c) at r (r + 1) / 2 + c
                            bool bad = false;
#pragma unroll
                            for (int r = 0; r < 6; r++)
#pragma unroll
                                for (int c = 0; c <= r; c++) Lb[r * (r + 1) / 2 + c] = S[(J + r) * n6 + J + c];
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                double d = Lb[c * (c + 1) / 2 + c];
#pragma unroll
                                for (int k = 0; k < c; k++) d -= Lb[c * (c + 1) / 2 + k] * Lb[c * (c + 1) / 2 + k];
                                if (!(d > 0)) bad = true;
                                inv[c] = rsqrt(d);
#pragma unroll
                                for (int r = c + 1; r < 6; r++) {
                                    double v = Lb[r * (r + 1) / 2 + c];
#pragma unroll
                                    for (int k = 0; k < c; k++) v -= Lb[r * (r + 1) / 2 + k] * Lb[c * (c + 1) / 2 + k];
                                    Lb[r * (r + 1) / 2 + c] = v * inv[c];
                                }
                            }
                            if (tid == 0) {
                                if (bad) s_bad = 1;
#pragma unroll
                                for (int c = 0; c < 6; c++) {
                                    dinv[J + c] = inv[c];
#pragma unroll
                                    for (int r = c + 1; r < 6; r++) S[(J + c) * n6 + J + r] = Lb[r * (r + 1) / 2 + c];
                                }
                                // M = L_JJ^-1 (lower triangular), column by column: the substitutions then advance a whole block per
                                // step (y_J = M r_J) instead of one unknown
                                double *M = linv + 6 * J;   // 36 doubles per block: J / 6 * 36
#pragma unroll
                                for (int c = 0; c < 6; c++) {
                                    double m[6];
                                    m[c] = inv[c];
#pragma unroll
                                    for (int r = c + 1; r < 6; r++) {
                                        double v = 0;
#pragma unroll
                                        for (int k = c; k < r; k++) v -= Lb[r * (r + 1) / 2 + k] * m[k];
                                        m[r] = v * inv[r];
                                    }
#pragma unroll
                                    for (int r = 0; r < 6; r++) M[6 * r + c] = r >= c ? m[r] : 0.0;
                                }
                            }
                            if (tid < below && !bad) {  // panel row i: x L_JJ^T = S(i, J..J+5)
                                const int i = J + 6 + tid;
                                double x[6];
#pragma unroll
                                for (int c = 0; c < 6; c++) {
                                    double v = S[i * n6 + J + c];
#pragma unroll
                                    for (int k = 0; k < c; k++) v -= x[k] * Lb[c * (c + 1) / 2 + k];
                                    x[c] = v * inv[c];
                                }
#pragma unroll
                                for (int c = 0; c < 6; c++) {
                                    S[i * n6 + J + c] = x[c];    // read by the trailing update
                                    S[(J + c) * n6 + i] = x[c];  // L^T for the substitutions
                                }
                            }
                        }
                        __syncthreads();
                        BA_T(8);
                        if (s_bad) { bad_pivot = 1; break; }
                        {
                            const int r0 = tid >> 3, c0 = tid & 7;
                            for (int i = J + 6 + r0; i < n6; i += BA_THREADS / 8) {
                                const double *li = S + i * n6 + J;
                                const double l0 = li[0], l1 = li[1], l2 = li[2], l3 = li[3], l4 = li[4], l5 = li[5];
                                for (int k = J + 6 + c0; k <= i; k += 8) {
                                    const double *lk = S + k * n6 + J;
                                    double v = S[i * n6 + k];
                                    v -= l0 * lk[0]; v -= l1 * lk[1]; v -= l2 * lk[2]; v -= l3 * lk[3]; v -= l4 * lk[4]; v -= l5 * lk[5];
                                    S[i * n6 + k] = v;
                                }
                            }
                        }
                        __syncthreads();
                        BA_T(9);
                    }
                    if (!bad_pivot && wid == 0) {
                        // substitutions by one warp, the unknowns in registers (n6 <= 96: three per lane), one 6 x 6 BLOCK per step:
                        // the block's six residuals are broadcast by shuffles, every lane forms y_J = M_J r_J (M_J = L_JJ^-1, 21
                        // multiply-adds) and subtracts L(i, J) y_J from the rows it owns below (forward) / above (backward) the
                        // block.  14 dependent steps instead of 168 (measured: the unknown-by-unknown form was 16 % of the kernel).
                        double x0 = lane < n6 ? xs[lane] : 0, x1 = lane + 32 < n6 ? xs[lane + 32] : 0, x2 = lane + 64 < n6 ? xs[lane + 64] : 0;
                        for (int B0 = 0; B0 < n6; B0 += 6) {  // forward: L y = b;  L(i, j) = S[j][i] for i > j
                            const double *M = linv + 6 * B0;
                            double r[6], y[6];
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                const int row = B0 + c;
                                r[c] = __shfl_sync(0xffffffffu, row < 32 ? x0 : row < 64 ? x1 : x2, row & 31);
                            }
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                double v = 0;
#pragma unroll
                                for (int d = 0; d <= c; d++) v += M[6 * c + d] * r[d];
                                y[c] = v;
                            }
                            // branch-free over the lane's three rows (three independent chains; a divergent form serialises them)
#pragma unroll
                            for (int part = 0; part < 3; part++) {
                                if (32 * part + 31 < B0 + 6 && 32 * part + 31 < B0) continue;   // warp-uniform: nothing of this part is at or below the block
                                const int i = lane + 32 * part;
                                double &x = part == 0 ? x0 : part == 1 ? x1 : x2;
                                const bool below = i >= B0 + 6 && i < n6;
                                const int ii = below ? i : 0;
                                double v = x;
#pragma unroll
                                for (int c = 0; c < 6; c++) v -= S[(B0 + c) * n6 + ii] * y[c];
                                x = below ? v : x;
#pragma unroll
                                for (int c = 0; c < 6; c++)
                                    if (i == B0 + c) x = y[c];
                            }
                        }
                        for (int B0 = n6 - 6; B0 >= 0; B0 -= 6) {  // backward: L^T x = y;  L(r, i) = S[r][i] for r > i
                            const double *M = linv + 6 * B0;
                            double r[6], y[6];
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                const int row = B0 + c;
                                r[c] = __shfl_sync(0xffffffffu, row < 32 ? x0 : row < 64 ? x1 : x2, row & 31);
                            }
#pragma unroll
                            for (int c = 0; c < 6; c++) {   // x_J = M^T r_J
                                double v = 0;
#pragma unroll
                                for (int d = c; d < 6; d++) v += M[6 * d + c] * r[d];
                                y[c] = v;
                            }
#pragma unroll
                            for (int part = 0; part < 3; part++) {
                                if (32 * part >= B0 + 6) continue;   // warp-uniform: this part lies below the block
                                const int i = lane + 32 * part;
                                double &x = part == 0 ? x0 : part == 1 ? x1 : x2;
                                const bool above = i < B0;
                                const int ii = above ? i : 0;
                                double v = x;
#pragma unroll
                                for (int c = 0; c < 6; c++) v -= S[(B0 + c) * n6 + ii] * y[c];
                                x = above ? v : x;
#pragma unroll
                                for (int c = 0; c < 6; c++)
                                    if (i == B0 + c) x = y[c];
                            }
                        }
                        if (lane < n6) xs[lane] = x0;
                        if (lane + 32 < n6) xs[lane + 32] = x1;
                        if (lane + 64 < n6) xs[lane + 64] = x2;
                    }
                    __syncthreads();
                    BA_T(10);
                    ok = !bad_pivot;
                    if (tid == 0) s_bad = 0;
                }
                BA_T(5);
                double scale = 0;
                if (ok) {
                    // ---- landmark increments: xl = Dinv (bl - sum_i Hpl(i,j)^T xs_i), then the state update
                    for (int j = tid; j < nl; j += BA_THREADS) {
                        if (fixed[j]) continue;
                        double *L = lm + LM_STRIDE * j;
                        double v[3] = {L[6], L[7], L[8]};
                        int es[BA_MAX_POSES];
#pragma unroll
                        for (int i = 0; i < BA_MAX_POSES; i++) es[i] = i < np ? edge_of[j * MP + i] : -1;
#pragma unroll
                        for (int i = 0; i < BA_MAX_POSES; i++) {
                            const int e = es[i];
                            if (e < 0) continue;
                            const double2 *P = reinterpret_cast<const double2 *>(hpl + 18 * e);  // v -= Hpl(i, j)^T xs_i
#pragma unroll
                            for (int m = 0; m < 3; m++) {
                                const double2 a2 = P[3 * m], b2 = P[3 * m + 1], c2 = P[3 * m + 2];
                                const double xa = xs[6 * i + 2 * m], xb = xs[6 * i + 2 * m + 1];
                                v[0] -= a2.x * xa; v[1] -= a2.y * xa; v[2] -= b2.x * xa;
                                v[0] -= b2.y * xb; v[1] -= c2.x * xb; v[2] -= c2.y * xb;
                            }
                        }
                        const double x0 = L[10] * v[0] + L[11] * v[1] + L[12] * v[2];
                        const double x1 = L[11] * v[0] + L[13] * v[1] + L[14] * v[2];
                        const double x2 = L[12] * v[0] + L[14] * v[1] + L[15] * v[2];
                        L[16] = x0; L[17] = x1; L[18] = x2;
                        scale += x0 * (lambda * x0 + L[6]) + x1 * (lambda * x1 + L[7]) + x2 * (lambda * x2 + L[8]);
                    }
                    for (int k = tid; k < n6; k += BA_THREADS) scale += xs[k] * (lambda * xs[k] + bp[k]);
                    scale = block_sum(scale, red);
                    // every Jacobian above used the un-updated state: only now move it
                    for (int j = tid; j < nl; j += BA_THREADS) {
                        if (fixed[j]) continue;
                        pts[3 * j] += lm[LM_STRIDE * j + 16]; pts[3 * j + 1] += lm[LM_STRIDE * j + 17]; pts[3 * j + 2] += lm[LM_STRIDE * j + 18];
                    }
                    for (int i = tid; i < np; i += BA_THREADS) pose_oplus(Rt + 12 * i, xs + 6 * i);
                    __syncthreads();
                }
                BA_T(6);
                double tempChi = compute_errors(c, ne, err, nullptr, red);
                BA_T(7);
                if (!ok) tempChi = 1.7976931348623157e308;
                rho = (currentChi - tempChi) / (scale + 1e-3);
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
                    for (int k = tid; k < 12 * np; k += BA_THREADS) Rt[k] = Rtb[k];   // pop(); err keeps the trial's values
                    for (int k = tid; k < 3 * nl; k += BA_THREADS) pts[k] = ptb[k];
                    __syncthreads();
                }
                qmax++;
            } while (rho < 0 && qmax < 10);
            lm_total++;
            if (qmax == 10 || rho == 0) terminated = true;
        }
        rounds++;
        // ---- inlier ratio on chi2 of the last evaluated errors (:216-227)
        int ci = 0, co = 0;
        for (int e = tid; e < ne; e += BA_THREADS) {
            const double c2 = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1];
            if (c2 > a.chi2_th) co++; else ci++;
        }
        inl = (int)(block_sum((double)ci, red) + 0.5);
        outl = (int)(block_sum((double)co, red) + 0.5);
        const double ratio = inl / (double)(inl + outl);
        if (ratio > 0.5) break;
        outer++;
    }
    // ---- outputs
    double *chi2 = a.chi2 + (size_t)w * a.MO;
    uint8_t *outlier = a.outlier + (size_t)w * a.MO;
    for (int e = tid; e < ne; e += BA_THREADS) {
        const double c2 = err[2 * e] * err[2 * e] + err[2 * e + 1] * err[2 * e + 1];
        chi2[e] = c2;
        outlier[e] = c2 > a.chi2_th;
    }
    for (int i = tid; i < np; i += BA_THREADS) {
        R_to_quat(Rt + 12 * i, poses + 7 * i);
        poses[7 * i + 4] = Rt[12 * i + 9]; poses[7 * i + 5] = Rt[12 * i + 10]; poses[7 * i + 6] = Rt[12 * i + 11];
    }
    if (tid == 0) { info[0] = rounds; info[1] = lm_total; info[2] = inl; info[3] = outl; }
}

// ================================================================================================
// host side
// ================================================================================================
static size_t ba_smem_bytes(int MP) {
    return sizeof(double) * (size_t)(12 * MP * 2 + 36 * MP + 6 * MP * 2 + 36 * MP * MP + 16 + 6 * MP + 36 * MP) + sizeof(int) * (size_t)(MP * (MP + 1) / 2 + 2);
}

static void free_ba(sb_ba *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_np, h->d_nl, h->d_ne, h->d_info, h->d_poses, h->d_points, h->d_uv, h->d_chi2, h->d_fixed,
                    h->d_outlier, h->d_op, h->d_ol, h->d_edge_of, h->d_next, h->d_lm, h->d_ptbak, h->d_err, h->d_hpl, h->d_ybd, h->d_pairs, h->d_plist};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->done) cudaEventDestroy(h->done);
    delete h;
}

extern "C" int sb_ba_create(sb_ba_t **out, int device, int max_windows, int max_poses, int max_points, int max_obs) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_windows >= 1 && max_windows <= 65535, "max_windows out of range [1, 65535]");
    SB_REQUIRE(max_poses >= 1 && max_poses <= BA_MAX_POSES, "max_poses out of range [1, 16]");
    SB_REQUIRE(max_points >= 1 && max_points <= (1 << 20), "max_points out of range");
    SB_REQUIRE(max_obs >= 1 && max_obs <= (1 << 24), "max_obs out of range");
    SB_TRY(sb_use_device(device));
    sb_ba *h = new sb_ba();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_windows = max_windows;
    h->max_poses = max_poses;
    h->max_points = max_points;
    h->max_obs = max_obs;
    const size_t W = max_windows, MP = max_poses, ML = max_points, MO = max_obs;
    cudaError_t e = cudaSuccess;
#define BA_ALLOC(ptr, bytes) \
    if (e == cudaSuccess) e = cudaMalloc((void **)&(ptr), (bytes))
    BA_ALLOC(h->d_np, W * 4);
    BA_ALLOC(h->d_nl, W * 4);
    BA_ALLOC(h->d_ne, W * 4);
    BA_ALLOC(h->d_info, W * 16);
    BA_ALLOC(h->d_poses, W * MP * 7 * 8);
    BA_ALLOC(h->d_points, W * ML * 3 * 8);
    BA_ALLOC(h->d_uv, W * MO * 2 * 8);
    BA_ALLOC(h->d_chi2, W * MO * 8);
    BA_ALLOC(h->d_fixed, W * ML);
    BA_ALLOC(h->d_outlier, W * MO);
    BA_ALLOC(h->d_op, W * MO * 4);
    BA_ALLOC(h->d_ol, W * MO * 4);
    BA_ALLOC(h->d_edge_of, W * ML * MP * 4);
    BA_ALLOC(h->d_next, W * MO * 4);
    BA_ALLOC(h->d_lm, W * ML * LM_STRIDE * 8);
    BA_ALLOC(h->d_ptbak, W * ML * 3 * 8);
    BA_ALLOC(h->d_err, W * MO * 4 * 8);
    BA_ALLOC(h->d_hpl, W * MO * 18 * 8);
    BA_ALLOC(h->d_ybd, W * MO * 10 * 8);
    BA_ALLOC(h->d_pairs, W * (MP * (MP + 1) / 2) * ML * sizeof(int4));
    BA_ALLOC(h->d_plist, W * MP * ML * sizeof(int2));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ba_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ba_smem_bytes(BA_MAX_POSES));
    if (e != cudaSuccess) {
        sb_set_error("sb_ba_create: %s", cudaGetErrorString(e));
        free_ba(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_ba_destroy(sb_ba_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_ba(h);
    }
    return SB_OK;
}

extern "C" int sb_ba_set_stream(sb_ba_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

static void quat7_to_ext(const double *p, double *R, double *t) {
    const double n = sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + p[3] * p[3]);
    const double x = p[0] / n, y = p[1] / n, z = p[2] / n, w = p[3] / n;
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
    t[0] = p[4]; t[1] = p[5]; t[2] = p[6];
}

extern "C" int sb_ba_solve_dev(sb_ba_t *h, int n_windows, const int32_t *d_n_poses, const int32_t *d_n_points,
                               const int32_t *d_n_obs, double *d_poses, double *d_points, const uint8_t *d_fixed,
                               const int32_t *d_obs_pose, const int32_t *d_obs_point, const double *d_uv, const double *K,
                               const double *cam_ext7, double huber_delta, double chi2_th, int outer_max,
                               int inner_iters, double *d_chi2, uint8_t *d_outlier, int32_t *d_info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(n_windows >= 1 && n_windows <= h->max_windows, "n_windows out of range [1, max_windows]");
    SB_REQUIRE(d_n_poses && d_n_points && d_n_obs && d_poses && d_points && d_fixed && d_obs_pose && d_obs_point && d_uv &&
                   d_chi2 && d_outlier && d_info && K && cam_ext7,
               "null pointer");
    SB_REQUIRE(huber_delta > 0 && outer_max >= 1 && inner_iters >= 1, "bad solver parameters");
    SB_TRY(sb_use_device(h->device));
    BaArgs a;
    a.np = d_n_poses; a.nl = d_n_points; a.ne = d_n_obs;
    a.poses = d_poses; a.points = d_points; a.fixed = d_fixed; a.op = d_obs_pose; a.ol = d_obs_point; a.uv = d_uv;
    a.chi2 = d_chi2; a.outlier = d_outlier; a.info = d_info;
    a.edge_of = h->d_edge_of; a.next_dup = h->d_next; a.lm = h->d_lm; a.ptbak = h->d_ptbak; a.err = h->d_err; a.hpl = h->d_hpl; a.ybd = h->d_ybd; a.pairs = h->d_pairs; a.plist = h->d_plist;
    a.MP = h->max_poses; a.ML = h->max_points; a.MO = h->max_obs;
    a.fx = K[0]; a.fy = K[1]; a.cx = K[2]; a.cy = K[3];
    quat7_to_ext(cam_ext7, a.extR, a.extT);
    a.delta = huber_delta; a.chi2_th = chi2_th; a.outer_max = outer_max; a.inner_iters = inner_iters;
    k_ba_solve<<<n_windows, BA_THREADS, ba_smem_bytes(h->max_poses), h->stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

static int ba_enqueue(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points, const int32_t *n_obs, double *poses,
                      double *points, const uint8_t *fixed, const int32_t *obs_pose, const int32_t *obs_point, const double *uv,
                      const double *K, const double *cam_ext7, double huber_delta, double chi2_th, int outer_max, int inner_iters,
                      double *chi2, uint8_t *outlier, int32_t *info);

// Asynchronous host-pointer form: the reference's Backend runs in its own thread beside the front end
// (src/backend.cpp:29-45), so the caller enqueues a batch of windows and collects it later.  Host arrays
// must stay valid (and should be pinned) until sb_ba_wait returns.
extern "C" int sb_ba_submit(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points,
                            const int32_t *n_obs, double *poses, double *points, const uint8_t *fixed,
                            const int32_t *obs_pose, const int32_t *obs_point, const double *uv, const double *K,
                            const double *cam_ext7, double huber_delta, double chi2_th, int outer_max, int inner_iters,
                            double *chi2, uint8_t *outlier, int32_t *info) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(!h->pending_info, "a batch is already in flight: call sb_ba_wait first");
    SB_REQUIRE(n_windows >= 1 && n_windows <= h->max_windows, "n_windows out of range [1, max_windows]");
    SB_REQUIRE(n_poses && n_points && n_obs && poses && points && fixed && obs_pose && obs_point && uv && chi2 && outlier && info,
               "null pointer");
    // everything sb_ba_solve_dev checks is checked BEFORE the first asynchronous copy touches the caller's buffers: an
    // argument error must not leave copies in flight behind a call that reported "nothing pending"
    SB_REQUIRE(K && cam_ext7, "null pointer");
    SB_REQUIRE(huber_delta > 0 && outer_max >= 1 && inner_iters >= 1, "bad solver parameters");
    SB_TRY(sb_use_device(h->device));
    const size_t W = n_windows, MP = h->max_poses, ML = h->max_points, MO = h->max_obs;
    cudaStream_t s = h->stream;
    const int rc = ba_enqueue(h, n_windows, n_poses, n_points, n_obs, poses, points, fixed, obs_pose, obs_point, uv, K, cam_ext7,
                              huber_delta, chi2_th, outer_max, inner_iters, chi2, outlier, info);
    if (rc != SB_OK) {   // a CUDA error in the middle of the sequence: drain the stream so that no copy still reads or
        cudaStreamSynchronize(s);   // writes the caller's buffers after the error return
        return rc;
    }
    h->pending_info = info;
    h->pending_windows = n_windows;
    return SB_OK;
}

static int ba_enqueue(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points, const int32_t *n_obs, double *poses,
                      double *points, const uint8_t *fixed, const int32_t *obs_pose, const int32_t *obs_point, const double *uv,
                      const double *K, const double *cam_ext7, double huber_delta, double chi2_th, int outer_max, int inner_iters,
                      double *chi2, uint8_t *outlier, int32_t *info) {
    const size_t W = n_windows, MP = h->max_poses, ML = h->max_points, MO = h->max_obs;
    cudaStream_t s = h->stream;
    SB_CUDA(cudaMemcpyAsync(h->d_np, n_poses, W * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_nl, n_points, W * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_ne, n_obs, W * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_poses, poses, W * MP * 56, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_points, points, W * ML * 24, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_fixed, fixed, W * ML, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_op, obs_pose, W * MO * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_ol, obs_point, W * MO * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_uv, uv, W * MO * 16, cudaMemcpyHostToDevice, s));
    SB_TRY(sb_ba_solve_dev(h, n_windows, h->d_np, h->d_nl, h->d_ne, h->d_poses, h->d_points, h->d_fixed, h->d_op, h->d_ol,
                           h->d_uv, K, cam_ext7, huber_delta, chi2_th, outer_max, inner_iters, h->d_chi2, h->d_outlier,
                           h->d_info));
    SB_CUDA(cudaMemcpyAsync(poses, h->d_poses, W * MP * 56, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(points, h->d_points, W * ML * 24, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(chi2, h->d_chi2, W * MO * 8, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(outlier, h->d_outlier, W * MO, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(info, h->d_info, W * 16, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaEventRecord(h->done, s));
    return SB_OK;
}

extern "C" int sb_ba_wait(sb_ba_t *h) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(h->pending_info, "no batch in flight");
    SB_TRY(sb_use_device(h->device));
    const int32_t *info = h->pending_info;
    const size_t W = (size_t)h->pending_windows;
    h->pending_info = nullptr;
    SB_CUDA(cudaEventSynchronize(h->done));
    for (size_t w = 0; w < W; w++)
        if (info[4 * w] < 0) {
            sb_set_error("window %zu: %s", w, "counts or indices out of range");
            return SB_ERR_INVALID;
        }
    return SB_OK;
}

extern "C" int sb_ba_solve(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points,
                           const int32_t *n_obs, double *poses, double *points, const uint8_t *fixed,
                           const int32_t *obs_pose, const int32_t *obs_point, const double *uv, const double *K,
                           const double *cam_ext7, double huber_delta, double chi2_th, int outer_max, int inner_iters,
                           double *chi2, uint8_t *outlier, int32_t *info) {
    SB_NVTX_FN();
    SB_TRY(sb_ba_submit(h, n_windows, n_poses, n_points, n_obs, poses, points, fixed, obs_pose, obs_point, uv, K, cam_ext7,
                        huber_delta, chi2_th, outer_max, inner_iters, chi2, outlier, info));
    return sb_ba_wait(h);
}

#ifdef BA_PROFILE
extern "C" int sb_ba_debug_profile(long long *out) {
    SB_NVTX_FN();
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_ba_prof, sizeof(long long) * 16);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_ba_prof, z, sizeof(z));
    return 0;
}
#endif
