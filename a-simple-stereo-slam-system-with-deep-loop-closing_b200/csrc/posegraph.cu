// posegraph.cu — loop-closure pose-graph optimisation on B200 (sm_100a), double precision.
//
// Replaces the g2o solve inside LoopClosing::PoseGraphOptimization (reference
// src/loopclosing.cpp:537-646): Levenberg-Marquardt (20 iterations) over the keyframe poses with
// EdgePoseGraph edges (include/myslam/g2o_types.h:157-190: error log(Z^-1 T0 T1^-1), information I6,
// NUMERIC Jacobians — central differences, step 1e-9, through VertexPose::oplusImpl — because the
// reference leaves linearizeOplus commented out), BlockSolver<6,6> without marginalisation.
//
// The reference's graph is a chain (keyframe -> previous keyframe) plus a handful of loop edges
// (17 on KITTI-00), so the LM system is block tridiagonal plus a low-rank term:
//        H + lambda I = T + J_loop^T J_loop,
// T from the chain edges and the edges with one fixed end, J_loop (6R x 6n) from the R long-range
// edges.  It is solved exactly (a direct method, like the reference's sparse Cholesky) by a block
// Thomas factorisation of T, a multi-right-hand-side substitution for T^-1 [b, J^T] and the
// Woodbury identity with a dense Cholesky of the 6R x 6R capacitance matrix — O(n) work instead of
// a general sparse factorisation.  One CTA runs the whole optimisation of one graph in one launch.
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "se3.cuh"

#define PG_THREADS 256
#define PG_MAX_LOOPS 64
#define PG_CHUNK 16
#define PG_PF 4     // chain steps whose right-hand sides are fetched ahead (PG_CHUNK is a multiple)

enum { PG_NONE = 0, PG_DIAG = 1, PG_CHAIN = 2, PG_LOOP = 3 };

#ifdef PG_PROFILE  // per-phase clock64 totals (tools/pg_profile.py); the product build has no trace of it
__device__ long long g_pg_prof[16];
#define PG_T(i) do { __syncthreads(); if (threadIdx.x == 0) { long long t_ = clock64(); g_pg_prof[i] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define PG_T(i) ((void)0)
#endif

struct sb_posegraph {
    int device, max_vertices, max_edges;
    cudaStream_t stream, own_stream;
    double *d_poses, *d_meas;
    uint8_t *d_fixed;
    int32_t *d_v0, *d_v1, *d_info;
    // workspace
    double *d_work;
    int32_t *d_iwork;
    size_t work_doubles, iwork_ints;
};

struct PgArgs {
    int n, m, iters;
    double *poses;
    const uint8_t *fixed;
    const int32_t *v0, *v1;
    const double *meas;
    int32_t *info;  // [4]: LM iterations, trials, free vertices, long-range edges  (info[0] < 0: error)
    double *stats;  // [2]: chi2 at start, chi2 at the end
    double *Rt, *Rtb, *Zinv, *err, *Ji, *Jj, *A, *B, *bvec, *hdiag, *Sinv, *G, *Q, *M, *yv, *x;
    int32_t *fidx, *vof, *ecls, *eloop, *inc_start, *inc_edge, *loop_edge;
};

static __device__ double pg_block_sum(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0;
    for (int k = 0; k < PG_THREADS / 32; k++) s += red[k];
    return s;
}
static __device__ double pg_block_max(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = red[0];
    for (int k = 1; k < PG_THREADS / 32; k++) s = fmax(s, red[k]);
    return s;
}

// EdgePoseGraph::computeError with explicit vertex poses
static __device__ __forceinline__ void pg_edge_error(const double *Zinv, const double *T0, const double *T1, double *e) {
    double T1i[12], P[12];
    se3_inv(T1, T1i);
    se3_mul(T0, T1i, P);
    se3_mul(Zinv, P, P);
    se3_log(P, e);
}

static __device__ double pg_errors(const PgArgs &a, double *red) {
    double chi = 0;
    for (int e = threadIdx.x; e < a.m; e += PG_THREADS) {
        double r[6];
        pg_edge_error(a.Zinv + 12 * e, a.Rt + 12 * a.v0[e], a.Rt + 12 * a.v1[e], r);
#pragma unroll
        for (int k = 0; k < 6; k++) { a.err[6 * e + k] = r[k]; chi += r[k] * r[k]; }
    }
    return pg_block_sum(chi, red);
}

__global__ void __launch_bounds__(PG_THREADS) k_posegraph(const __grid_constant__ PgArgs a) {
    __shared__ double red[16];
    __shared__ double w_B[36], w_G[36], w_M[36], w_I[36];  // 6x6 work blocks of the factorisation warp
    __shared__ __align__(16) double w_stage[2 * PG_CHUNK * 36];           // G / (S^-1, B) of a chunk of chain steps
    __shared__ int s_nf, s_R, s_bad;
    __shared__ double s_diag[6 * PG_MAX_LOOPS];  // diagonal of the capacitance matrix' Cholesky factor
    const int tid = threadIdx.x;
    const int n = a.n, m = a.m;

    // ---- set-up: free-vertex numbering, edge classes, vertex -> edge incidence (serial, once)
    if (tid == 0) {
        int nf = 0, bad = 0;
        for (int v = 0; v < n; v++) {
            if (a.fixed[v]) a.fidx[v] = -1;
            else { a.fidx[v] = nf; a.vof[nf] = v; nf++; }
        }
        for (int v = 0; v <= n; v++) a.inc_start[v] = 0;
        int R = 0;
        for (int e = 0; e < m; e++) {
            const int u0 = a.v0[e], u1 = a.v1[e];
            if (u0 < 0 || u0 >= n || u1 < 0 || u1 >= n || u0 == u1) { bad = 1; break; }
            const int p0 = a.fidx[u0], p1 = a.fidx[u1];
            int cls;
            if (p0 < 0 && p1 < 0) cls = PG_NONE;
            else if (p0 < 0 || p1 < 0) cls = PG_DIAG;
            else if (p0 - p1 == 1 || p1 - p0 == 1) cls = PG_CHAIN;
            else cls = PG_LOOP;
            a.ecls[e] = cls;
            a.eloop[e] = -1;
            if (cls == PG_LOOP) {
                if (R >= PG_MAX_LOOPS) { bad = 2; break; }
                a.eloop[e] = R;
                a.loop_edge[R] = e;
                R++;
            }
            a.inc_start[u0 + 1]++;
            a.inc_start[u1 + 1]++;
        }
        if (!bad)
            for (int v = 0; v < n; v++) a.inc_start[v + 1] += a.inc_start[v];
        s_nf = nf; s_R = R; s_bad = bad;
    }
    __syncthreads();
    if (s_bad) {
        if (tid == 0) { a.info[0] = -1; a.info[1] = -s_bad; a.info[2] = a.info[3] = 0; }
        return;
    }
    const int nf = s_nf, R = s_R, NC = 1 + 6 * R;
    // incidence lists: vertex v's edges in ascending edge index (serial fill keeps the order deterministic)
    if (tid == 0) {
        int *cursor = a.eloop + m;  // scratch of n ints behind eloop
        for (int v = 0; v < n; v++) cursor[v] = a.inc_start[v];
        for (int e = 0; e < m; e++) {
            a.inc_edge[cursor[a.v0[e]]++] = e;
            a.inc_edge[cursor[a.v1[e]]++] = e;
        }
    }
    for (int v = tid; v < n; v += PG_THREADS) {
        quat_to_R(a.poses + 7 * v, a.Rt + 12 * v);
        a.Rt[12 * v + 9] = a.poses[7 * v + 4]; a.Rt[12 * v + 10] = a.poses[7 * v + 5]; a.Rt[12 * v + 11] = a.poses[7 * v + 6];
    }
    for (int e = tid; e < m; e += PG_THREADS) {
        double Z[12];
        quat_to_R(a.meas + 7 * e, Z);
        Z[9] = a.meas[7 * e + 4]; Z[10] = a.meas[7 * e + 5]; Z[11] = a.meas[7 * e + 6];
        se3_inv(Z, a.Zinv + 12 * e);
    }
    __syncthreads();

    double lambda = 0, ni = 2, chi_start = 0, chi_end = 0;
    int lm_iters = 0, trials = 0;
    bool terminated = false;
#ifdef PG_PROFILE
    long long t_prev = clock64();
#endif
    for (int it = 0; it < a.iters && !terminated; it++) {
        PG_T(7);
        double currentChi = pg_errors(a, red);
        PG_T(0);
        if (it == 0) chi_start = currentChi;
        chi_end = currentChi;
        if (nf == 0) break;
        // ---- numeric Jacobians (g2o BaseBinaryEdge::linearizeOplus): one work item per (edge, vertex, column)
        for (int item = tid; item < m * 12; item += PG_THREADS) {
            const int e = item / 12, side = (item % 12) / 6, d = item % 6;
            const int u0 = a.v0[e], u1 = a.v1[e];
            if (a.fixed[side ? u1 : u0]) continue;
            const double delta = 1e-9, scalar = 1.0 / (2 * delta);
            double ep[6], em[6], T[12], step[6] = {0, 0, 0, 0, 0, 0};
            const double *base = a.Rt + 12 * (side ? u1 : u0);
#pragma unroll
            for (int k = 0; k < 12; k++) T[k] = base[k];
            step[d] = delta;
            pose_oplus(T, step);
            pg_edge_error(a.Zinv + 12 * e, side ? a.Rt + 12 * u0 : T, side ? T : a.Rt + 12 * u1, ep);
#pragma unroll
            for (int k = 0; k < 12; k++) T[k] = base[k];
            step[d] = -delta;
            pose_oplus(T, step);
            pg_edge_error(a.Zinv + 12 * e, side ? a.Rt + 12 * u0 : T, side ? T : a.Rt + 12 * u1, em);
            double *J = (side ? a.Jj : a.Ji) + 36 * e;
#pragma unroll
            for (int r = 0; r < 6; r++) J[6 * r + d] = scalar * (ep[r] - em[r]);
        }
        __syncthreads();
        PG_T(1);
        // ---- assembly, one thread per free vertex (gather over its incident edges, ascending edge index)
        double mx = 0;
        for (int p = tid; p < nf; p += PG_THREADS) {
            const int v = a.vof[p];
            double Ap[36], Bp[36], g[6], hd[6];
            for (int k = 0; k < 36; k++) { Ap[k] = 0; Bp[k] = 0; }
            for (int k = 0; k < 6; k++) { g[k] = 0; hd[k] = 0; }
            for (int q = a.inc_start[v]; q < a.inc_start[v + 1]; q++) {
                const int e = a.inc_edge[q];
                const int cls = a.ecls[e];
                const bool first = a.v0[e] == v;
                const double *Jv = (first ? a.Ji : a.Jj) + 36 * e;
                const double *er = a.err + 6 * e;
                for (int i = 0; i < 6; i++) {
                    double s = 0;
                    for (int r = 0; r < 6; r++) s += Jv[6 * r + i] * er[r];
                    g[i] -= s;
                    double dd = 0;
                    for (int r = 0; r < 6; r++) dd += Jv[6 * r + i] * Jv[6 * r + i];
                    hd[i] += dd;
                }
                if (cls == PG_DIAG || cls == PG_CHAIN) {
                    for (int i = 0; i < 6; i++)
                        for (int j = 0; j < 6; j++) {
                            double s = 0;
                            for (int r = 0; r < 6; r++) s += Jv[6 * r + i] * Jv[6 * r + j];
                            Ap[6 * i + j] += s;
                        }
                }
                if (cls == PG_CHAIN) {
                    const int other = first ? a.v1[e] : a.v0[e];
                    if (a.fidx[other] == p - 1) {  // block (p, p-1) = J_p^T J_{p-1}
                        const double *Jo = (first ? a.Jj : a.Ji) + 36 * e;
                        for (int i = 0; i < 6; i++)
                            for (int j = 0; j < 6; j++) {
                                double s = 0;
                                for (int r = 0; r < 6; r++) s += Jv[6 * r + i] * Jo[6 * r + j];
                                Bp[6 * i + j] += s;
                            }
                    }
                }
            }
            for (int k = 0; k < 36; k++) { a.A[36 * p + k] = Ap[k]; a.B[36 * p + k] = Bp[k]; }
            for (int k = 0; k < 6; k++) { a.bvec[6 * p + k] = g[k]; mx = fmax(mx, fabs(hd[k])); }
        }
        if (it == 0) {  // computeLambdaInit
            lambda = 1e-5 * pg_block_max(mx, red);
            ni = 2;
        }
        __syncthreads();

        double rho = 0;
        int qmax = 0;
        PG_T(2);
        do {
            PG_T(7);
            for (int k = tid; k < 12 * n; k += PG_THREADS) a.Rtb[k] = a.Rt[k];  // push()
            // ---- block Thomas factorisation of T = tridiag(B, A + lambda I, B^T): S_p = A_p - G_p B_p^T with
            //      G_p = B_p S_{p-1}^-1.  The recurrence is serial in p; inside a step one warp works on the 6x6
            //      blocks (18 lanes x 2 elements: rows r and r + 3 of column c), products and the Gauss-Jordan
            //      inverse go through shared memory with warp-level synchronisation only.
            if (tid < 32) {
                const int l = tid, r = l / 6, cc = l % 6;  // lanes 0..17 own (r, cc) and (r + 3, cc)
                const bool act = l < 18;
                int bad = 0;
                // the blocks of step p + 1 are loaded during step p: the recurrence is a chain of dependent 6 x 6 operations,
                // and a global load at the head of every step would sit on its critical path
                double nA0 = 0, nA1 = 0, nB0 = 0, nB1 = 0;
                if (act && nf > 0) { nA0 = a.A[6 * r + cc]; nA1 = a.A[6 * (r + 3) + cc]; }
                for (int p = 0; p < nf; p++) {
                    double e0 = 0, e1 = 0;
                    if (act) {
                        e0 = nA0 + (r == cc ? lambda : 0.0);
                        e1 = nA1 + (r + 3 == cc ? lambda : 0.0);
                        if (p > 0) { w_B[6 * r + cc] = nB0; w_B[6 * (r + 3) + cc] = nB1; }
                        if (p + 1 < nf) {
                            nA0 = a.A[36 * (p + 1) + 6 * r + cc]; nA1 = a.A[36 * (p + 1) + 6 * (r + 3) + cc];
                            nB0 = a.B[36 * (p + 1) + 6 * r + cc]; nB1 = a.B[36 * (p + 1) + 6 * (r + 3) + cc];
                        }
                    }
                    __syncwarp();
                    if (p > 0) {
                        if (act) {  // G = B S_{p-1}^-1
                            double g0 = 0, g1 = 0;
#pragma unroll
                            for (int k = 0; k < 6; k++) { g0 += w_B[6 * r + k] * w_I[6 * k + cc]; g1 += w_B[6 * (r + 3) + k] * w_I[6 * k + cc]; }
                            w_G[6 * r + cc] = g0; w_G[6 * (r + 3) + cc] = g1;
                            a.G[36 * p + 6 * r + cc] = g0; a.G[36 * p + 6 * (r + 3) + cc] = g1;
                        }
                        __syncwarp();
                        if (act) {  // S = A - G B^T
#pragma unroll
                            for (int k = 0; k < 6; k++) { e0 -= w_G[6 * r + k] * w_B[6 * cc + k]; e1 -= w_G[6 * (r + 3) + k] * w_B[6 * cc + k]; }
                        }
                    }
                    if (act) { w_M[6 * r + cc] = e0; w_M[6 * (r + 3) + cc] = e1; }
                    __syncwarp();
                    if (act) {  // symmetrise against round-off, start Gauss-Jordan on [M | I]
                        e0 = 0.5 * (w_M[6 * r + cc] + w_M[6 * cc + r]);
                        e1 = 0.5 * (w_M[6 * (r + 3) + cc] + w_M[6 * cc + r + 3]);
                    }
                    double i0 = r == cc ? 1.0 : 0.0, i1 = r + 3 == cc ? 1.0 : 0.0;
                    // Gauss-Jordan on [M | I] in registers: the pivot row and the pivot column travel by warp shuffles
                    // (element (r, cc) lives in lane 6 r + cc, rows 3..5 in the second register of lanes 0..17)
                    const int col_src = act ? 6 * r : 0;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const double rowm = k < 3 ? e0 : e1, rowi = k < 3 ? i0 : i1;
                        const double piv = __shfl_sync(0xffffffffu, rowm, 6 * (k % 3) + k);
                        double mk = __shfl_sync(0xffffffffu, rowm, 6 * (k % 3) + cc);
                        double ik = __shfl_sync(0xffffffffu, rowi, 6 * (k % 3) + cc);
                        const double f0 = __shfl_sync(0xffffffffu, e0, col_src + k), f1 = __shfl_sync(0xffffffffu, e1, col_src + k);
                        if (!(piv > 0)) bad = 1;  // positive definite blocks have positive pivots without pivoting
                        const double ip = 1.0 / piv;
                        mk *= ip;
                        ik *= ip;
                        e0 = r == k ? mk : e0 - f0 * mk;      i0 = r == k ? ik : i0 - f0 * ik;
                        e1 = r + 3 == k ? mk : e1 - f1 * mk;  i1 = r + 3 == k ? ik : i1 - f1 * ik;
                    }
                    __syncwarp();
                    if (act) { w_I[6 * r + cc] = i0; w_I[6 * (r + 3) + cc] = i1; }
                    __syncwarp();
                    if (act) { a.Sinv[36 * p + 6 * r + cc] = i0; a.Sinv[36 * p + 6 * (r + 3) + cc] = i1; }
                    if (__any_sync(0xffffffffu, bad)) break;
                }
                if (tid == 0) s_bad = bad;
            }
            PG_T(3);
            // ---- right-hand sides Q[p][c][6]: column 0 = b, column 1 + 6r + k = row k of J_loop,r (as a column of J^T)
            for (int idx = tid; idx < nf * NC * 6; idx += PG_THREADS) a.Q[idx] = 0;
            __syncthreads();
            for (int idx = tid; idx < nf * 6; idx += PG_THREADS) a.Q[(size_t)(idx / 6) * NC * 6 + idx % 6] = a.bvec[idx];
            for (int idx = tid; idx < R * 6 * 2; idx += PG_THREADS) {
                const int r = idx / 12, k = (idx % 12) / 2, side = idx % 2;
                const int e = a.loop_edge[r];
                const int p = a.fidx[side ? a.v1[e] : a.v0[e]];
                const double *J = (side ? a.Jj : a.Ji) + 36 * e;
                double *q = a.Q + ((size_t)p * NC + 1 + 6 * r + k) * 6;
#pragma unroll
                for (int i = 0; i < 6; i++) q[i] = J[6 * k + i];
            }
            __syncthreads();
            int ok = !s_bad;
            __syncthreads();
            PG_T(4);
            if (ok) {
                // ---- T^-1 [b, J^T]: one thread per column, sequential over the chain; the 6x6 blocks every column
                //      needs (G_p forward, S_p^-1 and B_{p+1} backward) are staged through shared memory in chunks.
                {
                    const int ncol_iters = (NC + PG_THREADS - 1) / PG_THREADS;
                    for (int ci = 0; ci < ncol_iters; ci++) {
                        const int c = ci * PG_THREADS + tid;
                        const bool on = c < NC;
                        // The recurrences only carry y (forward) and x (backward); the right-hand sides live in L2 (3.6 MB for
                        // KITTI-00), ~700 cycles away.  They are therefore fetched a block of PG_PF steps ahead of their use —
                        // without this every step waits for its own loads behind the previous step's stores.
                        double y[6] = {0, 0, 0, 0, 0, 0};
                        double cur[PG_PF][6], nxt[PG_PF][6];
#pragma unroll
                        for (int s4 = 0; s4 < PG_PF; s4++)
#pragma unroll
                            for (int i = 0; i < 6; i++) cur[s4][i] = (on && s4 < nf) ? a.Q[((size_t)s4 * NC + c) * 6 + i] : 0.0;
                        for (int p0 = 0; p0 < nf; p0 += PG_CHUNK) {
                            const int pn = min(PG_CHUNK, nf - p0);
                            __syncthreads();
                            for (int k = tid; k < pn * 36; k += PG_THREADS) w_stage[k] = a.G[36 * p0 + k];
                            __syncthreads();
                            if (on)
                                for (int pb = 0; pb < pn; pb += PG_PF) {
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++) {  // the block after this one
                                        const int pnx = p0 + pb + PG_PF + s4;
#pragma unroll
                                        for (int i = 0; i < 6; i++) nxt[s4][i] = pnx < nf ? a.Q[((size_t)pnx * NC + c) * 6 + i] : 0.0;
                                    }
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++) {
                                        const int pp = pb + s4, p = p0 + pp;
                                        if (pp < pn) {
                                            double v[6];
#pragma unroll
                                            for (int i = 0; i < 6; i++) v[i] = cur[s4][i];
                                            if (p > 0) {
                                                const double *Gp = w_stage + 36 * pp;
#pragma unroll
                                                for (int i = 0; i < 6; i++) {
                                                    double sacc = v[i];
#pragma unroll
                                                    for (int k = 0; k < 6; k++) sacc -= Gp[6 * i + k] * y[k];
                                                    v[i] = sacc;
                                                }
                                                double *q = a.Q + ((size_t)p * NC + c) * 6;
#pragma unroll
                                                for (int i = 0; i < 6; i++) q[i] = v[i];
                                            }
#pragma unroll
                                            for (int i = 0; i < 6; i++) y[i] = v[i];
                                        }
                                    }
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++)
#pragma unroll
                                        for (int i = 0; i < 6; i++) cur[s4][i] = nxt[s4][i];
                                }
                        }
                        double xn[6] = {0, 0, 0, 0, 0, 0};
                        // backward: block s4 = 0 is the highest step; the forward pass' stores are visible to this thread
#pragma unroll
                        for (int s4 = 0; s4 < PG_PF; s4++)
#pragma unroll
                            for (int i = 0; i < 6; i++) cur[s4][i] = (on && nf - 1 - s4 >= 0) ? a.Q[((size_t)(nf - 1 - s4) * NC + c) * 6 + i] : 0.0;
                        for (int pend = nf; pend > 0; pend -= PG_CHUNK) {
                            const int p0 = max(0, pend - PG_CHUNK), pn = pend - p0;
                            __syncthreads();
                            for (int k = tid; k < pn * 36; k += PG_THREADS) {
                                w_stage[k] = a.Sinv[36 * p0 + k];
                                w_stage[PG_CHUNK * 36 + k] = p0 + k / 36 + 1 < nf ? a.B[36 * (p0 + 1) + k] : 0.0;  // B_{p+1}
                            }
                            __syncthreads();
                            if (on)
                                for (int pb = pn - 1; pb >= 0; pb -= PG_PF) {
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++) {  // the block below this one
                                        const int pnx = p0 + pb - PG_PF - s4;
#pragma unroll
                                        for (int i = 0; i < 6; i++) nxt[s4][i] = pnx >= 0 ? a.Q[((size_t)pnx * NC + c) * 6 + i] : 0.0;
                                    }
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++) {
                                        const int pp = pb - s4, p = p0 + pp;
                                        if (pp >= 0) {
                                            double r6[6];
#pragma unroll
                                            for (int i = 0; i < 6; i++) r6[i] = cur[s4][i];
                                            if (p + 1 < nf) {
                                                const double *Bn = w_stage + PG_CHUNK * 36 + 36 * pp;  // (B_{p+1})^T x_{p+1}
#pragma unroll
                                                for (int i = 0; i < 6; i++) {
                                                    double sacc = 0;
#pragma unroll
                                                    for (int k = 0; k < 6; k++) sacc += Bn[6 * k + i] * xn[k];
                                                    r6[i] -= sacc;
                                                }
                                            }
                                            const double *Si = w_stage + 36 * pp;
#pragma unroll
                                            for (int i = 0; i < 6; i++) {
                                                double sacc = 0;
#pragma unroll
                                                for (int k = 0; k < 6; k++) sacc += Si[6 * i + k] * r6[k];
                                                xn[i] = sacc;
                                            }
                                            double *q = a.Q + ((size_t)p * NC + c) * 6;
#pragma unroll
                                            for (int i = 0; i < 6; i++) q[i] = xn[i];
                                        }
                                    }
#pragma unroll
                                    for (int s4 = 0; s4 < PG_PF; s4++)
#pragma unroll
                                        for (int i = 0; i < 6; i++) cur[s4][i] = nxt[s4][i];
                                }
                        }
                    }
                }
                __syncthreads();
                PG_T(5);
                const int n6r = 6 * R;
                if (R > 0) {
                    // ---- capacitance matrix M = I + J Z and right-hand side v = J x0
                    for (int idx = tid; idx < n6r * (n6r + 1); idx += PG_THREADS) {
                        const int row = idx / (n6r + 1), col = idx % (n6r + 1);  // col n6r = the vector v
                        const int r = row / 6, k = row % 6;
                        const int e = a.loop_edge[r];
                        const int pa = a.fidx[a.v0[e]], pb = a.fidx[a.v1[e]];
                        const int c = col == n6r ? 0 : 1 + col;
                        const double *za = a.Q + ((size_t)pa * NC + c) * 6, *zb = a.Q + ((size_t)pb * NC + c) * 6;
                        const double *Ja = a.Ji + 36 * e + 6 * k, *Jb = a.Jj + 36 * e + 6 * k;
                        double s = 0;
#pragma unroll
                        for (int i = 0; i < 6; i++) s += Ja[i] * za[i] + Jb[i] * zb[i];
                        if (col == n6r) a.yv[row] = s;
                        else a.M[(size_t)row * n6r + col] = s + (row == col ? 1.0 : 0.0);
                    }
                    __syncthreads();
                    // ---- dense Cholesky of M, right-looking with one barrier per column: every thread reads the pivot and
                    //      forms the scaled column entries it needs itself; column j of L goes to ROW j of the upper triangle
                    //      (M[j][i], i > j), the trailing update works on the lower triangle, the diagonal of L is kept apart.
                    {
                        const int r0 = tid >> 3, c0 = tid & 7;
                        for (int j = 0; j < n6r; j++) {
                            const double d = a.M[(size_t)j * n6r + j];
                            if (!(d > 0)) { if (tid == 0) s_bad = 1; break; }  // uniform: every thread reads the same entry
                            const double dj = sqrt(d), inv = 1.0 / dj;
                            if (tid == 0) s_diag[j] = dj;
                            for (int i = j + 1 + r0; i < n6r; i += PG_THREADS / 8) {
                                const double lij = a.M[(size_t)i * n6r + j] * inv;
                                if (c0 == 0) a.M[(size_t)j * n6r + i] = lij;
                                for (int k = j + 1 + c0; k <= i; k += 8) a.M[(size_t)i * n6r + k] -= lij * (a.M[(size_t)k * n6r + j] * inv);
                            }
                            __syncthreads();
                        }
                    }
                    __syncthreads();
                    ok = !s_bad;
                    __syncthreads();
                    if (ok && tid < 32) {  // M y = v by one warp: L(i, k) = M[k][i]
                        const int lane = tid;
                        for (int i = 0; i < n6r; i++) {
                            double v = 0;
                            for (int k = lane; k < i; k += 32) v += a.M[(size_t)k * n6r + i] * a.yv[k];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if (lane == 0) a.yv[i] = (a.yv[i] - v) / s_diag[i];
                            __syncwarp();
                        }
                        for (int i = n6r - 1; i >= 0; i--) {
                            double v = 0;
                            for (int k = i + 1 + lane; k < n6r; k += 32) v += a.M[(size_t)i * n6r + k] * a.yv[k];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if (lane == 0) a.yv[i] = (a.yv[i] - v) / s_diag[i];
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                }
                if (ok) {
                    // ---- x = x0 - Z y
                    for (int idx = tid; idx < nf * 6; idx += PG_THREADS) {
                        const int p = idx / 6, i = idx % 6;
                        const double *q = a.Q + (size_t)p * NC * 6;
                        double s = q[i];
                        for (int c = 0; c < n6r; c++) s -= q[(1 + c) * 6 + i] * a.yv[c];
                        a.x[idx] = s;
                    }
                    __syncthreads();
                }
            }
            __syncthreads();
            PG_T(6);
            if (tid == 0) s_bad = 0;
            double scale = 0;
            if (ok) {
                for (int idx = tid; idx < nf * 6; idx += PG_THREADS) scale += a.x[idx] * (lambda * a.x[idx] + a.bvec[idx]);
                scale = pg_block_sum(scale, red);
                for (int p = tid; p < nf; p += PG_THREADS) pose_oplus(a.Rt + 12 * a.vof[p], a.x + 6 * p);
                __syncthreads();
            }
            double tempChi = pg_errors(a, red);
            if (!ok) tempChi = 1.7976931348623157e308;
            rho = (currentChi - tempChi) / (scale + 1e-3);
            trials++;
            if (rho > 0 && isfinite(tempChi)) {
                double alpha = 1. - pow(2 * rho - 1, 3);
                alpha = fmin(alpha, 2. / 3.);
                lambda *= fmax(1. / 3., alpha);
                ni = 2;
                currentChi = tempChi;
                chi_end = tempChi;
            } else {
                lambda *= ni;
                ni *= 2;
                __syncthreads();
                for (int k = tid; k < 12 * n; k += PG_THREADS) a.Rt[k] = a.Rtb[k];  // pop()
                __syncthreads();
            }
            qmax++;
        } while (rho < 0 && qmax < 10);
        lm_iters++;
        if (qmax == 10 || rho == 0) terminated = true;
    }
    __syncthreads();
    for (int v = tid; v < n; v += PG_THREADS) {
        if (a.fixed[v]) continue;  // fixed vertices keep their input bits
        R_to_quat(a.Rt + 12 * v, a.poses + 7 * v);
        a.poses[7 * v + 4] = a.Rt[12 * v + 9]; a.poses[7 * v + 5] = a.Rt[12 * v + 10]; a.poses[7 * v + 6] = a.Rt[12 * v + 11];
    }
    if (tid == 0) {
        a.info[0] = lm_iters; a.info[1] = trials; a.info[2] = nf; a.info[3] = R;
        a.stats[0] = chi_start; a.stats[1] = chi_end;
    }
}

// ================================================================================================
// host side
// ================================================================================================
static void free_pg(sb_posegraph *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    void *ptrs[] = {h->d_poses, h->d_meas, h->d_fixed, h->d_v0, h->d_v1, h->d_info, h->d_work, h->d_iwork};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int sb_posegraph_create(sb_posegraph_t **out, int device, int max_vertices, int max_edges) {
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_vertices >= 2 && max_vertices <= (1 << 20), "max_vertices out of range");
    SB_REQUIRE(max_edges >= 1 && max_edges <= (1 << 22), "max_edges out of range");
    SB_TRY(sb_use_device(device));
    sb_posegraph *h = new sb_posegraph();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_vertices = max_vertices;
    h->max_edges = max_edges;
    const size_t n = max_vertices, m = max_edges, NC = 1 + 6 * PG_MAX_LOOPS, R6 = 6 * PG_MAX_LOOPS;
    h->work_doubles = 24 * n + 12 * m + 6 * m + 72 * m + 36 * n * 4 + 6 * n * 3 + n * NC * 6 + R6 * R6 + R6 + 8;
    h->iwork_ints = 2 * n + 3 * m + n + (n + 1) + 2 * m + PG_MAX_LOOPS + 16;
    cudaError_t e = cudaMalloc((void **)&h->d_poses, n * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_meas, m * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_fixed, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_v0, m * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_v1, m * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_info, 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_work, h->work_doubles * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_iwork, h->iwork_ints * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        sb_set_error("sb_posegraph_create: %s", cudaGetErrorString(e));
        free_pg(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_posegraph_destroy(sb_posegraph_t *h) {
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_pg(h);
    }
    return SB_OK;
}

extern "C" int sb_posegraph_set_stream(sb_posegraph_t *h, void *stream) {
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_posegraph_solve_dev(sb_posegraph_t *h, int n_vertices, double *d_poses, const uint8_t *d_fixed,
                                      int n_edges, const int32_t *d_v0, const int32_t *d_v1, const double *d_meas,
                                      int iters, int32_t *d_info, double *d_stats) {
    sb_clear_error();
    SB_REQUIRE(h && d_poses && d_fixed && d_v0 && d_v1 && d_meas && d_info && d_stats, "null pointer");
    SB_REQUIRE(n_vertices >= 1 && n_vertices <= h->max_vertices, "n_vertices out of range [1, max_vertices]");
    SB_REQUIRE(n_edges >= 0 && n_edges <= h->max_edges, "n_edges out of range [0, max_edges]");
    SB_REQUIRE(iters >= 1, "iters must be positive");
    SB_TRY(sb_use_device(h->device));
    const size_t n = h->max_vertices, m = h->max_edges, NC = 1 + 6 * PG_MAX_LOOPS, R6 = 6 * PG_MAX_LOOPS;
    PgArgs a;
    a.n = n_vertices; a.m = n_edges; a.iters = iters;
    a.poses = d_poses; a.fixed = d_fixed; a.v0 = d_v0; a.v1 = d_v1; a.meas = d_meas; a.info = d_info; a.stats = d_stats;
    double *w = h->d_work;
    a.Rt = w; w += 12 * n;
    a.Rtb = w; w += 12 * n;
    a.Zinv = w; w += 12 * m;
    a.err = w; w += 6 * m;
    a.Ji = w; w += 36 * m;
    a.Jj = w; w += 36 * m;
    a.A = w; w += 36 * n;
    a.B = w; w += 36 * n;
    a.Sinv = w; w += 36 * n;
    a.G = w; w += 36 * n;
    a.bvec = w; w += 6 * n;
    a.hdiag = w; w += 6 * n;
    a.x = w; w += 6 * n;
    a.Q = w; w += n * NC * 6;
    a.M = w; w += R6 * R6;
    a.yv = w; w += R6;
    int32_t *iw = h->d_iwork;
    a.fidx = iw; iw += n;
    a.vof = iw; iw += n;
    a.ecls = iw; iw += m;
    a.eloop = iw; iw += m + n;  // + n ints of cursor scratch
    a.inc_start = iw; iw += n + 1;
    a.inc_edge = iw; iw += 2 * m;
    a.loop_edge = iw; iw += PG_MAX_LOOPS;
    k_posegraph<<<1, PG_THREADS, 0, h->stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_posegraph_solve(sb_posegraph_t *h, int n_vertices, double *poses, const uint8_t *fixed, int n_edges,
                                  const int32_t *v0, const int32_t *v1, const double *meas, int iters, int32_t *info,
                                  double *stats) {
    sb_clear_error();
    SB_REQUIRE(h && poses && fixed && info && stats, "null pointer");
    SB_REQUIRE(n_edges == 0 || (v0 && v1 && meas), "null edge arrays");
    SB_REQUIRE(n_vertices >= 1 && n_vertices <= h->max_vertices, "n_vertices out of range [1, max_vertices]");
    SB_REQUIRE(n_edges >= 0 && n_edges <= h->max_edges, "n_edges out of range [0, max_edges]");
    SB_TRY(sb_use_device(h->device));
    cudaStream_t s = h->stream;
    SB_CUDA(cudaMemcpyAsync(h->d_poses, poses, (size_t)n_vertices * 56, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(h->d_fixed, fixed, (size_t)n_vertices, cudaMemcpyHostToDevice, s));
    if (n_edges > 0) {
        SB_CUDA(cudaMemcpyAsync(h->d_v0, v0, (size_t)n_edges * 4, cudaMemcpyHostToDevice, s));
        SB_CUDA(cudaMemcpyAsync(h->d_v1, v1, (size_t)n_edges * 4, cudaMemcpyHostToDevice, s));
        SB_CUDA(cudaMemcpyAsync(h->d_meas, meas, (size_t)n_edges * 56, cudaMemcpyHostToDevice, s));
    }
    double *d_stats = reinterpret_cast<double *>(h->d_info + 8);
    SB_TRY(sb_posegraph_solve_dev(h, n_vertices, h->d_poses, h->d_fixed, n_edges, h->d_v0, h->d_v1, h->d_meas, iters,
                                  h->d_info, d_stats));
    SB_CUDA(cudaMemcpyAsync(poses, h->d_poses, (size_t)n_vertices * 56, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(info, h->d_info, 16, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    if (info[0] < 0) {
        sb_set_error(info[1] == -2 ? "more than 64 long-range (non-chain) edges between free vertices" : "edge endpoints out of range");
        return info[1] == -2 ? SB_ERR_CAPACITY : SB_ERR_INVALID;
    }
    return SB_OK;
}

#ifdef PG_PROFILE
extern "C" int sb_posegraph_debug_profile(long long *out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_pg_prof, sizeof(long long) * 16);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_pg_prof, z, sizeof(z));
    return 0;
}
#endif
