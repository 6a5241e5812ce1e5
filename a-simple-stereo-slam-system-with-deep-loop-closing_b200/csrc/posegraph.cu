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
// edges.  It is solved exactly (a direct method, like the reference's sparse Cholesky):
//   * T^-1 [b, J^T] by BLOCK CYCLIC REDUCTION: at level l every second still-active vertex is eliminated
//     (6x6 inverse, Schur update of its two neighbours), so the 733-step serial recurrence of a block
//     Thomas solve becomes ceil(log2 n) = 10 levels of work that is parallel over vertices and over the
//     1 + 6R right-hand sides;
//   * the Woodbury identity with a dense Cholesky of the 6R x 6R capacitance matrix in shared memory.
// One thread-block CLUSTER (8 CTAs on 8 SMs, hardware cluster barrier between the phases) runs the whole
// optimisation of one graph in one launch: numeric Jacobians, assembly, all LM trials.  Every sum runs in a
// fixed order: results are bit-reproducible.  R is bounded only by the workspace chosen at create time.
#include <cooperative_groups.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "se3.cuh"

namespace cg = cooperative_groups;

#ifndef PG_THREADS
#define PG_THREADS 512
#endif
#ifndef PG_CLUSTER
#define PG_CLUSTER 16     // CTAs of the cluster: 16 needs the non-portable cluster size (one GPC of a B200 has the SMs for it)
#endif
#define PG_DEFAULT_LOOPS 64
#define PG_SMEM_LOOPS 26  // capacitance matrices up to (6 x 26)^2 doubles = 190 KB are factored in shared memory

enum { PG_NONE = 0, PG_DIAG = 1, PG_CHAIN = 2, PG_LOOP = 3 };

#ifdef PG_PROFILE  // per-phase clock64 totals (tools/pg_profile.py); the product build has no trace of it
__device__ long long g_pg_prof[16];
#define PG_T(i) do { cluster.sync(); if (gtid == 0) { long long t_ = clock64(); g_pg_prof[i] += t_ - t_prev; t_prev = t_; } } while (0)
#else
#define PG_T(i) ((void)0)
#endif

struct sb_posegraph {
    int device, max_vertices, max_edges, max_loops;
    cudaStream_t stream, own_stream;
    double *d_poses, *d_meas;
    uint8_t *d_fixed;
    int32_t *d_v0, *d_v1, *d_info;
    // workspace
    double *d_work;
    int32_t *d_iwork;
    size_t work_doubles, iwork_ints, smem_bytes;
};

struct PgArgs {
    int n, m, iters, max_loops, smem_loops;
    double *poses;
    const uint8_t *fixed;
    const int32_t *v0, *v1;
    const double *meas;
    int32_t *info;  // [4]: LM iterations, trials, free vertices, long-range edges  (info[0] < 0: error)
    double *stats;  // [2]: chi2 at start, chi2 at the end
    double *Rt, *Rtb, *Zinv, *err, *Ji, *Jj, *A, *B, *bvec, *D, *L0, *L1, *G, *U, *Q, *M, *yv, *x, *part;
    int32_t *fidx, *vof, *ecls, *eloop, *inc_start, *inc_edge, *loop_edge, *flags;
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
// cluster-wide reductions: per-CTA partials through global memory, double-buffered by `phase` so that one cluster
// barrier per reduction is enough; every CTA adds the partials in the same order -> identical values everywhere
static __device__ double pg_cluster_sum(cg::cluster_group &cluster, double v, double *red, double *part, int &phase) {
    const double s = pg_block_sum(v, red);
    double *buf = part + (phase & 1) * 2 * PG_CLUSTER;
    phase++;
    if (threadIdx.x == 0) { __stcg(buf + blockIdx.x, s); }
    __threadfence();
    cluster.sync();
    double t = 0;
    for (int b = 0; b < (int)gridDim.x; b++) t += __ldcg(buf + b);
    return t;
}
static __device__ double pg_cluster_max(cg::cluster_group &cluster, double v, double *red, double *part, int &phase) {
    const double s = pg_block_max(v, red);
    double *buf = part + (phase & 1) * 2 * PG_CLUSTER;
    phase++;
    if (threadIdx.x == 0) { __stcg(buf + blockIdx.x, s); }
    __threadfence();
    cluster.sync();
    double t = __ldcg(buf);
    for (int b = 1; b < (int)gridDim.x; b++) t = fmax(t, __ldcg(buf + b));
    return t;
}
// barrier between two phases that exchange data through global memory
static __device__ __forceinline__ void pg_sync(cg::cluster_group &cluster) {
    __threadfence();
    cluster.sync();
}

// EdgePoseGraph::computeError with explicit vertex poses
static __device__ __forceinline__ void pg_edge_error(const double *Zinv, const double *T0, const double *T1, double *e) {
    double T1i[12], P[12];
    se3_inv(T1, T1i);
    se3_mul(T0, T1i, P);
    se3_mul(Zinv, P, P);
    se3_log(P, e);
}

// inverse of a symmetric positive definite 6x6 block by Cholesky (lower triangle of D is read); returns false when a
// pivot is not positive
static __device__ bool pg_inv6(const double *D, double *G) {
    double Lc[21], Li[21];  // packed lower triangles, (r, c) at r (r + 1) / 2 + c
    bool ok = true;
#pragma unroll
    for (int c = 0; c < 6; c++) {
        double d = D[6 * c + c];
#pragma unroll
        for (int k = 0; k < c; k++) d -= Lc[c * (c + 1) / 2 + k] * Lc[c * (c + 1) / 2 + k];
        if (!(d > 0)) ok = false;
        const double inv = rsqrt(d);
        Lc[c * (c + 1) / 2 + c] = inv;  // the diagonal holds the reciprocal
#pragma unroll
        for (int r = c + 1; r < 6; r++) {
            double v = 0.5 * (D[6 * r + c] + D[6 * c + r]);  // symmetrise against round-off
#pragma unroll
            for (int k = 0; k < c; k++) v -= Lc[r * (r + 1) / 2 + k] * Lc[c * (c + 1) / 2 + k];
            Lc[r * (r + 1) / 2 + c] = v * inv;
        }
    }
    // Li = Lc^-1 (lower triangular)
#pragma unroll
    for (int c = 0; c < 6; c++) {
        Li[c * (c + 1) / 2 + c] = Lc[c * (c + 1) / 2 + c];
#pragma unroll
        for (int r = c + 1; r < 6; r++) {
            double v = 0;
#pragma unroll
            for (int k = c; k < r; k++) v -= Lc[r * (r + 1) / 2 + k] * Li[k * (k + 1) / 2 + c];
            Li[r * (r + 1) / 2 + c] = v * Lc[r * (r + 1) / 2 + r];
        }
    }
    // G = Li^T Li
#pragma unroll
    for (int i = 0; i < 6; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double v = 0;
#pragma unroll
            for (int k = i; k < 6; k++) v += Li[k * (k + 1) / 2 + i] * Li[k * (k + 1) / 2 + j];
            G[6 * i + j] = v;
            G[6 * j + i] = v;
        }
    return ok;
}

__global__ void __launch_bounds__(PG_THREADS) k_posegraph(const PgArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(16) double smM[];  // CTA 0: the capacitance matrix during its factorisation
    __shared__ double red[16];
    __shared__ double s_diag[6 * PG_SMEM_LOOPS];
    __shared__ int s_bad;
    const int tid = threadIdx.x;
    const int gtid = blockIdx.x * PG_THREADS + tid, gsz = gridDim.x * PG_THREADS;
    const int n = a.n, m = a.m;
    int phase = 0;

    // ---- set-up: free-vertex numbering, edge classes, vertex -> edge incidence (serial, once)
    if (gtid == 0) {
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
                if (R >= a.max_loops) { bad = 2; break; }
                a.eloop[e] = R;
                a.loop_edge[R] = e;
                R++;
            }
            a.inc_start[u0 + 1]++;
            a.inc_start[u1 + 1]++;
        }
        if (!bad) {
            for (int v = 0; v < n; v++) a.inc_start[v + 1] += a.inc_start[v];
            int *cursor = a.eloop + m;  // scratch of n ints behind eloop
            for (int v = 0; v < n; v++) cursor[v] = a.inc_start[v];
            for (int e = 0; e < m; e++) {  // ascending edge index per vertex: deterministic
                a.inc_edge[cursor[a.v0[e]]++] = e;
                a.inc_edge[cursor[a.v1[e]]++] = e;
            }
        }
        a.flags[0] = nf; a.flags[1] = R; a.flags[2] = bad;
    }
    pg_sync(cluster);
    const int nf = __ldcg(a.flags), R = __ldcg(a.flags + 1), bad0 = __ldcg(a.flags + 2);
    if (bad0) {
        if (gtid == 0) { a.info[0] = -1; a.info[1] = -bad0; a.info[2] = a.info[3] = 0; }
        return;
    }
    const int NC = 1 + 6 * R, n6r = 6 * R;
    for (int v = gtid; v < n; v += gsz) {
        quat_to_R(a.poses + 7 * v, a.Rt + 12 * v);
        a.Rt[12 * v + 9] = a.poses[7 * v + 4]; a.Rt[12 * v + 10] = a.poses[7 * v + 5]; a.Rt[12 * v + 11] = a.poses[7 * v + 6];
    }
    for (int e = gtid; e < m; e += gsz) {
        double Z[12];
        quat_to_R(a.meas + 7 * e, Z);
        Z[9] = a.meas[7 * e + 4]; Z[10] = a.meas[7 * e + 5]; Z[11] = a.meas[7 * e + 6];
        se3_inv(Z, a.Zinv + 12 * e);
    }
    pg_sync(cluster);

    double lambda = 0, ni = 2, chi_start = 0, chi_end = 0;
    int lm_iters = 0, trials = 0;
    bool terminated = false;
#ifdef PG_PROFILE
    long long t_prev = clock64();
#endif

    // computeActiveErrors + chi2, edges over the whole cluster
    auto errors = [&]() -> double {
        double chi = 0;
        for (int e = gtid; e < m; e += gsz) {
            double r[6];
            pg_edge_error(a.Zinv + 12 * e, a.Rt + 12 * a.v0[e], a.Rt + 12 * a.v1[e], r);
#pragma unroll
            for (int k = 0; k < 6; k++) { a.err[6 * e + k] = r[k]; chi += r[k] * r[k]; }
        }
        return pg_cluster_sum(cluster, chi, red, a.part, phase);
    };

    for (int it = 0; it < a.iters && !terminated; it++) {
        double currentChi = errors();
        PG_T(0);
        if (it == 0) chi_start = currentChi;
        chi_end = currentChi;
        if (nf == 0) break;
        // ---- numeric Jacobians (g2o BaseBinaryEdge::linearizeOplus): one work item per (edge, vertex, column)
        for (int item = gtid; item < m * 12; item += gsz) {
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
        pg_sync(cluster);
        PG_T(1);
        // ---- assembly, one thread per (free vertex, block row): gather over the incident edges in ascending edge index
        double mx = 0;
        for (int item = gtid; item < nf * 6; item += gsz) {
            const int p = item / 6, i = item - 6 * p;
            const int v = a.vof[p];
            double Ap[6] = {0, 0, 0, 0, 0, 0}, Bp[6] = {0, 0, 0, 0, 0, 0}, g = 0, hd = 0;
            for (int q = a.inc_start[v]; q < a.inc_start[v + 1]; q++) {
                const int e = a.inc_edge[q];
                const int cls = a.ecls[e];
                const bool first = a.v0[e] == v;
                const double *Jv = (first ? a.Ji : a.Jj) + 36 * e;
                const double *er = a.err + 6 * e;
                double ji[6];
#pragma unroll
                for (int r = 0; r < 6; r++) ji[r] = Jv[6 * r + i];
                double s = 0, dd = 0;
#pragma unroll
                for (int r = 0; r < 6; r++) { s += ji[r] * er[r]; dd += ji[r] * ji[r]; }
                g -= s;
                hd += dd;
                if (cls == PG_DIAG || cls == PG_CHAIN) {
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        double t = 0;
#pragma unroll
                        for (int r = 0; r < 6; r++) t += ji[r] * Jv[6 * r + j];
                        Ap[j] += t;
                    }
                }
                if (cls == PG_CHAIN) {
                    const int other = first ? a.v1[e] : a.v0[e];
                    if (a.fidx[other] == p - 1) {  // block (p, p-1) = J_p^T J_{p-1}
                        const double *Jo = (first ? a.Jj : a.Ji) + 36 * e;
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            double t = 0;
#pragma unroll
                            for (int r = 0; r < 6; r++) t += ji[r] * Jo[6 * r + j];
                            Bp[j] += t;
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 6; j++) { a.A[36 * p + 6 * i + j] = Ap[j]; a.B[36 * p + 6 * i + j] = Bp[j]; }
            a.bvec[6 * p + i] = g;
            mx = fmax(mx, fabs(hd));
        }
        if (it == 0) {  // computeLambdaInit
            lambda = 1e-5 * pg_cluster_max(cluster, mx, red, a.part, phase);
            ni = 2;
        } else {
            pg_sync(cluster);
        }

        double rho = 0;
        int qmax = 0;
        PG_T(2);
        do {
            // ---- push(), the damped block-tridiagonal system and the right-hand sides
            //      Q[p][c][6]: column 0 = b, column 1 + 6 r + k = row k of J_loop,r (as a column of J^T)
            for (int k = gtid; k < 12 * n; k += gsz) a.Rtb[k] = a.Rt[k];
            for (int k = gtid; k < 36 * nf; k += gsz) {
                const int rc = k % 36;
                a.D[k] = a.A[k] + ((rc / 6 == rc % 6) ? lambda : 0.0);
                a.L0[k] = a.B[k];
            }
            for (size_t idx = gtid; idx < (size_t)nf * NC * 6; idx += gsz) {
                const int p = (int)(idx / ((size_t)NC * 6)), rem = (int)(idx - (size_t)p * NC * 6);
                a.Q[idx] = rem < 6 ? a.bvec[6 * p + rem] : 0.0;
            }
            if (gtid == 0) a.flags[3] = 0;
            pg_sync(cluster);
            for (int idx = gtid; idx < R * 6 * 2; idx += gsz) {
                const int r = idx / 12, k = (idx % 12) / 2, side = idx % 2;
                const int e = a.loop_edge[r];
                const int p = a.fidx[side ? a.v1[e] : a.v0[e]];
                const double *J = (side ? a.Jj : a.Ji) + 36 * e;
                double *q = a.Q + ((size_t)p * NC + 1 + 6 * r + k) * 6;
#pragma unroll
                for (int i = 0; i < 6; i++) q[i] = J[6 * k + i];
            }
            PG_T(3);
            // ---- block cyclic reduction, forward.  Level l (stride s = 2^l): active vertices are p with (p + 1) % s == 0;
            //      those with (p + 1) % 2s == s are ELIMINATED, the others survive.  Row p reads
            //      Lp x_{p-s} + D_p x_p + L_{p+s}^T x_{p+s} = r_p   (L_p couples p to its left active neighbour).
            //      Phase a:  [right-hand sides of the previous level's survivors] + [G_e = D_e^-1, U_e = L_{e+s} saved]
            //      Phase b:  [y_e = G_e r_e for every column] + [D, L of the survivors]
            int levels = 0;
            for (int s = 1; s - 1 < nf; s <<= 1) levels++;
            for (int l = 0, s = 1; l < levels; l++, s <<= 1) {
                const double *Lc = (l & 1) ? a.L1 : a.L0;
                double *Ln = (l & 1) ? a.L0 : a.L1;
                // phase a (1): right-hand sides of the survivors of level l - 1 (stride s / 2), every column
                if (l > 0) {
                    const int sp = s >> 1;   // previous stride; survivors v = 2 sp - 1 + 2 sp k = s - 1 + s k
                    const double *Lp = ((l - 1) & 1) ? a.L1 : a.L0;
                    const int nsv = (nf - (s - 1) + s - 1) / s;   // vertices v = s - 1 + s k < nf
                    for (size_t item = gtid; item < (size_t)nsv * NC; item += gsz) {
                        const int k = (int)(item / NC), c = (int)(item - (size_t)k * NC);
                        const int v = s - 1 + s * k;
                        double *rv = a.Q + ((size_t)v * NC + c) * 6;
                        double acc[6];
#pragma unroll
                        for (int i = 0; i < 6; i++) acc[i] = rv[i];
                        {   // left eliminated neighbour v - sp (always exists)
                            const double *y = a.Q + ((size_t)(v - sp) * NC + c) * 6, *Lv = Lp + 36 * v;
#pragma unroll
                            for (int i = 0; i < 6; i++)
#pragma unroll
                                for (int j = 0; j < 6; j++) acc[i] -= Lv[6 * i + j] * y[j];
                        }
                        if (v + sp < nf) {   // right eliminated neighbour: its L couples it to v
                            const double *y = a.Q + ((size_t)(v + sp) * NC + c) * 6, *Le = Lp + 36 * (v + sp);
#pragma unroll
                            for (int i = 0; i < 6; i++)
#pragma unroll
                                for (int j = 0; j < 6; j++) acc[i] -= Le[6 * j + i] * y[j];
                        }
#pragma unroll
                        for (int i = 0; i < 6; i++) rv[i] = acc[i];
                    }
                }
                // phase a (2): inverse of the eliminated vertices' diagonal blocks, copy of their right neighbour's coupling
                const int nel = (nf - (s - 1) + 2 * s - 1) / (2 * s);   // vertices e = s - 1 + 2 s k < nf
                for (int k = gtid; k < nel; k += gsz) {
                    const int e = s - 1 + 2 * s * k;
                    if (!pg_inv6(a.D + 36 * e, a.G + 36 * e)) a.flags[3] = 1;
                    if (e + s < nf) {   // all 18 loads first, then the stores (the compiler must assume the two arrays alias)
                        const double2 *src = reinterpret_cast<const double2 *>(Lc + 36 * (e + s));
                        double2 *dst = reinterpret_cast<double2 *>(a.U + 36 * e);
                        double2 tmp[18];
#pragma unroll
                        for (int q = 0; q < 18; q++) tmp[q] = src[q];
#pragma unroll
                        for (int q = 0; q < 18; q++) dst[q] = tmp[q];
                    }
                }
                pg_sync(cluster);
                PG_T(8);
                // phase b (1): y_e = G_e r_e in place, every column
                for (size_t item = gtid; item < (size_t)nel * NC; item += gsz) {
                    const int k = (int)(item / NC), c = (int)(item - (size_t)k * NC);
                    const int e = s - 1 + 2 * s * k;
                    double *re = a.Q + ((size_t)e * NC + c) * 6;
                    const double *G = a.G + 36 * e;
                    double r[6], y[6];
#pragma unroll
                    for (int i = 0; i < 6; i++) r[i] = re[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) {
                        double v = 0;
#pragma unroll
                        for (int j = 0; j < 6; j++) v += G[6 * i + j] * r[j];
                        y[i] = v;
                    }
#pragma unroll
                    for (int i = 0; i < 6; i++) re[i] = y[i];
                }
                // phase b (2): Schur update of the survivors v = 2 s - 1 + 2 s k, one thread per (vertex, block row)
                const int nsv2 = (nf - (2 * s - 1) + 2 * s - 1) / (2 * s);
                for (int item = gtid; item < 6 * nsv2 && 2 * s - 1 < nf; item += gsz) {
                    const int k = item / 6, i = item - 6 * k;
                    const int v = 2 * s - 1 + 2 * s * k;
                    double drow[6], lrow[6] = {0, 0, 0, 0, 0, 0};
#pragma unroll
                    for (int j = 0; j < 6; j++) drow[j] = a.D[36 * v + 6 * i + j];
                    {   // left eliminated neighbour e1 = v - s:  T = L_v G_e1;  D -= T L_v^T;  L' = -T L_e1
                        const int e1 = v - s;
                        const double *Lv = Lc + 36 * v, *G = a.G + 36 * e1, *Le = Lc + 36 * e1;
                        double t[6];
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            double x = 0;
#pragma unroll
                            for (int q = 0; q < 6; q++) x += Lv[6 * i + q] * G[6 * q + j];
                            t[j] = x;
                        }
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            double x = 0;
#pragma unroll
                            for (int q = 0; q < 6; q++) x += t[q] * Lv[6 * j + q];
                            drow[j] -= x;
                        }
                        if (e1 - s >= 0) {
#pragma unroll
                            for (int j = 0; j < 6; j++) {
                                double x = 0;
#pragma unroll
                                for (int q = 0; q < 6; q++) x += t[q] * Le[6 * q + j];
                                lrow[j] = -x;
                            }
                        }
                    }
                    if (v + s < nf) {   // right eliminated neighbour e2 = v + s:  T = L_e2^T G_e2;  D -= T L_e2
                        const int e2 = v + s;
                        const double *Le = Lc + 36 * e2, *G = a.G + 36 * e2;
                        double t[6];
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            double x = 0;
#pragma unroll
                            for (int q = 0; q < 6; q++) x += Le[6 * q + i] * G[6 * q + j];
                            t[j] = x;
                        }
#pragma unroll
                        for (int j = 0; j < 6; j++) {
                            double x = 0;
#pragma unroll
                            for (int q = 0; q < 6; q++) x += t[q] * Le[6 * q + j];
                            drow[j] -= x;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 6; j++) { a.D[36 * v + 6 * i + j] = drow[j]; Ln[36 * v + 6 * i + j] = lrow[j]; }
                }
                pg_sync(cluster);
                PG_T(9);
            }
            int ok = !__ldcg(a.flags + 3);
            PG_T(4);
            // ---- backward: x_e = y_e - G_e (L_e x_{e-s} + U_e^T x_{e+s}), from the last level down; Q ends as T^-1 [b, J^T]
            for (int l = levels - 1, s = 1 << (levels - 1); l >= 0 && ok; l--, s >>= 1) {
                const double *Lc = (l & 1) ? a.L1 : a.L0;
                const int nel = (nf - (s - 1) + 2 * s - 1) / (2 * s);
                for (size_t item = gtid; item < (size_t)nel * NC; item += gsz) {
                    const int k = (int)(item / NC), c = (int)(item - (size_t)k * NC);
                    const int e = s - 1 + 2 * s * k;
                    double w[6] = {0, 0, 0, 0, 0, 0};
                    bool any = false;
                    if (e - s >= 0) {
                        const double *xa = a.Q + ((size_t)(e - s) * NC + c) * 6, *Le = Lc + 36 * e;
#pragma unroll
                        for (int i = 0; i < 6; i++)
#pragma unroll
                            for (int j = 0; j < 6; j++) w[i] += Le[6 * i + j] * xa[j];
                        any = true;
                    }
                    if (e + s < nf) {
                        const double *xc = a.Q + ((size_t)(e + s) * NC + c) * 6, *Ue = a.U + 36 * e;
#pragma unroll
                        for (int i = 0; i < 6; i++)
#pragma unroll
                            for (int j = 0; j < 6; j++) w[i] += Ue[6 * j + i] * xc[j];
                        any = true;
                    }
                    if (any) {
                        double *xe = a.Q + ((size_t)e * NC + c) * 6;
                        const double *G = a.G + 36 * e;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            double v = 0;
#pragma unroll
                            for (int j = 0; j < 6; j++) v += G[6 * i + j] * w[j];
                            xe[i] -= v;
                        }
                    }
                }
                pg_sync(cluster);
            }
            PG_T(5);
            if (ok && R > 0) {
                // ---- capacitance matrix M = I + J Z and right-hand side v = J x0
                for (int idx = gtid; idx < n6r * (n6r + 1); idx += gsz) {
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
                pg_sync(cluster);
                PG_T(10);
                // ---- dense Cholesky of M + the two substitutions by CTA 0 (in shared memory when it fits)
                if (blockIdx.x == 0) {
                    const bool in_smem = R <= a.smem_loops;
                    double *Mw = in_smem ? smM : a.M;
                    const int ld = in_smem ? n6r + 1 : n6r;   // odd leading dimension in shared memory: 6R is a multiple of 32 banks' worth for R = 16
                    if (in_smem)
                        for (int k = tid; k < n6r * n6r; k += PG_THREADS) smM[(k / n6r) * ld + k % n6r] = a.M[k];
                    if (tid == 0) s_bad = 0;
                    __syncthreads();
                    // right-looking, one barrier per column: every thread reads the pivot and forms the scaled column entries
                    // it needs itself; column j of L goes to ROW j of the upper triangle (M[j][i], i > j), the trailing update
                    // works on the lower triangle, the diagonal of L is kept apart.
                    double *diag = in_smem ? s_diag : a.yv + n6r;
                    const int r0 = tid >> 3, c0 = tid & 7;
                    for (int j = 0; j < n6r; j++) {
                        const double d = Mw[(size_t)j * ld + j];
                        if (!(d > 0)) { if (tid == 0) s_bad = 1; break; }  // uniform: every thread reads the same entry
                        const double inv = rsqrt(d);
                        if (tid == 0) diag[j] = inv;   // reciprocal pivots: no division in the substitutions
                        for (int i = j + 1 + r0; i < n6r; i += PG_THREADS / 8) {
                            const double lij = Mw[(size_t)i * ld + j] * inv;
                            if (c0 == 0) Mw[(size_t)j * ld + i] = lij;
                            for (int k = j + 1 + c0; k <= i; k += 8) Mw[(size_t)i * ld + k] -= lij * (Mw[(size_t)k * ld + j] * inv);
                        }
                        __syncthreads();
                    }
                    __syncthreads();
                    if (!s_bad && tid < 32 && in_smem) {
                        // M y = v by one warp, the unknowns in registers (lane l owns rows l, l + 32, ...): per step one broadcast of
                        // the solved unknown and one multiply-add per owned row; L(i, j) = M[j][i] (row j of the upper triangle).
                        const int lane = tid;
                        constexpr int NR = (6 * PG_SMEM_LOOPS + 31) / 32;
                        double xr[NR];
#pragma unroll
                        for (int q = 0; q < NR; q++) xr[q] = lane + 32 * q < n6r ? __ldcg(a.yv + lane + 32 * q) : 0.0;
                        for (int j = 0; j < n6r; j++) {   // forward, column oriented
                            double src = 0;
#pragma unroll
                            for (int q = 0; q < NR; q++) src = (j >> 5) == q ? xr[q] : src;
                            const double yj = __shfl_sync(0xffffffffu, src, j & 31) * diag[j];
                            const double *row = Mw + (size_t)j * ld;
#pragma unroll
                            for (int q = 0; q < NR; q++) {
                                const int i = lane + 32 * q;
                                if (i == j) xr[q] = yj;
                                else if (i > j && i < n6r) xr[q] -= row[i] * yj;
                            }
                        }
                        for (int j = n6r - 1; j >= 0; j--) {   // backward with L^T: L^T(i, j) = L(j, i) = M[i][j], i < j
                            double src = 0;
#pragma unroll
                            for (int q = 0; q < NR; q++) src = (j >> 5) == q ? xr[q] : src;
                            const double xj = __shfl_sync(0xffffffffu, src, j & 31) * diag[j];
#pragma unroll
                            for (int q = 0; q < NR; q++) {
                                const int i = lane + 32 * q;
                                if (i == j) xr[q] = xj;
                                else if (i < j) xr[q] -= Mw[(size_t)i * ld + j] * xj;
                            }
                        }
#pragma unroll
                        for (int q = 0; q < NR; q++)
                            if (lane + 32 * q < n6r) a.yv[lane + 32 * q] = xr[q];
                    } else if (!s_bad && tid < 32) {  // workspace in global memory (more loop edges than fit shared memory)
                        const int lane = tid;
                        for (int i = 0; i < n6r; i++) {
                            double v = 0;
                            for (int k = lane; k < i; k += 32) v += Mw[(size_t)k * ld + i] * a.yv[k];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if (lane == 0) a.yv[i] = (a.yv[i] - v) * diag[i];
                            __syncwarp();
                        }
                        for (int i = n6r - 1; i >= 0; i--) {
                            double v = 0;
                            for (int k = i + 1 + lane; k < n6r; k += 32) v += Mw[(size_t)i * ld + k] * a.yv[k];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
                            if (lane == 0) a.yv[i] = (a.yv[i] - v) * diag[i];
                            __syncwarp();
                        }
                    }
                    __syncthreads();
                    if (tid == 0 && s_bad) a.flags[3] = 1;
                }
                pg_sync(cluster);
                PG_T(11);
                ok = !__ldcg(a.flags + 3);
            }
            if (ok) {
                // ---- x = x0 - Z y
                for (int idx = gtid; idx < nf * 6; idx += gsz) {
                    const int p = idx / 6, i = idx % 6;
                    const double *q = a.Q + (size_t)p * NC * 6;
                    double s = q[i];
                    for (int c = 0; c < n6r; c++) s -= q[(1 + c) * 6 + i] * __ldcg(a.yv + c);
                    a.x[idx] = s;
                }
            }
            pg_sync(cluster);
            PG_T(6);
            double scale = 0;
            if (ok) {
                for (int idx = gtid; idx < nf * 6; idx += gsz) scale += a.x[idx] * (lambda * a.x[idx] + a.bvec[idx]);
                for (int p = gtid; p < nf; p += gsz) pose_oplus(a.Rt + 12 * a.vof[p], a.x + 6 * p);
            }
            scale = pg_cluster_sum(cluster, scale, red, a.part, phase);
            double tempChi = errors();
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
                for (int k = gtid; k < 12 * n; k += gsz) a.Rt[k] = a.Rtb[k];  // pop()
                pg_sync(cluster);
            }
            qmax++;
            PG_T(7);
        } while (rho < 0 && qmax < 10);
        lm_iters++;
        if (qmax == 10 || rho == 0) terminated = true;
    }
    pg_sync(cluster);
    for (int v = gtid; v < n; v += gsz) {
        if (a.fixed[v]) continue;  // fixed vertices keep their input bits
        R_to_quat(a.Rt + 12 * v, a.poses + 7 * v);
        a.poses[7 * v + 4] = a.Rt[12 * v + 9]; a.poses[7 * v + 5] = a.Rt[12 * v + 10]; a.poses[7 * v + 6] = a.Rt[12 * v + 11];
    }
    if (gtid == 0) {
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

static int pg_create(sb_posegraph_t **out, int device, int max_vertices, int max_edges, int max_loops) {
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_vertices >= 2 && max_vertices <= (1 << 20), "max_vertices out of range");
    SB_REQUIRE(max_edges >= 1 && max_edges <= (1 << 22), "max_edges out of range");
    SB_REQUIRE(max_loops >= 1 && max_loops <= 4096, "max_loops out of range [1, 4096]");
    const size_t n = max_vertices, m = max_edges, NC = 1 + 6 * (size_t)max_loops, R6 = 6 * (size_t)max_loops;
    SB_REQUIRE(n * NC * 6 * 8 <= ((size_t)48 << 30), "max_vertices x max_loops needs more than 48 GB of workspace");
    SB_TRY(sb_use_device(device));
    sb_posegraph *h = new sb_posegraph();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_vertices = max_vertices;
    h->max_edges = max_edges;
    h->max_loops = max_loops;
    h->work_doubles = 24 * n + 12 * m + 6 * m + 72 * m + 36 * n * 7 + 6 * n * 2 + n * NC * 6 + R6 * R6 + 2 * R6 + 4 * PG_CLUSTER + 8;
    h->iwork_ints = 2 * n + 3 * m + n + (n + 1) + 2 * m + max_loops + 16;
    const int sl = max_loops < PG_SMEM_LOOPS ? max_loops : PG_SMEM_LOOPS;
    h->smem_bytes = (size_t)6 * sl * (6 * sl + 1) * 8;
    cudaError_t e = cudaMalloc((void **)&h->d_poses, n * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_meas, m * 56);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_fixed, n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_v0, m * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_v1, m * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_info, 64);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_work, h->work_doubles * 8);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_iwork, h->iwork_ints * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_posegraph, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)6 * PG_SMEM_LOOPS * (6 * PG_SMEM_LOOPS + 1) * 8));
    if (e == cudaSuccess && PG_CLUSTER > 8) e = cudaFuncSetAttribute(k_posegraph, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
        sb_set_error("sb_posegraph_create: %s", cudaGetErrorString(e));
        free_pg(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_posegraph_create(sb_posegraph_t **out, int device, int max_vertices, int max_edges) {
    SB_NVTX_FN();
    return pg_create(out, device, max_vertices, max_edges, PG_DEFAULT_LOOPS);
}
// max_loops: the largest number of long-range (non-chain) edges between free vertices a solve may contain.  The reference
// re-adds every historical loop edge on each PoseGraphOptimization (src/loopclosing.cpp:585-599), so a long run needs a
// generous bound; the workspace grows with max_vertices x max_loops x 288 bytes.
extern "C" int sb_posegraph_create_loops(sb_posegraph_t **out, int device, int max_vertices, int max_edges, int max_loops) {
    SB_NVTX_FN();
    return pg_create(out, device, max_vertices, max_edges, max_loops);
}

extern "C" int sb_posegraph_destroy(sb_posegraph_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_pg(h);
    }
    return SB_OK;
}

extern "C" int sb_posegraph_set_stream(sb_posegraph_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_posegraph_solve_dev(sb_posegraph_t *h, int n_vertices, double *d_poses, const uint8_t *d_fixed,
                                      int n_edges, const int32_t *d_v0, const int32_t *d_v1, const double *d_meas,
                                      int iters, int32_t *d_info, double *d_stats) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && d_poses && d_fixed && d_v0 && d_v1 && d_meas && d_info && d_stats, "null pointer");
    SB_REQUIRE(n_vertices >= 1 && n_vertices <= h->max_vertices, "n_vertices out of range [1, max_vertices]");
    SB_REQUIRE(n_edges >= 0 && n_edges <= h->max_edges, "n_edges out of range [0, max_edges]");
    SB_REQUIRE(iters >= 1, "iters must be positive");
    SB_TRY(sb_use_device(h->device));
    const size_t n = h->max_vertices, m = h->max_edges, NC = 1 + 6 * (size_t)h->max_loops, R6 = 6 * (size_t)h->max_loops;
    PgArgs a;
    a.n = n_vertices; a.m = n_edges; a.iters = iters; a.max_loops = h->max_loops;
    a.smem_loops = h->max_loops < PG_SMEM_LOOPS ? h->max_loops : PG_SMEM_LOOPS;
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
    a.D = w; w += 36 * n;
    a.L0 = w; w += 36 * n;
    a.L1 = w; w += 36 * n;
    a.G = w; w += 36 * n;
    a.U = w; w += 36 * n;
    a.bvec = w; w += 6 * n;
    a.x = w; w += 6 * n;
    a.Q = w; w += n * NC * 6;
    a.M = w; w += R6 * R6;
    a.yv = w; w += 2 * R6;
    a.part = w; w += 4 * PG_CLUSTER;
    int32_t *iw = h->d_iwork;
    a.fidx = iw; iw += n;
    a.vof = iw; iw += n;
    a.ecls = iw; iw += m;
    a.eloop = iw; iw += m + n;  // + n ints of cursor scratch
    a.inc_start = iw; iw += n + 1;
    a.inc_edge = iw; iw += 2 * m;
    a.loop_edge = iw; iw += h->max_loops;
    a.flags = iw; iw += 8;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(PG_CLUSTER);
    cfg.blockDim = dim3(PG_THREADS);
    cfg.dynamicSmemBytes = h->smem_bytes;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PG_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SB_CUDA(cudaLaunchKernelEx(&cfg, k_posegraph, a));
    return SB_OK;
}

extern "C" int sb_posegraph_solve(sb_posegraph_t *h, int n_vertices, double *poses, const uint8_t *fixed, int n_edges,
                                  const int32_t *v0, const int32_t *v1, const double *meas, int iters, int32_t *info,
                                  double *stats) {
    SB_NVTX_FN();
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
        sb_set_error(info[1] == -2 ? "more long-range (non-chain) edges between free vertices than the handle's max_loops (sb_posegraph_create_loops)" : "edge endpoints out of range");
        return info[1] == -2 ? SB_ERR_CAPACITY : SB_ERR_INVALID;
    }
    return SB_OK;
}

#ifdef PG_PROFILE
extern "C" int sb_posegraph_debug_profile(long long *out) {
    SB_NVTX_FN();
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_pg_prof, sizeof(long long) * 16);
    long long z[16] = {0};
    cudaMemcpyToSymbol(g_pg_prof, z, sizeof(z));
    return 0;
}
#endif
