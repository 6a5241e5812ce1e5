// qt_core.inl — level-synchronous formulation of ORBextractor::DistributeOctTree
// (reference src/ORBextractor.cpp:586-810, ExtractorNode::DivideNode :526-582).
//
// The reference walks a std::list sequentially.  The same result is obtained here in bulk-
// synchronous "passes" that one CTA executes cooperatively:
//   * the node table IS the list: entry p is the p-th node from the front;
//   * a pass splits a set of nodes in a given processing order.  Children are push_front-ed in
//     creation order, so after the pass the M new children occupy positions M-1-c (c = creation
//     index) and the surviving old nodes follow in their old order — both are prefix sums;
//   * phase 1 (reference :641-717) processes every node with more than one key in list order;
//   * phase 2 (:719-780) processes the children of the previous pass sorted by (key count,
//     creation index) descending — the documented replacement of the reference's (size, heap
//     pointer) sort — and stops at the first split that brings the list to N nodes; the stop
//     point is the first crossing of a prefix sum over the sorted order;
//   * the final "best response per node, first wins" (:789-807) uses the fact that the candidate
//     order the reference sees (cell row, cell column, y, x) is a pure function of position.
//
// This file is included twice: by orb_kernels.cu (device, one CTA per (image, level)) and, with
// QT_HOST defined, by tests/qt_host_model.cpp where every QT_FOR runs serially — one valid
// interleaving — so the pass logic is checked on the CPU against the literal oracle.
//
// Candidate word: x | y << 12 | response << 24   (x, y border-relative, integer FAST coordinates).

#ifndef QT_CORE_INL
#define QT_CORE_INL

#ifdef QT_HOST
#include <stdint.h>
#include <string.h>
#define QT_DEV
#define QT_FOR(i, n) for (int i = 0; i < (n); ++i)
#define QT_SYNC() ((void)0)
#define QT_TID0 (true)
static inline int qt_atomic_add(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline void qt_atomic_max64(unsigned long long *p, unsigned long long v) { if (v > *p) *p = v; }
#else
#define QT_DEV __device__ __forceinline__
#define QT_FOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define QT_SYNC() __syncthreads()
#define QT_TID0 (threadIdx.x == 0)
static __device__ __forceinline__ int qt_atomic_add(int *p, int v) { return atomicAdd(p, v); }
static __device__ __forceinline__ void qt_atomic_max64(unsigned long long *p, unsigned long long v) { atomicMax(p, v); }
#endif

struct QtCtx {
    // candidates
    const uint32_t *cand;  // [n] packed words (read-only)
    int n;
    uint16_t *cnode;  // [n] table position of the node holding the candidate
    uint8_t *cq;      // [n] quadrant chosen in the current pass
    // node tables (ping-pong), capacity ncap
    int16_t *box[2];   // [ncap*4] ulx, uly, brx, bry
    uint16_t *cnt[2];  // [ncap]
    int ncap;
    // work arrays
    int *childcnt;   // [ncap*4] keys per child, then the child's new position
    int *ord;        // [ncap] processing order: table positions
    int *cpre;       // [ncap] creation-index prefix (exclusive) in processing order
    int *spre;       // [ncap] list size after each split (inclusive) in processing order
    int *cbase;      // [ncap] per position: creation index of first child, -1 if not split
    int *ubase;      // [ncap] per position: rank among surviving nodes
    uint32_t *keys;  // [pow2 >= ncap] sort buffer
    int *scratch;    // [64] scan scratch / scalars
    // geometry
    int width, height;        // maxBorder - minBorder
    int N;                    // wanted number of nodes
    int nCols, wCell, hCell;  // FAST grid, for the canonical candidate order
};

// scratch slots
#define QT_S_TOTAL 40
#define QT_S_K 41
#define QT_S_V 42
#define QT_S_NEXP 43

// ---- block-wide exclusive scan of a[0..len) in place; returns the total ------------------------
static QT_DEV int qt_scan_excl(int *a, int len, int *scratch) {
#ifdef QT_HOST
    int run = 0;
    for (int i = 0; i < len; i++) { int v = a[i]; a[i] = run; run += v; }
    (void)scratch;
    return run;
#else
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    int carry = 0;
    for (int base = 0; base < len; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < len ? a[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) scratch[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = lane < nw ? scratch[lane] : 0;
            int s = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(0xffffffffu, s, o);
                if (lane >= o) s += y;
            }
            if (lane < nw) scratch[lane] = s - w;  // exclusive warp offsets
            if (lane == 31) scratch[32] = s;       // tile total
        }
        __syncthreads();
        if (i < len) a[i] = carry + scratch[wid] + x - v;
        carry += scratch[32];
        __syncthreads();
    }
    return carry;
#endif
}

// ---- bitonic sort, descending, P a power of two ------------------------------------------------
static QT_DEV void qt_sort_desc(uint32_t *keys, int P) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            QT_FOR(i, P) {
                int ixj = i ^ j;
                if (ixj > i) {
                    uint32_t a = keys[i], b = keys[ixj];
                    bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            QT_SYNC();
        }
    }
}

static QT_DEV int qt_pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// ---- one pass: split nodes of table `cur` (size S) into table `cur^1`; returns the new size -----
// sorted == 0: phase 1, every node with cnt > 1, list order, no early stop.
// sorted == 1: phase 2, nodes at positions < Mprev with cnt > 1, (cnt, creation) descending,
//              stop after the first split that reaches N.
// *Mout = number of children created, *nexp = how many of them hold more than one key.
static QT_DEV int qt_pass(QtCtx &c, int cur, int S, int sorted, int Mprev, int *Mout, int *nexp) {
    int16_t *box = c.box[cur], *nbox = c.box[cur ^ 1];
    uint16_t *cnt = c.cnt[cur], *ncnt = c.cnt[cur ^ 1];
    const int limit = sorted ? Mprev : S;

    QT_FOR(p, S * 4) c.childcnt[p] = 0;
    if (QT_TID0) c.scratch[QT_S_NEXP] = 0;
    QT_SYNC();
    // quadrant of every key of a splittable node (DivideNode :528-570)
    QT_FOR(i, c.n) {
        int p = c.cnode[i];
        if (p < limit && cnt[p] > 1) {
            uint32_t w = c.cand[i];
            int x = w & 0xfff, y = (w >> 12) & 0xfff;
            int ulx = box[4 * p], uly = box[4 * p + 1], brx = box[4 * p + 2], bry = box[4 * p + 3];
            int halfX = (brx - ulx + 1) >> 1;  // ceil((UR.x-UL.x)/2), non-negative operands
            int halfY = (bry - uly + 1) >> 1;
            int q = (x < ulx + halfX) ? ((y < uly + halfY) ? 0 : 2) : ((y < uly + halfY) ? 1 : 3);
            c.cq[i] = (uint8_t)q;
            qt_atomic_add(&c.childcnt[4 * p + q], 1);
        }
    }
    QT_SYNC();

    // processing order
    int V;
    if (!sorted) {
        QT_FOR(p, S) c.ubase[p] = cnt[p] > 1 ? 1 : 0;
        QT_SYNC();
        V = qt_scan_excl(c.ubase, S, c.scratch);
        QT_FOR(p, S) if (cnt[p] > 1) c.ord[c.ubase[p]] = p;
        QT_SYNC();
    } else {
        const int P = qt_pow2ceil(Mprev);
        QT_FOR(k, P) c.keys[k] = (k < Mprev && cnt[k] > 1) ? (((uint32_t)cnt[k] << 16) | (uint32_t)(Mprev - 1 - k)) : 0u;
        QT_FOR(p, S) c.ubase[p] = (p < Mprev && cnt[p] > 1) ? 1 : 0;
        QT_SYNC();
        V = qt_scan_excl(c.ubase, S, c.scratch);
        qt_sort_desc(c.keys, P);
        QT_FOR(k, V) c.ord[k] = Mprev - 1 - (int)(c.keys[k] & 0xffffu);
        QT_SYNC();
    }

    // prefix sums in processing order: creation index and running list size
    QT_FOR(k, V) {
        int p = c.ord[k];
        int m = (c.childcnt[4 * p] > 0) + (c.childcnt[4 * p + 1] > 0) + (c.childcnt[4 * p + 2] > 0) +
                (c.childcnt[4 * p + 3] > 0);
        c.cpre[k] = m;
        c.spre[k] = m - 1;
    }
    QT_SYNC();
    int Mall = qt_scan_excl(c.cpre, V, c.scratch);
    int incall = qt_scan_excl(c.spre, V, c.scratch);  // exclusive; inclusive = excl + own
    int K = V;
    if (sorted) {
        if (QT_TID0) c.scratch[QT_S_K] = V;
        QT_SYNC();
        QT_FOR(k, V) {
            int before = S + c.spre[k];
            int own = (k + 1 < V ? c.spre[k + 1] : incall) - c.spre[k];
            if (before < c.N && before + own >= c.N) c.scratch[QT_S_K] = k + 1;  // unique first crossing
        }
        QT_SYNC();
        K = c.scratch[QT_S_K];
        QT_SYNC();
    }
    const int M = K < V ? c.cpre[K] : Mall;

    // mark split nodes, rank survivors
    QT_FOR(p, S) { c.cbase[p] = -1; c.ubase[p] = 1; }
    QT_SYNC();
    QT_FOR(k, K) { int p = c.ord[k]; c.cbase[p] = c.cpre[k]; c.ubase[p] = 0; }
    QT_SYNC();
    const int U = qt_scan_excl(c.ubase, S, c.scratch);
    const int newS = M + U;

    // build the new table
    QT_FOR(p, S) {
        int ulx = box[4 * p], uly = box[4 * p + 1], brx = box[4 * p + 2], bry = box[4 * p + 3];
        if (c.cbase[p] >= 0) {
            int halfX = (brx - ulx + 1) >> 1, halfY = (bry - uly + 1) >> 1;
            int mx = ulx + halfX, my = uly + halfY;
            int cidx = c.cbase[p];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                int k = c.childcnt[4 * p + q];
                if (k > 0) {
                    int np = M - 1 - cidx;
                    cidx++;
                    nbox[4 * np] = (int16_t)((q & 1) ? mx : ulx);
                    nbox[4 * np + 1] = (int16_t)((q & 2) ? my : uly);
                    nbox[4 * np + 2] = (int16_t)((q & 1) ? brx : mx);
                    nbox[4 * np + 3] = (int16_t)((q & 2) ? bry : my);
                    ncnt[np] = (uint16_t)k;
                    c.childcnt[4 * p + q] = np;
                    if (k > 1) qt_atomic_add(&c.scratch[QT_S_NEXP], 1);
                }
            }
        } else {
            int np = M + c.ubase[p];
            nbox[4 * np] = (int16_t)ulx; nbox[4 * np + 1] = (int16_t)uly;
            nbox[4 * np + 2] = (int16_t)brx; nbox[4 * np + 3] = (int16_t)bry;
            ncnt[np] = cnt[p];
            c.ubase[p] = np;
        }
    }
    QT_SYNC();
    QT_FOR(i, c.n) {
        int p = c.cnode[i];
        c.cnode[i] = (uint16_t)(c.cbase[p] >= 0 ? c.childcnt[4 * p + c.cq[i]] : c.ubase[p]);
    }
    QT_SYNC();
    *Mout = M;
    *nexp = c.scratch[QT_S_NEXP];
    QT_SYNC();
    return newS;
}

// ---- whole distribution.  out[0..ret) = selected candidate words in list order -------------------
// `out_cap` entries are available in `out`; the return value is the list size (<= ncap).
static QT_DEV int qt_distribute(QtCtx &c, uint32_t *out, int out_cap) {
    // roots (:590-617): nIni vertical strips
    const int nIni = (int)roundf((float)c.width / (float)c.height);
    if (nIni < 1 || c.n == 0) return 0;
    const float hX = (float)c.width / (float)nIni;
    int cur = 0;
    QT_FOR(p, nIni) {
        c.box[0][4 * p] = (int16_t)(int)(hX * (float)p);
        c.box[0][4 * p + 1] = 0;
        c.box[0][4 * p + 2] = (int16_t)(int)(hX * (float)(p + 1));
        c.box[0][4 * p + 3] = (int16_t)c.height;
        c.childcnt[p] = 0;
    }
    QT_SYNC();
    QT_FOR(i, c.n) {
        int x = c.cand[i] & 0xfff;
        int r = (int)((float)x / hX);
        if (r >= nIni) r = nIni - 1;
        c.cnode[i] = (uint16_t)r;
        qt_atomic_add(&c.childcnt[r], 1);
    }
    QT_SYNC();
    // drop empty roots (:621-632), keep order
    QT_FOR(p, nIni) c.ubase[p] = c.childcnt[p] > 0 ? 1 : 0;
    QT_SYNC();
    int S = qt_scan_excl(c.ubase, nIni, c.scratch);
    QT_FOR(p, nIni) {
        if (c.childcnt[p] > 0) {
            int np = c.ubase[p];
            for (int t = 0; t < 4; t++) c.box[1][4 * np + t] = c.box[0][4 * p + t];
            c.cnt[1][np] = (uint16_t)c.childcnt[p];
        }
    }
    QT_SYNC();
    QT_FOR(i, c.n) c.cnode[i] = (uint16_t)c.ubase[c.cnode[i]];
    QT_SYNC();
    cur = 1;

    bool finish = false;
    while (!finish) {  // :641
        int prevSize = S, M, nexp;
        S = qt_pass(c, cur, S, 0, 0, &M, &nexp);
        cur ^= 1;
        if (S >= c.N || S == prevSize) {
            finish = true;
        } else if (S + nexp * 3 > c.N) {  // :719
            while (!finish) {
                prevSize = S;
                int Mprev = M;
                S = qt_pass(c, cur, S, 1, Mprev, &M, &nexp);
                cur ^= 1;
                if (S >= c.N || S == prevSize) finish = true;
            }
        }
    }

    // best response per node, ties -> first in the reference's candidate order (:789-807)
    unsigned long long *best = (unsigned long long *)c.childcnt;  // ncap*4 ints >= ncap u64
    QT_FOR(p, S) best[p] = 0ull;
    QT_SYNC();
    QT_FOR(i, c.n) {
        uint32_t w = c.cand[i];
        int x = w & 0xfff, y = (w >> 12) & 0xfff, r = w >> 24;
        int ci = (y - 3) / c.hCell, cj = (x - 3) / c.wCell;
        uint32_t ordk = ((uint32_t)(ci * c.nCols + cj) << 12) | ((uint32_t)(y - 3 - ci * c.hCell) << 6) |
                        (uint32_t)(x - 3 - cj * c.wCell);
        qt_atomic_max64(&best[c.cnode[i]], ((unsigned long long)(r + 1) << 32) | (unsigned long long)(0xffffffffu - ordk));
    }
    QT_SYNC();
    QT_FOR(p, S) {
        if (p < out_cap) {
            unsigned long long b = best[p];
            uint32_t r = (uint32_t)(b >> 32) - 1u;
            uint32_t ordk = 0xffffffffu - (uint32_t)(b & 0xffffffffu);
            int cell = ordk >> 12, yin = (ordk >> 6) & 63, xin = ordk & 63;
            int ci = cell / c.nCols, cj = cell - ci * c.nCols;
            uint32_t x = (uint32_t)(cj * c.wCell + xin + 3), y = (uint32_t)(ci * c.hCell + yin + 3);
            out[p] = x | (y << 12) | (r << 24);
        }
    }
    QT_SYNC();
    return S;
}

#endif  // QT_CORE_INL
