// match.cu — brute-force Hamming 1-NN on 256-bit descriptors (sm_100a).
//
// Replaces cv::BFMatcher(NORM_HAMMING)::match as LoopClosing::MatchFeatures uses it (reference
// src/loopclosing.cpp:33,172): for every query row the nearest train row, ties -> lowest trainIdx.
// HBM traffic is negligible (64 KB per descriptor set); the work is 4 M distance evaluations per
// 2000 x 2000 problem, done as an exact int8 tensor-core GEMM (see below).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "umma.cuh"


struct sb_matcher {
    int device, max_batch, max_rows;
    cudaStream_t stream, own_stream;
    uint8_t *d_q, *d_t;
    int32_t *d_nq, *d_nt, *d_idx, *d_dist;
    // tensor-core operands: 0/1-byte expansions of the two descriptor sets, train keys, query popcounts
    int rows_pad;
    uint8_t *d_xq, *d_xt;
    uint32_t *d_tkey;
    int32_t *d_pq;
    CUtensorMap map_q, map_t;  // expanded operands as [rows][256] u8, 128-byte swizzle; boxes 128 x UM_M / 128 x UM_N
};

// ---------------------------------------------------------------------------------------------------
// Hamming distance on the tensor cores.  For bit vectors q, t:  |q xor t| = |q| + |t| - 2 <q, t>, and the
// 2000 x 2000 x 256 table of dot products is GEMM-shaped, so the bits are widened to 0/1 bytes once
// (k_expand) and the dot products come from the 5th-generation tensor cores (tcgen05.mma kind::i8,
// u8 x u8 -> s32: exact integer arithmetic).  Against the xor + POPC formulation this removes the POPC
// pipe (16 lanes/clk/SM) as the limit; against the mma.sync (IMMA m16n8k32) version this library used
// first it is 2.7x faster (199 -> 75 us for 64 problems of 2000 x 2000, profiles/).
// 1-NN: per query row minimise  |t| - 2 <q,t>  packed with the train index into one word
//   key = (|t| + 512 - 2 <q,t>) << 22 | index      (10 + 22 bits)
// so one IMAD + one unsigned min per accumulator keeps BFMatcher's "lowest trainIdx wins ties" rule,
// and partial results of train slices merge with atomicMin.
// ---------------------------------------------------------------------------------------------------
#define MATCH_KEY_SHIFT 22

// Expanded operand rows: xq / xt [set][rows_pad][256] bytes (0/1), rows >= n are zero;
// tkey[set][rows_pad] = (|t| + 512) << 22 | index  (0xffffffff for rows >= nt);  pq[set][rows_pad] = |q|.
struct ExpandSide {
    const uint8_t *src;
    long long set_stride;
    const int32_t *n_arr;
    int n_stride;
    uint8_t *dst;
    uint32_t *tkey;      // train side only
    int32_t *pq;         // query side only
    uint32_t *init_key;  // query side only: the atomicMin target of the query row
    long long init_stride;
};
// blockIdx.z = 0: query sets, 1: train sets (one launch for both)
__global__ void __launch_bounds__(256) k_expand(ExpandSide sq, ExpandSide st, int max_rows, int rows_pad) {
    sb_pdl_enter();
    const ExpandSide &S = blockIdx.z ? st : sq;
    const int set = blockIdx.y;
    const int n = min(S.n_arr[(long long)set * S.n_stride], max_rows);
    const int row = blockIdx.x * 32 + (threadIdx.x >> 3), part = threadIdx.x & 7;  // 8 threads per row, one 32-bit word each
    if (row >= rows_pad) return;
    uint32_t w = 0;
    if (row < n) w = reinterpret_cast<const uint32_t *>(S.src + (long long)set * S.set_stride + (long long)row * 32)[part];
    uint4 lo, hi;  // bits 0..15 and 16..31 of the word as 0/1 bytes
    lo.x = ((w >> 0) & 0xfu) * 0x00204081u & 0x01010101u;  lo.y = ((w >> 4) & 0xfu) * 0x00204081u & 0x01010101u;
    lo.z = ((w >> 8) & 0xfu) * 0x00204081u & 0x01010101u;  lo.w = ((w >> 12) & 0xfu) * 0x00204081u & 0x01010101u;
    hi.x = ((w >> 16) & 0xfu) * 0x00204081u & 0x01010101u; hi.y = ((w >> 20) & 0xfu) * 0x00204081u & 0x01010101u;
    hi.z = ((w >> 24) & 0xfu) * 0x00204081u & 0x01010101u; hi.w = ((w >> 28) & 0xfu) * 0x00204081u & 0x01010101u;
    uint4 *d = reinterpret_cast<uint4 *>(S.dst + ((long long)set * rows_pad + row) * 256 + part * 32);
    d[0] = lo;
    d[1] = hi;
    int pc = __popc(w);
    pc += __shfl_xor_sync(0xffffffffu, pc, 1);
    pc += __shfl_xor_sync(0xffffffffu, pc, 2);
    pc += __shfl_xor_sync(0xffffffffu, pc, 4);
    if (part == 0) {
        if (S.tkey) S.tkey[(long long)set * rows_pad + row] = row < n ? (((uint32_t)pc + 512u) << MATCH_KEY_SHIFT) | (uint32_t)row : 0xffffffffu;
        if (S.pq) S.pq[(long long)set * rows_pad + row] = pc;
        if (S.init_key && row < n) S.init_key[(long long)set * S.init_stride + row] = 0xffffffffu;
    }
}

// key -> (train index, distance = |q| + (key >> 22) - 512)
__global__ void k_hamming_decode(const int32_t *__restrict__ nq_arr, int nq_stride, int max_rows, int rows_pad,
                                 const int32_t *__restrict__ pq, int32_t *idx_key, int32_t *out_dist, long long out_stride) {
    sb_pdl_enter();
    const int set = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= min(nq_arr[(long long)set * nq_stride], max_rows)) return;
    const uint32_t key = (uint32_t)idx_key[(long long)set * out_stride + qi];
    const bool none = key == 0xffffffffu;
    idx_key[(long long)set * out_stride + qi] = none ? -1 : (int32_t)(key & ((1u << MATCH_KEY_SHIFT) - 1u));
    out_dist[(long long)set * out_stride + qi] = none ? -1 : (int32_t)(key >> MATCH_KEY_SHIFT) - 512 + pq[(long long)set * rows_pad + qi];
}

// ---------------------------------------------------------------------------------------------------
// The same GEMM on the 5th-generation tensor cores: tcgen05.mma kind::i8 (u8 x u8 -> s32), operands staged by TMA
// into 128-byte-swizzled shared memory, accumulators in tensor memory, 1-NN reduction fused into the epilogue.
//   CTA  = 128 query rows (the M of one UMMA) x one slice of the train set, walked in tiles of 256 train rows (N);
//   K    = 256 expanded bytes = 8 instructions of K = 32; an operand tile is two 128-byte-wide swizzle panels;
//   smem = the query tile (32 KB) + a 2-deep ring of train tiles (64 KB each) filled by TMA ahead of the MMAs;
//   TMEM = 2 accumulators of 128 lanes x 256 columns: the MMAs of tile t + 1 overlap the epilogue of tile t
//          (measured: 128-row tiles with a 5-deep ring are slower — the epilogue, not the operand feed, is the limit);
//   warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one lane), warps 2-17 = epilogue (lane quadrant = warp % 4,
//          column quarter = (warp - 2) / 4): a thread reads 64 columns of its accumulator row with tcgen05.ld (two loads
//          in flight) and folds key - (dot << 23) into a min; the tile's 256 train keys are staged in shared memory.
//          16 epilogue warps: the epilogue is ~1.75 instructions per (query, train) pair and must hide the TMEM latency.
// ---------------------------------------------------------------------------------------------------
#define UM_M 128
#define UM_N 256
#define UM_BST 2   // train tiles in flight in shared memory
#define UM_AST 2   // accumulators in tensor memory (UM_AST * UM_N = 512 columns)
#define UM_THREADS 576
#define UM_EPI_THREADS 512
#define UM_COLS_PER_WARP (UM_N / (UM_EPI_THREADS / 128))   // 64
#define UM_A_BYTES (UM_M * 256)
#define UM_B_BYTES (UM_N * 256)
#define UM_SMEM (UM_A_BYTES + UM_BST * UM_B_BYTES + 1024)

static __device__ __forceinline__ uint32_t um_fold32(uint32_t best, const uint32_t (&v)[32], const uint32_t *tk) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const uint4 k4 = *reinterpret_cast<const uint4 *>(tk + 4 * j);  // rows past the train set carry 0xffffffff and a zero dot product
        // key - (dot << 23) as one IMAD, two 3-input minima (VIMNMX3) per four columns
        const uint32_t NEG = 0u - (1u << 23);
        best = __vimin3_u32(best, v[4 * j] * NEG + k4.x, v[4 * j + 1] * NEG + k4.y);
        best = __vimin3_u32(best, v[4 * j + 2] * NEG + k4.z, v[4 * j + 3] * NEG + k4.w);
    }
    return best;
}

// grid = (rows_pad / UM_M, nslices, batch); out_key pre-set to 0xffffffff; rows_pad is a multiple of UM_N.
__global__ void __launch_bounds__(UM_THREADS, 1) k_hamming_umma(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_t,
                                                               const uint32_t *__restrict__ tkey, const int32_t *__restrict__ nq_arr,
                                                               int nq_stride, const int32_t *__restrict__ nt_arr, int nt_stride,
                                                               int max_rows, int rows_pad, uint32_t *__restrict__ out_key,
                                                               long long out_stride, int nslices) {
    sb_pdl_enter();
    extern __shared__ uint8_t um_raw[];
    __shared__ __align__(8) uint64_t a_full, b_full[UM_BST], b_empty[UM_BST], acc_full[UM_AST], acc_empty[UM_AST];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) uint32_t s_tk[UM_AST][UM_N];
    const int set = blockIdx.z;
    const int nq = min(nq_arr[(long long)set * nq_stride], max_rows);
    const int nt = min(nt_arr[(long long)set * nt_stride], max_rows);
    const int q0 = blockIdx.x * UM_M;
    if (q0 >= nq) return;
    const int per = (((nt + nslices - 1) / nslices) + UM_N - 1) / UM_N * UM_N;  // train rows per slice
    const int t_begin = blockIdx.y * per, t_end = min(nt, t_begin + per);
    if (t_begin >= t_end) return;
    const int ntiles = (t_end - t_begin + UM_N - 1) / UM_N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *smem = um_raw + ((1024u - (sb_smem_u32(um_raw) & 1023u)) & 1023u);  // swizzle panels need 1024-byte alignment
    uint8_t *sA = smem, *sB = smem + UM_A_BYTES;

    if (threadIdx.x == 0) {
        sb_mbar_init(&a_full, 1);
        for (int i = 0; i < UM_BST; i++) {
            sb_mbar_init(&b_full[i], 1);
            sb_mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < UM_AST; i++) {
            sb_mbar_init(&acc_full[i], 1);
            sb_mbar_init(&acc_empty[i], UM_EPI_THREADS / 32);
        }
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb_smem_u32(&tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer
            const int rowq = set * rows_pad + q0, rowt = set * rows_pad + t_begin;
            sb_mbar_expect_tx(&a_full, UM_A_BYTES);
            um_tma_load_2d(sA, &map_q, 0, rowq, &a_full);
            um_tma_load_2d(sA + UM_A_BYTES / 2, &map_q, 128, rowq, &a_full);
            for (int t = 0; t < ntiles; t++) {
                const int s = t % UM_BST, u = t / UM_BST;
                if (u >= 1) sb_mbar_wait(&b_empty[s], (u - 1) & 1);
                sb_mbar_expect_tx(&b_full[s], UM_B_BYTES);
                um_tma_load_2d(sB + s * UM_B_BYTES, &map_t, 0, rowt + t * UM_N, &b_full[s]);
                um_tma_load_2d(sB + s * UM_B_BYTES + UM_B_BYTES / 2, &map_t, 128, rowt + t * UM_N, &b_full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer
            // instruction descriptor: D = s32 (2 << 4), A = B = u8 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (2u << 4) | ((uint32_t)(UM_N >> 3) << 17) | ((uint32_t)(UM_M >> 4) << 24);
            sb_mbar_wait(&a_full, 0);
            for (int t = 0; t < ntiles; t++) {
                const int s = t % UM_BST, a = t % UM_AST, ua = t / UM_AST;
                sb_mbar_wait(&b_full[s], (t / UM_BST) & 1);
                if (ua >= 1) sb_mbar_wait(&acc_empty[a], (ua - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 8; k++) {
                    const uint64_t ad = um_smem_desc(sA + (k >> 2) * (UM_A_BYTES / 2) + (k & 3) * 32);
                    const uint64_t bd = um_smem_desc(sB + s * UM_B_BYTES + (k >> 2) * (UM_B_BYTES / 2) + (k & 3) * 32);
                    um_mma_i8(tmem + a * UM_N, ad, bd, idesc, k > 0);
                }
                um_commit(&b_empty[s]);
                um_commit(&acc_full[a]);
            }
        }
    } else {  // ===== epilogue: two threads per query row (128 columns each)
        const int quad = warp & 3, part = (warp - 2) >> 2;
        const int row = q0 + quad * 32 + lane;
        const int et = threadIdx.x - 64;  // 0 .. UM_EPI_THREADS - 1
        const uint32_t *TK = tkey + (long long)set * rows_pad + t_begin;
        uint32_t best = 0xffffffffu;
        uint32_t tk_next = et < UM_N ? __ldg(TK + et) : 0u;  // prefetched one tile ahead (rows_pad covers the last tile)
        for (int t = 0; t < ntiles; t++) {
            const int a = t % UM_AST;
            // the tile's train keys: s_tk[a] was last read for tile t - UM_AST, and every epilogue thread has passed the
            // named barriers of the tiles in between since then
            if (et < UM_N) {
                s_tk[a][et] = tk_next;
                if (t + 1 < ntiles) tk_next = __ldg(TK + (t + 1) * UM_N + et);
            }
            asm volatile("bar.sync 1, %0;" ::"n"(UM_EPI_THREADS) : "memory");
            sb_mbar_wait(&acc_full[a], (t / UM_AST) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tbase = tmem + (((uint32_t)quad * 32u) << 16) + (uint32_t)(a * UM_N + part * UM_COLS_PER_WARP);
            const uint32_t *tk = s_tk[a] + part * UM_COLS_PER_WARP;
            uint32_t va[32], vb[32];
            um_tmem_ld32_issue(tbase, va);
            um_tmem_ld32_issue(tbase + 32, vb);
            um_tmem_ld_wait(va);
            um_tmem_ld_wait(vb);
            best = um_fold32(best, va, tk);
            best = um_fold32(best, vb, tk + 32);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) um_mbar_arrive(&acc_empty[a]);
        }
        if (row < nq && best != 0xffffffffu) atomicMin(&out_key[(long long)set * out_stride + row], best);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

static void free_matcher(sb_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    void *ptrs[] = {m->d_q, m->d_t, m->d_nq, m->d_nt, m->d_idx, m->d_dist, m->d_xq, m->d_xt, m->d_tkey, m->d_pq};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

extern "C" int sb_matcher_create(sb_matcher_t **out, int device, int max_batch, int max_rows) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_batch >= 1 && max_batch <= 65535, "max_batch out of range [1, 65535]");
    SB_REQUIRE(max_rows >= 1 && max_rows <= (1 << 20), "max_rows out of range [1, 2^20]");
    SB_TRY(sb_use_device(device));
    sb_matcher *m = new sb_matcher();
    memset(m, 0, sizeof(*m));
    m->device = device;
    m->max_batch = max_batch;
    m->max_rows = max_rows;
    const size_t rows = (size_t)max_batch * max_rows;
    cudaError_t e = cudaMalloc((void **)&m->d_q, rows * 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_t, rows * 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_nq, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_nt, (size_t)max_batch * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_idx, rows * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_dist, rows * 4);
    m->rows_pad = (int)sb_align_up((size_t)max_rows, 256);  // UM_N: train tiles never cross a set
    const size_t prow = (size_t)max_batch * m->rows_pad;
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_xq, prow * 256);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_xt, prow * 256);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_tkey, prow * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_pq, prow * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hamming_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, UM_SMEM);
    if (e != cudaSuccess) {
        sb_set_error("sb_matcher_create: %s", cudaGetErrorString(e));
        free_matcher(m);
        return SB_ERR_CUDA;
    }
    {
        const uint64_t dims[2] = {256, (uint64_t)prow}, strides[1] = {256};
        const uint32_t boxq[2] = {128, UM_M}, boxt[2] = {128, UM_N};
        int rc = sb_make_tensor_map_u8_sw128(&m->map_q, m->d_xq, 2, dims, strides, boxq);
        if (rc == SB_OK) rc = sb_make_tensor_map_u8_sw128(&m->map_t, m->d_xt, 2, dims, strides, boxt);
        if (rc != SB_OK) {
            free_matcher(m);
            return rc;
        }
    }
    m->stream = m->own_stream;
    *out = m;
    return SB_OK;
}

extern "C" int sb_matcher_destroy(sb_matcher_t *m) {
    SB_NVTX_FN();
    if (m) {
        cudaSetDevice(m->device);
        cudaDeviceSynchronize();
        free_matcher(m);
    }
    return SB_OK;
}

extern "C" int sb_matcher_set_stream(sb_matcher_t *m, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(m, "null handle");
    m->stream = stream ? (cudaStream_t)stream : m->own_stream;
    return SB_OK;
}

extern "C" int sb_hamming_match_dev(sb_matcher_t *m, int batch, const uint8_t *d_q, int64_t q_set_stride,
                                    const int32_t *d_nq, int nq_stride, const uint8_t *d_t, int64_t t_set_stride,
                                    const int32_t *d_nt, int nt_stride, int max_rows, int32_t *d_train_idx,
                                    int32_t *d_dist, int64_t out_stride) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(m, "null handle");
    SB_REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    SB_REQUIRE(max_rows >= 1 && max_rows <= (1 << 20), "max_rows out of range [1, 2^20]");
    SB_REQUIRE(d_q && d_t && d_nq && d_nt && d_train_idx && d_dist, "null device pointer");
    SB_REQUIRE(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0 && (q_set_stride & 15) == 0 && (t_set_stride & 15) == 0,
               "descriptor sets must be 16-byte aligned");
    SB_TRY(sb_use_device(m->device));
    SB_REQUIRE(batch <= m->max_batch && max_rows <= m->max_rows, "batch / max_rows larger than given at create time");
    const int rp = m->rows_pad;
    const ExpandSide sq = {d_q, q_set_stride, d_nq, nq_stride, m->d_xq, nullptr, m->d_pq, reinterpret_cast<uint32_t *>(d_train_idx), out_stride};
    const ExpandSide st = {d_t, t_set_stride, d_nt, nt_stride, m->d_xt, m->d_tkey, nullptr, nullptr, 0};
    SB_CUDA(sb_launch_pdl(k_expand, dim3(sb_div_up(rp, 32), batch, 2), dim3(256), 0, m->stream, sq, st, max_rows, rp));
    // enough CTAs to fill 148 SMs: slice the train set when the batch alone does not
    const int qblocks = sb_div_up(max_rows, UM_M);
    int nslices = 1;  // one CTA per SM (160 KB of shared memory, all of TMEM)
    while (nslices < 16 && (long long)qblocks * batch * nslices < 148 && max_rows / (nslices * 2) >= UM_N) nslices *= 2;
    SB_CUDA(sb_launch_pdl(k_hamming_umma, dim3(qblocks, nslices, batch), dim3(UM_THREADS), UM_SMEM, m->stream, m->map_q, m->map_t, m->d_tkey, d_nq,
                          nq_stride, d_nt, nt_stride, max_rows, rp, reinterpret_cast<uint32_t *>(d_train_idx), out_stride, nslices));
    SB_CUDA(sb_launch_pdl(k_hamming_decode, dim3(sb_div_up(max_rows, 256), batch), dim3(256), 0, m->stream, d_nq, nq_stride, max_rows, rp, m->d_pq,
                          d_train_idx, d_dist, out_stride));
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_hamming_match(sb_matcher_t *m, int batch, const uint8_t *q, const int32_t *nq, const uint8_t *t,
                                const int32_t *nt, int cap, int32_t *train_idx, int32_t *dist) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(m, "null handle");
    SB_REQUIRE(batch >= 1 && batch <= m->max_batch, "batch out of range [1, max_batch]");
    SB_REQUIRE(cap >= 1 && cap <= m->max_rows, "cap out of range [1, max_rows]");
    SB_REQUIRE(q && nq && t && nt && train_idx && dist, "null pointer");
    for (int b = 0; b < batch; b++) SB_REQUIRE(nq[b] >= 0 && nq[b] <= cap && nt[b] >= 0 && nt[b] <= cap, "row count out of range [0, cap]");
    SB_TRY(sb_use_device(m->device));
    const size_t rows = (size_t)batch * cap;
    cudaStream_t s = m->stream;
    SB_CUDA(cudaMemcpyAsync(m->d_q, q, rows * 32, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(m->d_t, t, rows * 32, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(m->d_nq, nq, (size_t)batch * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemcpyAsync(m->d_nt, nt, (size_t)batch * 4, cudaMemcpyHostToDevice, s));
    SB_CUDA(cudaMemsetAsync(m->d_idx, 0xff, rows * 4, s));
    SB_CUDA(cudaMemsetAsync(m->d_dist, 0xff, rows * 4, s));
    SB_TRY(sb_hamming_match_dev(m, batch, m->d_q, (int64_t)cap * 32, m->d_nq, 1, m->d_t, (int64_t)cap * 32, m->d_nt, 1, cap,
                                m->d_idx, m->d_dist, cap));
    SB_CUDA(cudaMemcpyAsync(train_idx, m->d_idx, rows * 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaMemcpyAsync(dist, m->d_dist, rows * 4, cudaMemcpyDeviceToHost, s));
    SB_CUDA(cudaStreamSynchronize(s));
    return SB_OK;
}
