// match.cu — brute-force Hamming 1-NN on 256-bit descriptors (sm_100a).
//
// Replaces cv::BFMatcher(NORM_HAMMING)::match as LoopClosing::MatchFeatures uses it (reference
// src/loopclosing.cpp:33,172): for every query row the nearest train row, ties -> lowest trainIdx.
// The work is xor + population count; there is nothing to put on tensor cores and HBM traffic is
// negligible (64 KB per descriptor set), so the kernel is organised around the POPC issue rate:
// each thread owns QPT query rows in registers, the train set streams through shared memory in
// chunks and every shared-memory read is a warp-wide broadcast.
#include <string.h>

#include "common.cuh"

#define MATCH_THREADS 128
#define MATCH_QPT 2                              // query rows per thread
#define MATCH_QB (MATCH_THREADS * MATCH_QPT)     // query rows per CTA
#define MATCH_CHUNK 256                          // train rows per shared-memory chunk

struct sb_matcher {
    int device, max_batch, max_rows;
    cudaStream_t stream, own_stream;
    uint8_t *d_q, *d_t;
    int32_t *d_nq, *d_nt, *d_idx, *d_dist;
};

// Partial 1-NN over one slice of the train set.  The result of a slice is folded into out_idx[] with
// atomicMin on the key (distance << 22 | train index): the minimum key is the smallest distance and,
// among equal distances, the lowest train index — exactly BFMatcher's tie rule, independent of the
// order in which the slices finish.  k_hamming_decode then splits the key into (index, distance).
#define MATCH_KEY_SHIFT 22
__global__ void __launch_bounds__(MATCH_THREADS) k_hamming(const uint8_t *__restrict__ q, long long q_set_stride,
                                                          const int32_t *__restrict__ nq_arr, int nq_stride,
                                                          const uint8_t *__restrict__ t, long long t_set_stride,
                                                          const int32_t *__restrict__ nt_arr, int nt_stride, int max_rows,
                                                          uint32_t *__restrict__ out_key, long long out_stride, int nslices) {
    __shared__ uint4 ts[MATCH_CHUNK * 2];
    const int set = blockIdx.z;
    const int nq = min(nq_arr[(long long)set * nq_stride], max_rows);
    const int nt = min(nt_arr[(long long)set * nt_stride], max_rows);
    const int q0 = blockIdx.x * MATCH_QB;
    if (q0 >= nq) return;
    const int per = (((nt + nslices - 1) / nslices) + MATCH_CHUNK - 1) / MATCH_CHUNK * MATCH_CHUNK;  // rows per slice
    const int t_begin = blockIdx.y * per, t_end = min(nt, t_begin + per);
    if (t_begin >= t_end) return;
    const uint4 *Q = reinterpret_cast<const uint4 *>(q + (long long)set * q_set_stride);
    const uint4 *T = reinterpret_cast<const uint4 *>(t + (long long)set * t_set_stride);

    uint4 qa[MATCH_QPT], qb[MATCH_QPT];
    int best[MATCH_QPT], bidx[MATCH_QPT];
#pragma unroll
    for (int k = 0; k < MATCH_QPT; k++) {
        const int qi = q0 + k * MATCH_THREADS + threadIdx.x;
        const int qs = qi < nq ? qi : q0;  // idle lanes shadow a valid row; their result is not stored
        qa[k] = Q[2 * qs];
        qb[k] = Q[2 * qs + 1];
        best[k] = 0x7fffffff;
        bidx[k] = -1;
    }
    for (int c0 = t_begin; c0 < t_end; c0 += MATCH_CHUNK) {
        const int cn = min(MATCH_CHUNK, t_end - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cn * 2; i += MATCH_THREADS) ts[i] = T[2 * c0 + i];
        __syncthreads();
#pragma unroll 4
        for (int j = 0; j < cn; j++) {
            const uint4 ta = ts[2 * j], tb = ts[2 * j + 1];
#pragma unroll
            for (int k = 0; k < MATCH_QPT; k++) {
                const int d = __popc(qa[k].x ^ ta.x) + __popc(qa[k].y ^ ta.y) + __popc(qa[k].z ^ ta.z) +
                              __popc(qa[k].w ^ ta.w) + __popc(qb[k].x ^ tb.x) + __popc(qb[k].y ^ tb.y) +
                              __popc(qb[k].z ^ tb.z) + __popc(qb[k].w ^ tb.w);
                if (d < best[k]) {  // strict: the first (lowest) train index wins ties
                    best[k] = d;
                    bidx[k] = c0 + j;
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < MATCH_QPT; k++) {
        const int qi = q0 + k * MATCH_THREADS + threadIdx.x;
        if (qi < nq && bidx[k] >= 0)
            atomicMin(&out_key[(long long)set * out_stride + qi], ((uint32_t)best[k] << MATCH_KEY_SHIFT) | (uint32_t)bidx[k]);
    }
}

__global__ void k_hamming_init(const int32_t *__restrict__ nq_arr, int nq_stride, int max_rows, uint32_t *out_key,
                               long long out_stride) {
    const int set = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi < min(nq_arr[(long long)set * nq_stride], max_rows)) out_key[(long long)set * out_stride + qi] = 0xffffffffu;
}

__global__ void k_hamming_decode(const int32_t *__restrict__ nq_arr, int nq_stride, int max_rows, int32_t *idx_key,
                                 int32_t *out_dist, long long out_stride) {
    const int set = blockIdx.y, qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= min(nq_arr[(long long)set * nq_stride], max_rows)) return;
    const uint32_t key = (uint32_t)idx_key[(long long)set * out_stride + qi];
    const bool none = key == 0xffffffffu;
    idx_key[(long long)set * out_stride + qi] = none ? -1 : (int32_t)(key & ((1u << MATCH_KEY_SHIFT) - 1u));
    out_dist[(long long)set * out_stride + qi] = none ? -1 : (int32_t)(key >> MATCH_KEY_SHIFT);
}

static void free_matcher(sb_matcher *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    void *ptrs[] = {m->d_q, m->d_t, m->d_nq, m->d_nt, m->d_idx, m->d_dist};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (m->own_stream) cudaStreamDestroy(m->own_stream);
    delete m;
}

extern "C" int sb_matcher_create(sb_matcher_t **out, int device, int max_batch, int max_rows) {
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
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        sb_set_error("sb_matcher_create: %s", cudaGetErrorString(e));
        free_matcher(m);
        return SB_ERR_CUDA;
    }
    m->stream = m->own_stream;
    *out = m;
    return SB_OK;
}

extern "C" int sb_matcher_destroy(sb_matcher_t *m) {
    if (m) {
        cudaSetDevice(m->device);
        cudaDeviceSynchronize();
        free_matcher(m);
    }
    return SB_OK;
}

extern "C" int sb_matcher_set_stream(sb_matcher_t *m, void *stream) {
    SB_REQUIRE(m, "null handle");
    m->stream = stream ? (cudaStream_t)stream : m->own_stream;
    return SB_OK;
}

extern "C" int sb_hamming_match_dev(sb_matcher_t *m, int batch, const uint8_t *d_q, int64_t q_set_stride,
                                    const int32_t *d_nq, int nq_stride, const uint8_t *d_t, int64_t t_set_stride,
                                    const int32_t *d_nt, int nt_stride, int max_rows, int32_t *d_train_idx,
                                    int32_t *d_dist, int64_t out_stride) {
    sb_clear_error();
    SB_REQUIRE(m, "null handle");
    SB_REQUIRE(batch >= 1 && batch <= 65535, "batch out of range");
    SB_REQUIRE(max_rows >= 1 && max_rows <= (1 << 20), "max_rows out of range [1, 2^20]");
    SB_REQUIRE(d_q && d_t && d_nq && d_nt && d_train_idx && d_dist, "null device pointer");
    SB_REQUIRE(((uintptr_t)d_q & 15) == 0 && ((uintptr_t)d_t & 15) == 0 && (q_set_stride & 15) == 0 && (t_set_stride & 15) == 0,
               "descriptor sets must be 16-byte aligned");
    SB_TRY(sb_use_device(m->device));
    // enough CTAs to fill 148 SMs: slice the train set when the batch alone does not
    const int qblocks = sb_div_up(max_rows, MATCH_QB);
    int nslices = 1;
    while (nslices < 8 && (long long)qblocks * batch * nslices < 148 * 12 && max_rows / (nslices * 2) >= MATCH_CHUNK) nslices *= 2;
    k_hamming_init<<<dim3(sb_div_up(max_rows, 256), batch), 256, 0, m->stream>>>(d_nq, nq_stride, max_rows,
                                                                                 reinterpret_cast<uint32_t *>(d_train_idx), out_stride);
    k_hamming<<<dim3(qblocks, nslices, batch), MATCH_THREADS, 0, m->stream>>>(d_q, q_set_stride, d_nq, nq_stride, d_t, t_set_stride,
                                                                              d_nt, nt_stride, max_rows,
                                                                              reinterpret_cast<uint32_t *>(d_train_idx), out_stride, nslices);
    k_hamming_decode<<<dim3(sb_div_up(max_rows, 256), batch), 256, 0, m->stream>>>(d_nq, nq_stride, max_rows, d_train_idx, d_dist,
                                                                                   out_stride);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

extern "C" int sb_hamming_match(sb_matcher_t *m, int batch, const uint8_t *q, const int32_t *nq, const uint8_t *t,
                                const int32_t *nt, int cap, int32_t *train_idx, int32_t *dist) {
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
