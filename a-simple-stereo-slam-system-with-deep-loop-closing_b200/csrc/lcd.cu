// lcd.cu — DeepLCD descriptor scoring over the keyframe database (sm_100a).
//
// Replaces DeepLCD::score (reference src/deeplcd.cpp:35-39: the dot product of two L2-normalised
// 1064-float descriptors) and the database scan of LoopClosing::DetectLoop (src/loopclosing.cpp:124-161).
// The database lives in HBM as [capacity][1088] fp16 (north_star: batched fp16 GEMV, fp32 accumulate;
// rows padded to 1088 = 34 x 32 halves so that every lane reads 16-byte vectors) or fp32
// ([capacity][1064], bit-faithful storage of what the reference keeps).  One warp scores one
// (query, row) pair; the kernel is a pure stream over the database: 2 128 B (fp16) per score.
#include <cuda_fp16.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "umma.cuh"

#define LCD_DIM 1064
#define LCD_PAD 1088
#define LCD_WARPS 8

struct sb_lcd {
    int device, capacity, dtype;  // dtype: 0 = fp32 rows, 1 = fp16 rows
    int n;
    cudaStream_t stream, own_stream;
    void *d_db;
    float *d_query, *d_scores;   // staging: [max_queries][1064], [max_queries][capacity]
    int max_queries;
    int64_t *h_ids;              // keyframe id of every row, ascending (std::map order of _mvDatabase)
    float *h_scores;             // pinned
    __half *d_qh;                // fp16 database only: the queries of a batch as fp16 [max_queries rounded up to 128][1088]
    CUtensorMap map_q, map_db;   // ... and both operands as [rows][2176 bytes], 128-byte swizzle (tensor-core batch path)
    float *d_stage;              // bulk-load staging, LCD_STAGE_ROWS x 1064 fp32, allocated by the first sb_lcd_add_batch
};
#define LCD_STAGE_ROWS 256

// scores[q][r] = <query q, row r>, fp32 accumulate.  grid = (ceil(n / LCD_WARPS), nq)
template <bool HALF>
__global__ void __launch_bounds__(LCD_WARPS * 32) k_lcd_score(const void *__restrict__ db, int n, const float *__restrict__ queries,
                                                             float *__restrict__ scores, int score_stride) {
    __shared__ __align__(16) float q[LCD_PAD];
    const float *qg = queries + (size_t)blockIdx.y * LCD_DIM;
    for (int i = threadIdx.x; i < LCD_PAD; i += blockDim.x) q[i] = i < LCD_DIM ? qg[i] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, r = blockIdx.x * LCD_WARPS + (threadIdx.x >> 5);
    if (r >= n) return;
    float acc = 0.f;
    if (HALF) {
        const uint4 *row = reinterpret_cast<const uint4 *>(reinterpret_cast<const __half *>(db) + (size_t)r * LCD_PAD);
        // 1088 halves = 136 uint4; lane handles uint4 lane, lane + 32, ... (4.25 per lane)
        for (int v = lane; v < LCD_PAD / 8; v += 32) {
            const uint4 u = row[v];
            const __half2 *h = reinterpret_cast<const __half2 *>(&u);
            const float *qq = q + v * 8;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const float2 f = __half22float2(h[k]);
                acc = fmaf(f.x, qq[2 * k], acc);
                acc = fmaf(f.y, qq[2 * k + 1], acc);
            }
        }
    } else {
        const float4 *row = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(db) + (size_t)r * LCD_DIM);
        for (int v = lane; v < LCD_DIM / 4; v += 32) {  // 1064 / 4 = 266 float4 per row, rows are 16-byte aligned (4256 B)
            const float4 u = row[v];
            const float *qq = q + v * 4;
            acc = fmaf(u.x, qq[0], acc);
            acc = fmaf(u.y, qq[1], acc);
            acc = fmaf(u.z, qq[2], acc);
            acc = fmaf(u.w, qq[3], acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) scores[(size_t)blockIdx.y * score_stride + r] = acc;
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch scoring on the 5th-generation tensor cores (fp16 database, >= LCD_UMMA_MIN_Q queries): scores = Q . DB^T is a
// GEMM, [nq x 1088] x [1088 x n], fp16 operands, fp32 accumulate — tcgen05.mma kind::f16.  Both operands are K-major
// rows of 2176 bytes, so a K chunk of 64 halves is one 128-byte swizzle span: TMA (SWIZZLE_128B) drops a 128-query and a
// 256-row panel per chunk into a 4-deep shared-memory ring, one elected thread issues 4 UMMAs (M 128, N 256, K 16) per
// chunk into a 128 x 256 fp32 accumulator in tensor memory, and four epilogue warps read it back with tcgen05.ld and
// store the scores.  CTA = one 128 x 256 output tile; warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (TMEM lane quadrant = warp % 4).  The queries are rounded to fp16 first (k_lcd_q2h): against the
// GEMV path (fp32 queries) a score moves by a few 1e-4, inside the fp16 database's own tolerance (tests/test_gpu_lcd.py).
// ---------------------------------------------------------------------------------------------------------------------
#define LCD_UMMA_MIN_Q 64
#define LG_M 128
#define LG_N 256
#define LG_KC (LCD_PAD / 64)          // 17 chunks of 64 halves
#define LG_ST 4                       // ring depth
#define LG_A_BYTES (LG_M * 128)
#define LG_B_BYTES (LG_N * 128)
#define LG_STAGE (LG_A_BYTES + LG_B_BYTES)
#define LG_SMEM (LG_ST * LG_STAGE + 1024)
#define LG_THREADS 192

// queries fp32 [nq][1064] -> fp16 [gridDim.x][1088], rows >= nq and the padding columns are zero
__global__ void k_lcd_q2h(const float *__restrict__ q, int nq, __half *__restrict__ out) {
    const int r = blockIdx.x;
    for (int i = threadIdx.x; i < LCD_PAD; i += blockDim.x)
        out[(size_t)r * LCD_PAD + i] = __float2half_rn(r < nq && i < LCD_DIM ? q[(size_t)r * LCD_DIM + i] : 0.f);
}

// grid = (ceil(nq / 128), ceil(n / 256))
__global__ void __launch_bounds__(LG_THREADS, 1) k_lcd_score_umma(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_db,
                                                                 int nq, int n, float *__restrict__ scores, int stride) {
    extern __shared__ uint8_t lg_raw[];
    __shared__ __align__(8) uint64_t full[LG_ST], empty[LG_ST], acc_full;
    __shared__ uint32_t tmem_slot;
    const int q0 = blockIdx.x * LG_M, r0 = blockIdx.y * LG_N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *smem = lg_raw + ((1024u - (sb_smem_u32(lg_raw) & 1023u)) & 1023u);  // swizzle panels need 1024-byte alignment
    if (threadIdx.x == 0) {
        for (int i = 0; i < LG_ST; i++) {
            sb_mbar_init(&full[i], 1);
            sb_mbar_init(&empty[i], 1);
        }
        sb_mbar_init(&acc_full, 1);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sb_smem_u32(&tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer
            for (int kc = 0; kc < LG_KC; kc++) {
                const int s = kc % LG_ST, u = kc / LG_ST;
                if (u >= 1) sb_mbar_wait(&empty[s], (u - 1) & 1);
                sb_mbar_expect_tx(&full[s], LG_STAGE);
                um_tma_load_2d(smem + s * LG_STAGE, &map_q, kc * 128, q0, &full[s]);
                um_tma_load_2d(smem + s * LG_STAGE + LG_A_BYTES, &map_db, kc * 128, r0, &full[s]);  // rows past the capacity: zero-filled
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer
            // instruction descriptor: D = f32 (1 << 4), A = B = f16 (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(LG_N >> 3) << 17) | ((uint32_t)(LG_M >> 4) << 24);
            for (int kc = 0; kc < LG_KC; kc++) {
                const int s = kc % LG_ST;
                sb_mbar_wait(&full[s], (kc / LG_ST) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint64_t ad = um_smem_desc(smem + s * LG_STAGE + k * 32);
                    const uint64_t bd = um_smem_desc(smem + s * LG_STAGE + LG_A_BYTES + k * 32);
                    um_mma_f16(tmem, ad, bd, idesc, (kc | k) != 0);
                }
                um_commit(&empty[s]);
            }
            um_commit(&acc_full);
        }
    } else {  // ===== epilogue: one thread per query row
        const int quad = warp & 3;
        const int row = q0 + quad * 32 + lane;
        sb_mbar_wait(&acc_full, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float *out = scores + (size_t)row * stride + r0;
#pragma unroll 1
        for (int c0 = 0; c0 < LG_N; c0 += 64) {
            uint32_t va[32], vb[32];
            const uint32_t tbase = tmem + (((uint32_t)quad * 32u) << 16) + (uint32_t)c0;
            um_tmem_ld32_issue(tbase, va);
            um_tmem_ld32_issue(tbase + 32, vb);
            um_tmem_ld_wait(va);
            um_tmem_ld_wait(vb);
            if (row < nq) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (r0 + c0 + j < n) out[c0 + j] = __uint_as_float(va[j]);
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (r0 + c0 + 32 + j < n) out[c0 + 32 + j] = __uint_as_float(vb[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// one CTA per row: database row (row0 + blockIdx.x) <- src[blockIdx.x][1064] (fp32 -> fp16 with zero padding, or fp32 copy)
__global__ void k_lcd_store_row(void *db, int row0, const float *src_rows, int half) {
    const int row = row0 + blockIdx.x;
    const float *src = src_rows + (size_t)blockIdx.x * LCD_DIM;
    for (int i = threadIdx.x; i < LCD_PAD; i += blockDim.x) {
        if (half)
            reinterpret_cast<__half *>(db)[(size_t)row * LCD_PAD + i] = __float2half_rn(i < LCD_DIM ? src[i] : 0.f);
        else if (i < LCD_DIM)
            reinterpret_cast<float *>(db)[(size_t)row * LCD_DIM + i] = src[i];
    }
}

static void free_lcd(sb_lcd *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_db) cudaFree(h->d_db);
    if (h->d_query) cudaFree(h->d_query);
    if (h->d_scores) cudaFree(h->d_scores);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_qh) cudaFree(h->d_qh);
    if (h->h_scores) cudaFreeHost(h->h_scores);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    free(h->h_ids);
    delete h;
}

extern "C" int sb_lcd_create(sb_lcd_t **out, int device, int capacity, int dtype, int max_queries) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(capacity >= 1 && capacity <= (1 << 22), "capacity out of range");
    SB_REQUIRE(dtype == SB_LCD_FP32 || dtype == SB_LCD_FP16, "dtype must be SB_LCD_FP32 or SB_LCD_FP16");
    SB_REQUIRE(max_queries >= 1 && max_queries <= 65535, "max_queries out of range [1, 65535]");
    SB_TRY(sb_use_device(device));
    sb_lcd *h = new sb_lcd();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->capacity = capacity;
    h->dtype = dtype;
    h->max_queries = max_queries;
    h->h_ids = (int64_t *)malloc(sizeof(int64_t) * (size_t)capacity);
    const size_t row = dtype == SB_LCD_FP16 ? LCD_PAD * 2 : LCD_DIM * 4;
    cudaError_t e = cudaMalloc(&h->d_db, row * capacity);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_query, (size_t)max_queries * LCD_DIM * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_scores, (size_t)max_queries * capacity * 4);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_scores, (size_t)capacity * 4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    const size_t q_rows = sb_align_up((size_t)max_queries, LG_M);
    if (e == cudaSuccess && dtype == SB_LCD_FP16 && max_queries >= LCD_UMMA_MIN_Q) {   // tensor-core batch path
        e = cudaMalloc((void **)&h->d_qh, q_rows * LCD_PAD * 2);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_lcd_score_umma, cudaFuncAttributeMaxDynamicSharedMemorySize, LG_SMEM);
    }
    if (e != cudaSuccess) {
        sb_set_error("sb_lcd_create: %s", cudaGetErrorString(e));
        free_lcd(h);
        return SB_ERR_CUDA;
    }
    if (h->d_qh) {
        const uint64_t strides[1] = {LCD_PAD * 2};
        const uint64_t dq[2] = {LCD_PAD * 2, (uint64_t)q_rows}, dd[2] = {LCD_PAD * 2, (uint64_t)capacity};
        const uint32_t bq[2] = {128, LG_M}, bd[2] = {128, LG_N};
        int rc = sb_make_tensor_map_u8_sw128(&h->map_q, h->d_qh, 2, dq, strides, bq);
        if (rc == SB_OK) rc = sb_make_tensor_map_u8_sw128(&h->map_db, h->d_db, 2, dd, strides, bd);
        if (rc != SB_OK) {
            free_lcd(h);
            return rc;
        }
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_lcd_destroy(sb_lcd_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_lcd(h);
    }
    return SB_OK;
}

extern "C" int sb_lcd_set_stream(sb_lcd_t *h, void *stream) {
    SB_NVTX_FN();
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_lcd_size(const sb_lcd_t *h) { return h ? h->n : SB_ERR_INVALID; }

// LoopClosing::AddToDatabase (src/loopclosing.cpp:651-659): _mvDatabase is a std::map keyed by KF id;
// ids must arrive in ascending order (they do: keyframes are processed in creation order).
extern "C" int sb_lcd_add(sb_lcd_t *h, int64_t kf_id, const float *descr) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && descr, "null pointer");
    SB_REQUIRE(h->n < h->capacity, "database full");
    SB_REQUIRE(h->n == 0 || kf_id > h->h_ids[h->n - 1], "keyframe ids must be added in ascending order");
    SB_TRY(sb_use_device(h->device));
    SB_CUDA(cudaMemcpyAsync(h->d_query, descr, LCD_DIM * 4, cudaMemcpyHostToDevice, h->stream));
    k_lcd_store_row<<<1, 256, 0, h->stream>>>(h->d_db, h->n, h->d_query, h->dtype == SB_LCD_FP16);
    SB_CUDA(cudaGetLastError());
    SB_CUDA(cudaStreamSynchronize(h->stream));  // descr may be reused by the caller
    h->h_ids[h->n++] = kf_id;
    return SB_OK;
}

// Bulk load (replay): rows [n][1064] fp32, ids ascending.  fp32 database: the rows are copied straight into place
// (same layout), no launch.  fp16 database: one host->device copy and one launch per chunk of LCD_STAGE_ROWS rows.
// Nothing is added when any argument is rejected.
extern "C" int sb_lcd_add_batch(sb_lcd_t *h, int n, const int64_t *kf_ids, const float *descr) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && (n == 0 || (kf_ids && descr)), "null pointer");
    SB_REQUIRE(n >= 0 && h->n + n <= h->capacity, "database full");
    for (int i = 0; i < n; i++)
        SB_REQUIRE(i ? kf_ids[i] > kf_ids[i - 1] : (h->n == 0 || kf_ids[0] > h->h_ids[h->n - 1]),
                   "keyframe ids must be added in ascending order");
    SB_TRY(sb_use_device(h->device));
    if (h->dtype == SB_LCD_FP32) {
        if (n) SB_CUDA(cudaMemcpyAsync((float *)h->d_db + (size_t)h->n * LCD_DIM, descr, (size_t)n * LCD_DIM * 4, cudaMemcpyHostToDevice, h->stream));
    } else {
        if (!h->d_stage && n) SB_CUDA(cudaMalloc((void **)&h->d_stage, (size_t)LCD_STAGE_ROWS * LCD_DIM * 4));
        for (int i0 = 0; i0 < n; i0 += LCD_STAGE_ROWS) {
            const int m = n - i0 < LCD_STAGE_ROWS ? n - i0 : LCD_STAGE_ROWS;
            SB_CUDA(cudaMemcpyAsync(h->d_stage, descr + (size_t)i0 * LCD_DIM, (size_t)m * LCD_DIM * 4, cudaMemcpyHostToDevice, h->stream));
            k_lcd_store_row<<<m, 256, 0, h->stream>>>(h->d_db, h->n + i0, h->d_stage, 1);
            SB_CUDA(cudaGetLastError());
        }
    }
    SB_CUDA(cudaStreamSynchronize(h->stream));  // descr may be reused by the caller
    for (int i = 0; i < n; i++) h->h_ids[h->n + i] = kf_ids[i];
    h->n += n;
    return SB_OK;
}

// _mvDatabase.erase(id) (src/loopclosing.cpp:73-75): the row is dropped, later rows move up.
extern "C" int sb_lcd_remove(sb_lcd_t *h, int64_t kf_id) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_TRY(sb_use_device(h->device));
    int r = -1;
    for (int i = 0; i < h->n; i++)
        if (h->h_ids[i] == kf_id) { r = i; break; }
    SB_REQUIRE(r >= 0, "keyframe id not in the database");
    const size_t row = h->dtype == SB_LCD_FP16 ? LCD_PAD * 2 : LCD_DIM * 4;
    SB_CUDA(cudaStreamSynchronize(h->stream));
    for (int i = r; i + 1 < h->n; i++) {  // overlapping device copy is not allowed: row by row
        SB_CUDA(cudaMemcpyAsync((char *)h->d_db + row * i, (char *)h->d_db + row * (i + 1), row, cudaMemcpyDeviceToDevice, h->stream));
        h->h_ids[i] = h->h_ids[i + 1];
    }
    SB_CUDA(cudaStreamSynchronize(h->stream));
    h->n--;
    return SB_OK;
}

static int launch_scores(sb_lcd *h, int nq, const float *d_queries, float *d_scores, int stride) {
    if (h->n == 0) return SB_OK;
    if (h->d_qh && nq >= LCD_UMMA_MIN_Q && nq <= h->max_queries) {   // a batch of queries against the fp16 database: tensor cores
        const int q_rows = (int)sb_align_up((size_t)nq, LG_M);
        k_lcd_q2h<<<q_rows, 256, 0, h->stream>>>(d_queries, nq, h->d_qh);
        k_lcd_score_umma<<<dim3(q_rows / LG_M, sb_div_up(h->n, LG_N)), LG_THREADS, LG_SMEM, h->stream>>>(h->map_q, h->map_db, nq, h->n, d_scores, stride);
        SB_CUDA(cudaGetLastError());
        return SB_OK;
    }
    dim3 grid(sb_div_up(h->n, LCD_WARPS), nq);
    if (h->dtype == SB_LCD_FP16)
        k_lcd_score<true><<<grid, LCD_WARPS * 32, 0, h->stream>>>(h->d_db, h->n, d_queries, d_scores, stride);
    else
        k_lcd_score<false><<<grid, LCD_WARPS * 32, 0, h->stream>>>(h->d_db, h->n, d_queries, d_scores, stride);
    SB_CUDA(cudaGetLastError());
    return SB_OK;
}

// scores[q][r] for nq queries against all rows; device pointers, asynchronous.  score_stride >= size.
extern "C" int sb_lcd_score_dev(sb_lcd_t *h, int nq, const float *d_queries, float *d_scores, int score_stride) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && d_queries && d_scores, "null pointer");
    SB_REQUIRE(nq >= 1 && nq <= 65535 && score_stride >= h->n, "bad nq / score_stride");
    SB_TRY(sb_use_device(h->device));
    return launch_scores(h, nq, d_queries, d_scores, score_stride);
}

// DeepLCD::score of `nq` query descriptors against every database row: scores [nq][size] (host).
extern "C" int sb_lcd_score(sb_lcd_t *h, int nq, const float *queries, float *scores) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && queries && scores, "null pointer");
    SB_REQUIRE(nq >= 1 && nq <= h->max_queries, "nq out of range [1, max_queries]");
    SB_TRY(sb_use_device(h->device));
    if (h->n == 0) return SB_OK;
    SB_CUDA(cudaMemcpyAsync(h->d_query, queries, (size_t)nq * LCD_DIM * 4, cudaMemcpyHostToDevice, h->stream));
    SB_TRY(launch_scores(h, nq, h->d_query, h->d_scores, h->n));
    SB_CUDA(cudaMemcpyAsync(scores, h->d_scores, (size_t)nq * h->n * 4, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    return SB_OK;
}

// LoopClosing::DetectLoop (src/loopclosing.cpp:124-161): rows are visited in ascending keyframe id, the
// scan STOPS at the first row with cur_kf_id - id < min_gap (quirk Q12: break, not continue; the
// subtraction is unsigned in the reference), the first maximum wins, rows scoring above thres_low are
// counted; a loop candidate is reported iff max >= thres_high and count <= max_suspected.
extern "C" int sb_lcd_detect_loop(sb_lcd_t *h, int64_t cur_kf_id, const float *query, float thres_high, float thres_low,
                                  int min_gap, int max_suspected, int *found, int64_t *best_id, float *max_score,
                                  int *n_suspected) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && query && found && best_id && max_score && n_suspected, "null pointer");
    SB_TRY(sb_use_device(h->device));
    *found = 0;
    *best_id = 0;
    *max_score = 0.f;
    *n_suspected = 0;
    int n_scan = 0;
    while (n_scan < h->n && !((uint64_t)cur_kf_id - (uint64_t)h->h_ids[n_scan] < (uint64_t)min_gap)) n_scan++;
    if (n_scan == 0) return SB_OK;
    SB_CUDA(cudaMemcpyAsync(h->d_query, query, LCD_DIM * 4, cudaMemcpyHostToDevice, h->stream));
    const int n_all = h->n;
    h->n = n_scan;  // only the scanned prefix is scored
    int rc = launch_scores(h, 1, h->d_query, h->d_scores, n_all);
    h->n = n_all;
    SB_TRY(rc);
    SB_CUDA(cudaMemcpyAsync(h->h_scores, h->d_scores, (size_t)n_scan * 4, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaStreamSynchronize(h->stream));
    float mx = 0.f;
    int cnt = 0;
    int64_t bid = 0;
    for (int i = 0; i < n_scan; i++) {
        const float s = h->h_scores[i];
        if (s > mx) { mx = s; bid = h->h_ids[i]; }
        if (s > thres_low) cnt++;
    }
    *max_score = mx;
    *best_id = bid;
    *n_suspected = cnt;
    *found = !(mx < thres_high || cnt > max_suspected);
    return SB_OK;
}
