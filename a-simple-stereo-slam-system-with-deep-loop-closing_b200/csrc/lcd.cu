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
    if (h->h_scores) cudaFreeHost(h->h_scores);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    free(h->h_ids);
    delete h;
}

extern "C" int sb_lcd_create(sb_lcd_t **out, int device, int capacity, int dtype, int max_queries) {
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
    if (e != cudaSuccess) {
        sb_set_error("sb_lcd_create: %s", cudaGetErrorString(e));
        free_lcd(h);
        return SB_ERR_CUDA;
    }
    h->stream = h->own_stream;
    *out = h;
    return SB_OK;
}

extern "C" int sb_lcd_destroy(sb_lcd_t *h) {
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_lcd(h);
    }
    return SB_OK;
}

extern "C" int sb_lcd_set_stream(sb_lcd_t *h, void *stream) {
    SB_REQUIRE(h, "null handle");
    h->stream = stream ? (cudaStream_t)stream : h->own_stream;
    return SB_OK;
}

extern "C" int sb_lcd_size(const sb_lcd_t *h) { return h ? h->n : SB_ERR_INVALID; }

// LoopClosing::AddToDatabase (src/loopclosing.cpp:651-659): _mvDatabase is a std::map keyed by KF id;
// ids must arrive in ascending order (they do: keyframes are processed in creation order).
extern "C" int sb_lcd_add(sb_lcd_t *h, int64_t kf_id, const float *descr) {
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
    sb_clear_error();
    SB_REQUIRE(h && d_queries && d_scores, "null pointer");
    SB_REQUIRE(nq >= 1 && nq <= 65535 && score_stride >= h->n, "bad nq / score_stride");
    SB_TRY(sb_use_device(h->device));
    return launch_scores(h, nq, d_queries, d_scores, score_stride);
}

// DeepLCD::score of `nq` query descriptors against every database row: scores [nq][size] (host).
extern "C" int sb_lcd_score(sb_lcd_t *h, int nq, const float *queries, float *scores) {
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
