// stereo.cu — the per-frame front end as ONE call: ORB extraction on both views of a batch of stereo
// frames followed by left<->right Hamming matching, host buffers in / host buffers out.
//
// This is the drop-in entry point for BASELINE config 2 (ORBextractor::DetectAndCompute on the left and
// the right image, reference src/ORBextractor.cpp:922-985, then BFMatcher::match(query = left, train =
// right) with the semantics of src/loopclosing.cpp:172).  Descriptors never leave the device between
// the two stages.  submit() only enqueues (H2D copies, kernels, D2H copies) on the handle's stream;
// wait() synchronises and reports errors — two handles used alternately overlap one batch's copies
// with the other batch's kernels.
#include <string.h>

#include "common.cuh"

struct sb_stereo {
    int device, max_pairs, cap, max_w, max_h;
    sb_orb_t *orb;
    sb_matcher_t *mat;
    cudaStream_t stream;
    uint8_t *d_img;      // [max_pairs][2][h][row]
    sb_keypoint *d_kps;  // [max_pairs][2][cap]
    uint8_t *d_desc;     // [max_pairs][2][cap][32]
    int32_t *d_counts;   // [max_pairs][2]
    int32_t *d_midx, *d_mdist;  // [max_pairs][cap]
    int pending;
    int pending_probe;  // probe builds only
    cudaStream_t compute;        // sb_stereo_set_compute_stream: kernels here, copies on `stream`; null = everything on `stream`
    cudaEvent_t ev_in, ev_done;  // copy in finished / kernels finished
    int32_t *h_flags;            // page-locked copy of the extractor's device flags (compute-stream mode)
};

static void free_stereo(sb_stereo *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->orb) sb_orb_destroy(h->orb);
    if (h->mat) sb_matcher_destroy(h->mat);
    void *ptrs[] = {h->d_img, h->d_kps, h->d_desc, h->d_counts, h->d_midx, h->d_mdist};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_done) cudaEventDestroy(h->ev_done);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    delete h;
}

extern "C" int sb_stereo_create(sb_stereo_t **out, int device, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                                int minThFAST, int max_w, int max_h, int max_pairs) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(out, "null handle pointer");
    *out = nullptr;
    SB_REQUIRE(max_pairs >= 1 && max_pairs <= 2048, "max_pairs out of range [1, 2048]");
    SB_TRY(sb_use_device(device));
    sb_stereo *h = new sb_stereo();
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->max_pairs = max_pairs;
    h->max_w = max_w;
    h->max_h = max_h;
    int rc = sb_orb_create(&h->orb, device, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_w, max_h, 2 * max_pairs);
    if (rc != SB_OK) { free_stereo(h); return rc; }
    h->cap = sb_orb_capacity(h->orb);
    rc = sb_matcher_create(&h->mat, device, max_pairs, h->cap);
    if (rc != SB_OK) { free_stereo(h); return rc; }
    const size_t P = max_pairs, row = sb_align_up((size_t)max_w, 16);
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_img, P * 2 * row * max_h);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_kps, P * 2 * h->cap * sizeof(sb_keypoint));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_desc, P * 2 * h->cap * 32);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_counts, P * 2 * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_midx, P * h->cap * 4);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_mdist, P * h->cap * 4);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_flags, 16);
    if (e != cudaSuccess) {
        sb_set_error("sb_stereo_create: %s", cudaGetErrorString(e));
        free_stereo(h);
        return SB_ERR_CUDA;
    }
    sb_orb_set_stream(h->orb, h->stream);
    sb_matcher_set_stream(h->mat, h->stream);
    *out = h;
    return SB_OK;
}

extern "C" int sb_stereo_destroy(sb_stereo_t *h) {
    SB_NVTX_FN();
    if (h) {
        cudaSetDevice(h->device);
        cudaDeviceSynchronize();
        free_stereo(h);
    }
    return SB_OK;
}

extern "C" int sb_stereo_capacity(const sb_stereo_t *h) { return h ? h->cap : SB_ERR_INVALID; }

extern "C" int sb_stereo_set_compute_stream(sb_stereo_t *h, void *stream) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    SB_REQUIRE(!h->pending, "a batch is in flight");
    h->compute = (cudaStream_t)stream;
    cudaStream_t k = h->compute ? h->compute : h->stream;
    sb_orb_set_stream(h->orb, k);
    sb_matcher_set_stream(h->mat, k);
    return SB_OK;
}

// Enqueue one batch.  images: `pairs` frames, frame p = left plane then right plane, each `hgt` rows of
// `stride` bytes, consecutive frames `frame_pitch` bytes apart (host memory; pinned memory makes the
// copies asynchronous).  Outputs (host): kps [pairs][2][cap], desc [pairs][2][cap][32], counts [pairs][2],
// match_idx / match_dist [pairs][cap] (query = left row, train = right row; -1 where the right set is empty).
extern "C" int sb_stereo_submit(sb_stereo_t *h, int pairs, const uint8_t *images, int64_t frame_pitch, int64_t view_pitch,
                                int w, int hgt, int stride, sb_keypoint *kps, uint8_t *desc, int32_t *counts,
                                int32_t *match_idx, int32_t *match_dist) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h && images && kps && desc && counts && match_idx && match_dist, "null pointer");
    SB_REQUIRE(pairs >= 1 && pairs <= h->max_pairs, "pairs out of range [1, max_pairs]");
    SB_REQUIRE(w >= 62 && w <= h->max_w && hgt >= 62 && hgt <= h->max_h && stride >= w, "bad image size / stride");
    SB_REQUIRE(!h->pending, "previous batch not waited for");
    SB_TRY(sb_use_device(h->device));
    // Staging: when the planes are equally spaced the whole batch is ONE contiguous copy (rows keep the
    // caller's stride; the extractor's level-0 kernel re-pitches on the device — a 2-D copy with 1241-byte
    // rows would cost one DMA descriptor per row).  Otherwise plane-by-plane 2-D copies.
    const size_t row = sb_align_up((size_t)w, 16), plane_cap = sb_align_up((size_t)h->max_w, 16) * h->max_h;
    size_t plane = row * hgt;
    int dev_stride = (int)row;
    if ((size_t)view_pitch == (size_t)stride * hgt && (size_t)frame_pitch == 2 * (size_t)view_pitch &&
        (size_t)stride * hgt <= plane_cap) {
        plane = (size_t)stride * hgt;
        dev_stride = stride;
#ifdef SB_PROBE_SKIP_H2D  // tools/ probe builds only (csrc/Makefile EXTRA=...): where the end-to-end time goes — only the first batch is copied
        if (!h->pending_probe)
#endif
        SB_CUDA(cudaMemcpyAsync(h->d_img, images, (size_t)pairs * 2 * plane, cudaMemcpyHostToDevice, h->stream));
#ifdef SB_PROBE_SKIP_H2D
        h->pending_probe = 1;
#endif
    } else {
        for (int p = 0; p < pairs; p++)
            for (int v = 0; v < 2; v++)
                SB_CUDA(cudaMemcpy2DAsync(h->d_img + ((size_t)p * 2 + v) * plane, row, images + p * frame_pitch + v * view_pitch,
                                          (size_t)stride, (size_t)w, (size_t)hgt, cudaMemcpyHostToDevice, h->stream));
    }
    const int cap = h->cap;
    if (h->compute) {  // the kernels wait for this batch's images, not for the other handles' copies
        SB_CUDA(cudaEventRecord(h->ev_in, h->stream));
        SB_CUDA(cudaStreamWaitEvent(h->compute, h->ev_in, 0));
    }
    SB_TRY(sb_orb_detect_and_compute_dev(h->orb, 2 * pairs, h->d_img, (int64_t)plane, nullptr, 0, w, hgt, dev_stride, 0, h->d_kps,
                                         h->d_desc, h->d_counts, cap));
    SB_TRY(sb_hamming_match_dev(h->mat, pairs, h->d_desc, (int64_t)2 * cap * 32, h->d_counts, 2, h->d_desc + (size_t)cap * 32,
                                (int64_t)2 * cap * 32, h->d_counts + 1, 2, cap, h->d_midx, h->d_mdist, cap));
    if (h->compute) {  // the copies out wait for this batch's kernels; the status flags travel with them
        SB_TRY(sb_orb_status_async(h->orb, h->h_flags));
        SB_CUDA(cudaEventRecord(h->ev_done, h->compute));
        SB_CUDA(cudaStreamWaitEvent(h->stream, h->ev_done, 0));
    }
#ifndef SB_PROBE_SKIP_D2H
    SB_CUDA(cudaMemcpyAsync(kps, h->d_kps, (size_t)pairs * 2 * cap * sizeof(sb_keypoint), cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(desc, h->d_desc, (size_t)pairs * 2 * cap * 32, cudaMemcpyDeviceToHost, h->stream));
#endif
    SB_CUDA(cudaMemcpyAsync(counts, h->d_counts, (size_t)pairs * 2 * 4, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(match_idx, h->d_midx, (size_t)pairs * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    SB_CUDA(cudaMemcpyAsync(match_dist, h->d_mdist, (size_t)pairs * cap * 4, cudaMemcpyDeviceToHost, h->stream));
    h->pending = 1;
    return SB_OK;
}

// Wait for the submitted batch; SB_ERR_OVERFLOW / SB_ERR_CAPACITY as flagged on the device.
extern "C" int sb_stereo_wait(sb_stereo_t *h) {
    SB_NVTX_FN();
    sb_clear_error();
    SB_REQUIRE(h, "null handle");
    if (!h->pending) return SB_OK;
    h->pending = 0;
    if (h->compute) {
        SB_TRY(sb_use_device(h->device));
        SB_CUDA(cudaStreamSynchronize(h->stream));  // the copies out, hence this batch's kernels; later batches keep running
        return sb_orb_status_decode(h->h_flags);
    }
    return sb_orb_sync_status(h->orb);
}

extern "C" int sb_stereo_extract_match(sb_stereo_t *h, int pairs, const uint8_t *images, int64_t frame_pitch,
                                       int64_t view_pitch, int w, int hgt, int stride, sb_keypoint *kps, uint8_t *desc,
                                       int32_t *counts, int32_t *match_idx, int32_t *match_dist) {
    SB_NVTX_FN();
    SB_TRY(sb_stereo_submit(h, pairs, images, frame_pitch, view_pitch, w, hgt, stride, kps, desc, counts, match_idx, match_dist));
    return sb_stereo_wait(h);
}
