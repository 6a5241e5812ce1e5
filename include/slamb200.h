/*
 * slamb200.h — C ABI of the B200-native hot path of the stereo SLAM system.
 *
 * This is the drop-in boundary.  The reference (Mingrui-Yu/A-Simple-Stereo-SLAM-System-with-
 * Deep-Loop-Closing, paths below relative to its root) has no FFI: its operator surface is the
 * C++ classes compiled into libmyslam.so.  Each entry point here replaces the arithmetic behind
 * one of those methods; the C++ adaptor classes in
 *   a-simple-stereo-slam-system-with-deep-loop-closing_b200/host/
 * keep the reference's class/method names on top of these calls.
 *
 * Conventions
 *   - every function returns an int status: SB_OK (0) or a negative SB_ERR_*; nothing throws;
 *   - opaque handles own all device memory and one CUDA stream; a handle is not thread-safe,
 *     different handles are independent (the reference shares one ORBextractor between two
 *     threads, src/system.cpp:54,66 — the adaptor gives each calling thread its own handle);
 *   - "_dev" variants take DEVICE pointers, enqueue on the handle's stream and return without
 *     synchronising; the plain variants take HOST pointers, copy in/out and synchronise;
 *   - there is no CPU fallback: without a CUDA device every compute call returns SB_ERR_CUDA.
 */
#ifndef SLAMB200_H
#define SLAMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID (-1)   /* bad argument (null pointer, size out of range, ...) */
#define SB_ERR_CUDA (-2)      /* CUDA runtime error; sb_last_error() has the text */
#define SB_ERR_CAPACITY (-3)  /* a per-call capacity fixed at create time was exceeded */
#define SB_ERR_OVERFLOW (-4)  /* more FAST candidates on one pyramid level than the handle can hold */

/* Text of the last error on the calling thread ("" if none). */
const char *sb_last_error(void);
/* Library version string and compiled SM architecture. */
const char *sb_version(void);

/* Page-locked host memory for the buffers handed to the asynchronous host-pointer entry points (sb_stereo_submit,
 * sb_ba_submit): with pageable memory the copies — and therefore the submit call — block until they are done. */
int sb_host_alloc(void **ptr, size_t bytes);
int sb_host_free(void *ptr);

/* field-for-field cv::KeyPoint (28 bytes) */
typedef struct sb_keypoint {
    float x, y;      /* pt, level-0 pixel coordinates */
    float size;      /* 7 from FAST, or 31*scale[octave] */
    float angle;     /* degrees in [0,360), -1 if not computed */
    float response;  /* FAST corner score */
    int32_t octave;  /* pyramid level */
    int32_t class_id;
} sb_keypoint;

/* ---------------------------------------------------------------------------------------------
 * ORB extractor — replaces myslam::ORBextractor (include/myslam/ORBextractor.h:47-138,
 * src/ORBextractor.cpp:384-1265).  One handle processes up to max_batch images of up to
 * max_w x max_h pixels per call.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_orb sb_orb_t;

/* ORBextractor::ORBextractor (src/ORBextractor.cpp:384-445). */
int sb_orb_create(sb_orb_t **h, int device, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                  int minThFAST, int max_w, int max_h, int max_batch);
int sb_orb_destroy(sb_orb_t *h);
/* Use `stream` (a cudaStream_t) instead of the handle's own stream for all later calls. */
int sb_orb_set_stream(sb_orb_t *h, void *stream);
/* Waits for the handle's stream and reports what the asynchronous "_dev" calls since the last
 * check flagged on the device: SB_OK, SB_ERR_OVERFLOW or SB_ERR_CAPACITY. */
int sb_orb_sync_status(sb_orb_t *h);
/* The same check without a stream synchronisation: enqueues, on the handle's stream, a copy of the device flags into
 * host_flags (4 ints, page-locked) and their reset; after the caller has synchronised in its own way (an event, another
 * stream that waited for this one) sb_orb_status_decode(host_flags) gives SB_OK / SB_ERR_OVERFLOW / SB_ERR_CAPACITY. */
int sb_orb_status_async(sb_orb_t *h, int32_t *host_flags);
int sb_orb_status_decode(const int32_t *host_flags);
/* Per-image keypoint capacity the caller must provide to the calls below
 * (sum over levels of max(quota + 3, 4 * nIni) for the pyramid calls). */
int sb_orb_capacity(const sb_orb_t *h);
/* GetLevels / GetScaleFactors / GetInverseScaleFactors / GetScaleSigmaSquares /
 * GetInverseScaleSigmaSquares (ORBextractor.h:88-106) and mnFeaturesPerLevel; each array has
 * nlevels entries; any pointer may be null. */
int sb_orb_get_tables(const sb_orb_t *h, int *nlevels, float *scale, float *inv_scale, float *sigma2,
                      float *inv_sigma2, int *features_per_level);

/* ORBextractor::DetectAndCompute (src/ORBextractor.cpp:922-985) on `batch` images.
 *   img[b], mask[b] : CV_8UC1 planes, h rows of `stride` bytes (mask: `mstride`); mask == null or
 *                     mask[b] == null means "all 255" (the reference requires a mask).
 *   kps  [batch][cap], desc [batch][cap][32], counts [batch]; desc may be null
 *                     (== ORBextractor::DetectWithPyramid, :1135-1176).
 * Keypoints are level-major; inside a level in the reference's quadtree list order. */
int sb_orb_detect_and_compute(sb_orb_t *h, int batch, const uint8_t *const *img, const uint8_t *const *mask, int w,
                              int hgt, int stride, int mstride, sb_keypoint *kps, uint8_t *desc, int32_t *counts,
                              int cap);
/* Same, all pointers on the device; images are d_img + b * img_pitch_bytes (b < batch), masks
 * likewise (d_mask may be null).  Asynchronous on the handle's stream. */
int sb_orb_detect_and_compute_dev(sb_orb_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes,
                                  const uint8_t *d_mask, int64_t mask_pitch_bytes, int w, int hgt, int stride,
                                  int mstride, sb_keypoint *d_kps, uint8_t *d_desc, int32_t *d_counts, int cap);

/* ORBextractor::Detect (src/ORBextractor.cpp:989-1074): level 0 only, N = nfeatures, keypoints keep
 * FAST's size 7 / angle -1 / octave 0.  The reference returns silently on an empty mask; here a
 * null mask means "all 255". */
int sb_orb_detect(sb_orb_t *h, int batch, const uint8_t *const *img, const uint8_t *const *mask, int w, int hgt,
                  int stride, int mstride, sb_keypoint *kps, int32_t *counts, int cap);
int sb_orb_detect_dev(sb_orb_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes, const uint8_t *d_mask,
                      int64_t mask_pitch_bytes, int w, int hgt, int stride, int mstride, sb_keypoint *d_kps,
                      int32_t *d_counts, int cap);

/* ORBextractor::ScreenAndComputeKPsParams (src/ORBextractor.cpp:1083-1129) on one image.
 * `in` [n_in] is mutated exactly like the reference mutates its input vector (pt /= scale,
 * pt *= scale round trip); survivors are written to out[0 .. *n_out) in input order. */
int sb_orb_screen_params(sb_orb_t *h, const uint8_t *img, int w, int hgt, int stride, sb_keypoint *in, int n_in,
                         sb_keypoint *out, int32_t *n_out);
/* ORBextractor::CalcDescriptors (src/ORBextractor.cpp:1180-1226): desc [n][32], row i <-> kps[i]. */
int sb_orb_calc_descriptors(sb_orb_t *h, const uint8_t *img, int w, int hgt, int stride, const sb_keypoint *kps,
                            int n, uint8_t *desc);
/* LoopClosing::ProcessNewKF's two extractor calls (src/loopclosing.cpp:107-112) for a batch of keyframe images in one
 * pass: ScreenAndComputeKPsParams, then CalcDescriptors on its survivors.  in [batch][cap_in] (first n_in[b] live, mutated
 * like the reference's vector), out [batch][cap_in] survivors in input order, n_out [batch], desc [batch][cap_in][32]. */
int sb_orb_screen_describe(sb_orb_t *h, int batch, const uint8_t *const *img, int w, int hgt, int stride, sb_keypoint *in,
                           const int32_t *n_in, int cap_in, sb_keypoint *out, int32_t *n_out, uint8_t *desc);

/* Debug/inspection: copy pyramid level `level` of image `b` of the LAST call to the host.
 * which: 0 = mvImagePyramid, 1 = Gaussian-blurred working Mat, 2 = mvMaskPyramid.
 * out has lh rows of lw bytes (tight); lw/lh are returned. */
int sb_orb_debug_level(sb_orb_t *h, int b, int level, int which, uint8_t *out, int out_bytes, int *lw, int *lh);
/* Debug: FAST candidates handed to the quadtree for (image b, level): packed
 * x | y << 12 | response << 24 (border-relative x, y), unordered.  *n receives the count. */
int sb_orb_debug_candidates(sb_orb_t *h, int b, int level, uint32_t *out, int cap, int32_t *n);

/* Inspection: capture the TMA tile and response plane of one FAST cell of image 0 during the next call. */
int sb_orb_debug_fast_cell(sb_orb_t *h, int cell, uint8_t *out, int out_bytes, int *tile_bytes);

/* Per-stage timing with CUDA events recorded on the streams the kernels are launched on.
 * sb_orb_profile(h, 1) starts a fresh recording, sb_orb_profile(h, 0) stops it;
 * sb_orb_profile_read sums device milliseconds and kernel launches per stage. */
#define SB_ORB_STAGE_COPY 0      /* level 0 re-pitch                    */
#define SB_ORB_STAGE_RESIZE 1    /* pyramid levels 1 .. n-1             */
#define SB_ORB_STAGE_FAST 2      /* grid FAST + NMS + threshold select  */
#define SB_ORB_STAGE_QUADTREE 3  /* DistributeOctTree                   */
#define SB_ORB_STAGE_BLUR 4      /* Gaussian 7x7 on every level         */
#define SB_ORB_STAGE_DESCRIBE 5  /* IC_Angle + rBRIEF + output          */
#define SB_ORB_STAGE_COUNT 6
int sb_orb_profile(sb_orb_t *h, int enable);
int sb_orb_profile_read(sb_orb_t *h, float *ms, int32_t *launches, int nstages);

/* ---------------------------------------------------------------------------------------------
 * Brute-force Hamming 1-NN — replaces cv::BFMatcher(NORM_HAMMING)::match as used by
 * LoopClosing::MatchFeatures (src/loopclosing.cpp:33,172): for every query row the nearest train
 * row, ties -> lowest trainIdx.  `batch` independent (query set, train set) problems per call.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_matcher sb_matcher_t;
int sb_matcher_create(sb_matcher_t **m, int device, int max_batch, int max_rows);
int sb_matcher_destroy(sb_matcher_t *m);
int sb_matcher_set_stream(sb_matcher_t *m, void *stream);
/* q [batch][cap][32], t [batch][cap][32], nq/nt [batch]  ->  train_idx, dist [batch][cap]
 * (train_idx = -1, dist = -1 where the train set is empty). */
int sb_hamming_match(sb_matcher_t *m, int batch, const uint8_t *q, const int32_t *nq, const uint8_t *t,
                     const int32_t *nt, int cap, int32_t *train_idx, int32_t *dist);
/* Device variant with explicit strides so descriptor sets can live interleaved (e.g. the
 * [pair][left|right][cap][32] layout sb_orb_detect_and_compute_dev writes):
 *   query set i = d_q + i * q_set_stride (bytes), its row count = d_nq[i * nq_stride]; train alike;
 *   results at d_train_idx + i * out_stride (elements). */
int sb_hamming_match_dev(sb_matcher_t *m, int batch, const uint8_t *d_q, int64_t q_set_stride, const int32_t *d_nq,
                         int nq_stride, const uint8_t *d_t, int64_t t_set_stride, const int32_t *d_nt,
                         int nt_stride, int max_rows, int32_t *d_train_idx, int32_t *d_dist, int64_t out_stride);

/* ---------------------------------------------------------------------------------------------
 * Stereo front end in one call — ORBextractor::DetectAndCompute (src/ORBextractor.cpp:922-985) on the
 * left and the right view of `pairs` frames, then BFMatcher-Hamming match(query = left descriptors,
 * train = right descriptors) with the semantics of src/loopclosing.cpp:172; host buffers in and out,
 * descriptors stay on the device between the stages.  sb_stereo_submit only enqueues on the handle's
 * stream (truly asynchronous when the host buffers are pinned) and sb_stereo_wait synchronises, so two
 * handles used alternately overlap copies with kernels.
 *   images: frame p's left plane at images + p * frame_pitch, its right plane view_pitch bytes later,
 *           each `hgt` rows of `stride` bytes;
 *   kps [pairs][2][cap], desc [pairs][2][cap][32], counts [pairs][2], match_idx / match_dist [pairs][cap]
 *   with cap = sb_stereo_capacity().
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_stereo sb_stereo_t;
int sb_stereo_create(sb_stereo_t **h, int device, int nfeatures, float scaleFactor, int nlevels, int iniThFAST,
                     int minThFAST, int max_w, int max_h, int max_pairs);
int sb_stereo_destroy(sb_stereo_t *h);
int sb_stereo_capacity(const sb_stereo_t *h);
int sb_stereo_submit(sb_stereo_t *h, int pairs, const uint8_t *images, int64_t frame_pitch, int64_t view_pitch, int w,
                     int hgt, int stride, sb_keypoint *kps, uint8_t *desc, int32_t *counts, int32_t *match_idx,
                     int32_t *match_dist);
int sb_stereo_wait(sb_stereo_t *h);
/* Several handles in flight, ONE kernel stream: after this call the handle's kernels run on `stream` (a cudaStream_t shared
 * by all the handles of a pipeline, in submission order) while its host<->device copies stay on the handle's own stream,
 * ordered against the kernels by events.  Batches then follow one another on the device like a single-stream loop — a
 * concurrent fat-CTA kernel (local BA: one CTA = one whole SM) finds free SMs at every kernel boundary — and the copies of
 * neighbouring batches still overlap the kernels.  stream = NULL restores the default (everything on the handle's stream). */
int sb_stereo_set_compute_stream(sb_stereo_t *h, void *stream);
int sb_stereo_extract_match(sb_stereo_t *h, int pairs, const uint8_t *images, int64_t frame_pitch, int64_t view_pitch,
                            int w, int hgt, int stride, sb_keypoint *kps, uint8_t *desc, int32_t *counts,
                            int32_t *match_idx, int32_t *match_dist);

/* ---------------------------------------------------------------------------------------------
 * Local bundle adjustment — replaces the g2o solve inside Backend::OptimizeActiveMap
 * (src/backend.cpp:126-269; edge/vertex arithmetic include/myslam/g2o_types.h:25-59,106-153):
 * Levenberg-Marquardt with the landmarks marginalised (Schur), Huber kernel, up to `outer_max`
 * rounds of `inner_iters` iterations until more than half of the edges have chi2 <= chi2_th.
 * A call solves `n_windows` independent windows; every window owns a fixed-capacity slot:
 *   poses  [W][max_poses][7]   in/out  qx qy qz qw tx ty tz of T_cw (Sophus storage order); no pose is fixed
 *   points [W][max_points][3]  in/out  world positions
 *   fixed  [W][max_points]     1 = landmark kept constant (first observer outside the window, :175-177)
 *   obs_pose / obs_point [W][max_obs], uv [W][max_obs][2]  one EdgeProjection per entry (information I2)
 *   K = fx fy cx cy,  cam_ext7 = extrinsics of the observing camera (identity for the left camera)
 *   chi2 [W][max_obs]   e'e of each edge at the last error evaluation (what EdgeProjection::chi2() returns
 *                       after optimize()),  outlier [W][max_obs] = chi2 > chi2_th (:236-251)
 *   info [W][4] = outer rounds run, LM iterations run, inliers, outliers
 * The reference's values: huber_delta 5.991, chi2_th 5.991, outer_max 5, inner_iters 10.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_ba sb_ba_t;
int sb_ba_create(sb_ba_t **h, int device, int max_windows, int max_poses, int max_points, int max_obs);
int sb_ba_destroy(sb_ba_t *h);
int sb_ba_set_stream(sb_ba_t *h, void *stream);
int sb_ba_solve(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points, const int32_t *n_obs,
                double *poses, double *points, const uint8_t *fixed, const int32_t *obs_pose, const int32_t *obs_point,
                const double *uv, const double *K, const double *cam_ext7, double huber_delta, double chi2_th,
                int outer_max, int inner_iters, double *chi2, uint8_t *outlier, int32_t *info);
/* Asynchronous host-pointer form — the reference's Backend optimises in its own thread beside the front end
 * (src/backend.cpp:29-45, BackendLoop): sb_ba_submit enqueues copies + solve + copies back on the handle's stream and
 * returns; sb_ba_wait blocks until the batch is complete and reports invalid windows.  One batch in flight per handle;
 * host arrays must stay valid (pinned for true overlap) until sb_ba_wait returns.  sb_ba_solve = submit + wait. */
int sb_ba_submit(sb_ba_t *h, int n_windows, const int32_t *n_poses, const int32_t *n_points, const int32_t *n_obs,
                 double *poses, double *points, const uint8_t *fixed, const int32_t *obs_pose, const int32_t *obs_point,
                 const double *uv, const double *K, const double *cam_ext7, double huber_delta, double chi2_th,
                 int outer_max, int inner_iters, double *chi2, uint8_t *outlier, int32_t *info);
int sb_ba_wait(sb_ba_t *h);
/* Same with every array on the device (K and cam_ext7 stay host pointers); asynchronous. */
int sb_ba_solve_dev(sb_ba_t *h, int n_windows, const int32_t *d_n_poses, const int32_t *d_n_points,
                    const int32_t *d_n_obs, double *d_poses, double *d_points, const uint8_t *d_fixed,
                    const int32_t *d_obs_pose, const int32_t *d_obs_point, const double *d_uv, const double *K,
                    const double *cam_ext7, double huber_delta, double chi2_th, int outer_max, int inner_iters,
                    double *d_chi2, uint8_t *d_outlier, int32_t *d_info);

/* ---------------------------------------------------------------------------------------------
 * DeepLCD descriptor scoring — replaces DeepLCD::score (src/deeplcd.cpp:35-39) and the database
 * scan of LoopClosing::DetectLoop (src/loopclosing.cpp:124-161).  Descriptors are the 1064-float,
 * L2-normalised outputs of DeepLCD::calcDescr (src/deeplcd.cpp:55-91); the database keeps one row per
 * keyframe in ascending keyframe id (the reference's std::map order), stored fp32 or fp16 in HBM.
 * --------------------------------------------------------------------------------------------- */
#define SB_LCD_DIM 1064
#define SB_LCD_FP32 0
#define SB_LCD_FP16 1
typedef struct sb_lcd sb_lcd_t;
int sb_lcd_create(sb_lcd_t **h, int device, int capacity, int dtype, int max_queries);
int sb_lcd_destroy(sb_lcd_t *h);
int sb_lcd_set_stream(sb_lcd_t *h, void *stream);
int sb_lcd_size(const sb_lcd_t *h);
/* LoopClosing::AddToDatabase (src/loopclosing.cpp:651-659); ids ascending. */
int sb_lcd_add(sb_lcd_t *h, int64_t kf_id, const float *descr);
int sb_lcd_add_batch(sb_lcd_t *h, int n, const int64_t *kf_ids, const float *descr);
/* _mvDatabase.erase(id) (src/loopclosing.cpp:73-75). */
int sb_lcd_remove(sb_lcd_t *h, int64_t kf_id);
/* scores [nq][size]: DeepLCD::score of every query against every row (fp32 accumulate). */
int sb_lcd_score(sb_lcd_t *h, int nq, const float *queries, float *scores);
int sb_lcd_score_dev(sb_lcd_t *h, int nq, const float *d_queries, float *d_scores, int score_stride);
/* LoopClosing::DetectLoop: the reference's values are thres_high 0.94, thres_low 0.92
 * (config LCD.similarityScoreThreshold.{high,low}), min_gap 20, max_suspected 3. */
int sb_lcd_detect_loop(sb_lcd_t *h, int64_t cur_kf_id, const float *query, float thres_high, float thres_low,
                       int min_gap, int max_suspected, int *found, int64_t *best_id, float *max_score,
                       int *n_suspected);

/* ---------------------------------------------------------------------------------------------
 * Pose-graph optimisation — replaces the g2o solve inside LoopClosing::PoseGraphOptimization
 * (src/loopclosing.cpp:537-646; EdgePoseGraph include/myslam/g2o_types.h:157-190): Levenberg-
 * Marquardt, `iters` iterations (the reference: 20), numeric Jacobians (step 1e-9) as g2o computes
 * them for an edge without linearizeOplus, information I6, no robust kernel.
 *   poses [n][7] in/out (qx qy qz qw tx ty tz of T_cw), vertices in ascending keyframe id;
 *   fixed [n]: the reference fixes the active keyframes, the loop keyframe and keyframe 0 (:559-562);
 *   edge k: vertex v0[k] -> vertex v1[k] with measurement meas[k] = T_v0 * T_v1^-1
 *           (mRelativePoseToLastKF / mRelativePoseToLoopKF, :571-599).
 * The graph must be a chain plus long-range edges between free vertices (loop edges; KITTI-00: 17): at most 64 with
 * sb_posegraph_create, at most `max_loops` (<= 4096) with sb_posegraph_create_loops — the reference re-adds every
 * historical loop edge on each call (:585-599), so size max_loops for the whole run; the workspace is
 * max_vertices x max_loops x 288 bytes.  More long-range edges than that: SB_ERR_CAPACITY, nothing is modified.
 *   info [4] = LM iterations, LM trials, free vertices, long-range edges;  stats [2] = chi2 before, after.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_posegraph sb_posegraph_t;
int sb_posegraph_create(sb_posegraph_t **h, int device, int max_vertices, int max_edges);
int sb_posegraph_create_loops(sb_posegraph_t **h, int device, int max_vertices, int max_edges, int max_loops);
int sb_posegraph_destroy(sb_posegraph_t *h);
int sb_posegraph_set_stream(sb_posegraph_t *h, void *stream);
int sb_posegraph_solve(sb_posegraph_t *h, int n_vertices, double *poses, const uint8_t *fixed, int n_edges,
                       const int32_t *v0, const int32_t *v1, const double *meas, int iters, int32_t *info,
                       double *stats);
int sb_posegraph_solve_dev(sb_posegraph_t *h, int n_vertices, double *d_poses, const uint8_t *d_fixed, int n_edges,
                           const int32_t *d_v0, const int32_t *d_v1, const double *d_meas, int iters, int32_t *d_info,
                           double *d_stats);

/* ---------------------------------------------------------------------------------------------
 * Pose-only optimisation (SURVEY §8f "next" row 1) — replaces the g2o solve inside
 * Frontend::EstimateCurrentPose (src/frontend.cpp:176-276) and LoopClosing::OptimizeCurrentPose
 * (src/loopclosing.cpp:339-433): one pose, one EdgeProjectionPoseOnly (include/myslam/g2o_types.h:63-102)
 * per matched map point, Huber kernel (g2o default delta 1.0), `pre_rounds` plain optimize(inner_iters)
 * calls (loop closing: 1, front end: 0) and then `rounds` (4) rounds of optimize(inner_iters) + chi2
 * classification (outliers sit out the next round; the robust kernels are removed after round rounds-2).
 * A call solves `n_frames` independent frames:
 *   poses [F][7] in/out (qx qy qz qw tx ty tz), points [F][max_obs][3] world positions of the matched map
 *   points, uv [F][max_obs][2] their pixel observations, K = fx fy cx cy,
 *   outlier [F][max_obs] out, info [F][4] = inliers (the functions' return value), LM iterations, rounds, 0.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_pose sb_pose_t;
int sb_pose_create(sb_pose_t **h, int device, int max_frames, int max_obs);
int sb_pose_destroy(sb_pose_t *h);
int sb_pose_set_stream(sb_pose_t *h, void *stream);
int sb_pose_solve(sb_pose_t *h, int n_frames, const int32_t *n_obs, double *poses, const double *points, const double *uv,
                  const double *K, double huber_delta, double chi2_th, int pre_rounds, int rounds, int inner_iters,
                  uint8_t *outlier, int32_t *info);
int sb_pose_solve_dev(sb_pose_t *h, int n_frames, const int32_t *d_n_obs, double *d_poses, const double *d_points,
                      const double *d_uv, const double *K, double huber_delta, double chi2_th, int pre_rounds, int rounds,
                      int inner_iters, uint8_t *d_outlier, int32_t *d_info);

/* ---------------------------------------------------------------------------------------------
 * Stereo triangulation (SURVEY §8f "next" row 3) — replaces myslam::triangulation
 * (include/myslam/algorithm.h:16-33) plus the callers' acceptance test (src/frontend.cpp:403,474) for n
 * left/right correspondences that share the two camera poses (Camera::Pose() of the left and right
 * camera): DLT by SVD of the 4x4 system, ok = sigma4 / sigma3 < ratio_th (reference: 1e-2) && z > 0;
 * accepted and rejected points are both written (like the reference computes pt_world before testing),
 * mapped by T_wc7 when given (currentPoseTwc * pcamera, src/frontend.cpp:477).
 *   uv_left / uv_right [n][2] float pixels, K = fx fy cx cy, poses = qx qy qz qw tx ty tz.
 * --------------------------------------------------------------------------------------------- */
int sb_triangulate(int device, int n, const float *uv_left, const float *uv_right, const double *K_left,
                   const double *K_right, const double *pose_left7, const double *pose_right7, const double *T_wc7,
                   double ratio_th, double *points, uint8_t *ok);
int sb_triangulate_dev(int device, void *stream, int n, const float *d_uv_left, const float *d_uv_right, const double *K_left,
                       const double *K_right, const double *pose_left7, const double *pose_right7, const double *T_wc7,
                       double ratio_th, double *d_points, uint8_t *d_ok);

/* ---------------------------------------------------------------------------------------------
 * Pyramidal Lucas-Kanade tracking (SURVEY §8f "next" row 1) — replaces cv::calcOpticalFlowPyrLK as called
 * by Frontend::TrackLastFrame (src/frontend.cpp:150-153) and Frontend::FindFeaturesInRight (:358-361):
 * winSize (win, win) = (11, 11), maxLevel (create time) = 3, criteria COUNT+EPS (max_count 30, eps 0.01),
 * OPTFLOW_USE_INITIAL_FLOW, minEigThreshold 1e-4.  `batch` image pairs per call, up to max_pts points each:
 *   prev_pts [batch][max_pts][2] float; next_pts [batch][max_pts][2] in (initial guess when
 *   use_initial_flow) / out; status [batch][max_pts] (1 = tracked); n_pts [batch].
 * The error output of the OpenCV call is not produced (the reference ignores it).
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_lk sb_lk_t;
int sb_lk_create(sb_lk_t **h, int device, int max_w, int max_h, int max_batch, int max_pts, int max_level);
int sb_lk_destroy(sb_lk_t *h);
int sb_lk_set_stream(sb_lk_t *h, void *stream);
int sb_lk_track(sb_lk_t *h, int batch, const uint8_t *const *prev, const uint8_t *const *next, int w, int hgt, int stride,
                const int32_t *n_pts, const float *prev_pts, float *next_pts, uint8_t *status, int win, int max_count,
                double eps, int use_initial_flow, float min_eig_th);
int sb_lk_track_dev(sb_lk_t *h, int batch, const uint8_t *d_prev_img, const uint8_t *d_next_img, int64_t img_pitch_bytes, int w,
                    int hgt, int stride, const int32_t *d_n_pts, const float *d_prev_pts, float *d_next_pts, uint8_t *d_status,
                    int win, int max_count, double eps, int use_initial_flow, float min_eig_th);

/* ---------------------------------------------------------------------------------------------
 * DeepLCD whole-image descriptor (SURVEY §8f "next" row 2) — replaces DeepLCD::calcDescrOriginalImg
 * (src/deeplcd.cpp:43-52: 7x7 sigma-0 Gaussian blur IN PLACE on the caller's image, resize to 160x120)
 * and DeepLCD::calcDescr (:55-91: u8 -> float / 255, Caffe Net::Forward, descriptor /= norm) for a
 * batch of images.  The network is data: `layers` restates deploy.prototxt (Convolution / ReLU /
 * Pooling MAX / LRN across channels, Caffe's shape rules; the trailing Flatten is implicit) and
 * `weights` is one flat fp32 buffer in Caffe's blob order — per Convolution layer W [Cout][Cin][k][k]
 * then bias [Cout] — i.e. the contents of calc.caffemodel (DeepLCD::DeepLCD, src/deeplcd.cpp:10-31).
 * The reference's network (1 x 120 x 160 in, 1064 out) is listed in INTEGRATION.md.
 * --------------------------------------------------------------------------------------------- */
#define SB_CALC_CONV 0
#define SB_CALC_RELU 1
#define SB_CALC_POOL_MAX 2
#define SB_CALC_LRN 3
typedef struct sb_calc_layer {
    int32_t type;                          /* SB_CALC_* */
    int32_t num_output, kernel, stride, pad; /* Convolution; Pooling uses kernel, stride, pad */
    int32_t local_size;                    /* LRN */
    float alpha, beta, k;                  /* LRN */
} sb_calc_layer;
typedef struct sb_calc sb_calc_t;
int sb_calc_create(sb_calc_t **h, int device, int in_h, int in_w, const sb_calc_layer *layers, int n_layers,
                   const float *weights, int64_t n_weights, int max_batch, int max_img_w, int max_img_h);
int sb_calc_destroy(sb_calc_t *h);
int sb_calc_set_stream(sb_calc_t *h, void *stream);
/* autoencoder_output->channels() (src/deeplcd.cpp:72): 1064 for the reference's network. */
int sb_calc_descr_dim(const sb_calc_t *h);
/* the network's input size (rows, cols): what sb_calc_descr expects */
int sb_calc_input_size(const sb_calc_t *h, int *in_h, int *in_w);
/* calcDescrOriginalImg: img = `batch` pointers to w x hgt u8 images; descr [batch][dim];
 * blurred_out = null, or `batch` pointers (entries may be null) that receive the blurred image —
 * pass the input pointers to reproduce the reference's in-place blur of KeyFrame::mImageLeft. */
int sb_calc_descr_original(sb_calc_t *h, int batch, const uint8_t *const *img, int w, int hgt, int stride, float *descr,
                           uint8_t *const *blurred_out);
/* calcDescr: images that already have the net's input size. */
int sb_calc_descr(sb_calc_t *h, int batch, const uint8_t *const *img, int stride, float *descr);
int sb_calc_descr_original_dev(sb_calc_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes, int w, int hgt, int stride,
                               float *d_descr, uint8_t *d_blurred);
int sb_calc_descr_dev(sb_calc_t *h, int batch, const uint8_t *d_img, int64_t img_pitch_bytes, int stride, float *d_descr);
/* DeepLCD::DeepLCD(network_definition_file, pre_trained_model_file, gpu_id) (src/deeplcd.cpp:10-31): reads
 * deploy.prototxt (protobuf text format) and calc.caffemodel (protobuf wire format of caffe.proto's
 * NetParameter; weights matched to layers by name like CopyTrainedLayersFrom) without Caffe or protobuf.
 * sb_calc_parse_caffe needs no device: layers / weights may be null to query n_layers / n_weights;
 * shape [2] = net input height, width. */
int sb_calc_parse_caffe(const char *prototxt_path, const char *caffemodel_path, sb_calc_layer *layers, int cap_layers, int *n_layers,
                        float *weights, int64_t cap_weights, int64_t *n_weights, int *shape);
int sb_calc_create_from_caffe(sb_calc_t **h, int device, const char *prototxt_path, const char *caffemodel_path, int max_batch,
                              int max_img_w, int max_img_h);

/* ---------------------------------------------------------------------------------------------
 * PnP with RANSAC (SURVEY §8f "next" row 4) — replaces cv::solvePnPRansac as LoopClosing::ComputeCorrectPose
 * calls it (src/loopclosing.cpp:259-268: objectPoints = map points seen by the loop keyframe, imagePoints =
 * matched keypoints of the current keyframe, K, no distortion, no extrinsic guess, iterationsCount 100,
 * reprojectionError 5.991, confidence 0.99, SOLVEPNP_ITERATIVE).  A call solves `n_problems` loop candidates:
 *   obj [P][max_points][3] float (cv::Point3f), img [P][max_points][2] float (cv::Point2f), K = fx fy cx cy;
 *   `iterations` hypotheses from minimal samples (P3P + a fourth point), all evaluated in parallel (OpenCV
 *   stops early at `confidence`: a subset of these), inlier <=> squared reprojection error <= reproj_err^2,
 *   best = most inliers, then Levenberg-Marquardt on the inliers of the best model;
 *   pose7 [P][7] = qx qy qz qw tx ty tz of T_cw (what cv::Rodrigues + Sophus::SE3d(R, t) give, :269-272),
 *   rvec_tvec [P][6] = cv's rvec, tvec; inlier [P][max_points] = 1 for the inliers of the best hypothesis;
 *   info [P][4] = found (0: fewer than 4 points or no hypothesis with >= 4 inliers — the reference's
 *   try/catch path, :262-267), inliers, hypotheses that had a solution, refinement iterations.
 * The samples come from a counter-based generator seeded by `seed`: the result is a deterministic function of
 * the arguments (cv::RNG state is not).  The follow-up OptimizeCurrentPose (:339-433) is sb_pose_solve.
 * --------------------------------------------------------------------------------------------- */
typedef struct sb_pnp sb_pnp_t;
int sb_pnp_create(sb_pnp_t **h, int device, int max_problems, int max_points);
int sb_pnp_destroy(sb_pnp_t *h);
int sb_pnp_set_stream(sb_pnp_t *h, void *stream);
int sb_pnp_ransac(sb_pnp_t *h, int n_problems, const int32_t *n_points, const float *obj, const float *img, const double *K,
                  int iterations, double reproj_err, uint64_t seed, double *pose7, double *rvec_tvec, uint8_t *inlier,
                  int32_t *info);
int sb_pnp_ransac_dev(sb_pnp_t *h, int n_problems, const int32_t *d_n_points, const float *d_obj, const float *d_img,
                      int max_points, const double *K, int iterations, double reproj_err, uint64_t seed, double *d_pose7,
                      double *d_rvec_tvec, uint8_t *d_inlier, int32_t *d_info);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU exchange step (BASELINE config 5).  The per-frame path shards by frame /
 * window with no collective; the one exchange is an all-gather of the keyframe poses every rank owns
 * before LoopClosing::PoseGraphOptimization (src/loopclosing.cpp:537-646), which reads the pose of
 * EVERY keyframe (:547-566).  The reference is single-process, so there is no reference signature:
 * these follow SURVEY.md 8b.  One ncclAllGather on the CALLER's communicator and stream; each rank
 * contributes a fixed-size record ([cap][7] doubles + its count), so no second collective.
 * NCCL is bound at run time from the caller's process (no link-time dependency; SB_ERR_CUDA with
 * the loader's text when absent).
 *   nccl_comm  ncclComm_t of the caller          stream  cudaStream_t (NULL = default stream)
 *   local      [cap][7] poses (qx,qy,qz,qw,tx,ty,tz), first n_local rows valid
 *   all        [world][cap][7]; counts [world]   (keyframe k of a round-robin shard = all[k % world][k / world])
 * sb_allgather_kf_poses: HOST pointers, stages through the communicator's device, synchronises.
 * sb_allgather_kf_poses_dev: DEVICE pointers, enqueue only; d_scratch = (1 + world) * (cap * 7 + 1) doubles.
 * ------------------------------------------------------------------------------------------------ */
int sb_allgather_kf_poses(void *nccl_comm, void *stream, const double *local, int n_local, double *all, int *counts,
                          int cap);
int sb_allgather_kf_poses_dev(void *nccl_comm, void *stream, const double *d_local, int n_local, double *d_all,
                              int32_t *d_counts, int cap, double *d_scratch);
/* Bootstrap for a host without a communicator: rank 0 makes the 128-byte id and ships it to the others over its
 * own channel; every rank calls sb_nccl_comm_init (ncclCommInitRank) on its device. */
int sb_nccl_version(int *version);
int sb_nccl_unique_id(uint8_t id128[128]);
int sb_nccl_comm_init(void **nccl_comm, int device, int world, int rank, const uint8_t id128[128]);
int sb_nccl_comm_destroy(void *nccl_comm);

#ifdef __cplusplus
}
#endif
#endif /* SLAMB200_H */
