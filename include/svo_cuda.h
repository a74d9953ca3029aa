/* include/svo_cuda.h -- C ABI of the B200-native stereo front-end (libsvo_b200.so).
 *
 * The reference (liuzhenboo/Stereo-Visual-Odometry) has no FFI/plugin layer: its boundary is the C++ class surface
 * lzb_vio::ORBextractor / lzb_vio::Tracking.  This header is the thin extern "C" layer those
 * classes call instead of OpenCV; every entry point names the reference call site it replaces.  Plain pointers and
 * sizes only; host pointers in, host pointers out unless a name ends in _dev; no exceptions cross this ABI.
 *
 * Return value: SVO_OK (0) or a negative svo_status; svo_last_error(ctx) gives the message.  A context is bound to one
 * CUDA device and one stream and is NOT re-entrant (like the reference's ORBextractor, whose mvImagePyramid is
 * per-object scratch, tracking.cpp:508-509): use one context per host thread / GPU.
 *
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device svo_ctx_create fails with SVO_ERR_CUDA.
 */
#ifndef SVO_CUDA_H
#define SVO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_ABI_VERSION 5

typedef enum svo_status {
    SVO_OK = 0,
    SVO_ERR_INVALID = -1,        /* bad argument (null pointer, size out of range, wrong state) */
    SVO_ERR_CUDA = -2,           /* CUDA runtime error / no device */
    SVO_ERR_CAPACITY = -3,       /* an internal or caller-provided capacity was exceeded (never silently truncated) */
    SVO_ERR_NOT_CONFIGURED = -4  /* svo_orb_configure() has not been called */
} svo_status;

typedef struct svo_ctx svo_ctx;

/* Binary-compatible with cv::KeyPoint (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id. */
typedef struct svo_keypoint {
    float x, y, size, angle, response;
    int32_t octave, class_id;
} svo_keypoint;

/* ---- context ------------------------------------------------------------------------------------------------- */
/* max_images: how many images one batched call may carry (device scratch is sized for it). */
int svo_ctx_create(svo_ctx** ctx, int device, int max_w, int max_h, int max_images);
void svo_ctx_destroy(svo_ctx* ctx);
const char* svo_last_error(const svo_ctx* ctx);
int svo_abi_version(void);
/* Debug: with SVO_GUARD=1 in the environment every device buffer of the library carries a 4 KB guard band on both sides; returns the
 * number of band bytes overwritten so far (0 = no write overrun), -1 when the library runs without SVO_GUARD. */
long long svo_guard_check(void);
/* Debug: deliberately overruns a scratch buffer by 3 bytes and returns what svo_guard_check() reports for it (3; -1 without SVO_GUARD). */
long long svo_guard_selftest(void);
/* Run all work of this context on the caller's CUDA stream (a cudaStream_t); NULL restores the context's own. */
int svo_ctx_set_stream(svo_ctx* ctx, void* cuda_stream);
int svo_ctx_synchronize(svo_ctx* ctx);
/* Kernels launched by this context since creation (bench.py's gpu_launches claim). */
int64_t svo_ctx_launch_count(const svo_ctx* ctx);
/* Device time of pipeline stages, measured with CUDA events on the context's stream while
 * svo_ctx_enable_stage_timing(ctx, 1).  Events are kept per timing slot (svo_ctx_set_timing_slot, 0..SVO_TIMING_SLOTS-1)
 * so that a bench can time K back-to-back steps without synchronising in between and read all of them afterwards.
 * stage: 0 pyramid(K1) 1 fast+cells(K2) 2 compact(K2b) 3 octree(K3) 4 angle+descriptor(K4/K5) 5 match(K6)
 *        6 fullframe-fast(K9) 7 lk-pyramid(K7) 8..11 lk-track calls 1..4 (K8) 12 pose (K13, both lists).
 * svo_ctx_stage_ms returns milliseconds, <0 if the stage did not run in that slot. */
int svo_ctx_enable_stage_timing(svo_ctx* ctx, int on);
int svo_ctx_set_timing_slot(svo_ctx* ctx, int slot);
float svo_ctx_stage_ms(const svo_ctx* ctx, int slot, int stage);
#define SVO_NUM_STAGES 13
#define SVO_TIMING_SLOTS 64

/* ---- ORB extractor ------------------------------------------------------------------------------------------- */
/* Replaces ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (include/lzb_vio/ORBextractor.h:33-34, src/ORBextractor.cpp:359-418): scale tables, per-level quotas, pattern, umax. */
int svo_orb_configure(svo_ctx* ctx, int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST);
/* Getters behind GetScaleFactors()/GetInverseScaleFactors()/GetScaleSigmaSquares()/GetInverseScaleSigmaSquares()
 * (ORBextractor.h:45-73) and mnFeaturesPerLevel; each array has nlevels entries, any pointer may be NULL. */
int svo_orb_tables(const svo_ctx* ctx, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                   int32_t* features_per_level);
/* Upper bound on keypoints returned per image: sum over levels of max(N_l + 3, 4 * nIni). */
int svo_orb_capacity(const svo_ctx* ctx, int w, int h);
/* Level geometry for a w x h input (ORBextractor::ComputePyramid, ORBextractor.cpp:1065-1066). */
int svo_orb_level_size(const svo_ctx* ctx, int w, int h, int level, int* lw, int* lh);

/* Note on the octree (ORBextractor.cpp:624-689): the reference breaks ties between equal-count nodes by heap ADDRESS; this library
 * uses creation order, which is what the reference does under an allocator that hands out increasing addresses.  Under glibc
 * malloc 1-3 % of the reference's own selected keypoints vary with the heap's history.
 *
 * Replaces ORBextractor::operator()(image, mask, keypoints, descriptors) (ORBextractor.h:41-43,
 * ORBextractor.cpp:990-1055).  img: 8-bit single channel, w x h, row stride in bytes.  kps/desc: caller arrays of
 * `cap` keypoints / cap*32 bytes; *n_out = number written.  An empty image (w==0||h==0||!img) returns SVO_OK with
 * *n_out = 0 (the reference returns silently, :993-994). */
int svo_orb_extract(svo_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, svo_keypoint* kps, uint8_t* desc,
                    int cap, int* n_out);
/* n_imgs images of identical size in one pass (image i -> kps + i*cap_per_img, desc + i*cap_per_img*32, n_out[i]). */
int svo_orb_extract_batch(svo_ctx* ctx, const uint8_t* const* imgs, int n_imgs, int w, int h, size_t stride,
                          svo_keypoint* kps, uint8_t* desc, int cap_per_img, int* n_out);
/* Read back pyramid level `level` of image `img_slot` of the last extract call, without border (public member
 * mvImagePyramid, ORBextractor.h:75; also used by the parity tests).  dst may be NULL to query the size. */
int svo_orb_read_level(svo_ctx* ctx, int img_slot, int level, uint8_t* dst, size_t dst_stride, int* w, int* h);
/* Read back the output of the per-cell FAST loop (vToDistributeKeys, ORBextractor.cpp:742-781) for one level, in the
 * reference's order, coordinates relative to minBorder: x, y, response filled; size=7, angle=-1, octave=0. */
int svo_orb_read_candidates(svo_ctx* ctx, int img_slot, int level, svo_keypoint* out, int cap, int* n_out);

/* ---- full-frame FAST ----------------------------------------------------------------------------------------- */
/* Replaces cv::FAST(image, keypoints, threshold, nonmaxSuppression) at tracking.cpp:101 (Detect_OpenCVFASTFeatures).
 * Output in row-major order: x, y, size=7, angle=-1, response=score, octave=0, class_id=-1. nms must be non-zero. */
int svo_fast_detect(svo_ctx* ctx, const uint8_t* img, int w, int h, size_t stride, int threshold, int nms,
                    svo_keypoint* kps, int cap, int* n_out);

/* ---- brute-force Hamming ------------------------------------------------------------------------------------- */
/* Replaces DescriptorMatcher::create("BruteForce-Hamming")->match(query, train, matches) (tracking.cpp:539-544):
 * for query i, train_idx[i] = lowest index with minimum popcount distance, dist[i] = that distance (DMatch.distance
 * as an integer).  nt == 0 gives train_idx = -1 (OpenCV returns no matches for an empty train set). */
int svo_match_hamming(svo_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* train_idx,
                      int32_t* dist);

/* ---- pyramidal Lucas-Kanade ---------------------------------------------------------------------------------- */
/* Replaces cv::calcOpticalFlowPyrLK(prev, next, prevPts, nextPts, status, err, Size(win,win), max_level,
 * TermCriteria(COUNT+EPS, max_iter, eps), 0, min_eig) (tracking.cpp:593,600,607,613).  win must be 21.
 * prev_xy / next_xy: n x 2 floats; status: n bytes; err: n floats or NULL. */
int svo_lk_track(svo_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t stride,
                 const float* prev_xy, int n, float* next_xy, uint8_t* status, float* err, int win, int max_level,
                 int max_iter, double eps, double min_eig);

/* ---- the Tracking post-filters (device; the same kernels the throughput mode runs) ----------------------------------------- */
/* The tail of Tracking::ORB_Robust_Find_MuliImage_MatchedFeatures after the two match() calls (tracking.cpp:546-577): of the first
 * des_index = min(n_last_left, n_last_right, n_cur_left) queries (quirk Q3) those with both distances <= max(2 * min_dist, 30) and
 * |y_left - y_right| < feature_match_error are accepted, in query order.  idx1/dist1: match(last_left, last_right), idx2/dist2:
 * match(last_left, cur_left).  Outputs: cap x 2 floats each (what the reference push_backs), *n_out accepted matches. */
int svo_orb_match_filter(svo_ctx* ctx, const svo_keypoint* last_left, int n_last_left, const svo_keypoint* last_right, int n_last_right,
                         const svo_keypoint* cur_left, int n_cur_left, const int32_t* idx1, const int32_t* dist1, const int32_t* idx2,
                         const int32_t* dist2, double feature_match_error, float* t2_left, float* t1_left, float* t1_right, int cap,
                         int* n_out);
/* Tracking::deleteBadmatchFeatures (tracking.cpp:623-660): a track is dropped if one of its five points has a negative coordinate,
 * one of its four status bytes is 0, or |p0.y - p1.y| or |p2.y - p3.y| exceeds feature_match_error.  The five n x 2 lists are
 * compacted IN PLACE (order kept), *n_out survivors. */
int svo_lk_chain_filter(svo_ctx* ctx, int n, float* p0, float* p1, float* p2, float* p3, float* p0_return, const uint8_t* status0,
                        const uint8_t* status1, const uint8_t* status2, const uint8_t* status3, double feature_match_error, int* n_out);

/* ---- triangulation ------------------------------------------------------------------------------------------- */
/* Replaces cv::triangulatePoints(P1, P2, pts1, pts2, points4D) + cv::convertPointsFromHomogeneous (tracking.cpp:190-192,
 * 292-294).  proj1/proj2: 3x4 row-major doubles; pts: n x 2 floats; points4d (may be NULL): n x 4 floats (column i of
 * OpenCV's 4 x n result); points3d: n x 3 floats.  Agreement with OpenCV: sign of the 4-vector is arbitrary, 3-D points
 * match to float rounding. */
int svo_triangulate_points(svo_ctx* ctx, const double* proj1, const double* proj2, const float* pts1, const float* pts2,
                           int n, float* points4d, float* points3d);

/* ---- pose from 3-D / 2-D correspondences ---------------------------------------------------------------------- */
/* Replaces cv::solvePnPRansac(points3D, points2D, K, zeros(4), rvec, tvec, useExtrinsicGuess = true, iterations, reproj_error,
 * confidence, inliers, SOLVEPNP_ITERATIVE) + cv::Rodrigues(rvec, rotation) at tracking.cpp:486-489 (OpenCV_EstimatePose_PnP), for
 * n_problems independent lists in ONE launch (one CTA per list): 5-point EPnP hypotheses drawn with cv::RNG((uint64)-1), consensus
 * at reproj_error pixels, RANSACUpdateNumIters, Levenberg-Marquardt refinement over the consensus set.  K: 3 x 3 row-major.
 * Problem i: counts[i] points at pts3d + i*cap*3 / pts2d + i*cap*2.  Outputs per problem: rvec[3], tvec[3], rotation[9] (row-major,
 * may be NULL), inlier_mask[cap] bytes (may be NULL; cv's `inliers` = the indices whose byte is 1), n_inliers (may be NULL),
 * info[2] = {index of the winning hypothesis, hypotheses evaluated} (may be NULL), ok = solvePnPRansac's return value.
 * ok = 0 (no model with more than 4 inliers, or fewer than 5 points): rvec = tvec = 0, rotation = identity, empty inlier set.
 * Agreement with OpenCV: DESIGN.md section 5 (identical inlier sets and poses whenever the winning 5-point hypothesis is one whose
 * EPnP solution does not depend on the basis OpenCV's SVD picks in an exactly singular subspace; else a close consensus set). */
int svo_pnp_ransac(svo_ctx* ctx, const double* K, const float* pts3d, const float* pts2d, const int32_t* counts, int n_problems, int cap,
                   int iterations, float reproj_error, double confidence, double* rvec, double* tvec, double* rotation,
                   uint8_t* inlier_mask, int32_t* n_inliers, int32_t* info, int32_t* ok);

/* ---- throughput mode: batches of consecutive stereo frames ------------------------------------------------------ */
/* One unit of work = one new stereo frame t:
 *   SVO_MODE_ORB: operator() on left_t and right_t (Detect_MyORBFeatures, tracking.cpp:502-532), then
 *                 match(left_{t-1}, right_{t-1}) and match(left_{t-1}, left_t) (tracking.cpp:543-544);
 *   SVO_MODE_LK : cv::FAST(left_t, 20, true) (tracking.cpp:101) and the four chained LK calls over frames t-1, t
 *                 seeded with the FAST corners of left_{t-1} (tracking.cpp:268-271, 593-618).
 * The context carries frame t-1 between calls (Tracking::last_frame_); after svo_frontend_reset the first frame only
 * detects (Tracking::StereoInit_f2f, tracking.cpp:78-92) and reports zero queries / seeds. */
#define SVO_MODE_ORB 1
#define SVO_MODE_LK 2

typedef struct svo_frontend_out {
    /* capacities, filled by the caller */
    int32_t kp_cap; /* keypoints per image  (>= svo_orb_capacity)          */
    int32_t pt_cap; /* FAST corners / LK seeds per frame                    */
    /* ORB mode; frame f uses [f*kp_cap, (f+1)*kp_cap).  Any pointer may be NULL (that output stays on the device). */
    svo_keypoint* kps_left;
    uint8_t* desc_left;
    int32_t* n_left;
    svo_keypoint* kps_right;
    uint8_t* desc_right;
    int32_t* n_right;
    int32_t* n_query;       /* per frame: #queries = n_left of frame f-1 (0 for the first frame after reset) */
    int32_t* stereo_idx;    /* match(left_{f-1}, right_{f-1}).trainIdx,  kp_cap per frame */
    int32_t* stereo_dist;
    int32_t* temporal_idx;  /* match(left_{f-1}, left_f).trainIdx */
    int32_t* temporal_dist;
    /* LK mode; frame f uses [f*pt_cap, ...) */
    svo_keypoint* fast_kps; /* FAST corners of left_f */
    int32_t* n_fast;
    int32_t* n_seed;        /* per frame: #LK seeds = n_fast of frame f-1 */
    float* lk_p1;           /* t-1 right  (call 1: left_{t-1} -> right_{t-1}),  pt_cap*2 floats per frame */
    float* lk_p2;           /* t   right  (call 2) */
    float* lk_p3;           /* t   left   (call 3) */
    float* lk_p0r;          /* t-1 left, returned (call 4) */
    uint8_t* lk_status;     /* 4 x pt_cap per frame: status1..status4 */
    /* device-side post-filters: what the reference hands to triangulation / PnP */
    float* orb_t2_left;     /* ORB_Robust_Find_MuliImage_MatchedFeatures outputs (tracking.cpp:534-581), kp_cap*2 floats per frame */
    float* orb_t1_left;
    float* orb_t1_right;
    int32_t* n_orb_matched; /* per frame: number of accepted matches */
    float* lkf_t1_left;     /* LK_Robust_Find_MuliImage_MatchedFeatures outputs after deleteBadmatchFeatures (tracking.cpp:583-660) */
    float* lkf_t1_right;
    float* lkf_t2_right;
    float* lkf_t2_left;
    int32_t* n_lk_matched;  /* per frame: number of surviving tracks */
    /* cv::triangulatePoints + convertPointsFromHomogeneous of (t1_left, t1_right) (tracking.cpp:190-192, 292-294); filled only
     * after svo_frontend_set_projection(); 3 floats per point */
    float* orb_points3d;    /* kp_cap*3 floats per frame, n_orb_matched valid */
    float* lk_points3d;     /* pt_cap*3 floats per frame, n_lk_matched valid */
    /* the `err` output of the four calcOpticalFlowPyrLK calls (tracking.cpp:593-618 pass it; unused by the reference afterwards):
     * 4 x pt_cap floats per frame (err1..err4), computed unless svo_frontend_set_lk_err(ctx, 0) */
    float* lk_err;
    /* pose of frame f relative to frame f-1 from each filtered list (OpenCV_EstimatePose_PnP, tracking.cpp:464-501), filled only
     * after svo_frontend_set_pnp(ctx, 1, ...): 16 doubles per frame = rvec[3], tvec[3], rotation[9] row-major, 1 spare;
     * 4 ints per frame = {ok, inliers, winning hypothesis, hypotheses evaluated}; one byte per list entry = inlier flag */
    double* orb_pose;
    int32_t* orb_pnp_info;
    uint8_t* orb_inlier_mask;  /* kp_cap per frame, n_orb_matched valid */
    double* lk_pose;
    int32_t* lk_pnp_info;
    uint8_t* lk_inlier_mask;   /* pt_cap per frame, n_lk_matched valid */
} svo_frontend_out;

int svo_frontend_configure(svo_ctx* ctx, int w, int h, int mode, int max_frames);
int svo_frontend_reset(svo_ctx* ctx);
/* Tracking::feature_match_error_ (parameter.cpp:53, config/default.yaml:66; default 3) used by the device-side post-filters */
int svo_frontend_set_match_error(svo_ctx* ctx, double feature_match_error);
/* Both modes in one batch: run the ORB tail (octree, orientation + descriptors, the two Hamming matches, the association filter)
 * on a side stream concurrently with the FAST + LK chain (they share no data; the big kernels still mostly serialise, the small
 * ones fill each other's tails: ~1 % at 1241x376).  On by default; 0 = one stream. */
int svo_frontend_set_overlap(svo_ctx* ctx, int enable);
/* The reference passes `err` to every calcOpticalFlowPyrLK call, so OpenCV computes it; the LK chain does the same by default
 * (svo_frontend_out.lk_err).  0 skips that work (the reference never reads err). */
int svo_frontend_set_lk_err(svo_ctx* ctx, int enable);
/* Sensors::projMatr1_ / projMatr2_ (3x4, row-major doubles; sensors.cpp:8-24): enables the device-side triangulation of
 * the filtered lists.  Call after svo_frontend_configure(). */
int svo_frontend_set_projection(svo_ctx* ctx, const double* proj_left, const double* proj_right);
/* Tracking::iterationsCount_ / reprojectionError_ / confidence_ (parameter.cpp:57-59; config/default.yaml:80-82: 500, 0.5, 0.99):
 * enables the device-side pose step on the triangulated lists (K = left 3 x 3 of proj_left, tracking.cpp:476-477).  Call after
 * svo_frontend_set_projection(); a later svo_frontend_configure() switches it off again.  enable = 0 frees its buffers. */
int svo_frontend_set_pnp(svo_ctx* ctx, int enable, int iterations, float reproj_error, double confidence);
/* host images -> device (left/right: n_frames contiguous images of h rows, `stride` bytes per row) */
int svo_frontend_upload(svo_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride);
/* blocks until the host-to-device copies of the last svo_frontend_upload() have completed (the host buffers may be reused) */
int svo_frontend_wait_upload(svo_ctx* ctx);
/* same, but left/right are DEVICE pointers (images already resident in HBM, e.g. written by a decoder or a torch tensor) */
int svo_frontend_upload_dev(svo_ctx* ctx, const uint8_t* left_dev, const uint8_t* right_dev, int n_frames, size_t stride);
/* all kernels for the uploaded frames; inputs and outputs stay in HBM */
int svo_frontend_run(svo_ctx* ctx);
/* device -> host for every non-NULL field of out, for the OLDEST run not yet delivered (runs are delivered in order; with
 * max_images >= 2 * (2 * max_frames + 1) one more run may be in flight while this one is downloaded on its own stream).
 * The run leaves the queue only when it has been delivered; SVO_ERR_CAPACITY because a caller capacity (kp_cap / pt_cap) is too
 * small writes nothing and keeps the run queued, so the call can be repeated with larger buffers (or svo_frontend_discard).
 * A DEVICE capacity overflow (SVO_ERR_CAPACITY, "device capacity exceeded") retires the run: its results are incomplete.
 * out == NULL retires the run without copying anything (device-resident pipelines); such a caller MUST call
 * svo_frontend_check() before trusting the device-side results, because nothing else reports a capacity overflow. */
int svo_frontend_download(svo_ctx* ctx, svo_frontend_out* out);
/* SVO_ERR_CAPACITY if any run retired with svo_frontend_download(ctx, NULL) since the last check overflowed a device capacity
 * (one 4-byte read-back and a stream synchronisation). */
int svo_frontend_check(svo_ctx* ctx);
/* drop the oldest undelivered run without reading anything */
int svo_frontend_discard(svo_ctx* ctx);
/* upload + run + download */
int svo_frontend_batch(svo_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_frames, size_t stride,
                       svo_frontend_out* out);

/* ---- frame source (host code; replaces System::NextFrame_kitti, System.cpp:75-104) ----------------------------------- */
/* Reads <dir>/image_0/%06d.png (left) and <dir>/image_1/%06d.png (right) from index first_index on, like the reference's
 * cv::imread(..., IMREAD_GRAYSCALE); the run ends at the first missing pair (the reference returns a null frame there).
 * 8-bit greyscale non-interlaced PNG (the KITTI odometry grey sequences) and binary PGM (%06d.pgm) are decoded; any other
 * format or a size other than w x h is an error.  `threads` decoder threads fill a ring of `ring_batches` (>= 3) batches of
 * `batch` frames in PINNED host memory ahead of the consumer. */
typedef struct svo_source svo_source;
int svo_source_open(svo_source** src, const char* dir, int first_index, int w, int h, int batch, int ring_batches, int threads);
/* Number of consecutive stereo pairs found from first_index on. */
int svo_source_frame_count(const svo_source* src);
/* 1 if the ring is pinned (page-locked) memory; 0 on a host without a CUDA device (ordinary memory: decoder tests / tools). */
int svo_source_is_pinned(const svo_source* src);
/* Next batch in sequence order: *left / *right point at *n_frames contiguous w x h images (stride w).  A batch stays valid until
 * the SECOND following call: svo_frontend_upload() copies asynchronously, so the batch must outlive the request for the next one;
 * a caller that has downloaded (or svo_frontend_wait_upload-ed) batch k before it asks for batch k + 2 is safe -- the pipelined
 * loop upload(k+1), run(k+1), download(k), next() is.  *n_frames == 0: end of the run.
 * Blocks until the batch is decoded.  SVO_ERR_INVALID on a decode error (svo_source_error). */
int svo_source_next(svo_source* src, const uint8_t** left, const uint8_t** right, int* n_frames);
const char* svo_source_error(const svo_source* src);
void svo_source_close(svo_source* src);
/* Writers for tests and benchmarks (synthetic sequences on disk in the reference's dataset layout). */
int svo_write_pgm(const char* path, const uint8_t* img, int w, int h, size_t stride);
int svo_write_png(const char* path, const uint8_t* img, int w, int h, size_t stride, int zlib_level);

/* ---- deterministic synthetic stereo sequences (bench / tests input; host code, needs no GPU) --------------------- */
/* Frame `frame` of sequence `seed`: textured canvas, 8-px disparity, 2-px/frame motion, canvas
 * re-seeded every 16 frames.  left/right: h rows of `stride` bytes.  Lives in its own host-only library, libsvo_synth.so
 * (no CUDA in it), so that the CPU reference arm of bench.py never maps the CUDA library. */
int svo_synth_stereo_frame(uint64_t seed, int frame, int w, int h, uint8_t* left, uint8_t* right, size_t stride);

#ifdef __cplusplus
}
#endif
#endif /* SVO_CUDA_H */
