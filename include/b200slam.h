/* b200slam.h - C ABI of libb200slam.so: the B200-native (sm_100a) ALIKED + LightGlue
 * frontend that stands behind SimpleSLAM's `slam/core/features_utils.py`.
 *
 * The reference has no FFI layer: its seam is the Python adapter
 * `/root/reference/slam/core/features_utils.py`, which constructs and calls the
 * un-vendored `lightglue` package.  Each entry point below names the reference call it
 * replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *  - every function returns 0 on success or a negative B2S_E* code; the message is
 *    available from b2s_last_error() (thread local).  Nothing throws or aborts.
 *  - "_dev" pointers are device pointers on the handle's device; the caller owns all I/O
 *    buffers; work is enqueued on `stream` (a cudaStream_t passed as void*); the
 *    library owns weights and workspaces.  Functions whose name ends in `_host` take host
 *    pointers, copy in/out and synchronise before returning.
 *  - a handle is bound to one device and is not thread-safe (one handle per device+stream).
 *  - there is NO CPU fallback: without a CUDA device every create call fails.
 */
#ifndef B200SLAM_H_
#define B200SLAM_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_VERSION 100

enum {
  B2S_OK = 0,
  B2S_EINVAL = -1,   /* bad argument / malformed weight blob */
  B2S_ECUDA = -2,    /* CUDA runtime error (message has the cudaError string) */
  B2S_ENOMEM = -3,
  B2S_ENODEV = -4,   /* no usable sm_100 device */
  B2S_ESIZE = -5,    /* input larger than the handle supports */
  B2S_ERANGE = -6    /* ALIKED: an activation left the fp16 range of the two-plane convolutions (re-create with B2S_ALIKED_CONV_NP=3) */
};

/* Arithmetic of the LightGlue matcher.
 *   B2S_FP32      fp32-faithful on the tensor cores, fast form: every fp32 operand is carried as TWO fp16 planes
 *                 x = h0 + 2^-11 h1 (22 significant bits at any magnitude) and each contraction issues the three
 *                 significant cross products into fp32 TMEM accumulators (operand error below the fp32 rounding noise
 *                 of the accumulation; identical match sets vs an fp32 reference in every parity test).  fp16 limits
 *                 the operand range to |x| < 65504: every producer checks it on the device, and a launch sequence
 *                 that left the range is re-run on the B2S_FP32X3 engine (b2s_lg_fallback; the host entry points do
 *                 so themselves, device entry points report n_matches = B2S_LG_RANGE);
 *   B2S_FP32X3    fp32-faithful, any range: three bf16 planes, six cross products (twice the tensor-core work);
 *   B2S_BF16      operands rounded once to bf16 (fastest; >= 99 % match-set agreement).
 * The input projection and the assignment head are fp32-faithful in every mode. */
enum { B2S_FP32 = 0, B2S_BF16 = 1, B2S_FP32X3 = 2 };
enum { B2S_LG_RANGE = -2 };   /* n_matches of a pair whose launch sequence left the fp16 operand range (B2S_FP32) */
enum { B2S_IMG_BGR_U8_HWC = 0, B2S_IMG_RGB_F32_CHW = 1 };

typedef struct b2s_aliked b2s_aliked;
typedef struct b2s_lg b2s_lg;

/* replaces the keyword arguments of `ALIKED(max_num_keypoints=...)`, features_utils.py:25,
 * plus upstream's default_conf (model_name, detection_threshold, nms_radius, resize). */
typedef struct {
  int model;            /* 0 = aliked-n16 (M=16), 1 = aliked-n32 (M=32) */
  int max_kp;           /* n_limit; <=0 means upstream's 20000 */
  float det_thresh;     /* 0.2 */
  int nms_radius;       /* 2 (only 2 is supported) */
  int resize_long;      /* 1024; <=0 disables resizing */
  int precision;        /* B2S_FP32 | B2S_BF16 */
} b2s_aliked_cfg;

/* replaces `LightGlue(features='aliked')`, features_utils.py:26, plus upstream's default_conf. */
typedef struct {
  int n_layers;         /* 9 */
  int heads;            /* 4 */
  int dim;              /* 256 */
  int in_dim;           /* 128 */
  float depth_conf;     /* 0.95; <=0 disables early exit */
  float width_conf;     /* 0.99; <=0 disables point pruning */
  float filter_thresh;  /* 0.1 */
  int pruning_min_kpts; /* prune a side only while it has more points than this; -1 = the
                           reference's CPU setting (always), 1536 = its CUDA+flash setting */
  int precision;        /* B2S_FP32 | B2S_BF16 */
  int max_kp;           /* initial workspace size per image (grows on demand) */
} b2s_lg_cfg;

int b2s_version(void);
const char* b2s_last_error(void);
int b2s_device_count(void);

void b2s_aliked_default_cfg(b2s_aliked_cfg* cfg);
void b2s_lg_default_cfg(b2s_lg_cfg* cfg);

/* weights: flat "B2SW" blob of fp32 tensors named as in the upstream state dict
 * (layout in opencv-simpleslam_b200/weights.py::pack_state). */
int b2s_aliked_create(const b2s_aliked_cfg* cfg, const void* weights, size_t nbytes, int device,
                      b2s_aliked** out);
void b2s_aliked_destroy(b2s_aliked* h);

/* replaces `detector.extract(t0)` (features_utils.py:94) and, for B2S_IMG_BGR_U8_HWC,
 * also `_bgr_to_tensor` (features_utils.py:219-222).
 *   img_dev    : device image; u8 BGR HWC with `row_stride` bytes per row, or f32 RGB CHW
 *                planar in [0,1] (row_stride ignored)
 *   kpts_dev   : [max_kp,2] f32 keypoints in ORIGINAL-image pixels (x,y)
 *   desc_dev   : [max_kp,128] f32 unit-norm descriptors
 *   scores_dev : [max_kp] f32 upstream "keypoint_scores" (nullable)
 *   n_out_dev  : device int32, number of valid rows  */
int b2s_aliked_extract(b2s_aliked* h, const void* img_dev, int img_format, int H, int W,
                       int row_stride, void* stream, float* kpts_dev, float* desc_dev,
                       float* scores_dev, int32_t* n_out_dev);

/* Batched extraction of B same-shape frames (imgs_dev: HOST array of B device image pointers).  The frames are spread
 * over n_lanes (<= 8; 0 = 4) internal extractor lanes - own workspace and stream each - that run concurrently; `stream`
 * forks into the lanes and joins them, nothing synchronises with the host.  Outputs are slabs:
 * kpts_dev [B, max_kp, 2], desc_dev [B, max_kp, 128], scores_dev (nullable) [B, max_kp], n_out_dev [B]. */
int b2s_aliked_extract_batch(b2s_aliked* h, const void* const* imgs_dev, int B, int img_format, int H, int W,
                             int row_stride, int n_lanes, void* stream, float* kpts_dev, float* desc_dev,
                             float* scores_dev, int32_t* n_out_dev);
/* eps > 0: b2s_aliked_extract_batch also applies `des /= (||des||_2 + eps)` (features_utils.py:100, eps = 1e-8) on the
 * device, as b2s_aliked_extract_host_ex does for single frames; 0 (default) switches it off. */
int b2s_aliked_set_batch_renorm(b2s_aliked* h, float eps);

/* same, host buffers; copies in/out, synchronises, returns the count in *n_out. */
int b2s_aliked_extract_host(b2s_aliked* h, const void* img_host, int img_format, int H, int W,
                            int row_stride, float* kpts_host, float* desc_host,
                            float* scores_host, int32_t* n_out);

/* same, and additionally applies the reference caller's re-normalisation `des0 /= (||des0||_2 + eps)`
 * (features_utils.py:100) on the device when desc_renorm_eps > 0 (the reference uses 1e-8). */
int b2s_aliked_extract_host_ex(b2s_aliked* h, const void* img_host, int img_format, int H, int W,
                               int row_stride, float* kpts_host, float* desc_host,
                               float* scores_host, int32_t* n_out, float desc_renorm_eps);

/* Split form of b2s_aliked_extract_host_ex for callers that build per-keypoint host objects (cv2.KeyPoint lists,
 * features_utils.py:61-63): _begin uploads the frame and enqueues the whole extraction; _keypoints blocks only until
 * the detector has finished (the descriptor head is still running) and returns keypoints + count; _finish waits for
 * the descriptors (n rows) and scores.  One extraction may be pending per handle. */
int b2s_aliked_extract_host_begin(b2s_aliked* h, const void* img_host, int img_format, int H, int W,
                                  int row_stride, float desc_renorm_eps);
int b2s_aliked_extract_host_keypoints(b2s_aliked* h, float* kpts_host, int32_t* n_out);
int b2s_aliked_extract_host_finish(b2s_aliked* h, float* desc_host, float* scores_host);

/* Device-side copy of the last *_host extraction's n keypoints / descriptors (exactly what the caller received) into
 * caller-owned device buffers: a host-API caller can keep a frame's features on the GPU and skip their re-upload when the
 * frame is matched next (the drop-in's feature cache, opencv-simpleslam_b200/features_utils.py). */
int b2s_aliked_copy_last_features(b2s_aliked* h, float* kpts_dst_dev, float* desc_dst_dev, int n, void* stream);

int b2s_lightglue_create(const b2s_lg_cfg* cfg, const void* weights, size_t nbytes, int device,
                         b2s_lg** out);
void b2s_lg_destroy(b2s_lg* h);

/* replaces `matcher({'image0':{...}, 'image1':{...}})` (features_utils.py:157-161, :237).
 *   k0/k1 : [m,2]/[n,2] f32 keypoints (pixels); d0/d1 : [m,128]/[n,128] f32 descriptors
 *   size0/size1 : HOST pointers to (W,H) or NULL -> upstream's bounding-extent normalisation
 *                 (NULL is what the reference's split path does, features_utils.py:158-161)
 *   matches_dev  : [min(m,n),2] int32 (i in image0, j in image1), ascending i
 *   mscores_dev  : [min(m,n)] f32
 *   n_matches_dev: device int32
 *   optional (nullable): matches0_dev [m] / matches1_dev [n] int32 (-1 = unmatched),
 *   ms0_dev [m] / ms1_dev [n] f32, prune0_dev [m] / prune1_dev [n] int32.
 *   stop_layer_dev (device int32, nullable) receives the number of layers executed (upstream 'stop').
 * The whole match is enqueued on `stream` without any host synchronisation: the decisions upstream
 * takes on the host after every layer (early exit, point pruning) are taken on the device, and the
 * kernels of layers behind an exit return at once.  (A stream sync happens only when the workspace
 * has to grow because max(m,n) exceeds the handle's capacity.) */
int b2s_lightglue_match(b2s_lg* h, const float* k0_dev, const float* d0_dev, int m,
                        const float* k1_dev, const float* d1_dev, int n, const float* size0,
                        const float* size1, void* stream, int32_t* matches_dev,
                        float* mscores_dev, int32_t* n_matches_dev, int32_t* stop_layer_dev,
                        int32_t* matches0_dev, int32_t* matches1_dev, float* ms0_dev,
                        float* ms1_dev, int32_t* prune0_dev, int32_t* prune1_dev);

int b2s_lightglue_match_host(b2s_lg* h, const float* k0, const float* d0, int m, const float* k1,
                             const float* d1, int n, const float* size0, const float* size1,
                             int32_t* matches, float* mscores, int32_t* n_matches,
                             int32_t* stop_layer, int32_t* matches0, int32_t* matches1,
                             float* ms0, float* ms1, int32_t* prune0, int32_t* prune1);

/* Batched matching of P independent pairs: the keyframe-window all-pairs loops of the reference
 * (`select_keyframe`, keyframe_utils.py:153; `triangulate_between_kfs_2view`, triangulation_utils.py:131 - BASELINE
 * config 3) or the consecutive pairs of a frame stream.  The pair is a grid dimension of every kernel: ONE launch
 * sequence serves up to b2s_lg_max_batch() pairs (larger batches are processed in such chunks), each pair with its own
 * device-resident sizes, early-exit flag and pruning state; results are identical to P calls of b2s_lightglue_match
 * with size0 = size1 = NULL.
 * Keypoints/descriptors of F frames are packed; frame f owns rows [cu[f], cu[f+1]).
 * pair p matches frame pair_i[p] against pair_j[p]; outputs for pair p start at row
 * p*stride of matches_dev/mscores_dev, counts in n_matches_dev[p]. Host arrays: cu,
 * pair_i, pair_j.  _ex: counts (host, nullable) [F]: frame f owns rows [cu[f], cu[f] + counts[f]) - frames in fixed-size
 * slots, as an all_gather of per-rank records leaves them; counts_dev (DEVICE, nullable) [F]: the true count of frame f is
 * min(host count, counts_dev[f]) read on the device - the extractor's n_out_dev goes straight in and a frame stream never
 * waits for a count to reach the host (the host count is then just the slot capacity); max_batch > 0 lowers the pairs
 * per launch sequence; stop_layers_dev (nullable) [P]. */
int b2s_lightglue_match_batch(b2s_lg* h, const float* kpts_dev, const float* desc_dev,
                              const int32_t* cu, int n_frames, const int32_t* pair_i,
                              const int32_t* pair_j, int n_pairs, void* stream, int stride,
                              int32_t* matches_dev, float* mscores_dev, int32_t* n_matches_dev);
int b2s_lightglue_match_batch_ex(b2s_lg* h, const float* kpts_dev, const float* desc_dev,
                                 const int32_t* cu, const int32_t* counts, const int32_t* counts_dev, int n_frames,
                                 const int32_t* pair_i,
                                 const int32_t* pair_j, int n_pairs, void* stream, int stride, int max_batch,
                                 int32_t* matches_dev, float* mscores_dev, int32_t* n_matches_dev,
                                 int32_t* stop_layers_dev);
int b2s_lg_max_batch(void);
/* Device workspace of a matcher handle for `pairs` pairs per launch sequence of up to max_kp keypoints per image
 * (bytes; about 110 MB per pair at 2048 keypoints in B2S_FP32), and the call that allocates it ahead of time (otherwise
 * the workspace grows on demand, which synchronises the stream). */
size_t b2s_lg_workspace_bytes(const b2s_lg_cfg* cfg, int max_kp, int pairs);
int b2s_lg_reserve(b2s_lg* h, int max_kp, int pairs);
/* B2S_FP32: the B2S_FP32X3 matcher that re-runs pairs reported as B2S_LG_RANGE (created on first use; *out = NULL for the
 * other precisions), how often it was asked for, and the operand planes of a handle (1 bf16, 2 fp16x2, 3 bf16x3). */
int b2s_lg_fallback(b2s_lg* h, b2s_lg** out);
long long b2s_lg_range_fallbacks(const b2s_lg* h);
int b2s_lg_planes(const b2s_lg* h);

/* ---------------------------------------------------------------------------------------------
 * Rows behind the matcher: epipolar outlier rejection and frame ingest.
 * --------------------------------------------------------------------------------------------- */
typedef struct b2s_fm b2s_fm;
typedef struct b2s_remap b2s_remap;

/* Fundamental-matrix RANSAC.  Replaces `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC, thresh, 0.99)` inside
 * `filter_matches_ransac` (features_utils.py:185-200; callers main_revamped.py:126, keyframe_utils.py:154,
 * triangulation_utils.py:132).  n_hyp minimal 7-point samples are drawn (splitmix64 streams of `seed`), solved and
 * scored in parallel with OpenCV's error (max of the two squared point-to-epipolar-line distances <= thresh^2); the
 * model with the largest consensus wins (lowest sample index on ties).  OpenCV's sequential loop stops after at most
 * 2000 samples; a fixed n_hyp >= 2000 therefore never looks at fewer hypotheses than cv2 would for the same data.
 *   pts1_dev / pts2_dev : [*,2] f32 pixel coordinates; pairs_dev (nullable) [n,2] int32 selects row pairs[i][0] of
 *                         pts1 and pairs[i][1] of pts2 (the matcher's `matches` output) - NULL means row i of both
 *   mask_dev   : [n] u8, 1 = inlier           F_dev : [9] f64 row-major, unit Frobenius norm (x2^T F x1 = 0)
 *   result_dev : [2] int32 = { inlier count, winning model index (sample*3 + root) or -1 if every sample was degenerate } */
int b2s_fm_create(int device, int max_points, int max_hypotheses, b2s_fm** out);
void b2s_fm_destroy(b2s_fm* h);
int b2s_fm_ransac(b2s_fm* h, const float* pts1_dev, const float* pts2_dev, const int32_t* pairs_dev, int n,
                  float thresh, int n_hyp, uint64_t seed, void* stream, uint8_t* mask_dev, double* F_dev,
                  int32_t* result_dev);
/* host buffers; copies in/out on the handle's own stream and synchronises */
int b2s_fm_ransac_host(b2s_fm* h, const float* pts1, const float* pts2, int n, float thresh, int n_hyp,
                       uint64_t seed, uint8_t* mask, double* F, int32_t* n_inliers, int32_t* model_index);
/* cv2-IDENTICAL form: the same mask and the same F (to rounding) as `cv2.findFundamentalMat(pts1, pts2, cv2.FM_RANSAC,
 * thresh, confidence)` (features_utils.py:195-196) for n >= 15 - below that OpenCV silently runs LMedS, whose winner for
 * n <= 13 is decided by the rounding noise of near-zero residuals and is left to cv2 itself by the Python drop-in.
 * OpenCV's loop (calib3d ptsetreg.cpp RANSACPointSetRegistrator::run) is reproduced, not approximated: the host draws
 * the iterations' subsets from cv::RNG((uint64)-1) ahead of the models (getSubset + haveCollinearPoints - the stream does
 * not depend on them; waves of 64 / 256 / the rest, so a high inlier ratio stops after the first wave), the device solves
 * every subset of a wave with fundam.cpp's run7Point (including the null-space basis
 * OpenCV's SVD returns, which fixes the order of the up-to-three models) and counts inliers with computeError's float32
 * errors, and the host replays the strictly-greater update and RANSACUpdateNumIters over the counts; a last kernel
 * writes the winner's mask.  max_iters <= the handle's max_hypotheses (cv2's default is 1000).
 *   mask [n] u8, F [9] f64 row-major (F[8] = 1), n_inliers; info4 = { winning iteration, model within it (-1: no
 *   model - cv2 returns None), final iteration bound, subsets drawn }.  Restated in oracle/cv_ransac.py (pinned against
 *   cv2), tests/test_gpu_geometry.py compares with cv2.findFundamentalMat directly. */
int b2s_fm_cv_ransac_host(b2s_fm* h, const float* pts1, const float* pts2, int n, double thresh, double confidence,
                          int max_iters, uint8_t* mask, double* F, int32_t* n_inliers, int32_t* info4);
/* test hook: candidate models (pixel coordinates, [3][9]), their number and consensus counts of one sample */
int b2s_fm_debug_models(b2s_fm* h, int hyp, double* models27, int32_t* n_models, int32_t* counts3);
long long b2s_fm_launch_count(const b2s_fm* h);

/* Frame ingest.  Replaces `cv2.remap(img, mapx, mapy, cv2.INTER_LINEAR)` on u8 BGR frames with the maps of
 * `cv2.initUndistortRectifyMap(..., cv2.CV_32FC1)` (main_revamped.py:313-315, :323-324).  Bit-exact with OpenCV's
 * fixed-point bilinear remap (5 fractional coordinate bits, 15-bit weights, constant 0 border).  The maps are uploaded
 * once at create; the undistorted frame can stay on the device and go straight into b2s_aliked_extract. */
int b2s_remap_create(int device, const float* mapx_host, const float* mapy_host, int dst_h, int dst_w, int src_h,
                     int src_w, b2s_remap** out);
void b2s_remap_destroy(b2s_remap* h);
int b2s_remap_bgr(b2s_remap* h, const uint8_t* src_dev, int src_stride, void* stream, uint8_t* dst_dev, int dst_stride);
int b2s_remap_bgr_host(b2s_remap* h, const uint8_t* src, int src_stride, uint8_t* dst, int dst_stride);
const uint8_t* b2s_remap_output_dev(const b2s_remap* h);
void b2s_remap_dims(const b2s_remap* h, int* dst_h, int* dst_w, int* src_h, int* src_w);
/* Attach (or detach with NULL) an ingest stage to an extractor: the *_host extraction entry points then take the raw
 * distorted frame, undistort it on the device and extract from the result - the reference's
 * `img = cv2.remap(img, mapx, mapy, INTER_LINEAR); feature_extractor(args, img, detector)` in one upload.  Keypoints are in
 * undistorted-image pixels.  The remap handle is borrowed, not owned. */
int b2s_aliked_set_undistort(b2s_aliked* h, b2s_remap* r);
long long b2s_remap_launch_count(const b2s_remap* h);

/* Projected-landmark window matching.  Replaces the body of `reproject_and_match_2d3d` (pnp_utils.py:224-304, called
 * once per tracked frame at main_revamped.py:460): project every map point with T_cw, keep those in front of the camera
 * and inside the image, collect the keypoints within radius_px of the projection (the reference's cKDTree ball query),
 * score each by the smallest descriptor distance to the point's most recent (<= max_obs, the reference checks 6)
 * observations, and give every point - in map order, as the reference's loop does - its best keypoint not taken by
 * an earlier point, if that distance is <= max_dist.
 *   Xw_dev [P,3] f64 | mp_desc_dev [rows,max_obs,128] f32 | mp_nobs_dev [rows] i32 (0: point skipped, as when the
 *   reference finds no descriptor) | mp_row_dev (nullable) [P] i32: table row of point i (NULL: row i) | K_host [9] f64 row-major | Tcw_host [16] f64 row-major | kps_dev [N,2] f32 | des_dev [N,128] f32
 *   distance = L2 between float descriptors (what the reference computes whatever `use_cosine` says: its
 *   _best_mp_distance_to_cur_desc always passes metric="auto", pnp_utils.py:121)
 *   kp_of_point_dev [P] i32: matched keypoint index or -1 | uv_dev (nullable) [P,2] f32 projections
 *   flags_dev (nullable) i32: bit 0 set if a window held more keypoints than the handle's cand_cap (result then unreliable) */
typedef struct b2s_reproj b2s_reproj;
int b2s_reproj_create(int device, int max_points, int max_kps, int cand_cap, b2s_reproj** out);
void b2s_reproj_destroy(b2s_reproj* h);
int b2s_reproj_match(b2s_reproj* h, const double* Xw_dev, const float* mp_desc_dev, const int32_t* mp_nobs_dev,
                     const int32_t* mp_row_dev, int P, int max_obs, const double* K_host, const double* Tcw_host, const float* kps_dev,
                     const float* des_dev, int N, int img_w, int img_h, double radius_px, double max_dist,
                     void* stream, int32_t* kp_of_point_dev, float* uv_dev, int32_t* flags_dev);
long long b2s_reproj_launch_count(const b2s_reproj* h);

/* Test hooks: copy a named intermediate of the most recent call to the host (fp32).
 * Returns the number of floats available in *n (and copies min(*n, cap)). */
int b2s_aliked_debug_get(b2s_aliked* h, const char* name, float* out, size_t cap, size_t* n);
int b2s_lg_set_debug(b2s_lg* h, int on);
int b2s_lg_debug_get(b2s_lg* h, const char* name, float* out, size_t cap, size_t* n);

/* Per-kernel-class timing with CUDA events recorded on the launching stream (used by bench.py
 * for roofline.achieved).  cls: 0 = attention kernel, 1 = GEMM kernels.  profile_read
 * synchronises on the recorded events, returns the summed kernel time (ms) and the number of
 * launches, and clears the records. */
int b2s_lg_profile(b2s_lg* h, int on);
int b2s_lg_profile_read(b2s_lg* h, int cls, double* ms, long long* n_launches);
/* Executed attention work since b2s_lg_profile(h, 1): the sum of nq*nk over every live self-attention problem
 * and over every live cross-attention problem (device-side counters: pruning / early exit are accounted for).
 * Synchronises the device. */
int b2s_lg_profile_work(b2s_lg* h, double* self_pairs, double* cross_pairs);

/* Unit-test entry points for the tcgen05 kernels (host buffers).  The plain names round the
 * operands to bf16; the "3" variants carry them as three bf16 planes (fp32-faithful):
 *   b2s_test_gemm_tc* : C[M,N] = A[M,K] W[N,K]^T + bias  (fp32 out; N, K multiples of 64)
 *   b2s_test_attn_tc* : ctx[nq,256] = softmax(q k^T / 8) v per head (4 heads x 64) */
int b2s_test_gemm_tc(const float* A, const float* W, const float* bias, int M, int N, int K, float* C);
int b2s_test_attn_tc(const float* q, const float* k, const float* v, int nq, int nk, float* ctx);
int b2s_test_gemm_tc3(const float* A, const float* W, const float* bias, int M, int N, int K, float* C);
int b2s_test_attn_tc3(const float* q, const float* k, const float* v, int nq, int nk, float* ctx);
/* device-only timing of one self-block attention launch (2 problems x 4 heads, nq x nk), mean ms per launch */
int b2s_bench_attn_tc(int nq, int nk, int iters, float* ms_out);
int b2s_bench_attn_tc3(int nq, int nk, int iters, float* ms_out);
/* same, plus SM-clock stamps of CTA (0,0,0) of one extra launch: trace_out [3 roles][64 tiles][8 events] (tools/trace_attn.py) */
int b2s_trace_attn_tc3(int nq, int nk, int iters, float* ms_out, long long* trace_out);

/* device-only timing of one bf16x3 layer-GEMM shape, `iters` launches back to back (PDL), clusters of `cl` CTAs sharing
 * the A tile (1 = none); ts_out (nullable) receives 6 globaltimer stamps (ns, relative) per CTA of one extra launch */
int b2s_bench_gemm_tc3(int M, int N, int K, int cl, int iters, float* ms_out, long long* ts_out, int* n_cta_out);
/* the same unit-test / timing entry points for the fp16x2 operand scheme of B2S_FP32 */
int b2s_test_gemm_h2(const float* A, const float* W, const float* bias, int M, int N, int K, float* C);
int b2s_test_attn_h2(const float* q, const float* k, const float* v, int nq, int nk, float* ctx);
int b2s_bench_attn_h2(int nq, int nk, int iters, float* ms_out);
int b2s_trace_attn_h2(int nq, int nk, int iters, float* ms_out, long long* trace_out);
int b2s_bench_gemm_h2(int M, int N, int K, int cl, int iters, float* ms_out, long long* ts_out, int* n_cta_out);

/* Number of CUDA kernels this library launched on behalf of the handle so far. */
long long b2s_aliked_launch_count(const b2s_aliked* h);
long long b2s_lg_launch_count(const b2s_lg* h);

#ifdef __cplusplus
}
#endif
#endif /* B200SLAM_H_ */
