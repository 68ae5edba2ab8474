/* svo_cuda.h — C ABI of the B200-native direct front-end hot path of SVO Pro.
 *
 * Drop-in boundary: the reference has no FFI; its front-end modules are C++ classes
 * new-ed inside FrameHandlerBase (src/svo/src/frame_handler_base.cpp:125,135,145). The C++ facades in
 * svo_pro_universal_b200/host/ keep those class names and signatures and call the entry points below.
 * Every entry point cites the reference interface it replaces.
 *
 * Conventions
 *   - every call returns an svo_status (0 = ok, negative = error; nothing throws across the ABI);
 *   - one svo_cuda_ctx is bound to one GPU and one CUDA stream; calls on a ctx are serialised on that
 *     stream, different contexts are independent (one ctx per GPU per host thread);
 *   - `mem` says where the I/O arrays of a batched call live: SVO_MEM_HOST pointers are staged through the
 *     context (async copies on its stream) and the call returns after the outputs have landed;
 *     SVO_MEM_DEVICE pointers are used in place and the call returns without synchronising;
 *   - transformations are 7 doubles (qw qx qy qz tx ty tz), T_a_b maps b-coordinates into a;
 *   - images are 8-bit, pyramids live in device memory inside an svo_cuda_pyr (a batch of frames).
 * There is no CPU fallback: without a CUDA device svo_cuda_ctx_create fails.
 */
#ifndef SVO_CUDA_H_
#define SVO_CUDA_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVO_MAX_LEVELS 8
#define SVO_MAX_CAMS 4

typedef enum {
  SVO_OK = 0,
  SVO_ERR_INVALID_ARG = -1,
  SVO_ERR_CUDA = -2,
  SVO_ERR_NO_DEVICE = -3,
  SVO_ERR_OUT_OF_MEMORY = -4,
  SVO_ERR_TOO_MANY_FEATURES = -5,
  SVO_ERR_UNSUPPORTED = -6
} svo_status;

typedef enum { SVO_MEM_HOST = 0, SVO_MEM_DEVICE = 1 } svo_mem;

typedef struct svo_cuda_ctx svo_cuda_ctx;
typedef struct svo_cuda_pyr svo_cuda_pyr;

/* ---- context ------------------------------------------------------------------------------- */
int svo_cuda_ctx_create(int device, svo_cuda_ctx** out);
int svo_cuda_ctx_destroy(svo_cuda_ctx* ctx);
/* Use an existing cudaStream_t (e.g. torch's current stream); NULL restores the context's own stream. */
int svo_cuda_ctx_set_stream(svo_cuda_ctx* ctx, void* cuda_stream);
int svo_cuda_ctx_synchronize(svo_cuda_ctx* ctx);
/* Human-readable description of the last error on this context (never NULL). */
const char* svo_cuda_last_error(const svo_cuda_ctx* ctx);
/* Number of kernels this library has launched on the context since creation (bench bookkeeping). */
long long svo_cuda_launch_count(const svo_cuda_ctx* ctx);
int svo_cuda_device_count(void);
/* sizeof() of a POD struct of this header by name (e.g. "svo_align_result"), -1 if unknown: lets bindings verify their layout. */
int svo_cuda_sizeof(const char* struct_name);

/* Page-locked host memory for the I/O arrays of SVO_MEM_HOST calls (asynchronous copies need it to overlap and to reach the link's
 * bandwidth). write_combined != 0 allocates write-combined pages: the CPU should only WRITE them (reads are uncached and slow), the
 * DMA engine reads them without snooping the CPU caches — meant for upload staging such as camera frames. */
int svo_cuda_host_alloc(svo_cuda_ctx* ctx, size_t bytes, int write_combined, void** out);
int svo_cuda_host_free(svo_cuda_ctx* ctx, void* ptr);

/* ---- camera model -------------------------------------------------------------------------- */
/* Pinhole with optional radial-tangential distortion:
 * vk::cameras::PinholeProjection<NoDistortion|RadialTangentialDistortion>
 * (src/vikit/vikit_cameras/include/vikit/cameras/implementation/pinhole_projection.hpp:30-76,
 *  radial_tangential_distortion.h:34-95). */
typedef struct {
  double fx, fy, cx, cy;
  double k1, k2, p1, p2;
  int width, height;
  int distortion; /* 0 = none, 1 = radial-tangential */
  int _pad;
} svo_camera;

/* ---- (a1) image pyramid: svo::Frame::img_pyr_ built by frame_utils::createImgPyramid
 *      (src/svo_common/src/frame.cpp:372-386) with vk::halfSample (src/vikit/vikit_common/src/vision.cpp:19-111).
 * halfsample_mode: -1 = the reference's x86 rule per level (SSE2 rounding formula when cols % 16 == 0, truncating
 * mean otherwise), 0 = always the truncating mean (non-SSE build). */
int svo_cuda_pyr_create(svo_cuda_ctx* ctx, int n_frames, int width, int height, int n_levels, int halfsample_mode,
                        svo_cuda_pyr** out);
int svo_cuda_pyr_destroy(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr);
/* Copy `count` level-0 images into frames [first, first+count). src rows are src_pitch bytes apart, frames
 * src_frame_stride bytes apart. `mem` tells whether src is host or device memory. Asynchronous on the ctx stream
 * (host memory should be pinned for the copy to overlap). */
int svo_cuda_pyr_upload(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count, const uint8_t* src, size_t src_pitch,
                        size_t src_frame_stride, svo_mem mem);
/* Build levels 1..n-1 of frames [first, first+count) from their level 0 (one fused launch for up to 5 levels). */
int svo_cuda_pyr_build(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count);
/* Copy one level of one frame out (tight dst_pitch >= cols). Synchronises when dst is host memory. */
int svo_cuda_pyr_download(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, uint8_t* dst, size_t dst_pitch,
                          svo_mem mem);
/* Geometry + device address of a level (frame 0); frames of one level are frame_stride bytes apart. */
int svo_cuda_pyr_level_info(const svo_cuda_pyr* pyr, int level, int* cols, int* rows, size_t* pitch, size_t* frame_stride,
                            void** device_ptr);

/* ---- (a2-a5) pyramidal FAST detector with grid-cell non-max suppression:
 *      svo::feature_detection_utils::fastDetector (src/svo_direct/src/feature_detection_utils.cpp:145-194),
 *      fast::fast_corner_detect_10[_sse2] / fast_corner_score_10 / fast_nonmax_3x3
 *      (src/fast_neon/include/fast/fast.h:20-41), OccupandyGrid2D::getCellIndex
 *      (src/svo_common/include/svo/common/occupancy_grid_2d.h:82-95). */
typedef struct { /* svo::Corner, src/svo_direct/include/svo/direct/feature_detection_types.h:17-29 */
  int x, y, level;
  float score, angle;
} svo_corner;

typedef struct { /* svo::DetectorOptions subset, feature_detection_types.h:49-84 */
  int threshold;   /* threshold_primary (FAST barrier), default 10 */
  int border;      /* default 8 */
  int min_level;   /* default 0 */
  int max_level;   /* default 2 */
  int cell_size;   /* default 30 */
  int arc_length;  /* 10 = what the reference front-end runs on x86; 9 = the ARM/NEON variant */
} svo_detector_options;

/* Number of grid cells per frame for a w x h image: ceil(w/cell) * ceil(h/cell). */
int svo_cuda_grid_cells(int width, int height, int cell_size, int* n_cols, int* n_rows);

/* For frames [first, first+count) (pyramid already built): per-cell best corner after per-level FAST detection,
 * scoring, 3x3 non-max, border test and occupancy skip. corners_out: [count][n_cells]; cells without a corner hold
 * (0,0,level 0,score=threshold,angle 0), exactly what fastDetector leaves in its pre-filled `corners`.
 * occupancy_in: [count][n_cells] bytes (non-zero = cell already holds a feature) or NULL. */
int svo_cuda_fast_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                         const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem);
/* Convenience variant (a1 + a2-a5): svo_cuda_pyr_build of the frames followed by svo_cuda_fast_detect on the same stream — two launches,
 * no host work in between; the detector re-reads levels 0-2 (it is bound by instruction issue, not by that traffic). */
int svo_cuda_pyramid_fast_detect(svo_cuda_ctx* ctx, svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                                 const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem);
/* Raw per-level stages, for parity tests against fast::* (a2, a3, a4): dense maps for one level of one frame.
 * score_map[rows][cols] int16: 0 where the pixel is not a corner at `threshold`, else fast_corner_score_10;
 * nonmax_map[rows][cols] uint8: 1 where the corner survives fast_nonmax_3x3. Either may be NULL. */
int svo_cuda_fast_level_maps(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int threshold, int arc_length,
                             int16_t* score_map, uint8_t* nonmax_map, svo_mem mem);

/* The fast:: leaf functions with their list-shaped results (src/fast_neon/include/fast/fast.h:11-41): corners of ONE level of ONE
 * frame in raster order, as fast_corner_detect_10[_sse2] (arc_length 10) / fast_corner_detect_9[_sse2] (9) push them. */
typedef struct { short x, y; } svo_fast_xy; /* fast::fast_xy, fast.h:11-15 */
/* fast_corner_detect_* + fast_corner_score_* + fast_nonmax_3x3 of a level in one call. n_out (HOST pointer in both memory modes) = number
 * of corners found; the first min(n_out, max_corners) are written: xy_out[i], scores_out[i] = fast_corner_score_10 at `threshold`,
 * nonmax_out[i] = 1 if fast_nonmax_3x3 keeps corner i (the reference returns the indices of these). Any output may be NULL. The call
 * synchronises the context's stream. */
int svo_cuda_fast_corner_list(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int threshold, int arc_length,
                              int max_corners, svo_fast_xy* xy_out, int* scores_out, uint8_t* nonmax_out, int* n_out, svo_mem mem);
/* fast_corner_score_10 (fast.h:38; src/fast_neon/src/fast_10_score.cpp:3150-3178) for a caller-supplied list: scores_out[i] = the largest
 * barrier at which xy[i] is still a corner, `threshold` if it is none above it (pixels closer than 3 to the border: `threshold`). */
int svo_cuda_fast_corner_score(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int frame, int level, int n, const svo_fast_xy* xy,
                               int threshold, int arc_length, int* scores_out, svo_mem mem);
/* fast_nonmax_3x3 (fast.h:41; src/fast_neon/src/nonmax_3x3.cpp:17-112) on a raster-ordered list: nonmax_idx_out (capacity n) receives the
 * indices, ascending, of the corners no 8-neighbour in the list matches or beats; n_out (HOST pointer) their number. Synchronises. */
int svo_cuda_fast_nonmax_3x3(svo_cuda_ctx* ctx, int n, const svo_fast_xy* xy, const int* scores, int* nonmax_idx_out, int* n_out,
                             svo_mem mem);

/* ---- (f2) edgelet detector and the FastGrad combination (the reference's default detector, svo_factory.cpp:292-295) ------ */
/* feature_detection_utils::edgeletDetector_V2 (src/svo_direct/include/svo/direct/feature_detection_utils.h:75-82;
 * src/svo_direct/src/feature_detection_utils.cpp:313-385) with getAngleAtPixelUsingHistogram (:831-839, 945-1009) for the
 * winners, i.e. the device part of GradientDetectorGrid::detect (src/svo_direct/src/feature_detection.cpp:130-151).
 * Works on pyramid level 1 (the pyramid needs >= 2 levels, already built); per-cell result in corners_out [count][n_cells]:
 * level 0, px = 2 * the level-1 pixel, score = gradient magnitude (float), angle = dominant histogram angle; cells without an
 * edgelet hold (0,0,0,score=threshold,0). threshold = DetectorOptions::threshold_secondary as the reference's int.
 * border >= 4 is required: the reference's neighbour test reads 4 rows up and down (it offsets a float pointer by a byte step). */
int svo_cuda_edgelet_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, int threshold, int border,
                            int cell_size, const uint8_t* occupancy_in, svo_corner* corners_out, svo_mem mem);
/* FastGradDetector::detect (feature_detection.h:133-150; feature_detection.cpp:154-194) up to fillFeatures' sort: FAST corners
 * per cell (opt, as svo_cuda_fast_detect) into corners_out, then edgelets (threshold_secondary) into edgelets_out for the cells
 * that neither occupancy_in nor a FAST corner occupies; the edgelet stage is skipped for frames whose corners already reach
 * max_n_features (:176-177). Both outputs are [count][n_cells]. */
int svo_cuda_fastgrad_detect(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, int first, int count, const svo_detector_options* opt,
                             int threshold_secondary, int max_n_features, const uint8_t* occupancy_in, svo_corner* corners_out,
                             svo_corner* edgelets_out, svo_mem mem);
/* Raw stage for parity tests: angle_hist::angleHistogram's bin (feature_detection_utils.cpp:945-962) of every central-difference
 * gradient (gx, gy) in [-255, 255]^2; bins_out [511][511] int8, row gy + 255, column gx + 255. */
int svo_cuda_angle_histogram_bins(svo_cuda_ctx* ctx, int8_t* bins_out, svo_mem mem);

/* ---- (b) svo::SparseImgAlign (src/svo_img_align/include/svo/img_align/sparse_img_align.h:30-77,
 *      sparse_img_align_base.h:37-163; run(): src/svo_img_align/src/sparse_img_align.cpp:34-113) ---------- */
typedef struct {
  /* SparseImgAlignOptions (sparse_img_align_base.h:37-46) */
  int max_level, min_level;
  int estimate_illumination_gain, estimate_illumination_offset;
  int use_distortion_jacobian, robustification;
  double weight_scale;
  /* vk::solver::MiniLeastSquaresSolverOptions (GaussNewton; sparse_img_align_base.cpp:35-42) */
  int max_iter;
  int _pad;
  double eps;
  /* setAlphaInitialValue / setBetaInitialValue */
  double alpha_init, beta_init;
  /* setWeightedPrior lambdas (used when a prior array is passed) */
  double lambda_rot, lambda_trans, lambda_alpha, lambda_beta;
} svo_sparse_align_options;

typedef struct { /* setWeightedPrior(T_cur_ref_prior, alpha_prior, beta_prior, ...) per pair */
  double T[7];
  double alpha, beta;
} svo_align_prior;

typedef struct {
  double T_icur_iref[7];           /* optimised state */
  double T_f_w[SVO_MAX_CAMS][7];   /* cur frames' T_f_w_ = T_cam_imu * T_icur_iref * T_iref_world */
  double alpha, beta;
  double chi2;                     /* getError() */
  double H[64];                    /* getHessian(), row-major 8x8, last evaluated */
  int n_tracked;                   /* run() return value */
  int iters[SVO_MAX_LEVELS];       /* evaluateError calls per level, from max_level down */
  int stop;                        /* solver stop_ flag (NaN in dx) */
} svo_align_result;

/* B independent frame-bundle pairs, each with n_cams cameras (1 = mono, 2 = stereo).
 *   ref_pyr/cur_pyr[c]      : pyramid batches of camera c; pair i uses frame ref_frame_idx[i*n_cams+c]
 *                             (NULL index array = frame i).
 *   cams[c], T_cam_imu[c*7] : camera model and extrinsics of camera c (shared by all pairs).
 *   T_imu_world_ref/cur     : [B][7]; cur is the initial guess (frame_handler_base.cpp:346-358).
 *   n_features[i*n_cams+c]  : features of ref frame c of pair i, stored at [(i*n_cams+c)*max_features + k] in
 *   px [..][2], f [..][3] (unit bearing, Frame::f_vec_), depth [..] (distance landmark/seed -> ref camera centre),
 *   eligible [..] (1 = has landmark or seed reference and is not a MapPoint type, sparse_img_align.cpp:242-248).
 *   priors                  : NULL or [B].
 * All arrays follow `mem`. */
int svo_cuda_sparse_align(svo_cuda_ctx* ctx, int n_cams, const svo_cuda_pyr* const* ref_pyr, const svo_cuda_pyr* const* cur_pyr,
                          const int* ref_frame_idx, const int* cur_frame_idx, const svo_camera* cams, const double* T_cam_imu,
                          int B, const double* T_imu_world_ref, const double* T_imu_world_cur, const int* n_features,
                          int max_features, const double* px, const double* f, const double* depth, const uint8_t* eligible,
                          const svo_sparse_align_options* opt, const svo_align_prior* priors, svo_align_result* results,
                          svo_mem mem);

/* ---- (c) svo::feature_alignment + svo::Matcher ------------------------------------------------------------ */
typedef struct { /* svo::Matcher::Options, src/svo_direct/include/svo/direct/matcher.h:39-54 */
  int align_1d, align_max_iter;
  int max_epi_search_steps;
  int subpix_refinement, epi_search_edgelet_filtering, scan_on_unit_sphere;
  double epi_search_edgelet_max_angle;
  int affine_est_offset, affine_est_gain;
  double max_patch_diff_ratio;
} svo_matcher_options;

typedef struct { /* the parts of svo::FeatureWrapper the matcher reads (feature_wrapper.h:34-45) */
  double px[2];
  double f[3];
  double grad[2];
  int type;  /* svo::FeatureType */
  int level;
} svo_feature;

typedef struct { /* Matcher public members read back by callers (matcher.h:70-79) + MatchResult */
  double px_cur[2];
  double f_cur[3];
  double A_cur_ref[4];
  double h_inv;
  double epi_length_pyramid;
  double depth;
  int result; /* svo::Matcher::MatchResult */
  int search_level;
  int reject;
  int _pad;
  double epi_image[2]; /* Matcher::epi_image_ (matcher.h:73): px_A - px_B, set by findEpipolarMatchDirect */
} svo_match_out;

/* feature_alignment::align2D (src/svo_direct/include/svo/direct/feature_alignment.h:34-43; .cpp:212-391).
 * M patches against level `level` of frame frame_idx[i] of `pyr`: patch_with_border [M][100] u8, px [M][2] in/out
 * (level coordinates), converged [M] u8. */
int svo_cuda_align2d(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const int* frame_idx, const int* level, int M,
                     const uint8_t* patch_with_border, int n_iter, int affine_est_offset, int affine_est_gain, double* px,
                     uint8_t* converged, svo_mem mem);
/* feature_alignment::align1D (feature_alignment.h:23-32; .cpp:31-209): dir [M][2]; h_inv [M] out (may be NULL). */
int svo_cuda_align1d(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr, const int* frame_idx, const int* level, int M,
                     const double* dir, const uint8_t* patch_with_border, int n_iter, int affine_est_offset,
                     int affine_est_gain, double* px, double* h_inv, uint8_t* converged, svo_mem mem);
/* feature_alignment::alignPyr2D / alignPyr2DVec (feature_alignment.h:57-80; .cpp:731-973): pyramidal KLT of M features from
 * (ref frame ref_frame_idx[i]) into (cur frame cur_frame_idx[i]); px_ref_level_0 int [M][2]; px_cur double [M][2] in/out
 * (level-0 px); patch_sizes [n_levels] (8 or 16 per level, indexed by level); status [M] = converged. */
int svo_cuda_align_pyr2d(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const int* ref_frame_idx,
                         const int* cur_frame_idx, int M, const int* px_ref_level_0, double* px_cur, int max_level,
                         int min_level, const int* patch_sizes, int n_iter, float min_update_squared, uint8_t* status,
                         svo_mem mem);
/* warp::getWarpMatrixAffine + getBestSearchLevel + warpAffine (src/svo_direct/src/patch_warp.cpp:20-60,97-156) followed by
 * the 10x10 -> 8x8 crop; outputs A [M][4], search_level [M], patch_with_border [M][100], ok [M]. Mostly for parity tests. */
int svo_cuda_warp_affine(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const int* ref_frame_idx, const svo_camera* cam_ref,
                         const svo_camera* cam_cur, const double* T_cur_ref, const int* T_idx, int M, const svo_feature* ftrs,
                         const double* depth, double* A_out, int* search_level_out, uint8_t* patch_with_border_out,
                         uint8_t* ok_out, svo_mem mem);
/* Matcher::findMatchDirect (matcher.h:84-89; matcher.cpp:31-141) for M features. Feature i lives in ref frame
 * ref_frame_idx[i], is searched in cur frame cur_frame_idx[i] with T_cur_ref[T_idx[i]] ([..][7]);
 * px_cur_guess [M][2] is the caller's estimate (level-0 px). */
int svo_cuda_find_match_direct(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr,
                               const int* ref_frame_idx, const int* cur_frame_idx, const svo_camera* cam_ref,
                               const svo_camera* cam_cur, const double* T_cur_ref, const int* T_idx, int M,
                               const svo_feature* ftrs, const double* ref_depth, const double* px_cur_guess,
                               const svo_matcher_options* opt, svo_match_out* out, svo_mem mem);
/* Matcher::findEpipolarMatchDirect (matcher.h:100-108; matcher.cpp:157-241): d_inv [M][3] = (estimate, min, max)
 * inverse depths in the order the reference passes them (d_estimate_inv, d_min_inv, d_max_inv). */
int svo_cuda_find_epipolar_match_direct(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr,
                                        const int* ref_frame_idx, const int* cur_frame_idx, const svo_camera* cam_ref,
                                        const svo_camera* cam_cur, const double* T_cur_ref, const int* T_idx, int M,
                                        const svo_feature* ftrs, const double* d_inv, const svo_matcher_options* opt,
                                        svo_match_out* out, svo_mem mem);

/* Matcher::scanEpipolarLine (matcher.h:111-122; matcher.cpp:324-488) on its own, for M independent scans: segment A~C~B [M][3] each
 * (points in the cur camera frame), the 8x8 reference patch [M][64] the PatchScore is built from, patch_level [M], the Matcher member
 * epi_length_pyramid_ [M] the scan length derives from, zmssd_best [M] in/out (the caller's starting value, PatchScore::threshold() in
 * the reference's own call), image_best [M][2] out. opt: scan_on_unit_sphere and max_epi_search_steps are read. */
int svo_cuda_scan_epipolar_line(svo_cuda_ctx* ctx, const svo_cuda_pyr* cur_pyr, const int* cur_frame_idx, const svo_camera* cam_cur, int M,
                                const double* A, const double* B, const double* C, const uint8_t* patch, const int* patch_level,
                                const double* epi_length_pyramid, const svo_matcher_options* opt, double* image_best, int* zmssd_best,
                                svo_mem mem);

/* ---- (f3) svo::StereoTriangulation ------------------------------------------------------------------------------- */
typedef enum { SVO_STEREO_NOT_REACHED = 0, SVO_STEREO_FAILED = 1, SVO_STEREO_SUCCESS = 2 } svo_stereo_status;
typedef struct { /* what StereoTriangulation::compute writes per matched feature (stereo_triangulation.cpp:102-129) */
  double px_cur[2], f_cur[3];   /* matcher.px_cur_, matcher.f_cur_ -> frame1's px_vec_ / f_vec_ column */
  double grad_cur[2];           /* (A_cur_ref * ref grad).normalized() -> frame1's grad_vec_ column */
  double xyz_world[3];          /* the new Point: frame0->T_world_cam() * (f * depth) */
  double depth;
  int status;                   /* svo_stereo_status */
  int slot;                     /* success: feature slot in frame1 (num_features_ at that moment) */
  int match_result;             /* Matcher::MatchResult, -1 = not reached */
  int level, type;              /* success: copied from the frame0 feature */
  int _pad;
} svo_stereo_result;
typedef struct { int n_succeeded, n_failed; } svo_stereo_stats;

/* StereoTriangulation::compute (src/svo/include/svo/stereo_triangulation.h:20-37; src/svo/src/stereo_triangulation.cpp:87-137: the
 * matching loop) for B independent stereo pairs: frame0 of pair b = frame frame0_idx[b] of pyr0 (NULL = b), frame1 likewise in pyr1;
 * T_f1f0 [7] = frame1->T_cam_body_ * frame0->T_body_cam_ (the rig's extrinsics, shared); T_world_cam0 [B][7]. Pair b tries the
 * features ftrs[feat_begin[b] .. feat_begin[b+1]) of frame0 IN THIS ORDER (the caller has detected them and applied the reference's
 * two std::random_shuffle calls, :34-79) with Matcher::findEpipolarMatchDirect (align_1d = isEdgelet(type), the inverse-depth
 * range given) until n_desired[b] = triangulate_n_features - frame0->numLandmarks() of them succeeded; n_features_in_frame1[b] =
 * frame1->num_features_ before the call. mopt: the Matcher options (the reference sets max_epi_search_steps = 500 and
 * subpix_refinement = true, :90-91; its align_1d field is ignored). results [n_features], stats [B]. n_features = feat_begin[B].
 * Entries whose `type` is negative are holes: they are not matched, not counted and come back SVO_STEREO_NOT_REACHED, so that a
 * device-resident caller can pass fixed-shape lists (e.g. one slot per grid cell) without compacting them on the host. */
int svo_cuda_stereo_triangulate(svo_cuda_ctx* ctx, const svo_cuda_pyr* pyr0, const svo_cuda_pyr* pyr1, const int* frame0_idx,
                                const int* frame1_idx, const svo_camera* cam0, const svo_camera* cam1, const double* T_f1f0,
                                const double* T_world_cam0, int B, const int* feat_begin, int n_features, const svo_feature* ftrs,
                                const int* n_desired, const int* n_features_in_frame1, double mean_depth_inv, double min_depth_inv,
                                double max_depth_inv, const svo_matcher_options* mopt, svo_stereo_result* results,
                                svo_stereo_stats* stats, svo_mem mem);

/* ---- (d) svo::DepthFilter seed update ---------------------------------------------------------------------- */
/* depth_filter_utils::updateFilterVogiatzis (src/svo_direct/include/svo/direct/depth_filter.h:201-205;
 * src/svo_direct/src/depth_filter.cpp:501-552): n independent updates; state [n][4] = (mu, sigma2, a, b) in/out;
 * ok [n] (0 = the reference returns false). */
int svo_cuda_update_filter_vogiatzis(svo_cuda_ctx* ctx, int n, const double* z, const double* tau2, const double* mu_range,
                                     double* state, uint8_t* ok, svo_mem mem);
/* n_obs ORDERED updates per seed in ONE launch (the filter state stays in registers): z, tau2 [n_obs][n] (observation-major),
 * state [n][4] in/out, ok [n_obs][n] (may be NULL). Equivalent to n_obs successive calls of updateFilterVogiatzis (gaussian == 0;
 * depth_filter.h:201-205, depth_filter.cpp:501-552) or updateFilterGaussian (gaussian != 0; depth_filter.h:207-211,
 * depth_filter.cpp:554-578; mu_range is not read and may be NULL) on the same seed. The Vogiatzis quotients are formed with shared
 * reciprocals: results agree with the reference's to a few ulp per update, not bit for bit. */
int svo_cuda_update_filter_seq(svo_cuda_ctx* ctx, int n, int n_obs, const double* z, const double* tau2, const double* mu_range,
                               double* state, uint8_t* ok, int gaussian, svo_mem mem);
/* depth_filter_utils::computeTau (depth_filter.cpp:580-596): T_ref_cur [n][7], f [n][3], z [n] -> tau [n]. */
int svo_cuda_compute_tau(svo_cuda_ctx* ctx, int n, const double* T_ref_cur, const double* f, const double* z,
                         double px_error_angle, double* tau, svo_mem mem);

typedef struct {
  double seed_convergence_sigma2_thresh;      /* DepthFilterOptions, depth_filter.h:33 (200) */
  double mappoint_convergence_sigma2_thresh;  /* depth_filter.h:37 (500) */
  double px_error_angle;  /* cam.getAngleError(1.0): the function-static of depth_filter.cpp:383-384; <= 0 = derive from cam_cur */
  int check_visibility, check_convergence, use_vogiatzis_update;
  int _pad;
} svo_depth_filter_options;

/* DepthFilter::updateSeeds / depth_filter_utils::updateSeed (depth_filter.h:118-120,190-199; depth_filter.cpp:200-249,
 * 367-499): S seeds, seed s lives in ref frame ref_frame_idx[s] of ref_pyr with mu range seed_mu_range[s];
 * it is observed, in order, by n_obs frames: observation o of seed s uses cur frame obs_frame_idx[o*S+s] and
 * T_cur_ref[obs_T_idx[o*S+s]]; a negative obs_frame_idx skips that observation (e.g. cur == ref frame).
 * types [S] (uint8 FeatureType) and state [S][4] are updated in place; n_success[0] counts successful updates;
 * match_results (optional) [n_obs][S] receives each Matcher::MatchResult (-1 = no match attempted). */
int svo_cuda_update_seeds(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const svo_camera* cam_ref,
                          const svo_camera* cam_cur, int S, const int* ref_frame_idx, const svo_feature* ftrs, uint8_t* types,
                          double* state, const double* seed_mu_range, int n_obs, const int* obs_frame_idx,
                          const int* obs_T_idx, const double* T_cur_ref, const svo_matcher_options* mopt,
                          const svo_depth_filter_options* dopt, int* n_success, int* match_results, svo_mem mem);

/* ---- (f1) svo::Reprojector candidate matching ----------------------------------------------------------------- */
/* reprojector_utils::getCandidate / projectPointAndCheckVisibility / sortCandidatesByReprojStats / sortCandidatesByNumObs /
 * matchCandidates / matchCandidate (src/svo/include/svo/reprojector.h:168-200; src/svo/src/reprojector.cpp:310-543) with
 * Frame::isVisible (src/svo_common/src/frame.cpp:229-257) and Point::getCloseViewObs (src/svo_common/src/point.cpp:83-129).
 * The map is handed over as flat tables: the keyframes' feature columns (Frame SoA, frame.h:62-73) and the landmarks'
 * bookkeeping (point.h:82-91). The tables are read only; everything the reference mutates comes back per entry. */
typedef struct {
  int n_kfs, n_feat, n_points, n_obs;
  const double* kf_T_f_w;          /* [n_kfs][7] Frame::T_f_w_ */
  const double* kf_seed_mu_range;  /* [n_kfs] Frame::seed_mu_range_ */
  const int* kf_frame_idx;         /* [n_kfs] frame of keyframe k inside ref_pyr (NULL = k) */
  const svo_feature* feat;         /* [n_feat] px_vec_, f_vec_, grad_vec_, type_vec_, level_vec_ of all keyframes, back to back */
  const double* feat_score;        /* [n_feat] score_vec_ */
  const double* feat_seed_state;   /* [n_feat][4] invmu_sigma2_a_b_vec_ */
  const int* feat_point;           /* [n_feat] landmark id, -1 = landmark_vec_[i] == nullptr (a seed) */
  const int* feat_kf;              /* [n_feat] keyframe that holds the feature */
  const double* pt_pos;            /* [n_points][3] Point::pos_ */
  const int* pt_n_failed;          /* [n_points] Point::n_failed_reproj_ */
  const int* pt_n_succeeded;       /* [n_points] Point::n_succeeded_reproj_ */
  const int* pt_obs_begin;         /* [n_points+1] Point::obs_ of point p = obs_feat[begin[p] .. begin[p+1]) */
  const int* obs_feat;             /* [n_obs] feature index (into feat) of each observation */
} svo_reproj_map;

typedef struct { /* ReprojectorOptions subset (reprojector.h:27-70) + the arguments of matchCandidates */
  int cell_size;                   /* default 30 */
  int max_n_features;              /* matchCandidates' max_n_features_per_frame; 0 = unlimited and occupancy ignored */
  int affine_est_offset, affine_est_gain;
  int sort_by_num_obs;             /* 0 = sortCandidatesByReprojStats, 1 = sortCandidatesByNumObs */
  int _pad;
  double seed_sigma2_thresh;       /* default 200 */
  double px_error_angle;           /* updateSeed's function-static (depth_filter.cpp:383-384); <= 0 = derive from cam_cur */
} svo_reprojector_options;

typedef enum {
  SVO_REPROJ_NOT_CANDIDATE = 0,    /* getCandidate returned false (not visible in the current frame) */
  SVO_REPROJ_NOT_REACHED = 1,      /* behind the point where matchCandidates stopped: stays in candidates_ */
  SVO_REPROJ_SKIPPED = 2,          /* its grid cell was occupied at its turn */
  SVO_REPROJ_FAILED = 3,           /* tried, no match */
  SVO_REPROJ_MATCHED = 4
} svo_reproj_status;

typedef struct {
  double cur_px[2];                /* Candidate::cur_px */
  double px[2], f[3], grad[2];     /* matched: what matchCandidate writes into the current frame's feature slot */
  double seed_state[4];            /* the ref feature's seed state after the call (changed when an unconverged seed was tried) */
  int status;                      /* svo_reproj_status */
  int order;                       /* position in the sorted candidate list, -1 = not a candidate */
  int slot;                        /* matched: feature slot in the current frame (num_features_ at that moment) */
  int level;                       /* matched: matcher.search_level_ */
  int type_out;                    /* the ref feature's type after the call (updateSeed may converge it / mark it an outlier) */
  int match_result;                /* Matcher::MatchResult of the attempt, -1 = none */
  int d_failed, d_succeeded;       /* increments of the landmark's n_failed_reproj_ / n_succeeded_reproj_ */
} svo_reproj_result;

typedef struct { int n_candidates, n_trials, n_matches, n_consumed; } svo_reproj_stats;

/* F current frames at once (independent reprojections, e.g. the frames of a batch of sequences): frame j lives in
 * cur_pyr frame cur_frame_idx[j] (NULL = j) with pose cur_T_f_w[j], holds n_features_in[j] features already, and reprojects
 * the entries entry_feat[entry_begin[j] .. entry_begin[j+1]) (feature indices into map->feat, in the reference's visiting
 * order: the features of the visible keyframes that pass the map bookkeeping of Reprojector::reprojectFrames).
 * n_entries = entry_begin[F]. occupancy [F][n_cells] is the grid, in/out (n_cells from svo_cuda_grid_cells with cam_cur's
 * size); results [n_entries]; stats [F]. Candidates that compare equal under the reference's sort keep their visiting order (the reference's
 * std::sort leaves their order unspecified). At most 4096 entries per frame and 4096 grid cells. All arrays, including those inside `map`,
 * follow `mem`. */
int svo_cuda_reproject_match(svo_cuda_ctx* ctx, const svo_cuda_pyr* ref_pyr, const svo_cuda_pyr* cur_pyr, const svo_camera* cam_ref,
                             const svo_camera* cam_cur, const svo_reproj_map* map, int F, const int* cur_frame_idx,
                             const double* cur_T_f_w, const int* n_features_in, const int* entry_begin, int n_entries,
                             const int* entry_feat, uint8_t* occupancy, const svo_reprojector_options* opt, svo_reproj_result* results,
                             svo_reproj_stats* stats, svo_mem mem);

/* ---- (f4) svo::PoseOptimizer ------------------------------------------------------------------------------------ */
typedef struct { /* PoseOptimizer::getDefaultSolverOptions (src/svo/src/pose_optimizer.cpp:22-29) + run()'s arguments */
  int err_type;             /* PoseOptimizer::ErrorType: 0 kUnitPlane (default), 1 kBearingVectorDiff, 2 kImagePlane */
  int max_iter;             /* 10 */
  double eps;               /* 0.000001 */
  double reproj_thresh_px;  /* run(frame_bundle, reproj_thresh_px): outlier threshold, default poseoptim_thresh = 2.0 */
  double prior_lambda;      /* setRotationPrior(R_frame_world, lambda); used when prior_q is passed */
} svo_pose_optimizer_options;

typedef struct {
  double T_imu_world[7];           /* optimised state */
  double T_f_w[SVO_MAX_CAMS][7];   /* frame->T_f_w_ = T_cam_imu * T_imu_world of every camera (pose_optimizer.cpp:58-61) */
  double measurement_sigma;        /* MAD scale of the start errors */
  double reproj_error_before, reproj_error_after;  /* stats_ (medians, scaled by the focal length for kUnitPlane) */
  double chi2;
  int n_meas_final;                /* run()'s return value: measurements minus removed outliers */
  int n_meas;
  int iters;                       /* iterCount() */
  int stop;
} svo_pose_opt_result;

/* PoseOptimizer::run (src/svo/include/svo/pose_optimizer.h:20-103; src/svo/src/pose_optimizer.cpp:39-94) for B independent
 * frame bundles of n_cams cameras (cams[c], T_cam_imu [n_cams][7] shared by all bundles). Bundle b starts at
 * T_imu_world[b] and owns the features [feat_begin[b], feat_begin[b+1]) of ftrs (px, f, grad, level, type of the current
 * frames' columns), observed by camera feat_cam[i] (NULL = camera 0), with the 3-D point xyz_world[i] where has_xyz[i] (the
 * landmark's position, or the seed's position from its reference keyframe; features with neither are skipped as in
 * evaluateErrorImpl, :123-137). prior_q: NULL or [B][4], the rotation prior R_frame_world of setRotationPrior.
 * outlier [n_features] receives 1 where removeOutliers marks the feature kOutlier (:283-290; the caller resets the landmark /
 * seed reference). At most 2048 features per bundle. n_features = feat_begin[B]. */
int svo_cuda_pose_optimize(svo_cuda_ctx* ctx, int n_cams, const svo_camera* cams, const double* T_cam_imu, int B,
                           const double* T_imu_world, const int* feat_begin, int n_features, const svo_feature* ftrs,
                           const int* feat_cam, const double* xyz_world, const uint8_t* has_xyz, const double* prior_q,
                           const svo_pose_optimizer_options* opt, svo_pose_opt_result* results, uint8_t* outlier, svo_mem mem);

/* Point::optimize (src/svo_common/include/svo/common/point.h:155, 170-204; src/svo_common/src/point.cpp:216-325) for P independent
 * points, as FrameHandlerBase::optimizeStructure runs it over the landmarks of a frame (src/svo/src/frame_handler_base.cpp:785-825):
 * point i (pos [P][3], in/out) is observed by obs_frame[o] with unit bearing obs_f[o] for o in [obs_begin[i], obs_begin[i+1]);
 * T_f_w [n_frames][7] are the observing frames' poses. Points with fewer than two observations are left alone (:255-259).
 * using_bearing_vector selects the unit-sphere residual (omni cameras), else the unit plane. iters_out [P] (may be NULL) receives
 * the number of Gauss-Newton iterations started. n_obs = obs_begin[P]. */
int svo_cuda_optimize_points(svo_cuda_ctx* ctx, int P, double* pos, const int* obs_begin, int n_obs, const int* obs_frame,
                             const double* obs_f, int n_frames, const double* T_f_w, int n_iter, int using_bearing_vector,
                             int* iters_out, svo_mem mem);

#ifdef __cplusplus
}
#endif
#endif /* SVO_CUDA_H_ */
