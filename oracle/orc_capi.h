/* ORACLE — TEST INFRASTRUCTURE ONLY.
 * C interface of the CPU oracle (liborc.so), loaded by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs through ctypes. The product never links it.
 * All pointers are host pointers. Transformations are 7 doubles (qw qx qy qz tx ty tz).
 */
#ifndef ORC_CAPI_H_
#define ORC_CAPI_H_
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 8
#define ORC_MAX_CAMS 4

typedef struct {
  int x, y, level;
  float score, angle;
} orc_corner;

/* One camera frame: image pyramid (tight or strided), camera model, poses and (for ref frames of
 * sparse alignment) the feature SoA run() reads. */
typedef struct {
  const uint8_t* level_data[ORC_MAX_LEVELS];
  int level_cols[ORC_MAX_LEVELS];
  int level_rows[ORC_MAX_LEVELS];
  int level_step[ORC_MAX_LEVELS];
  int n_levels;
  double cam[8]; /* fx fy cx cy k1 k2 p1 p2 */
  int width, height, distortion; /* distortion: 0 none, 1 radtan */
  double T_cam_imu[7];
  double T_imu_world[7];
  int n_features;
  const double* px;        /* [n][2] */
  const double* f;         /* [n][3] */
  const double* depth;     /* [n]    */
  const uint8_t* eligible; /* [n]    */
} orc_frame;

typedef struct {
  int max_level, min_level;
  int estimate_illumination_gain, estimate_illumination_offset;
  int use_distortion_jacobian, robustification;
  double weight_scale;
  int max_iter;
  double eps;
  double alpha_init, beta_init;
  int have_prior;
  double prior_T[7];
  double prior_alpha, prior_beta;
  double lambda_rot, lambda_trans, lambda_alpha, lambda_beta;
} orc_align_options;

typedef struct {
  int n_tracked;
  double T_icur_iref[7];
  double alpha, beta, chi2;
  double H[64];
  int iters[ORC_MAX_LEVELS]; /* evaluateError calls per level, from max_level down */
  double T_f_w[ORC_MAX_CAMS][7];
  int stop;
} orc_align_result;

typedef struct {
  int type; /* svo::FeatureType value */
  double px[2];
  double f[3];
  double grad[2];
  int level;
} orc_feature;

typedef struct {
  int align_1d, align_max_iter;
  int max_epi_search_steps;
  int subpix_refinement, epi_search_edgelet_filtering, scan_on_unit_sphere;
  double epi_search_edgelet_max_angle;
  int affine_est_offset, affine_est_gain;
  double max_patch_diff_ratio;
} orc_matcher_options;

typedef struct {
  int result; /* Matcher::MatchResult */
  double px_cur[2];
  double f_cur[3];
  int search_level;
  double A_cur_ref[4]; /* row-major 2x2 */
  double h_inv;
  double epi_length_pyramid;
  int reject;
  double depth;
  uint8_t patch_with_border[100];
  double epi_image[2]; /* Matcher::epi_image_ (matcher.h:73) */
} orc_match_out;

/* a1 */
void orc_half_sample(const uint8_t* in, int cols, int rows, int stride, uint8_t* out, int out_stride, int mode);
size_t orc_create_img_pyramid(const uint8_t* img0, int cols, int rows, int n_levels, uint8_t* out_levels_1_up, int mode);
/* a2-a4 (restated) */
int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int barrier, int arc, short* xy, int cap);
void orc_fast_score10(const uint8_t* img, int stride, const short* xy, int n, int threshold, int* scores);
int orc_fast_nonmax3x3(const short* xy, const int* scores, int n, int* idx_out);
/* a5: corners_out has n_cols*n_rows entries, initialised inside to (0,0,score=threshold,0,0). */
void orc_fast_detector(const uint8_t* img0, int cols, int rows, int n_levels, int pyr_mode, int threshold, int border,
                       int min_level, int max_level, int cell_size, const uint8_t* occupancy, orc_corner* corners_out);
/* a5+a6: FastDetector::detect; returns the number of features written (<= max_n). occupancy may be NULL. */
int orc_fast_detect_features(const uint8_t* img0, int cols, int rows, int n_levels, int pyr_mode, double threshold, int border,
                             int min_level, int max_level, int cell_size, const uint8_t* occupancy, int max_n,
                             double* px_out, double* score_out, int* level_out);
/* f2: edgelet detector + the detector classes. Pyramid levels are passed in (n_levels pointers / cols / rows / step). */
void orc_gaussian_blur3x3(const uint8_t* img, int cols, int rows, int step, uint8_t* out);
void orc_scharr3x3(const uint8_t* img, int cols, int rows, int step, int16_t* dx, int16_t* dy);
void orc_edgelet_detector_v2(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step,
                             int threshold, int border, int cell_size, const uint8_t* occupancy, orc_corner* corners_out);
double orc_angle_at_pixel_histogram(const uint8_t* img, int cols, int rows, int step, int x, int y, int halfpatch_size);
int orc_angle_histogram_bin(int gx, int gy);
void orc_fast_detector_pyr(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step, int threshold,
                           int border, int min_level, int max_level, int cell_size, const uint8_t* occupancy, orc_corner* corners_out);
int orc_detect_features(int detector_type, int n_levels, const uint8_t* const* data, const int* cols, const int* rows,
                        const int* step, double threshold_primary, double threshold_secondary, int border, int min_level,
                        int max_level, int cell_size, const uint8_t* occupancy, int max_n, double* px_out, double* score_out,
                        int* level_out, double* grad_out, int* type_out);
/* f3 (StereoTriangulation::compute after detection and shuffling): entry i = feature ftrs[i] of frame0 in visiting order.
 * status: 0 = not reached (n_desired successes came first), 1 = tried and failed, 2 = success. */
typedef struct {
  double px_cur[2], f_cur[3], grad_cur[2], xyz_world[3], depth;
  int status, slot, match_result, level, type, _pad;
} orc_stereo_result;
int orc_stereo_triangulate(const orc_frame* frame0, const orc_frame* frame1, int n, const orc_feature* ftrs, int n_desired,
                           int n_features_in_frame1, double mean_depth_inv, double min_depth_inv, double max_depth_inv,
                           orc_stereo_result* results, int* n_failed);
/* f4 (second half): Point::optimize on one point with n_obs observations (T_f_w [n_obs][7], f [n_obs][3]); pos in/out.
 * Returns the number of iterations started. */
int orc_point_optimize(int n_obs, const double* T_f_w, const double* f, double pos[3], int n_iter, int using_bearing_vector);
/* b */
int orc_sparse_align(int n_cams, const orc_frame* ref, const orc_frame* cur, const orc_align_options* opt, orc_align_result* res);
/* B independent problems: ref/cur hold B*n_cams frames; n_threads worker threads (one problem per task). */
int orc_sparse_align_batch(int B, int n_cams, const orc_frame* ref, const orc_frame* cur, const orc_align_options* opt,
                           orc_align_result* res, int n_threads);
/* c */
int orc_warp_affine(const double A_cur_ref[4], const uint8_t* img, int cols, int rows, int step, const double px_ref[2],
                    int level_ref, int search_level, int halfpatch_size, uint8_t* patch);
void orc_get_warp_matrix_affine(const orc_frame* ref, const orc_frame* cur, const double px_ref[2], const double f_ref[3],
                                double depth_ref, const double T_cur_ref[7], int level_ref, double A_out[4]);
int orc_get_best_search_level(const double A[4], int max_level);
int orc_zmssd(const uint8_t* ref_patch64, const uint8_t* cur, int stride);
int orc_align2d(const uint8_t* img, int cols, int rows, int step, const uint8_t* patch_with_border, int n_iter,
                int est_offset, int est_gain, double px[2]);
int orc_align1d(const uint8_t* img, int cols, int rows, int step, const double dir[2], const uint8_t* patch_with_border,
                int n_iter, int est_offset, int est_gain, double px[2], double* h_inv);
int orc_find_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                          double ref_depth, const double px_cur_in[2], const orc_matcher_options* opt, orc_match_out* out);
int orc_find_epipolar_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                                   double d_estimate_inv, double d_min_inv, double d_max_inv, const orc_matcher_options* opt,
                                   orc_match_out* out);
/* Matcher::scanEpipolarLine on its own (matcher.h:111-122; matcher.cpp:324-488): segment A~C~B in the cur camera frame, the 8x8
 * reference patch, the member epi_length_pyramid_ the scan length derives from; zmssd_best in/out, image_best out. */
void orc_scan_epipolar_line(const orc_frame* cur, const double A[3], const double B[3], const double C[3], const uint8_t* patch64,
                            int patch_level, double epi_length_pyramid, const orc_matcher_options* opt, double image_best[2],
                            int* zmssd_best);
/* M features sharing (ref, cur, T_cur_ref); threaded. */
int orc_find_match_direct_batch(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], int M,
                                const orc_feature* ftrs, const double* ref_depth, const double* px_cur_in,
                                const orc_matcher_options* opt, orc_match_out* out, int n_threads);
int orc_find_epipolar_match_direct_batch(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], int M,
                                         const orc_feature* ftrs, const double* d_inv3 /* [M][3] est,min,max */,
                                         const orc_matcher_options* opt, orc_match_out* out, int n_threads);
/* d */
int orc_update_filter_vogiatzis(double z, double tau2, double mu_range, double state[4]);
int orc_update_filter_gaussian(double z, double tau2, double state[4]);
void orc_update_filter_vogiatzis_batch(int n, const double* z, const double* tau2, const double* mu_range, double* state,
                                       uint8_t* ok, int n_threads);
double orc_compute_tau(const double T_ref_cur[7], const double f[3], double z, double px_error_angle);
double orc_px_error_angle(const orc_frame* frame, double px_noise);
/* S seeds of one ref frame observed, in order, by n_obs cur frames. types/states are in/out.
 * match_results (optional) is [n_obs][S]; success (optional) is [n_obs][S]. */
int orc_update_seeds(const orc_frame* ref, int n_obs, const orc_frame* cur_frames, const double* T_cur_ref /* [n_obs][7] */,
                     int S, const orc_feature* ftrs, uint8_t* types, double* states /* [S][4] */, double seed_mu_range,
                     const orc_matcher_options* opt, double sigma2_convergence_threshold,
                     double mappoint_sigma2_convergence_threshold, double px_error_angle,
                     int check_visibility, int check_convergence, int use_vogiatzis, int* match_results, uint8_t* success,
                     int n_threads);

/* f1: Reprojector candidate matching (src/svo/src/reprojector.cpp:342-543). Flattened map tables: the keyframes' feature SoA
 * columns and the landmarks' bookkeeping, exactly the members the reference reads. */
typedef struct {
  int n_kfs;
  const orc_frame* kfs;             /* [K] pyramids, camera, T_cam_imu / T_imu_world (T_f_w = T_cam_imu * T_imu_world) */
  const double* kf_seed_mu_range;   /* [K] Frame::seed_mu_range_ */
  const int* kf_feat_begin;         /* [K+1] features of keyframe k are the global features [begin[k], begin[k+1]) */
  const orc_feature* feat;          /* [NF] px, f, grad, level, type */
  const double* feat_score;         /* [NF] score_vec_ */
  const double* feat_seed_state;    /* [NF][4] invmu_sigma2_a_b_vec_ */
  const int* feat_point;            /* [NF] landmark id, -1 = landmark_vec_[i] == nullptr */
  const int* feat_kf;               /* [NF] keyframe of the feature */
  int n_points;
  const double* pt_pos;             /* [P][3] Point::pos_ */
  const int* pt_n_failed;           /* [P] n_failed_reproj_ */
  const int* pt_n_succeeded;        /* [P] n_succeeded_reproj_ */
  const int* pt_obs_begin;          /* [P+1] Point::obs_ of point p = obs_feat[begin[p] .. begin[p+1]) */
  const int* obs_feat;              /* [NO] global feature index of each observation */
} orc_reproj_map;

typedef struct {
  int cell_size;
  int max_n_features;               /* matchCandidates' max_n_features_per_frame (0 = unlimited, occupancy ignored) */
  int affine_est_offset, affine_est_gain;
  int sort_by_num_obs;              /* 0 = sortCandidatesByReprojStats, 1 = sortCandidatesByNumObs */
  double seed_sigma2_thresh;
  double px_error_angle;            /* updateSeed's function-static (depth_filter.cpp:383-384) */
} orc_reproj_options;

enum { ORC_REPROJ_NOT_CANDIDATE = 0, ORC_REPROJ_NOT_REACHED = 1, ORC_REPROJ_SKIPPED = 2, ORC_REPROJ_FAILED = 3, ORC_REPROJ_MATCHED = 4 };

typedef struct {
  double cur_px[2];                 /* Candidate::cur_px (projection into the current frame) */
  double px[2];                     /* matched: feature.px / f / grad written into the current frame's slot */
  double f[3];
  double grad[2];
  double seed_state[4];             /* the ref feature's seed state after the call (updated for unconverged seeds that were tried) */
  int status;
  int order;                        /* position in the sorted candidate list, -1 = not a candidate */
  int slot;                         /* matched: feature slot in the current frame */
  int level;                        /* matched: matcher.search_level_ */
  int type_out;                     /* the ref feature's type after the call (updateSeed may converge it or make it an outlier) */
  int match_result;                 /* Matcher::MatchResult of the attempt, -1 = none */
  int d_failed, d_succeeded;        /* increments of the landmark's n_failed_reproj_ / n_succeeded_reproj_ */
} orc_reproj_result;

typedef struct { int n_candidates, n_trials, n_matches, n_consumed; } orc_reproj_stats;

/* One current frame: getCandidate for the E entries (global feature indices, in the reference's visiting order), sort,
 * matchCandidates. occupancy [n_cells] in/out; results [E]; n_features_in = frame->num_features_ before the call. */
int orc_reproject_match(const orc_reproj_map* map, const orc_frame* cur, int E, const int* entry_feat, int n_features_in,
                        uint8_t* occupancy, const orc_reproj_options* opt, orc_reproj_result* results, orc_reproj_stats* stats);

/* f4: PoseOptimizer::run (src/svo/src/pose_optimizer.cpp). One frame bundle: frames[c] carry camera c and T_cam_imu; the state
 * starts at frames[0].T_imu_world. N features over all cameras: ftrs (px, f, grad, level, type), feat_cam, xyz_world [N][3],
 * has_xyz [N] (landmark or corner/edgelet seed reference present). stats = (measurement sigma, median error before, after
 * [px-equivalent], GN iterations, n_meas, chi2). Returns n_meas - deleted outliers. */
typedef struct {
  int err_type;            /* 0 kUnitPlane, 1 kBearingVectorDiff, 2 kImagePlane */
  int max_iter;
  double eps;
  double reproj_thresh_px;
  int have_prior;          /* setRotationPrior(R_frame_world, lambda) */
  double prior_q[4];
  double prior_lambda;
} orc_pose_opt_options;
int orc_pose_optimize(int n_cams, const orc_frame* frames, int N, const orc_feature* ftrs, const int* feat_cam, const double* xyz_world,
                      const uint8_t* has_xyz, const orc_pose_opt_options* opt, double T_imu_world_out[7], uint8_t* outlier, double stats[6]);

/* f3: alignPyr2D for M features sharing the two pyramids; px_ref_level_0 int [M][2]; px_cur [M][2] in/out; status [M] */
void orc_align_pyr2d(const orc_frame* ref, const orc_frame* cur, int max_level, int min_level, const int* patch_sizes, int n_iter,
                     float min_update_squared, int M, const int* px_ref_level_0, double* px_cur, uint8_t* status, int n_threads);
/* small pieces exposed one by one so that tests can pin them against the compiled reference (oracle/_ref/libdirect_ref.so) */
void orc_tukey_weight(float b, const float* err, int n, float* w);
/* which: 0 distort, 1 undistort, 2 jacobian (jac_out [n][4] = J00, J01, J10, J11) */
void orc_radtan(double k1, double k2, double p1, double p2, int which, double* xy, int n, double* jac_out);
/* out[0..5] = getDepth, getInvMinDepth, getInvMaxDepth, isConverged, getSigma2FromDepthSigma, getInitSigma2FromMuRange */
void orc_seed_helpers(const double* state4, double mu_range, double sigma2_convergence_threshold, double depth, double depth_sigma, double* out);
void orc_grid_cell_index(int cell_size, int n_cols, const int* xy, const int* scale, int n, long long* idx);
void orc_patch_from_patch_with_border(const uint8_t* patch_with_border, int patch_size, uint8_t* patch);

#ifdef __cplusplus
}
#endif
#endif
