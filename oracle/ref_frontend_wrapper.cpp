// ORACLE — TEST INFRASTRUCTURE ONLY.
// C wrapper around the REFERENCE's own SparseImgAlign, compiled from where it lies under /root/reference into
// oracle/_ref/libfrontend_ref.so by oracle/Makefile:
//   src/svo_img_align/src/sparse_img_align.cpp, sparse_img_align_base.cpp            (run, caches, residuals, H/g, update, prior)
//   src/vikit/vikit_solver/include/vikit/solver/implementation/mini_least_squares_solver.hpp   (Gauss-Newton driver)
//   src/vikit/vikit_solver/src/robust_cost.cpp                                         (Tukey weights)
//   src/vikit/vikit_cameras/include/vikit/cameras/{camera_geometry*, pinhole_projection, radial_tangential_distortion}
//   3rd/minkindr/include/kindr/minimal/*                                               (SE3 compose / inverse / exp / log)
//   src/svo_direct/src/patch_warp.cpp                                                  (getWarpMatrixAffine, getBestSearchLevel, warpAffine)
//   src/svo_direct/src/matcher.cpp, feature_alignment.cpp                              (findMatchDirect, findEpipolarMatchDirect, epipolar scans,
//                                                                                       triangulation, align1D / align2D)
//   src/svo_direct/src/depth_filter.cpp                                                (updateSeed, updateFilterVogiatzis / Gaussian, computeTau)
// Eigen, OpenCV and glog resolve to the container-only stand-ins in oracle/shim (Eigen's LDLT / quaternion / small-matrix
// arithmetic is restated there, see shim_eigen.hpp); svo::Frame / FrameBundle / Point are the reduced classes of
// oracle/shim/svo_fake (same members, names and accessor semantics as the reference's). No reference source is copied here.
// The structs are the oracle's own C API types, so tests feed the restatement and the compiled reference identically.
#include <svo/img_align/sparse_img_align.h>
#include <svo/direct/matcher.h>
#include <svo/direct/patch_warp.h>
#include <svo/direct/patch_score.h>
#include <svo/direct/depth_filter.h>
#include <svo/direct/feature_detection_utils.h>
#include <svo/common/frame.h>
#include <svo/common/camera.h>
#include <svo/common/occupancy_grid_2d.h>
#include <svo/reprojector.h>
#include <svo/pose_optimizer.h>
#include <cstring>
#include "orc_capi.h"

namespace {

using svo::Transformation;

Transformation toT(const double* a) {  // (qw qx qy qz tx ty tz)
  return Transformation(svo::Quaternion(a[0], a[1], a[2], a[3]), Eigen::Vector3d(a[4], a[5], a[6]));
}
void fromT(const Transformation& T, double* a) {
  const auto& q = T.getRotation().toImplementation();
  a[0] = q.w(); a[1] = q.x(); a[2] = q.y(); a[3] = q.z();
  a[4] = T.getPosition()[0]; a[5] = T.getPosition()[1]; a[6] = T.getPosition()[2];
}

svo::CameraPtr makeCamera(const orc_frame& f) {
  using namespace vk::cameras;
  if (f.distortion == 0) {
    typedef PinholeProjection<NoDistortion> P;
    return std::make_shared<CameraGeometry<P>>(f.width, f.height, P(f.cam[0], f.cam[1], f.cam[2], f.cam[3], NoDistortion()));
  }
  typedef PinholeProjection<RadialTangentialDistortion> P;
  return std::make_shared<CameraGeometry<P>>(f.width, f.height,
                                             P(f.cam[0], f.cam[1], f.cam[2], f.cam[3], RadialTangentialDistortion(f.cam[4], f.cam[5], f.cam[6], f.cam[7])));
}

svo::FramePtr makeFrame(const orc_frame& f) {
  auto fr = std::make_shared<svo::Frame>();
  fr->cam_ = makeCamera(f);
  for (int l = 0; l < f.n_levels; ++l)
    fr->img_pyr_.emplace_back(f.level_rows[l], f.level_cols[l], CV_8UC1, const_cast<uint8_t*>(f.level_data[l]), (size_t)f.level_step[l]);
  fr->set_T_cam_imu(toT(f.T_cam_imu));
  fr->T_f_w_ = fr->T_cam_imu() * toT(f.T_imu_world);
  const int n = f.px ? f.n_features : 0;
  fr->resizeFeatureStorage(n);
  fr->num_features_ = n;
  const Transformation T_world_cam = fr->T_world_cam();
  for (int i = 0; i < n; ++i) {
    fr->px_vec_.col(i) = Eigen::Vector2d(f.px[2 * i], f.px[2 * i + 1]);
    const Eigen::Vector3d bearing(f.f[3 * i], f.f[3 * i + 1], f.f[3 * i + 2]);
    fr->f_vec_.col(i) = bearing;
    fr->type_vec_[i] = svo::FeatureType::kCorner;
    // the reference reads the depth back as |landmark - camera centre| (sparse_img_align.cpp:283-286)
    if (f.eligible[i]) fr->landmark_vec_[i] = std::make_shared<svo::Point>(T_world_cam * Eigen::Vector3d(bearing * f.depth[i]));
  }
  return fr;
}

struct Probe : public svo::SparseImgAlign {  // read access to the solver's protected counters
  using svo::SparseImgAlign::SparseImgAlign;
  size_t lastIter() const { return iter_; }
  bool stopped() const { return stop_; }
};

}  // namespace

extern "C" int ref_sparse_align(int n_cams, const orc_frame* ref, const orc_frame* cur, const orc_align_options* o, orc_align_result* res) {
  std::vector<svo::FramePtr> rf, cf;
  for (int c = 0; c < n_cams; ++c) { rf.push_back(makeFrame(ref[c])); cf.push_back(makeFrame(cur[c])); }
  auto rb = std::make_shared<svo::FrameBundle>(rf), cb = std::make_shared<svo::FrameBundle>(cf);
  svo::SparseImgAlignOptions opt;
  opt.max_level = o->max_level; opt.min_level = o->min_level;
  opt.estimate_illumination_gain = o->estimate_illumination_gain != 0;
  opt.estimate_illumination_offset = o->estimate_illumination_offset != 0;
  opt.use_distortion_jacobian = o->use_distortion_jacobian != 0;
  opt.robustification = o->robustification != 0;
  opt.weight_scale = o->weight_scale;
  svo::SparseImgAlignBase::SolverOptions so = svo::SparseImgAlignBase::getDefaultSolverOptions();
  so.max_iter = o->max_iter;
  so.eps = o->eps;
  Probe aligner(so, opt);
  aligner.reset();  // callers always reset() first (frame_handler_base.cpp:621)
  aligner.setAlphaInitialValue(o->alpha_init);
  aligner.setBetaInitialValue(o->beta_init);
  if (o->have_prior)
    aligner.setWeightedPrior(toT(o->prior_T), o->prior_alpha, o->prior_beta, o->lambda_rot, o->lambda_trans, o->lambda_alpha, o->lambda_beta);
  std::memset(res, 0, sizeof(*res));
  res->n_tracked = (int)aligner.run(rb, cb);
  // T_f_w_ = T_cam_imu * T_icur_iref * T_iref_world  (sparse_img_align.cpp:102-106)
  const Transformation T_icur_iref = cf[0]->T_imu_cam() * cf[0]->T_f_w_ * rf[0]->T_imu_world().inverse();
  fromT(T_icur_iref, res->T_icur_iref);
  for (int c = 0; c < n_cams; ++c) fromT(cf[c]->T_f_w_, res->T_f_w[c]);
  res->chi2 = aligner.getError();
  const auto& H = aligner.getHessian();
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) res->H[8 * i + j] = H(i, j);
  res->stop = aligner.stopped() ? 1 : 0;
  return res->n_tracked;
}

// One "frame pair step" as bench.py defines it, on the REFERENCE's own code: the pyramid of the new frame with vk::halfSample
// in the loop of frame_utils::createImgPyramid (src/svo_common/src/frame.cpp:372-386: level 0 is the image, level i = halfSample
// of level i-1 into a fresh (rows/2, cols/2) Mat), then SparseImgAlign::run. B independent pairs on n_threads std::threads
// (the reference runs one pair per call on one thread; the batch only keeps every host core busy with such calls).
#include <atomic>
#include <thread>
#include <vikit/vision.h>
extern "C" int ref_pyramid_align_batch(int B, int n_levels, const uint8_t* const* cur_l0, int cols, int rows, const orc_frame* ref,
                                       const orc_frame* cur, const orc_align_options* opt, orc_align_result* res, int n_threads) {
  std::atomic<int> next(0);
  auto work = [&]() {
    for (int i = next.fetch_add(1); i < B; i = next.fetch_add(1)) {
      std::vector<cv::Mat> pyr(n_levels);
      pyr[0] = cv::Mat(rows, cols, CV_8UC1, const_cast<uint8_t*>(cur_l0[i]), (size_t)cols).clone();  // frame_handler_base.cpp:184-186
      for (int l = 1; l < n_levels; ++l) {
        pyr[l] = cv::Mat(pyr[l - 1].rows / 2, pyr[l - 1].cols / 2, CV_8U);
        vk::halfSample(pyr[l - 1], pyr[l]);
      }
      orc_frame cf = cur[i];
      cf.n_levels = n_levels;
      for (int l = 0; l < n_levels; ++l) {
        cf.level_data[l] = pyr[l].data;
        cf.level_cols[l] = pyr[l].cols;
        cf.level_rows[l] = pyr[l].rows;
        cf.level_step[l] = (int)pyr[l].step;
      }
      ref_sparse_align(1, ref + i, &cf, opt, res + i);
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return 0;
}

// ---- (c) matcher ----------------------------------------------------------------------------------------------------------
namespace {

// The oracle's C API hands T_cur_ref over explicitly; the reference derives it from the frames' poses
// (T_cur_ref = cur.T_f_w_ * ref.T_f_w_^-1, matcher.cpp:41, :166), so the ref frame sits at the identity.
void poseFrames(svo::Frame& rf, svo::Frame& cf, const double* T_cur_ref) {
  rf.T_f_w_ = Transformation();
  cf.T_f_w_ = toT(T_cur_ref);
}
void setOptions(svo::Matcher& m, const orc_matcher_options* o) {
  m.options_.align_1d = o->align_1d != 0;
  m.options_.align_max_iter = o->align_max_iter;
  m.options_.max_epi_search_steps = o->max_epi_search_steps;
  m.options_.subpix_refinement = o->subpix_refinement != 0;
  m.options_.epi_search_edgelet_filtering = o->epi_search_edgelet_filtering != 0;
  m.options_.scan_on_unit_sphere = o->scan_on_unit_sphere != 0;
  m.options_.epi_search_edgelet_max_angle = o->epi_search_edgelet_max_angle;
  m.options_.affine_est_offset_ = o->affine_est_offset != 0;
  m.options_.affine_est_gain_ = o->affine_est_gain != 0;
  m.options_.max_patch_diff_ratio = o->max_patch_diff_ratio;
}
void fillOut(const svo::Matcher& m, svo::Matcher::MatchResult r, double depth, orc_match_out* out) {
  out->result = int(r);
  out->px_cur[0] = m.px_cur_[0]; out->px_cur[1] = m.px_cur_[1];
  out->f_cur[0] = m.f_cur_[0]; out->f_cur[1] = m.f_cur_[1]; out->f_cur[2] = m.f_cur_[2];
  out->search_level = m.search_level_;
  out->A_cur_ref[0] = m.A_cur_ref_(0, 0); out->A_cur_ref[1] = m.A_cur_ref_(0, 1);
  out->A_cur_ref[2] = m.A_cur_ref_(1, 0); out->A_cur_ref[3] = m.A_cur_ref_(1, 1);
  out->h_inv = m.h_inv_;
  out->epi_length_pyramid = m.epi_length_pyramid_;
  out->reject = m.reject_;
  out->depth = depth;
  std::memcpy(out->patch_with_border, m.patch_with_border_, 100);
  out->epi_image[0] = m.epi_image_[0]; out->epi_image[1] = m.epi_image_[1];
}
// a one-feature frame around the reference feature (FeatureWrapper binds to the frame's SoA columns)
void setFeature(svo::Frame& rf, const orc_feature* f) {
  rf.resizeFeatureStorage(1);
  rf.num_features_ = 1;
  rf.px_vec_.col(0) = Eigen::Vector2d(f->px[0], f->px[1]);
  rf.f_vec_.col(0) = Eigen::Vector3d(f->f[0], f->f[1], f->f[2]);
  rf.grad_vec_.col(0) = Eigen::Vector2d(f->grad[0], f->grad[1]);
  rf.type_vec_[0] = static_cast<svo::FeatureType>(f->type);
  rf.level_vec_(0) = f->level;
}
void initMatcher(svo::Matcher& m) {  // the members the reference leaves uninitialised are reported as zeros
  std::memset(m.patch_, 0, sizeof(m.patch_));
  std::memset(m.patch_with_border_, 0, sizeof(m.patch_with_border_));
  m.A_cur_ref_.setZero(); m.epi_image_.setZero(); m.px_cur_.setZero(); m.f_cur_.setZero();
  m.epi_length_pyramid_ = 0; m.h_inv_ = 0; m.search_level_ = 0; m.reject_ = false;
}

}  // namespace

extern "C" {

int ref_find_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                          double ref_depth, const double px_cur_in[2], const orc_matcher_options* opt, orc_match_out* out) {
  svo::FramePtr rf = makeFrame(*ref), cf = makeFrame(*cur);
  poseFrames(*rf, *cf, T_cur_ref);
  setFeature(*rf, ftr);
  svo::Matcher m;
  initMatcher(m);
  setOptions(m, opt);
  svo::Keypoint px(px_cur_in[0], px_cur_in[1]);
  m.px_cur_ = px;
  svo::FeatureWrapper fw = rf->getFeatureWrapper(0);
  const svo::Matcher::MatchResult r = m.findMatchDirect(*rf, *cf, fw, ref_depth, px);
  fillOut(m, r, 0.0, out);
  return int(r);
}

int ref_find_epipolar_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                                   double d_estimate_inv, double d_min_inv, double d_max_inv, const orc_matcher_options* opt,
                                   orc_match_out* out) {
  svo::FramePtr rf = makeFrame(*ref), cf = makeFrame(*cur);
  poseFrames(*rf, *cf, T_cur_ref);
  setFeature(*rf, ftr);
  svo::Matcher m;
  initMatcher(m);
  setOptions(m, opt);
  double depth = 0.0;
  svo::FeatureWrapper fw = rf->getFeatureWrapper(0);
  const svo::Matcher::MatchResult r = m.findEpipolarMatchDirect(*rf, *cf, toT(T_cur_ref), fw, d_estimate_inv, d_min_inv, d_max_inv, depth);
  fillOut(m, r, depth, out);
  return int(r);
}

void ref_scan_epipolar_line(const orc_frame* cur, const double A[3], const double B[3], const double C[3], const uint8_t* patch64,
                            int patch_level, double epi_length_pyramid, const orc_matcher_options* opt, double image_best[2],
                            int* zmssd_best) {
  svo::FramePtr cf = makeFrame(*cur);
  svo::Matcher m;
  initMatcher(m);
  setOptions(m, opt);
  m.epi_length_pyramid_ = epi_length_pyramid;
  std::memcpy(m.patch_, patch64, 64);
  svo::Matcher::PatchScore patch_score(m.patch_);
  svo::Keypoint best(0.0, 0.0);
  m.scanEpipolarLine(*cf, Eigen::Vector3d(A[0], A[1], A[2]), Eigen::Vector3d(B[0], B[1], B[2]), Eigen::Vector3d(C[0], C[1], C[2]),
                     patch_score, patch_level, &best, zmssd_best);
  image_best[0] = best[0]; image_best[1] = best[1];
}

void ref_get_warp_matrix_affine(const orc_frame* ref, const orc_frame* cur, const double px_ref[2], const double f_ref[3],
                                double depth_ref, const double T_cur_ref[7], int level_ref, double A_out[4]) {
  svo::CameraPtr cr = makeCamera(*ref), cc = makeCamera(*cur);
  svo::Keypoint px(px_ref[0], px_ref[1]);
  svo::BearingVector f(f_ref[0], f_ref[1], f_ref[2]);
  svo::warp::AffineTransformation2 A;
  svo::warp::getWarpMatrixAffine(cr, cc, px, f, depth_ref, toT(T_cur_ref), level_ref, &A);
  A_out[0] = A(0, 0); A_out[1] = A(0, 1); A_out[2] = A(1, 0); A_out[3] = A(1, 1);
}

int ref_get_best_search_level(const double A[4], int max_level) {
  svo::warp::AffineTransformation2 M;
  M << A[0], A[1], A[2], A[3];
  return svo::warp::getBestSearchLevel(M, max_level);
}

int ref_warp_affine(const double A_cur_ref[4], const uint8_t* img, int cols, int rows, int step, const double px_ref[2],
                    int level_ref, int search_level, int halfpatch_size, uint8_t* patch) {
  svo::warp::AffineTransformation2 M;
  M << A_cur_ref[0], A_cur_ref[1], A_cur_ref[2], A_cur_ref[3];
  cv::Mat im(rows, cols, CV_8UC1, const_cast<uint8_t*>(img), (size_t)step);
  svo::Keypoint px(px_ref[0], px_ref[1]);
  return svo::warp::warpAffine(M, im, px, level_ref, search_level, halfpatch_size, patch) ? 1 : 0;
}

}  // extern "C"

// ---- (d) depth filter -----------------------------------------------------------------------------------------------------
// (DepthFilter's constructor names the detector factory: the reference's own makeDetector is linked in, see f3 below.)

extern "C" {

int ref_update_filter_vogiatzis(double z, double tau2, double mu_range, double state[4]) {
  svo::SeedState s;
  s << state[0], state[1], state[2], state[3];
  bool ok;
  {
    Eigen::Ref<svo::SeedState> r(s);
    ok = svo::depth_filter_utils::updateFilterVogiatzis(z, tau2, mu_range, r);
  }
  for (int i = 0; i < 4; ++i) state[i] = s[i];
  return ok ? 1 : 0;
}

int ref_update_filter_gaussian(double z, double tau2, double state[4]) {
  svo::SeedState s;
  s << state[0], state[1], state[2], state[3];
  bool ok;
  {
    Eigen::Ref<svo::SeedState> r(s);
    ok = svo::depth_filter_utils::updateFilterGaussian(z, tau2, r);
  }
  for (int i = 0; i < 4; ++i) state[i] = s[i];
  return ok ? 1 : 0;
}

double ref_compute_tau(const double T_ref_cur[7], const double f[3], double z, double px_error_angle) {
  return svo::depth_filter_utils::computeTau(toT(T_ref_cur), svo::BearingVector(f[0], f[1], f[2]), z, px_error_angle);
}

double ref_px_error_angle(const orc_frame* frame, double px_noise) { return makeCamera(*frame)->getAngleError(px_noise); }

// S seeds of one ref frame observed, in order, by n_obs cur frames (same contract as orc_update_seeds). The reference keeps
// px_error_angle in a function-local static initialised from the first camera it sees (depth_filter.cpp:383-384): every call
// of one process must therefore use the same camera intrinsics.
int ref_update_seeds(const orc_frame* ref, int n_obs, const orc_frame* cur_frames, const double* T_cur_ref, int S,
                     const orc_feature* ftrs, uint8_t* types, double* states, double seed_mu_range,
                     const orc_matcher_options* opt, double sigma2_convergence_threshold,
                     double mappoint_sigma2_convergence_threshold, int check_visibility, int check_convergence,
                     int use_vogiatzis, int* match_results, uint8_t* success) {
  svo::FramePtr rf = makeFrame(*ref);
  rf->id_ = 0;
  rf->T_f_w_ = Transformation();
  rf->resizeFeatureStorage(S);
  rf->num_features_ = S;
  rf->seed_mu_range_ = seed_mu_range;
  for (int s = 0; s < S; ++s) {
    rf->px_vec_.col(s) = Eigen::Vector2d(ftrs[s].px[0], ftrs[s].px[1]);
    rf->f_vec_.col(s) = Eigen::Vector3d(ftrs[s].f[0], ftrs[s].f[1], ftrs[s].f[2]);
    rf->grad_vec_.col(s) = Eigen::Vector2d(ftrs[s].grad[0], ftrs[s].grad[1]);
    rf->level_vec_(s) = ftrs[s].level;
    rf->type_vec_[s] = static_cast<svo::FeatureType>(types[s]);
    for (int k = 0; k < 4; ++k) rf->invmu_sigma2_a_b_vec_(k, s) = states[4 * s + k];
  }
  std::vector<svo::FramePtr> cfs;
  for (int o = 0; o < n_obs; ++o) {
    cfs.push_back(makeFrame(cur_frames[o]));
    cfs.back()->id_ = o + 1;
    cfs.back()->T_f_w_ = toT(T_cur_ref + 7 * o);
  }
  int n_success = 0;
  for (int s = 0; s < S; ++s) {
    svo::Matcher m;
    initMatcher(m);
    setOptions(m, opt);
    for (int o = 0; o < n_obs; ++o) {
      // DepthFilter::updateSeeds picks the threshold by seed type (depth_filter.cpp:214-221)
      const svo::FeatureType type = rf->type_vec_[s];
      const double thresh = svo::isMapPointSeed(type) ? mappoint_sigma2_convergence_threshold : sigma2_convergence_threshold;
      const bool ok = svo::depth_filter_utils::updateSeed(*cfs[o], *rf, (size_t)s, m, thresh, check_visibility != 0,
                                                          check_convergence != 0, use_vogiatzis != 0);
      if (success) success[(size_t)o * S + s] = ok;
      (void)match_results;
      if (ok) ++n_success;
    }
  }
  for (int s = 0; s < S; ++s) {
    types[s] = static_cast<uint8_t>(rf->type_vec_[s]);
    for (int k = 0; k < 4; ++k) states[4 * s + k] = rf->invmu_sigma2_a_b_vec_(k, s);
  }
  return n_success;
}

}  // extern "C"

// ---- (f1) Reprojector candidate flow ------------------------------------------------------------------------------------------
// The reference's own reprojector.cpp (src/svo/src/reprojector.cpp, compiled unmodified): getCandidate for every entry in
// visiting order, sortCandidatesByReprojStats / sortCandidatesByNumObs, matchCandidates. The keyframes, their feature columns
// and the landmarks with their observation lists are rebuilt from the flat tables as real svo::Frame / svo::Point objects.
namespace {
struct RefMap {
  std::vector<svo::FramePtr> kfs;
  std::vector<svo::PointPtr> pts;
};
RefMap buildRefMap(const orc_reproj_map* map) {
  using namespace svo;
  RefMap m;
  for (int k = 0; k < map->n_kfs; ++k) {
    orc_frame f = map->kfs[k];
    f.px = nullptr;  // features come from the tables below
    FramePtr fr = makeFrame(f);
    fr->id_ = k + 1;
    const int b = map->kf_feat_begin[k], n = map->kf_feat_begin[k + 1] - b;
    fr->resizeFeatureStorage(n);
    fr->num_features_ = n;
    fr->seed_mu_range_ = map->kf_seed_mu_range[k];
    for (int i = 0; i < n; ++i) {
      const orc_feature& q = map->feat[b + i];
      fr->px_vec_.col(i) = Eigen::Vector2d(q.px[0], q.px[1]);
      fr->f_vec_.col(i) = Eigen::Vector3d(q.f[0], q.f[1], q.f[2]);
      fr->grad_vec_.col(i) = Eigen::Vector2d(q.grad[0], q.grad[1]);
      fr->level_vec_(i) = q.level;
      fr->type_vec_[i] = static_cast<FeatureType>(q.type);
      fr->score_vec_(i) = map->feat_score[b + i];
      for (int c = 0; c < 4; ++c) fr->invmu_sigma2_a_b_vec_(c, i) = map->feat_seed_state[4 * size_t(b + i) + c];
    }
    m.kfs.push_back(fr);
  }
  for (int p = 0; p < map->n_points; ++p) {
    auto pt = std::make_shared<Point>(Eigen::Vector3d(map->pt_pos[3 * p], map->pt_pos[3 * p + 1], map->pt_pos[3 * p + 2]));
    pt->id_ = p;
    pt->n_failed_reproj_ = map->pt_n_failed[p];
    pt->n_succeeded_reproj_ = map->pt_n_succeeded[p];
    for (int o = map->pt_obs_begin[p]; o < map->pt_obs_begin[p + 1]; ++o) {
      const int fi = map->obs_feat[o], k = map->feat_kf[fi];
      pt->obs_.emplace_back(m.kfs[k], size_t(fi - map->kf_feat_begin[k]));
    }
    m.pts.push_back(pt);
  }
  const int n_feat = map->kf_feat_begin[map->n_kfs];
  for (int fi = 0; fi < n_feat; ++fi)
    if (map->feat_point[fi] >= 0) {
      const int k = map->feat_kf[fi];
      m.kfs[k]->landmark_vec_[fi - map->kf_feat_begin[k]] = m.pts[map->feat_point[fi]];
    }
  return m;
}
}  // namespace

extern "C" int ref_reproject_match(const orc_reproj_map* map, const orc_frame* cur, int E, const int* entry_feat, int n_features_in,
                                   uint8_t* occupancy, const orc_reproj_options* opt, orc_reproj_result* results,
                                   orc_reproj_stats* stats) {
  using namespace svo;
  RefMap rm = buildRefMap(map);
  std::vector<FramePtr>& kfs = rm.kfs;
  std::vector<PointPtr>& pts = rm.pts;
  const int n_feat = map->kf_feat_begin[map->n_kfs];
  orc_frame cf = *cur;
  cf.px = nullptr;
  FramePtr frame = makeFrame(cf);
  frame->id_ = 1000;
  const size_t cap = size_t(n_features_in) + size_t(E) + 1;
  frame->resizeFeatureStorage(cap);
  frame->num_features_ = n_features_in;
  for (int i = 0; i < n_features_in; ++i) frame->type_vec_[i] = FeatureType::kCorner;

  OccupandyGrid2D grid(opt->cell_size, OccupandyGrid2D::getNCell(frame->cam()->imageWidth(), opt->cell_size),
                       OccupandyGrid2D::getNCell(frame->cam()->imageHeight(), opt->cell_size));
  grid.reset();
  for (size_t c = 0; c < grid.occupancy_.size(); ++c) grid.occupancy_[c] = occupancy[c] != 0;
  const std::vector<bool> occ_in = grid.occupancy_;

  std::vector<int> entry_of_feat(n_feat, -1);
  Reprojector::Candidates candidates;
  for (int e = 0; e < E; ++e) {
    const int fi = entry_feat[e], k = map->feat_kf[fi];
    entry_of_feat[fi] = e;
    orc_reproj_result& r = results[e];
    std::memset(&r, 0, sizeof(r));
    r.status = ORC_REPROJ_NOT_CANDIDATE; r.order = -1; r.slot = -1; r.match_result = -1;
    Reprojector::Candidate c;
    if (reprojector_utils::getCandidate(frame, kfs[k], size_t(fi - map->kf_feat_begin[k]), c)) {
      r.cur_px[0] = c.cur_px[0]; r.cur_px[1] = c.cur_px[1];
      candidates.push_back(c);
    }
  }
  if (opt->sort_by_num_obs) reprojector_utils::sortCandidatesByNumObs(candidates);
  else reprojector_utils::sortCandidatesByReprojStats(candidates);
  auto featOf = [&](const Reprojector::Candidate& c) {
    int k = 0;
    while (kfs[k].get() != c.ref_frame.get()) ++k;
    return map->kf_feat_begin[k] + int(c.ref_index);
  };
  const Reprojector::Candidates sorted = candidates;
  for (size_t p = 0; p < sorted.size(); ++p) {
    orc_reproj_result& r = results[entry_of_feat[featOf(sorted[p])]];
    r.order = int(p);
    r.status = ORC_REPROJ_NOT_REACHED;
  }
  Reprojector::Statistics st;
  reprojector_utils::matchCandidates(frame, size_t(opt->max_n_features), opt->affine_est_offset != 0, opt->affine_est_gain != 0,
                                     candidates, grid, st, opt->seed_sigma2_thresh);
  stats->n_candidates = int(sorted.size());
  stats->n_trials = int(st.n_trials);
  stats->n_matches = int(st.n_matches);
  stats->n_consumed = int(sorted.size() - candidates.size());
  // which candidate filled which new slot: landmarks by feature.landmark, seeds by feature.seed_ref
  for (size_t s = size_t(n_features_in); s < frame->num_features_; ++s) {
    int fi = -1;
    for (int e = 0; e < E && fi < 0; ++e) {
      const int g = entry_feat[e], k = map->feat_kf[g];
      if (results[e].order < 0 || results[e].order >= stats->n_consumed) continue;
      if (map->feat_point[g] >= 0) { if (frame->landmark_vec_[s] == pts[map->feat_point[g]]) fi = g; }
      else if (frame->seed_ref_vec_[s].keyframe == kfs[k] && frame->seed_ref_vec_[s].seed_id == g - map->kf_feat_begin[k]) fi = g;
    }
    if (fi < 0) continue;
    orc_reproj_result& r = results[entry_of_feat[fi]];
    r.status = ORC_REPROJ_MATCHED;
    r.slot = int(s);
    r.px[0] = frame->px_vec_(0, s); r.px[1] = frame->px_vec_(1, s);
    r.f[0] = frame->f_vec_(0, s); r.f[1] = frame->f_vec_(1, s); r.f[2] = frame->f_vec_(2, s);
    if (isEdgelet(frame->type_vec_[s])) { r.grad[0] = frame->grad_vec_(0, s); r.grad[1] = frame->grad_vec_(1, s); }
    r.level = frame->level_vec_(s);
  }
  // consumed candidates that did not match: skipped when their cell was occupied at their turn, tried (and failed) otherwise
  std::vector<bool> occ = occ_in;
  for (int p = 0; p < stats->n_consumed; ++p) {
    orc_reproj_result& r = results[entry_of_feat[featOf(sorted[p])]];
    const size_t cell = grid.getCellIndex(sorted[p].cur_px.x(), sorted[p].cur_px.y(), 1);
    if (opt->max_n_features > 0 && occ[cell]) { r.status = ORC_REPROJ_SKIPPED; continue; }
    if (r.status == ORC_REPROJ_MATCHED) occ[cell] = true;
    else r.status = ORC_REPROJ_FAILED;
  }
  for (size_t c = 0; c < grid.occupancy_.size(); ++c) occupancy[c] = grid.occupancy_[c] ? 1 : 0;
  for (int e = 0; e < E; ++e) {
    const int fi = entry_feat[e], k = map->feat_kf[fi], i = fi - map->kf_feat_begin[k];
    for (int c = 0; c < 4; ++c) results[e].seed_state[c] = kfs[k]->invmu_sigma2_a_b_vec_(c, i);
    results[e].type_out = int(kfs[k]->type_vec_[i]);
    if (map->feat_point[fi] >= 0) {
      results[e].d_failed = pts[map->feat_point[fi]]->n_failed_reproj_ - map->pt_n_failed[map->feat_point[fi]];
      results[e].d_succeeded = pts[map->feat_point[fi]]->n_succeeded_reproj_ - map->pt_n_succeeded[map->feat_point[fi]];
    }
  }
  return stats->n_matches;
}

// The whole Reprojector::reprojectFrames (reprojector.cpp:28-310) of the reference on the same tables: the first n_visible
// keyframes are `visible_kfs` (in order). Outputs: the features the call appended to the current frame (slots 0..n-1), the grid,
// the statistics, the landmarks' counters, the keyframes' seed states / types afterwards and the number of trashed points.
extern "C" int ref_reproject_frames(const orc_reproj_map* map, int n_visible, const orc_frame* cur, int max_n_features_per_frame,
                                    int reproject_unconverged_seeds, double max_unconverged_seeds_ratio, int min_required_features,
                                    int remove_unconstrained_points, int* out_type, double* out_px, int* out_level, int* out_point,
                                    int* out_seed_feat, double* out_state, double* out_f, double* out_grad, double* out_score,
                                    uint8_t* occupancy_out, int* stats_out /* n_trials, n_matches, n_trash */, int* pt_counters_out,
                                    double* feat_state_out, int* feat_type_out) {
  using namespace svo;
  RefMap rm = buildRefMap(map);
  orc_frame cf = *cur;
  cf.px = nullptr;
  FramePtr frame = makeFrame(cf);
  frame->id_ = 1000;
  ReprojectorOptions o;
  o.max_n_features_per_frame = size_t(max_n_features_per_frame);
  o.reproject_unconverged_seeds = reproject_unconverged_seeds != 0;
  o.max_unconverged_seeds_ratio = max_unconverged_seeds_ratio;
  o.min_required_features = size_t(min_required_features);
  o.remove_unconstrained_points = remove_unconstrained_points != 0;
  Reprojector rp(o, 0);
  std::vector<FramePtr> visible(rm.kfs.begin(), rm.kfs.begin() + n_visible);
  std::vector<PointPtr> trash;
  rp.reprojectFrames(frame, visible, trash);
  const int n = int(frame->num_features_);
  for (int s = 0; s < n; ++s) {
    out_type[s] = int(frame->type_vec_[s]);
    out_px[2 * s] = frame->px_vec_(0, s); out_px[2 * s + 1] = frame->px_vec_(1, s);
    out_level[s] = frame->level_vec_(s);
    out_point[s] = frame->landmark_vec_[s] ? frame->landmark_vec_[s]->id() : -1;
    out_seed_feat[s] = -1;
    if (frame->seed_ref_vec_[s].keyframe) {
      int k = 0;
      while (rm.kfs[k].get() != frame->seed_ref_vec_[s].keyframe.get()) ++k;
      out_seed_feat[s] = map->kf_feat_begin[k] + frame->seed_ref_vec_[s].seed_id;
    }
    for (int c = 0; c < 4; ++c) out_state[4 * s + c] = frame->invmu_sigma2_a_b_vec_(c, s);
    for (int c = 0; c < 3; ++c) out_f[3 * s + c] = frame->f_vec_(c, s);
    const bool e = isEdgelet(frame->type_vec_[s]);
    out_grad[2 * s] = e ? frame->grad_vec_(0, s) : 0.0; out_grad[2 * s + 1] = e ? frame->grad_vec_(1, s) : 0.0;
    out_score[s] = frame->score_vec_(s);
  }
  for (size_t c = 0; c < rp.grid_->occupancy_.size(); ++c) occupancy_out[c] = rp.grid_->occupancy_[c] ? 1 : 0;
  stats_out[0] = int(rp.stats_.n_trials); stats_out[1] = int(rp.stats_.n_matches); stats_out[2] = int(trash.size());
  for (int p = 0; p < map->n_points; ++p) {
    pt_counters_out[2 * p] = rm.pts[p]->n_failed_reproj_; pt_counters_out[2 * p + 1] = rm.pts[p]->n_succeeded_reproj_;
  }
  const int n_feat = map->kf_feat_begin[map->n_kfs];
  for (int fi = 0; fi < n_feat; ++fi) {
    const int k = map->feat_kf[fi], i = fi - map->kf_feat_begin[k];
    for (int c = 0; c < 4; ++c) feat_state_out[4 * fi + c] = rm.kfs[k]->invmu_sigma2_a_b_vec_(c, i);
    feat_type_out[fi] = int(rm.kfs[k]->type_vec_[i]);
  }
  return n;
}

// ---- (f4) PoseOptimizer -------------------------------------------------------------------------------------------------------
// The reference's own pose_optimizer.cpp (src/svo/src/pose_optimizer.cpp, compiled unmodified) on a bundle rebuilt from the flat
// arrays: every feature with has_xyz gets a landmark at xyz_world (the seed branch of evaluateErrorImpl computes the same point
// from the seed's keyframe); the others keep landmark == nullptr and a non-seed type, i.e. they are skipped.
extern "C" int ref_pose_optimize(int n_cams, const orc_frame* frames, int N, const orc_feature* ftrs, const int* feat_cam,
                                 const double* xyz_world, const uint8_t* has_xyz, const orc_pose_opt_options* opt,
                                 double T_imu_world_out[7], uint8_t* outlier, double stats[6]) {
  using namespace svo;
  std::vector<FramePtr> fr;
  std::vector<std::vector<int>> idx(n_cams);
  for (int i = 0; i < N; ++i) idx[feat_cam[i]].push_back(i);
  for (int c = 0; c < n_cams; ++c) {
    orc_frame f = frames[c];
    f.px = nullptr;
    for (int k = 0; k < 7; ++k) f.T_imu_world[k] = frames[0].T_imu_world[k];
    FramePtr p = makeFrame(f);
    const int n = int(idx[c].size());
    p->resizeFeatureStorage(n);
    p->num_features_ = n;
    for (int j = 0; j < n; ++j) {
      const int i = idx[c][j];
      const orc_feature& q = ftrs[i];
      p->px_vec_.col(j) = Eigen::Vector2d(q.px[0], q.px[1]);
      p->f_vec_.col(j) = Eigen::Vector3d(q.f[0], q.f[1], q.f[2]);
      p->grad_vec_.col(j) = Eigen::Vector2d(q.grad[0], q.grad[1]);
      p->level_vec_(j) = q.level;
      p->type_vec_[j] = static_cast<FeatureType>(q.type);
      if (has_xyz[i]) p->landmark_vec_[j] = std::make_shared<Point>(Eigen::Vector3d(xyz_world[3 * i], xyz_world[3 * i + 1], xyz_world[3 * i + 2]));
    }
    fr.push_back(p);
  }
  auto bundle = std::make_shared<FrameBundle>(fr);
  PoseOptimizer::SolverOptions so = PoseOptimizer::getDefaultSolverOptions();
  so.max_iter = opt->max_iter;
  so.eps = opt->eps;
  PoseOptimizer po(so);
  po.reset();  // frame_handler_base.cpp:757
  po.setErrorType(static_cast<PoseOptimizer::ErrorType>(opt->err_type));
  if (opt->have_prior) po.setRotationPrior(svo::Quaternion(opt->prior_q[0], opt->prior_q[1], opt->prior_q[2], opt->prior_q[3]), opt->prior_lambda);
  const size_t n = po.run(bundle, opt->reproj_thresh_px);
  fromT(fr[0]->T_imu_world(), T_imu_world_out);
  for (int c = 0; c < n_cams; ++c)
    for (size_t j = 0; j < idx[c].size(); ++j) outlier[idx[c][j]] = fr[c]->type_vec_[j] == FeatureType::kOutlier && ftrs[idx[c][j]].type != int(FeatureType::kOutlier);
  stats[0] = po.measurement_sigma_; stats[1] = po.stats_.reproj_error_before; stats[2] = po.stats_.reproj_error_after;
  stats[3] = double(po.iterCount()); stats[4] = 0.0; stats[5] = po.getError();
  return int(n);
}

// ---- f3: StereoTriangulation::compute (src/svo/src/stereo_triangulation.cpp:23-139), the whole function: detection with the detector
// makeDetector builds, bearing vectors, the two std::random_shuffle calls (srand(seed) first, so that ref_stereo_shuffle_order
// reproduces the visiting order), the epipolar matching loop and the bookkeeping of both frames.
#include <svo/stereo_triangulation.h>
#include <svo/direct/feature_detection.h>
#include <svo/direct/feature_detection_utils.h>
#include <algorithm>
#include <numeric>

extern "C" void ref_stereo_shuffle_order(unsigned seed, int n_old, int n_corners, int n_new, int* order) {
  std::vector<size_t> indices(static_cast<size_t>(n_new));
  std::iota(indices.begin(), indices.end(), n_old);
  srand(seed);
  std::random_shuffle(indices.begin(), indices.begin() + n_corners);
  std::random_shuffle(indices.begin() + n_corners, indices.end());
  for (int i = 0; i < n_new; ++i) order[i] = (int)indices[i];
}

// Outputs: frame0's new feature columns (cap0 entries of room) and, per feature frame1 received, its columns, the landmark position
// and the frame0 feature index its landmark was created from. Returns frame1's feature count; *n0_out = frame0's feature count.
extern "C" int ref_stereo_triangulation_compute(const orc_frame* f0, const orc_frame* f1, int detector_type, double threshold_primary,
                                                double threshold_secondary, int triangulate_n_features, double mean_depth_inv,
                                                double min_depth_inv, double max_depth_inv, unsigned seed, int cap, int* n0_out,
                                                double* px0, int* level0, int* type0, double* score0, double* grad0, double* px1,
                                                double* fv1, double* grad1, int* level1, int* type1, double* score1, double* xyz1,
                                                int* ref_index1) {
  svo::FramePtr frame0 = makeFrame(*f0), frame1 = makeFrame(*f1);
  frame0->id_ = 1; frame1->id_ = 2;
  svo::DetectorOptions o;
  o.detector_type = static_cast<svo::DetectorType>(detector_type);
  o.threshold_primary = threshold_primary; o.threshold_secondary = threshold_secondary;
  svo::StereoTriangulationOptions so;
  so.triangulate_n_features = (size_t)triangulate_n_features;
  so.mean_depth_inv = mean_depth_inv; so.min_depth_inv = min_depth_inv; so.max_depth_inv = max_depth_inv;
  svo::StereoTriangulation st(so, svo::feature_detection_utils::makeDetector(o, frame0->cam_));
  srand(seed);
  st.compute(frame0, frame1);
  const int n0 = (int)frame0->num_features_, n1 = (int)frame1->num_features_;
  *n0_out = n0;
  for (int i = 0; i < n0 && i < cap; ++i) {
    px0[2 * i] = frame0->px_vec_(0, i); px0[2 * i + 1] = frame0->px_vec_(1, i);
    level0[i] = frame0->level_vec_(i); type0[i] = (int)frame0->type_vec_[i]; score0[i] = frame0->score_vec_(i);
    grad0[2 * i] = frame0->grad_vec_(0, i); grad0[2 * i + 1] = frame0->grad_vec_(1, i);
  }
  for (int i = 0; i < n1 && i < cap; ++i) {
    px1[2 * i] = frame1->px_vec_(0, i); px1[2 * i + 1] = frame1->px_vec_(1, i);
    for (int k = 0; k < 3; ++k) fv1[3 * i + k] = frame1->f_vec_(k, i);
    grad1[2 * i] = frame1->grad_vec_(0, i); grad1[2 * i + 1] = frame1->grad_vec_(1, i);
    level1[i] = frame1->level_vec_(i); type1[i] = (int)frame1->type_vec_[i]; score1[i] = frame1->score_vec_(i);
    const svo::PointPtr& p = frame1->landmark_vec_[i];
    for (int k = 0; k < 3; ++k) xyz1[3 * i + k] = p ? p->pos_[k] : 0.0;
    ref_index1[i] = p && !p->obs_.empty() ? (int)p->obs_[0].keypoint_index_ : -1;
  }
  return n1;
}

// ---- f3 (tracker part): FeatureTracker::trackAndDetect (src/svo_tracker/src/feature_tracker.cpp:14-248) over a mono sequence: detector
// from makeDetector, pyramidal KLT of every active track from its first observation (alignPyr2D), track bookkeeping, re-detection when
// fewer than min_tracks tracks survive. Per frame k the frame's columns after the call are written to row k of the outputs (cap each).
#include <svo/tracker/feature_tracker.h>
#include <svo/tracker/feature_tracking_types.h>

extern "C" int ref_feature_tracker_sequence(int n_frames, const orc_frame* frames, int detector_type, double threshold_primary,
                                            double threshold_secondary, int min_tracks_to_detect, int reset_before_detection,
                                            int template_is_first, int cap, int* n_features, double* px, int* track_id, double* score,
                                            int* n_active, int* n_terminated, double* disparity) {
  svo::FeatureTrackerOptions to;
  to.min_tracks_to_detect_new_features = (size_t)min_tracks_to_detect;
  to.reset_before_detection = reset_before_detection != 0;
  to.klt_template_is_first_observation = template_is_first != 0;
  svo::DetectorOptions o;
  o.detector_type = static_cast<svo::DetectorType>(detector_type);
  o.threshold_primary = threshold_primary; o.threshold_secondary = threshold_secondary;
  auto cams = std::make_shared<vk::cameras::NCamera>(std::vector<std::shared_ptr<vk::cameras::CameraGeometryBase>>{makeCamera(frames[0])});
  svo::FeatureTracker tracker(to, o, cams);
  std::vector<svo::FrameBundlePtr> keep;  // tracks reference their bundles
  int first_id = -1;
  for (int k = 0; k < n_frames; ++k) {
    svo::FramePtr fr = makeFrame(frames[k]);
    fr->id_ = k + 1;
    auto bundle = std::make_shared<svo::FrameBundle>(std::vector<svo::FramePtr>{fr});
    keep.push_back(bundle);
    tracker.trackAndDetect(bundle);
    const int n = (int)fr->num_features_;
    n_features[k] = n;
    for (int i = 0; i < n && i < cap; ++i) {
      if (first_id < 0) first_id = fr->track_id_vec_(i);
      px[((size_t)k * cap + i) * 2] = fr->px_vec_(0, i); px[((size_t)k * cap + i) * 2 + 1] = fr->px_vec_(1, i);
      track_id[(size_t)k * cap + i] = fr->track_id_vec_(i) - first_id;  // ids come from a process-wide counter
      score[(size_t)k * cap + i] = fr->score_vec_(i);
    }
    n_active[k] = (int)tracker.getTotalActiveTracks();
    n_terminated[k] = (int)tracker.terminated_tracks_.at(0).size();
    std::vector<size_t> nt; std::vector<double> disp;
    tracker.getNumTrackedAndDisparityPerFrame(0.5, &nt, &disp);
    disparity[k] = disp.at(0);
  }
  return 0;
}
