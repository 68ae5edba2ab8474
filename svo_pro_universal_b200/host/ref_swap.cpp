// Link-time swap of the reference's direct front-end hot path.
//
// This translation unit is compiled AGAINST THE REFERENCE'S OWN HEADERS and defines the hot-path functions those headers declare
// by calling the C ABI of libsvo_cuda.so (include/svo_cuda.h). Linked in place of the reference's sparse_img_align.cpp and
// matcher.cpp — and in front of the same-named functions of depth_filter.cpp, feature_alignment.cpp and
// feature_detection_utils.cpp — the rest of the reference (FrameHandlerBase's `new SparseImgAlign(...)`,
// src/svo/src/frame_handler_base.cpp:125,135,145; Reprojector, DepthFilter, the detectors) runs on the GPU path unchanged:
//
//   svo::SparseImgAlign::SparseImgAlign / run        src/svo_img_align/include/svo/img_align/sparse_img_align.h:30-77
//   svo::Matcher::findMatchDirect                    src/svo_direct/include/svo/direct/matcher.h:84-89
//   svo::Matcher::findEpipolarMatchDirect (x2)       matcher.h:92-108
//   svo::Matcher::scanEpipolarLine, getResultString  matcher.h:111-122
//   svo::depth_filter_utils::updateSeed              src/svo_direct/include/svo/direct/depth_filter.h:190-199
//   svo::depth_filter_utils::updateFilterVogiatzis / updateFilterGaussian / computeTau   depth_filter.h:201-216
//   svo::feature_alignment::align1D / align2D        src/svo_direct/include/svo/direct/feature_alignment.h:23-43
//   svo::feature_detection_utils::fastDetector       src/svo_direct/include/svo/direct/feature_detection_utils.h
//
// Frames carry their pyramid as host images (svo::Frame::img_pyr_); the device copy is kept in a small cache keyed by the
// content of level 0, so repeated calls on the same frames (one Matcher call per feature in the reference's loops) upload once.
// Every call here is a batch of ONE through the batched C ABI: this file proves the boundary, throughput comes from handing the
// reference's loops over as batches (host/svo_b200.cpp: Reprojector, DepthFilter::updateSeeds, FeatureTracker).
#include <svo/img_align/sparse_img_align.h>
#include <svo/direct/matcher.h>
#include <svo/direct/depth_filter.h>
#include <svo/direct/feature_alignment.h>
#include <svo/direct/feature_detection_utils.h>
#include <svo/direct/patch_utils.h>
#include <svo/direct/patch_score.h>
#include <svo/common/frame.h>
#include <svo/common/camera.h>
#include <svo/common/seed.h>
#include <svo/common/logging.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/svo_cuda.h"

namespace b200swap {

svo_cuda_ctx* context() {
  static svo_cuda_ctx* ctx = [] {
    svo_cuda_ctx* c = nullptr;
    const char* dev = std::getenv("SVO_B200_DEVICE");
    const int rc = svo_cuda_ctx_create(dev ? std::atoi(dev) : 0, &c);
    if (rc != SVO_OK) throw std::runtime_error("svo_cuda_ctx_create failed (status " + std::to_string(rc) + "): no CUDA device, and there is no CPU fallback");
    return c;
  }();
  return ctx;
}

void check(int rc, const char* what) {
  if (rc != SVO_OK) throw std::runtime_error(std::string(what) + ": " + svo_cuda_last_error(context()));
}

void toArray(const svo::Transformation& T, double* a) {
  const auto& q = T.getRotation().toImplementation();
  a[0] = q.w(); a[1] = q.x(); a[2] = q.y(); a[3] = q.z();
  a[4] = T.getPosition()[0]; a[5] = T.getPosition()[1]; a[6] = T.getPosition()[2];
}
svo::Transformation fromArray(const double* a) {
  return svo::Transformation(svo::Quaternion(a[0], a[1], a[2], a[3]), Eigen::Vector3d(a[4], a[5], a[6]));
}

svo_camera toCamera(const svo::Camera& cam) {
  svo_camera c;
  std::memset(&c, 0, sizeof(c));
  const Eigen::VectorXd k = cam.getIntrinsicParameters();
  const Eigen::VectorXd d = cam.getDistortionParameters();
  c.fx = k[0]; c.fy = k[1]; c.cx = k[2]; c.cy = k[3];
  c.width = (int)cam.imageWidth(); c.height = (int)cam.imageHeight();
  if (d.size() >= 4 && (d[0] != 0.0 || d[1] != 0.0 || d[2] != 0.0 || d[3] != 0.0)) {
    c.distortion = 1;
    c.k1 = d[0]; c.k2 = d[1]; c.p1 = d[2]; c.p2 = d[3];
  }
  return c;
}

// Device pyramid of a frame: level 0 is uploaded and the levels are rebuilt on the device (bit-identical to vk::halfSample,
// tests/test_gpu_detect.py). The cache key is a hash of level 0's bytes: a frame is an image, whatever object holds it.
class PyramidCache {
 public:
  const svo_cuda_pyr* get(const svo::Frame& f) {
    const cv::Mat& im = f.img_pyr_.at(0);
    const int n_levels = (int)f.img_pyr_.size();
    unsigned long long h = 1469598103934665603ull ^ ((unsigned long long)im.cols << 32) ^ (unsigned long long)im.rows ^ ((unsigned long long)n_levels << 56);
    for (int y = 0; y < im.rows; ++y) {
      const unsigned char* row = im.data + (size_t)y * im.step;
      int x = 0;
      for (; x + 8 <= im.cols; x += 8) { unsigned long long w; std::memcpy(&w, row + x, 8); h = (h ^ w) * 1099511628211ull; h ^= h >> 29; }
      for (; x < im.cols; ++x) h = (h ^ row[x]) * 1099511628211ull;
    }
    std::lock_guard<std::mutex> lock(mu_);
    for (auto it = entries_.begin(); it != entries_.end(); ++it)
      if (it->hash == h && it->cols == im.cols && it->rows == im.rows && it->n_levels == n_levels) {
        entries_.splice(entries_.begin(), entries_, it);
        return entries_.front().pyr;
      }
    svo_cuda_pyr* p = nullptr;
    check(svo_cuda_pyr_create(context(), 1, im.cols, im.rows, n_levels, -1, &p), "svo_cuda_pyr_create");
    check(svo_cuda_pyr_upload(context(), p, 0, 1, im.data, im.step, im.step * (size_t)im.rows, SVO_MEM_HOST), "svo_cuda_pyr_upload");
    check(svo_cuda_pyr_build(context(), p, 0, 1), "svo_cuda_pyr_build");
    check(svo_cuda_ctx_synchronize(context()), "svo_cuda_ctx_synchronize");  // the host image may go away after this call
    entries_.push_front(Entry{h, im.cols, im.rows, n_levels, p});
    if (entries_.size() > 24) {
      svo_cuda_pyr_destroy(context(), entries_.back().pyr);
      entries_.pop_back();
    }
    return p;
  }

 private:
  struct Entry { unsigned long long hash; int cols, rows, n_levels; svo_cuda_pyr* pyr; };
  std::list<Entry> entries_;
  std::mutex mu_;
};
PyramidCache& pyramids() { static PyramidCache c; return c; }

svo_feature toFeature(const svo::FeatureWrapper& f) {
  svo_feature o;
  std::memset(&o, 0, sizeof(o));
  o.px[0] = f.px[0]; o.px[1] = f.px[1];
  o.f[0] = f.f[0]; o.f[1] = f.f[1]; o.f[2] = f.f[2];
  o.grad[0] = f.grad[0]; o.grad[1] = f.grad[1];
  o.type = (int)f.type;
  o.level = (int)f.level;
  return o;
}

svo_matcher_options toOptions(const svo::Matcher::Options& m) {
  svo_matcher_options o;
  std::memset(&o, 0, sizeof(o));
  o.align_1d = m.align_1d; o.align_max_iter = m.align_max_iter;
  o.max_epi_search_steps = (int)m.max_epi_search_steps;
  o.subpix_refinement = m.subpix_refinement; o.epi_search_edgelet_filtering = m.epi_search_edgelet_filtering;
  o.scan_on_unit_sphere = m.scan_on_unit_sphere;
  o.epi_search_edgelet_max_angle = m.epi_search_edgelet_max_angle;
  o.affine_est_offset = m.affine_est_offset_; o.affine_est_gain = m.affine_est_gain_;
  o.max_patch_diff_ratio = m.max_patch_diff_ratio;
  return o;
}

// the warped reference patch (Matcher::patch_with_border_ / patch_) of one feature, as the matcher kernels form it
bool warpedPatch(const svo::Frame& ref_frame, const svo::Frame& cur_frame, const double T_cur_ref[7], const svo_feature& ft, double depth,
                 unsigned char* pwb, unsigned char* patch) {
  const svo_camera cr = toCamera(*ref_frame.cam()), cc = toCamera(*cur_frame.cam());
  double A[4];
  int sl = 0;
  unsigned char ok = 0;
  check(svo_cuda_warp_affine(context(), pyramids().get(ref_frame), nullptr, &cr, &cc, T_cur_ref, nullptr, 1, &ft, &depth, A, &sl, pwb, &ok, SVO_MEM_HOST),
        "svo_cuda_warp_affine");
  for (int y = 0; y < 8; ++y) std::memcpy(patch + 8 * y, pwb + 10 * (y + 1) + 1, 8);  // patch_utils::createPatchFromPatchWithBorder
  return ok != 0;
}

}  // namespace b200swap

namespace svo {

// ---- SparseImgAlign ------------------------------------------------------------------------------------------------------
SparseImgAlign::SparseImgAlign(SolverOptions optimization_options, SparseImgAlignOptions options)
    : SparseImgAlignBase(optimization_options, options) {
  setPatchSize<SparseImgAlign>(4);  // the device kernel is built for the reference's 4x4 patches (sparse_img_align.cpp:31)
}

size_t SparseImgAlign::run(const FrameBundle::Ptr& ref_frames, const FrameBundle::Ptr& cur_frames) {
  CHECK(!ref_frames->empty());
  CHECK_EQ(ref_frames->size(), cur_frames->size());
  CHECK_EQ(patch_size_, 4) << "libsvo_cuda's sparse alignment kernel implements the reference's 4x4 patches";
  const int n_cams = (int)ref_frames->size();
  CHECK_LE(n_cams, SVO_MAX_CAMS);
  size_t max_f = 1;
  for (const auto& f : ref_frames->frames_) max_f = std::max(max_f, f->num_features_);
  std::vector<const svo_cuda_pyr*> rp(n_cams), cp(n_cams);
  std::vector<svo_camera> cams(n_cams);
  std::vector<double> T_cam_imu(7 * n_cams), px(2 * max_f * n_cams, 0.0), fv(3 * max_f * n_cams, 0.0), depth(max_f * n_cams, 1.0);
  std::vector<uint8_t> eligible(max_f * n_cams, 0);
  std::vector<int> n_features(n_cams), idx(n_cams, 0);
  for (int c = 0; c < n_cams; ++c) {
    const Frame& rf = *ref_frames->at(c);
    rp[c] = b200swap::pyramids().get(rf);
    cp[c] = b200swap::pyramids().get(*cur_frames->at(c));
    cams[c] = b200swap::toCamera(*rf.cam());
    b200swap::toArray(rf.T_cam_imu(), &T_cam_imu[7 * c]);
    n_features[c] = (int)rf.num_features_;
    const Eigen::Vector3d ref_pos = rf.pos();
    for (size_t i = 0; i < rf.num_features_; ++i) {
      const size_t k = c * max_f + i;
      px[2 * k] = rf.px_vec_(0, i); px[2 * k + 1] = rf.px_vec_(1, i);
      fv[3 * k] = rf.f_vec_(0, i); fv[3 * k + 1] = rf.f_vec_(1, i); fv[3 * k + 2] = rf.f_vec_(2, i);
      // sparse_img_align.cpp:242-248: needs a landmark or a seed reference, and must not be a MapPoint type
      if ((rf.landmark_vec_[i] == nullptr && rf.seed_ref_vec_[i].keyframe == nullptr) || isMapPoint(rf.type_vec_[i])) continue;
      eligible[k] = 1;
      if (rf.landmark_vec_[i]) {  // :281-291: the depth is read back as |landmark - camera centre|
        depth[k] = (rf.landmark_vec_[i]->pos_ - ref_pos).norm();
      } else {
        const SeedRef& sr = rf.seed_ref_vec_[i];
        const Position pos = sr.keyframe->T_world_cam() * sr.keyframe->getSeedPosInFrame(sr.seed_id);
        depth[k] = (pos - ref_pos).norm();
      }
    }
  }
  double T_ref[7], T_cur[7];
  b200swap::toArray(ref_frames->at(0)->T_imu_world(), T_ref);
  b200swap::toArray(cur_frames->at(0)->T_imu_world(), T_cur);
  svo_sparse_align_options o;
  std::memset(&o, 0, sizeof(o));
  o.max_level = options_.max_level; o.min_level = options_.min_level;
  o.estimate_illumination_gain = options_.estimate_illumination_gain;
  o.estimate_illumination_offset = options_.estimate_illumination_offset;
  o.use_distortion_jacobian = options_.use_distortion_jacobian;
  o.robustification = options_.robustification;
  o.weight_scale = options_.weight_scale;
  o.max_iter = (int)solver_options_.max_iter;
  o.eps = solver_options_.eps;
  o.alpha_init = alpha_init_; o.beta_init = beta_init_;
  o.lambda_rot = prior_lambda_rot_; o.lambda_trans = prior_lambda_trans_; o.lambda_alpha = prior_lambda_alpha_; o.lambda_beta = prior_lambda_beta_;
  svo_align_prior prior;
  std::memset(&prior, 0, sizeof(prior));
  if (have_prior_) {
    b200swap::toArray(prior_.T_icur_iref, prior.T);
    prior.alpha = prior_.alpha; prior.beta = prior_.beta;
  }
  svo_align_result res;
  std::memset(&res, 0, sizeof(res));
  b200swap::check(svo_cuda_sparse_align(b200swap::context(), n_cams, rp.data(), cp.data(), idx.data(), idx.data(), cams.data(), T_cam_imu.data(), 1,
                                        T_ref, T_cur, n_features.data(), (int)max_f, px.data(), fv.data(), depth.data(), eligible.data(), &o,
                                        have_prior_ ? &prior : nullptr, &res, SVO_MEM_HOST),
                  "svo_cuda_sparse_align");
  if (res.n_tracked == 0) {
    SVO_ERROR_STREAM("SparseImgAlign: no features to track!");
    return 0;  // the frames' poses stay untouched (:53-57)
  }
  ref_frames_ = ref_frames;
  cur_frames_ = cur_frames;
  T_iref_world_ = ref_frames->at(0)->T_imu_world();
  for (int c = 0; c < n_cams; ++c) cur_frames->at(c)->T_f_w_ = b200swap::fromArray(res.T_f_w[c]);  // :103-106
  chi2_ = res.chi2;
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) H_(i, j) = res.H[8 * i + j];
  stop_ = res.stop != 0;
  level_ = options_.min_level;
  const int n_lv = options_.max_level - options_.min_level;
  iter_ = (n_lv >= 0 && n_lv < SVO_MAX_LEVELS) ? (size_t)std::max(0, res.iters[n_lv] - 1) : 0;
  alpha_init_ = 0.0;  // :108-110
  beta_init_ = 0.0;
  return (size_t)res.n_tracked;
}

// The solver's per-iteration hooks are not used: the whole Gauss-Newton run happens in one kernel launch.
double SparseImgAlign::evaluateError(const SparseImgAlignState&, HessianMatrix*, GradientVector*) {
  LOG(FATAL) << "SparseImgAlign::evaluateError: the GPU path has no per-iteration host callback";
  return 0.0;
}
void SparseImgAlign::update(const SparseImgAlignState& old_model, const UpdateVector& dx, SparseImgAlignState& new_model) {
  SparseImgAlignBase::update(old_model, dx, new_model);
}
void SparseImgAlign::applyPrior(const SparseImgAlignState& current_model) { SparseImgAlignBase::applyPrior(current_model); }
void SparseImgAlign::finishIteration() {}

// ---- Matcher -------------------------------------------------------------------------------------------------------------
namespace {
void fillMembers(Matcher& m, const svo_match_out& o, bool px_valid) {
  m.A_cur_ref_(0, 0) = o.A_cur_ref[0]; m.A_cur_ref_(0, 1) = o.A_cur_ref[1];
  m.A_cur_ref_(1, 0) = o.A_cur_ref[2]; m.A_cur_ref_(1, 1) = o.A_cur_ref[3];
  m.search_level_ = o.search_level;
  if (px_valid) {
    m.px_cur_ = Keypoint(o.px_cur[0], o.px_cur[1]);
    m.f_cur_ = BearingVector(o.f_cur[0], o.f_cur[1], o.f_cur[2]);
  }
}
}  // namespace

Matcher::MatchResult Matcher::findMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr, const FloatType& ref_depth,
                                               Keypoint& px_cur) {
  CHECK(options_.use_affine_warp_) << "the GPU matcher implements the affine warp (the reference's default)";
  const svo_camera cr = b200swap::toCamera(*ref_frame.cam()), cc = b200swap::toCamera(*cur_frame.cam());
  double T[7];
  b200swap::toArray(cur_frame.T_cam_world() * ref_frame.T_world_cam(), T);
  const svo_feature ft = b200swap::toFeature(ref_ftr);
  const svo_matcher_options mo = b200swap::toOptions(options_);
  const double depth = ref_depth, guess[2] = {px_cur[0], px_cur[1]};
  svo_match_out out;
  std::memset(&out, 0, sizeof(out));
  b200swap::check(svo_cuda_find_match_direct(b200swap::context(), b200swap::pyramids().get(ref_frame), b200swap::pyramids().get(cur_frame), nullptr, nullptr,
                                             &cr, &cc, T, nullptr, 1, &ft, &depth, guess, &mo, &out, SVO_MEM_HOST),
                  "svo_cuda_find_match_direct");
  const MatchResult r = static_cast<MatchResult>(out.result);
  if (r == MatchResult::kFailVisibility) return r;  // nothing was computed (matcher.cpp:38-44)
  fillMembers(*this, out, r == MatchResult::kSuccess);
  if (r != MatchResult::kFailWarp) b200swap::warpedPatch(ref_frame, cur_frame, T, ft, depth, patch_with_border_, patch_);
  if (isEdgelet(ref_ftr.type)) h_inv_ = out.h_inv;
  if (r == MatchResult::kSuccess) px_cur = px_cur_;
  return r;
}

Matcher::MatchResult Matcher::findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr,
                                                       const double d_estimate_inv, const double d_min_inv, const double d_max_inv, double& depth) {
  const Transformation T_cur_ref = cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse();  // matcher.cpp:148-155
  return findEpipolarMatchDirect(ref_frame, cur_frame, T_cur_ref, ref_ftr, d_estimate_inv, d_min_inv, d_max_inv, depth);
}

Matcher::MatchResult Matcher::findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const Transformation& T_cur_ref,
                                                       const FeatureWrapper& ref_ftr, const double d_estimate_inv, const double d_min_inv,
                                                       const double d_max_inv, double& depth) {
  const svo_camera cr = b200swap::toCamera(*ref_frame.cam()), cc = b200swap::toCamera(*cur_frame.cam());
  double T[7];
  b200swap::toArray(T_cur_ref, T);
  const svo_feature ft = b200swap::toFeature(ref_ftr);
  const svo_matcher_options mo = b200swap::toOptions(options_);
  const double d3[3] = {d_estimate_inv, d_min_inv, d_max_inv};
  svo_match_out out;
  std::memset(&out, 0, sizeof(out));
  b200swap::check(svo_cuda_find_epipolar_match_direct(b200swap::context(), b200swap::pyramids().get(ref_frame), b200swap::pyramids().get(cur_frame), nullptr,
                                                      nullptr, &cr, &cc, T, nullptr, 1, &ft, d3, &mo, &out, SVO_MEM_HOST),
                  "svo_cuda_find_epipolar_match_direct");
  const MatchResult r = static_cast<MatchResult>(out.result);
  reject_ = out.reject != 0;
  epi_image_ = Eigen::Vector2d(out.epi_image[0], out.epi_image[1]);  // matcher.cpp:176: set before every return
  A_cur_ref_(0, 0) = out.A_cur_ref[0]; A_cur_ref_(0, 1) = out.A_cur_ref[1]; A_cur_ref_(1, 0) = out.A_cur_ref[2]; A_cur_ref_(1, 1) = out.A_cur_ref[3];
  if (r == MatchResult::kFailAngle) return r;  // matcher.cpp:181-192: returns before the search level is chosen
  search_level_ = out.search_level;
  epi_length_pyramid_ = out.epi_length_pyramid;
  if (r != MatchResult::kFailWarp)
    b200swap::warpedPatch(ref_frame, cur_frame, T, ft, 1.0 / std::max(0.000001, d_estimate_inv), patch_with_border_, patch_);
  if (r == MatchResult::kFailWarp) return r;
  px_cur_ = Keypoint(out.px_cur[0], out.px_cur[1]);  // the scan / the mid-point sets px_cur_ before the later checks (:209-229)
  h_inv_ = out.h_inv;
  if (r == MatchResult::kSuccess || r == MatchResult::kFailTriangulation) f_cur_ = BearingVector(out.f_cur[0], out.f_cur[1], out.f_cur[2]);
  if (r == MatchResult::kSuccess) depth = out.depth;
  return r;
}

void Matcher::scanEpipolarLine(const Frame& frame, const Eigen::Vector3d& A, const Eigen::Vector3d& B, const Eigen::Vector3d& C,
                               const PatchScore& patch_score, const int patch_level, Keypoint* image_best, int* zmssd_best) {
  const svo_camera cc = b200swap::toCamera(*frame.cam());
  const svo_matcher_options mo = b200swap::toOptions(options_);
  double best[2] = {0.0, 0.0};
  b200swap::check(svo_cuda_scan_epipolar_line(b200swap::context(), b200swap::pyramids().get(frame), nullptr, &cc, 1, A.data(), B.data(), C.data(),
                                              patch_score.ref_patch_, &patch_level, &epi_length_pyramid_, &mo, best, zmssd_best, SVO_MEM_HOST),
                  "svo_cuda_scan_epipolar_line");
  *image_best = Keypoint(best[0], best[1]);
}

std::string Matcher::getResultString(const Matcher::MatchResult& result) {  // matcher.cpp:243-260
  switch (result) {
    case MatchResult::kSuccess: return "Success";
    case MatchResult::kFailScore: return "FailScore";
    case MatchResult::kFailTriangulation: return "FailTriangulation";
    case MatchResult::kFailVisibility: return "FailVisibility";
    case MatchResult::kFailWarp: return "FailWarp";
    case MatchResult::kFailAlignment: return "FailAlignment";
    case MatchResult::kFailRange: return "FailRange";
    case MatchResult::kFailAngle: return "FailAngle";
    case MatchResult::kFailCloseView: return "FailCloseView";
    case MatchResult::kFailLock: return "FailLock";
    case MatchResult::kFailTooFar: return "FailTooFar";
  }
  return "unknown";
}

// ---- depth filter --------------------------------------------------------------------------------------------------------
namespace depth_filter_utils {

bool updateSeed(const Frame& cur_frame, Frame& ref_frame, const size_t& seed_index, Matcher& matcher, const FloatType sigma2_convergence_threshold,
                const bool check_visibility, const bool check_convergence, const bool use_vogiatzis_update) {
  if (cur_frame.id() == ref_frame.id()) {
    SVO_WARN_STREAM_THROTTLE(1.0, "update seed with ref frame");
    return false;
  }
  static double px_error_angle = cur_frame.getAngleError(1.0);  // the reference's function-static (depth_filter.cpp:383-384)
  const svo_camera cr = b200swap::toCamera(*ref_frame.cam()), cc = b200swap::toCamera(*cur_frame.cam());
  FeatureWrapper ref_ftr = ref_frame.getFeatureWrapper(seed_index);
  const svo_feature ft = b200swap::toFeature(ref_ftr);
  uint8_t type = (uint8_t)ref_frame.type_vec_[seed_index];
  double state[4];
  for (int k = 0; k < 4; ++k) state[k] = ref_frame.invmu_sigma2_a_b_vec_(k, seed_index);
  const double mu_range = ref_frame.seed_mu_range_;
  double T[7];
  b200swap::toArray(cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse(), T);
  svo_matcher_options mo = b200swap::toOptions(matcher.options_);
  svo_depth_filter_options dopt;
  std::memset(&dopt, 0, sizeof(dopt));
  // both thresholds carry the caller's value: DepthFilter::updateSeeds has already picked it by seed type (depth_filter.cpp:214-221)
  dopt.seed_convergence_sigma2_thresh = sigma2_convergence_threshold;
  dopt.mappoint_convergence_sigma2_thresh = sigma2_convergence_threshold;
  dopt.px_error_angle = px_error_angle;
  dopt.check_visibility = check_visibility; dopt.check_convergence = check_convergence; dopt.use_vogiatzis_update = use_vogiatzis_update;
  const int zero = 0;
  int n_success = 0, match_result = -1;
  b200swap::check(svo_cuda_update_seeds(b200swap::context(), b200swap::pyramids().get(ref_frame), b200swap::pyramids().get(cur_frame), &cr, &cc, 1, &zero, &ft,
                                        &type, state, &mu_range, 1, &zero, &zero, T, &mo, &dopt, &n_success, &match_result, SVO_MEM_HOST),
                  "svo_cuda_update_seeds");
  // What the reference leaves behind: the seed's state and type, the matcher's align_1d option (:423-427) and — read by
  // Reprojector::matchCandidate right after updateSeed (reprojector.cpp:412-440) — the Matcher's public result members. The
  // batched seed update keeps the match on the device, so the one match this call made is repeated through the Matcher entry
  // point to fill them (a batch of one either way; the batched callers never need the members).
  if (match_result >= 0) {
    matcher.options_.align_1d = (ft.type == (int)FeatureType::kEdgeletSeed || ft.type == (int)FeatureType::kEdgeletSeedConverged);
    Eigen::Ref<SeedState> before = ref_frame.invmu_sigma2_a_b_vec_.col(seed_index);  // still the state the match was made with
    double depth = 0.0;
    matcher.findEpipolarMatchDirect(ref_frame, cur_frame, cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse(), ref_ftr, seed::getInvDepth(before),
                                    seed::getInvMinDepth(before), seed::getInvMaxDepth(before), depth);
  }
  for (int k = 0; k < 4; ++k) ref_frame.invmu_sigma2_a_b_vec_(k, seed_index) = state[k];
  ref_frame.type_vec_[seed_index] = static_cast<FeatureType>(type);
  return n_success != 0;
}

bool updateFilterVogiatzis(const FloatType z, const FloatType tau2, const FloatType z_range, Eigen::Ref<SeedState>& seed) {
  double s[4] = {seed(0), seed(1), seed(2), seed(3)};
  const double zz = z, tt = tau2, rr = z_range;
  uint8_t ok = 0;
  b200swap::check(svo_cuda_update_filter_vogiatzis(b200swap::context(), 1, &zz, &tt, &rr, s, &ok, SVO_MEM_HOST), "svo_cuda_update_filter_vogiatzis");
  for (int k = 0; k < 4; ++k) seed(k) = s[k];
  return ok != 0;
}

bool updateFilterGaussian(const FloatType z, const FloatType tau2, Eigen::Ref<SeedState>& seed) {
  double s[4] = {seed(0), seed(1), seed(2), seed(3)};
  const double zz = z, tt = tau2;
  uint8_t ok = 0;
  b200swap::check(svo_cuda_update_filter_seq(b200swap::context(), 1, 1, &zz, &tt, nullptr, s, &ok, 1, SVO_MEM_HOST), "svo_cuda_update_filter_seq");
  for (int k = 0; k < 4; ++k) seed(k) = s[k];
  return ok != 0;
}

double computeTau(const Transformation& T_ref_cur, const BearingVector& f, const FloatType z, const FloatType px_error_angle) {
  double T[7], fv[3] = {f[0], f[1], f[2]}, zz = z, tau = 0.0;
  b200swap::toArray(T_ref_cur, T);
  b200swap::check(svo_cuda_compute_tau(b200swap::context(), 1, T, fv, &zz, px_error_angle, &tau, SVO_MEM_HOST), "svo_cuda_compute_tau");
  return tau;
}

}  // namespace depth_filter_utils
}  // namespace svo
