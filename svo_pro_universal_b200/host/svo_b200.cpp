// Host facades over the C ABI (see svo_b200.h). Every method packs its arguments into the POD batches of
// include/svo_cuda.h, calls ONE entry point with host buffers (SVO_MEM_HOST) and unpacks the results; nothing here
// computes the hot path on the CPU.
#include "svo_b200.h"
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <numeric>

namespace svo {

// ---- Transformation (kindr::minimal::QuatTransformation semantics; host-side glue only) --------------------------------
static std::array<double, 3> rotate(const std::array<double, 4>& q, const std::array<double, 3>& v) {
  // Eigen QuaternionBase::_transformVector
  const double qx = q[1], qy = q[2], qz = q[3], w = q[0];
  double ux = qy * v[2] - qz * v[1], uy = qz * v[0] - qx * v[2], uz = qx * v[1] - qy * v[0];
  ux += ux; uy += uy; uz += uz;
  return {v[0] + w * ux + (qy * uz - qz * uy), v[1] + w * uy + (qz * ux - qx * uz), v[2] + w * uz + (qx * uy - qy * ux)};
}
Transformation Transformation::inverse() const {  // quat-transformation-inl.h:209-213
  Transformation r;
  r.q = {q[0], -q[1], -q[2], -q[3]};
  const auto v = rotate(r.q, t);
  r.t = {-v[0], -v[1], -v[2]};
  return r;
}
Transformation Transformation::operator*(const Transformation& b) const {  // quat-transformation-inl.h:150-156
  Transformation r;
  const auto& a = q;
  r.q = {a[0] * b.q[0] - a[1] * b.q[1] - a[2] * b.q[2] - a[3] * b.q[3], a[0] * b.q[1] + a[1] * b.q[0] + a[2] * b.q[3] - a[3] * b.q[2],
         a[0] * b.q[2] + a[2] * b.q[0] + a[3] * b.q[1] - a[1] * b.q[3], a[0] * b.q[3] + a[3] * b.q[0] + a[1] * b.q[2] - a[2] * b.q[1]};
  const double n2 = r.q[0] * r.q[0] + r.q[1] * r.q[1] + r.q[2] * r.q[2] + r.q[3] * r.q[3];
  if (std::abs(n2 - 1.0) > 1e-4) { const double n = std::sqrt(n2); for (double& c : r.q) c /= n; }
  const auto v = rotate(q, b.t);
  r.t = {t[0] + v[0], t[1] + v[1], t[2] + v[2]};
  return r;
}

KeypointIdentifier::KeypointIdentifier(const FramePtr& _frame, const size_t _feature_index)  // point.cpp:19-23
    : frame(_frame), frame_id(_frame->id_), keypoint_index_(_feature_index) {}

size_t Frame::numTrackedFeatures() const {  // frame.h:153-163
  size_t count = 0;
  for (size_t i = 0; i < num_features_; ++i)
    if ((isValidLandmark(i) && !isFixedLandmark(type_vec_[i]) && !isMapPoint(type_vec_[i])) || isCornerEdgeletSeed(type_vec_[i])) ++count;
  return count;
}

void Frame::clearFeatureStorage() {
  px_vec_.clear(); f_vec_.clear(); grad_vec_.clear(); score_vec_.clear(); level_vec_.clear(); type_vec_.clear();
  depth_vec_.clear(); invmu_sigma2_a_b_vec_.clear(); landmark_vec_.clear(); seed_ref_vec_.clear(); track_id_vec_.clear();
  num_features_ = 0;
}

// ---- device plumbing -----------------------------------------------------------------------------------------------------
namespace b200 {
static void check(int rc, const char* what) {
  if (rc != SVO_OK) throw Error(std::string(what) + " failed with status " + std::to_string(rc) + ": " + svo_cuda_last_error(context()));
}
// One context (= one stream) per host thread, on the device named by SVO_B200_DEVICE (default 0); it is destroyed when its thread
// exits. Objects that outlive a thread (device pyramids cached on frames, the Reprojector's map copy) do not keep a context: they are
// released through the releasing thread's context, and cudaFree synchronises the whole device.
namespace {
struct ThreadContext {
  svo_cuda_ctx* ctx = nullptr;
  ~ThreadContext() { if (ctx) svo_cuda_ctx_destroy(ctx); }
};
int facadeDevice() {
  static const int dev = [] { const char* e = std::getenv("SVO_B200_DEVICE"); return e ? std::atoi(e) : 0; }();
  return dev;
}
std::mutex& frameUploadMutex() { static std::mutex m; return m; }
}  // namespace
svo_cuda_ctx* context() {
  static thread_local ThreadContext tc;
  if (!tc.ctx) {
    const int rc = svo_cuda_ctx_create(facadeDevice(), &tc.ctx);
    if (rc != SVO_OK) throw Error("svo_cuda_ctx_create failed (no CUDA device? this front-end has no CPU fallback), status " + std::to_string(rc));
  }
  return tc.ctx;
}
GpuPyramid::GpuPyramid(int width, int height, int n_levels) : width_(width), height_(height), n_levels_(n_levels) {
  check(svo_cuda_pyr_create(context(), 1, width, height, n_levels, -1, &pyr_), "svo_cuda_pyr_create");
}
GpuPyramid::~GpuPyramid() {
  if (pyr_) svo_cuda_pyr_destroy(context(), pyr_);
}
// Frames are shared between threads (the reference's tracking thread and its DepthFilter thread both hold FramePtrs), each with its
// own stream: the lazy upload is serialised, and the copy is published only after the uploading stream has finished building it, so
// whichever stream reads frame.gpu_ afterwards sees complete levels.
const GpuPyramid& ensureGpu(const Frame& frame) {
  std::lock_guard<std::mutex> lock(frameUploadMutex());
  if (!frame.gpu_) {
    if (frame.img_pyr_.empty() || frame.img_pyr_[0].empty()) throw Error("frame has no image");
    const Image& l0 = frame.img_pyr_[0];
    auto g = std::make_shared<GpuPyramid>(l0.cols, l0.rows, int(frame.img_pyr_.size()));
    check(svo_cuda_pyr_upload(context(), g->handle(), 0, 1, l0.data, l0.step, l0.step * l0.rows, SVO_MEM_HOST), "svo_cuda_pyr_upload");
    check(svo_cuda_pyr_build(context(), g->handle(), 0, 1), "svo_cuda_pyr_build");
    check(svo_cuda_ctx_synchronize(context()), "svo_cuda_ctx_synchronize");
    frame.gpu_ = g;
  }
  return *frame.gpu_;
}
}  // namespace b200

namespace frame_utils {
void createImgPyramid(const Image& img_level_0, int n_levels, ImgPyr& pyr, std::shared_ptr<b200::GpuPyramid>* gpu) {
  if (img_level_0.empty() || img_level_0.rows <= 0 || img_level_0.cols <= 0 || n_levels <= 0)
    throw b200::Error("createImgPyramid: invalid image or level count");  // CHECKs of frame.cpp:374-377
  auto g = std::make_shared<b200::GpuPyramid>(img_level_0.cols, img_level_0.rows, n_levels);
  svo_cuda_ctx* ctx = b200::context();
  b200::check(svo_cuda_pyr_upload(ctx, g->handle(), 0, 1, img_level_0.data, img_level_0.step, img_level_0.step * img_level_0.rows, SVO_MEM_HOST),
              "svo_cuda_pyr_upload");
  b200::check(svo_cuda_pyr_build(ctx, g->handle(), 0, 1), "svo_cuda_pyr_build");
  pyr.resize(n_levels);
  pyr[0] = img_level_0;
  if (!pyr[0].storage.empty()) pyr[0].data = pyr[0].storage.data();
  for (int l = 1; l < n_levels; ++l) {
    int cols, rows;
    svo_cuda_pyr_level_info(g->handle(), l, &cols, &rows, nullptr, nullptr, nullptr);
    pyr[l] = Image(rows, cols);
    b200::check(svo_cuda_pyr_download(ctx, g->handle(), 0, l, pyr[l].data, pyr[l].step, SVO_MEM_HOST), "svo_cuda_pyr_download");
  }
  if (gpu) *gpu = g;
}
}  // namespace frame_utils

// ---- SparseImgAlign ---------------------------------------------------------------------------------------------------------
SparseImgAlignBase::SolverOptions SparseImgAlignBase::getDefaultSolverOptions() {
  SolverOptions options;
  options.max_iter = 10;
  options.eps = 0.0005;
  return options;
}
void SparseImgAlignBase::setWeightedPrior(const Transformation& T_cur_ref_prior, double alpha_prior, double beta_prior, double lambda_rot,
                                          double lambda_trans, double lambda_alpha, double lambda_beta) {
  prior_lambda_rot_ = lambda_rot; prior_lambda_trans_ = lambda_trans; prior_lambda_alpha_ = lambda_alpha; prior_lambda_beta_ = lambda_beta;
  T_cur_ref_prior.toArray(prior_.T);
  prior_.alpha = alpha_prior; prior_.beta = beta_prior;
  have_prior_ = true;
}
void SparseImgAlignBase::reset() {
  have_prior_ = false;
  chi2_ = 1e10;
}

size_t SparseImgAlign::run(const FrameBundle::Ptr& ref_frames, const FrameBundle::Ptr& cur_frames) {
  if (!ref_frames || !cur_frames || ref_frames->empty() || ref_frames->size() != cur_frames->size() || ref_frames->size() > SVO_MAX_CAMS)
    throw b200::Error("SparseImgAlign::run: invalid frame bundles");  // CHECKs of sparse_img_align.cpp:38-39
  const int n_cams = int(ref_frames->size());
  size_t max_f = 1;
  for (const auto& f : ref_frames->frames_) max_f = std::max(max_f, f->num_features_);
  std::vector<const svo_cuda_pyr*> rp(n_cams), cp(n_cams);
  std::vector<svo_camera> cams(n_cams);
  std::vector<double> T_cam_imu(7 * n_cams), px(2 * max_f * n_cams), fv(3 * max_f * n_cams), depth(max_f * n_cams, 1.0);
  std::vector<uint8_t> eligible(max_f * n_cams, 0);
  std::vector<int> n_features(n_cams);
  for (int c = 0; c < n_cams; ++c) {
    const Frame& rf = *ref_frames->at(c);
    const Frame& cf = *cur_frames->at(c);
    rp[c] = b200::ensureGpu(rf).handle();
    cp[c] = b200::ensureGpu(cf).handle();
    cams[c] = rf.cam_->model;
    rf.T_cam_imu_.toArray(&T_cam_imu[7 * c]);
    n_features[c] = int(rf.num_features_);
    for (size_t i = 0; i < rf.num_features_; ++i) {
      const size_t k = c * max_f + i;
      px[2 * k] = rf.px_vec_[i][0]; px[2 * k + 1] = rf.px_vec_[i][1];
      fv[3 * k] = rf.f_vec_[i][0]; fv[3 * k + 1] = rf.f_vec_[i][1]; fv[3 * k + 2] = rf.f_vec_[i][2];
      const bool has_depth = i < rf.depth_vec_.size() && rf.depth_vec_[i] > 0.0;
      depth[k] = has_depth ? rf.depth_vec_[i] : 1.0;
      // sparse_img_align.cpp:242-248: needs a landmark or seed reference, and must not be a MapPoint type
      eligible[k] = has_depth && !isMapPoint(rf.type_vec_[i]);
    }
  }
  double T_ref[7], T_cur[7];
  ref_frames->at(0)->T_imu_world().toArray(T_ref);
  cur_frames->at(0)->T_imu_world().toArray(T_cur);
  svo_sparse_align_options o{};
  o.max_level = options_.max_level; o.min_level = options_.min_level;
  o.estimate_illumination_gain = options_.estimate_illumination_gain;
  o.estimate_illumination_offset = options_.estimate_illumination_offset;
  o.use_distortion_jacobian = options_.use_distortion_jacobian;
  o.robustification = options_.robustification;
  o.weight_scale = options_.weight_scale;
  o.max_iter = int(solver_options_.max_iter);
  o.eps = solver_options_.eps;
  o.alpha_init = alpha_init_; o.beta_init = beta_init_;
  o.lambda_rot = prior_lambda_rot_; o.lambda_trans = prior_lambda_trans_; o.lambda_alpha = prior_lambda_alpha_; o.lambda_beta = prior_lambda_beta_;
  svo_align_result res{};
  const int zero = 0;
  std::vector<int> idx(n_cams, zero);
  b200::check(svo_cuda_sparse_align(b200::context(), n_cams, rp.data(), cp.data(), idx.data(), idx.data(), cams.data(), T_cam_imu.data(), 1,
                                    T_ref, T_cur, n_features.data(), int(max_f), px.data(), fv.data(), depth.data(), eligible.data(), &o,
                                    have_prior_ ? &prior_ : nullptr, &res, SVO_MEM_HOST),
              "svo_cuda_sparse_align");
  if (res.n_tracked == 0) return 0;  // "SparseImgAlign: no features to track!" — poses are left untouched (:53-57)
  for (int c = 0; c < n_cams; ++c) cur_frames->at(c)->T_f_w_ = Transformation::fromArray(res.T_f_w[c]);  // :103-106
  chi2_ = res.chi2;
  std::copy(res.H, res.H + 64, H_.begin());
  alpha_init_ = 0.0;  // :108-110
  beta_init_ = 0.0;
  return size_t(res.n_tracked);
}

// ---- feature alignment ------------------------------------------------------------------------------------------------------
namespace {
struct OneLevel {  // a single host image as a 1-frame, 1-level device pyramid
  b200::GpuPyramid g;
  explicit OneLevel(const Image& img) : g(img.cols, img.rows, 1) {
    b200::check(svo_cuda_pyr_upload(b200::context(), g.handle(), 0, 1, img.data, img.step, img.step * img.rows, SVO_MEM_HOST), "svo_cuda_pyr_upload");
  }
};
}  // namespace
namespace feature_alignment {
bool align1D(const Image& cur_img, const GradientVector& dir, uint8_t* ref_patch_with_border, uint8_t* /*ref_patch*/, const int n_iter,
             const bool affine_est_offset, const bool affine_est_gain, Keypoint* cur_px_estimate, double* h_inv) {
  if (!cur_px_estimate) throw b200::Error("align1D: cur_px_estimate is null");  // CHECK_NOTNULL (:42)
  OneLevel lv(cur_img);
  const int zero = 0;
  uint8_t conv = 0;
  double h = 0.0;
  b200::check(svo_cuda_align1d(b200::context(), lv.g.handle(), &zero, &zero, 1, dir.data(), ref_patch_with_border, n_iter, affine_est_offset,
                               affine_est_gain, cur_px_estimate->data(), &h, &conv, SVO_MEM_HOST), "svo_cuda_align1d");
  if (h_inv) *h_inv = h;
  return conv != 0;
}
bool align2D(const Image& cur_img, uint8_t* ref_patch_with_border, uint8_t* /*ref_patch*/, const int n_iter, const bool affine_est_offset,
             const bool affine_est_gain, Keypoint& cur_px_estimate, bool /*no_simd*/) {
  OneLevel lv(cur_img);
  const int zero = 0;
  uint8_t conv = 0;
  b200::check(svo_cuda_align2d(b200::context(), lv.g.handle(), &zero, &zero, 1, ref_patch_with_border, n_iter, affine_est_offset, affine_est_gain,
                               cur_px_estimate.data(), &conv, SVO_MEM_HOST), "svo_cuda_align2d");
  return conv != 0;
}
}  // namespace feature_alignment

// ---- Matcher -----------------------------------------------------------------------------------------------------------------
svo_matcher_options Matcher::cOptions() const {
  svo_matcher_options o{};
  o.align_1d = options_.align_1d; o.align_max_iter = options_.align_max_iter;
  o.max_epi_search_steps = int(options_.max_epi_search_steps);
  o.subpix_refinement = options_.subpix_refinement;
  o.epi_search_edgelet_filtering = options_.epi_search_edgelet_filtering;
  o.scan_on_unit_sphere = options_.scan_on_unit_sphere;
  o.epi_search_edgelet_max_angle = options_.epi_search_edgelet_max_angle;
  o.affine_est_offset = options_.affine_est_offset_; o.affine_est_gain = options_.affine_est_gain_;
  o.max_patch_diff_ratio = options_.max_patch_diff_ratio;
  return o;
}
static svo_feature cFeature(const FeatureWrapper& f) {
  svo_feature c{};
  c.px[0] = f.px[0]; c.px[1] = f.px[1];
  c.f[0] = f.f[0]; c.f[1] = f.f[1]; c.f[2] = f.f[2];
  c.grad[0] = f.grad[0]; c.grad[1] = f.grad[1];
  c.type = int(f.type);
  c.level = f.level;
  return c;
}
// Matcher::patch_with_border_ / patch_ (matcher.h:70-71): the warped reference patch of one feature, as the matcher kernels form it
// (patch_warp.cpp:97-156 + patch_utils::createPatchFromPatchWithBorder) — one more single-feature launch, only for the callers
// that read the members.
static void warpedPatch(Matcher& m, const Frame& ref_frame, const Frame& cur_frame, const double T_cur_ref[7], const svo_feature& ft, double depth) {
  double A[4];
  int sl = 0;
  uint8_t ok = 0;
  b200::check(svo_cuda_warp_affine(b200::context(), b200::ensureGpu(ref_frame).handle(), nullptr, &ref_frame.cam_->model, &cur_frame.cam_->model,
                                   T_cur_ref, nullptr, 1, &ft, &depth, A, &sl, m.patch_with_border_, &ok, SVO_MEM_HOST),
              "svo_cuda_warp_affine");
  for (int y = 0; y < 8; ++y) std::memcpy(m.patch_ + 8 * y, m.patch_with_border_ + 10 * (y + 1) + 1, 8);
}
static void unpack(Matcher& m, const svo_match_out& o) {
  std::copy(o.A_cur_ref, o.A_cur_ref + 4, m.A_cur_ref_.begin());
  m.epi_length_pyramid_ = o.epi_length_pyramid;
  m.h_inv_ = o.h_inv;
  m.search_level_ = o.search_level;
  m.reject_ = o.reject != 0;
  m.px_cur_ = {o.px_cur[0], o.px_cur[1]};
  m.f_cur_ = {o.f_cur[0], o.f_cur[1], o.f_cur[2]};
}
Matcher::MatchResult Matcher::findMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr,
                                              const FloatType& ref_depth, Keypoint& px_cur) {
  const svo_feature ft = cFeature(ref_ftr);
  const svo_matcher_options o = cOptions();
  double T[7];
  (cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse()).toArray(T);  // cur.T_cam_world() * ref.T_world_cam() (matcher.cpp:49)
  svo_match_out out{};
  const int zero = 0;
  const double d = ref_depth;
  b200::check(svo_cuda_find_match_direct(b200::context(), b200::ensureGpu(ref_frame).handle(), b200::ensureGpu(cur_frame).handle(), &zero, &zero,
                                         &ref_frame.cam_->model, &cur_frame.cam_->model, T, &zero, 1, &ft, &d, px_cur.data(), &o, &out, SVO_MEM_HOST),
              "svo_cuda_find_match_direct");
  unpack(*this, out);
  const MatchResult r = static_cast<MatchResult>(out.result);
  if (r != MatchResult::kFailVisibility && r != MatchResult::kFailWarp) warpedPatch(*this, ref_frame, cur_frame, T, ft, d);
  if (out.result == 0) px_cur = px_cur_;
  return r;
}
Matcher::MatchResult Matcher::findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const FeatureWrapper& ref_ftr,
                                                      const double d_estimate_inv, const double d_min_inv, const double d_max_inv, double& depth) {
  return findEpipolarMatchDirect(ref_frame, cur_frame, cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse(), ref_ftr, d_estimate_inv, d_min_inv,
                                 d_max_inv, depth);  // matcher.cpp:143-155
}
Matcher::MatchResult Matcher::findEpipolarMatchDirect(const Frame& ref_frame, const Frame& cur_frame, const Transformation& T_cur_ref,
                                                      const FeatureWrapper& ref_ftr, const double d_estimate_inv, const double d_min_inv,
                                                      const double d_max_inv, double& depth) {
  const svo_feature ft = cFeature(ref_ftr);
  const svo_matcher_options o = cOptions();
  double T[7];
  T_cur_ref.toArray(T);
  const double d3[3] = {d_estimate_inv, d_min_inv, d_max_inv};
  svo_match_out out{};
  const int zero = 0;
  b200::check(svo_cuda_find_epipolar_match_direct(b200::context(), b200::ensureGpu(ref_frame).handle(), b200::ensureGpu(cur_frame).handle(), &zero,
                                                  &zero, &ref_frame.cam_->model, &cur_frame.cam_->model, T, &zero, 1, &ft, d3, &o, &out, SVO_MEM_HOST),
              "svo_cuda_find_epipolar_match_direct");
  unpack(*this, out);
  epi_image_ = {out.epi_image[0], out.epi_image[1]};  // matcher.cpp:176: set before every return
  const MatchResult r = static_cast<MatchResult>(out.result);
  if (r != MatchResult::kFailAngle && r != MatchResult::kFailWarp) warpedPatch(*this, ref_frame, cur_frame, T, ft, 1.0 / std::max(0.000001, d_estimate_inv));
  if (out.result == 0) depth = out.depth;
  return r;
}
void Matcher::scanEpipolarLine(const Frame& frame, const BearingVector& A, const BearingVector& B, const BearingVector& C,
                               const PatchScore& patch_score, const int patch_level, Keypoint* image_best, int* zmssd_best) {
  const svo_matcher_options o = cOptions();
  double best[2] = {0.0, 0.0};
  b200::check(svo_cuda_scan_epipolar_line(b200::context(), b200::ensureGpu(frame).handle(), nullptr, &frame.cam_->model, 1, A.data(), B.data(),
                                          C.data(), patch_score.ref_patch_, &patch_level, &epi_length_pyramid_, &o, best, zmssd_best, SVO_MEM_HOST),
              "svo_cuda_scan_epipolar_line");
  *image_best = {best[0], best[1]};
}
std::string Matcher::getResultString(const MatchResult& result) {  // matcher.cpp:243-260
  switch (result) {
    case MatchResult::kSuccess: return "success";
    case MatchResult::kFailScore: return "fail score";
    case MatchResult::kFailTriangulation: return "fail triangulation";
    case MatchResult::kFailVisibility: return "fail visibility";
    case MatchResult::kFailWarp: return "fail warp";
    case MatchResult::kFailAlignment: return "fail alignment";
    case MatchResult::kFailRange: return "fail range";
    case MatchResult::kFailAngle: return "fail angle";
    case MatchResult::kFailCloseView: return "fail close view";
    case MatchResult::kFailLock: return "fail lock";
    default: return "unknown";
  }
}

// ---- DepthFilter ---------------------------------------------------------------------------------------------------------------
namespace depth_filter_utils {
static svo_depth_filter_options cDepthOptions(double seed_thresh, double map_thresh, bool check_visibility, bool check_convergence, bool vogiatzis) {
  svo_depth_filter_options d{};
  d.seed_convergence_sigma2_thresh = seed_thresh;
  d.mappoint_convergence_sigma2_thresh = map_thresh;
  d.px_error_angle = 0.0;  // derived from the cur camera: getAngleError(1.0), depth_filter.cpp:383-384
  d.check_visibility = check_visibility; d.check_convergence = check_convergence; d.use_vogiatzis_update = vogiatzis;
  return d;
}
// Seeds `indices` of ref_frame against cur_frame in one launch; writes states/types back. Returns #successes.
static size_t updateSeedsOfFrame(const Frame& cur_frame, Frame& ref_frame, const std::vector<size_t>& indices, Matcher& matcher,
                                 const svo_depth_filter_options& dopt) {
  if (cur_frame.id_ == ref_frame.id_ || indices.empty()) return 0;  // "update seed with ref frame" (depth_filter.cpp:377-381)
  const int S = int(indices.size());
  std::vector<svo_feature> ft(S);
  std::vector<uint8_t> types(S);
  std::vector<double> state(4 * S), mu_range(S, ref_frame.seed_mu_range_);
  std::vector<int> zeros(S, 0);
  for (int k = 0; k < S; ++k) {
    const size_t i = indices[k];
    ft[k] = cFeature(FeatureWrapper{ref_frame.type_vec_[i], ref_frame.px_vec_[i], ref_frame.f_vec_[i], ref_frame.grad_vec_[i], ref_frame.level_vec_[i]});
    types[k] = uint8_t(ref_frame.type_vec_[i]);
    std::copy(ref_frame.invmu_sigma2_a_b_vec_[i].begin(), ref_frame.invmu_sigma2_a_b_vec_[i].end(), &state[4 * k]);
  }
  double T[7];
  (cur_frame.T_f_w_ * ref_frame.T_f_w_.inverse()).toArray(T);  // depth_filter.cpp:406
  const svo_matcher_options mo = matcher.cOptions();
  int n_success = 0;
  b200::check(svo_cuda_update_seeds(b200::context(), b200::ensureGpu(ref_frame).handle(), b200::ensureGpu(cur_frame).handle(), &ref_frame.cam_->model,
                                    &cur_frame.cam_->model, S, zeros.data(), ft.data(), types.data(), state.data(), mu_range.data(), 1, zeros.data(),
                                    zeros.data(), T, &mo, &dopt, &n_success, nullptr, SVO_MEM_HOST),
              "svo_cuda_update_seeds");
  for (int k = 0; k < S; ++k) {
    const size_t i = indices[k];
    ref_frame.type_vec_[i] = FeatureType(types[k]);
    std::copy(&state[4 * k], &state[4 * k] + 4, ref_frame.invmu_sigma2_a_b_vec_[i].begin());
  }
  return size_t(n_success);
}
bool updateSeed(const Frame& cur_frame, Frame& ref_frame, const size_t& seed_index, Matcher& matcher, const FloatType sigma2_convergence_threshold,
                const bool check_visibility, const bool check_convergence, const bool use_vogiatzis_update) {
  const svo_depth_filter_options d =
      cDepthOptions(sigma2_convergence_threshold, sigma2_convergence_threshold, check_visibility, check_convergence, use_vogiatzis_update);
  return updateSeedsOfFrame(cur_frame, ref_frame, {seed_index}, matcher, d) == 1;
}
bool updateFilterVogiatzis(const FloatType z, const FloatType tau2, const FloatType z_range, SeedState& seed) {
  uint8_t ok = 0;
  b200::check(svo_cuda_update_filter_vogiatzis(b200::context(), 1, &z, &tau2, &z_range, seed.data(), &ok, SVO_MEM_HOST), "svo_cuda_update_filter_vogiatzis");
  return ok != 0;
}
bool updateFilterGaussian(const FloatType z, const FloatType tau2, SeedState& seed) {
  uint8_t ok = 0;
  b200::check(svo_cuda_update_filter_seq(b200::context(), 1, 1, &z, &tau2, nullptr, seed.data(), &ok, 1, SVO_MEM_HOST), "svo_cuda_update_filter_seq");
  return ok != 0;
}
double computeTau(const Transformation& T_ref_cur, const BearingVector& f, const FloatType z, const FloatType px_error_angle) {
  double T[7], tau = 0.0;
  T_ref_cur.toArray(T);
  b200::check(svo_cuda_compute_tau(b200::context(), 1, T, f.data(), &z, px_error_angle, &tau, SVO_MEM_HOST), "svo_cuda_compute_tau");
  return tau;
}
}  // namespace depth_filter_utils

DepthFilter::DepthFilter(const DepthFilterOptions& options) : options_(options) {
  matcher_.options_.scan_on_unit_sphere = options.scan_epi_unit_sphere;  // depth_filter.cpp:35-41
  matcher_.options_.affine_est_offset_ = options.affine_est_offset;
  matcher_.options_.affine_est_gain_ = options.affine_est_gain;
}
size_t DepthFilter::updateSeedsOfRefFrame(const FramePtr& ref_frame, const FramePtr& cur_frame) {
  const svo_depth_filter_options d = depth_filter_utils::cDepthOptions(options_.seed_convergence_sigma2_thresh,
                                                                       options_.mappoint_convergence_sigma2_thresh, true, false, true);
  std::vector<size_t> seeds;
  for (size_t i = 0; i < ref_frame->num_features_; ++i)
    if (isSeed(ref_frame->type_vec_[i])) seeds.push_back(i);
  return depth_filter_utils::updateSeedsOfFrame(*cur_frame, *ref_frame, seeds, matcher_, d);
}
size_t DepthFilter::updateSeeds(const std::vector<FramePtr>& ref_frames_with_seeds, const FramePtr& cur_frame) {
  size_t n_success = 0;
  if (!thread_) {
    for (const FramePtr& ref_frame : ref_frames_with_seeds) n_success += updateSeedsOfRefFrame(ref_frame, cur_frame);
  } else {  // depth_filter.cpp:235-249
    std::unique_lock<std::mutex> lock(jobs_mut_);
    for (const FramePtr& ref_frame : ref_frames_with_seeds) {
      Job j;
      j.type = Job::UPDATE; j.cur_frame = cur_frame; j.ref_frame = ref_frame;
      jobs_.push(j);
    }
    jobs_condvar_.notify_all();
  }
  return n_success;
}
DepthFilter::~DepthFilter() { stopThread(); }
void DepthFilter::startThread() {
  if (thread_) return;  // "Thread already started!"
  quit_thread_ = false;
  thread_.reset(new std::thread(&DepthFilter::updateSeedsLoop, this));
}
void DepthFilter::stopThread() {
  if (!thread_) return;
  {
    std::unique_lock<std::mutex> lock(jobs_mut_);
    quit_thread_ = true;
  }
  jobs_condvar_.notify_all();
  thread_->join();
  thread_.reset();
}
void DepthFilter::reset() {
  std::unique_lock<std::mutex> lock(jobs_mut_);
  while (!jobs_.empty()) jobs_.pop();
}
void DepthFilter::waitForJobs() {
  std::unique_lock<std::mutex> lock(jobs_mut_);
  idle_condvar_.wait(lock, [this] { return !thread_ || (jobs_.empty() && !busy_); });
}
void DepthFilter::updateSeedsLoop() {  // depth_filter.cpp:145-198
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lock(jobs_mut_);
      busy_ = false;
      idle_condvar_.notify_all();
      while (jobs_.empty() && !quit_thread_) jobs_condvar_.wait(lock);
      if (quit_thread_) return;
      job = jobs_.front();
      jobs_.pop();
      busy_ = true;
    }
    try {
      if (job.type == Job::SEED_INIT) {
        std::unique_lock<std::mutex> lock(feature_detector_mut_);
        depth_filter_utils::initializeSeeds(job.cur_frame, feature_detector_, options_.max_n_seeds_per_frame, float(job.min_depth),
                                            float(job.max_depth), float(job.mean_depth));
      } else {
        updateSeedsOfRefFrame(job.ref_frame, job.cur_frame);
      }
    } catch (const std::exception& e) {  // a worker thread must not take the process down through an uncaught exception
      std::fprintf(stderr, "svo::DepthFilter worker: %s\n", e.what());
    }
  }
}

// ---- Reprojector -----------------------------------------------------------------------------------------------------------------
namespace {
// The map tables of svo_reproj_map, flattened from the keyframes a reprojection can touch: the visible keyframes and the frames
// holding the other observations of their landmarks (Point::getCloseViewObs may pick any of them, point.cpp:83-129).
struct FlatMap {
  std::vector<FramePtr> kfs;
  std::map<const Frame*, int> index;
  std::vector<int> begin;  // [K+1]
  std::vector<double> T_f_w, mu_range, score, state, pt_pos;
  std::vector<svo_feature> feat;
  std::vector<int> feat_point, feat_kf, pt_failed, pt_succeeded, obs_begin, obs_feat;
  std::vector<PointPtr> pts;
  std::map<const Point*, int> pt_index;
  std::map<int, int> cand_type;                       // candidate type at projection time (unconverged-seed pass)
  std::vector<std::array<double, 2>> remaining;       // cur_px of the candidates the last pass left in candidates_
  svo_cuda_pyr* pyr = nullptr;

  int addFrame(const FramePtr& f) {
    auto it = index.find(f.get());
    if (it != index.end()) return it->second;
    const int k = int(kfs.size());
    index[f.get()] = k;
    kfs.push_back(f);
    return k;
  }
  ~FlatMap() { if (pyr) svo_cuda_pyr_destroy(b200::context(), pyr); }
};

void buildFlatMap(const std::vector<FramePtr>& visible_kfs, FlatMap& m) {
  for (const FramePtr& f : visible_kfs) m.addFrame(f);
  for (size_t k = 0; k < m.kfs.size(); ++k) {  // grows while the observation frames of the landmarks are added
    const FramePtr f = m.kfs[k];
    for (size_t i = 0; i < f->num_features_; ++i)
      if (f->isValidLandmark(i))
        for (const KeypointIdentifier& obs : f->landmark_vec_[i]->obs_)
          if (FramePtr of = obs.frame.lock()) m.addFrame(of);
  }
  const int K = int(m.kfs.size());
  m.begin.assign(1, 0);
  for (int k = 0; k < K; ++k) {
    const FramePtr& f = m.kfs[k];
    double T[7];
    f->T_f_w_.toArray(T);
    m.T_f_w.insert(m.T_f_w.end(), T, T + 7);
    m.mu_range.push_back(f->seed_mu_range_);
    for (size_t i = 0; i < f->num_features_; ++i) {
      svo_feature q{};
      q.px[0] = f->px_vec_[i][0]; q.px[1] = f->px_vec_[i][1];
      for (int c = 0; c < 3; ++c) q.f[c] = f->f_vec_[i][c];
      q.grad[0] = f->grad_vec_[i][0]; q.grad[1] = f->grad_vec_[i][1];
      q.type = int(f->type_vec_[i]);
      q.level = f->level_vec_[i];
      m.feat.push_back(q);
      m.score.push_back(f->score_vec_[i]);
      for (int c = 0; c < 4; ++c) m.state.push_back(f->invmu_sigma2_a_b_vec_[i][c]);
      m.feat_kf.push_back(k);
      int pid = -1;
      if (f->isValidLandmark(i)) {
        const PointPtr& p = f->landmark_vec_[i];
        auto it = m.pt_index.find(p.get());
        if (it == m.pt_index.end()) {
          pid = int(m.pts.size());
          m.pt_index[p.get()] = pid;
          m.pts.push_back(p);
        } else {
          pid = it->second;
        }
      }
      m.feat_point.push_back(pid);
    }
    m.begin.push_back(int(m.feat.size()));
  }
  m.obs_begin.assign(1, 0);
  for (const PointPtr& p : m.pts) {
    for (int c = 0; c < 3; ++c) m.pt_pos.push_back(p->pos_[c]);
    m.pt_failed.push_back(p->n_failed_reproj_);
    m.pt_succeeded.push_back(p->n_succeeded_reproj_);
    for (const KeypointIdentifier& obs : p->obs_)
      if (FramePtr of = obs.frame.lock()) m.obs_feat.push_back(m.begin[m.index[of.get()]] + int(obs.keypoint_index_));
    m.obs_begin.push_back(int(m.obs_feat.size()));
  }
  // one pyramid batch holding every keyframe of the set: level 0 is copied device to device, the levels are rebuilt (same bytes)
  svo_cuda_ctx* ctx = b200::context();
  const b200::GpuPyramid& g0 = b200::ensureGpu(*m.kfs[0]);
  b200::check(svo_cuda_pyr_create(ctx, K, g0.width(), g0.height(), g0.n_levels(), -1, &m.pyr), "svo_cuda_pyr_create");
  for (int k = 0; k < K; ++k) {
    const b200::GpuPyramid& g = b200::ensureGpu(*m.kfs[k]);
    size_t pitch = 0, stride = 0;
    void* ptr = nullptr;
    svo_cuda_pyr_level_info(g.handle(), 0, nullptr, nullptr, &pitch, &stride, &ptr);
    b200::check(svo_cuda_pyr_upload(ctx, m.pyr, k, 1, static_cast<const uint8_t*>(ptr), pitch, stride, SVO_MEM_DEVICE), "svo_cuda_pyr_upload");
  }
  b200::check(svo_cuda_pyr_build(ctx, m.pyr, 0, K), "svo_cuda_pyr_build");
}
}  // namespace

void Reprojector::reprojectFrames(const FramePtr& cur_frame, const std::vector<FramePtr>& visible_kfs, std::vector<PointPtr>& trash_points) {
  if (options_.max_n_features_per_frame == 0) throw b200::Error("Reprojector: max_n_features_per_frame must be > 0");  // CHECK_GT, :38
  const svo_camera& cam = cur_frame->cam_->model;
  if (!grid_) {
    int n_cols = 0, n_rows = 0;
    svo_cuda_grid_cells(cam.width, cam.height, int(options_.cell_size), &n_cols, &n_rows);
    grid_.reset(new OccupandyGrid2D(int(options_.cell_size), n_cols, n_rows));
  }
  grid_->reset();
  stats_.reset();
  if (visible_kfs.empty()) return;
  const size_t max_total_n_features = options_.max_n_features_per_frame;  // + max_fixed_landmarks only with the global map
  FlatMap m;
  buildFlatMap(visible_kfs, m);
  cur_frame->landmark_vec_.resize(cur_frame->num_features_);
  cur_frame->seed_ref_vec_.resize(cur_frame->num_features_);
  const b200::GpuPyramid& cur_gpu = b200::ensureGpu(*cur_frame);
  double T_cur[7];
  cur_frame->T_f_w_.toArray(T_cur);
  const auto cur_pos = cur_frame->pos();

  svo_reproj_map tables{};
  tables.n_kfs = int(m.kfs.size()); tables.n_feat = int(m.feat.size()); tables.n_points = int(m.pts.size()); tables.n_obs = int(m.obs_feat.size());
  auto bindTables = [&]() {
    tables.kf_T_f_w = m.T_f_w.data(); tables.kf_seed_mu_range = m.mu_range.data(); tables.kf_frame_idx = nullptr;
    tables.feat = m.feat.data(); tables.feat_score = m.score.data(); tables.feat_seed_state = m.state.data();
    tables.feat_point = m.feat_point.data(); tables.feat_kf = m.feat_kf.data(); tables.pt_pos = m.pt_pos.data();
    tables.pt_n_failed = m.pt_failed.data(); tables.pt_n_succeeded = m.pt_succeeded.data(); tables.pt_obs_begin = m.obs_begin.data();
    tables.obs_feat = m.obs_feat.data();
  };
  // One matchCandidates round (getCandidate + sort + match) on the GPU; project_only = only setGridCellsOccupied(candidates)
  // (reprojector.cpp:545-555), done by handing over a fully occupied grid so that nothing is tried.
  auto pass = [&](const std::vector<int>& entries, size_t max_n, bool project_only, Statistics* st) {
    if (entries.empty()) return;
    bindTables();
    svo_reprojector_options o{};
    o.cell_size = int(options_.cell_size); o.max_n_features = int(max_n);
    o.affine_est_offset = options_.affine_est_offset; o.affine_est_gain = options_.affine_est_gain;
    o.seed_sigma2_thresh = options_.seed_sigma2_thresh;
    std::vector<uint8_t> occ = grid_->occupancy_;
    if (project_only) std::fill(occ.begin(), occ.end(), 1);
    const int n_in = int(cur_frame->num_features_);
    const int eb[2] = {0, int(entries.size())};
    std::vector<svo_reproj_result> res(entries.size());
    svo_reproj_stats rs{};
    b200::check(svo_cuda_reproject_match(b200::context(), m.pyr, cur_gpu.handle(), &cam, &cam, &tables, 1, nullptr, T_cur, &n_in, eb,
                                         int(entries.size()), entries.data(), occ.data(), &o, res.data(), &rs, SVO_MEM_HOST),
                "svo_cuda_reproject_match");
    if (project_only) {
      for (const svo_reproj_result& r : res)
        if (r.status != SVO_REPROJ_NOT_CANDIDATE) grid_->occupancy_[grid_->getCellIndex(int(r.cur_px[0]), int(r.cur_px[1]), 1)] = 1;
      return;
    }
    grid_->occupancy_ = occ;
    if (st) { st->n_trials += size_t(rs.n_trials); st->n_matches += size_t(rs.n_matches); }
    // everything the reference mutates: landmark counters, seed states / types of the keyframes, the new features of the frame
    std::vector<int> matched;
    for (size_t e = 0; e < res.size(); ++e) {
      const svo_reproj_result& r = res[e];
      const int fi = entries[e], k = m.feat_kf[fi], i = fi - m.begin[k], pid = m.feat_point[fi];
      if (pid >= 0) {
        m.pts[pid]->n_failed_reproj_ += r.d_failed; m.pts[pid]->n_succeeded_reproj_ += r.d_succeeded;
        m.pt_failed[pid] = m.pts[pid]->n_failed_reproj_; m.pt_succeeded[pid] = m.pts[pid]->n_succeeded_reproj_;
      } else {
        for (int c = 0; c < 4; ++c) { m.kfs[k]->invmu_sigma2_a_b_vec_[i][c] = r.seed_state[c]; m.state[4 * size_t(fi) + c] = r.seed_state[c]; }
        m.kfs[k]->type_vec_[i] = FeatureType(r.type_out);
        m.feat[fi].type = r.type_out;
      }
      if (r.status == SVO_REPROJ_MATCHED) matched.push_back(int(e));
    }
    std::sort(matched.begin(), matched.end(), [&](int a, int b) { return res[a].slot < res[b].slot; });
    for (int e : matched) {  // matchCandidate, reprojector.cpp:462-484 (the candidate's type is the one it had when it was projected)
      const svo_reproj_result& r = res[e];
      const int fi = entries[e], k = m.feat_kf[fi], i = fi - m.begin[k], pid = m.feat_point[fi];
      cur_frame->type_vec_.push_back(pid >= 0 ? m.kfs[k]->type_vec_[i] : FeatureType(m.cand_type.count(fi) ? m.cand_type[fi] : r.type_out));
      cur_frame->px_vec_.push_back({r.px[0], r.px[1]});
      cur_frame->f_vec_.push_back({r.f[0], r.f[1], r.f[2]});
      cur_frame->grad_vec_.push_back({r.grad[0], r.grad[1]});
      cur_frame->level_vec_.push_back(r.level);
      cur_frame->score_vec_.push_back(m.score[fi]);
      cur_frame->invmu_sigma2_a_b_vec_.push_back({r.seed_state[0], r.seed_state[1], r.seed_state[2], r.seed_state[3]});
      cur_frame->landmark_vec_.push_back(pid >= 0 ? m.pts[pid] : nullptr);
      SeedRef sr;
      std::array<double, 3> xyz_w;
      if (pid >= 0) {
        xyz_w = m.pts[pid]->pos_;
      } else {
        sr.keyframe = m.kfs[k]; sr.seed_id = i;
        const Transformation T_w_ref = m.kfs[k]->T_f_w_.inverse();
        const double d = 1.0 / r.seed_state[0];
        const auto v = rotate(T_w_ref.q, {m.feat[fi].f[0] * d, m.feat[fi].f[1] * d, m.feat[fi].f[2] * d});
        xyz_w = {v[0] + T_w_ref.t[0], v[1] + T_w_ref.t[1], v[2] + T_w_ref.t[2]};
      }
      cur_frame->seed_ref_vec_.push_back(sr);
      cur_frame->depth_vec_.push_back(std::sqrt((xyz_w[0] - cur_pos[0]) * (xyz_w[0] - cur_pos[0]) + (xyz_w[1] - cur_pos[1]) * (xyz_w[1] - cur_pos[1]) +
                                                (xyz_w[2] - cur_pos[2]) * (xyz_w[2] - cur_pos[2])));
      ++cur_frame->num_features_;
    }
    // candidates the loop never reached stay in candidates_ (reprojector.cpp:380): remembered for setGridCellsOccupied
    m.remaining.clear();
    for (const svo_reproj_result& r : res)
      if (r.status == SVO_REPROJ_NOT_REACHED) m.remaining.push_back({r.cur_px[0], r.cur_px[1]});
  };
  auto occupyRemaining = [&]() {
    for (const auto& px : m.remaining) grid_->occupancy_[grid_->getCellIndex(int(px[0]), int(px[1]), 1)] = 1;
  };

  // ---- landmarks (reprojector.cpp:133-190)
  std::vector<int> entries;
  for (const FramePtr& ref_frame : visible_kfs) {
    const int k = m.index[ref_frame.get()];
    for (size_t i = 0; i < ref_frame->num_features_; ++i) {
      const FeatureType type = ref_frame->type_vec_[i];
      if (!ref_frame->isValidLandmark(i) || type == FeatureType::kOutlier || isMapPoint(type) || isFixedLandmark(type)) continue;
      const PointPtr& point = ref_frame->landmark_vec_[i];
      if (point->n_failed_reproj_ > 10) { trash_points.push_back(point); continue; }
      if (point->last_projected_kf_id_.at(camera_index_) == cur_frame->id_) continue;
      point->last_projected_kf_id_[camera_index_] = cur_frame->id_;
      if (point->obs_.size() < 2 && options_.remove_unconstrained_points) { trash_points.push_back(point); continue; }
      entries.push_back(m.begin[k] + int(i));
    }
  }
  Statistics lm_stats;
  pass(entries, max_total_n_features, false, &lm_stats);
  stats_.add(lm_stats);
  if (doesFrameHaveEnoughFeatures(cur_frame)) occupyRemaining();

  // ---- converged seeds (:192-235)
  auto seedEntries = [&](bool converged) {
    entries.clear();
    for (const FramePtr& ref_frame : visible_kfs) {
      const int k = m.index[ref_frame.get()];
      for (size_t i = 0; i < ref_frame->num_features_; ++i)
        if (converged ? isConvergedCornerEdgeletSeed(ref_frame->type_vec_[i]) : isUnconvergedCornerEdgeletSeed(ref_frame->type_vec_[i]))
          entries.push_back(m.begin[k] + int(i));
    }
  };
  seedEntries(true);
  if (doesFrameHaveEnoughFeatures(cur_frame)) {
    pass(entries, max_total_n_features, true, nullptr);
    return;
  }
  m.remaining.clear();
  Statistics sd_stats;
  pass(entries, max_total_n_features, false, &sd_stats);
  stats_.add(sd_stats);
  if (doesFrameHaveEnoughFeatures(cur_frame) || !options_.reproject_unconverged_seeds) {
    occupyRemaining();
    return;
  }

  // ---- unconverged seeds (:237-300)
  seedEntries(false);
  size_t max_allowed_total = max_total_n_features;
  if (options_.max_unconverged_seeds_ratio > 0) {
    const double min_lm_seeds_ratio = 1 - options_.max_unconverged_seeds_ratio;
    const size_t alt = static_cast<size_t>(cur_frame->numTrackedFeatures() / min_lm_seeds_ratio);
    if (max_allowed_total > alt) max_allowed_total = alt;
  }
  if (max_allowed_total < options_.min_required_features) max_allowed_total = options_.min_required_features;
  m.remaining.clear();
  Statistics un_sd_stats;
  // the candidate's type is read before updateSeed may converge it (matchCandidate: feature.type = c.type)
  m.cand_type.clear();
  for (int fi : entries) m.cand_type[fi] = m.feat[fi].type;
  pass(entries, max_allowed_total, false, &un_sd_stats);
  stats_.add(un_sd_stats);
  if (doesFrameHaveEnoughFeatures(cur_frame)) occupyRemaining();
}

// ---- PoseOptimizer -----------------------------------------------------------------------------------------------------------------
PoseOptimizer::SolverOptions PoseOptimizer::getDefaultSolverOptions() {
  SolverOptions options;
  options.max_iter = 10;
  options.eps = 0.000001;
  return options;
}
void PoseOptimizer::setRotationPrior(const std::array<double, 4>& R_frame_world, double lambda) {
  have_prior_ = true;
  prior_q_ = R_frame_world;
  prior_lambda_ = lambda;
}
size_t PoseOptimizer::run(const FrameBundle::Ptr& frame_bundle, double reproj_thresh_px) {
  if (!frame_bundle || frame_bundle->empty()) throw b200::Error("PoseOptimizer: FrameBundle is empty");  // CHECK, :41
  const size_t n_cams = frame_bundle->size();
  if (n_cams > SVO_MAX_CAMS) throw b200::Error("PoseOptimizer: too many cameras");
  std::vector<svo_camera> cams;
  std::vector<double> T_cam_imu, xyz;
  std::vector<svo_feature> ftrs;
  std::vector<int> feat_cam;
  std::vector<uint8_t> has;
  for (size_t c = 0; c < n_cams; ++c) {
    const Frame& f = *frame_bundle->at(c);
    cams.push_back(f.cam_->model);
    double T[7];
    f.T_cam_imu_.toArray(T);
    T_cam_imu.insert(T_cam_imu.end(), T, T + 7);
    for (size_t i = 0; i < f.num_features_; ++i) {
      svo_feature q{};
      q.px[0] = f.px_vec_[i][0]; q.px[1] = f.px_vec_[i][1];
      for (int k = 0; k < 3; ++k) q.f[k] = f.f_vec_[i][k];
      q.grad[0] = f.grad_vec_[i][0]; q.grad[1] = f.grad_vec_[i][1];
      q.type = int(f.type_vec_[i]);
      q.level = f.level_vec_[i];
      std::array<double, 3> p{{0, 0, 0}};
      uint8_t h = 0;
      if (f.isValidLandmark(i)) {  // evaluateErrorImpl, :123-137
        p = f.landmark_vec_[i]->pos_;
        h = 1;
      } else if (isCornerEdgeletSeed(f.type_vec_[i]) && i < f.seed_ref_vec_.size() && f.seed_ref_vec_[i].keyframe) {
        const Frame& kf = *f.seed_ref_vec_[i].keyframe;
        const int id = f.seed_ref_vec_[i].seed_id;
        const Transformation T_w_kf = kf.T_f_w_.inverse();
        const double d = 1.0 / kf.invmu_sigma2_a_b_vec_[id][0];
        const auto v = rotate(T_w_kf.q, {kf.f_vec_[id][0] * d, kf.f_vec_[id][1] * d, kf.f_vec_[id][2] * d});
        p = {v[0] + T_w_kf.t[0], v[1] + T_w_kf.t[1], v[2] + T_w_kf.t[2]};
        h = 1;
      }
      ftrs.push_back(q); feat_cam.push_back(int(c)); has.push_back(h);
      xyz.insert(xyz.end(), p.begin(), p.end());
    }
  }
  if (ftrs.empty()) throw b200::Error("PoseOptimizer: No features in frames");  // CHECK_GT, :42
  double T0[7];
  frame_bundle->at(0)->T_imu_world().toArray(T0);
  const int feat_begin[2] = {0, int(ftrs.size())};
  svo_pose_optimizer_options o{};
  o.err_type = int(err_type_); o.max_iter = int(solver_options_.max_iter); o.eps = solver_options_.eps;
  o.reproj_thresh_px = reproj_thresh_px; o.prior_lambda = prior_lambda_;
  svo_pose_opt_result r{};
  std::vector<uint8_t> outlier(ftrs.size());
  b200::check(svo_cuda_pose_optimize(b200::context(), int(n_cams), cams.data(), T_cam_imu.data(), 1, T0, feat_begin, int(ftrs.size()), ftrs.data(),
                                     feat_cam.data(), xyz.data(), has.data(), have_prior_ ? prior_q_.data() : nullptr, &o, &r, outlier.data(),
                                     SVO_MEM_HOST), "svo_cuda_pose_optimize");
  size_t k = 0;
  for (size_t c = 0; c < n_cams; ++c) {
    Frame& f = *frame_bundle->at(c);
    f.T_f_w_ = Transformation::fromArray(r.T_f_w[c]);
    for (size_t i = 0; i < f.num_features_; ++i, ++k)
      if (outlier[k]) {  // removeOutliers, :283-290
        f.type_vec_[i] = FeatureType::kOutlier;
        if (i < f.seed_ref_vec_.size()) f.seed_ref_vec_[i].keyframe.reset();
        if (i < f.landmark_vec_.size()) f.landmark_vec_[i] = nullptr;
      }
  }
  measurement_sigma_ = r.measurement_sigma;
  stats_.reproj_error_before = r.reproj_error_before;
  stats_.reproj_error_after = r.reproj_error_after;
  iter_ = size_t(r.iters);
  return size_t(r.n_meas_final);
}

// ---- FAST detector ----------------------------------------------------------------------------------------------------------------
namespace feature_detection_utils {
void fastDetector(const b200::GpuPyramid& gpu, const int threshold, const int border, const size_t min_level, const size_t max_level,
                  Corners& corners, OccupandyGrid2D& grid) {
  if (corners.size() != grid.occupancy_.size() || int(max_level) > gpu.n_levels() - 1)
    throw b200::Error("fastDetector: corners/grid size mismatch or max_level beyond the pyramid");  // CHECKs of :154-155
  svo_detector_options o{threshold, border, int(min_level), int(max_level), grid.cell_size, 10};
  std::vector<svo_corner> out(corners.size());
  b200::check(svo_cuda_fast_detect(b200::context(), gpu.handle(), 0, 1, &o, grid.occupancy_.data(), out.data(), SVO_MEM_HOST), "svo_cuda_fast_detect");
  for (size_t k = 0; k < out.size(); ++k)
    if (out[k].score > corners[k].score) corners[k] = Corner(out[k].x, out[k].y, out[k].score, out[k].level, out[k].angle);
}
void edgeletDetector_V2(const b200::GpuPyramid& gpu, const int threshold, const int border, const int /*min_level*/, const int /*max_level*/,
                        Corners& corners, OccupandyGrid2D& grid) {
  if (corners.size() != grid.occupancy_.size()) throw b200::Error("edgeletDetector_V2: corners/grid size mismatch");  // CHECK_EQ of :322
  std::vector<svo_corner> out(corners.size());
  b200::check(svo_cuda_edgelet_detect(b200::context(), gpu.handle(), 0, 1, threshold, border, grid.cell_size, grid.occupancy_.data(),
                                      out.data(), SVO_MEM_HOST), "svo_cuda_edgelet_detect");
  for (size_t k = 0; k < out.size(); ++k)
    if (out[k].score > corners[k].score) corners[k] = Corner(out[k].x, out[k].y, out[k].score, out[k].level, out[k].angle);
}

void fillFeatures(const Corners& corners, const FeatureType& type, const double& threshold, const size_t max_n_features,
                  Keypoints& keypoints, Scores& scores, Levels& levels, Gradients& gradients, FeatureTypes& types, OccupandyGrid2D& grid) {
  std::vector<size_t> idx;
  for (size_t k = 0; k < corners.size(); ++k)
    if (corners[k].score > threshold) {
      idx.push_back(k);
      grid.occupancy_[grid.getCellIndex(corners[k].x, corners[k].y)] = 1;
    }
  std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) { return corners[a].score > corners[b].score; });
  if (idx.size() > max_n_features) idx.resize(max_n_features);
  for (size_t k : idx) {
    const Corner& c = corners[k];
    keypoints.push_back({double(c.x), double(c.y)});
    gradients.push_back({std::cos(c.angle), std::sin(c.angle)});  // float overloads, as feature_detection_utils.cpp:101
    scores.push_back(c.score);
    levels.push_back(c.level);
    types.push_back(type);
  }
}

AbstractDetector::Ptr makeDetector(const DetectorOptions& options, const CameraPtr& cam) {
  switch (options.detector_type) {
    case DetectorType::kFast: return std::make_shared<FastDetector>(options, cam);
    case DetectorType::kFastGrad: return std::make_shared<FastGradDetector>(options, cam);
    case DetectorType::kGridGrad: return std::make_shared<GradientDetectorGrid>(options, cam);
    default: throw b200::Error("makeDetector: only kFast, kFastGrad and kGridGrad are implemented on the B200 path");
  }
}
}  // namespace feature_detection_utils

AbstractDetector::AbstractDetector(const DetectorOptions& options, const CameraPtr& cam)
    : options_(options),
      grid_(int(options.cell_size), int(std::ceil(double(cam->imageWidth()) / options.cell_size)),
            int(std::ceil(double(cam->imageHeight()) / options.cell_size))),
      // feature_detection.cpp:34-39: cell_size / sec_grid_fineness, ceil(width / (cell_size / fineness)) x ceil(height / ...)
      closeness_check_grid_(int(options.cell_size / std::max<size_t>(options.sec_grid_fineness, 1)),
                            int(std::ceil(double(cam->imageWidth()) / double(options.cell_size / std::max<size_t>(options.sec_grid_fineness, 1)))),
                            int(std::ceil(double(cam->imageHeight()) / double(options.cell_size / std::max<size_t>(options.sec_grid_fineness, 1))))) {}

void FastDetector::detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec, Levels& level_vec,
                          Gradients& grad_vec, FeatureTypes& types_vec) {
  Corners corners(size_t(grid_.n_cols) * grid_.n_rows, Corner(0, 0, float(options_.threshold_primary), 0, 0.0f));
  feature_detection_utils::fastDetector(gpu, int(options_.threshold_primary), options_.border, options_.min_level, options_.max_level,
                                        corners, grid_);
  feature_detection_utils::fillFeatures(corners, FeatureType::kCorner, options_.threshold_primary, max_n_features, px_vec, score_vec,
                                        level_vec, grad_vec, types_vec, grid_);
  resetGrid();
}

void GradientDetectorGrid::detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec,
                                  Levels& level_vec, Gradients& grad_vec, FeatureTypes& types_vec) {
  Corners corners(size_t(grid_.n_cols) * grid_.n_rows, Corner(0, 0, float(options_.threshold_secondary), 0, 0.0f));
  feature_detection_utils::edgeletDetector_V2(gpu, int(options_.threshold_secondary), options_.border, options_.min_level,
                                              options_.max_level, corners, grid_);
  feature_detection_utils::fillFeatures(corners, FeatureType::kEdgelet, options_.threshold_secondary, max_n_features, px_vec, score_vec,
                                        level_vec, grad_vec, types_vec, grid_);
  resetGrid();
}

void FastGradDetector::detect(const b200::GpuPyramid& gpu, const size_t max_n_features, Keypoints& px_vec, Scores& score_vec,
                              Levels& level_vec, Gradients& grad_vec, FeatureTypes& types_vec) {
  // One device call covers both stages (FAST -> cells with a corner become occupied -> edgelets in the remaining cells, skipped when
  // the corners already reach max_n_features); the two fillFeatures passes of feature_detection.cpp:166-190 follow on the host.
  const size_t n_cells = size_t(grid_.n_cols) * grid_.n_rows;
  if (int(options_.max_level) > gpu.n_levels() - 1) throw b200::Error("FastGradDetector: max_level beyond the pyramid");
  svo_detector_options o{int(options_.threshold_primary), options_.border, options_.min_level, options_.max_level, grid_.cell_size, 10};
  std::vector<svo_corner> fast(n_cells), edge(n_cells);
  b200::check(svo_cuda_fastgrad_detect(b200::context(), gpu.handle(), 0, 1, &o, int(options_.threshold_secondary), int(max_n_features),
                                       grid_.occupancy_.data(), fast.data(), edge.data(), SVO_MEM_HOST), "svo_cuda_fastgrad_detect");
  Corners corners(n_cells, Corner(0, 0, float(options_.threshold_primary), 0, 0.0f));
  for (size_t k = 0; k < n_cells; ++k)
    if (fast[k].score > corners[k].score) corners[k] = Corner(fast[k].x, fast[k].y, fast[k].score, fast[k].level, fast[k].angle);
  const size_t n_before = px_vec.size();
  feature_detection_utils::fillFeatures(corners, FeatureType::kCorner, options_.threshold_primary, max_n_features, px_vec, score_vec,
                                        level_vec, grad_vec, types_vec, grid_);
  const long max_features = long(max_n_features) - long(px_vec.size() - n_before) - long(n_before);
  if (max_features > 0) {
    Corners edgelets(n_cells, Corner(0, 0, float(options_.threshold_secondary), 0, 0.0f));
    for (size_t k = 0; k < n_cells; ++k)
      if (edge[k].score > edgelets[k].score) edgelets[k] = Corner(edge[k].x, edge[k].y, edge[k].score, edge[k].level, edge[k].angle);
    feature_detection_utils::fillFeatures(edgelets, FeatureType::kEdgelet, options_.threshold_secondary, size_t(max_features), px_vec,
                                          score_vec, level_vec, grad_vec, types_vec, grid_);
  }
  resetGrid();
}

// frame_utils::computeNormalizedBearingVectors (frame.cpp:427-439) for one keypoint: f = normalize(backProject3(px)); <= n_cells
// keypoints per keyframe, host glue (pinhole_projection.hpp:30-41, radial_tangential_distortion.h:80-95)
static BearingVector normalizedBearing(const svo_camera& cm, const Keypoint& px) {
  double x = (px[0] - cm.cx) * (1.0 / cm.fx), y = (px[1] - cm.cy) * (1.0 / cm.fy);
  if (cm.distortion) {
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; ++it) {
      const double xx = x * x, yy = y * y, xy = x * y, xy2 = 2 * xy, r2 = xx + yy;
      const double icdist = 1.0 / (1.0 + (cm.k1 + cm.k2 * r2) * r2);
      const double ddx = cm.p1 * xy2 + cm.p2 * (r2 + 2.0 * xx), ddy = cm.p2 * xy2 + cm.p1 * (r2 + 2.0 * yy);
      x = (x0 - ddx) * icdist;
      y = (y0 - ddy) * icdist;
    }
  }
  const double n = std::sqrt(x * x + y * y + 1.0);
  return {x / n, y / n, 1.0 / n};
}

void AbstractDetector::detect(const FramePtr& frame) {
  // feature_detection.cpp:40-50: detect into the frame's columns, then frame_utils::computeNormalizedBearingVectors
  const size_t n_old = frame->px_vec_.size();
  detect(b200::ensureGpu(*frame), grid_.size(), frame->px_vec_, frame->score_vec_, frame->level_vec_, frame->grad_vec_, frame->type_vec_);
  const svo_camera& cm = frame->cam_->model;
  for (size_t i = n_old; i < frame->px_vec_.size(); ++i) {
    frame->depth_vec_.push_back(-1.0);
    frame->invmu_sigma2_a_b_vec_.push_back({0, 0, 0, 0});
    frame->f_vec_.push_back(normalizedBearing(cm, frame->px_vec_[i]));
  }
  frame->num_features_ = frame->px_vec_.size();
}

// ---- DepthFilter::addKeyframe / initializeSeeds ---------------------------------------------------------------------------------------
DepthFilter::DepthFilter(const DepthFilterOptions& options, const DetectorOptions& detector_options, const CameraPtr& cam)
    : DepthFilter(options) {
  feature_detector_ = feature_detection_utils::makeDetector(detector_options, cam);
}

void DepthFilter::addKeyframe(const FramePtr& frame, const double depth_mean, const double depth_min, const double depth_max) {
  if (!feature_detector_) throw b200::Error("DepthFilter::addKeyframe: no feature detector (use the detector-building constructor)");
  if (!thread_) {
    std::unique_lock<std::mutex> lock(feature_detector_mut_);
    depth_filter_utils::initializeSeeds(frame, feature_detector_, options_.max_n_seeds_per_frame, float(depth_min), float(depth_max),
                                        float(depth_mean));
  } else {  // depth_filter.cpp:125-133: clear all other jobs, this one has priority
    std::unique_lock<std::mutex> lock(jobs_mut_);
    while (!jobs_.empty()) jobs_.pop();
    Job j;
    j.type = Job::SEED_INIT; j.cur_frame = frame; j.min_depth = depth_min; j.max_depth = depth_max; j.mean_depth = depth_mean;
    jobs_.push(j);
    jobs_condvar_.notify_all();
  }
}

namespace depth_filter_utils {
void initializeSeeds(const FramePtr& frame, const std::shared_ptr<AbstractDetector>& feature_detector, const size_t max_n_seeds,
                     const float depth_min, const float depth_max, const float depth_mean) {
  const int max_n_features = int(max_n_seeds) - int(frame->numFeatures());
  if (max_n_features <= 0) return;  // "Have already enough features."
  Keypoints new_px; Scores new_scores; Levels new_levels; Gradients new_grads; FeatureTypes new_types;
  feature_detector->detect(b200::ensureGpu(*frame), size_t(max_n_features), new_px, new_scores, new_levels, new_grads, new_types);
  // (the reference writes straight into the frame when it has no features yet and through temporaries otherwise, :283-320; the
  // resulting columns are the same)
  (void)depth_max;
  frame->seed_mu_range_ = 1.0 / double(depth_min);                                 // seed::getMeanRangeFromDepthMinMax (seed.h:135-138)
  const double mu = 1.0 / double(depth_mean);                                       // seed::getMeanFromDepth (seed.h:130-133)
  const double sigma2 = frame->seed_mu_range_ * frame->seed_mu_range_ / 36.0;       // seed::getInitSigma2FromMuRange (seed.h:140-143)
  frame->landmark_vec_.resize(frame->numFeatures(), nullptr);
  frame->seed_ref_vec_.resize(frame->numFeatures());
  for (size_t i = 0; i < new_px.size(); ++i) {
    FeatureType t;
    if (new_types[i] == FeatureType::kCorner) t = FeatureType::kCornerSeed;
    else if (new_types[i] == FeatureType::kEdgelet) t = FeatureType::kEdgeletSeed;
    else if (new_types[i] == FeatureType::kMapPoint) t = FeatureType::kMapPointSeed;
    else throw b200::Error("initializeSeeds: unknown feature type");
    frame->px_vec_.push_back(new_px[i]);
    frame->f_vec_.push_back(normalizedBearing(frame->cam_->model, new_px[i]));
    frame->grad_vec_.push_back(new_grads[i]);
    frame->score_vec_.push_back(new_scores[i]);
    frame->level_vec_.push_back(new_levels[i]);
    frame->type_vec_.push_back(t);
    frame->depth_vec_.push_back(-1.0);
    frame->invmu_sigma2_a_b_vec_.push_back({mu, sigma2, 10.0, 10.0});
    frame->landmark_vec_.push_back(nullptr);
    frame->seed_ref_vec_.push_back(SeedRef());
  }
  frame->num_features_ = frame->px_vec_.size();
}
}  // namespace depth_filter_utils

// ---- Point::optimize / optimizeStructure ------------------------------------------------------------------------------------------
static void optimizePointsOnDevice(const std::vector<Point*>& pts, size_t n_iter, bool sphere) {
  std::vector<double> pos, obs_f, T_f_w;
  std::vector<int> begin{0}, obs_frame;
  std::vector<const Frame*> frames;  // observing frames, deduplicated
  for (Point* pt : pts) {
    for (const KeypointIdentifier& o : pt->obs_) {
      const FramePtr fr = o.frame.lock();
      if (!fr) continue;  // "could not unlock weak_ptr<Frame>" (point.cpp:292): the observation is skipped
      size_t k = 0;
      while (k < frames.size() && frames[k] != fr.get()) ++k;
      if (k == frames.size()) {
        frames.push_back(fr.get());
        double T[7];
        fr->T_f_w_.toArray(T);
        T_f_w.insert(T_f_w.end(), T, T + 7);
      }
      obs_frame.push_back(int(k));
      const BearingVector& f = fr->f_vec_.at(o.keypoint_index_);
      obs_f.insert(obs_f.end(), f.begin(), f.end());
    }
    // the reference tests obs_.size() < 2 (expired frames included): keep such a point out by giving it no observations
    if (pt->obs_.size() < 2) { obs_frame.resize(size_t(begin.back())); obs_f.resize(size_t(begin.back()) * 3); }
    begin.push_back(int(obs_frame.size()));
    pos.insert(pos.end(), pt->pos_.begin(), pt->pos_.end());
  }
  if (pts.empty()) return;
  b200::check(svo_cuda_optimize_points(b200::context(), int(pts.size()), pos.data(), begin.data(), int(obs_frame.size()), obs_frame.data(),
                                       obs_f.data(), int(frames.size()), T_f_w.data(), int(n_iter), sphere ? 1 : 0, nullptr, SVO_MEM_HOST),
              "svo_cuda_optimize_points");
  for (size_t i = 0; i < pts.size(); ++i) pts[i]->pos_ = {pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
}

void Point::optimize(const size_t n_iter, bool using_bearing_vector) { optimizePointsOnDevice({this}, n_iter, using_bearing_vector); }

void optimizeStructure(const FrameBundle::Ptr& frames, int max_n_pts, int max_iter, bool optimize_on_sphere) {
  if (max_n_pts == 0) return;
  for (const FramePtr& frame : frames->frames_) {
    std::vector<Point*> pts;
    for (size_t i = 0; i < frame->num_features_; ++i) {
      if (!frame->isValidLandmark(i) || isEdgelet(frame->type_vec_[i])) continue;
      pts.push_back(frame->landmark_vec_[i].get());
    }
    if (max_n_pts > 0) {  // favour points that have not been optimised in a while (only an ordering in the reference, see the header)
      const size_t n = std::min(size_t(max_n_pts), pts.size());
      std::nth_element(pts.begin(), pts.begin() + n, pts.end(),
                       [](const Point* l, const Point* r) { return l->last_structure_optim_ < r->last_structure_optim_; });
    }
    // a landmark seen twice in one frame would be optimised twice in a row by the reference; the batch runs every point once per
    // occurrence in order, so duplicates go through separate calls
    std::vector<Point*> batch;
    for (Point* p : pts) {
      if (std::find(batch.begin(), batch.end(), p) != batch.end()) { optimizePointsOnDevice(batch, size_t(max_iter), optimize_on_sphere); batch.clear(); }
      batch.push_back(p);
    }
    optimizePointsOnDevice(batch, size_t(max_iter), optimize_on_sphere);
    for (Point* p : pts) p->last_structure_optim_ = frame->id_;
  }
}

// ---- FeatureTracker -------------------------------------------------------------------------------------------------------------------
namespace feature_alignment {
void alignPyr2DVec(const Frame& ref_frame, const Frame& cur_frame, int max_level, int min_level, const std::vector<int>& patch_sizes,
                   int n_iter, float min_update_squared, const std::vector<std::array<int, 2>>& px_ref, std::vector<Keypoint>& px_cur,
                   std::vector<uint8_t>& status) {
  const int M = int(px_ref.size());
  status.assign(size_t(M), 0);
  if (M == 0) return;
  std::vector<int> ps(SVO_MAX_LEVELS, 8);
  for (size_t l = 0; l < patch_sizes.size() && l < ps.size(); ++l) ps[l] = patch_sizes[l];
  b200::check(svo_cuda_align_pyr2d(b200::context(), b200::ensureGpu(ref_frame).handle(), b200::ensureGpu(cur_frame).handle(), nullptr, nullptr, M,
                                   &px_ref[0][0], &px_cur[0][0], max_level, min_level, ps.data(), n_iter, min_update_squared, status.data(),
                                   SVO_MEM_HOST), "svo_cuda_align_pyr2d");
}
}  // namespace feature_alignment

int PointIdProvider::getNewPointId() {
  static std::atomic<int> last_id{0};
  return last_id.fetch_add(1);
}

double FeatureTrack::getDisparity() const {
  const Keypoint &a = front().getPx(), &b = back().getPx();
  return std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]));
}

namespace feature_tracking_utils {
double getTracksDisparityPercentile(const FeatureTracks& tracks, double pivot_ratio) {
  if (!(pivot_ratio > 0.0 && pivot_ratio < 1.0)) throw b200::Error("getTracksDisparityPercentile: pivot_ratio needs to be in (0,1)");
  if (tracks.empty()) return 0.0;
  std::vector<double> disparities;
  disparities.reserve(tracks.size());
  for (const FeatureTrack& track : tracks) disparities.push_back(track.getDisparity());
  const size_t pivot = size_t(std::floor(pivot_ratio * disparities.size()));
  std::nth_element(disparities.begin(), disparities.begin() + pivot, disparities.end(), std::greater<double>());
  return disparities[pivot];
}
}  // namespace feature_tracking_utils

FeatureTracker::FeatureTracker(const FeatureTrackerOptions& options, const DetectorOptions& detector_options, const std::vector<CameraPtr>& cams)
    : options_(options), bundle_size_(cams.size()), active_tracks_(cams.size()), terminated_tracks_(cams.size()) {
  for (const CameraPtr& cam : cams) detectors_.push_back(feature_detection_utils::makeDetector(detector_options, cam));
}

void FeatureTracker::trackAndDetect(const FrameBundle::Ptr& nframe_kp1) {
  const size_t n_tracked = trackFrameBundle(nframe_kp1);
  if (n_tracked < options_.min_tracks_to_detect_new_features) {
    if (options_.reset_before_detection) {
      resetActiveTracks();
      for (const FramePtr& frame : nframe_kp1->frames_) frame->clearFeatureStorage();
    }
    initializeNewTracks(nframe_kp1);
  }
}

size_t FeatureTracker::trackFrameBundle(const FrameBundle::Ptr& nframe_kp1) {
  resetTerminatedTracks();
  for (size_t frame_index = 0; frame_index < bundle_size_; ++frame_index) {
    FeatureTracks& tracks = active_tracks_.at(frame_index);
    const FramePtr& cur_frame = nframe_kp1->frames_.at(frame_index);
    const size_t n = tracks.size();
    // KLT of all tracks, batched per template frame
    std::vector<Keypoint> cur_px(n);
    std::vector<uint8_t> success(n, 0);
    std::vector<const Frame*> ref_frames(n);
    for (size_t t = 0; t < n; ++t) {
      const FeatureRef& ref = options_.klt_template_is_first_observation ? tracks[t].front() : tracks[t].back();
      ref_frames[t] = ref.getFrame().get();
      cur_px[t] = tracks[t].back().getPx();
    }
    std::vector<uint8_t> done(n, 0);
    for (size_t t0 = 0; t0 < n; ++t0) {
      if (done[t0]) continue;
      std::vector<size_t> sel;
      std::vector<std::array<int, 2>> px_ref;
      std::vector<Keypoint> px_cur;
      for (size_t t = t0; t < n; ++t)
        if (!done[t] && ref_frames[t] == ref_frames[t0]) {
          const FeatureRef& ref = options_.klt_template_is_first_observation ? tracks[t].front() : tracks[t].back();
          sel.push_back(t);
          px_ref.push_back({int(ref.getPx()[0]), int(ref.getPx()[1])});  // getPx().cast<int>() (:80)
          px_cur.push_back(cur_px[t]);
          done[t] = 1;
        }
      std::vector<uint8_t> st;
      feature_alignment::alignPyr2DVec(*ref_frames[t0], *cur_frame, options_.klt_max_level, options_.klt_min_level, options_.klt_patch_sizes,
                                       options_.klt_max_iter, float(options_.klt_min_update_squared), px_ref, px_cur, st);
      for (size_t k = 0; k < sel.size(); ++k) { cur_px[sel[k]] = px_cur[k]; success[sel[k]] = st[k]; }
    }
    // bookkeeping in track order (:67-115)
    cur_frame->clearFeatureStorage();
    FeatureTracks kept;
    kept.reserve(n);
    for (size_t t = 0; t < n; ++t) {
      FeatureTrack& track = tracks[t];
      if (success[t]) {
        const FeatureRef& ref = options_.klt_template_is_first_observation ? track.front() : track.back();
        const size_t slot = cur_frame->px_vec_.size();
        cur_frame->px_vec_.push_back(cur_px[t]);
        cur_frame->score_vec_.push_back(ref.getFrame()->score_vec_.at(ref.getFeatureIndex()));
        cur_frame->track_id_vec_.push_back(track.getTrackId());
        // resizeFeatureStorage's initial values for the other columns (frame.cpp:94-123)
        cur_frame->grad_vec_.push_back({0.0, 0.0});
        cur_frame->level_vec_.push_back(0);
        cur_frame->type_vec_.push_back(FeatureType::kCorner);
        cur_frame->depth_vec_.push_back(-1.0);
        cur_frame->invmu_sigma2_a_b_vec_.push_back({0, 0, 0, 0});
        cur_frame->f_vec_.push_back(normalizedBearing(cur_frame->cam_->model, cur_px[t]));  // computeNormalizedBearingVectors (:119-121)
        track.pushBack(nframe_kp1, frame_index, slot);
        kept.push_back(track);
      } else {
        terminated_tracks_.at(frame_index).push_back(track);
      }
    }
    tracks.swap(kept);
    cur_frame->num_features_ = cur_frame->px_vec_.size();
  }
  return getTotalActiveTracks();
}

size_t FeatureTracker::initializeNewTracks(const FrameBundle::Ptr& nframe) {
  for (size_t frame_index = 0; frame_index < bundle_size_; ++frame_index) {
    const FramePtr& frame = nframe->frames_.at(frame_index);
    AbstractDetector& det = *detectors_.at(frame_index);
    det.resetGrid();
    for (const Keypoint& px : frame->px_vec_) det.grid_.occupancy_.at(det.grid_.getCellIndex(int(px[0]), int(px[1]), 1)) = 1;  // fillWithKeypoints
    Keypoints new_px; Scores new_scores; Levels new_levels; Gradients new_grads; FeatureTypes new_types;
    det.detect(b200::ensureGpu(*frame), det.grid_.size(), new_px, new_scores, new_levels, new_grads, new_types);
    const size_t n_old = frame->num_features_;
    FeatureTracks& tracks = active_tracks_.at(frame_index);
    for (size_t i = 0; i < new_px.size(); ++i) {
      frame->px_vec_.push_back(new_px[i]);
      frame->f_vec_.push_back(normalizedBearing(frame->cam_->model, new_px[i]));
      frame->grad_vec_.push_back(new_grads[i]);
      frame->score_vec_.push_back(new_scores[i]);
      frame->level_vec_.push_back(new_levels[i]);
      frame->type_vec_.push_back(FeatureType::kCorner);  // the reference leaves type_vec_ at its initial value ("TODO(cfo)", :169)
      frame->depth_vec_.push_back(-1.0);
      frame->invmu_sigma2_a_b_vec_.push_back({0, 0, 0, 0});
      const int new_track_id = PointIdProvider::getNewPointId();
      tracks.emplace_back(new_track_id);
      tracks.back().pushBack(nframe, frame_index, n_old + i);
      frame->track_id_vec_.resize(n_old + i, -1);
      frame->track_id_vec_.push_back(new_track_id);
    }
    frame->num_features_ = frame->px_vec_.size();
  }
  return getTotalActiveTracks();
}

size_t FeatureTracker::getTotalActiveTracks() const {
  size_t n = 0;
  for (const FeatureTracks& t : active_tracks_) n += t.size();
  return n;
}

void FeatureTracker::getNumTrackedAndDisparityPerFrame(double pivot_ratio, std::vector<size_t>* num_tracked, std::vector<double>* disparity) const {
  num_tracked->resize(bundle_size_);
  disparity->resize(bundle_size_);
  for (size_t i = 0; i < bundle_size_; ++i) {
    num_tracked->at(i) = active_tracks_[i].size();
    disparity->at(i) = feature_tracking_utils::getTracksDisparityPercentile(active_tracks_[i], pivot_ratio);
  }
}

void FeatureTracker::reset() {
  resetActiveTracks();
  resetTerminatedTracks();
  for (auto& d : detectors_) d->resetGrid();
}

// ---- StereoTriangulation ------------------------------------------------------------------------------------------------------------
// libstdc++'s std::random_shuffle(first, last) (bits/stl_algo.h; removed from C++17 but what the reference's build runs):
// for i = 1 .. n-1: swap(a[i], a[std::rand() % (i + 1)]).
template <class It>
static void randomShuffleLikeReference(It first, It last) {
  if (first == last) return;
  for (It i = first + 1; i != last; ++i) {
    It j = first + std::rand() % ((i - first) + 1);
    if (i != j) std::iter_swap(i, j);
  }
}

void StereoTriangulation::compute(const FramePtr& frame0, const FramePtr& frame1) {
  if (frame0->numLandmarks() >= options_.triangulate_n_features) return;  // :27-32
  // detect new features (:34-47), bearing vectors (:49-51), append to frame0 (:53-66)
  Keypoints new_px; Scores new_scores; Levels new_levels; Gradients new_grads; FeatureTypes new_types;
  const size_t max_n_features = feature_detector_->grid_.size();
  feature_detector_->detect(b200::ensureGpu(*frame0), max_n_features, new_px, new_scores, new_levels, new_grads, new_types);
  if (new_px.empty()) return;
  const size_t n_old = frame0->numFeatures(), n_new = new_px.size();
  frame0->landmark_vec_.resize(n_old, nullptr);
  frame0->seed_ref_vec_.resize(n_old);
  for (size_t i = 0; i < n_new; ++i) {
    frame0->px_vec_.push_back(new_px[i]);
    frame0->f_vec_.push_back(normalizedBearing(frame0->cam_->model, new_px[i]));
    frame0->grad_vec_.push_back(new_grads[i]);
    frame0->score_vec_.push_back(new_scores[i]);
    frame0->level_vec_.push_back(new_levels[i]);
    frame0->type_vec_.push_back(new_types[i]);
    frame0->depth_vec_.push_back(-1.0);
    frame0->invmu_sigma2_a_b_vec_.push_back({0, 0, 0, 0});
    frame0->landmark_vec_.push_back(nullptr);
    frame0->seed_ref_vec_.push_back(SeedRef());
  }
  frame0->num_features_ += n_new;
  // visiting order (:68-79)
  std::vector<size_t> indices(n_new);
  std::iota(indices.begin(), indices.end(), n_old);
  const long n_corners = std::count_if(new_types.begin(), new_types.end(), [](const FeatureType& t) { return t == FeatureType::kCorner; });
  randomShuffleLikeReference(indices.begin(), indices.begin() + n_corners);
  randomShuffleLikeReference(indices.begin() + n_corners, indices.end());
  const size_t n_desired = options_.triangulate_n_features - frame0->numLandmarks();
  // the matching loop (:87-133) in one device call
  std::vector<svo_feature> ftrs(n_new);
  for (size_t k = 0; k < n_new; ++k) {
    const size_t i = indices[k];
    svo_feature q{};
    q.px[0] = frame0->px_vec_[i][0]; q.px[1] = frame0->px_vec_[i][1];
    for (int a = 0; a < 3; ++a) q.f[a] = frame0->f_vec_[i][a];
    q.grad[0] = frame0->grad_vec_[i][0]; q.grad[1] = frame0->grad_vec_[i][1];
    q.type = int(frame0->type_vec_[i]); q.level = frame0->level_vec_[i];
    ftrs[k] = q;
  }
  Matcher matcher;
  matcher.options_.max_epi_search_steps = 500;
  matcher.options_.subpix_refinement = true;
  const svo_matcher_options mo = matcher.cOptions();
  double T_f1f0[7], T_w_c0[7];
  (frame1->T_cam_imu_ * frame0->T_cam_imu_.inverse()).toArray(T_f1f0);  // frame1->T_cam_body_ * frame0->T_body_cam_ (:92)
  frame0->T_f_w_.inverse().toArray(T_w_c0);
  const int begin[2] = {0, int(n_new)}, want = int(n_desired), first_slot = int(frame1->numFeatures()), zero = 0;
  std::vector<svo_stereo_result> res(n_new);
  svo_stereo_stats stats{};
  b200::check(svo_cuda_stereo_triangulate(b200::context(), b200::ensureGpu(*frame0).handle(), b200::ensureGpu(*frame1).handle(), &zero, &zero,
                                          &frame0->cam_->model, &frame1->cam_->model, T_f1f0, T_w_c0, 1, begin, int(n_new), ftrs.data(), &want,
                                          &first_slot, options_.mean_depth_inv, options_.min_depth_inv, options_.max_depth_inv, &mo,
                                          res.data(), &stats, SVO_MEM_HOST), "svo_cuda_stereo_triangulate");
  // bookkeeping of both frames (:102-129), in visiting order = slot order
  frame1->landmark_vec_.resize(frame1->numFeatures(), nullptr);
  frame1->seed_ref_vec_.resize(frame1->numFeatures());
  for (size_t k = 0; k < n_new; ++k) {
    const svo_stereo_result& r = res[k];
    if (r.status != SVO_STEREO_SUCCESS) continue;
    const size_t i_ref = indices[k];
    auto new_point = std::make_shared<Point>();
    new_point->pos_ = {r.xyz_world[0], r.xyz_world[1], r.xyz_world[2]};
    frame0->landmark_vec_[i_ref] = new_point;
    new_point->obs_.emplace_back(frame0, i_ref);
    const size_t i_cur = frame1->num_features_;
    frame1->type_vec_.push_back(FeatureType(r.type));
    frame1->level_vec_.push_back(r.level);
    frame1->px_vec_.push_back({r.px_cur[0], r.px_cur[1]});
    frame1->f_vec_.push_back({r.f_cur[0], r.f_cur[1], r.f_cur[2]});
    frame1->score_vec_.push_back(frame0->score_vec_[i_ref]);
    frame1->grad_vec_.push_back({r.grad_cur[0], r.grad_cur[1]});
    frame1->depth_vec_.push_back(-1.0);
    frame1->invmu_sigma2_a_b_vec_.push_back({0, 0, 0, 0});
    frame1->landmark_vec_.push_back(new_point);
    frame1->seed_ref_vec_.push_back(SeedRef());
    new_point->obs_.emplace_back(frame1, i_cur);
    frame1->num_features_++;
  }
}

}  // namespace svo

// ---- fast:: leaves (fast.h:20-41) -----------------------------------------------------------------------------------------------------
namespace fast {
namespace {
using svo::b200::check;
using svo::b200::context;
struct OneLevel {  // a one-level device pyramid holding the caller's image
  svo_cuda_pyr* pyr = nullptr;
  OneLevel(const fast_byte* img, int w, int h, int stride) {
    check(svo_cuda_pyr_create(context(), 1, w, h, 1, -1, &pyr), "svo_cuda_pyr_create");
    const int rc = svo_cuda_pyr_upload(context(), pyr, 0, 1, img, size_t(stride), size_t(stride) * h, SVO_MEM_HOST);
    if (rc == SVO_OK) svo_cuda_ctx_synchronize(context());
    if (rc != SVO_OK) { svo_cuda_pyr_destroy(context(), pyr); pyr = nullptr; check(rc, "svo_cuda_pyr_upload"); }
  }
  ~OneLevel() { if (pyr) svo_cuda_pyr_destroy(context(), pyr); }
};
void detect(const fast_byte* img, int w, int h, int stride, short barrier, int arc, std::vector<fast_xy>& corners) {
  if (w < 7 || h < 7) return;  // the reference's loops over [3, w-3) x [3, h-3) are empty
  OneLevel lvl(img, w, h, stride);
  std::vector<svo_fast_xy> xy(std::max(1024, w * h / 16));
  int n = 0;
  for (;;) {
    check(svo_cuda_fast_corner_list(context(), lvl.pyr, 0, 0, barrier, arc, int(xy.size()), xy.data(), nullptr, nullptr, &n, SVO_MEM_HOST),
          "svo_cuda_fast_corner_list");
    if (n <= int(xy.size())) break;
    xy.resize(size_t(n));
  }
  corners.reserve(corners.size() + size_t(n));
  for (int i = 0; i < n; ++i) corners.push_back(fast_xy(xy[i].x, xy[i].y));  // appended, as the reference's push_back
}
}  // namespace
void fast_corner_detect_9(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners) { detect(img, w, h, stride, barrier, 9, corners); }
void fast_corner_detect_9_sse2(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners) { detect(img, w, h, stride, barrier, 9, corners); }
void fast_corner_detect_10(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners) { detect(img, w, h, stride, barrier, 10, corners); }
void fast_corner_detect_10_sse2(const fast_byte* img, int w, int h, int stride, short barrier, std::vector<fast_xy>& corners) { detect(img, w, h, stride, barrier, 10, corners); }

void fast_corner_score_10(const fast_byte* img, const int img_stride, const std::vector<fast_xy>& corners, const int threshold, std::vector<int>& scores) {
  scores.resize(corners.size());
  if (corners.empty()) return;
  int max_x = 0, max_y = 0;
  for (const fast_xy& c : corners) { max_x = std::max<int>(max_x, c.x); max_y = std::max<int>(max_y, c.y); }
  const int w = std::min(max_x + 4, img_stride), h = max_y + 4;  // every pixel fast_10_score.cpp:3158-3177 reads
  OneLevel lvl(img, w, h, img_stride);
  static_assert(sizeof(fast_xy) == sizeof(svo_fast_xy), "fast_xy layout");
  check(svo_cuda_fast_corner_score(context(), lvl.pyr, 0, 0, int(corners.size()), reinterpret_cast<const svo_fast_xy*>(corners.data()), threshold, 10,
                                   scores.data(), SVO_MEM_HOST),
        "svo_cuda_fast_corner_score");
}

void fast_nonmax_3x3(const std::vector<fast_xy>& corners, const std::vector<int>& scores, std::vector<int>& nonmax_corners) {
  nonmax_corners.clear();
  if (corners.empty()) return;
  nonmax_corners.resize(corners.size());
  int n = 0;
  check(svo_cuda_fast_nonmax_3x3(context(), int(corners.size()), reinterpret_cast<const svo_fast_xy*>(corners.data()), scores.data(),
                                 nonmax_corners.data(), &n, SVO_MEM_HOST),
        "svo_cuda_fast_nonmax_3x3");
  nonmax_corners.resize(size_t(n));
}
}  // namespace fast
