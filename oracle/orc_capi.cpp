// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
// extern "C" surface of the CPU oracle; see orc_capi.h.
#include "orc_capi.h"
#include <thread>
#include <atomic>
#include <vector>
#include <cstring>
#include "orc_math.hpp"
#include "orc_detect.hpp"
#include "orc_sparse_align.hpp"
#include "orc_matcher.hpp"
#include "orc_depth_filter.hpp"
#include "orc_reprojector.hpp"
#include "orc_pose_optimizer.hpp"
#include "orc_point.hpp"

using namespace orc;

namespace {

Camera camOf(const orc_frame* f) {
  Camera c;
  c.fx = f->cam[0]; c.fy = f->cam[1]; c.cx = f->cam[2]; c.cy = f->cam[3];
  c.k1 = f->cam[4]; c.k2 = f->cam[5]; c.p1 = f->cam[6]; c.p2 = f->cam[7];
  c.width = f->width; c.height = f->height; c.distortion = f->distortion;
  return c;
}
std::vector<Img> pyrOf(const orc_frame* f) {
  std::vector<Img> v(f->n_levels);
  for (int i = 0; i < f->n_levels; ++i) v[i] = Img{f->level_data[i], f->level_cols[i], f->level_rows[i], f->level_step[i]};
  return v;
}
AlignFrame alignFrameOf(const orc_frame* f) {
  AlignFrame a;
  a.img_pyr = pyrOf(f);
  a.cam = camOf(f);
  a.T_cam_imu = se3FromArray(f->T_cam_imu);
  a.T_imu_world = se3FromArray(f->T_imu_world);
  a.px.resize(f->n_features); a.f.resize(f->n_features); a.depth.resize(f->n_features); a.eligible.resize(f->n_features);
  for (int i = 0; i < f->n_features; ++i) {
    a.px[i] = {f->px[2 * i], f->px[2 * i + 1]};
    a.f[i] = {f->f[3 * i], f->f[3 * i + 1], f->f[3 * i + 2]};
    a.depth[i] = f->depth[i];
    a.eligible[i] = f->eligible ? f->eligible[i] : 1;
  }
  return a;
}
MatchFrame matchFrameOf(const orc_frame* f) {
  MatchFrame m;
  m.img_pyr = pyrOf(f);
  m.cam = camOf(f);
  return m;
}
FeatureRef featureOf(const orc_feature* f) {
  FeatureRef r;
  r.type = static_cast<FeatureType>(f->type);
  r.px = {f->px[0], f->px[1]};
  r.f = {f->f[0], f->f[1], f->f[2]};
  r.grad = {f->grad[0], f->grad[1]};
  r.level = f->level;
  return r;
}
void setMatcherOptions(Matcher& m, const orc_matcher_options* o) {
  m.options_.align_1d = o->align_1d;
  m.options_.align_max_iter = o->align_max_iter;
  m.options_.max_epi_search_steps = o->max_epi_search_steps;
  m.options_.subpix_refinement = o->subpix_refinement;
  m.options_.epi_search_edgelet_filtering = o->epi_search_edgelet_filtering;
  m.options_.scan_on_unit_sphere = o->scan_on_unit_sphere;
  m.options_.epi_search_edgelet_max_angle = o->epi_search_edgelet_max_angle;
  m.options_.affine_est_offset_ = o->affine_est_offset;
  m.options_.affine_est_gain_ = o->affine_est_gain;
  m.options_.max_patch_diff_ratio = o->max_patch_diff_ratio;
}
void fillMatchOut(const Matcher& m, Matcher::MatchResult r, double depth, orc_match_out* out) {
  out->result = int(r);
  out->px_cur[0] = m.px_cur_.x; out->px_cur[1] = m.px_cur_.y;
  out->f_cur[0] = m.f_cur_.x; out->f_cur[1] = m.f_cur_.y; out->f_cur[2] = m.f_cur_.z;
  out->search_level = m.search_level_;
  out->A_cur_ref[0] = m.A_cur_ref_[0][0]; out->A_cur_ref[1] = m.A_cur_ref_[0][1];
  out->A_cur_ref[2] = m.A_cur_ref_[1][0]; out->A_cur_ref[3] = m.A_cur_ref_[1][1];
  out->h_inv = m.h_inv_;
  out->epi_length_pyramid = m.epi_length_pyramid_;
  out->reject = m.reject_;
  out->depth = depth;
  std::memcpy(out->patch_with_border, m.patch_with_border_, 100);
  out->epi_image[0] = m.epi_image_.x; out->epi_image[1] = m.epi_image_.y;
}

template <class F>
void parallelFor(int n, int n_threads, F&& fn) {
  if (n_threads <= 1 || n <= 1) { for (int i = 0; i < n; ++i) fn(i); return; }
  std::atomic<int> next{0};
  std::vector<std::thread> th;
  const int nt = std::min(n_threads, n);
  for (int t = 0; t < nt; ++t)
    th.emplace_back([&]() { for (;;) { const int i = next.fetch_add(1); if (i >= n) break; fn(i); } });
  for (auto& t : th) t.join();
}

}  // namespace

extern "C" {

void orc_half_sample(const uint8_t* in, int cols, int rows, int stride, uint8_t* out, int out_stride, int mode) {
  halfSample(in, cols, rows, stride, out, out_stride, mode);
}

size_t orc_create_img_pyramid(const uint8_t* img0, int cols, int rows, int n_levels, uint8_t* out, int mode) {
  Pyramid pyr;
  createImgPyramid(img0, cols, rows, cols, n_levels, pyr, mode);
  size_t off = 0;
  for (int i = 1; i < n_levels; ++i) {
    if (out) std::memcpy(out + off, pyr.store[i].data(), pyr.store[i].size());
    off += pyr.store[i].size();
  }
  return off;
}

int orc_fast_detect(const uint8_t* img, int w, int h, int stride, int barrier, int arc, short* xy, int cap) {
  std::vector<FastXY> c;
  if (arc == 9) fastCornerDetect<9>(img, w, h, stride, barrier, c);
  else fastCornerDetect<10>(img, w, h, stride, barrier, c);
  const int n = std::min<int>(cap, int(c.size()));
  for (int i = 0; i < n; ++i) { xy[2 * i] = c[i].x; xy[2 * i + 1] = c[i].y; }
  return int(c.size());
}

void orc_fast_score10(const uint8_t* img, int stride, const short* xy, int n, int threshold, int* scores) {
  std::vector<FastXY> c(n);
  for (int i = 0; i < n; ++i) c[i] = FastXY{xy[2 * i], xy[2 * i + 1]};
  std::vector<int> s;
  fastCornerScore10(img, stride, c, threshold, s);
  for (int i = 0; i < n; ++i) scores[i] = s[i];
}

int orc_fast_nonmax3x3(const short* xy, const int* scores, int n, int* idx_out) {
  std::vector<FastXY> c(n);
  for (int i = 0; i < n; ++i) c[i] = FastXY{xy[2 * i], xy[2 * i + 1]};
  std::vector<int> s(scores, scores + n), nm;
  fastNonmax3x3(c, s, nm);
  for (size_t i = 0; i < nm.size(); ++i) idx_out[i] = nm[i];
  return int(nm.size());
}

void orc_fast_detector(const uint8_t* img0, int cols, int rows, int n_levels, int pyr_mode, int threshold, int border,
                       int min_level, int max_level, int cell_size, const uint8_t* occupancy, orc_corner* corners_out) {
  Pyramid pyr;
  createImgPyramid(img0, cols, rows, cols, n_levels, pyr, pyr_mode);
  const int n_cols = int(std::ceil(double(cols) / cell_size));
  const int n_rows = int(std::ceil(double(rows) / cell_size));
  std::vector<Corner> corners(size_t(n_cols) * n_rows, Corner{0, 0, 0, float(threshold), 0.0f});
  std::vector<uint8_t> occ(size_t(n_cols) * n_rows, 0);
  if (occupancy) occ.assign(occupancy, occupancy + occ.size());
  fastDetector(pyr.lv, threshold, border, min_level, max_level, corners, occ, cell_size, n_cols);
  for (size_t i = 0; i < corners.size(); ++i)
    corners_out[i] = orc_corner{corners[i].x, corners[i].y, corners[i].level, corners[i].score, corners[i].angle};
}

int orc_fast_detect_features(const uint8_t* img0, int cols, int rows, int n_levels, int pyr_mode, double threshold, int border,
                             int min_level, int max_level, int cell_size, const uint8_t* occupancy, int max_n,
                             double* px_out, double* score_out, int* level_out) {
  // ref: src/svo_direct/src/feature_detection.cpp:53-74 (FastDetector::detect)
  Pyramid pyr;
  createImgPyramid(img0, cols, rows, cols, n_levels, pyr, pyr_mode);
  const int n_cols = int(std::ceil(double(cols) / cell_size));
  const int n_rows = int(std::ceil(double(rows) / cell_size));
  std::vector<Corner> corners(size_t(n_cols) * n_rows, Corner{0, 0, 0, float(threshold), 0.0f});
  std::vector<uint8_t> occ(size_t(n_cols) * n_rows, 0);
  if (occupancy) occ.assign(occupancy, occupancy + occ.size());
  fastDetector(pyr.lv, int(threshold), border, min_level, max_level, corners, occ, cell_size, n_cols);
  FeatureSoA out;
  fillFeatures(corners, nullptr, 0, threshold, size_t(max_n), out, occ, cell_size, n_cols);
  const int n = int(out.score.size());
  for (int i = 0; i < n; ++i) {
    px_out[2 * i] = out.px[2 * i]; px_out[2 * i + 1] = out.px[2 * i + 1];
    score_out[i] = out.score[i];
    level_out[i] = out.level[i];
  }
  return n;
}

static std::vector<Img> pyrOf(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step) {
  std::vector<Img> pyr;
  for (int l = 0; l < n_levels; ++l) pyr.push_back(Img{data[l], cols[l], rows[l], step[l]});
  return pyr;
}

void orc_gaussian_blur3x3(const uint8_t* img, int cols, int rows, int step, uint8_t* out) {
  std::vector<uint8_t> o;
  gaussianBlur3x3(Img{img, cols, rows, step}, o);
  std::memcpy(out, o.data(), o.size());
}

void orc_scharr3x3(const uint8_t* img, int cols, int rows, int step, int16_t* dx, int16_t* dy) {
  std::vector<uint8_t> tight(size_t(cols) * rows);
  for (int y = 0; y < rows; ++y) std::memcpy(&tight[size_t(y) * cols], img + size_t(y) * step, cols);
  std::vector<int16_t> a, b;
  scharr3x3(tight.data(), cols, rows, a, b);
  std::memcpy(dx, a.data(), a.size() * 2);
  std::memcpy(dy, b.data(), b.size() * 2);
}

void orc_edgelet_detector_v2(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step,
                             int threshold, int border, int cell_size, const uint8_t* occupancy, orc_corner* corners_out) {
  const int n_cols = int(std::ceil(double(cols[0]) / cell_size));
  const int n_rows = int(std::ceil(double(rows[0]) / cell_size));
  std::vector<Corner> corners(size_t(n_cols) * n_rows, Corner{0, 0, 0, float(threshold), 0.0f});
  std::vector<uint8_t> occ(corners.size(), 0);
  if (occupancy) occ.assign(occupancy, occupancy + occ.size());
  edgeletDetectorV2(pyrOf(n_levels, data, cols, rows, step), threshold, border, corners, occ, cell_size, n_cols);
  for (size_t i = 0; i < corners.size(); ++i)
    corners_out[i] = orc_corner{corners[i].x, corners[i].y, corners[i].level, corners[i].score, corners[i].angle};
}

double orc_angle_at_pixel_histogram(const uint8_t* img, int cols, int rows, int step, int x, int y, int halfpatch_size) {
  return angleAtPixelUsingHistogram(Img{img, cols, rows, step}, x, y, halfpatch_size);
}

int orc_angle_histogram_bin(int gx, int gy) { return angleHistogramBin(gx, gy); }

void orc_fast_detector_pyr(int n_levels, const uint8_t* const* data, const int* cols, const int* rows, const int* step, int threshold,
                           int border, int min_level, int max_level, int cell_size, const uint8_t* occupancy, orc_corner* corners_out) {
  const int n_cols = int(std::ceil(double(cols[0]) / cell_size));
  const int n_rows = int(std::ceil(double(rows[0]) / cell_size));
  std::vector<Corner> corners(size_t(n_cols) * n_rows, Corner{0, 0, 0, float(threshold), 0.0f});
  std::vector<uint8_t> occ(corners.size(), 0);
  if (occupancy) occ.assign(occupancy, occupancy + occ.size());
  fastDetector(pyrOf(n_levels, data, cols, rows, step), threshold, border, min_level, max_level, corners, occ, cell_size, n_cols);
  for (size_t i = 0; i < corners.size(); ++i)
    corners_out[i] = orc_corner{corners[i].x, corners[i].y, corners[i].level, corners[i].score, corners[i].angle};
}

int orc_detect_features(int detector_type, int n_levels, const uint8_t* const* data, const int* cols, const int* rows,
                        const int* step, double threshold_primary, double threshold_secondary, int border, int min_level,
                        int max_level, int cell_size, const uint8_t* occupancy, int max_n, double* px_out, double* score_out,
                        int* level_out, double* grad_out, int* type_out) {
  const int n_cols = int(std::ceil(double(cols[0]) / cell_size));
  const int n_rows = int(std::ceil(double(rows[0]) / cell_size));
  std::vector<uint8_t> occ(size_t(n_cols) * n_rows, 0);
  if (occupancy) occ.assign(occupancy, occupancy + occ.size());
  DetectedFeatures out;
  detectFeatures(detector_type, pyrOf(n_levels, data, cols, rows, step), threshold_primary, threshold_secondary, border, min_level,
                 max_level, cell_size, occ, size_t(max_n), out);
  const int n = int(out.score.size());
  for (int i = 0; i < n && i < max_n; ++i) {
    px_out[2 * i] = out.px[2 * i]; px_out[2 * i + 1] = out.px[2 * i + 1];
    grad_out[2 * i] = out.grad[2 * i]; grad_out[2 * i + 1] = out.grad[2 * i + 1];
    score_out[i] = out.score[i];
    level_out[i] = out.level[i];
    type_out[i] = out.type[i];
  }
  return n;
}

int orc_sparse_align(int n_cams, const orc_frame* ref, const orc_frame* cur, const orc_align_options* opt, orc_align_result* res) {
  std::vector<AlignFrame> rf, cf;
  for (int i = 0; i < n_cams; ++i) { rf.push_back(alignFrameOf(&ref[i])); cf.push_back(alignFrameOf(&cur[i])); }
  SolverOptions so;
  so.max_iter = opt->max_iter;
  so.eps = opt->eps;
  SparseImgAlignOptions o;
  o.max_level = opt->max_level; o.min_level = opt->min_level;
  o.estimate_illumination_gain = opt->estimate_illumination_gain;
  o.estimate_illumination_offset = opt->estimate_illumination_offset;
  o.use_distortion_jacobian = opt->use_distortion_jacobian;
  o.robustification = opt->robustification;
  o.weight_scale = opt->weight_scale;
  SparseImgAlign sia(so, o);
  sia.reset();  // callers always reset() first (src/svo/src/frame_handler_base.cpp:621)
  if (opt->have_prior)
    sia.setWeightedPrior(se3FromArray(opt->prior_T), opt->prior_alpha, opt->prior_beta, opt->lambda_rot, opt->lambda_trans,
                         opt->lambda_alpha, opt->lambda_beta);
  sia.alpha_init_ = opt->alpha_init;
  sia.beta_init_ = opt->beta_init;
  const SparseImgAlignResult r = sia.run(rf, cf);
  std::memset(res, 0, sizeof(*res));
  res->n_tracked = int(r.n_fts_to_track);
  se3ToArray(r.T_icur_iref, res->T_icur_iref);
  res->alpha = r.alpha; res->beta = r.beta; res->chi2 = r.chi2;
  if (r.n_fts_to_track) for (int a = 0; a < 8; ++a) for (int b = 0; b < 8; ++b) res->H[8 * a + b] = r.H[a][b];
  for (size_t i = 0; i < r.iters_per_level.size() && i < ORC_MAX_LEVELS; ++i) res->iters[i] = r.iters_per_level[i];
  for (size_t i = 0; i < r.T_f_w.size() && i < ORC_MAX_CAMS; ++i) se3ToArray(r.T_f_w[i], res->T_f_w[i]);
  res->stop = r.stop;
  return res->n_tracked;
}

int orc_sparse_align_batch(int B, int n_cams, const orc_frame* ref, const orc_frame* cur, const orc_align_options* opt,
                           orc_align_result* res, int n_threads) {
  parallelFor(B, n_threads, [&](int i) { orc_sparse_align(n_cams, ref + size_t(i) * n_cams, cur + size_t(i) * n_cams, opt, res + i); });
  return 0;
}

int orc_warp_affine(const double A[4], const uint8_t* img, int cols, int rows, int step, const double px_ref[2],
                    int level_ref, int search_level, int halfpatch_size, uint8_t* patch) {
  const double Am[2][2] = {{A[0], A[1]}, {A[2], A[3]}};
  return warpAffine(Am, Img{img, cols, rows, step}, V2{px_ref[0], px_ref[1]}, level_ref, search_level, halfpatch_size, patch) ? 1 : 0;
}

void orc_get_warp_matrix_affine(const orc_frame* ref, const orc_frame* cur, const double px_ref[2], const double f_ref[3],
                                double depth_ref, const double T_cur_ref[7], int level_ref, double A_out[4]) {
  double A[2][2];
  getWarpMatrixAffine(camOf(ref), camOf(cur), V2{px_ref[0], px_ref[1]}, V3{f_ref[0], f_ref[1], f_ref[2]}, depth_ref,
                      se3FromArray(T_cur_ref), level_ref, A);
  A_out[0] = A[0][0]; A_out[1] = A[0][1]; A_out[2] = A[1][0]; A_out[3] = A[1][1];
}

int orc_get_best_search_level(const double A[4], int max_level) {
  const double Am[2][2] = {{A[0], A[1]}, {A[2], A[3]}};
  return getBestSearchLevel(Am, max_level);
}

int orc_zmssd(const uint8_t* ref_patch64, const uint8_t* cur, int stride) {
  return ZMSSD(ref_patch64).computeScore(cur, stride);
}

int orc_align2d(const uint8_t* img, int cols, int rows, int step, const uint8_t* pwb, int n_iter, int est_offset, int est_gain,
                double px[2]) {
  uint8_t patch[64];
  createPatchFromPatchWithBorder(pwb, 8, patch);
  V2 p{px[0], px[1]};
  const bool r = align2D(Img{img, cols, rows, step}, pwb, patch, n_iter, est_offset != 0, est_gain != 0, p);
  px[0] = p.x; px[1] = p.y;
  return r ? 1 : 0;
}

int orc_align1d(const uint8_t* img, int cols, int rows, int step, const double dir[2], const uint8_t* pwb, int n_iter,
                int est_offset, int est_gain, double px[2], double* h_inv) {
  uint8_t patch[64];
  createPatchFromPatchWithBorder(pwb, 8, patch);
  V2 p{px[0], px[1]};
  const bool r = align1D(Img{img, cols, rows, step}, V2{dir[0], dir[1]}, pwb, patch, n_iter, est_offset != 0, est_gain != 0, &p, h_inv);
  px[0] = p.x; px[1] = p.y;
  return r ? 1 : 0;
}

int orc_find_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                          double ref_depth, const double px_cur_in[2], const orc_matcher_options* opt, orc_match_out* out) {
  const MatchFrame rf = matchFrameOf(ref), cf = matchFrameOf(cur);
  Matcher m;
  setMatcherOptions(m, opt);
  V2 px{px_cur_in[0], px_cur_in[1]};
  m.px_cur_ = px;
  const Matcher::MatchResult r = m.findMatchDirect(rf, cf, se3FromArray(T_cur_ref), featureOf(ftr), ref_depth, px);
  fillMatchOut(m, r, 0.0, out);
  return int(r);
}

int orc_find_epipolar_match_direct(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], const orc_feature* ftr,
                                   double d_estimate_inv, double d_min_inv, double d_max_inv, const orc_matcher_options* opt,
                                   orc_match_out* out) {
  const MatchFrame rf = matchFrameOf(ref), cf = matchFrameOf(cur);
  Matcher m;
  setMatcherOptions(m, opt);
  double depth = 0.0;
  const Matcher::MatchResult r = m.findEpipolarMatchDirect(rf, cf, se3FromArray(T_cur_ref), featureOf(ftr), d_estimate_inv,
                                                           d_min_inv, d_max_inv, depth);
  fillMatchOut(m, r, depth, out);
  return int(r);
}

void orc_scan_epipolar_line(const orc_frame* cur, const double A[3], const double B[3], const double C[3], const uint8_t* patch64,
                            int patch_level, double epi_length_pyramid, const orc_matcher_options* opt, double image_best[2],
                            int* zmssd_best) {
  const MatchFrame cf = matchFrameOf(cur);
  Matcher m;
  setMatcherOptions(m, opt);
  m.epi_length_pyramid_ = epi_length_pyramid;
  const ZMSSD patch_score(patch64);
  V2 best{0, 0};
  const V3 a{A[0], A[1], A[2]}, b{B[0], B[1], B[2]}, c{C[0], C[1], C[2]};
  if (m.options_.scan_on_unit_sphere) m.scanEpipolarUnitSphere(cf, a, b, c, patch_score, patch_level, &best, zmssd_best);  // matcher.cpp:335-338
  else m.scanEpipolarUnitPlane(cf, a, b, c, patch_score, patch_level, &best, zmssd_best);
  image_best[0] = best.x; image_best[1] = best.y;
}

// StereoTriangulation::compute from the matching loop on — ref: src/svo/src/stereo_triangulation.cpp:87-137.
// (Detection, bearing vectors and the two std::random_shuffle calls, :34-79, happen before: `ftrs` is the shuffled visiting order.)
int orc_stereo_triangulate(const orc_frame* frame0, const orc_frame* frame1, int n, const orc_feature* ftrs, int n_desired,
                           int n_features_in_frame1, double mean_depth_inv, double min_depth_inv, double max_depth_inv,
                           orc_stereo_result* results, int* n_failed_out) {
  const MatchFrame f0 = matchFrameOf(frame0), f1 = matchFrameOf(frame1);
  const SE3 T_c0_w = se3FromArray(frame0->T_cam_imu) * se3FromArray(frame0->T_imu_world);
  const SE3 T_c1_w = se3FromArray(frame1->T_cam_imu) * se3FromArray(frame1->T_imu_world);
  const SE3 T_f1f0 = se3FromArray(frame1->T_cam_imu) * inverse(se3FromArray(frame0->T_cam_imu));  // T_cam_body(1) * T_body_cam(0), :92
  const SE3 T_w_c0 = inverse(T_c0_w);
  (void)T_c1_w;
  Matcher matcher;
  matcher.options_.max_epi_search_steps = 500;  // :90-91
  matcher.options_.subpix_refinement = true;
  int n_succeeded = 0, n_failed = 0;
  for (int i = 0; i < n; ++i) { results[i] = orc_stereo_result{}; results[i].match_result = -1; results[i].slot = -1; }
  for (int i = 0; i < n; ++i) {
    const FeatureRef ft = featureOf(&ftrs[i]);
    matcher.options_.align_1d = isEdgelet(ft.type);  // :95
    double depth = 0.0;
    const Matcher::MatchResult res = matcher.findEpipolarMatchDirect(f0, f1, T_f1f0, ft, mean_depth_inv, min_depth_inv, max_depth_inv, depth);
    orc_stereo_result& r = results[i];
    r.match_result = int(res);
    if (res == Matcher::MatchResult::kSuccess) {
      const V3 xyz = T_w_c0 * (ft.f * depth);  // :104-105
      r.status = 2;
      r.slot = n_features_in_frame1 + n_succeeded;  // frame1->num_features_ at that moment, :111
      r.depth = depth;
      r.xyz_world[0] = xyz.x; r.xyz_world[1] = xyz.y; r.xyz_world[2] = xyz.z;
      r.px_cur[0] = matcher.px_cur_.x; r.px_cur[1] = matcher.px_cur_.y;
      r.f_cur[0] = matcher.f_cur_.x; r.f_cur[1] = matcher.f_cur_.y; r.f_cur[2] = matcher.f_cur_.z;
      const double gx = matcher.A_cur_ref_[0][0] * ft.grad.x + matcher.A_cur_ref_[0][1] * ft.grad.y;  // :118-119
      const double gy = matcher.A_cur_ref_[1][0] * ft.grad.x + matcher.A_cur_ref_[1][1] * ft.grad.y;
      const double n2 = gx * gx + gy * gy;
      const double nn = n2 > 0.0 ? std::sqrt(n2) : 1.0;  // Eigen's normalized() leaves a zero vector alone
      r.grad_cur[0] = gx / nn; r.grad_cur[1] = gy / nn;
      r.level = ft.level; r.type = int(ft.type);
      ++n_succeeded;
    } else {
      r.status = 1;
      ++n_failed;
    }
    if (n_succeeded >= n_desired) break;  // :131-132
  }
  if (n_failed_out) *n_failed_out = n_failed;
  return n_succeeded;
}

int orc_point_optimize(int n_obs, const double* T_f_w, const double* f, double pos[3], int n_iter, int using_bearing_vector) {
  std::vector<PointObservation> obs;
  for (int i = 0; i < n_obs; ++i) obs.push_back(PointObservation{se3FromArray(T_f_w + 7 * i), V3{f[3 * i], f[3 * i + 1], f[3 * i + 2]}});
  V3 p{pos[0], pos[1], pos[2]};
  const int it = pointOptimize(obs, p, size_t(n_iter), using_bearing_vector != 0);
  pos[0] = p.x; pos[1] = p.y; pos[2] = p.z;
  return it;
}

int orc_find_match_direct_batch(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], int M,
                                const orc_feature* ftrs, const double* ref_depth, const double* px_cur_in,
                                const orc_matcher_options* opt, orc_match_out* out, int n_threads) {
  const MatchFrame rf = matchFrameOf(ref), cf = matchFrameOf(cur);
  const SE3 T = se3FromArray(T_cur_ref);
  parallelFor(M, n_threads, [&](int i) {
    Matcher m;
    setMatcherOptions(m, opt);
    V2 px{px_cur_in[2 * i], px_cur_in[2 * i + 1]};
    m.px_cur_ = px;
    const Matcher::MatchResult r = m.findMatchDirect(rf, cf, T, featureOf(&ftrs[i]), ref_depth[i], px);
    fillMatchOut(m, r, 0.0, &out[i]);
  });
  return 0;
}

int orc_find_epipolar_match_direct_batch(const orc_frame* ref, const orc_frame* cur, const double T_cur_ref[7], int M,
                                         const orc_feature* ftrs, const double* d_inv3, const orc_matcher_options* opt,
                                         orc_match_out* out, int n_threads) {
  const MatchFrame rf = matchFrameOf(ref), cf = matchFrameOf(cur);
  const SE3 T = se3FromArray(T_cur_ref);
  parallelFor(M, n_threads, [&](int i) {
    Matcher m;
    setMatcherOptions(m, opt);
    double depth = 0.0;
    const Matcher::MatchResult r = m.findEpipolarMatchDirect(rf, cf, T, featureOf(&ftrs[i]), d_inv3[3 * i], d_inv3[3 * i + 1],
                                                             d_inv3[3 * i + 2], depth);
    fillMatchOut(m, r, depth, &out[i]);
  });
  return 0;
}

int orc_update_filter_vogiatzis(double z, double tau2, double mu_range, double state[4]) {
  return updateFilterVogiatzis(z, tau2, mu_range, state) ? 1 : 0;
}
int orc_update_filter_gaussian(double z, double tau2, double state[4]) { return updateFilterGaussian(z, tau2, state) ? 1 : 0; }

void orc_update_filter_vogiatzis_batch(int n, const double* z, const double* tau2, const double* mu_range, double* state,
                                       uint8_t* ok, int n_threads) {
  const int chunk = 4096;
  const int n_chunks = (n + chunk - 1) / chunk;
  parallelFor(n_chunks, n_threads, [&](int c) {
    const int e = std::min(n, (c + 1) * chunk);
    for (int i = c * chunk; i < e; ++i) {
      const bool r = updateFilterVogiatzis(z[i], tau2[i], mu_range[i], state + 4 * size_t(i));
      if (ok) ok[i] = r;
    }
  });
}

double orc_compute_tau(const double T_ref_cur[7], const double f[3], double z, double px_error_angle) {
  return computeTau(se3FromArray(T_ref_cur), V3{f[0], f[1], f[2]}, z, px_error_angle);
}

double orc_px_error_angle(const orc_frame* frame, double px_noise) { return camOf(frame).getAngleError(px_noise); }

int orc_update_seeds(const orc_frame* ref, int n_obs, const orc_frame* cur_frames, const double* T_cur_ref, int S,
                     const orc_feature* ftrs, uint8_t* types, double* states, double seed_mu_range,
                     const orc_matcher_options* opt, double sigma2_convergence_threshold,
                     double mappoint_sigma2_convergence_threshold, double px_error_angle,
                     int check_visibility, int check_convergence, int use_vogiatzis, int* match_results, uint8_t* success,
                     int n_threads) {
  const MatchFrame rf = matchFrameOf(ref);
  std::vector<MatchFrame> cfs;
  std::vector<SE3> Ts;
  for (int o = 0; o < n_obs; ++o) { cfs.push_back(matchFrameOf(&cur_frames[o])); Ts.push_back(se3FromArray(T_cur_ref + 7 * o)); }
  std::atomic<int> n_success{0};
  // Seeds are independent; the observations of one seed are applied in order.
  parallelFor(S, n_threads, [&](int s) {
    Matcher m;
    setMatcherOptions(m, opt);
    FeatureType type = static_cast<FeatureType>(types[s]);
    for (int o = 0; o < n_obs; ++o) {
      int mr = -1;
      // DepthFilter::updateSeeds picks the threshold by seed type (src/svo_direct/src/depth_filter.cpp:214-221)
      const double cur_thresh = (type == FeatureType::kMapPointSeed || type == FeatureType::kMapPointSeedConverged)
                                    ? mappoint_sigma2_convergence_threshold : sigma2_convergence_threshold;
      const bool ok = updateSeed(cfs[o], rf, Ts[o], featureOf(&ftrs[s]), type, states + 4 * size_t(s), seed_mu_range, m,
                                 cur_thresh, px_error_angle, check_visibility != 0, check_convergence != 0,
                                 use_vogiatzis != 0, &mr);
      if (match_results) match_results[size_t(o) * S + s] = mr;
      if (success) success[size_t(o) * S + s] = ok;
      if (ok) n_success.fetch_add(1);
    }
    types[s] = static_cast<uint8_t>(type);
  });
  return n_success.load();
}

}  // extern "C"

extern "C" {
void orc_align_pyr2d(const orc_frame* ref, const orc_frame* cur, int max_level, int min_level, const int* patch_sizes, int n_iter,
                     float min_update_squared, int M, const int* px_ref_level_0, double* px_cur, uint8_t* status, int n_threads) {
  std::vector<Img> pr, pc;
  for (int l = 0; l < ref->n_levels; ++l) {
    pr.push_back(Img{ref->level_data[l], ref->level_cols[l], ref->level_rows[l], ref->level_step[l]});
    pc.push_back(Img{cur->level_data[l], cur->level_cols[l], cur->level_rows[l], cur->level_step[l]});
  }
  const std::vector<int> ps(patch_sizes, patch_sizes + ref->n_levels);
  parallelFor(M, n_threads, [&](int i) {
    V2 px{px_cur[2 * i], px_cur[2 * i + 1]};
    status[i] = alignPyr2D(pr, pc, max_level, min_level, ps, n_iter, min_update_squared, px_ref_level_0 + 2 * i, px) ? 1 : 0;
    px_cur[2 * i] = px.x; px_cur[2 * i + 1] = px.y;
  });
}
void orc_tukey_weight(float b, const float* err, int n, float* w) {
  TukeyWeightFunction f(b);
  for (int i = 0; i < n; ++i) w[i] = f.weight(err[i]);
}
void orc_radtan(double k1, double k2, double p1, double p2, int which, double* xy, int n, double* jac_out) {
  Camera c{};
  c.k1 = k1; c.k2 = k2; c.p1 = p1; c.p2 = p2; c.distortion = 1;
  for (int i = 0; i < n; ++i) {
    if (which == 0) { double xd, yd; c.distort(xy[2 * i], xy[2 * i + 1], xd, yd); xy[2 * i] = xd; xy[2 * i + 1] = yd; }
    else if (which == 1) c.undistort(xy[2 * i], xy[2 * i + 1]);
    else {
      double J[2][2];
      c.distJacobian(xy[2 * i], xy[2 * i + 1], J);
      jac_out[4 * i] = J[0][0]; jac_out[4 * i + 1] = J[0][1]; jac_out[4 * i + 2] = J[1][0]; jac_out[4 * i + 3] = J[1][1];
    }
  }
}
void orc_seed_helpers(const double* s, double mu_range, double sigma2_convergence_threshold, double depth, double depth_sigma, double* out) {
  out[0] = seed::getDepth(s);
  out[1] = seed::getInvMinDepth(s);
  out[2] = seed::getInvMaxDepth(s);
  out[3] = seed::isConverged(s, mu_range, sigma2_convergence_threshold) ? 1.0 : 0.0;
  out[4] = seed::getSigma2FromDepthSigma(depth, depth_sigma);
  out[5] = seed::getInitSigma2FromMuRange(mu_range);
}
void orc_grid_cell_index(int cell_size, int n_cols, const int* xy, const int* scale, int n, long long* idx) {
  for (int i = 0; i < n; ++i) idx[i] = (long long)gridCellIndex(xy[2 * i], xy[2 * i + 1], scale[i], cell_size, n_cols);
}
void orc_patch_from_patch_with_border(const uint8_t* patch_with_border, int patch_size, uint8_t* patch) {
  createPatchFromPatchWithBorder(patch_with_border, patch_size, patch);
}
}  // extern "C"

// One "frame pair step" of the hot path as the bench defines it: build the pyramid of the NEW (cur) frame from its level-0
// image (frame_utils::createImgPyramid), then SparseImgAlign::run against the already-built ref frame. B independent pairs,
// n_threads workers. cur[i] supplies camera/poses; its level pointers are ignored and rebuilt from cur_l0[i].
extern "C" int orc_pyramid_align_batch(int B, int n_levels, const uint8_t* const* cur_l0, int cols, int rows, int pyr_mode,
                                       const orc_frame* ref, const orc_frame* cur, const orc_align_options* opt,
                                       orc_align_result* res, int n_threads) {
  parallelFor(B, n_threads, [&](int i) {
    Pyramid pyr;
    createImgPyramid(cur_l0[i], cols, rows, cols, n_levels, pyr, pyr_mode);
    orc_frame cf = cur[i];
    cf.n_levels = n_levels;
    for (int l = 0; l < n_levels; ++l) {
      cf.level_data[l] = pyr.lv[l].data;
      cf.level_cols[l] = pyr.lv[l].cols;
      cf.level_rows[l] = pyr.lv[l].rows;
      cf.level_step[l] = pyr.lv[l].step;
    }
    orc_sparse_align(1, ref + i, &cf, opt, res + i);
  });
  return 0;
}

// f1
extern "C" int orc_reproject_match(const orc_reproj_map* map, const orc_frame* cur, int E, const int* entry_feat, int n_features_in,
                                   uint8_t* occupancy, const orc_reproj_options* opt, orc_reproj_result* results, orc_reproj_stats* stats) {
  std::vector<ReprojFrame> kfs(map->n_kfs);
  for (int k = 0; k < map->n_kfs; ++k) {
    kfs[k].mf = matchFrameOf(&map->kfs[k]);
    kfs[k].T_f_w = se3FromArray(map->kfs[k].T_cam_imu) * se3FromArray(map->kfs[k].T_imu_world);
  }
  ReprojFrame c;
  c.mf = matchFrameOf(cur);
  c.T_f_w = se3FromArray(cur->T_cam_imu) * se3FromArray(cur->T_imu_world);
  reprojectMatch(*map, kfs, c, E, entry_feat, n_features_in, occupancy, *opt, results, stats);
  return stats->n_matches;
}

// f4
extern "C" int orc_pose_optimize(int n_cams, const orc_frame* frames, int N, const orc_feature* ftrs, const int* feat_cam,
                                 const double* xyz_world, const uint8_t* has_xyz, const orc_pose_opt_options* opt,
                                 double T_imu_world_out[7], uint8_t* outlier, double stats[6]) {
  std::vector<PoseOptCam> cams(n_cams);
  for (int c = 0; c < n_cams; ++c) { cams[c].cam = camOf(&frames[c]); cams[c].T_cam_imu = se3FromArray(frames[c].T_cam_imu); }
  std::vector<PoseOptFeature> fts(N);
  for (int i = 0; i < N; ++i) {
    const FeatureRef r = featureOf(&ftrs[i]);
    fts[i].px = r.px; fts[i].f = r.f; fts[i].grad = r.grad; fts[i].level = r.level; fts[i].type = r.type;
    fts[i].xyz_world = {xyz_world[3 * i], xyz_world[3 * i + 1], xyz_world[3 * i + 2]};
    fts[i].has_xyz = has_xyz[i] != 0;
    fts[i].cam = feat_cam[i];
  }
  PoseOptimizer po;
  po.err_type = opt->err_type; po.max_iter = opt->max_iter; po.eps = opt->eps;
  if (opt->have_prior) {
    po.have_prior = true;
    po.prior.q = {opt->prior_q[0], opt->prior_q[1], opt->prior_q[2], opt->prior_q[3]};
    po.prior.t = {0, 0, 0};
    po.prior_lambda = opt->prior_lambda;
  }
  SE3 T = se3FromArray(frames[0].T_imu_world);
  const size_t n = po.run(fts, cams, T, opt->reproj_thresh_px, outlier, stats);
  se3ToArray(T, T_imu_world_out);
  return int(n);
}
