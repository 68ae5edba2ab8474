// ORACLE — TEST INFRASTRUCTURE ONLY.
// Thin C wrapper around the REFERENCE's own direct front-end sources, compiled from where they lie under
// /root/reference into oracle/_ref/libdirect_ref.so by oracle/Makefile:
//   src/vikit/vikit_common/src/vision.cpp            vk::halfSample (+ halfSampleSSE2)
//   src/svo_direct/src/feature_alignment.cpp         svo::feature_alignment::align1D / align2D / alignPyr2D
//   src/svo_direct/include/svo/direct/patch_score.h  svo::patch_score::ZMSSD<4> (header only)
//   src/svo_direct/include/svo/direct/patch_utils.h  createPatchFromPatchWithBorder (header only)
//   src/vikit/vikit_solver/src/robust_cost.cpp       vk::solver::TukeyWeightFunction::weight
//   src/vikit/vikit_cameras/include/vikit/cameras/radial_tangential_distortion.h  distort / undistort / jacobian (header only)
//   src/svo_common/include/svo/common/seed.h         seed::* accessors, isConverged, getSigma2FromDepthSigma (header only)
//   src/svo_common/include/svo/common/occupancy_grid_2d.h  OccupandyGrid2D::getCellIndex (header only)
// Their OpenCV / Eigen / glog includes resolve to the container-only stand-ins in oracle/shim/ (none of those libraries
// is installed here). No reference source is copied into this repository; this file only declares C entry points.
#include <vikit/vision.h>
#include <svo/direct/feature_alignment.h>
#include <svo/direct/patch_score.h>
#include <svo/direct/patch_utils.h>
#include <svo/common/seed.h>
#include <svo/common/occupancy_grid_2d.h>
#include <vikit/solver/robust_cost.h>
#include <vikit/cameras/radial_tangential_distortion.h>
#include <cstring>
#include <vector>

namespace {
// An image container as the reference creates them: cv::Mat::create -> continuous, 64-byte aligned (frame.cpp:380-385).
cv::Mat makeMat(const unsigned char* data, int w, int h, int stride) {
  cv::Mat m(h, w, CV_8UC1);
  for (int r = 0; r < h; ++r) std::memcpy(m.data + (size_t)r * m.step.p[0], data + (size_t)r * stride, w);
  return m;
}
}  // namespace

extern "C" {

// frame_utils::createImgPyramid (src/svo_common/src/frame.cpp:372-386) is four lines around vk::halfSample; the loop is
// re-typed here because frame.cpp itself needs the whole Frame class. levels_out[l] receives level l (l >= 1),
// (h >> l) x (w >> l) bytes, contiguous.
void ref_create_img_pyramid(const unsigned char* img, int w, int h, int stride, int n_levels, unsigned char** levels_out) {
  std::vector<cv::Mat> pyr(n_levels);
  pyr[0] = makeMat(img, w, h, stride);
  for (int i = 1; i < n_levels; ++i) {
    pyr[i] = cv::Mat(pyr[i - 1].rows / 2, pyr[i - 1].cols / 2, CV_8U);
    vk::halfSample(pyr[i - 1], pyr[i]);
    for (int r = 0; r < pyr[i].rows; ++r)
      std::memcpy(levels_out[i] + (size_t)r * pyr[i].cols, pyr[i].data + (size_t)r * pyr[i].step.p[0], pyr[i].cols);
  }
}

// one vk::halfSample call on caller-provided buffers (alignment / stride chosen by the caller, so that both branches of
// vision.cpp:72-111 can be exercised); out is (h/2) x (w/2) with row stride out_stride
void ref_half_sample(unsigned char* in, int w, int h, int stride, unsigned char* out, int out_stride) {
  cv::Mat min(h, w, CV_8UC1, in, stride), mout(h / 2, w / 2, CV_8UC1, out, out_stride);
  vk::halfSample(min, mout);
}

int ref_align2d(const unsigned char* img, int w, int h, int stride, unsigned char* patch_with_border, unsigned char* patch,
                int n_iter, int est_offset, int est_gain, double* px) {
  cv::Mat m(h, w, CV_8UC1, const_cast<unsigned char*>(img), stride);
  svo::Keypoint kp(px[0], px[1]);
  const bool ok = svo::feature_alignment::align2D(m, patch_with_border, patch, n_iter, est_offset != 0, est_gain != 0, kp);
  px[0] = kp[0]; px[1] = kp[1];
  return ok ? 1 : 0;
}

int ref_align1d(const unsigned char* img, int w, int h, int stride, const double* dir, unsigned char* patch_with_border,
                unsigned char* patch, int n_iter, int est_offset, int est_gain, double* px, double* h_inv) {
  cv::Mat m(h, w, CV_8UC1, const_cast<unsigned char*>(img), stride);
  svo::Keypoint kp(px[0], px[1]);
  svo::GradientVector d(dir[0], dir[1]);
  const bool ok = svo::feature_alignment::align1D(m, d, patch_with_border, patch, n_iter, est_offset != 0, est_gain != 0, &kp, h_inv);
  px[0] = kp[0]; px[1] = kp[1];
  return ok ? 1 : 0;
}

// alignPyr2D (pyramidal KLT, feature_alignment.cpp:761-973); the pyramids are given as n_levels pointers
int ref_align_pyr2d(unsigned char* const* ref_levels, unsigned char* const* cur_levels, const int* ws, const int* hs, const int* strides,
                    int n_levels, int max_level, int min_level, const int* patch_sizes, int n_patch_sizes, int n_iter,
                    float min_update_squared, const int* px_ref_level_0, double* px_cur_level_0) {
  std::vector<cv::Mat> pr, pc;
  for (int l = 0; l < n_levels; ++l) {
    pr.emplace_back(hs[l], ws[l], CV_8UC1, ref_levels[l], strides[l]);
    pc.emplace_back(hs[l], ws[l], CV_8UC1, cur_levels[l], strides[l]);
  }
  std::vector<int> ps(patch_sizes, patch_sizes + n_patch_sizes);
  Eigen::Vector2i pref(px_ref_level_0[0], px_ref_level_0[1]);
  svo::Keypoint kp(px_cur_level_0[0], px_cur_level_0[1]);
  const bool ok = svo::feature_alignment::alignPyr2D(pr, pc, max_level, min_level, ps, n_iter, min_update_squared, pref, kp);
  px_cur_level_0[0] = kp[0]; px_cur_level_0[1] = kp[1];
  return ok ? 1 : 0;
}

// ZMSSD<4>: score of the 8x8 reference patch against n candidate top-left corners cur + offsets[i] (row stride `stride`)
void ref_zmssd(unsigned char* ref_patch, unsigned char* cur, int stride, const int* offsets, int n, int* scores, int* threshold) {
  svo::patch_score::ZMSSD<4> z(ref_patch);
  for (int i = 0; i < n; ++i) scores[i] = z.computeScore(cur + offsets[i], stride);
  if (threshold) *threshold = svo::patch_score::ZMSSD<4>::threshold();
}

void ref_patch_from_patch_with_border(const unsigned char* patch_with_border, int patch_size, unsigned char* patch) {
  svo::patch_utils::createPatchFromPatchWithBorder(patch_with_border, patch_size, patch);
}

// vk::solver::TukeyWeightFunction(b).weight(error) for n errors (robust_cost.cpp:44-60; default b = 4.6851, robust_cost.h:70)
void ref_tukey_weight(float b, const float* err, int n, float* w) {
  vk::solver::TukeyWeightFunction f(b);
  for (int i = 0; i < n; ++i) w[i] = f.weight(err[i]);
}

// RadialTangentialDistortion: which = 0 distort, 1 undistort, 2 jacobian (out = J00, J01, J10, J11); xy in/out [n][2]
void ref_radtan(double k1, double k2, double p1, double p2, int which, double* xy, int n, double* jac_out) {
  vk::cameras::RadialTangentialDistortion d(k1, k2, p1, p2);
  for (int i = 0; i < n; ++i) {
    if (which == 0) {  // the Vector2d overload: the one PinholeProjection::project3 calls (pinhole_projection.hpp:54-55)
      const Eigen::Vector2d r = d.distort(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
      xy[2 * i] = r[0]; xy[2 * i + 1] = r[1];
    }
    else if (which == 1) d.undistort(xy[2 * i], xy[2 * i + 1]);
    else {
      const Eigen::Matrix2d J = d.jacobian(Eigen::Vector2d(xy[2 * i], xy[2 * i + 1]));
      jac_out[4 * i] = J(0, 0); jac_out[4 * i + 1] = J(0, 1); jac_out[4 * i + 2] = J(1, 0); jac_out[4 * i + 3] = J(1, 1);
    }
  }
}

// seed.h (inverse-depth parametrisation): out[0..5] = getDepth, getInvMinDepth, getInvMaxDepth, isConverged,
// getSigma2FromDepthSigma(depth, depth_sigma), getInitSigma2FromMuRange(mu_range)
void ref_seed_helpers(const double* state4, double mu_range, double sigma2_convergence_threshold, double depth, double depth_sigma, double* out) {
  svo::SeedState s;
  s << state4[0], state4[1], state4[2], state4[3];
  out[0] = svo::seed::getDepth(s);
  out[1] = svo::seed::getInvMinDepth(s);
  out[2] = svo::seed::getInvMaxDepth(s);
  out[3] = svo::seed::isConverged(s, mu_range, sigma2_convergence_threshold) ? 1.0 : 0.0;
  out[4] = svo::seed::getSigma2FromDepthSigma(depth, depth_sigma);
  out[5] = svo::seed::getInitSigma2FromMuRange(mu_range);
}

// OccupandyGrid2D(cell_size, n_cols, n_rows).getCellIndex(x, y, scale) for n corners (occupancy_grid_2d.h:82-95)
void ref_grid_cell_index(int cell_size, int n_cols, int n_rows, const int* xy, const int* scale, int n, long long* idx) {
  svo::OccupandyGrid2D grid(cell_size, n_cols, n_rows);
  for (int i = 0; i < n; ++i) idx[i] = (long long)grid.getCellIndex(xy[2 * i], xy[2 * i + 1], scale[i]);
}

}  // extern "C"
