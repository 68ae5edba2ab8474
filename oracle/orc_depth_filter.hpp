// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Rows d1-d4 of SURVEY.md §8: depth-filter seed update (Vogiatzis Gaussian x Beta filter,
// inverse-depth parametrisation) driven by the epipolar matcher.
// Parity status: "parity unpinned" (restatement; no reference vectors exist for this path).
#pragma once
#include <cmath>
#include "orc_math.hpp"
#include "orc_matcher.hpp"

namespace orc {

// ref: src/svo_common/include/svo/common/seed.h:110-169 (inverse depth parametrisation)
namespace seed {
inline double getDepth(const double* s) { return 1.0 / s[0]; }
inline double getInvDepth(const double* s) { return s[0]; }
inline double getInvMinDepth(const double* s) { return s[0] + std::sqrt(s[1]); }
inline double getInvMaxDepth(const double* s) { return std::max(s[0] - std::sqrt(s[1]), 0.00000001); }
inline double getMeanFromDepth(double depth) { return 1.0 / depth; }
inline double getMeanRangeFromDepthMinMax(double depth_min, double /*depth_max*/) { return 1.0 / depth_min; }
inline double getInitSigma2FromMuRange(double mu_range) { return mu_range * mu_range / 36.0; }
inline bool isConverged(const double* s, double mu_range, double sigma2_convergence_threshold) {
  const double thresh = mu_range / sigma2_convergence_threshold;
  return (s[1] < thresh * thresh);
}
inline double getSigma2FromDepthSigma(double depth, double depth_sigma) {
  const double sigma = 0.5 * (1.0 / std::max(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));
  return sigma * sigma;
}
inline void increaseOutlierProbability(double* s) { s[3] += 1; }
}  // namespace seed

// ref: src/vikit/vikit_common/include/vikit/math_utils.h:186-194
inline double normPdf(const double x, const double mean, const double sigma) {
  double exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sigma * sigma;
  double result = std::exp(exponent);
  result /= sigma * std::sqrt(2 * M_PI);
  return result;
}

// d3. ref: src/svo_direct/src/depth_filter.cpp:501-552
inline bool updateFilterVogiatzis(const double z, const double tau2, const double mu_range, double* mu_sigma2_a_b) {
  double& mu = mu_sigma2_a_b[0];
  double& sigma2 = mu_sigma2_a_b[1];
  double& a = mu_sigma2_a_b[2];
  double& b = mu_sigma2_a_b[3];
  const double norm_scale = std::sqrt(sigma2 + tau2);
  if (std::isnan(norm_scale)) return false;
  const double oldsigma2 = sigma2;
  const double s2 = 1.0 / (1.0 / sigma2 + 1.0 / tau2);
  const double m = s2 * (mu / sigma2 + z / tau2);
  const double uniform_x = 1.0 / mu_range;
  double C1 = a / (a + b) * normPdf(z, mu, norm_scale);
  double C2 = b / (a + b) * uniform_x;
  const double normalization_constant = C1 + C2;
  C1 /= normalization_constant;
  C2 /= normalization_constant;
  const double f = C1 * (a + 1.0) / (a + b + 1.0) + C2 * a / (a + b + 1.0);
  const double e = C1 * (a + 1.0) * (a + 2.0) / ((a + b + 1.0) * (a + b + 2.0))
                 + C2 * a * (a + 1.0) / ((a + b + 1.0) * (a + b + 2.0));
  const double mu_new = C1 * m + C2 * mu;
  sigma2 = C1 * (s2 + m * m) + C2 * (sigma2 + mu * mu) - mu_new * mu_new;
  mu = mu_new;
  a = (e - f) / (f - e / f);
  b = a * (1.0 - f) / f;
  if (sigma2 < 0.0) sigma2 = oldsigma2;
  if (mu < 0.0) {
    mu = 1.0;
    return false;
  }
  return true;
}

// ref: depth_filter.cpp:554-578
inline bool updateFilterGaussian(const double z, const double tau2, double* mu_sigma2_a_b) {
  double& mu = mu_sigma2_a_b[0];
  double& sigma2 = mu_sigma2_a_b[1];
  const double norm_scale = std::sqrt(sigma2 + tau2);
  if (std::isnan(norm_scale)) return false;
  const double denom = (sigma2 + tau2);
  mu = (sigma2 * z + tau2 * mu) / denom;
  sigma2 = sigma2 * tau2 / denom;
  return true;
}

// d2. ref: depth_filter.cpp:580-596
inline double computeTau(const SE3& T_ref_cur, const V3& f, const double z, const double px_error_angle) {
  const V3 t = T_ref_cur.t;
  const V3 a = f * z - t;
  const double t_norm = norm(t);
  const double a_norm = norm(a);
  const double alpha = std::acos(dot(f, t) / t_norm);
  const double beta = std::acos(dot(a, -t) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = M_PI - alpha - beta_plus;
  const double z_plus = t_norm * std::sin(beta_plus) / std::sin(gamma_plus);
  return (z_plus - z);
}

// d1. depth_filter_utils::updateSeed — ref: depth_filter.cpp:367-499
// `type` is in/out (kOutlier on filter failure, *Converged on convergence); `state` is the 4-vector
// (inv-mu, sigma2, a, b) updated in place. The same-frame-id test (:377-381) is the caller's job here.
// `px_error_angle` is the reference's function-static computed from the first cur frame's camera (:384).
inline bool updateSeed(const MatchFrame& cur_frame, const MatchFrame& ref_frame, const SE3& T_cur_ref,
                       FeatureRef ref_ftr, FeatureType& type, double* state, const double seed_mu_range,
                       Matcher& matcher, const double sigma2_convergence_threshold, const double px_error_angle,
                       const bool check_visibility = true, const bool check_convergence = false,
                       const bool use_vogiatzis_update = true, int* match_result = nullptr) {
  if (match_result) *match_result = -1;
  if (type == FeatureType::kOutlier) return false;
  if ((type == FeatureType::kCornerSeedConverged || type == FeatureType::kEdgeletSeedConverged
       || type == FeatureType::kMapPointSeedConverged) && check_convergence)
    return false;
  ref_ftr.type = type;
  if (check_visibility) {
    const V3 xyz_f = T_cur_ref * (ref_ftr.f * seed::getDepth(state));
    const V2 px = cur_frame.cam.project3(xyz_f);
    if (!cur_frame.cam.isKeypointVisible(px.x, px.y)) return false;
    const int pxi0 = int(px.x), pxi1 = int(px.y);
    const int boundary = 9;
    if (!cur_frame.cam.isKeypointVisibleWithMarginInt(pxi0, pxi1, boundary)) return false;
  }
  if (ref_ftr.type == FeatureType::kEdgeletSeed || ref_ftr.type == FeatureType::kEdgeletSeedConverged)
    matcher.options_.align_1d = true;
  else
    matcher.options_.align_1d = false;

  double depth;
  const Matcher::MatchResult res = matcher.findEpipolarMatchDirect(
      ref_frame, cur_frame, T_cur_ref, ref_ftr, seed::getInvDepth(state), seed::getInvMinDepth(state),
      seed::getInvMaxDepth(state), depth);
  if (match_result) *match_result = int(res);
  if (res != Matcher::MatchResult::kSuccess) {
    if (!matcher.reject_) seed::increaseOutlierProbability(state);
    return false;
  }
  const double depth_sigma = computeTau(inverse(T_cur_ref), ref_ftr.f, depth, px_error_angle);
  if (use_vogiatzis_update) {
    if (!updateFilterVogiatzis(seed::getMeanFromDepth(depth), seed::getSigma2FromDepthSigma(depth, depth_sigma), seed_mu_range, state)) {
      type = FeatureType::kOutlier;
      return false;
    }
  } else {
    if (!updateFilterGaussian(seed::getMeanFromDepth(depth), seed::getSigma2FromDepthSigma(depth, depth_sigma), state)) {
      type = FeatureType::kOutlier;
      return false;
    }
  }
  if (seed::isConverged(state, seed_mu_range, sigma2_convergence_threshold)) {
    if (type == FeatureType::kCornerSeed) type = FeatureType::kCornerSeedConverged;
    else if (type == FeatureType::kEdgeletSeed) type = FeatureType::kEdgeletSeedConverged;
    else if (type == FeatureType::kMapPointSeed) type = FeatureType::kMapPointSeedConverged;
  }
  return true;
}

}  // namespace orc
