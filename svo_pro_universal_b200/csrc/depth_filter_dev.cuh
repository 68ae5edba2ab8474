// Device functions of the depth-filter seed update, shared by depth_filter.cu (d) and reprojector.cu (f1: the Reprojector
// updates unconverged seeds through depth_filter_utils::updateSeed, src/svo/src/reprojector.cpp:412-419).
//
// ref: src/svo_direct/src/depth_filter.cpp:367-499 (updateSeed), :501-552 (updateFilterVogiatzis), :554-578
//      (updateFilterGaussian), :580-596 (computeTau); src/svo_common/include/svo/common/seed.h:110-169;
//      src/vikit/vikit_common/include/vikit/math_utils.h:186-194 (normPdf)
#pragma once
#include "matcher_dev.cuh"

namespace svo_dev {

SVO_D double normPdf(double x, double mean, double sigma) {
  double exponent = x - mean;
  exponent *= -exponent;
  exponent /= 2 * sigma * sigma;
  double result = exp(exponent);
  result /= sigma * sqrt(2 * 3.14159265358979323846);
  return result;
}

// depth_filter.cpp:501-552; s = (mu, sigma2, a, b) in/out
SVO_D bool updateFilterVogiatzis(double z, double tau2, double mu_range, double s[4]) {
  double mu = s[0], sigma2 = s[1], a = s[2], b = s[3];
  const double norm_scale = sqrt(sigma2 + tau2);
  if (norm_scale != norm_scale) return false;
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
  const double e = C1 * (a + 1.0) * (a + 2.0) / ((a + b + 1.0) * (a + b + 2.0)) + C2 * a * (a + 1.0) / ((a + b + 1.0) * (a + b + 2.0));
  const double mu_new = C1 * m + C2 * mu;
  sigma2 = C1 * (s2 + m * m) + C2 * (sigma2 + mu * mu) - mu_new * mu_new;
  mu = mu_new;
  a = (e - f) / (f - e / f);
  b = a * (1.0 - f) / f;
  bool ok = true;
  if (sigma2 < 0.0) sigma2 = oldsigma2;
  if (mu < 0.0) { mu = 1.0; ok = false; }
  s[0] = mu; s[1] = sigma2; s[2] = a; s[3] = b;
  return ok;
}

// depth_filter.cpp:554-578
SVO_D bool updateFilterGaussian(double z, double tau2, double s[4]) {
  const double norm_scale = sqrt(s[1] + tau2);
  if (norm_scale != norm_scale) return false;
  const double denom = s[1] + tau2;
  s[0] = (s[1] * z + tau2 * s[0]) / denom;
  s[1] = s[1] * tau2 / denom;
  return true;
}

// depth_filter.cpp:580-596
SVO_D double computeTau(const V3d& t, const V3d& f, double z, double px_error_angle) {
  const V3d a = f * z - t;
  const double t_norm = norm3(t);
  const double a_norm = norm3(a);
  const double alpha = acos(dot3(f, t) / t_norm);
  const double beta = acos(dot3(a, -t) / (t_norm * a_norm));
  const double beta_plus = beta + px_error_angle;
  const double gamma_plus = 3.14159265358979323846 - alpha - beta_plus;
  const double z_plus = t_norm * sin(beta_plus) / sin(gamma_plus);
  return z_plus - z;
}

// depth_filter_utils::updateSeed for ONE observation (depth_filter.cpp:387-499; the same-frame test :377-381 is the caller's).
// `type` and `st` = (inverse mu, sigma2, a, b) are the seed's in/out state; returns true when the filter was updated.
// *match_result receives the Matcher::MatchResult (-1 when no match was attempted).
SVO_D bool updateSeedOnce(const Group& g, const PyrView& ref_pyr, int ref_frame, const PyrView& cur_pyr, int cur_frame,
                          const svo_camera& cam_ref, const svo_camera& cam_cur, const SE3d& T, svo_feature& ft, int& type, double st[4],
                          double mu_range, double sigma2_convergence_threshold, double px_error_angle, bool check_visibility,
                          bool check_convergence, bool use_vogiatzis_update, const svo_matcher_options& mopt, uint8_t* pwb,
                          MatchState& m, int* match_result) {
  *match_result = -1;
  if (type == kOutlier) return false;  // :387-392
  if ((type == kCornerSeedConverged || type == kEdgeletSeedConverged || type == kMapPointSeedConverged) && check_convergence)
    return false;  // :394-399
  const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
  if (check_visibility) {  // :406-420
    const V3d xyz_f = se3Apply(T, f_ref * (1.0 / st[0]));
    const V2d px = camProject3(cam_cur, xyz_f);
    if (!(px.x >= 0.0 && px.y >= 0.0 && px.x < (double)cam_cur.width && px.y < (double)cam_cur.height)) return false;
    const int pxi0 = (int)px.x, pxi1 = (int)px.y;
    const int boundary = 9;
    if (!(pxi0 >= boundary && pxi1 >= boundary && pxi0 < cam_cur.width - boundary && pxi1 < cam_cur.height - boundary)) return false;
  }
  const bool align_1d = (type == kEdgeletSeed || type == kEdgeletSeedConverged);  // :423-427
  ft.type = type;
  initMatchState(m);
  double depth = 0.0;
  // seed.h:115-128: d_estimate_inv = mu, d_min_inv = mu + sigma, d_max_inv = max(mu - sigma, 1e-8)
  const double sig = sqrt(st[1]);
  const int mr = findEpipolarMatchDirect(g, ref_pyr, ref_frame, cur_pyr, cur_frame, cam_ref, cam_cur, T, ft, st[0], st[0] + sig,
                                         fmax(st[0] - sig, 0.00000001), mopt, align_1d, pwb, m, &depth);
  *match_result = mr;
  if (mr != kSuccess) {
    if (!m.reject) st[3] += 1;  // seed::increaseOutlierProbability, :445-450
    return false;
  }
  const SE3d T_ref_cur = se3Inv(T);
  const double depth_sigma = computeTau(T_ref_cur.t, f_ref, depth, px_error_angle);  // :459
  const double zi = 1.0 / depth;
  // seed::getSigma2FromDepthSigma (seed.h:155-160)
  const double sg = 0.5 * (1.0 / fmax(0.000000000001, depth - depth_sigma) - 1.0 / (depth + depth_sigma));
  const double tau2 = sg * sg;
  const bool ok = use_vogiatzis_update ? updateFilterVogiatzis(zi, tau2, mu_range, st) : updateFilterGaussian(zi, tau2, st);
  if (!ok) {
    type = kOutlier;  // :470-471, :481-482
    return false;
  }
  const double thresh = mu_range / sigma2_convergence_threshold;  // seed::isConverged, seed.h:145-153
  if (st[1] < thresh * thresh) {
    if (type == kCornerSeed) type = kCornerSeedConverged;
    else if (type == kEdgeletSeed) type = kEdgeletSeedConverged;
    else if (type == kMapPointSeed) type = kMapPointSeedConverged;
  }
  return true;
}

}  // namespace svo_dev
