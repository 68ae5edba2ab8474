// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Row f4 of SURVEY.md §8: PoseOptimizer::run — Gauss-Newton on the reprojection residuals of one frame bundle with a MAD scale
// estimate, Tukey weights, optional rotation prior and outlier removal.
// ref: src/svo/src/pose_optimizer.cpp:18-336 (run, evaluateErrorImpl, removeOutliers, update, applyPrior), :338-629 (residuals),
//      src/vikit/vikit_solver/include/vikit/solver/implementation/mini_least_squares_solver.hpp:42-107,230-262,
//      src/vikit/vikit_solver/src/robust_cost.cpp:19-26 (MADScaleEstimator), :44-60 (TukeyWeightFunction, b = 4.6851),
//      src/svo_common/include/svo/common/frame.h:342-397 (Jacobians), src/vikit/vikit_common/include/vikit/math_utils.h:165-172.
// Parity status: pinned — the reference's own pose_optimizer.cpp compiles into oracle/_ref/libfrontend_ref.so
// (ref_pose_optimize in ref_frontend_wrapper.cpp) and this restatement is checked against it.
#pragma once
#include <algorithm>
#include <cmath>
#include <vector>
#include "orc_capi.h"
#include "orc_math.hpp"
#include "orc_sparse_align.hpp"

namespace orc {

struct PoseOptFeature {
  V2 px; V3 f; V2 grad; int level; FeatureType type; V3 xyz_world; bool has_xyz; int cam;
};
struct PoseOptCam { Camera cam; SE3 T_cam_imu; };

inline M3 skew(const V3& v) {  // vikit/math_utils.h:85-92
  M3 m;
  m.m[0][0] = 0; m.m[0][1] = -v.z; m.m[0][2] = v.y;
  m.m[1][0] = v.z; m.m[1][1] = 0; m.m[1][2] = -v.x;
  m.m[2][0] = -v.y; m.m[2][1] = v.x; m.m[2][2] = 0;
  return m;
}

// J[r][c] = sum_k A[r][k] * B[k][c] in Eigen's evaluation order (k ascending)
template <int R, int K, int Cn>
inline void matMul(const double (&A)[R][K], const double (&B)[K][Cn], double (&out)[R][Cn]) {
  for (int r = 0; r < R; ++r)
    for (int c = 0; c < Cn; ++c) {
      double s = A[r][0] * B[0][c];
      for (int k = 1; k < K; ++k) s += A[r][k] * B[k][c];
      out[r][c] = s;
    }
}
inline void generators(const V3& p_in_imu, double (&G)[3][6]) {  // G_x = [I | -skew(p)]
  const M3 S = skew(p_in_imu);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) { G[r][c] = r == c ? 1.0 : 0.0; G[r][3 + c] = -S.m[r][c]; }
}
// frame.h:342-357
inline void jacobian_xyz2uv_imu(const SE3& T_cam_imu, const V3& p_in_imu, double (&J)[2][6]) {
  double G[3][6];
  generators(p_in_imu, G);
  const V3 p = T_cam_imu * p_in_imu;
  const double Jp[2][3] = {{1, 0, -p.x / p.z}, {0, 1, -p.y / p.z}};
  const M3 R = quatToMatrix(T_cam_imu.q);
  const double s = -1.0 / p.z;
  double sJ[2][3], sJR[2][3];
  for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) sJ[r][c] = s * Jp[r][c];   // (-1/z * J_proj) ...
  matMul(sJ, R.m, sJR);                                                               // ... * R ...
  matMul(sJR, G, J);                                                                  // ... * G_x
}
// frame.h:360-371
inline void jacobian_xyz2img_imu(const SE3& T_cam_imu, const V3& p_in_imu, const double (&J_cam)[2][3], double (&J)[2][6]) {
  double G[3][6];
  generators(p_in_imu, G);
  const M3 R = quatToMatrix(T_cam_imu.q);
  double JR[2][3];
  matMul(J_cam, R.m, JR);
  matMul(JR, G, J);
}
// frame.h:374-397
inline void jacobian_xyz2f_imu(const SE3& T_cam_imu, const V3& p_in_imu, double (&J)[3][6]) {
  double G[3][6];
  generators(p_in_imu, G);
  const V3 p = T_cam_imu * p_in_imu;
  const double x2 = p.x * p.x, y2 = p.y * p.y, z2 = p.z * p.z, xy = p.x * p.y, yz = p.y * p.z, zx = p.z * p.x;
  double Jn[3][3] = {{y2 + z2, -xy, -zx}, {-xy, x2 + z2, -yz}, {-zx, -yz, x2 + y2}};
  const double s = 1 / std::pow(x2 + y2 + z2, 1.5);
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Jn[r][c] *= s;
  const M3 R = quatToMatrix(T_cam_imu.q);
  double JR[3][3];
  matMul(Jn, R.m, JR);
  matMul(JR, G, J);
}

struct PoseOptimizer {
  enum ErrorType { kUnitPlane = 0, kBearingVectorDiff = 1, kImagePlane = 2 };
  int max_iter = 10;
  double eps = 0.000001;
  int err_type = kUnitPlane;
  bool have_prior = false;
  SE3 prior;
  double prior_lambda = 0.0;
  double I_prior[6] = {0, 0, 0, 0, 0, 0};  // diagonal: the reference's I_prior_ is zero outside the rotation block's diagonal
  TukeyWeightFunction robust_weight;      // default b = 4.6851
  double measurement_sigma = 1.0, focal_length = 1.0;
  double H[6][6], g[6];
  size_t n_meas = 0, iter = 0;
  bool stop = false;
  double chi2 = 1e10;

  // accumulate one residual: e (dimension DIM, already whitened), Jacobian rows J (DIM x 6, already whitened)
  template <int DIM>
  void accumulate(const double (&J)[DIM][6], const double (&e)[DIM], double weight) {
    for (int a = 0; a < 6; ++a)
      for (int b = 0; b < 6; ++b) {
        double s = J[0][a] * J[0][b];
        for (int k = 1; k < DIM; ++k) s += J[k][a] * J[k][b];
        H[a][b] += s * weight;
      }
    for (int a = 0; a < 6; ++a) {
      double s = J[0][a] * e[0];
      for (int k = 1; k < DIM; ++k) s += J[k][a] * e[k];
      g[a] -= s * weight;
    }
  }

  // one feature (pose_optimizer.cpp:338-629); returns false when the feature carries no 3-D point
  bool residual(const PoseOptFeature& ft, const PoseOptCam& pc, const SE3& T_imu_world, double measurement_sigma_, bool with_jacobian,
                double* unwhitened_error, double* chi2_error) {
    const V3 xyz_in_imu = T_imu_world * ft.xyz_world;
    const V3 xyz_in_cam = pc.T_cam_imu * xyz_in_imu;
    const bool edgelet = isEdgelet(ft.type);
    const double R = 1.0 / measurement_sigma_;
    if (!edgelet && err_type == kUnitPlane) {
      double e[2] = {ft.f.x / ft.f.z - xyz_in_cam.x / xyz_in_cam.z, ft.f.y / ft.f.z - xyz_in_cam.y / xyz_in_cam.z};
      *unwhitened_error = std::sqrt(e[0] * e[0] + e[1] * e[1]);
      e[0] *= R; e[1] *= R;
      const double weight = robust_weight.weight(float(std::sqrt(e[0] * e[0] + e[1] * e[1])));
      *chi2_error = 0.5 * (e[0] * e[0] + e[1] * e[1]) * weight;
      if (with_jacobian) {
        double J[2][6];
        jacobian_xyz2uv_imu(pc.T_cam_imu, xyz_in_imu, J);
        for (auto& row : J) for (double& v : row) v *= R;
        accumulate(J, e, weight);
      }
    } else if (!edgelet && err_type == kImagePlane) {
      double J_cam[2][3];
      const V2 px_est = pc.cam.project3(xyz_in_cam, J_cam);
      double e[2] = {ft.px.x - px_est.x, ft.px.y - px_est.y};
      *unwhitened_error = std::sqrt(e[0] * e[0] + e[1] * e[1]);
      e[0] *= R; e[1] *= R;
      const double weight = robust_weight.weight(float(std::sqrt(e[0] * e[0] + e[1] * e[1])));
      *chi2_error = 0.5 * (e[0] * e[0] + e[1] * e[1]) * weight;
      if (with_jacobian) {
        double J[2][6];
        jacobian_xyz2img_imu(pc.T_cam_imu, xyz_in_imu, J_cam, J);
        for (auto& row : J) for (double& v : row) v = ((-1.0) * v) * R;
        accumulate(J, e, weight);
      }
    } else if (!edgelet) {  // kBearingVectorDiff
      const V3 fe = normalized(xyz_in_cam);
      double e[3] = {ft.f.x - fe.x, ft.f.y - fe.y, ft.f.z - fe.z};
      *unwhitened_error = std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
      for (double& v : e) v *= R;
      const double weight = robust_weight.weight(float(std::sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2])));
      *chi2_error = 0.5 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * weight;
      if (with_jacobian) {
        double J[3][6];
        jacobian_xyz2f_imu(pc.T_cam_imu, xyz_in_imu, J);
        for (auto& row : J) for (double& v : row) v = ((-1.0) * v) * R;
        accumulate(J, e, weight);
      }
    } else if (err_type == kUnitPlane) {  // edgelets: the error along the gradient direction
      double e = ft.grad.x * (ft.f.x / ft.f.z - xyz_in_cam.x / xyz_in_cam.z) + ft.grad.y * (ft.f.y / ft.f.z - xyz_in_cam.y / xyz_in_cam.z);
      *unwhitened_error = std::abs(e);
      e *= R;
      const double weight = robust_weight.weight(float(e));
      *chi2_error = 0.5 * e * e * weight;
      if (with_jacobian) {
        double Jp[2][6], J[1][6];
        jacobian_xyz2uv_imu(pc.T_cam_imu, xyz_in_imu, Jp);
        for (int c = 0; c < 6; ++c) J[0][c] = (ft.grad.x * Jp[0][c] + ft.grad.y * Jp[1][c]) * R;
        const double ee[1] = {e};
        accumulate(J, ee, weight);
      }
    } else if (err_type == kImagePlane) {
      double J_cam[2][3];
      const V2 px_est = pc.cam.project3(xyz_in_cam, J_cam);
      double e = ft.grad.x * (ft.px.x - px_est.x) + ft.grad.y * (ft.px.y - px_est.y);
      *unwhitened_error = std::abs(e);
      e *= R;
      const double weight = robust_weight.weight(float(e));
      *chi2_error = 0.5 * e * e * weight;
      if (with_jacobian) {
        double Jp[2][6], J[1][6];
        jacobian_xyz2img_imu(pc.T_cam_imu, xyz_in_imu, J_cam, Jp);
        for (int c = 0; c < 6; ++c) J[0][c] = ((ft.grad.x * (-1.0)) * Jp[0][c] + (ft.grad.y * (-1.0)) * Jp[1][c]) * R;
        const double ee[1] = {e};
        accumulate(J, ee, weight);
      }
    } else {  // edgelet, kBearingVectorDiff (pose_optimizer.cpp:560-627)
      double J_cam[2][3];
      const V2 px_est = pc.cam.project3(xyz_in_cam, J_cam);
      const double pd[2] = {ft.px.x - px_est.x, ft.px.y - px_est.y};
      const double pd2 = pd[0] * pd[0] + pd[1] * pd[1];
      const V3 fe = normalized(xyz_in_cam);
      const double fd[3] = {ft.f.x - fe.x, ft.f.y - fe.y, ft.f.z - fe.z};
      const double fd2 = fd[0] * fd[0] + fd[1] * fd[1] + fd[2] * fd[2];
      const double e_img = ft.grad.x * pd[0] + ft.grad.y * pd[1];
      const double scale_ratio = std::sqrt(fd2) / std::sqrt(pd2);
      double e = e_img * scale_ratio;
      *unwhitened_error = std::abs(e);
      e *= R;
      const double weight = robust_weight.weight(float(e));
      *chi2_error = 0.5 * e * e * weight;
      if (with_jacobian) {
        double Jp[2][6], Jb[3][6], J[1][6];
        jacobian_xyz2img_imu(pc.T_cam_imu, xyz_in_imu, J_cam, Jp);
        jacobian_xyz2f_imu(pc.T_cam_imu, xyz_in_imu, Jb);
        const double k = (0.5) * (1.0 / (scale_ratio)) * (1 / (pd2 * pd2));
        for (int c = 0; c < 6; ++c) {
          const double J_img = (ft.grad.x * (-1.0)) * Jp[0][c] + (ft.grad.y * (-1.0)) * Jp[1][c];
          const double J_ftf = ((2 * fd[0]) * (-1.0)) * Jb[0][c] + ((2 * fd[1]) * (-1.0)) * Jb[1][c] + ((2 * fd[2]) * (-1.0)) * Jb[2][c];
          const double J_ptp = ((2 * pd[0]) * (-1.0)) * Jp[0][c] + ((2 * pd[1]) * (-1.0)) * Jp[1][c];
          const double J_ratio = k * (J_ftf * pd2 - J_ptp * fd2);
          J[0][c] = (e_img * J_ratio + scale_ratio * J_img) * R;
        }
        const double ee[1] = {e};
        accumulate(J, ee, weight);
      }
    }
    return true;
  }

  // pose_optimizer.cpp:96-196
  double evaluateErrorImpl(const std::vector<PoseOptFeature>& fts, const std::vector<PoseOptCam>& cams, const SE3& T_imu_world,
                           bool with_jacobian, std::vector<float>* unwhitened_errors) {
    double chi2_error_sum = 0.0;
    for (const PoseOptFeature& ft : fts) {
      if (!ft.has_xyz) continue;
      const int scale = (1 << ft.level);
      double ms = measurement_sigma * scale;
      if (isEdgelet(ft.type)) ms *= 2.0;  // kEdgeletSigmaExtraFactor
      double ue, ce;
      residual(ft, cams[ft.cam], T_imu_world, ms, with_jacobian, &ue, &ce);
      if (unwhitened_errors) unwhitened_errors->push_back(float(ue / scale));
      chi2_error_sum += ce;
      ++n_meas;
    }
    return chi2_error_sum;
  }

  // run (pose_optimizer.cpp:39-94) after reset(); outlier[i] = 1 where removeOutliers marks the feature kOutlier
  size_t run(const std::vector<PoseOptFeature>& fts, const std::vector<PoseOptCam>& cams, SE3& T_imu_world, double reproj_thresh_px,
             uint8_t* outlier, double* stats) {
    focal_length = cams[0].cam.errorMultiplier();
    n_meas = 0; iter = 0; stop = false; chi2 = 1e10;
    std::vector<float> start_errors;
    evaluateErrorImpl(fts, cams, T_imu_world, false, &start_errors);
    {  // MADScaleEstimator::compute
      std::vector<float> e = start_errors;
      auto it = e.begin() + std::floor(e.size() / 2);
      std::nth_element(e.begin(), it, e.end());
      measurement_sigma = 1.48f * (*it);
    }
    // optimizeGaussNewton (mini_least_squares_solver.hpp:42-107)
    SE3 state = T_imu_world, old_state = T_imu_world;
    for (iter = 0; iter < size_t(max_iter); ++iter) {
      for (auto& row : H) for (double& v : row) v = 0.0;
      for (double& v : g) v = 0.0;
      n_meas = 0;
      const double new_chi2 = evaluateErrorImpl(fts, cams, state, true, nullptr);
      if (have_prior) {  // applyPrior (pose_optimizer.cpp:311-334)
        if (iter == 0) {
          double H_max_diag = 0;
          for (int j = 3; j < 6; ++j) H_max_diag = std::max(H_max_diag, std::fabs(H[j][j]));
          for (int j = 0; j < 6; ++j) I_prior[j] = (j >= 3 ? 1.0 : 0.0) * (H_max_diag * prior_lambda);
        }
        double l[6];
        se3Log(state * inverse(prior), l);
        for (int j = 0; j < 6; ++j) { H[j][j] += I_prior[j]; g[j] -= I_prior[j] * l[j]; }
      }
      double dx[6];
      ldltSolve<6>(H, g, dx);
      if (std::isnan(dx[0])) stop = true;
      if (stop) { state = old_state; break; }
      // update (pose_optimizer.cpp:300-309): T_new = exp(dx) * T_old, quaternion normalised
      SE3 new_state = se3Exp(dx) * state;
      const double n = std::sqrt(new_state.q.w * new_state.q.w + new_state.q.x * new_state.q.x + new_state.q.y * new_state.q.y + new_state.q.z * new_state.q.z);
      new_state.q = {new_state.q.w / n, new_state.q.x / n, new_state.q.y / n, new_state.q.z / n};
      old_state = state;
      state = new_state;
      chi2 = new_chi2;
      double x_norm = 0.0;
      for (double v : dx) x_norm = std::max(x_norm, std::fabs(v));
      if (x_norm < eps) break;
    }
    T_imu_world = state;
    // removeOutliers (pose_optimizer.cpp:198-298) with measurement_sigma = 0 (only the unwhitened error is used)
    double outlier_threshold = reproj_thresh_px;
    if (err_type == kUnitPlane) outlier_threshold = reproj_thresh_px / focal_length;
    else if (err_type == kBearingVectorDiff) outlier_threshold = std::fabs(2 * std::sin(0.5 * cams[0].cam.getAngleError(reproj_thresh_px)));
    size_t n_deleted = 0;
    std::vector<double> final_errors;
    for (size_t i = 0; i < fts.size(); ++i) {
      outlier[i] = 0;
      if (!fts[i].has_xyz) continue;
      double ue, ce;
      residual(fts[i], cams[fts[i].cam], state, 0.0, false, &ue, &ce);
      ue *= 1.0 / (1 << fts[i].level);
      final_errors.push_back(ue);
      if (std::fabs(ue) > outlier_threshold) { outlier[i] = 1; ++n_deleted; }
    }
    const double error_scale = (err_type == kUnitPlane) ? focal_length : 1.0;
    auto med = [](auto v) { auto it = v.begin() + std::floor(v.size() / 2); std::nth_element(v.begin(), it, v.end()); return double(*it); };
    stats[0] = measurement_sigma; stats[1] = med(start_errors) * error_scale; stats[2] = med(final_errors) * error_scale;
    stats[3] = double(iter); stats[4] = double(n_meas); stats[5] = chi2;
    return n_meas - n_deleted;
  }
};

}  // namespace orc
