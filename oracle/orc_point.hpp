// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// SURVEY.md §8 row f4, second half: svo::Point::optimize — Gauss-Newton refinement of a 3-D point from its observations.
// ref: src/svo_common/src/point.cpp:216-325 (updateHessianGradientUnitPlane / UnitSphere, optimize),
//      src/svo_common/include/svo/common/point.h:170-204 (jacobian_xyz2uv, jacobian_xyz2f).
// Pinned against the reference's own point.cpp compiled into oracle/_ref/libpoint_ref.so.
#pragma once
#include <cmath>
#include <vector>
#include "orc_math.hpp"
#include "orc_sparse_align.hpp"  // ldltSolve (Eigen's pivoted LDL^T)

namespace orc {

struct PointObservation {
  SE3 T_f_w;  // the observing frame's pose
  V3 f;       // its unit bearing vector of the point
};

// Returns the number of iterations started (the reference has no such output; used by tests only).
inline int pointOptimize(const std::vector<PointObservation>& obs, V3& pos, size_t n_iter, bool using_bearing_vector) {
  V3 old_point = pos;
  double chi2 = 0.0;
  if (obs.size() < 2) return 0;  // point.cpp:255-259
  const double eps = 0.0000000001;
  int iters = 0;
  for (size_t it = 0; it < n_iter; ++it) {
    ++iters;
    double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0};
    double new_chi2 = 0.0;
    for (const PointObservation& o : obs) {
      const V3 p = o.T_f_w * pos;
      const M3 R = quatToMatrix(o.T_f_w.q);
      if (!using_bearing_vector) {
        // jacobian_xyz2uv: J = -[1/z 0 -x/z^2; 0 1/z -y/z^2] * R_f_w ; e = project2(f) - project2(p)
        const double z_inv = 1.0 / p.z, z_inv_sq = z_inv * z_inv;
        const double J0[2][3] = {{z_inv, 0.0, -p.x * z_inv_sq}, {0.0, z_inv, -p.y * z_inv_sq}};
        double J[2][3];
        for (int r = 0; r < 2; ++r)
          for (int c = 0; c < 3; ++c) J[r][c] = -((J0[r][0] * R.m[0][c] + J0[r][1] * R.m[1][c]) + J0[r][2] * R.m[2][c]);
        const double e[2] = {o.f.x / o.f.z - p.x / p.z, o.f.y / o.f.z - p.y / p.z};
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) A[r][c] += J[0][r] * J[0][c] + J[1][r] * J[1][c];
          b[r] -= J[0][r] * e[0] + J[1][r] * e[1];
        }
        new_chi2 += e[0] * e[0] + e[1] * e[1];
      } else {
        // jacobian_xyz2f: J = -(1 / |p|^3) [y2+z2 -xy -zx; -xy x2+z2 -yz; -zx -yz x2+y2] * R_f_w ; e = f - p / |p|
        const double x2 = p.x * p.x, y2 = p.y * p.y, z2 = p.z * p.z, xy = p.x * p.y, yz = p.y * p.z, zx = p.z * p.x;
        const double s = 1.0 / std::pow(x2 + y2 + z2, 1.5);
        const double N[3][3] = {{(y2 + z2) * s, -xy * s, -zx * s}, {-xy * s, (x2 + z2) * s, -yz * s}, {-zx * s, -yz * s, (x2 + y2) * s}};
        double J[3][3];
        for (int r = 0; r < 3; ++r)
          for (int c = 0; c < 3; ++c) J[r][c] = ((-1.0 * N[r][0]) * R.m[0][c] + (-1.0 * N[r][1]) * R.m[1][c]) + (-1.0 * N[r][2]) * R.m[2][c];
        const double n = std::sqrt(x2 + y2 + z2);
        const double e[3] = {o.f.x - p.x / n, o.f.y - p.y / n, o.f.z - p.z / n};
        for (int r = 0; r < 3; ++r) {
          for (int c = 0; c < 3; ++c) A[r][c] += (J[0][r] * J[0][c] + J[1][r] * J[1][c]) + J[2][r] * J[2][c];
          b[r] -= (J[0][r] * e[0] + J[1][r] * e[1]) + J[2][r] * e[2];
        }
        new_chi2 += (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2];
      }
    }
    double dp[3];
    ldltSolve<3>(A, b, dp);  // A.ldlt().solve(b)
    if ((it > 0 && new_chi2 > chi2) || std::isnan(dp[0])) {
      pos = old_point;  // roll-back
      break;
    }
    old_point = pos;
    pos = V3{pos.x + dp[0], pos.y + dp[1], pos.z + dp[2]};
    chi2 = new_chi2;
    if (std::max(std::fabs(dp[0]), std::max(std::fabs(dp[1]), std::fabs(dp[2]))) <= eps) break;  // vk::norm_max(dp) <= eps
  }
  return iters;
}

}  // namespace orc
