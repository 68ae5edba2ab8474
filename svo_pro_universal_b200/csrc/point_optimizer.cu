// (f4, second half) svo::Point::optimize — Gauss-Newton refinement of 3-D points from their observations, one thread per point.
//
// ref: src/svo_common/src/point.cpp:216-325 (updateHessianGradientUnitPlane / UnitSphere, optimize),
//      src/svo_common/include/svo/common/point.h:170-204 (jacobian_xyz2uv, jacobian_xyz2f); caller:
//      FrameHandlerBase::optimizeStructure, src/svo/src/frame_handler_base.cpp:785-825 (the landmarks of a frame, max_iter 5).
//
// A point has 2-20 observations and a 3x3 normal system: there is nothing to share between threads, so a thread walks its
// point's observation list (frame pose gathered through the read-only cache, 7 doubles shared by all points of a frame), keeps
// A, b in registers and solves with Eigen's pivoted LDL^T. Arithmetic follows the reference operation by operation (this file is
// compiled without FMA contraction): results are bit-equal to the reference's compiled point.cpp in the tests.
#include "common.cuh"

namespace {

// dx = A.ldlt().solve(b) for 3x3: Eigen's pivoted in-place LDL^T (Eigen/src/Cholesky/LDLT.h), zero pivots -> 0
SVO_D void ldltSolve3(const double (&Hin)[3][3], const double (&g)[3], double (&dx)[3]) {
  constexpr int D = 3;
  double A[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) A[i][j] = Hin[i][j];
  int transp[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    int piv = k;
    double big = fabs(A[k][k]);
    for (int i = k + 1; i < D; ++i) if (fabs(A[i][i]) > big) { big = fabs(A[i][i]); piv = i; }
    transp[k] = piv;
    if (piv != k) {
      for (int j = 0; j < k; ++j) { const double t = A[k][j]; A[k][j] = A[piv][j]; A[piv][j] = t; }
      for (int i = piv + 1; i < D; ++i) { const double t = A[i][k]; A[i][k] = A[i][piv]; A[i][piv] = t; }
      { const double t = A[k][k]; A[k][k] = A[piv][piv]; A[piv][piv] = t; }
      for (int i = k + 1; i < piv; ++i) { const double t = A[i][k]; A[i][k] = A[piv][i]; A[piv][i] = t; }
    }
    double temp[D];
    for (int j = 0; j < k; ++j) temp[j] = A[j][j] * A[k][j];
    for (int j = 0; j < k; ++j) A[k][k] -= A[k][j] * temp[j];
    for (int i = k + 1; i < D; ++i) for (int j = 0; j < k; ++j) A[i][k] -= A[i][j] * temp[j];
    const double akk = A[k][k];
    if (fabs(akk) > 0.0) for (int i = k + 1; i < D; ++i) A[i][k] /= akk;
  }
  double x[D];
  for (int i = 0; i < D; ++i) x[i] = g[i];
  for (int k = 0; k < D; ++k) { const double t = x[k]; x[k] = x[transp[k]]; x[transp[k]] = t; }
  for (int i = 0; i < D; ++i) for (int j = 0; j < i; ++j) x[i] -= A[i][j] * x[j];
  const double tolerance = 1.0 / 1.7976931348623157e308;
  for (int i = 0; i < D; ++i) { if (fabs(A[i][i]) > tolerance) x[i] /= A[i][i]; else x[i] = 0.0; }
  for (int i = D - 1; i >= 0; --i) for (int j = i + 1; j < D; ++j) x[i] -= A[j][i] * x[j];
  for (int k = D - 1; k >= 0; --k) { const double t = x[k]; x[k] = x[transp[k]]; x[transp[k]] = t; }
  for (int i = 0; i < D; ++i) dx[i] = x[i];
}

__global__ void __launch_bounds__(128) optimize_points_kernel(int P, double* __restrict__ pos, const int* __restrict__ obs_begin,
                                                             const int* __restrict__ obs_frame, const double* __restrict__ obs_f,
                                                             const double* __restrict__ T_f_w, int n_iter, int sphere,
                                                             int* __restrict__ iters_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const int lo = obs_begin[i], hi = obs_begin[i + 1];
  V3d p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
  V3d old_point = p;
  double chi2 = 0.0;
  int iters = 0;
  if (hi - lo >= 2) {  // point.cpp:255-259
    for (int it = 0; it < n_iter; ++it) {
      ++iters;
      double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, b[3] = {0, 0, 0};
      double new_chi2 = 0.0;
      for (int o = lo; o < hi; ++o) {
        const SE3d T = se3Load(T_f_w + 7 * (size_t)obs_frame[o]);
        const V3d f{obs_f[3 * (size_t)o], obs_f[3 * (size_t)o + 1], obs_f[3 * (size_t)o + 2]};
        const V3d q = se3Apply(T, p);
        const M3d R = quatToMatrix(T.q);
        if (!sphere) {
          const double z_inv = 1.0 / q.z, z_inv_sq = z_inv * z_inv;
          const double J0[2][3] = {{z_inv, 0.0, -q.x * z_inv_sq}, {0.0, z_inv, -q.y * z_inv_sq}};
          double J[2][3];
#pragma unroll
          for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r][c] = -((J0[r][0] * R.m[0][c] + J0[r][1] * R.m[1][c]) + J0[r][2] * R.m[2][c]);
          const double e0 = f.x / f.z - q.x / q.z, e1 = f.y / f.z - q.y / q.z;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) A[r][c] += J[0][r] * J[0][c] + J[1][r] * J[1][c];
            b[r] -= J[0][r] * e0 + J[1][r] * e1;
          }
          new_chi2 += e0 * e0 + e1 * e1;
        } else {
          const double x2 = q.x * q.x, y2 = q.y * q.y, z2 = q.z * q.z, xy = q.x * q.y, yz = q.y * q.z, zx = q.z * q.x;
          const double s = 1.0 / pow(x2 + y2 + z2, 1.5);
          const double N[3][3] = {{(y2 + z2) * s, -xy * s, -zx * s}, {-xy * s, (x2 + z2) * s, -yz * s}, {-zx * s, -yz * s, (x2 + y2) * s}};
          double J[3][3];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) J[r][c] = ((-1.0 * N[r][0]) * R.m[0][c] + (-1.0 * N[r][1]) * R.m[1][c]) + (-1.0 * N[r][2]) * R.m[2][c];
          const double n = sqrt(x2 + y2 + z2);
          const double e0 = f.x - q.x / n, e1 = f.y - q.y / n, e2 = f.z - q.z / n;
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) A[r][c] += (J[0][r] * J[0][c] + J[1][r] * J[1][c]) + J[2][r] * J[2][c];
            b[r] -= (J[0][r] * e0 + J[1][r] * e1) + J[2][r] * e2;
          }
          new_chi2 += (e0 * e0 + e1 * e1) + e2 * e2;
        }
      }
      double dp[3];
      ldltSolve3(A, b, dp);
      if ((it > 0 && new_chi2 > chi2) || isnan(dp[0])) { p = old_point; break; }  // roll-back
      old_point = p;
      p = V3d{p.x + dp[0], p.y + dp[1], p.z + dp[2]};
      chi2 = new_chi2;
      if (fmax(fabs(dp[0]), fmax(fabs(dp[1]), fabs(dp[2]))) <= 0.0000000001) break;
    }
  }
  pos[3 * i] = p.x; pos[3 * i + 1] = p.y; pos[3 * i + 2] = p.z;
  if (iters_out) iters_out[i] = iters;
}

}  // namespace

extern "C" int svo_cuda_optimize_points(svo_cuda_ctx* ctx, int P, double* pos, const int* obs_begin, int n_obs, const int* obs_frame,
                                        const double* obs_f, int n_frames, const double* T_f_w, int n_iter, int using_bearing_vector,
                                        int* iters_out, svo_mem mem) {
  if (!ctx || P < 0 || n_obs < 0 || n_frames < 0 || n_iter < 0 || (P > 0 && (!pos || !obs_begin)) || (n_obs > 0 && (!obs_frame || !obs_f || !T_f_w)))
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_optimize_points: bad arguments");
  if (P == 0) return SVO_OK;
  cudaSetDevice(ctx->device);
  Stager st(ctx, mem);
  double* d_pos = st.inout(pos, (size_t)P * 3);
  const int* d_begin = st.in(obs_begin, (size_t)P + 1);
  const int* d_frame = st.in(obs_frame, (size_t)n_obs);
  const double* d_f = st.in(obs_f, (size_t)n_obs * 3);
  const double* d_T = st.in(T_f_w, (size_t)n_frames * 7);
  int* d_it = st.out(iters_out, (size_t)P);
  if (!st.send()) return st.finish();
  optimize_points_kernel<<<(P + 127) / 128, 128, 0, ctx->stream>>>(P, d_pos, d_begin, d_frame, d_f, d_T, n_iter, using_bearing_vector, d_it);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}
