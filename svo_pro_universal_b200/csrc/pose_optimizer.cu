// (f4) PoseOptimizer: Gauss-Newton on the reprojection residuals of a frame bundle, batched over B independent bundles.
//
// ref: src/svo/src/pose_optimizer.cpp:39-94 (run), :96-196 (evaluateErrorImpl), :198-298 (removeOutliers), :300-334 (update,
//      applyPrior), :338-629 (the six residual / Jacobian functions)
//      src/vikit/vikit_solver/include/vikit/solver/implementation/mini_least_squares_solver.hpp:42-107 (Gauss-Newton driver),
//      :253-262 (dx = H.ldlt().solve(g)); src/vikit/vikit_solver/src/robust_cost.cpp:19-26 (MAD scale), :44-60 (Tukey, b = 4.6851)
//      src/svo_common/include/svo/common/frame.h:342-397 (xyz -> uv / image / bearing Jacobians w.r.t. the IMU pose)
//
// One 128-thread CTA per bundle runs the WHOLE run(): start errors -> median (rank counting in shared memory) -> measurement
// sigma -> every Gauss-Newton iteration (threads walk the bundle's features, 21 + 6 + 1 partial sums per thread, warp-shuffle +
// shared-memory block reduction, thread 0 adds the prior, solves the 6x6 system with the pivoted LDL^T of Eigen and updates
// the pose) -> outlier flags and the final median. No host round trip; all arithmetic FP64 like the reference, the Tukey
// weights and the MAD scale in float as in the reference's signatures.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxFeat = 2048;   // features per bundle (shared-memory error buffer for the two medians)
constexpr int kNS = 28;          // 21 upper-triangle H + 6 g + chi2

enum { kUnitPlane = 0, kBearingVectorDiff = 1, kImagePlane = 2 };
enum { kEdgeletSeed = 0, kEdgeletSeedConverged = 3, kEdgelet = 6 };
SVO_D bool isEdgeletT(int t) { return t == kEdgelet || t == kEdgeletSeed || t == kEdgeletSeedConverged; }

struct PoseOptParams {
  int n_cams, B;
  svo_camera cams[SVO_MAX_CAMS];
  double T_cam_imu[SVO_MAX_CAMS][7];
  svo_pose_optimizer_options opt;
  const double* T_imu_world;
  const int* feat_begin;
  const svo_feature* ftrs;
  const int* feat_cam;
  const double* xyz_world;
  const uint8_t* has_xyz;
  const double* prior_q;
  svo_pose_opt_result* results;
  uint8_t* outlier;
};

SVO_D float tukeyWeight(float error) {  // robust_cost.cpp:48-60 with b = 4.6851f
  const float b_square = 4.6851f * 4.6851f;
  const float x_square = error * error;
  if (x_square <= b_square) {
    const float tmp = 1.0f - x_square / b_square;
    return tmp * tmp;
  }
  return 0.0f;
}

// out[R][C] = A[R][K] * B[K][C], products summed in ascending k (Eigen's coefficient-based product)
template <int R, int K, int C>
SVO_D void matMul(const double (&A)[R][K], const double (&Bm)[K][C], double (&out)[R][C]) {
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int c = 0; c < C; ++c) {
      double s = A[r][0] * Bm[0][c];
#pragma unroll
      for (int k = 1; k < K; ++k) s += A[r][k] * Bm[k][c];
      out[r][c] = s;
    }
}
SVO_D void generators(const V3d& p, double (&G)[3][6]) {  // G_x = [I | -skew(p_in_imu)]
  const double S[3][3] = {{0, -p.z, p.y}, {p.z, 0, -p.x}, {-p.y, p.x, 0}};
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { G[r][c] = r == c ? 1.0 : 0.0; G[r][3 + c] = -S[r][c]; }
}
struct CamPose { svo_camera cam; SE3d T; M3d R; };

SVO_D void jacUv(const CamPose& cp, const V3d& p_imu, const V3d& p, double (&J)[2][6]) {  // frame.h:342-357
  double G[3][6];
  generators(p_imu, G);
  const double s = -1.0 / p.z;
  const double sJ[2][3] = {{s * 1, s * 0, s * (-p.x / p.z)}, {s * 0, s * 1, s * (-p.y / p.z)}};
  double sJR[2][3];
  matMul(sJ, cp.R.m, sJR);
  matMul(sJR, G, J);
}
SVO_D void jacImg(const CamPose& cp, const V3d& p_imu, const double (&J_cam)[2][3], double (&J)[2][6]) {  // frame.h:360-371
  double G[3][6], JR[2][3];
  generators(p_imu, G);
  matMul(J_cam, cp.R.m, JR);
  matMul(JR, G, J);
}
SVO_D void jacBearing(const CamPose& cp, const V3d& p_imu, const V3d& p, double (&J)[3][6]) {  // frame.h:374-397
  double G[3][6], JR[3][3];
  generators(p_imu, G);
  const double x2 = p.x * p.x, y2 = p.y * p.y, z2 = p.z * p.z, xy = p.x * p.y, yz = p.y * p.z, zx = p.z * p.x;
  double Jn[3][3] = {{y2 + z2, -xy, -zx}, {-xy, x2 + z2, -yz}, {-zx, -yz, x2 + y2}};
  const double s = 1 / pow(x2 + y2 + z2, 1.5);
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) Jn[r][c] *= s;
  matMul(Jn, cp.R.m, JR);
  matMul(JR, G, J);
}

// acc[0..20] upper triangle of H (row-major a <= b), acc[21..26] g, acc[27] chi2
template <int DIM>
SVO_D void accumulate(const double (&J)[DIM][6], const double (&e)[DIM], double weight, double (&acc)[kNS]) {
  int idx = 0;
#pragma unroll
  for (int a = 0; a < 6; ++a)
#pragma unroll
    for (int b = a; b < 6; ++b) {
      double s = J[0][a] * J[0][b];
#pragma unroll
      for (int k = 1; k < DIM; ++k) s += J[k][a] * J[k][b];
      acc[idx++] += s * weight;
    }
#pragma unroll
  for (int a = 0; a < 6; ++a) {
    double s = J[0][a] * e[0];
#pragma unroll
    for (int k = 1; k < DIM; ++k) s += J[k][a] * e[k];
    acc[21 + a] -= s * weight;
  }
}

// One feature (pose_optimizer.cpp:338-629). JAC = false: only the unwhitened error (start errors, removeOutliers).
template <bool JAC>
SVO_D double residual(const svo_feature& ft, const V3d& xyz_world, const CamPose& cp, const SE3d& T_imu_world, int err_type,
                      double measurement_sigma, double (&acc)[kNS]) {
  const V3d p_imu = se3Apply(T_imu_world, xyz_world);
  const V3d p = se3Apply(cp.T, p_imu);
  const bool edgelet = isEdgeletT(ft.type);
  const double R = 1.0 / measurement_sigma;
  const V3d f{ft.f[0], ft.f[1], ft.f[2]};
  double ue;
  if (!edgelet && err_type == kUnitPlane) {
    double e[2] = {f.x / f.z - p.x / p.z, f.y / f.z - p.y / p.z};
    ue = sqrt(e[0] * e[0] + e[1] * e[1]);
    if (JAC) {
      e[0] *= R; e[1] *= R;
      const double weight = tukeyWeight((float)sqrt(e[0] * e[0] + e[1] * e[1]));
      acc[27] += 0.5 * (e[0] * e[0] + e[1] * e[1]) * weight;
      double J[2][6];
      jacUv(cp, p_imu, p, J);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) J[r][c] *= R;
      accumulate(J, e, weight, acc);
    }
  } else if (!edgelet && err_type == kImagePlane) {
    const V2d px_est = camProject3(cp.cam, p);
    double e[2] = {ft.px[0] - px_est.x, ft.px[1] - px_est.y};
    ue = sqrt(e[0] * e[0] + e[1] * e[1]);
    if (JAC) {
      e[0] *= R; e[1] *= R;
      const double weight = tukeyWeight((float)sqrt(e[0] * e[0] + e[1] * e[1]));
      acc[27] += 0.5 * (e[0] * e[0] + e[1] * e[1]) * weight;
      double J_cam[2][3], J[2][6];
      camProject3Jac(cp.cam, p, J_cam);
      jacImg(cp, p_imu, J_cam, J);
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) J[r][c] = ((-1.0) * J[r][c]) * R;
      accumulate(J, e, weight, acc);
    }
  } else if (!edgelet) {  // kBearingVectorDiff
    const V3d fe = normalized3(p);
    double e[3] = {f.x - fe.x, f.y - fe.y, f.z - fe.z};
    ue = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]);
    if (JAC) {
      e[0] *= R; e[1] *= R; e[2] *= R;
      const double weight = tukeyWeight((float)sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2]));
      acc[27] += 0.5 * (e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) * weight;
      double J[3][6];
      jacBearing(cp, p_imu, p, J);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 6; ++c) J[r][c] = ((-1.0) * J[r][c]) * R;
      accumulate(J, e, weight, acc);
    }
  } else if (err_type == kUnitPlane) {
    double e = ft.grad[0] * (f.x / f.z - p.x / p.z) + ft.grad[1] * (f.y / f.z - p.y / p.z);
    ue = fabs(e);
    if (JAC) {
      e *= R;
      const double weight = tukeyWeight((float)e);
      acc[27] += 0.5 * e * e * weight;
      double Jp[2][6], J[1][6];
      jacUv(cp, p_imu, p, Jp);
#pragma unroll
      for (int c = 0; c < 6; ++c) J[0][c] = (ft.grad[0] * Jp[0][c] + ft.grad[1] * Jp[1][c]) * R;
      const double ee[1] = {e};
      accumulate(J, ee, weight, acc);
    }
  } else if (err_type == kImagePlane) {
    const V2d px_est = camProject3(cp.cam, p);
    double e = ft.grad[0] * (ft.px[0] - px_est.x) + ft.grad[1] * (ft.px[1] - px_est.y);
    ue = fabs(e);
    if (JAC) {
      e *= R;
      const double weight = tukeyWeight((float)e);
      acc[27] += 0.5 * e * e * weight;
      double J_cam[2][3], Jp[2][6], J[1][6];
      camProject3Jac(cp.cam, p, J_cam);
      jacImg(cp, p_imu, J_cam, Jp);
#pragma unroll
      for (int c = 0; c < 6; ++c) J[0][c] = ((ft.grad[0] * (-1.0)) * Jp[0][c] + (ft.grad[1] * (-1.0)) * Jp[1][c]) * R;
      const double ee[1] = {e};
      accumulate(J, ee, weight, acc);
    }
  } else {  // edgelet, kBearingVectorDiff (pose_optimizer.cpp:560-627)
    const V2d px_est = camProject3(cp.cam, p);
    const double pd[2] = {ft.px[0] - px_est.x, ft.px[1] - px_est.y};
    const double pd2 = pd[0] * pd[0] + pd[1] * pd[1];
    const V3d fe = normalized3(p);
    const double fd[3] = {f.x - fe.x, f.y - fe.y, f.z - fe.z};
    const double fd2 = fd[0] * fd[0] + fd[1] * fd[1] + fd[2] * fd[2];
    const double e_img = ft.grad[0] * pd[0] + ft.grad[1] * pd[1];
    const double scale_ratio = sqrt(fd2) / sqrt(pd2);
    double e = e_img * scale_ratio;
    ue = fabs(e);
    if (JAC) {
      e *= R;
      const double weight = tukeyWeight((float)e);
      acc[27] += 0.5 * e * e * weight;
      double J_cam[2][3], Jp[2][6], Jb[3][6], J[1][6];
      camProject3Jac(cp.cam, p, J_cam);
      jacImg(cp, p_imu, J_cam, Jp);
      jacBearing(cp, p_imu, p, Jb);
      const double k = (0.5) * (1.0 / (scale_ratio)) * (1 / (pd2 * pd2));
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double J_img = (ft.grad[0] * (-1.0)) * Jp[0][c] + (ft.grad[1] * (-1.0)) * Jp[1][c];
        const double J_ftf = ((2 * fd[0]) * (-1.0)) * Jb[0][c] + ((2 * fd[1]) * (-1.0)) * Jb[1][c] + ((2 * fd[2]) * (-1.0)) * Jb[2][c];
        const double J_ptp = ((2 * pd[0]) * (-1.0)) * Jp[0][c] + ((2 * pd[1]) * (-1.0)) * Jp[1][c];
        const double J_ratio = k * (J_ftf * pd2 - J_ptp * fd2);
        J[0][c] = (e_img * J_ratio + scale_ratio * J_img) * R;
      }
      const double ee[1] = {e};
      accumulate(J, ee, weight, acc);
    }
  }
  return ue;
}

// dx = H.ldlt().solve(g): Eigen's pivoted in-place LDL^T (Eigen/src/Cholesky/LDLT.h) on the 6x6 system, zero pivots -> 0
SVO_D void ldltSolve6(const double (&Hin)[6][6], const double (&g)[6], double (&dx)[6]) {
  constexpr int D = 6;
  double A[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) A[i][j] = Hin[i][j];
  int transp[D];
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

// k-th smallest (k = floor(m / 2)) of s_err[0..m): the element whose rank (ties broken by index) equals k
template <class T>
SVO_D void blockMedian(const T* s_err, int m, T* s_out) {
  const int k = m / 2;
  for (int i = threadIdx.x; i < m; i += kThreads) {
    const T v = s_err[i];
    int rank = 0;
    for (int j = 0; j < m; ++j) {
      const T w = s_err[j];
      rank += (w < v) || (w == v && j < i);
    }
    if (rank == k) *s_out = v;
  }
}

// Resident CTAs per SM the register allocation is held to. Measured on the B200 (8192 two-camera bundles of the front-end chain):
// 1 (202 registers, 2 CTAs / SM) 2.52 ms, 3 (168) 2.20 ms, 4 (128 registers, 4 CTAs / SM) 2.03 ms, 5 (96) 2.15 ms.
#ifndef SVO_POSEOPT_MINB
#define SVO_POSEOPT_MINB 4
#endif
__global__ void __launch_bounds__(kThreads, SVO_POSEOPT_MINB) pose_optimize_kernel(const PoseOptParams P) {
  __shared__ float s_err[kMaxFeat];     // start errors (float, as the reference's std::vector<float>)
  __shared__ double s_errd[kMaxFeat];   // final errors (double)
  __shared__ double s_mediand;
  __shared__ double s_red[kWarps][kNS];
  __shared__ CamPose s_cam[SVO_MAX_CAMS];
  __shared__ SE3d s_T, s_T_old, s_T_iw[SVO_MAX_CAMS];
  __shared__ double s_sigma, s_I_prior[6], s_chi2;
  __shared__ float s_median;
  __shared__ int s_m, s_brk, s_iter, s_stop, s_deleted;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f0 = P.feat_begin[b];
  const int n = min(P.feat_begin[b + 1] - f0, kMaxFeat);
  const int err_type = P.opt.err_type;
  if (tid < P.n_cams) {
    s_cam[tid].cam = P.cams[tid];
    s_cam[tid].T = se3Load(P.T_cam_imu[tid]);
    s_cam[tid].R = quatToMatrix(s_cam[tid].T.q);
  }
  if (tid == 0) { s_T = se3Load(P.T_imu_world + 7 * (size_t)b); s_T_old = s_T; s_m = 0; s_iter = 0; s_stop = 0; s_chi2 = 1e10; s_deleted = 0; s_median = 0.f; }
  __syncthreads();
  double acc[kNS];

  // ---- start errors -> MAD scale (pose_optimizer.cpp:48-53)
  for (int i = tid; i < n; i += kThreads) {
    if (!P.has_xyz[f0 + i]) continue;
    const svo_feature ft = P.ftrs[f0 + i];
    const int c = P.feat_cam ? P.feat_cam[f0 + i] : 0;
    const V3d X{P.xyz_world[3 * (size_t)(f0 + i)], P.xyz_world[3 * (size_t)(f0 + i) + 1], P.xyz_world[3 * (size_t)(f0 + i) + 2]};
    const double ue = residual<false>(ft, X, s_cam[c], s_T, err_type, 1.0, acc);
    s_err[atomicAdd(&s_m, 1)] = (float)(ue / (1 << ft.level));
  }
  __syncthreads();
  const int m = s_m;
  if (m > 0) blockMedian(s_err, m, &s_median);
  __syncthreads();
  const float err_before = s_median;
  if (tid == 0) s_sigma = (double)(1.48f * s_median);
  __syncthreads();
  const double sigma = s_sigma;

  // ---- optimizeGaussNewton (mini_least_squares_solver.hpp:42-107)
  const int max_iter = P.opt.max_iter;
  for (int iter = 0; iter < max_iter && m > 0; ++iter) {
#pragma unroll
    for (int k = 0; k < kNS; ++k) acc[k] = 0.0;
    const SE3d T = s_T;
    for (int i = tid; i < n; i += kThreads) {
      if (!P.has_xyz[f0 + i]) continue;
      const svo_feature ft = P.ftrs[f0 + i];
      const int c = P.feat_cam ? P.feat_cam[f0 + i] : 0;
      const V3d X{P.xyz_world[3 * (size_t)(f0 + i)], P.xyz_world[3 * (size_t)(f0 + i) + 1], P.xyz_world[3 * (size_t)(f0 + i) + 2]};
      double ms = sigma * (1 << ft.level);
      if (isEdgeletT(ft.type)) ms *= 2.0;  // kEdgeletSigmaExtraFactor
      residual<true>(ft, X, s_cam[c], T, err_type, ms, acc);
    }
#pragma unroll
    for (int k = 0; k < kNS; ++k) {
      double v = acc[k];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      if (lane == 0) s_red[warp][k] = v;
    }
    __syncthreads();
    if (tid == 0) {
      double tot[kNS];
      for (int k = 0; k < kNS; ++k) { double v = 0.0; for (int w = 0; w < kWarps; ++w) v += s_red[w][k]; tot[k] = v; }
      double H[6][6], g[6], dx[6];
      int idx = 0;
      for (int a = 0; a < 6; ++a) for (int c2 = a; c2 < 6; ++c2) { H[a][c2] = tot[idx]; H[c2][a] = tot[idx]; ++idx; }
      for (int a = 0; a < 6; ++a) g[a] = tot[21 + a];
      const double new_chi2 = tot[27];
      if (P.prior_q) {  // applyPrior (pose_optimizer.cpp:311-334): prior_ = (R_frame_world, 0), information on the rotation block
        if (iter == 0) {
          double H_max_diag = 0;
          for (int j = 3; j < 6; ++j) H_max_diag = fmax(H_max_diag, fabs(H[j][j]));
          for (int j = 0; j < 6; ++j) s_I_prior[j] = (j >= 3 ? 1.0 : 0.0) * (H_max_diag * P.opt.prior_lambda);
        }
        SE3d prior;
        prior.q = Quatd{P.prior_q[4 * (size_t)b], P.prior_q[4 * (size_t)b + 1], P.prior_q[4 * (size_t)b + 2], P.prior_q[4 * (size_t)b + 3]};
        prior.t = V3d{0, 0, 0};
        const SE3d E = se3Mul(s_T, se3Inv(prior));
        const V3d lr = quatLog(E.q);
        const double l[6] = {E.t.x, E.t.y, E.t.z, lr.x, lr.y, lr.z};
        for (int j = 0; j < 6; ++j) { H[j][j] += s_I_prior[j]; g[j] -= s_I_prior[j] * l[j]; }
      }
      ldltSolve6(H, g, dx);
      int brk = 0;
      if (dx[0] != dx[0]) {  // solve() failed: stop, roll back
        s_stop = 1;
        s_T = s_T_old;
        brk = 1;
      } else {
        // update (pose_optimizer.cpp:300-309): T_new = exp(dx) * T_old, quaternion normalised
        SE3d inc;
        inc.q = quatExp(V3d{dx[3], dx[4], dx[5]});
        inc.t = V3d{dx[0], dx[1], dx[2]};
        SE3d Tn = se3Mul(inc, s_T);
        const double nq = sqrt(Tn.q.w * Tn.q.w + Tn.q.x * Tn.q.x + Tn.q.y * Tn.q.y + Tn.q.z * Tn.q.z);
        Tn.q = Quatd{Tn.q.w / nq, Tn.q.x / nq, Tn.q.y / nq, Tn.q.z / nq};
        s_T_old = s_T;
        s_T = Tn;
        s_chi2 = new_chi2;
        double x_norm = 0.0;
        for (int j = 0; j < 6; ++j) x_norm = fmax(x_norm, fabs(dx[j]));
        if (x_norm < P.opt.eps) brk = 1;
      }
      s_iter = iter + (brk ? 0 : 1);  // iter_ keeps the index of the iteration that broke out, max_iter when the loop ran out
      s_brk = brk;
    }
    __syncthreads();
    if (s_brk) break;
  }

  // ---- frames' poses, removeOutliers (pose_optimizer.cpp:58-70, 198-298), statistics
  if (tid < P.n_cams) {
    const SE3d T_f_w = se3Mul(s_cam[tid].T, s_T);
    se3Store(T_f_w, P.results[b].T_f_w[tid]);
    s_T_iw[tid] = se3Mul(se3Inv(s_cam[tid].T), T_f_w);  // frame->T_imu_world() = T_imu_cam * T_f_w_
  } else if (tid < SVO_MAX_CAMS) {
    for (int k = 0; k < 7; ++k) P.results[b].T_f_w[tid][k] = 0.0;
  }
  if (tid == 0) s_m = 0;
  __syncthreads();
  const double focal = fabs(s_cam[0].cam.fx);  // getErrorMultiplier() of the first camera (pinhole_projection.hpp:66-70)
  double thr = P.opt.reproj_thresh_px;
  if (err_type == kUnitPlane) thr = P.opt.reproj_thresh_px / focal;
  else if (err_type == kBearingVectorDiff) thr = fabs(2 * sin(0.5 * camAngleError(s_cam[0].cam, P.opt.reproj_thresh_px)));
  int deleted = 0;
  for (int i = tid; i < P.feat_begin[b + 1] - f0; i += kThreads) {
    uint8_t out = 0;
    if (i < n && P.has_xyz[f0 + i]) {
      const svo_feature ft = P.ftrs[f0 + i];
      const int c = P.feat_cam ? P.feat_cam[f0 + i] : 0;
      const V3d X{P.xyz_world[3 * (size_t)(f0 + i)], P.xyz_world[3 * (size_t)(f0 + i) + 1], P.xyz_world[3 * (size_t)(f0 + i) + 2]};
      double ue = residual<false>(ft, X, s_cam[c], s_T_iw[c], err_type, 1.0, acc);
      ue *= 1.0 / (1 << ft.level);
      s_errd[atomicAdd(&s_m, 1)] = ue;
      if (fabs(ue) > thr) { out = 1; ++deleted; }
    }
    P.outlier[f0 + i] = out;
  }
  if (deleted) atomicAdd(&s_deleted, deleted);
  __syncthreads();
  if (tid == 0) s_mediand = 0.0;
  __syncthreads();
  if (s_m > 0) blockMedian(s_errd, s_m, &s_mediand);
  __syncthreads();
  if (tid == 0) {
    svo_pose_opt_result& r = P.results[b];
    se3Store(s_T, r.T_imu_world);
    const double error_scale = err_type == kUnitPlane ? focal : 1.0;
    r.measurement_sigma = sigma;
    r.reproj_error_before = (double)err_before * error_scale;
    r.reproj_error_after = s_mediand * error_scale;
    r.chi2 = s_chi2;
    r.n_meas = m;
    r.n_meas_final = m - s_deleted;
    r.iters = s_iter;
    r.stop = s_stop;
  }
}

}  // namespace

extern "C" int svo_cuda_pose_optimize(svo_cuda_ctx* ctx, int n_cams, const svo_camera* cams, const double* T_cam_imu, int B,
                                      const double* T_imu_world, const int* feat_begin, int n_features, const svo_feature* ftrs,
                                      const int* feat_cam, const double* xyz_world, const uint8_t* has_xyz, const double* prior_q,
                                      const svo_pose_optimizer_options* opt, svo_pose_opt_result* results, uint8_t* outlier,
                                      svo_mem mem) {
  if (!ctx || n_cams < 1 || n_cams > SVO_MAX_CAMS || !cams || !T_cam_imu || B < 0 || !T_imu_world || !feat_begin || n_features < 0 ||
      (n_features > 0 && (!ftrs || !xyz_world || !has_xyz || !outlier)) || !opt || !results)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pose_optimize: bad arguments");
  if (opt->err_type < 0 || opt->err_type > 2 || opt->max_iter < 0)
    return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pose_optimize: bad error type / max_iter");
  if (B == 0) return SVO_OK;
  if (mem == SVO_MEM_HOST) {
    if (feat_begin[B] != n_features) return SVO_FAIL(ctx, SVO_ERR_INVALID_ARG, "svo_cuda_pose_optimize: n_features != feat_begin[B]");
    for (int b = 0; b < B; ++b)
      if (feat_begin[b + 1] - feat_begin[b] > kMaxFeat || feat_begin[b + 1] < feat_begin[b])
        return SVO_FAIL(ctx, SVO_ERR_TOO_MANY_FEATURES, "svo_cuda_pose_optimize: more than 2048 features in one bundle");
  }
  cudaSetDevice(ctx->device);
  PoseOptParams P;
  memset(&P, 0, sizeof(P));
  P.n_cams = n_cams; P.B = B;
  for (int c = 0; c < n_cams; ++c) {
    P.cams[c] = cams[c];
    for (int k = 0; k < 7; ++k) P.T_cam_imu[c][k] = T_cam_imu[7 * c + k];
  }
  P.opt = *opt;
  Stager st(ctx, mem);
  P.T_imu_world = st.in(T_imu_world, (size_t)B * 7);
  P.feat_begin = st.in(feat_begin, (size_t)B + 1);
  P.ftrs = st.in(ftrs, (size_t)n_features);
  P.feat_cam = st.in(feat_cam, (size_t)n_features);
  P.xyz_world = st.in(xyz_world, (size_t)n_features * 3);
  P.has_xyz = st.in(has_xyz, (size_t)n_features);
  P.prior_q = st.in(prior_q, (size_t)B * 4);
  P.results = st.out(results, (size_t)B);
  P.outlier = st.out(outlier, (size_t)n_features);
  if (!st.send()) return st.finish();
  pose_optimize_kernel<<<B, kThreads, 0, ctx->stream>>>(P);
  SVO_LAUNCH_CHECK(ctx);
  return st.finish();
}
