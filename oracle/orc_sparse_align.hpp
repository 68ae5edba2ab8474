// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Rows b1-b9 of SURVEY.md §8: svo::SparseImgAlign (inverse-compositional sparse image alignment)
// with the vk::solver::MiniLeastSquaresSolver Gauss-Newton driver.
// Parity status: "parity unpinned" (restatement; the reference needs Eigen/OpenCV/glog to build and
// its tests hold no vectors for this path).
#pragma once
#include <vector>
#include <cmath>
#include <cstdio>
#include "orc_math.hpp"

namespace orc {

// ref: src/svo_img_align/include/svo/img_align/sparse_img_align_base.h:37-46
struct SparseImgAlignOptions {
  int max_level = 4;
  int min_level = 1;
  bool estimate_illumination_gain = false;
  bool estimate_illumination_offset = false;
  bool use_distortion_jacobian = false;
  bool robustification = false;
  double weight_scale = 10;
};
// ref: src/vikit/vikit_solver/include/vikit/solver/mini_least_squares_solver.h:20-47 with the
// SparseImgAlign defaults of src/svo_img_align/src/sparse_img_align_base.cpp:35-42.
struct SolverOptions {
  size_t max_iter = 10;
  double eps = 0.0005;
  bool stop_when_error_increases = false;
};
// ref: sparse_img_align_base.h:49-55
struct SparseImgAlignState {
  SE3 T_icur_iref;
  double alpha = 0.0;
  double beta = 0.0;
};

// One camera of a frame bundle, reduced to what run() reads.
struct AlignFrame {
  std::vector<Img> img_pyr;   // frame.img_pyr_
  Camera cam;                 // frame.cam()
  SE3 T_cam_imu;              // frame.T_cam_imu(); T_imu_cam() is its inverse
  SE3 T_imu_world;            // frame.T_imu_world() (only cam 0 is read, sparse_img_align.cpp:62,75)
  // ref-frame features (px_vec_, f_vec_, depth of landmark/seed from the ref camera centre,
  // eligibility = (landmark || seed ref) && !isMapPoint(type), sparse_img_align.cpp:242-248):
  std::vector<V2> px;
  std::vector<V3> f;
  std::vector<double> depth;
  std::vector<uint8_t> eligible;
};

// ref: src/vikit/vikit_solver/src/robust_cost.cpp:44-60, robust_cost.h:67-75 (b = 4.6851f)
struct TukeyWeightFunction {
  float b_square_;
  explicit TukeyWeightFunction(const float b = 4.6851f) : b_square_(b * b) {}
  float weight(const float& error) const {
    const float x_square = error * error;
    if (x_square <= b_square_) {
      const float tmp = 1.0f - x_square / b_square_;
      return tmp * tmp;
    }
    return 0.0f;
  }
};

// 8x8 symmetric solve following Eigen::LDLT (pivoted, in place, lower) + solve().
// ref: src/vikit/vikit_solver/include/vikit/solver/implementation/mini_least_squares_solver.hpp:253-262
// (dx = H.ldlt().solve(g)); algorithm per Eigen/src/Cholesky/LDLT.h (ldlt_inplace<Lower>::unblocked,
// LDLT::_solve_impl): symmetric diagonal pivoting on max |A_kk|, zero pivots -> solution component 0.
template <int D>
inline void ldltSolve(const double Hin[D][D], const double g[D], double dx[D]) {
  double A[D][D];
  for (int i = 0; i < D; ++i) for (int j = 0; j < D; ++j) A[i][j] = Hin[i][j];
  int transp[D];
  for (int k = 0; k < D; ++k) {
    int piv = k; double big = std::abs(A[k][k]);
    for (int i = k + 1; i < D; ++i) if (std::abs(A[i][i]) > big) { big = std::abs(A[i][i]); piv = i; }
    transp[k] = piv;
    if (piv != k) {
      // symmetric swap of rows/cols k and piv on the lower triangle (LDLT.h:298-311)
      for (int j = 0; j < k; ++j) std::swap(A[k][j], A[piv][j]);
      for (int i = piv + 1; i < D; ++i) std::swap(A[i][k], A[i][piv]);
      std::swap(A[k][k], A[piv][piv]);
      for (int i = k + 1; i < piv; ++i) { const double tmp = A[i][k]; A[i][k] = A[piv][i]; A[piv][i] = tmp; }
    }
    // A[k][k] -= A10 * D * A10^T ; A21 -= A20 * (D A10^T) ; A21 /= A[k][k]   (LDLT.h:320-338)
    double temp[D];
    for (int j = 0; j < k; ++j) temp[j] = A[j][j] * A[k][j];
    for (int j = 0; j < k; ++j) A[k][k] -= A[k][j] * temp[j];
    for (int i = k + 1; i < D; ++i) for (int j = 0; j < k; ++j) A[i][k] -= A[i][j] * temp[j];
    const double akk = A[k][k];
    if (std::abs(akk) > 0.0) for (int i = k + 1; i < D; ++i) A[i][k] /= akk;
  }
  // solve: dst = P b; L^-1; D^-1 (pseudo-inverse); L^-T; P^T   (LDLT.h:568-600)
  double x[D];
  for (int i = 0; i < D; ++i) x[i] = g[i];
  for (int k = 0; k < D; ++k) std::swap(x[k], x[transp[k]]);
  for (int i = 0; i < D; ++i) for (int j = 0; j < i; ++j) x[i] -= A[i][j] * x[j];
  const double tolerance = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < D; ++i) {
    if (std::abs(A[i][i]) > tolerance) x[i] /= A[i][i];
    else x[i] = 0.0;
  }
  for (int i = D - 1; i >= 0; --i) for (int j = i + 1; j < D; ++j) x[i] -= A[j][i] * x[j];
  for (int k = D - 1; k >= 0; --k) std::swap(x[k], x[transp[k]]);
  for (int i = 0; i < D; ++i) dx[i] = x[i];
}

struct SparseImgAlignResult {
  size_t n_fts_to_track = 0;
  SE3 T_icur_iref;
  double alpha = 0, beta = 0;
  double chi2 = 0;                 // getError()
  double H[8][8];                  // getHessian()
  std::vector<int> iters_per_level;  // number of evaluateError calls at each level (max -> min)
  std::vector<SE3> T_f_w;          // per cur frame
  bool stop = false;
};

class SparseImgAlign {
 public:
  typedef double FloatType;  // ref: src/svo_common/include/svo/common/types.h:16
  SolverOptions solver_options_;
  SparseImgAlignOptions options_;
  TukeyWeightFunction tukey_;
  float weight_scale_f_;  // passed as `const float weight_scale` (sparse_img_align.h:121)

  // solver state (mini_least_squares_solver.h:171-186)
  double H_[8][8]; double g_[8]; double dx_[8];
  bool have_prior_ = false;
  SparseImgAlignState prior_;
  double I_prior_[8][8];
  double chi2_ = 0.0;
  bool stop_ = false;
  size_t iter_ = 0;
  double prior_lambda_rot_ = 0, prior_lambda_trans_ = 0, prior_lambda_alpha_ = 0, prior_lambda_beta_ = 0;
  double alpha_init_ = 0.0, beta_init_ = 0.0;

  // caches (sparse_img_align.h:51-58), all FloatType = double
  int patch_size_ = 4, border_size_ = 1, patch_size_with_border_ = 6, patch_area_ = 16;
  std::vector<std::vector<size_t>> fts_vec_;
  std::vector<double> uv_cache_, xyz_ref_cache_, jacobian_proj_cache_, jacobian_cache_, residual_cache_, ref_patch_cache_;
  std::vector<uint8_t> visibility_mask_;
  bool have_cache_ = false;
  int level_ = 0;
  const std::vector<AlignFrame>* ref_frames_ = nullptr;
  const std::vector<AlignFrame>* cur_frames_ = nullptr;
  std::vector<int> iters_per_level_;

  SparseImgAlign(const SolverOptions& so, const SparseImgAlignOptions& o)
      : solver_options_(so), options_(o), weight_scale_f_(static_cast<float>(o.weight_scale)) {}

  // ref: mini_least_squares_solver.hpp:240-250
  void reset() {
    have_prior_ = false;
    chi2_ = 1e10;
    iter_ = 0;
    stop_ = false;
  }
  // ref: src/svo_img_align/src/sparse_img_align_base.cpp:44-62
  void setWeightedPrior(const SE3& T_cur_ref_prior, double alpha_prior, double beta_prior,
                        double lambda_rot, double lambda_trans, double lambda_alpha, double lambda_beta) {
    prior_lambda_rot_ = lambda_rot;
    prior_lambda_trans_ = lambda_trans;
    prior_lambda_alpha_ = lambda_alpha;
    prior_lambda_beta_ = lambda_beta;
    prior_.T_icur_iref = T_cur_ref_prior;
    prior_.alpha = alpha_prior;
    prior_.beta = beta_prior;
    have_prior_ = true;
    for (auto& r : I_prior_) for (double& v : r) v = 0.0;
  }

  // b2. ref: src/svo_img_align/src/sparse_img_align.cpp:209-260
  static void extractFeaturesSubset(const AlignFrame& ref_frame, int max_level, int patch_size_wb, std::vector<size_t>& fts) {
    const FloatType scale = 1.0f / (1 << max_level);
    const Img& ref_img = ref_frame.img_pyr.at(max_level);
    const int rows_minus_two = ref_img.rows - 2;
    const int cols_minus_two = ref_img.cols - 2;
    const FloatType patch_center_wb = (patch_size_wb - 1) / 2.0f;
    for (size_t i = 0; i < ref_frame.px.size(); ++i) {
      if (!ref_frame.eligible[i]) continue;
      const FloatType u_tl = ref_frame.px[i].x * scale - patch_center_wb;
      const FloatType v_tl = ref_frame.px[i].y * scale - patch_center_wb;
      const int u_tl_i = std::floor(u_tl);
      const int v_tl_i = std::floor(v_tl);
      if (!(u_tl_i < 0 || v_tl_i < 0 || u_tl_i + patch_size_wb >= cols_minus_two || v_tl_i + patch_size_wb >= rows_minus_two))
        fts.push_back(i);
    }
  }

  // vk::skew + Frame::jacobian_xyz2uv_imu — ref: src/svo_common/include/svo/common/frame.h:342-357
  static void jacobian_xyz2uv_imu(const SE3& T_cam_imu, const V3& p_in_imu, double J[2][6]) {
    double G[3][6] = {{1, 0, 0, 0, p_in_imu.z, -p_in_imu.y},
                      {0, 1, 0, -p_in_imu.z, 0, p_in_imu.x},
                      {0, 0, 1, p_in_imu.y, -p_in_imu.x, 0}};  // [I | -skew(p)]
    const V3 pc = T_cam_imu * p_in_imu;
    const double Jp[2][3] = {{1, 0, -pc.x / pc.z}, {0, 1, -pc.y / pc.z}};
    const M3 R = quatToMatrix(T_cam_imu.q);
    const double s = -1.0 / pc.z;
    double A[2][3], B[2][3];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c) A[r][c] = s * Jp[r][c];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c)
      B[r][c] = A[r][0] * R.m[0][c] + A[r][1] * R.m[1][c] + A[r][2] * R.m[2][c];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 6; ++c)
      J[r][c] = B[r][0] * G[0][c] + B[r][1] * G[1][c] + B[r][2] * G[2][c];
  }
  // Frame::jacobian_xyz2image_imu — ref: src/svo_common/src/frame.cpp:274-290
  static void jacobian_xyz2image_imu(const Camera& cam, const SE3& T_cam_imu, const V3& p_in_imu, double J[2][6]) {
    double G[3][6] = {{1, 0, 0, 0, p_in_imu.z, -p_in_imu.y},
                      {0, 1, 0, -p_in_imu.z, 0, p_in_imu.x},
                      {0, 0, 1, p_in_imu.y, -p_in_imu.x, 0}};
    const V3 pc = T_cam_imu * p_in_imu;
    double Jp[2][3];
    cam.project3(pc, Jp);
    const M3 R = quatToMatrix(T_cam_imu.q);
    double B[2][3];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 3; ++c)
      B[r][c] = Jp[r][0] * R.m[0][c] + Jp[r][1] * R.m[1][c] + Jp[r][2] * R.m[2][c];
    for (int r = 0; r < 2; ++r) for (int c = 0; c < 6; ++c)
      J[r][c] = B[r][0] * G[0][c] + B[r][1] * G[1][c] + B[r][2] * G[2][c];
  }

  // b3. ref: sparse_img_align.cpp:262-317
  void precomputeBaseCaches(const AlignFrame& ref_frame, const std::vector<size_t>& fts, bool use_distortion_jac, size_t& feature_counter) {
    const double focal_length = ref_frame.cam.errorMultiplier();
    const SE3 T_imu_cam = inverse(ref_frame.T_cam_imu);
    const SE3& T_cam_imu = ref_frame.T_cam_imu;
    for (const size_t i : fts) {
      uv_cache_[2 * feature_counter + 0] = ref_frame.px[i].x;
      uv_cache_[2 * feature_counter + 1] = ref_frame.px[i].y;
      const FloatType depth = ref_frame.depth[i];
      const V3 xyz_ref = ref_frame.f[i] * depth;
      xyz_ref_cache_[3 * feature_counter + 0] = xyz_ref.x;
      xyz_ref_cache_[3 * feature_counter + 1] = xyz_ref.y;
      xyz_ref_cache_[3 * feature_counter + 2] = xyz_ref.z;
      const V3 xyz_in_imu = T_imu_cam * xyz_ref;
      double frame_jac[2][6];
      if (!use_distortion_jac /* && camera type == pinhole: the only type this oracle models */) {
        jacobian_xyz2uv_imu(T_cam_imu, xyz_in_imu, frame_jac);
        for (auto& r : frame_jac) for (double& v : r) v *= focal_length;
      } else {
        jacobian_xyz2image_imu(ref_frame.cam, T_cam_imu, xyz_in_imu, frame_jac);
        for (auto& r : frame_jac) for (double& v : r) v *= (-1.0);
      }
      const size_t col_index = 2 * feature_counter;
      for (int c = 0; c < 6; ++c) {
        jacobian_proj_cache_[6 * col_index + c] = frame_jac[0][c];
        jacobian_proj_cache_[6 * (col_index + 1) + c] = frame_jac[1][c];
      }
      ++feature_counter;
    }
  }

  // b4. ref: sparse_img_align.cpp:319-403
  void precomputeJacobiansAndRefPatches(const AlignFrame& ref_frame, size_t level, int patch_size, size_t nr_features,
                                        bool estimate_alpha, bool estimate_beta, size_t& feature_counter) {
    const Img& ref_img = ref_frame.img_pyr.at(level);
    const int stride = ref_img.step;
    const FloatType scale = 1.0f / (1 << level);
    const int patch_area = patch_size * patch_size;
    const int border = 1;
    const int patch_size_wb = patch_size + 2 * border;
    const int patch_area_wb = patch_size_wb * patch_size_wb;
    const FloatType patch_center_wb = (patch_size_wb - 1) / 2.0f;
    std::vector<FloatType> interp_patch_array(patch_area_wb);

    for (size_t i = 0; i < nr_features; ++i, ++feature_counter) {
      const FloatType u_tl = uv_cache_[2 * feature_counter + 0] * scale - patch_center_wb;
      const FloatType v_tl = uv_cache_[2 * feature_counter + 1] * scale - patch_center_wb;
      const int u_tl_i = std::floor(u_tl);
      const int v_tl_i = std::floor(v_tl);
      const FloatType subpix_u_tl = u_tl - u_tl_i;
      const FloatType subpix_v_tl = v_tl - v_tl_i;
      const FloatType wtl = (1.0 - subpix_u_tl) * (1.0 - subpix_v_tl);
      const FloatType wtr = subpix_u_tl * (1.0 - subpix_v_tl);
      const FloatType wbl = (1.0 - subpix_u_tl) * subpix_v_tl;
      const FloatType wbr = subpix_u_tl * subpix_v_tl;
      const int jacobian_proj_col = 2 * feature_counter;

      size_t pixel_counter = 0;
      for (int y = 0; y < patch_size_wb; ++y) {
        const uint8_t* r = ref_img.data + (v_tl_i + y) * stride + u_tl_i;
        for (int x = 0; x < patch_size_wb; ++x, ++r, ++pixel_counter)
          interp_patch_array[pixel_counter] = wtl * r[0] + wtr * r[1] + wbl * r[stride] + wbr * r[stride + 1];
      }
      pixel_counter = 0;
      const double* Jp0 = &jacobian_proj_cache_[6 * jacobian_proj_col];
      const double* Jp1 = &jacobian_proj_cache_[6 * (jacobian_proj_col + 1)];
      for (int y = 0; y < patch_size; ++y) {
        for (int x = 0; x < patch_size; ++x, ++pixel_counter) {
          const int offset_center = (x + border) + patch_size_wb * (y + border);
          ref_patch_cache_[patch_area * feature_counter + pixel_counter] = interp_patch_array[offset_center];
          const FloatType dx = 0.5f * (interp_patch_array[offset_center + 1] - interp_patch_array[offset_center - 1]);
          const FloatType dy = 0.5f * (interp_patch_array[offset_center + patch_size_wb] - interp_patch_array[offset_center - patch_size_wb]);
          const size_t jacobian_col = feature_counter * patch_area + pixel_counter;
          double* Jc = &jacobian_cache_[8 * jacobian_col];
          for (int c = 0; c < 6; ++c) Jc[c] = (dx * Jp0[c] + dy * Jp1[c]) * scale;
          Jc[6] = estimate_alpha ? -(interp_patch_array[offset_center]) : 0.0;
          Jc[7] = estimate_beta ? -1.0 : 0.0;
        }
      }
    }
  }

  // b5. ref: sparse_img_align.cpp:405-498
  void computeResidualsOfFrame(const AlignFrame& cur_frame, size_t level, int patch_size, size_t nr_features,
                               const SE3& T_cur_ref, const float alpha, const float beta, size_t& feature_counter) {
    const Img& cur_img = cur_frame.img_pyr.at(level);
    const int stride = cur_img.step;
    const FloatType scale = 1.0f / (1 << level);
    const int patch_area = patch_size * patch_size;
    const FloatType patch_center = (patch_size - 1) / 2.0f;

    for (size_t i = 0; i < nr_features; ++i, ++feature_counter) {
      const V3 xyz_ref{xyz_ref_cache_[3 * feature_counter], xyz_ref_cache_[3 * feature_counter + 1], xyz_ref_cache_[3 * feature_counter + 2]};
      const V3 xyz_cur = T_cur_ref * xyz_ref;
      if (/* pinhole && */ xyz_cur.z < 0.0) {
        visibility_mask_[feature_counter] = false;
        continue;
      }
      const V2 uv_cur = cur_frame.cam.project3(xyz_cur);
      const FloatType uv_cur_pyr0 = uv_cur.x * scale, uv_cur_pyr1 = uv_cur.y * scale;
      const FloatType u_tl = uv_cur_pyr0 - patch_center;
      const FloatType v_tl = uv_cur_pyr1 - patch_center;
      if (u_tl < 0.0 || v_tl < 0.0 || u_tl + patch_size + 2.0 >= cur_img.cols || v_tl + patch_size + 2.0 >= cur_img.rows) {
        visibility_mask_[feature_counter] = false;
        continue;
      } else {
        visibility_mask_[feature_counter] = true;
      }
      const int u_tl_i = std::floor(u_tl);
      const int v_tl_i = std::floor(v_tl);
      const FloatType subpix_u_tl = u_tl - u_tl_i;
      const FloatType subpix_v_tl = v_tl - v_tl_i;
      const FloatType wtl = (1.0 - subpix_u_tl) * (1.0 - subpix_v_tl);
      const FloatType wtr = subpix_u_tl * (1.0 - subpix_v_tl);
      const FloatType wbl = (1.0 - subpix_u_tl) * subpix_v_tl;
      const FloatType wbr = subpix_u_tl * subpix_v_tl;

      size_t pixel_counter = 0;
      for (int y = 0; y < patch_size; ++y) {
        const uint8_t* cur_img_ptr = cur_img.data + (v_tl_i + y) * stride + u_tl_i;
        for (int x = 0; x < patch_size; ++x, ++pixel_counter, ++cur_img_ptr) {
          const FloatType intensity_cur = wtl * cur_img_ptr[0] + wtr * cur_img_ptr[1] + wbl * cur_img_ptr[stride] + wbr * cur_img_ptr[stride + 1];
          const FloatType res = static_cast<FloatType>(intensity_cur * (1.0 + alpha) + beta)
                                - ref_patch_cache_[patch_area * feature_counter + pixel_counter];
          residual_cache_[patch_area * feature_counter + pixel_counter] = res;
        }
      }
    }
  }

  // b6. ref: sparse_img_align.cpp:500-541
  FloatType computeHessianAndGradient(double H[8][8], double g[8]) {
    float chi2 = 0.0;
    size_t n_meas = 0;
    const size_t patch_area = patch_area_;
    const size_t mask_size = visibility_mask_.size();
    for (size_t i = 0; i < mask_size; ++i) {
      if (visibility_mask_[i] == true) {
        const size_t patch_offset = i * patch_area;
        for (size_t j = 0; j < patch_area; ++j) {
          const FloatType res = residual_cache_[patch_offset + j];
          float weight = 1.0;
          if (options_.robustification) weight = tukey_.weight(res / weight_scale_f_);
          chi2 += res * res * weight;
          ++n_meas;
          const double* J = &jacobian_cache_[8 * (patch_offset + j)];
          for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) H[r][c] += J[r] * J[c] * weight;
          for (int r = 0; r < 8; ++r) g[r] -= J[r] * res * weight;
        }
      }
    }
    return chi2 / n_meas;
  }

  // b9. ref: sparse_img_align.cpp:115-156
  double evaluateError(const SparseImgAlignState& state, double H[8][8], double g[8]) {
    if (!have_cache_) {
      size_t feature_counter = 0;
      for (size_t i = 0; i < ref_frames_->size(); ++i)
        precomputeJacobiansAndRefPatches(ref_frames_->at(i), level_, patch_size_, fts_vec_.at(i).size(),
                                         options_.estimate_illumination_gain, options_.estimate_illumination_offset, feature_counter);
      have_cache_ = true;
    }
    size_t feature_counter = 0;
    for (size_t i = 0; i < ref_frames_->size(); ++i) {
      const SE3 T_cur_ref = cur_frames_->at(i).T_cam_imu * state.T_icur_iref * inverse(ref_frames_->at(i).T_cam_imu);
      computeResidualsOfFrame(cur_frames_->at(i), level_, patch_size_, fts_vec_.at(i).size(), T_cur_ref,
                              static_cast<float>(state.alpha), static_cast<float>(state.beta), feature_counter);
    }
    const float chi2 = computeHessianAndGradient(H, g);
    return chi2;
  }

  // b8. ref: sparse_img_align_base.cpp:64-75
  static void update(const SparseImgAlignState& state_old, const double dx[8], SparseImgAlignState& state_new) {
    const double mdx[6] = {-dx[0], -dx[1], -dx[2], -dx[3], -dx[4], -dx[5]};
    state_new.T_icur_iref = state_old.T_icur_iref * se3Exp(mdx);
    state_new.alpha = (state_old.alpha - dx[6]) / (1.0 + dx[6]);
    state_new.beta = (state_old.beta - dx[7]) / (1.0 + dx[6]);
    quatNormalize(state_new.T_icur_iref.q);
  }

  // b8. ref: sparse_img_align_base.cpp:77-107
  void applyPrior(const SparseImgAlignState& state) {
    if (iter_ == 0) {
      double H_max_diag_trans = 0;
      for (size_t j = 0; j < 3; ++j) H_max_diag_trans = std::max(H_max_diag_trans, std::fabs(H_[j][j]));
      double H_max_diag_rot = 0;
      for (size_t j = 3; j < 6; ++j) H_max_diag_rot = std::max(H_max_diag_rot, std::fabs(H_[j][j]));
      const double I_alpha = prior_lambda_alpha_ * H_[6][6];
      const double I_beta = prior_lambda_beta_ * H_[7][7];
      for (auto& r : I_prior_) for (double& v : r) v = 0.0;
      for (int j = 0; j < 3; ++j) I_prior_[j][j] = 1.0 * prior_lambda_trans_ * H_max_diag_trans;
      for (int j = 3; j < 6; ++j) I_prior_[j][j] = 1.0 * prior_lambda_rot_ * H_max_diag_rot;
      I_prior_[6][6] = I_alpha;
      I_prior_[7][7] = I_beta;
    }
    for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) H_[r][c] += I_prior_[r][c];
    double l[6];
    se3Log(inverse(prior_.T_icur_iref) * state.T_icur_iref, l);
    for (int r = 0; r < 6; ++r) {
      double s = 0;
      for (int c = 0; c < 6; ++c) s += I_prior_[r][c] * l[c];
      g_[r] += s;
    }
    g_[6] += I_prior_[6][6] * (prior_.alpha - state.alpha);
    g_[7] += I_prior_[7][7] * (prior_.beta - state.beta);
  }

  // b7. ref: mini_least_squares_solver.hpp:42-107 (optimizeGaussNewton)
  void optimizeGaussNewton(SparseImgAlignState& state) {
    SparseImgAlignState old_state = state;
    int n_eval = 0;
    for (iter_ = 0; iter_ < solver_options_.max_iter; ++iter_) {
      for (auto& r : H_) for (double& v : r) v = 0.0;
      for (double& v : g_) v = 0.0;
      const double new_chi2 = evaluateError(state, H_, g_);
      ++n_eval;
      if (have_prior_) applyPrior(state);
      // solveDefaultImpl (:253-262)
      ldltSolve<8>(H_, g_, dx_);
      if (std::isnan(dx_[0])) stop_ = true;
      if ((iter_ > 0 && new_chi2 > chi2_ && solver_options_.stop_when_error_increases) || stop_) {
        state = old_state;  // rollback
        break;
      }
      SparseImgAlignState new_state;
      update(state, dx_, new_state);
      old_state = state;
      state = new_state;
      chi2_ = new_chi2;
      double x_norm = -1;
      for (int i = 0; i < 8; ++i) { const double a = std::fabs(dx_[i]); if (a > x_norm) x_norm = a; }
      if (x_norm < solver_options_.eps) break;
    }
    iters_per_level_.push_back(n_eval);
  }

  // b9. ref: sparse_img_align.cpp:34-113
  SparseImgAlignResult run(const std::vector<AlignFrame>& ref_frames, const std::vector<AlignFrame>& cur_frames) {
    SparseImgAlignResult res;
    fts_vec_.clear();
    iters_per_level_.clear();
    size_t n_fts_to_track = 0;
    for (const AlignFrame& frame : ref_frames) {
      std::vector<size_t> fts;
      extractFeaturesSubset(frame, options_.max_level, patch_size_with_border_, fts);
      n_fts_to_track += fts.size();
      fts_vec_.push_back(fts);
    }
    res.n_fts_to_track = n_fts_to_track;
    res.T_icur_iref = cur_frames.at(0).T_imu_world * inverse(ref_frames.at(0).T_imu_world);
    if (n_fts_to_track == 0) return res;

    ref_frames_ = &ref_frames;
    cur_frames_ = &cur_frames;
    const SE3 T_iref_world = ref_frames.at(0).T_imu_world;

    uv_cache_.assign(2 * n_fts_to_track, 0);
    xyz_ref_cache_.assign(3 * n_fts_to_track, 0);
    jacobian_proj_cache_.assign(6 * 2 * n_fts_to_track, 0);
    jacobian_cache_.assign(8 * n_fts_to_track * patch_area_, 0);
    residual_cache_.assign(patch_area_ * n_fts_to_track, 0);
    visibility_mask_.assign(n_fts_to_track, 0);
    ref_patch_cache_.assign(patch_area_ * n_fts_to_track, 0);

    SparseImgAlignState state;
    state.T_icur_iref = cur_frames.at(0).T_imu_world * inverse(T_iref_world);
    state.alpha = alpha_init_;
    state.beta = beta_init_;

    size_t feature_counter = 0;
    for (size_t i = 0; i < ref_frames.size(); ++i)
      precomputeBaseCaches(ref_frames.at(i), fts_vec_.at(i), options_.use_distortion_jacobian, feature_counter);

    for (level_ = options_.max_level; level_ >= options_.min_level; --level_) {
      have_cache_ = false;
      optimizeGaussNewton(state);
    }
    for (const AlignFrame& f : cur_frames) res.T_f_w.push_back(f.T_cam_imu * state.T_icur_iref * T_iref_world);
    alpha_init_ = 0.0;
    beta_init_ = 0.0;

    res.T_icur_iref = state.T_icur_iref;
    res.alpha = state.alpha;
    res.beta = state.beta;
    res.chi2 = chi2_;
    for (int r = 0; r < 8; ++r) for (int c = 0; c < 8; ++c) res.H[r][c] = H_[r][c];
    res.iters_per_level = iters_per_level_;
    res.stop = stop_;
    return res;
  }
};

}  // namespace orc
