// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.hpp header for the rules).
//
// Rows c1-c7 of SURVEY.md §8: affine patch warp, ZMSSD score, align1D / align2D,
// Matcher::findMatchDirect and Matcher::findEpipolarMatchDirect (unit-plane and unit-sphere scans).
// Parity status: "parity unpinned" (restatement; no reference vectors exist for this path).
#pragma once
#include <vector>
#include <cmath>
#include "orc_math.hpp"

namespace orc {

// ---------------------------------------------------------------------------
// c1. warp::getWarpMatrixAffine — ref: src/svo_direct/src/patch_warp.cpp:20-60 (pinhole branch)
inline void getWarpMatrixAffine(const Camera& cam_ref, const Camera& cam_cur, const V2& px_ref, const V3& f_ref,
                                const double depth_ref, const SE3& T_cur_ref, const int level_ref, double A_cur_ref[2][2]) {
  const int kHalfPatchSize = 5;
  const V3 xyz_ref = f_ref * depth_ref;
  V3 xyz_du_ref = cam_ref.backProject3({px_ref.x + double(kHalfPatchSize) * (1 << level_ref), px_ref.y + 0.0 * (1 << level_ref)});
  V3 xyz_dv_ref = cam_ref.backProject3({px_ref.x + 0.0 * (1 << level_ref), px_ref.y + double(kHalfPatchSize) * (1 << level_ref)});
  xyz_du_ref = xyz_du_ref * xyz_ref.z;
  xyz_dv_ref = xyz_dv_ref * xyz_ref.z;
  const V2 px_cur = cam_cur.project3(T_cur_ref * xyz_ref);
  const V2 px_du_cur = cam_cur.project3(T_cur_ref * xyz_du_ref);
  const V2 px_dv_cur = cam_cur.project3(T_cur_ref * xyz_dv_ref);
  A_cur_ref[0][0] = (px_du_cur.x - px_cur.x) / kHalfPatchSize;
  A_cur_ref[1][0] = (px_du_cur.y - px_cur.y) / kHalfPatchSize;
  A_cur_ref[0][1] = (px_dv_cur.x - px_cur.x) / kHalfPatchSize;
  A_cur_ref[1][1] = (px_dv_cur.y - px_cur.y) / kHalfPatchSize;
}

// c1. warp::getBestSearchLevel — ref: patch_warp.cpp:97-110
inline int getBestSearchLevel(const double A[2][2], const int max_level) {
  int search_level = 0;
  double D = A[0][0] * A[1][1] - A[1][0] * A[0][1];  // Eigen 2x2 determinant
  while (D > 3.0 && search_level < max_level) {
    search_level += 1;
    D *= 0.25;
  }
  return search_level;
}

// c1. warp::warpAffine — ref: patch_warp.cpp:112-156
// Float arithmetic; compile this oracle with -ffp-contract=off so no FMA is formed.
inline bool warpAffine(const double A_cur_ref[2][2], const Img& img_ref, const V2& px_ref, const int level_ref,
                       const int search_level, const int halfpatch_size, uint8_t* patch) {
  // Eigen 2x2 inverse: adj / det via invdet multiply (Eigen/src/LU/InverseImpl.h compute_inverse_size2_helper)
  const double det = A_cur_ref[0][0] * A_cur_ref[1][1] - A_cur_ref[1][0] * A_cur_ref[0][1];
  const double invdet = 1.0 / det;
  const double Ai[2][2] = {{A_cur_ref[1][1] * invdet, -A_cur_ref[0][1] * invdet},
                           {-A_cur_ref[1][0] * invdet, A_cur_ref[0][0] * invdet}};
  const float sl = float(1 << search_level);
  const float A_ref_cur[2][2] = {{float(Ai[0][0]) * sl, float(Ai[0][1]) * sl}, {float(Ai[1][0]) * sl, float(Ai[1][1]) * sl}};
  if (std::isnan(A_ref_cur[0][0])) return false;

  uint8_t* patch_ptr = patch;
  const float lr = float(1 << level_ref);
  const float px_ref_pyr[2] = {float(px_ref.x) / lr, float(px_ref.y) / lr};
  const int stride = img_ref.step;
  for (int y = -halfpatch_size; y < halfpatch_size; ++y) {
    for (int x = -halfpatch_size; x < halfpatch_size; ++x, ++patch_ptr) {
      const float px_patch[2] = {float(x), float(y)};
      const float px0 = (A_ref_cur[0][0] * px_patch[0] + A_ref_cur[0][1] * px_patch[1]) + px_ref_pyr[0];
      const float px1 = (A_ref_cur[1][0] * px_patch[0] + A_ref_cur[1][1] * px_patch[1]) + px_ref_pyr[1];
      const int xi = std::floor(px0);
      const int yi = std::floor(px1);
      if (xi < 0 || yi < 0 || xi + 1 >= img_ref.cols || yi + 1 >= img_ref.rows) return false;
      const float subpix_x = px0 - xi;
      const float subpix_y = px1 - yi;
      const float w00 = (1.0f - subpix_x) * (1.0f - subpix_y);
      const float w01 = (1.0f - subpix_x) * subpix_y;
      const float w10 = subpix_x * (1.0f - subpix_y);
      const float w11 = 1.0f - w00 - w01 - w10;
      const uint8_t* const ptr = img_ref.data + yi * stride + xi;
      *patch_ptr = static_cast<uint8_t>(w00 * ptr[0] + w01 * ptr[stride] + w10 * ptr[1] + w11 * ptr[stride + 1]);
    }
  }
  return true;
}

// c2. patch_utils::createPatchFromPatchWithBorder — ref: src/svo_direct/include/svo/direct/patch_utils.h:18-30
inline void createPatchFromPatchWithBorder(const uint8_t* patch_with_border, const int patch_size, uint8_t* patch) {
  uint8_t* patch_ptr = patch;
  for (int y = 1; y < patch_size + 1; ++y, patch_ptr += patch_size) {
    const uint8_t* ref_patch_border_ptr = patch_with_border + y * (patch_size + 2) + 1;
    for (int x = 0; x < patch_size; ++x) patch_ptr[x] = ref_patch_border_ptr[x];
  }
}

// c5. patch_score::ZMSSD<4> — ref: src/svo_direct/include/svo/direct/patch_score.h:43-109 (ctor, plain loop),
// :264-283 (strided computeScore, the branch the reference's -msse3 flags build).
struct ZMSSD {
  static const int patch_size_ = 8;
  static const int patch_area_ = 64;
  static const int threshold_ = 2000 * patch_area_;
  const uint8_t* ref_patch_;
  int sumA_, sumAA_;
  explicit ZMSSD(const uint8_t* ref_patch) : ref_patch_(ref_patch) {
    uint32_t sumA_uint = 0, sumAA_uint = 0;
    for (int r = 0; r < patch_area_; r++) {
      const uint8_t n = ref_patch_[r];
      sumA_uint += n;
      sumAA_uint += n * n;
    }
    sumA_ = sumA_uint;
    sumAA_ = sumAA_uint;
  }
  static int threshold() { return threshold_; }
  int computeScore(const uint8_t* cur_patch, int stride) const {
    uint32_t sumB_uint = 0, sumBB_uint = 0, sumAB_uint = 0;
    for (int y = 0, r = 0; y < patch_size_; ++y) {
      const uint8_t* cur_patch_ptr = cur_patch + y * stride;
      for (int x = 0; x < patch_size_; ++x, ++r) {
        const uint8_t cur_px = cur_patch_ptr[x];
        sumB_uint += cur_px;
        sumBB_uint += cur_px * cur_px;
        sumAB_uint += cur_px * ref_patch_[r];
      }
    }
    const int sumB = sumB_uint, sumBB = sumBB_uint, sumAB = sumAB_uint;
    return sumAA_ - 2 * sumAB + sumBB - (sumA_ * sumA_ - 2 * sumA_ * sumB + sumB * sumB) / patch_area_;
  }
};

// Eigen::Matrix3f::inverse() (cofactor form, Eigen/src/LU/InverseImpl.h compute_inverse_size3_helper)
inline void inverse3f(const float m[3][3], float r[3][3]) {
  auto cof = [&](int i, int j) -> float {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  const float c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
  const float det = c00 * m[0][0] + c10 * m[1][0] + c20 * m[2][0];
  const float invdet = 1.0f / det;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r[i][j] = cof(j, i) * invdet;
}
// Eigen::Matrix4f::inverse(): general cofactor expansion (scalar path; the SSE path is algebraically identical).
inline void inverse4f(const float m[4][4], float r[4][4]) {
  auto det3 = [&](int r0, int r1, int r2, int c0, int c1, int c2) -> float {
    return m[r0][c0] * (m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1])
         - m[r0][c1] * (m[r1][c0] * m[r2][c2] - m[r1][c2] * m[r2][c0])
         + m[r0][c2] * (m[r1][c0] * m[r2][c1] - m[r1][c1] * m[r2][c0]);
  };
  float cofm[4][4];
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
    int rr[3], cc[3]; int a = 0, b = 0;
    for (int k = 0; k < 4; ++k) { if (k != i) rr[a++] = k; if (k != j) cc[b++] = k; }
    const float d = det3(rr[0], rr[1], rr[2], cc[0], cc[1], cc[2]);
    cofm[i][j] = ((i + j) & 1) ? -d : d;
  }
  const float det = m[0][0] * cofm[0][0] + m[0][1] * cofm[0][1] + m[0][2] * cofm[0][2] + m[0][3] * cofm[0][3];
  const float invdet = 1.0f / det;
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r[i][j] = cofm[j][i] * invdet;
}

// f3. feature_alignment::alignPyr2D — ref: src/svo_direct/src/feature_alignment.cpp:761-973 (non-NEON path)
// Pyramidal inverse-compositional KLT on integer gradients: the template is cut at the truncated reference position,
// the current image is sampled with 7-bit fixed-point bilinear weights (rounded descale), H and Jres are float sums in
// raster order, update = Hinv * Jres * 2 (the gradients are twice the central difference).
inline bool alignPyr2D(const std::vector<Img>& img_pyr_ref, const std::vector<Img>& img_pyr_cur, const int max_level,
                       const int min_level, const std::vector<int>& patch_sizes, const int n_iter,
                       const float min_update_squared, const int px_ref_level_0[2], V2& px_cur_level_0) {
  const int max_patch_area = patch_sizes[0] * patch_sizes[0];
  std::vector<uint8_t> ref_patch(max_patch_area);
  std::vector<int16_t> ref_patch_dx(max_patch_area), ref_patch_dy(max_patch_area);
  bool converged = false;
  for (int level = max_level; level >= min_level; --level) {
    const int patch_size = patch_sizes[level];
    const int halfpatch_size = patch_size / 2;
    const int scale = (1 << level);
    const Img& img_ref = img_pyr_ref[level];
    const Img& img_cur = img_pyr_cur[level];
    const int width = img_ref.cols;
    const int height = img_ref.rows;
    const int step = img_ref.step;
    const float px_ref_flt[2] = {float(px_ref_level_0[0]) / scale - float(halfpatch_size),
                                 float(px_ref_level_0[1]) / scale - float(halfpatch_size)};
    const int px_ref[2] = {int(px_ref_flt[0]), int(px_ref_flt[1])};
    const float px_ref_offset[2] = {px_ref_flt[0] - float(px_ref[0]), px_ref_flt[1] - float(px_ref[1])};
    if (px_ref[0] < 1 || px_ref[1] < 1 || px_ref[0] >= width - patch_size - 1 || px_ref[1] >= height - patch_size - 1) continue;
    uint8_t* it_patch = ref_patch.data();
    int16_t* it_dx = ref_patch_dx.data();
    int16_t* it_dy = ref_patch_dy.data();
    float H[2][2] = {{0, 0}, {0, 0}};
    for (int y = 0; y < patch_size; ++y) {
      const uint8_t* it = img_ref.data + (px_ref[1] + y) * step + (px_ref[0]);
      for (int x = 0; x < patch_size; ++x, ++it, ++it_patch, ++it_dx, ++it_dy) {
        *it_patch = *it;
        *it_dx = static_cast<int16_t>(it[1]) - it[-1];
        *it_dy = static_cast<int16_t>(it[step]) - it[-step];
        const float J[2] = {float(*it_dx), float(*it_dy)};
        for (int r = 0; r < 2; ++r) for (int c = 0; c < 2; ++c) H[r][c] += J[r] * J[c];
      }
    }
    // Eigen::Matrix2f::inverse(): compute_inverse_size2_helper (invdet multiply)
    const float invdet = 1.0f / (H[0][0] * H[1][1] - H[1][0] * H[0][1]);
    const float Hinv[2][2] = {{H[1][1] * invdet, -H[0][1] * invdet}, {-H[1][0] * invdet, H[0][0] * invdet}};
    float u = float(px_cur_level_0.x / scale - halfpatch_size - px_ref_offset[0]);
    float v = float(px_cur_level_0.y / scale - halfpatch_size - px_ref_offset[1]);
    float update[2] = {0, 0};
    bool go_to_next_level = false;
    const int SHIFT_BITS = 7;
    converged = false;
    for (int iter = 0; iter < n_iter; ++iter) {
      if (std::isnan(u) || std::isnan(v)) return false;
      go_to_next_level = false;
      const int u_r = std::floor(u);
      const int v_r = std::floor(v);
      if (u_r < 0 || v_r < 0 || u_r >= width - patch_size || v_r >= height - patch_size) {
        go_to_next_level = true;
        break;
      }
      const float subpix_x = u - u_r;
      const float subpix_y = v - v_r;
      const uint16_t wTL = static_cast<uint16_t>((1.0f - subpix_x) * (1.0f - subpix_y) * (1 << SHIFT_BITS));
      const uint16_t wTR = static_cast<uint16_t>(subpix_x * (1.0f - subpix_y) * (1 << SHIFT_BITS));
      const uint16_t wBL = static_cast<uint16_t>((1.0f - subpix_x) * subpix_y * (1 << SHIFT_BITS));
      const uint16_t wBR = (1 << SHIFT_BITS) - wTL - wTR - wBL;
      const uint8_t* it_ref = ref_patch.data();
      float Jres[2] = {0, 0};
      const int16_t* it_ref_dx = ref_patch_dx.data();
      const int16_t* it_ref_dy = ref_patch_dy.data();
      for (int y = 0; y < patch_size; ++y) {
        const uint8_t* it = img_cur.data + (v_r + y) * step + (u_r);
        for (int x = 0; x < patch_size; ++x, ++it, ++it_ref, ++it_ref_dx, ++it_ref_dy) {
          const uint16_t cur = ((wTL * it[0] + wTR * it[1] + wBL * it[step] + wBR * it[step + 1]) + (1 << (SHIFT_BITS - 1))) >> SHIFT_BITS;
          const float res = static_cast<float>(cur) - *it_ref;
          Jres[0] -= res * (*it_ref_dx);
          Jres[1] -= res * (*it_ref_dy);
        }
      }
      update[0] = (Hinv[0][0] * Jres[0] + Hinv[0][1] * Jres[1]) * 2.0f;
      update[1] = (Hinv[1][0] * Jres[0] + Hinv[1][1] * Jres[1]) * 2.0f;
      u += update[0];
      v += update[1];
      if (update[0] * update[0] + update[1] * update[1] < min_update_squared) {
        converged = true;
        break;
      }
    }
    px_cur_level_0 = V2{double((u + halfpatch_size + px_ref_offset[0]) * scale), double((v + halfpatch_size + px_ref_offset[1]) * scale)};
    if (!converged && !go_to_next_level) return false;
  }
  return converged;
}

// c4. feature_alignment::align1D — ref: src/svo_direct/src/feature_alignment.cpp:31-209
inline bool align1D(const Img& cur_img, const V2& dir, const uint8_t* ref_patch_with_border, const uint8_t* ref_patch,
                    const int n_iter, const bool affine_est_offset, const bool affine_est_gain,
                    V2* cur_px_estimate, double* h_inv = nullptr) {
  constexpr int kHalfPatchSize = 4;
  constexpr int kPatchSize = 2 * kHalfPatchSize;
  constexpr int kPatchArea = kPatchSize * kPatchSize;
  bool converged = false;
  float ref_patch_dv[kPatchArea];
  float H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  constexpr int ref_step = kPatchSize + 2;
  float* it_dv = ref_patch_dv;
  for (int y = 0; y < kPatchSize; ++y) {
    const uint8_t* it = ref_patch_with_border + (y + 1) * ref_step + 1;
    for (int x = 0; x < kPatchSize; ++x, ++it, ++it_dv) {
      float J[3];
      const float dx = static_cast<float>(it[1]) - static_cast<float>(it[-1]);
      const float dy = static_cast<float>(it[ref_step]) - static_cast<float>(it[-ref_step]);
      J[0] = 0.5f * (dir.x * dx + dir.y * dy);
      J[1] = affine_est_offset ? 1.0f : 0.0f;
      J[2] = affine_est_gain ? -1.0f * it[0] : 0.0f;
      *it_dv = J[0];
      for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) H[r][c] += J[r] * J[c];
    }
  }
  if (!affine_est_offset) H[1][1] = 1.0;
  if (!affine_est_gain) H[2][2] = 1.0;
  if (h_inv) *h_inv = 1.0 / H[0][0] * kPatchSize * kPatchSize;
  float Hinv[3][3];
  inverse3f(H, Hinv);
  float mean_diff = 0;
  float alpha = 1.0;
  float u = cur_px_estimate->x;
  float v = cur_px_estimate->y;
  const float min_update_squared = 0.03 * 0.03;
  const int cur_step = cur_img.step;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = std::floor(u);
    const int v_r = std::floor(v);
    if (u_r < kHalfPatchSize || v_r < kHalfPatchSize || u_r >= cur_img.cols - kHalfPatchSize || v_r >= cur_img.rows - kHalfPatchSize)
      break;
    if (std::isnan(u) || std::isnan(v)) return false;
    const float subpix_x = u - u_r;
    const float subpix_y = v - v_r;
    const float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    const float wTR = subpix_x * (1.0 - subpix_y);
    const float wBL = (1.0 - subpix_x) * subpix_y;
    const float wBR = subpix_x * subpix_y;
    const uint8_t* it_ref = ref_patch;
    const float* it_ref_dv = ref_patch_dv;
    float Jres[3] = {0, 0, 0};
    for (int y = 0; y < kPatchSize; ++y) {
      const uint8_t* it = cur_img.data + (v_r + y - kHalfPatchSize) * cur_step + u_r - kHalfPatchSize;
      for (int x = 0; x < kPatchSize; ++x, ++it, ++it_ref, ++it_ref_dv) {
        const float cur_intensity = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        const float res = cur_intensity - alpha * (*it_ref) + mean_diff;
        Jres[0] -= res * (*it_ref_dv);
        if (affine_est_offset) Jres[1] -= res;
        if (affine_est_gain) Jres[2] -= (-1) * res * (*it_ref);
      }
    }
    if (!affine_est_offset) Jres[1] = 0.0;
    if (!affine_est_gain) Jres[2] = 0.0;
    float update[3];
    for (int r = 0; r < 3; ++r) update[r] = Hinv[r][0] * Jres[0] + Hinv[r][1] * Jres[1] + Hinv[r][2] * Jres[2];
    u += update[0] * dir.x;
    v += update[0] * dir.y;
    mean_diff += update[1];
    alpha += update[2];
    if (update[0] * update[0] < min_update_squared) {
      converged = true;
      break;
    }
  }
  cur_px_estimate->x = u;
  cur_px_estimate->y = v;
  return converged;
}

// c3. feature_alignment::align2D — ref: src/svo_direct/src/feature_alignment.cpp:212-391 (float path)
inline bool align2D(const Img& cur_img, const uint8_t* ref_patch_with_border, const uint8_t* ref_patch, const int n_iter,
                    const bool affine_est_offset, const bool affine_est_gain, V2& cur_px_estimate) {
  const int halfpatch_size_ = 4;
  const int patch_size_ = 8;
  const int patch_area_ = 64;
  bool converged = false;
  float ref_patch_dx[patch_area_];
  float ref_patch_dy[patch_area_];
  float H[4][4];
  for (auto& r : H) for (float& v : r) v = 0;
  const int ref_step = patch_size_ + 2;
  float* it_dx = ref_patch_dx;
  float* it_dy = ref_patch_dy;
  for (int y = 0; y < patch_size_; ++y) {
    const uint8_t* it = ref_patch_with_border + (y + 1) * ref_step + 1;
    for (int x = 0; x < patch_size_; ++x, ++it, ++it_dx, ++it_dy) {
      float J[4];
      J[0] = 0.5 * (it[1] - it[-1]);
      J[1] = 0.5 * (it[ref_step] - it[-ref_step]);
      J[2] = affine_est_offset ? 1.0 : 0.0;
      J[3] = affine_est_gain ? -1.0 * it[0] : 0.0;
      *it_dx = J[0];
      *it_dy = J[1];
      for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) H[r][c] += J[r] * J[c];
    }
  }
  if (!affine_est_offset) H[2][2] = 1.0;
  if (!affine_est_gain) H[3][3] = 1.0;
  float Hinv[4][4];
  inverse4f(H, Hinv);
  float mean_diff = 0;
  float alpha = 1.0;
  float u = cur_px_estimate.x;
  float v = cur_px_estimate.y;
  const float min_update_squared = 0.03 * 0.03;
  const int cur_step = cur_img.step;
  float update[4] = {0, 0, 0, 0};
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = std::floor(u);
    const int v_r = std::floor(v);
    if (u_r < halfpatch_size_ || v_r < halfpatch_size_ || u_r >= cur_img.cols - halfpatch_size_ || v_r >= cur_img.rows - halfpatch_size_)
      break;
    if (std::isnan(u) || std::isnan(v)) return false;
    const float subpix_x = u - u_r;
    const float subpix_y = v - v_r;
    const float wTL = (1.0 - subpix_x) * (1.0 - subpix_y);
    const float wTR = subpix_x * (1.0 - subpix_y);
    const float wBL = (1.0 - subpix_x) * subpix_y;
    const float wBR = subpix_x * subpix_y;
    const uint8_t* it_ref = ref_patch;
    const float* it_ref_dx = ref_patch_dx;
    const float* it_ref_dy = ref_patch_dy;
    float Jres[4] = {0, 0, 0, 0};
    for (int y = 0; y < patch_size_; ++y) {
      const uint8_t* it = cur_img.data + (v_r + y - halfpatch_size_) * cur_step + u_r - halfpatch_size_;
      for (int x = 0; x < patch_size_; ++x, ++it, ++it_ref, ++it_ref_dx, ++it_ref_dy) {
        const float search_pixel = wTL * it[0] + wTR * it[1] + wBL * it[cur_step] + wBR * it[cur_step + 1];
        const float res = search_pixel - alpha * (*it_ref) + mean_diff;
        Jres[0] -= res * (*it_ref_dx);
        Jres[1] -= res * (*it_ref_dy);
        if (affine_est_offset) Jres[2] -= res;
        if (affine_est_gain) Jres[3] -= (-1) * res * (*it_ref);
      }
    }
    if (!affine_est_offset) Jres[2] = 0.0;
    if (!affine_est_gain) Jres[3] = 0.0;
    for (int r = 0; r < 4; ++r)
      update[r] = Hinv[r][0] * Jres[0] + Hinv[r][1] * Jres[1] + Hinv[r][2] * Jres[2] + Hinv[r][3] * Jres[3];
    u += update[0];
    v += update[1];
    mean_diff += update[2];
    alpha += update[3];
    if (update[0] * update[0] + update[1] * update[1] < min_update_squared) {
      converged = true;
      break;
    }
  }
  cur_px_estimate.x = u;
  cur_px_estimate.y = v;
  return converged;
}

// ---------------------------------------------------------------------------
// c6/c7. svo::Matcher — ref: src/svo_direct/include/svo/direct/matcher.h:28-140, src/svo_direct/src/matcher.cpp
enum class FeatureType : uint8_t {  // ref: src/svo_common/include/svo/common/types.h:60-73
  kEdgeletSeed = 0, kCornerSeed = 1, kMapPointSeed = 2, kEdgeletSeedConverged = 3, kCornerSeedConverged = 4,
  kMapPointSeedConverged = 5, kEdgelet = 6, kCorner = 7, kMapPoint = 8, kFixedLandmark = 9, kOutlier = 10
};
inline bool isEdgelet(FeatureType t) {  // ref: types.h (isEdgelet)
  return t == FeatureType::kEdgelet || t == FeatureType::kEdgeletSeed || t == FeatureType::kEdgeletSeedConverged;
}
inline bool isSeed(FeatureType t) { return static_cast<uint8_t>(t) < 6; }

struct MatchFrame {   // the parts of svo::Frame the matcher reads
  std::vector<Img> img_pyr;
  Camera cam;
};
struct FeatureRef {   // the parts of svo::FeatureWrapper the matcher reads
  FeatureType type;
  V2 px;
  V3 f;
  V2 grad;
  int level;
};

class Matcher {
 public:
  static const int kHalfPatchSize = 4;
  static const int kPatchSize = 8;
  struct Options {  // ref: matcher.h:39-54
    bool align_1d = false;
    int align_max_iter = 10;
    double max_epi_length_optim = 2.0;
    size_t max_epi_search_steps = 100;
    bool subpix_refinement = true;
    bool epi_search_edgelet_filtering = true;
    bool scan_on_unit_sphere = true;
    double epi_search_edgelet_max_angle = 0.7;
    bool use_affine_warp_ = true;
    bool affine_est_offset_ = true;
    bool affine_est_gain_ = false;
    double max_patch_diff_ratio = 2.0;
  } options_;
  enum class MatchResult {  // ref: matcher.h:56-68
    kSuccess, kFailScore, kFailTriangulation, kFailVisibility, kFailWarp, kFailAlignment,
    kFailRange, kFailAngle, kFailCloseView, kFailLock, kFailTooFar
  };
  uint8_t patch_[kPatchSize * kPatchSize];
  uint8_t patch_with_border_[(kPatchSize + 2) * (kPatchSize + 2)];
  double A_cur_ref_[2][2] = {{0, 0}, {0, 0}};
  V2 epi_image_{0, 0};
  double epi_length_pyramid_ = 0;
  double h_inv_ = 0;
  int search_level_ = 0;
  bool reject_ = false;
  V2 px_cur_{0, 0};
  V3 f_cur_{0, 0, 0};

  // c6. ref: matcher.cpp:31-141
  MatchResult findMatchDirect(const MatchFrame& ref_frame, const MatchFrame& cur_frame, const SE3& T_cur_ref,
                              const FeatureRef& ref_ftr, const double& ref_depth, V2& px_cur) {
    const int pxi0 = int(ref_ftr.px.x) / (1 << ref_ftr.level);  // Vector2i / int (matcher.cpp:38)
    const int pxi1 = int(ref_ftr.px.y) / (1 << ref_ftr.level);
    const int boundary = kHalfPatchSize + 2;
    if (pxi0 < boundary || pxi1 < boundary
        || pxi0 >= static_cast<int>(ref_frame.cam.width / (1 << ref_ftr.level)) - boundary
        || pxi1 >= static_cast<int>(ref_frame.cam.height / (1 << ref_ftr.level)) - boundary)
      return MatchResult::kFailVisibility;
    getWarpMatrixAffine(ref_frame.cam, cur_frame.cam, ref_ftr.px, ref_ftr.f, ref_depth, T_cur_ref, ref_ftr.level, A_cur_ref_);
    search_level_ = getBestSearchLevel(A_cur_ref_, int(ref_frame.img_pyr.size()) - 1);
    if (!warpAffine(A_cur_ref_, ref_frame.img_pyr[ref_ftr.level], ref_ftr.px, ref_ftr.level, search_level_, kHalfPatchSize + 1, patch_with_border_))
      return MatchResult::kFailWarp;
    createPatchFromPatchWithBorder(patch_with_border_, kPatchSize, patch_);
    V2 px_scaled{px_cur.x / (1 << search_level_), px_cur.y / (1 << search_level_)};
    const V2 px_scaled_start = px_scaled;
    if (isEdgelet(ref_ftr.type)) {
      V2 dir_cur{A_cur_ref_[0][0] * ref_ftr.grad.x + A_cur_ref_[0][1] * ref_ftr.grad.y,
                 A_cur_ref_[1][0] * ref_ftr.grad.x + A_cur_ref_[1][1] * ref_ftr.grad.y};
      dir_cur = normalized(dir_cur);
      if (align1D(cur_frame.img_pyr[search_level_], dir_cur, patch_with_border_, patch_, options_.align_max_iter,
                  options_.affine_est_offset_, options_.affine_est_gain_, &px_scaled, &h_inv_)) {
        const double ddx = px_scaled.x - px_scaled_start.x, ddy = px_scaled.y - px_scaled_start.y;
        if (std::sqrt(ddx * ddx + ddy * ddy) > options_.max_patch_diff_ratio * kPatchSize) return MatchResult::kFailTooFar;
        px_cur = {px_scaled.x * (1 << search_level_), px_scaled.y * (1 << search_level_)};
        px_cur_ = px_cur;
        f_cur_ = normalized(cur_frame.cam.backProject3(px_cur_));
        return MatchResult::kSuccess;
      }
    } else {
      const bool res = align2D(cur_frame.img_pyr[search_level_], patch_with_border_, patch_, options_.align_max_iter,
                               options_.affine_est_offset_, options_.affine_est_gain_, px_scaled);
      if (res) {
        const double ddx = px_scaled.x - px_scaled_start.x, ddy = px_scaled.y - px_scaled_start.y;
        if (std::sqrt(ddx * ddx + ddy * ddy) > options_.max_patch_diff_ratio * kPatchSize) return MatchResult::kFailTooFar;
        px_cur = {px_scaled.x * (1 << search_level_), px_scaled.y * (1 << search_level_)};
        px_cur_ = px_cur;
        f_cur_ = normalized(cur_frame.cam.backProject3(px_cur_));
        return MatchResult::kSuccess;
      }
    }
    return MatchResult::kFailAlignment;
  }

  // ref: matcher.cpp:262-289
  MatchResult findLocalMatch(const MatchFrame& frame, const V2& direction, const int patch_level, V2& px_cur) {
    V2 px_scaled{px_cur.x / (1 << patch_level), px_cur.y / (1 << patch_level)};
    bool res;
    if (options_.align_1d)
      res = align1D(frame.img_pyr[patch_level], direction, patch_with_border_, patch_, options_.align_max_iter,
                    options_.affine_est_offset_, options_.affine_est_gain_, &px_scaled, &h_inv_);
    else
      res = align2D(frame.img_pyr[patch_level], patch_with_border_, patch_, options_.align_max_iter,
                    options_.affine_est_offset_, options_.affine_est_gain_, px_scaled);
    if (!res) return MatchResult::kFailAlignment;
    px_cur = {px_scaled.x * (1 << patch_level), px_scaled.y * (1 << patch_level)};
    return MatchResult::kSuccess;
  }

  // ref: matcher.cpp:292-312
  bool updateZMSSD(const MatchFrame& frame, const int pxi[2], const int patch_level, const ZMSSD& patch_score, int* zmssd_best) {
    const Img& im = frame.img_pyr[patch_level];
    const uint8_t* cur_patch_ptr = im.data + (pxi[1] - kHalfPatchSize) * im.step + (pxi[0] - kHalfPatchSize);
    const int zmssd = patch_score.computeScore(cur_patch_ptr, im.step);
    if (zmssd < *zmssd_best) {
      *zmssd_best = zmssd;
      return true;
    }
    return false;
  }
  // ref: matcher.cpp:314-322
  bool isPatchWithinImage(const MatchFrame& frame, const int pxi[2], const int patch_level) {
    return !(pxi[0] < kPatchSize || pxi[1] < kPatchSize
             || pxi[0] >= (static_cast<int>(frame.cam.width / (1 << patch_level)) - kPatchSize)
             || pxi[1] >= (static_cast<int>(frame.cam.height / (1 << patch_level)) - kPatchSize));
  }

  // ref: matcher.cpp:340-413
  void scanEpipolarUnitPlane(const MatchFrame& frame, const V3& A, const V3& B, const V3& C, const ZMSSD& patch_score,
                             const int patch_level, V2* image_best, int* zmssd_best) {
    size_t n_steps = epi_length_pyramid_ / 0.7;
    V2 step{(A.x / A.z - B.x / B.z) / n_steps, (A.y / A.z - B.y / B.z) / n_steps};
    if (n_steps > options_.max_epi_search_steps) n_steps = options_.max_epi_search_steps;
    const V2 uv_C{C.x / C.z, C.y / C.z};
    V2 uv = uv_C;
    V2 uv_best = uv;
    bool forward = true;
    int last_checked_pxi[2] = {0, 0};
    for (size_t i = 0; i < n_steps; ++i, uv.x += step.x, uv.y += step.y) {
      const V2 px = frame.cam.project3({uv.x, uv.y, 1.0});
      const int pxi[2] = {int(px.x / (1 << patch_level) + 0.5), int(px.y / (1 << patch_level) + 0.5)};
      if (pxi[0] == last_checked_pxi[0] && pxi[1] == last_checked_pxi[1]) continue;
      last_checked_pxi[0] = pxi[0]; last_checked_pxi[1] = pxi[1];
      if (!isPatchWithinImage(frame, pxi, patch_level)) {
        if (forward) {
          i = n_steps * 0.5;
          step = {-step.x, -step.y};
          uv = uv_C;
          forward = false;
          continue;
        } else {
          break;
        }
      }
      if (updateZMSSD(frame, pxi, patch_level, patch_score, zmssd_best)) uv_best = uv;
      if (forward && i > n_steps * 0.5) {
        step = {-step.x, -step.y};
        uv = uv_C;
        forward = false;
      }
    }
    *image_best = frame.cam.project3({uv_best.x, uv_best.y, 1.0});
  }

  // Eigen::AngleAxisd::toRotationMatrix() * v (kindr AngleAxis::rotate, angle-axis-inl.h:192-195)
  static V3 angleAxisRotate(const V3& axis, double angle, const V3& v) {
    M3 res;
    const double s = std::sin(angle), c = std::cos(angle);
    const V3 sin_axis = axis * s;
    const V3 cos1_axis = axis * (1.0 - c);
    double tmp;
    tmp = cos1_axis.x * axis.y; res.m[0][1] = tmp - sin_axis.z; res.m[1][0] = tmp + sin_axis.z;
    tmp = cos1_axis.x * axis.z; res.m[0][2] = tmp + sin_axis.y; res.m[2][0] = tmp - sin_axis.y;
    tmp = cos1_axis.y * axis.z; res.m[1][2] = tmp - sin_axis.x; res.m[2][1] = tmp + sin_axis.x;
    res.m[0][0] = cos1_axis.x * axis.x + c;
    res.m[1][1] = cos1_axis.y * axis.y + c;
    res.m[2][2] = cos1_axis.z * axis.z + c;
    return res * v;
  }

  // ref: matcher.cpp:415-488
  void scanEpipolarUnitSphere(const MatchFrame& frame, const V3& A, const V3& B, const V3& C, const ZMSSD& patch_score,
                              const int patch_level, V2* image_best, int* zmssd_best) {
    size_t n_steps = epi_length_pyramid_ / 0.7;
    n_steps = n_steps > options_.max_epi_search_steps ? options_.max_epi_search_steps : n_steps;
    const size_t half_steps = n_steps / 2;
    const V3 f_A = normalized(A);
    const V3 f_B = normalized(B);
    const double step = std::acos(dot(f_A, f_B)) / n_steps;
    const V3 axis = normalized(cross(f_B, f_A));
    const V3 f_C = normalized(C);
    V3 f = f_C;
    V3 f_best = f_C;
    int last_checked_pxi[2] = {0, 0};
    for (size_t i = 0; i < n_steps; i++) {
      double angle = 0.0;
      if (i < half_steps) angle = i * step;
      else angle = (i - half_steps) * (-step);
      f = angleAxisRotate(axis, angle, f_C);
      const V2 px = frame.cam.project3(f);
      const int pxi[2] = {int(px.x / (1 << patch_level) + 0.5), int(px.y / (1 << patch_level) + 0.5)};
      if (pxi[0] == last_checked_pxi[0] && pxi[1] == last_checked_pxi[1]) continue;
      last_checked_pxi[0] = pxi[0]; last_checked_pxi[1] = pxi[1];
      if (!isPatchWithinImage(frame, pxi, patch_level)) {
        if (i < half_steps) {
          i = half_steps;
          continue;
        } else {
          break;
        }
      }
      if (updateZMSSD(frame, pxi, patch_level, patch_score, zmssd_best)) f_best = f;
    }
    *image_best = frame.cam.project3(f_best);
  }

  // matcher_utils::depthFromTriangulation — ref: matcher.cpp:492-505
  static MatchResult depthFromTriangulation(const SE3& T_search_ref, const V3& f_ref, const V3& f_cur, double* depth) {
    const V3 a0 = quatRotate(T_search_ref.q, f_ref);
    const V3 a1 = f_cur;
    const double AtA[2][2] = {{dot(a0, a0), dot(a0, a1)}, {dot(a1, a0), dot(a1, a1)}};
    const double det = AtA[0][0] * AtA[1][1] - AtA[1][0] * AtA[0][1];
    if (det < 0.000001) return MatchResult::kFailTriangulation;
    const double invdet = 1.0 / det;
    const double Inv[2][2] = {{AtA[1][1] * invdet, -AtA[0][1] * invdet}, {-AtA[1][0] * invdet, AtA[0][0] * invdet}};
    // depth2 = -AtA^-1 * A^T * t, Eigen evaluates ((-AtA^-1) * A^T) * t
    const V3 t = T_search_ref.t;
    const V3 row0 = a0 * (-Inv[0][0]) + a1 * (-Inv[0][1]);
    (*depth) = std::fabs(dot(row0, t));
    return MatchResult::kSuccess;
  }

  // c7. ref: matcher.cpp:157-241
  MatchResult findEpipolarMatchDirect(const MatchFrame& ref_frame, const MatchFrame& cur_frame, const SE3& T_cur_ref,
                                      const FeatureRef& ref_ftr, const double d_estimate_inv, const double d_min_inv,
                                      const double d_max_inv, double& depth) {
    int zmssd_best = ZMSSD::threshold();
    const V3 Rf = quatRotate(T_cur_ref.q, ref_ftr.f);
    const V3 A = Rf + T_cur_ref.t * d_min_inv;
    const V3 B = Rf + T_cur_ref.t * d_max_inv;
    const V2 px_A = cur_frame.cam.project3(A);
    const V2 px_B = cur_frame.cam.project3(B);
    epi_image_ = {px_A.x - px_B.x, px_A.y - px_B.y};
    getWarpMatrixAffine(ref_frame.cam, cur_frame.cam, ref_ftr.px, ref_ftr.f, 1.0 / std::max(0.000001, d_estimate_inv),
                        T_cur_ref, ref_ftr.level, A_cur_ref_);
    reject_ = false;
    if (isEdgelet(ref_ftr.type) && options_.epi_search_edgelet_filtering) {
      const V2 grad_cur = normalized(V2{A_cur_ref_[0][0] * ref_ftr.grad.x + A_cur_ref_[0][1] * ref_ftr.grad.y,
                                        A_cur_ref_[1][0] * ref_ftr.grad.x + A_cur_ref_[1][1] * ref_ftr.grad.y});
      const V2 en = normalized(epi_image_);
      const double cosangle = std::fabs(grad_cur.x * en.x + grad_cur.y * en.y);
      if (cosangle < options_.epi_search_edgelet_max_angle) {
        reject_ = true;
        return MatchResult::kFailAngle;
      }
    }
    search_level_ = getBestSearchLevel(A_cur_ref_, int(ref_frame.img_pyr.size()) - 1);
    epi_length_pyramid_ = std::sqrt(epi_image_.x * epi_image_.x + epi_image_.y * epi_image_.y) / (1 << search_level_);
    const V2 epi_dir_image = normalized(epi_image_);
    if (!warpAffine(A_cur_ref_, ref_frame.img_pyr[ref_ftr.level], ref_ftr.px, ref_ftr.level, search_level_, kHalfPatchSize + 1, patch_with_border_))
      return MatchResult::kFailWarp;
    createPatchFromPatchWithBorder(patch_with_border_, kPatchSize, patch_);

    if (epi_length_pyramid_ < 2.0) {
      px_cur_ = {(px_A.x + px_B.x) / 2.0, (px_A.y + px_B.y) / 2.0};
      const MatchResult res = findLocalMatch(cur_frame, epi_dir_image, search_level_, px_cur_);
      if (res != MatchResult::kSuccess) return res;
      f_cur_ = normalized(cur_frame.cam.backProject3(px_cur_));
      return depthFromTriangulation(T_cur_ref, ref_ftr.f, f_cur_, &depth);
    }
    const ZMSSD patch_score(patch_);
    const V3 C = Rf + T_cur_ref.t * d_estimate_inv;
    if (options_.scan_on_unit_sphere) scanEpipolarUnitSphere(cur_frame, A, B, C, patch_score, search_level_, &px_cur_, &zmssd_best);
    else scanEpipolarUnitPlane(cur_frame, A, B, C, patch_score, search_level_, &px_cur_, &zmssd_best);
    if (zmssd_best < ZMSSD::threshold()) {
      if (options_.subpix_refinement) {
        const MatchResult res = findLocalMatch(cur_frame, epi_dir_image, search_level_, px_cur_);
        if (res != MatchResult::kSuccess) return res;
      }
      f_cur_ = normalized(cur_frame.cam.backProject3(px_cur_));
      return depthFromTriangulation(T_cur_ref, ref_ftr.f, f_cur_, &depth);
    }
    return MatchResult::kFailScore;
  }
};

}  // namespace orc
