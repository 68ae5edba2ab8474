// Device functions of the patch matcher, shared by the (c) kernels in matcher.cu and the (d) seed-update kernel in
// depth_filter.cu. One 8-lane group works on one feature: lane r owns row r of the 8x8 patch, group-wide sums use
// __shfl_xor_sync with the group's own mask so the four groups of a warp may diverge (different iteration counts,
// different epipolar scan lengths).
//
// ref: src/svo_direct/src/patch_warp.cpp:20-60 (getWarpMatrixAffine), :97-110 (getBestSearchLevel), :112-156 (warpAffine)
//      src/svo_direct/include/svo/direct/patch_utils.h:18-30, patch_score.h:43-109,264-283 (ZMSSD)
//      src/svo_direct/src/feature_alignment.cpp:31-209 (align1D), :212-391 (align2D)
//      src/svo_direct/src/matcher.cpp:31-141 (findMatchDirect), :157-241 (findEpipolarMatchDirect), :262-338,
//      :340-413 (scanEpipolarUnitPlane), :415-488 (scanEpipolarUnitSphere), :492-505 (depthFromTriangulation)
#pragma once
#include "common.cuh"

namespace svo_dev {

constexpr int kGroup = 8;            // lanes per feature
constexpr int kPatchBytes = 112;     // the 10x10 patch (100 bytes rounded up to 16)
constexpr int kRedStride = 68;       // floats per ordered-sum row: 64 terms + 4 of padding (conflict-free 128-bit reads)
constexpr int kRedRows = 4;          // sums formed side by side (lanes 0..3 of the group)
constexpr int kPwbPitch = kPatchBytes + kRedRows * kRedStride * 4;  // bytes of shared memory per group: patch + ordered-sum scratch

enum MatchResult {  // svo::Matcher::MatchResult, matcher.h:56-68
  kSuccess = 0, kFailScore, kFailTriangulation, kFailVisibility, kFailWarp, kFailAlignment, kFailRange, kFailAngle,
  kFailCloseView, kFailLock, kFailTooFar
};
enum FeatureType {  // svo::FeatureType, src/svo_common/include/svo/common/types.h:60-73
  kEdgeletSeed = 0, kCornerSeed = 1, kMapPointSeed = 2, kEdgeletSeedConverged = 3, kCornerSeedConverged = 4,
  kMapPointSeedConverged = 5, kEdgelet = 6, kCorner = 7, kMapPoint = 8, kFixedLandmark = 9, kOutlier = 10
};
SVO_HD bool isEdgeletType(int t) { return t == kEdgelet || t == kEdgeletSeed || t == kEdgeletSeedConverged; }

struct Group {
  unsigned mask;  // the 8 lanes of this group
  int r;          // lane within the group = patch row
};
SVO_D Group makeGroup() {
  const int lane = threadIdx.x & 31;
  Group g;
  g.r = lane & (kGroup - 1);
  g.mask = 0xFFu << (lane & ~(kGroup - 1));
  return g;
}
template <class T>
SVO_D T groupSum(const Group& g, T v) {
  v += __shfl_xor_sync(g.mask, v, 1);
  v += __shfl_xor_sync(g.mask, v, 2);
  v += __shfl_xor_sync(g.mask, v, 4);
  return v;
}
SVO_D bool groupAny(const Group& g, bool p) { return (__ballot_sync(g.mask, p) & g.mask) != 0u; }

struct ImgView {
  const uint8_t* data;
  int cols, rows, pitch;
};
SVO_D ImgView levelView(const PyrView& p, int frame, int level) {
  return ImgView{p.level(frame, level), p.cols[level], p.rows[level], p.pitch[level]};
}

// ---- c1 ------------------------------------------------------------------------------------------------------
SVO_D void getWarpMatrixAffine(const svo_camera& cam_ref, const svo_camera& cam_cur, double pxr_x, double pxr_y, const V3d& f_ref,
                               double depth_ref, const SE3d& T_cur_ref, int level_ref, double A[2][2]) {
  const int kHalf = 5;
  const V3d xyz_ref = f_ref * depth_ref;
  V3d du = camBackProject3(cam_ref, pxr_x + double(kHalf) * (1 << level_ref), pxr_y);
  V3d dv = camBackProject3(cam_ref, pxr_x, pxr_y + double(kHalf) * (1 << level_ref));
  du = du * xyz_ref.z;
  dv = dv * xyz_ref.z;
  const V2d pc = camProject3(cam_cur, se3Apply(T_cur_ref, xyz_ref));
  const V2d pdu = camProject3(cam_cur, se3Apply(T_cur_ref, du));
  const V2d pdv = camProject3(cam_cur, se3Apply(T_cur_ref, dv));
  A[0][0] = (pdu.x - pc.x) / kHalf; A[1][0] = (pdu.y - pc.y) / kHalf;
  A[0][1] = (pdv.x - pc.x) / kHalf; A[1][1] = (pdv.y - pc.y) / kHalf;
}
SVO_D int getBestSearchLevel(const double A[2][2], int max_level) {
  int sl = 0;
  double D = A[0][0] * A[1][1] - A[1][0] * A[0][1];
  while (D > 3.0 && sl < max_level) { sl += 1; D *= 0.25; }
  return sl;
}
// warpAffine with halfpatch_size = 5: the group fills pwb[100]; float arithmetic is spelled with round-to-nearest
// intrinsics so no FMA is formed and the truncated u8 values equal the reference's.
SVO_D bool warpAffine10(const Group& g, const double A[2][2], const ImgView& img, double pxr_x, double pxr_y, int level_ref,
                        int search_level, uint8_t* pwb) {
  const double det = A[0][0] * A[1][1] - A[1][0] * A[0][1];
  const double invdet = 1.0 / det;
  const float sl = (float)(1 << search_level);
  const float a00 = __fmul_rn((float)(A[1][1] * invdet), sl), a01 = __fmul_rn((float)(-A[0][1] * invdet), sl);
  const float a10 = __fmul_rn((float)(-A[1][0] * invdet), sl), a11 = __fmul_rn((float)(A[0][0] * invdet), sl);
  if (a00 != a00) return false;
  const float lr = (float)(1 << level_ref);
  const float p0 = __fdiv_rn((float)pxr_x, lr), p1 = __fdiv_rn((float)pxr_y, lr);
  bool bad = false;
  __syncwarp(g.mask);  // every lane is done reading the previous patch in pwb
  for (int i = g.r; i < 100; i += kGroup) {
    const int yy = i / 10, xx = i - yy * 10;
    const float x = (float)(xx - 5), y = (float)(yy - 5);
    const float px0 = __fadd_rn(__fadd_rn(__fmul_rn(a00, x), __fmul_rn(a01, y)), p0);
    const float px1 = __fadd_rn(__fadd_rn(__fmul_rn(a10, x), __fmul_rn(a11, y)), p1);
    const int xi = (int)floorf(px0), yi = (int)floorf(px1);
    if (xi < 0 || yi < 0 || xi + 1 >= img.cols || yi + 1 >= img.rows || !(px0 == px0) || !(px1 == px1)) {
      bad = true;
      continue;
    }
    const float sx = __fsub_rn(px0, (float)xi), sy = __fsub_rn(px1, (float)yi);
    const float w00 = __fmul_rn(__fsub_rn(1.0f, sx), __fsub_rn(1.0f, sy));
    const float w01 = __fmul_rn(__fsub_rn(1.0f, sx), sy);
    const float w10 = __fmul_rn(sx, __fsub_rn(1.0f, sy));
    const float w11 = __fsub_rn(__fsub_rn(__fsub_rn(1.0f, w00), w01), w10);
    const uint8_t* p = img.data + (size_t)yi * img.pitch + xi;
    const float v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w00, (float)p[0]), __fmul_rn(w01, (float)p[img.pitch])),
                                        __fmul_rn(w10, (float)p[1])), __fmul_rn(w11, (float)p[img.pitch + 1]));
    pwb[i] = (uint8_t)__float2int_rz(v);
  }
  __syncwarp(g.mask);
  return !groupAny(g, bad);
}

// ---- small float inverses (Eigen cofactor forms) -------------------------------------------------------------
SVO_D void inverse3f(const float m[3][3], float r[3][3]) {
#define COF3(i, j) (m[(i + 1) % 3][(j + 1) % 3] * m[(i + 2) % 3][(j + 2) % 3] - m[(i + 1) % 3][(j + 2) % 3] * m[(i + 2) % 3][(j + 1) % 3])
  const float c00 = COF3(0, 0), c10 = COF3(1, 0), c20 = COF3(2, 0);
  const float det = c00 * m[0][0] + c10 * m[1][0] + c20 * m[2][0];
  const float invdet = 1.0f / det;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) r[i][j] = COF3(j, i) * invdet;
#undef COF3
}
SVO_D float det3f(const float m[4][4], int r0, int r1, int r2, int c0, int c1, int c2) {
  return m[r0][c0] * (m[r1][c1] * m[r2][c2] - m[r1][c2] * m[r2][c1]) - m[r0][c1] * (m[r1][c0] * m[r2][c2] - m[r1][c2] * m[r2][c0]) +
         m[r0][c2] * (m[r1][c0] * m[r2][c1] - m[r1][c1] * m[r2][c0]);
}
SVO_D void inverse4f(const float m[4][4], float r[4][4]) {
  float cof[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r0 = (i == 0) ? 1 : 0, r1 = (i <= 1) ? 2 : 1, r2 = (i <= 2) ? 3 : 2;
      const int c0 = (j == 0) ? 1 : 0, c1 = (j <= 1) ? 2 : 1, c2 = (j <= 2) ? 3 : 2;
      const float d = det3f(m, r0, r1, r2, c0, c1, c2);
      cof[i][j] = ((i + j) & 1) ? -d : d;
    }
  const float det = m[0][0] * cof[0][0] + m[0][1] * cof[0][1] + m[0][2] * cof[0][2] + m[0][3] * cof[0][3];
  const float invdet = 1.0f / det;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) r[i][j] = cof[j][i] * invdet;
}

// 9 bytes x0..x0+8 of a 4-byte-aligned row: bytes 0..7 in (a,b), byte 8 returned in c's low byte.
SVO_D void loadRow9(const uint8_t* row, int x0, unsigned& a, unsigned& b, unsigned& c) {
  const unsigned* w = reinterpret_cast<const unsigned*>(row + (x0 & ~3));
  const unsigned w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
  const unsigned sh = (x0 & 3) * 8;
  a = __funnelshift_r(w0, w1, sh);
  b = __funnelshift_r(w1, w2, sh);
  c = __funnelshift_r(w2, w3, sh) & 0xffu;
}
SVO_D unsigned byte9(unsigned a, unsigned b, unsigned c, int i) { return i < 4 ? byteOf(a, i) : (i < 8 ? byteOf(b, i - 4) : c); }

// ---- c3/c4: align2D / align1D --------------------------------------------------------------------------------
// Bit-exactness: the reference accumulates H and Jres in FLOAT, pixel by pixel in raster order. Lane r computes the
// interpolated intensities / residuals of its patch row in parallel, then the running sums are carried through the
// rows in order (owner lane adds its 8 pixels, result is broadcast with a shuffle), so every float operation happens in
// the reference's order. The translation unit is compiled with -fmad=false so no multiply-add is contracted.
// Ordered float sums. Each lane holds 8 consecutive terms (its patch row) of each of NK sums over the 64 patch pixels. The
// terms go through the group's shared-memory scratch so that lane k owns sum k and adds its 64 terms one by one in raster
// order — the NK chains run side by side on NK lanes instead of one after the other on the row's owner lane, and every
// float addition still happens in the reference's order. SUB: acc = acc - term (Jres), else acc = acc + term (H).
// term(k, x): the x-th term of sum k in this lane's row; `first` = the sum handled by lane 0 (sums first..first+NK-1).
template <int NK, bool SUB, class F>
SVO_D void orderedSums(const Group& g, float* red, F term, float (&acc)[NK]) {
  static_assert(NK <= kRedRows, "one sum per lane, kRedRows rows of scratch");
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    float4 lo, hi;
    lo.x = term(k, 0); lo.y = term(k, 1); lo.z = term(k, 2); lo.w = term(k, 3);
    hi.x = term(k, 4); hi.y = term(k, 5); hi.z = term(k, 6); hi.w = term(k, 7);
    float4* dst = reinterpret_cast<float4*>(red + k * kRedStride + g.r * 8);
    dst[0] = lo; dst[1] = hi;
  }
  __syncwarp(g.mask);
  float a = 0.f;
  if (g.r < NK) {
    const float4* src = reinterpret_cast<const float4*>(red + g.r * kRedStride);
#pragma unroll 4  // the chain of 64 dependent additions is latency bound; a short loop body keeps the kernels' code small
    for (int i = 0; i < 16; ++i) {
      const float4 v = src[i];
      if (SUB) { a = a - v.x; a = a - v.y; a = a - v.z; a = a - v.w; }
      else { a = a + v.x; a = a + v.y; a = a + v.z; a = a + v.w; }
    }
  }
#pragma unroll
  for (int k = 0; k < NK; ++k) acc[k] = __shfl_sync(g.mask, a, k, kGroup);
  __syncwarp(g.mask);  // the scratch may be rewritten
}

// pwb: the group's 10x10 patch in shared memory, followed by the ordered-sum scratch. px in/out (level px). Returns converged.
SVO_D bool align2D(const Group& g, const ImgView& img, uint8_t* pwb, int n_iter, bool est_offset, bool est_gain,
                   double& px_x, double& px_y) {
  const uint8_t* it = pwb + (g.r + 1) * 10 + 1;
  float* red = reinterpret_cast<float*>(pwb + kPatchBytes);
  float rdx[8], rdy[8], rref[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    rdx[x] = (float)(0.5 * ((int)it[x + 1] - (int)it[x - 1]));    // feature_alignment.cpp:252-253
    rdy[x] = (float)(0.5 * ((int)it[x + 10] - (int)it[x - 10]));
    rref[x] = (float)it[x];
  }
  const float J2 = est_offset ? 1.0f : 0.0f;
  // H += J*J^T pixel by pixel in raster order (:261); 00 01 02 03 11 12 13 22 | 23 33
  float h[10];
  {
    float ha[4], hb[4], hc[2];
    orderedSums<4, false>(g, red, [&](int k, int x) {
      const float J0 = rdx[x], J1 = rdy[x], J3 = est_gain ? -1.0f * rref[x] : 0.0f;
      return k == 0 ? J0 * J0 : k == 1 ? J0 * J1 : k == 2 ? J0 * J2 : J0 * J3;
    }, ha);
    orderedSums<4, false>(g, red, [&](int k, int x) {
      const float J1 = rdy[x], J3 = est_gain ? -1.0f * rref[x] : 0.0f;
      return k == 0 ? J1 * J1 : k == 1 ? J1 * J2 : k == 2 ? J1 * J3 : J2 * J2;
    }, hb);
    orderedSums<2, false>(g, red, [&](int k, int x) {
      const float J3 = est_gain ? -1.0f * rref[x] : 0.0f;
      return k == 0 ? J2 * J3 : J3 * J3;
    }, hc);
#pragma unroll
    for (int k = 0; k < 4; ++k) { h[k] = ha[k]; h[4 + k] = hb[k]; }
    h[8] = hc[0]; h[9] = hc[1];
  }
  float H[4][4] = {{h[0], h[1], h[2], h[3]}, {h[1], h[4], h[5], h[6]}, {h[2], h[5], h[7], h[8]}, {h[3], h[6], h[8], h[9]}};
  if (!est_offset) H[2][2] = 1.0f;
  if (!est_gain) H[3][3] = 1.0f;
  float Hinv[4][4];
  inverse4f(H, Hinv);
  float mean_diff = 0.f, alpha = 1.0f;
  float u = (float)px_x, v = (float)px_y;
  const float min_update_squared = (float)(0.03 * 0.03);
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floorf(u), v_r = (int)floorf(v);
    if (u_r < 4 || v_r < 4 || u_r >= img.cols - 4 || v_r >= img.rows - 4) break;
    if (u != u || v != v) return false;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy)), wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy), wBR = sx * sy;
    const uint8_t* row = img.data + (size_t)(v_r + g.r - 4) * img.pitch;
    unsigned a0, b0, c0, a1, b1, c1;
    loadRow9(row, u_r - 4, a0, b0, c0);
    loadRow9(row + img.pitch, u_r - 4, a1, b1, c1);
    float res[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const float sp = wTL * (float)byte9(a0, b0, c0, x) + wTR * (float)byte9(a0, b0, c0, x + 1) +
                       wBL * (float)byte9(a1, b1, c1, x) + wBR * (float)byte9(a1, b1, c1, x + 1);
      res[x] = sp - alpha * rref[x] + mean_diff;  // :322-323
    }
    // Jres -= J * res in raster order (:325-328); the sums of disabled parameters are not formed (they are zeroed below)
    float j[4];
    orderedSums<4, true>(g, red, [&](int k, int x) {
      return k == 0 ? res[x] * rdx[x] : k == 1 ? res[x] * rdy[x] : k == 2 ? (est_offset ? res[x] : 0.f)
             : (est_gain ? (-1.0f * res[x]) * rref[x] : 0.f);
    }, j);
    if (!est_offset) j[2] = 0.f;
    if (!est_gain) j[3] = 0.f;
    const float up0 = Hinv[0][0] * j[0] + Hinv[0][1] * j[1] + Hinv[0][2] * j[2] + Hinv[0][3] * j[3];
    const float up1 = Hinv[1][0] * j[0] + Hinv[1][1] * j[1] + Hinv[1][2] * j[2] + Hinv[1][3] * j[3];
    const float up2 = Hinv[2][0] * j[0] + Hinv[2][1] * j[1] + Hinv[2][2] * j[2] + Hinv[2][3] * j[3];
    const float up3 = Hinv[3][0] * j[0] + Hinv[3][1] * j[1] + Hinv[3][2] * j[2] + Hinv[3][3] * j[3];
    u += up0; v += up1; mean_diff += up2; alpha += up3;
    if (up0 * up0 + up1 * up1 < min_update_squared) { converged = true; break; }
  }
  px_x = u; px_y = v;
  return converged;
}

SVO_D bool align1D(const Group& g, const ImgView& img, double dir_x, double dir_y, uint8_t* pwb, int n_iter, bool est_offset,
                   bool est_gain, double& px_x, double& px_y, double* h_inv) {
  const uint8_t* it = pwb + (g.r + 1) * 10 + 1;
  float* red = reinterpret_cast<float*>(pwb + kPatchBytes);
  float rdv[8], rref[8];
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    const float dx = (float)it[x + 1] - (float)it[x - 1];         // feature_alignment.cpp:63-65
    const float dy = (float)it[x + 10] - (float)it[x - 10];
    rdv[x] = (float)(0.5f * (dir_x * dx + dir_y * dy));
    rref[x] = (float)it[x];
  }
  const float J1 = est_offset ? 1.0f : 0.0f;
  float h[6];  // 00 01 02 11 | 12 22
  {
    float ha[4], hb[2];
    orderedSums<4, false>(g, red, [&](int k, int x) {
      const float J0 = rdv[x], J2 = est_gain ? -1.0f * rref[x] : 0.0f;
      return k == 0 ? J0 * J0 : k == 1 ? J0 * J1 : k == 2 ? J0 * J2 : J1 * J1;
    }, ha);
    orderedSums<2, false>(g, red, [&](int k, int x) {
      const float J2 = est_gain ? -1.0f * rref[x] : 0.0f;
      return k == 0 ? J1 * J2 : J2 * J2;
    }, hb);
#pragma unroll
    for (int k = 0; k < 4; ++k) h[k] = ha[k];
    h[4] = hb[0]; h[5] = hb[1];
  }
  float H[3][3] = {{h[0], h[1], h[2]}, {h[1], h[3], h[4]}, {h[2], h[4], h[5]}};
  if (!est_offset) H[1][1] = 1.0f;
  if (!est_gain) H[2][2] = 1.0f;
  if (h_inv) *h_inv = 1.0 / H[0][0] * 8 * 8;
  float Hinv[3][3];
  inverse3f(H, Hinv);
  float mean_diff = 0.f, alpha = 1.0f;
  float u = (float)px_x, v = (float)px_y;
  const float min_update_squared = (float)(0.03 * 0.03);
  bool converged = false;
  for (int iter = 0; iter < n_iter; ++iter) {
    const int u_r = (int)floorf(u), v_r = (int)floorf(v);
    if (u_r < 4 || v_r < 4 || u_r >= img.cols - 4 || v_r >= img.rows - 4) break;
    if (u != u || v != v) return false;
    const float sx = u - u_r, sy = v - v_r;
    const float wTL = (float)((1.0 - sx) * (1.0 - sy)), wTR = (float)(sx * (1.0 - sy));
    const float wBL = (float)((1.0 - sx) * sy), wBR = sx * sy;
    const uint8_t* row = img.data + (size_t)(v_r + g.r - 4) * img.pitch;
    unsigned a0, b0, c0, a1, b1, c1;
    loadRow9(row, u_r - 4, a0, b0, c0);
    loadRow9(row + img.pitch, u_r - 4, a1, b1, c1);
    float res[8];
#pragma unroll
    for (int x = 0; x < 8; ++x) {
      const float ci = wTL * (float)byte9(a0, b0, c0, x) + wTR * (float)byte9(a0, b0, c0, x + 1) +
                       wBL * (float)byte9(a1, b1, c1, x) + wBR * (float)byte9(a1, b1, c1, x + 1);
      res[x] = ci - alpha * rref[x] + mean_diff;  // :137-139
    }
    float j[3];
    orderedSums<3, true>(g, red, [&](int k, int x) {
      return k == 0 ? res[x] * rdv[x] : k == 1 ? (est_offset ? res[x] : 0.f) : (est_gain ? (-1.0f * res[x]) * rref[x] : 0.f);
    }, j);
    if (!est_offset) j[1] = 0.f;
    if (!est_gain) j[2] = 0.f;
    const float up0 = Hinv[0][0] * j[0] + Hinv[0][1] * j[1] + Hinv[0][2] * j[2];
    const float up1 = Hinv[1][0] * j[0] + Hinv[1][1] * j[1] + Hinv[1][2] * j[2];
    const float up2 = Hinv[2][0] * j[0] + Hinv[2][1] * j[1] + Hinv[2][2] * j[2];
    u = (float)(u + up0 * dir_x);
    v = (float)(v + up0 * dir_y);
    mean_diff += up1;
    alpha += up2;
    if (up0 * up0 < min_update_squared) { converged = true; break; }
  }
  px_x = u; px_y = v;
  return converged;
}

// ---- matcher state -------------------------------------------------------------------------------------------
struct MatchState {  // the public Matcher members (matcher.h:70-79)
  double A[2][2];
  double epi_x, epi_y;
  double epi_length_pyramid;
  double h_inv;
  double px_x, px_y;
  V3d f_cur;
  int search_level;
  int reject;
};
SVO_D void initMatchState(MatchState& m) {
  m.A[0][0] = m.A[0][1] = m.A[1][0] = m.A[1][1] = 0.0;
  m.epi_x = m.epi_y = m.epi_length_pyramid = m.h_inv = m.px_x = m.px_y = 0.0;
  m.f_cur = V3d{0, 0, 0};
  m.search_level = 0;
  m.reject = 0;
}

// c6. Matcher::findMatchDirect (matcher.cpp:31-141)
SVO_D int findMatchDirect(const Group& g, const PyrView& ref_pyr, int ref_frame, const PyrView& cur_pyr, int cur_frame,
                          const svo_camera& cam_ref, const svo_camera& cam_cur, const SE3d& T_cur_ref, const svo_feature& ft,
                          double ref_depth, double guess_x, double guess_y, const svo_matcher_options& opt, uint8_t* pwb,
                          MatchState& m) {
  const int pxi0 = (int)ft.px[0] / (1 << ft.level), pxi1 = (int)ft.px[1] / (1 << ft.level);
  const int boundary = 4 + 2;
  if (pxi0 < boundary || pxi1 < boundary || pxi0 >= cam_ref.width / (1 << ft.level) - boundary ||
      pxi1 >= cam_ref.height / (1 << ft.level) - boundary)
    return kFailVisibility;
  const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
  getWarpMatrixAffine(cam_ref, cam_cur, ft.px[0], ft.px[1], f_ref, ref_depth, T_cur_ref, ft.level, m.A);
  m.search_level = getBestSearchLevel(m.A, ref_pyr.n_levels - 1);
  if (!warpAffine10(g, m.A, levelView(ref_pyr, ref_frame, ft.level), ft.px[0], ft.px[1], ft.level, m.search_level, pwb))
    return kFailWarp;
  const double sc = (double)(1 << m.search_level);
  double ps_x = guess_x / sc, ps_y = guess_y / sc;
  const double start_x = ps_x, start_y = ps_y;
  const ImgView cur = levelView(cur_pyr, cur_frame, m.search_level);
  bool ok;
  if (isEdgeletType(ft.type)) {
    V2d d{m.A[0][0] * ft.grad[0] + m.A[0][1] * ft.grad[1], m.A[1][0] * ft.grad[0] + m.A[1][1] * ft.grad[1]};
    d = normalized2(d);
    ok = align1D(g, cur, d.x, d.y, pwb, opt.align_max_iter, opt.affine_est_offset != 0, opt.affine_est_gain != 0, ps_x, ps_y, &m.h_inv);
  } else {
    ok = align2D(g, cur, pwb, opt.align_max_iter, opt.affine_est_offset != 0, opt.affine_est_gain != 0, ps_x, ps_y);
  }
  if (!ok) return kFailAlignment;
  const double ddx = ps_x - start_x, ddy = ps_y - start_y;
  if (sqrt(ddx * ddx + ddy * ddy) > opt.max_patch_diff_ratio * 8) return kFailTooFar;
  m.px_x = ps_x * sc; m.px_y = ps_y * sc;
  m.f_cur = normalized3(camBackProject3(cam_cur, m.px_x, m.px_y));
  return kSuccess;
}

// ---- c5: ZMSSD<4> (patch_score.h:43-109, 264-283) -------------------------------------------------------------------
// The epipolar scans score ONE candidate position per lane: a lane holds the whole 8x8 reference patch as 16 packed words and
// forms sumB, sumBB, sumAB of its own candidate window with byte dot products (IDP4A), so the 8 lanes of a group evaluate 8
// consecutive scan steps at once and no shuffle is needed per score. All sums are exact integers, as in the reference.
struct ZmssdRef {
  unsigned ref[16];  // row y of the reference patch = bytes of ref[2y], ref[2y+1]
  int sumA, sumAA;
};
SVO_D ZmssdRef makeZmssdRef(const Group& g, uint8_t* pwb) {
  const uint8_t* it = pwb + (g.r + 1) * 10 + 1;
  const unsigned lo = it[0] | (it[1] << 8) | (it[2] << 16) | (it[3] << 24);
  const unsigned hi = it[4] | (it[5] << 8) | (it[6] << 16) | (it[7] << 24);
  // the group's rows go through its ordered-sum scratch (16-byte aligned, unused during the scan)
  unsigned* sh = reinterpret_cast<unsigned*>(pwb + kPatchBytes);
  __syncwarp(g.mask);
  sh[2 * g.r] = lo; sh[2 * g.r + 1] = hi;
  __syncwarp(g.mask);
  ZmssdRef z;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint4 v = reinterpret_cast<const uint4*>(sh)[k];
    z.ref[4 * k] = v.x; z.ref[4 * k + 1] = v.y; z.ref[4 * k + 2] = v.z; z.ref[4 * k + 3] = v.w;
  }
  __syncwarp(g.mask);  // the scratch may be rewritten
  const int sA = (int)__dp4a(lo, 0x01010101u, __dp4a(hi, 0x01010101u, 0u));
  const int sAA = (int)__dp4a(lo, lo, __dp4a(hi, hi, 0u));
  z.sumA = groupSum(g, sA);
  z.sumAA = groupSum(g, sAA);
  return z;
}
// score of the 8x8 window whose top-left pixel is (x0, y0), computed by ONE lane
SVO_D int zmssdScore(const ZmssdRef& z, const ImgView& img, int x0, int y0) {
  const uint8_t* row = img.data + (size_t)y0 * img.pitch;
  unsigned sB = 0u, sBB = 0u, sAB = 0u;
#pragma unroll
  for (int y = 0; y < 8; ++y) {
    unsigned a, b;
    loadRow8(row + (size_t)y * img.pitch, x0, a, b);
    sB = __dp4a(a, 0x01010101u, sB); sB = __dp4a(b, 0x01010101u, sB);
    sBB = __dp4a(a, a, sBB); sBB = __dp4a(b, b, sBB);
    sAB = __dp4a(a, z.ref[2 * y], sAB); sAB = __dp4a(b, z.ref[2 * y + 1], sAB);
  }
  const int B = (int)sB;
  return z.sumAA - 2 * (int)sAB + (int)sBB - (z.sumA * z.sumA - 2 * z.sumA * B + B * B) / 64;
}

SVO_D V3d angleAxisRotate(const V3d& axis, double angle, const V3d& v) {  // Eigen::AngleAxisd::toRotationMatrix() * v
  double s, c;
  sincos(angle, &s, &c);
  const V3d sin_axis = axis * s;
  const V3d cos1_axis = axis * (1.0 - c);
  M3d R;
  double tmp;
  tmp = cos1_axis.x * axis.y; R.m[0][1] = tmp - sin_axis.z; R.m[1][0] = tmp + sin_axis.z;
  tmp = cos1_axis.x * axis.z; R.m[0][2] = tmp + sin_axis.y; R.m[2][0] = tmp - sin_axis.y;
  tmp = cos1_axis.y * axis.z; R.m[1][2] = tmp - sin_axis.x; R.m[2][1] = tmp + sin_axis.x;
  R.m[0][0] = cos1_axis.x * axis.x + c;
  R.m[1][1] = cos1_axis.y * axis.y + c;
  R.m[2][2] = cos1_axis.z * axis.z + c;
  return R * v;
}

// matcher_utils::depthFromTriangulation (matcher.cpp:492-505)
SVO_D int depthFromTriangulation(const SE3d& T_search_ref, const V3d& f_ref, const V3d& f_cur, double* depth) {
  const V3d a0 = quatRotate(T_search_ref.q, f_ref);
  const V3d a1 = f_cur;
  const double m00 = dot3(a0, a0), m01 = dot3(a0, a1), m10 = dot3(a1, a0), m11 = dot3(a1, a1);
  const double det = m00 * m11 - m10 * m01;
  if (det < 0.000001) return kFailTriangulation;
  const double invdet = 1.0 / det;
  const V3d row0 = a0 * (-(m11 * invdet)) + a1 * (-(-m01 * invdet));
  *depth = fabs(dot3(row0, T_search_ref.t));
  return kSuccess;
}

// ---- c7. Matcher::findEpipolarMatchDirect (matcher.cpp:157-241) in three pieces -----------------------------------------
//   epiSetup   scalar geometry of one feature: epipolar segment, affine warp matrix, edgelet angle gate, search level and
//              the scan parameters (one thread suffices; the phased seed / match pipelines run it one thread per feature);
//   epiMatch   the 8-lane group part: affine warp of the reference patch, ZMSSD scan along the segment, sub-pixel alignment;
//   epiFinish  scalar: bearing vector of the match and the triangulated depth.
// findEpipolarMatchDirect() below chains the three inside one group for the callers that want everything in one kernel.
struct EpiSetup {
  double A[2][2];                          // A_cur_ref_
  double epi_x, epi_y, epi_length_pyramid; // epi_image_, epi_length_pyramid_
  double dir_x, dir_y;                     // epi_image_.normalized()
  double px0_x, px0_y;                     // short segment: mid-point of its end points, the local match starts there
  int search_level, reject;
  int early;                               // < 0: go on with epiMatch, else the MatchResult findEpipolarMatchDirect returns at once
  int short_epi;
  int n_steps, half_steps;                 // scan length (after the max_epi_search_steps cap); unit sphere: n_steps / 2
  double step;                             // unit sphere: angle per step
  V3d axis, f_C;                           // unit sphere: rotation axis and the bearing of the depth estimate
  double step_x, step_y, uvC_x, uvC_y;     // unit plane: step and start on the plane z = 1
};

// The scan parameters Matcher::scanEpipolarUnitSphere / scanEpipolarUnitPlane derive from the segment A~C~B and the member
// epi_length_pyramid_ (matcher.cpp:340-362, 415-441); e.epi_length_pyramid must be set.
SVO_D void epiScanSetup(const V3d& A, const V3d& B, const V3d& C, const svo_matcher_options& opt, EpiSetup& e) {
  size_t n_steps = (size_t)(e.epi_length_pyramid / 0.7);
  if (opt.scan_on_unit_sphere) {  // matcher.cpp:415-441
    n_steps = n_steps > (size_t)opt.max_epi_search_steps ? (size_t)opt.max_epi_search_steps : n_steps;
    const V3d f_A = normalized3(A), f_B = normalized3(B);
    e.step = acos(dot3(f_A, f_B)) / n_steps;
    e.axis = normalized3(cross3(f_B, f_A));
    e.f_C = normalized3(C);
    e.n_steps = (int)n_steps;
    e.half_steps = (int)(n_steps / 2);
  } else {  // matcher.cpp:340-362
    e.step_x = (A.x / A.z - B.x / B.z) / n_steps; e.step_y = (A.y / A.z - B.y / B.z) / n_steps;
    if (n_steps > (size_t)opt.max_epi_search_steps) n_steps = (size_t)opt.max_epi_search_steps;
    e.uvC_x = C.x / C.z; e.uvC_y = C.y / C.z;
    e.n_steps = (int)n_steps;
  }
}

SVO_D void epiSetup(const svo_camera& cam_ref, const svo_camera& cam_cur, const SE3d& T_cur_ref, const svo_feature& ft, double d_estimate_inv,
                    double d_min_inv, double d_max_inv, const svo_matcher_options& opt, int max_level, EpiSetup& e) {
  const V3d f_ref{ft.f[0], ft.f[1], ft.f[2]};
  const V3d Rf = quatRotate(T_cur_ref.q, f_ref);
  const V3d A = Rf + T_cur_ref.t * d_min_inv;
  const V3d B = Rf + T_cur_ref.t * d_max_inv;
  const V2d px_A = camProject3(cam_cur, A), px_B = camProject3(cam_cur, B);
  e.epi_x = px_A.x - px_B.x; e.epi_y = px_A.y - px_B.y;
  getWarpMatrixAffine(cam_ref, cam_cur, ft.px[0], ft.px[1], f_ref, 1.0 / fmax(0.000001, d_estimate_inv), T_cur_ref, ft.level, e.A);
  e.reject = 0; e.early = -1; e.short_epi = 0;
  e.search_level = 0; e.epi_length_pyramid = 0.0; e.dir_x = e.dir_y = e.px0_x = e.px0_y = 0.0;
  e.n_steps = e.half_steps = 0; e.step = 0.0; e.axis = V3d{0, 0, 0}; e.f_C = V3d{0, 0, 0};
  e.step_x = e.step_y = e.uvC_x = e.uvC_y = 0.0;
  if (isEdgeletType(ft.type) && opt.epi_search_edgelet_filtering) {
    const V2d gc = normalized2(V2d{e.A[0][0] * ft.grad[0] + e.A[0][1] * ft.grad[1], e.A[1][0] * ft.grad[0] + e.A[1][1] * ft.grad[1]});
    const V2d en = normalized2(V2d{e.epi_x, e.epi_y});
    const double cosangle = fabs(gc.x * en.x + gc.y * en.y);
    if (cosangle < opt.epi_search_edgelet_max_angle) { e.reject = 1; e.early = kFailAngle; return; }
  }
  e.search_level = getBestSearchLevel(e.A, max_level);
  e.epi_length_pyramid = sqrt(e.epi_x * e.epi_x + e.epi_y * e.epi_y) / (1 << e.search_level);
  const V2d epi_dir = normalized2(V2d{e.epi_x, e.epi_y});
  e.dir_x = epi_dir.x; e.dir_y = epi_dir.y;
  // matcher.cpp:209-218: a short epipolar segment goes straight to the local (sub-pixel) match at its mid-point
  if (e.epi_length_pyramid < 2.0) {
    e.short_epi = 1;
    e.px0_x = (px_A.x + px_B.x) / 2.0; e.px0_y = (px_A.y + px_B.y) / 2.0;
    return;
  }
  const V3d C = Rf + T_cur_ref.t * d_estimate_inv;
  epiScanSetup(A, B, C, opt, e);
}

// Group arg-min of the lanes' scores; ties go to the lowest lane = the earliest scan step (the reference keeps the first
// minimum: strict <).
SVO_D void groupArgMin(const Group& g, int& z, int& lane) {
#pragma unroll
  for (int o = 1; o < kGroup; o <<= 1) {
    const int oz = __shfl_xor_sync(g.mask, z, o), ol = __shfl_xor_sync(g.mask, lane, o);
    if (oz < z || (oz == z && ol < lane)) { z = oz; lane = ol; }
  }
}
SVO_D double groupBroadcast(const Group& g, double v, int src) { return __shfl_sync(g.mask, v, src, kGroup); }
// index of the first lane of the group whose predicate holds, kGroup if none
SVO_D int groupFirst(const Group& g, bool p) {
  const unsigned b = (__ballot_sync(g.mask, p) & g.mask) >> ((threadIdx.x & 31) & ~(kGroup - 1));
  return b ? __ffs((int)b) - 1 : kGroup;
}

struct WithinBox { int x_hi, y_hi; };  // matcher.cpp:314-322 with the divisions hoisted: 8 <= p < hi
SVO_D WithinBox makeWithinBox(const svo_camera& cam, int patch_level) {
  return WithinBox{cam.width / (1 << patch_level) - 8, cam.height / (1 << patch_level) - 8};
}
SVO_D bool withinBox(const WithinBox& b, int px, int py) { return !(px < 8 || py < 8 || px >= b.x_hi || py >= b.y_hi); }

// Matcher::scanEpipolarUnitSphere (matcher.cpp:415-488). Lane r of the group evaluates loop iteration i_base + r: the angle,
// the rotated bearing and its pixel depend on i only. The loop's sequential rules are applied per chunk of 8 iterations: an
// iteration is skipped when its pixel equals the previous iteration's pixel (`last` always holds the previous iteration's
// pixel); the first out-of-image pixel of a chunk ends the chunk (first half: jump to half_steps + 1, second half: stop) and
// the lanes behind it are discarded. Returns the best ZMSSD score and the pixel of the best bearing.
SVO_D int scanEpipolarUnitSphere(const Group& g, const EpiSetup& e, const svo_camera& cam_cur, const ImgView& cur, int pl,
                                 const ZmssdRef& zref, double& px_x, double& px_y, const int zmssd_init = 2000 * 64) {
  const double inv_pl = 1.0 / (double)(1 << pl);  // exact: x / 2^pl == x * 2^-pl
  const WithinBox box = makeWithinBox(cam_cur, pl);
  const double neg_step = -e.step;
  int zmssd_best = zmssd_init;  // PatchScore::threshold() in findEpipolarMatchDirect (matcher.cpp:167)
  V3d f_best = e.f_C;
  int last_x = 0, last_y = 0;
  int i_base = 0;
  while (i_base < e.n_steps) {
    const int i = i_base + g.r;
    const bool valid = i < e.n_steps;
    const double angle = i < e.half_steps ? (double)i * e.step : (double)(i - e.half_steps) * neg_step;
    const V3d f = angleAxisRotate(e.axis, angle, e.f_C);
    const V2d px = camProject3(cam_cur, f);
    const int pxi0 = (int)(px.x * inv_pl + 0.5), pxi1 = (int)(px.y * inv_pl + 0.5);
    int prev_x = __shfl_up_sync(g.mask, pxi0, 1, kGroup), prev_y = __shfl_up_sync(g.mask, pxi1, 1, kGroup);
    if (g.r == 0) { prev_x = last_x; prev_y = last_y; }
    const bool differs = !(pxi0 == prev_x && pxi1 == prev_y);
    const bool within = withinBox(box, pxi0, pxi1);
    const int r_out = groupFirst(g, valid && differs && !within);
    int z = 0x7fffffff, zl = g.r;
    if (valid && differs && within && g.r < r_out) z = zmssdScore(zref, cur, pxi0 - 4, pxi1 - 4);
    groupArgMin(g, z, zl);
    if (z < zmssd_best) {
      zmssd_best = z;
      f_best = V3d{groupBroadcast(g, f.x, zl), groupBroadcast(g, f.y, zl), groupBroadcast(g, f.z, zl)};
    }
    const int src = r_out < kGroup ? r_out : kGroup - 1;
    last_x = __shfl_sync(g.mask, pxi0, src, kGroup); last_y = __shfl_sync(g.mask, pxi1, src, kGroup);
    if (r_out < kGroup) {
      if (i_base + r_out < e.half_steps) i_base = e.half_steps + 1;  // `i = half_steps; continue;` -> ++i
      else break;
    } else {
      i_base += kGroup;
    }
  }
  const V2d pb = camProject3(cam_cur, f_best);
  px_x = pb.x; px_y = pb.y;
  return zmssd_best;
}

// Matcher::scanEpipolarUnitPlane (matcher.cpp:340-413). The position on the plane z = 1 is a running sum uv += step, so a
// chunk forms the 8 prefix sums by the same sequence of additions (every lane forms all eight and keeps its own); besides
// the rules above, the first scored iteration with i > n_steps / 2 of the forward pass reverses the scan after its score.
SVO_D int scanEpipolarUnitPlane(const Group& g, const EpiSetup& e, const svo_camera& cam_cur, const ImgView& cur, int pl,
                                const ZmssdRef& zref, double& px_x, double& px_y, const int zmssd_init = 2000 * 64) {
  const double inv_pl = 1.0 / (double)(1 << pl);
  const WithinBox box = makeWithinBox(cam_cur, pl);
  const double half_n = e.n_steps * 0.5;
  int zmssd_best = zmssd_init;  // PatchScore::threshold() in findEpipolarMatchDirect (matcher.cpp:167)
  double step_x = e.step_x, step_y = e.step_y;
  double base_x = e.uvC_x, base_y = e.uvC_y, best_x = e.uvC_x, best_y = e.uvC_y;
  bool forward = true;
  int last_x = 0, last_y = 0;
  int i_base = 0;
  while (i_base < e.n_steps) {
    const int i = i_base + g.r;
    const bool valid = i < e.n_steps;
    double uv_x = base_x, uv_y = base_y, run_x = base_x, run_y = base_y;
#pragma unroll
    for (int t = 1; t <= kGroup; ++t) {
      run_x += step_x; run_y += step_y;
      if (t == g.r) { uv_x = run_x; uv_y = run_y; }
    }
    double ud, vd;
    camDistort(cam_cur, uv_x, uv_y, ud, vd);  // project3 of (uv, 1): 1 / 1.0 and the products with it are exact
    const double pxx = cam_cur.fx * ud + cam_cur.cx, pxy = cam_cur.fy * vd + cam_cur.cy;
    const int pxi0 = (int)(pxx * inv_pl + 0.5), pxi1 = (int)(pxy * inv_pl + 0.5);
    int prev_x = __shfl_up_sync(g.mask, pxi0, 1, kGroup), prev_y = __shfl_up_sync(g.mask, pxi1, 1, kGroup);
    if (g.r == 0) { prev_x = last_x; prev_y = last_y; }
    const bool differs = !(pxi0 == prev_x && pxi1 == prev_y);
    const bool within = withinBox(box, pxi0, pxi1);
    const int r_out = groupFirst(g, valid && differs && !within);
    const int r_flip = groupFirst(g, forward && valid && differs && within && (double)i > half_n);
    int z = 0x7fffffff, zl = g.r;
    if (valid && differs && within && g.r < r_out && g.r <= r_flip) z = zmssdScore(zref, cur, pxi0 - 4, pxi1 - 4);
    groupArgMin(g, z, zl);
    if (z < zmssd_best) {
      zmssd_best = z;
      best_x = groupBroadcast(g, uv_x, zl); best_y = groupBroadcast(g, uv_y, zl);
    }
    const int r_end = r_out < r_flip ? r_out : r_flip;
    const int src = r_end < kGroup ? r_end : kGroup - 1;
    last_x = __shfl_sync(g.mask, pxi0, src, kGroup); last_y = __shfl_sync(g.mask, pxi1, src, kGroup);
    if (r_end < kGroup) {
      if (r_out < r_flip) {  // left the image
        if (!forward) break;
        i_base = (int)(size_t)(e.n_steps * 0.5) + 1;  // `i = n_steps * 0.5; ...; continue;` -> ++i
      } else {               // scored, then reversed
        i_base = i_base + r_flip + 1;
      }
      step_x = -step_x; step_y = -step_y;
      base_x = e.uvC_x + step_x; base_y = e.uvC_y + step_y;  // uv = uvC, then the loop's uv += step
      forward = false;
    } else {
      base_x = run_x; base_y = run_y;  // uv of iteration i_base + 8
      i_base += kGroup;
    }
  }
  double ud, vd;
  camDistort(cam_cur, best_x, best_y, ud, vd);
  px_x = cam_cur.fx * ud + cam_cur.cx; px_y = cam_cur.fy * vd + cam_cur.cy;
  return zmssd_best;
}

// findLocalMatch (matcher.cpp:262-289): px in/out (level-0 pixels)
SVO_D int findLocalMatch(const Group& g, const PyrView& cur_pyr, int cur_frame, double dir_x, double dir_y, int patch_level,
                         const svo_matcher_options& opt, bool align_1d, uint8_t* pwb, double& px_x, double& px_y, double* h_inv) {
  const double sc = (double)(1 << patch_level);
  double ps_x = px_x / sc, ps_y = px_y / sc;
  const ImgView cur = levelView(cur_pyr, cur_frame, patch_level);
  bool res;
  if (align_1d) res = align1D(g, cur, dir_x, dir_y, pwb, opt.align_max_iter, opt.affine_est_offset != 0, opt.affine_est_gain != 0, ps_x, ps_y, h_inv);
  else res = align2D(g, cur, pwb, opt.align_max_iter, opt.affine_est_offset != 0, opt.affine_est_gain != 0, ps_x, ps_y);
  if (!res) return kFailAlignment;
  px_x = ps_x * sc; px_y = ps_y * sc;
  return kSuccess;
}

// The group part of findEpipolarMatchDirect (matcher.cpp:196-229): warp, scan, sub-pixel refinement. px_x / px_y receive px_cur_.
// SCAN: 1 = unit sphere, 0 = unit plane, 2 = opt.scan_on_unit_sphere at run time (kernels whose launch knows the mode compile one scan).
template <int SCAN = 2>
SVO_D int epiMatch(const Group& g, const PyrView& ref_pyr, int ref_frame, const PyrView& cur_pyr, int cur_frame, const svo_camera& cam_cur,
                   const svo_feature& ft, const EpiSetup& e, const svo_matcher_options& opt, bool align_1d, uint8_t* pwb, double& px_x,
                   double& px_y, double* h_inv) {
  if (!warpAffine10(g, e.A, levelView(ref_pyr, ref_frame, ft.level), ft.px[0], ft.px[1], ft.level, e.search_level, pwb))
    return kFailWarp;
  if (e.short_epi) {
    px_x = e.px0_x; px_y = e.px0_y;
  } else {
    const ZmssdRef zref = makeZmssdRef(g, pwb);
    const ImgView cur = levelView(cur_pyr, cur_frame, e.search_level);
    const bool sphere = SCAN == 2 ? opt.scan_on_unit_sphere != 0 : SCAN == 1;
    const int zmssd_best = sphere ? scanEpipolarUnitSphere(g, e, cam_cur, cur, e.search_level, zref, px_x, px_y)
                                  : scanEpipolarUnitPlane(g, e, cam_cur, cur, e.search_level, zref, px_x, px_y);
    if (!(zmssd_best < 2000 * 64)) return kFailScore;
  }
  // both ways end in ONE findLocalMatch call site (code size: these kernels stall on instruction fetch otherwise, profiles/)
  if (e.short_epi || opt.subpix_refinement) return findLocalMatch(g, cur_pyr, cur_frame, e.dir_x, e.dir_y, e.search_level, opt, align_1d, pwb, px_x, px_y, h_inv);
  return kSuccess;
}

// matcher.cpp:231-240: bearing of the match and its depth by triangulation
SVO_D int epiFinish(const svo_camera& cam_cur, const SE3d& T_cur_ref, const V3d& f_ref, double px_x, double px_y, V3d& f_cur, double* depth) {
  f_cur = normalized3(camBackProject3(cam_cur, px_x, px_y));
  return depthFromTriangulation(T_cur_ref, f_ref, f_cur, depth);
}

template <int SCAN = 2>
SVO_D int findEpipolarMatchDirect(const Group& g, const PyrView& ref_pyr, int ref_frame, const PyrView& cur_pyr, int cur_frame,
                                  const svo_camera& cam_ref, const svo_camera& cam_cur, const SE3d& T_cur_ref, const svo_feature& ft,
                                  double d_estimate_inv, double d_min_inv, double d_max_inv, const svo_matcher_options& opt,
                                  bool align_1d, uint8_t* pwb, MatchState& m, double* depth) {
  EpiSetup e;
  epiSetup(cam_ref, cam_cur, T_cur_ref, ft, d_estimate_inv, d_min_inv, d_max_inv, opt, ref_pyr.n_levels - 1, e);
  m.A[0][0] = e.A[0][0]; m.A[0][1] = e.A[0][1]; m.A[1][0] = e.A[1][0]; m.A[1][1] = e.A[1][1];
  m.epi_x = e.epi_x; m.epi_y = e.epi_y;
  m.reject = e.reject;
  if (e.early >= 0) return e.early;
  m.search_level = e.search_level;
  m.epi_length_pyramid = e.epi_length_pyramid;
  const int res = epiMatch<SCAN>(g, ref_pyr, ref_frame, cur_pyr, cur_frame, cam_cur, ft, e, opt, align_1d, pwb, m.px_x, m.px_y, &m.h_inv);
  if (res != kSuccess) return res;
  return epiFinish(cam_cur, T_cur_ref, V3d{ft.f[0], ft.f[1], ft.f[2]}, m.px_x, m.px_y, m.f_cur, depth);
}

}  // namespace svo_dev
