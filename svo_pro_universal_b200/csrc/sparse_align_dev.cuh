// Device helpers shared by the sparse-alignment kernels (sparse_align.cu: one CTA per pair; sparse_align_pp.cu: two pairs per CTA,
// ping-pong with a solver warp): warp reductions, the Newton reciprocal, tap loaders, the small-angle SE3 update, the LDL^T factor /
// solve, the patch sums behind H and the per-patch Jacobian rows. See sparse_align.cu for the reference citations.
#pragma once
#include "common.cuh"
#include <cstdlib>

namespace svo_align {

constexpr int kFixedSlots = 180;                  // compile-time stride for the common <= 180-feature case (max_fts)
constexpr int kCamBlk = 36;                       // per camera: R_cam_imu (9) | R_imu_cam (9) | t_cam_imu (3) | t_imu_cam (3) | T_cur_ref (12)

struct AlignParams {
  int n_cams, B, max_features, slots;
  PyrView ref_pyr[SVO_MAX_CAMS], cur_pyr[SVO_MAX_CAMS];
  const int* ref_frame_idx;
  const int* cur_frame_idx;
  svo_camera cams[SVO_MAX_CAMS];
  double T_cam_imu[SVO_MAX_CAMS][7];
  const double* T_imu_world_ref;
  const double* T_imu_world_cur;
  const int* n_features;
  const double* px;
  const double* f;
  const double* depth;
  const uint8_t* eligible;
  svo_sparse_align_options opt;
  const svo_align_prior* priors;
  svo_align_result* results;
};

// Storage type of the per-level reference patch cache (32 values per feature, the bulk of a pair's shared memory). FP32 halves it (and
// the shared-memory traffic of every iteration: 0.875 -> 0.821 ms per 4096 pairs on the B200); the values are rounded ONCE per level (relative 2^-24, at most 7.6e-6 grey
// levels) and every operation on them stays FP64. The rounding is a fixed perturbation of the reference data (not noise per iteration):
// the converged pose moves by ~1e-10 rad / m (measured, tests/), six orders inside the 1e-4 tolerance, and the iteration counts of
// every parity case are unchanged. -DSVO_ALIGN_PATCH_F64 restores the FP64 cache (A/B builds).
#ifdef SVO_ALIGN_PATCH_F64
typedef double PatchT;
SVO_D double patchLoad(const double* p) { return *p; }
#else
typedef float PatchT;
// exact float -> double of a non-negative normal float (or 0, which comes out as 2^-127) with integer instructions only: the FP64
// pipe is the busy one, and F2F.F64.F32 shares the quarter-rate conversion pipe
SVO_D double patchLoad(const float* p) {
  const unsigned u = __float_as_uint(*p);
  return __hiloint2double((int)((u >> 3) + 0x38000000u), (int)(u << 29));
}
#endif

SVO_D double warpSum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NP (power of two <= 32) per-lane values over the warp with a transposing butterfly: every step halves the number of
// values a lane still carries, so NP values cost NP - 1 + log2(32 / NP) shuffled doubles instead of 5 * NP.
// Returns, in every lane, the warp total of value number `lane / (32 / NP)` (NP == 32: value `lane`).
template <int NP>
SVO_D double warpSumMulti(double* v, int lane) {
  int off = 16;
#pragma unroll
  for (int n = NP; n > 1; n >>= 1, off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const double keep = up ? v[i + n / 2] : v[i];
      const double send = up ? v[i] : v[i + n / 2];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// 1/d to within an ulp: hardware seed (relative error <= 2^-20) + one cubic Newton step (error e^3 < 2^-60): 3 dependent FMAs
// behind the MUFU; the IEEE-correct division the compiler emits is about 80 cycles deep and sits on the serial path of every iteration.
SVO_D double fastRcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  const double e = fma(-d, x, 1.0);
  return fma(x, fma(e, e, e), x);
}

// Five taps x0..x0+4 of a row always lie inside two aligned 32-bit words: taps 0-3 in `a`, tap 4 in byte 0 of `b`.
SVO_D void loadRow5(const uint8_t* row, int x0, unsigned& a, unsigned& b) {
  const unsigned* w = reinterpret_cast<const unsigned*>(row + (x0 & ~3));
  const unsigned w0 = __ldg(w), w1 = __ldg(w + 1);
  const unsigned sh = (x0 & 3) * 8;
  a = __funnelshift_r(w0, w1, sh);
  b = w1 >> sh;
}

// exact u8 -> double on the FP64 pipe (2^52 + b has b in its low mantissa bits), instead of I2F on the quarter-rate XU pipe
SVO_D double u8ToDouble(unsigned b) { return __hiloint2double(0x43300000, (int)b) - 4503599627370496.0; }
SVO_D unsigned byteAt(unsigned w, int i) { return __byte_perm(w, 0u, 0x4440 | i); }
// byte i of w as a double. SVO_ALIGN_I2F_MASK (A/B builds) selects, per tap column, the conversion instruction (I2F.F64.U8 with a byte
// selector, conversion pipe) instead of the byte extraction + FP64 add above.
#ifndef SVO_ALIGN_I2F_MASK
#define SVO_ALIGN_I2F_MASK 31   // measured: 0.818 (none) -> 0.810 ms (all five columns) per 4096 pairs: the FP64 pipe is the shared one
#endif
template <int COL>
SVO_D double tapToDouble(unsigned w, int i) {
  if ((SVO_ALIGN_I2F_MASK >> COL) & 1) return (double)((w >> (8 * i)) & 0xffu);
  return u8ToDouble(byteAt(w, i));
}

// cos(x) and sin(x)/x for y = x^2 <= 0.25 (Taylor to y^9: truncation < 1e-24), evaluated pairwise to keep the dependent chain short.
SVO_D void cosSinc(double y, double& c, double& sc) {
  const double y2 = y * y, y4 = y2 * y2, y8 = y4 * y4;
  const double c01 = fma(y, -1.0 / 2.0, 1.0), c23 = fma(y, -1.0 / 720.0, 1.0 / 24.0);
  const double c45 = fma(y, -1.0 / 3628800.0, 1.0 / 40320.0), c67 = fma(y, -1.0 / 87178291200.0, 1.0 / 479001600.0);
  const double c89 = fma(y, -1.0 / 6402373705728000.0, 1.0 / 20922789888000.0);
  c = fma(y8, c89, fma(y4, fma(y2, c67, c45), fma(y2, c23, c01)));
  const double s01 = fma(y, -1.0 / 6.0, 1.0), s23 = fma(y, -1.0 / 5040.0, 1.0 / 120.0);
  const double s45 = fma(y, -1.0 / 39916800.0, 1.0 / 362880.0), s67 = fma(y, -1.0 / 1307674368000.0, 1.0 / 6227020800.0);
  const double s89 = fma(y, -1.0 / 121645100408832000.0, 1.0 / 355687428096000.0);
  sc = fma(y8, s89, fma(y4, fma(y2, s67, s45), fma(y2, s23, s01)));
}
// quatExp (common.cuh) with the polynomial above for |dx| <= 1 rad; same small-angle branch as the reference.
SVO_D Quatd quatExpFast(const V3d& dx) {
  const double th2 = dot3(dx, dx);
  if (th2 > 1.0) return quatExp(dx);
  double ct, sc;
  cosSinc(0.25 * th2, ct, sc);
  const double na = th2 < SVO_EPS4ROOT * SVO_EPS4ROOT ? 0.5 + th2 * (1.0 / 48.0) : 0.5 * sc;
  return {ct, dx.x * na, dx.y * na, dx.z * na};
}
// q / |q| for a quaternion that is already unit up to rounding: 1/sqrt(1+e) = 1 - e/2 + 3e^2/8 (|e| < 1e-6 -> error < 1e-18)
SVO_D void quatNormalizeFast(Quatd& q) {
  const double e = quatSqNorm(q) - 1.0;
  if (fabs(e) < 1e-6) {
    const double sN = fma(e, fma(e, 0.375, -0.5), 1.0);
    q.w *= sN; q.x *= sN; q.y *= sN; q.z *= sN;
  } else {
    quatNormalize(q);
  }
}

// ref: src/vikit/vikit_solver/src/robust_cost.cpp:48-60 with b = 4.6851f (robust_cost.h:70)
SVO_D float tukeyWeight(float error) {
  const float b_square = 4.6851f * 4.6851f;
  const float x_square = error * error;
  if (x_square <= b_square) {
    const float tmp = 1.0f - x_square / b_square;
    return tmp * tmp;
  }
  return 0.0f;
}

// Symmetric solve H dx = g by LDL^T without pivoting, split into factor (once per H) and solve (every iteration).
// A zero pivot (an all-zero row/column of the PSD normal matrix: illumination parameters switched off) yields dx_k = 0,
// which is what Eigen's pivoted LDLT::solve returns for those rows (mini_least_squares_solver.hpp:258). `tri` holds the
// upper triangle row-major ((0,0),(0,1)..(0,D-1),(1,1)..), `diag_add` is added on the diagonal. Lf is the strict lower
// triangle row-major (Lf[i*(i-1)/2 + j], j < i), rd the pivot reciprocals (0 for a zero pivot). Everything unrolls at
// compile time; one Newton reciprocal per pivot.
template <int D>
SVO_D void ldltFactor(const double* tri, const double* diag_add, double* Lf, double* rdf) {
  double L[D][D], dd[D], rd[D];
#pragma unroll
  for (int k = 0; k < D; ++k) {
    double w[D];
    double d = tri[k * D - (k * (k - 1)) / 2] + diag_add[k];
#pragma unroll
    for (int j = 0; j < k; ++j) { w[j] = L[k][j] * dd[j]; d -= L[k][j] * w[j]; }
    dd[k] = d;
    const bool ok = fabs(d) > 2.2250738585072014e-308;
    rd[k] = ok ? fastRcp(d) : 0.0;
#pragma unroll
    for (int i = k + 1; i < D; ++i) {
      double s = tri[k * D - (k * (k - 1)) / 2 + (i - k)];
#pragma unroll
      for (int j = 0; j < k; ++j) s -= L[i][j] * w[j];
      L[i][k] = ok ? s * rd[k] : s;
    }
  }
#pragma unroll
  for (int i = 0; i < D; ++i) {
    rdf[i] = rd[i];
#pragma unroll
    for (int j = 0; j < i; ++j) Lf[(i * (i - 1)) / 2 + j] = L[i][j];
  }
}
template <int D>
SVO_D void ldltSolve(const double* Lf, const double* rdf, const double* g, double* dx) {
  double L[D][D], x[D];
#pragma unroll
  for (int i = 0; i < D; ++i)
#pragma unroll
    for (int j = 0; j < i; ++j) L[i][j] = Lf[(i * (i - 1)) / 2 + j];
#pragma unroll
  for (int i = 0; i < D; ++i) {
    double s = g[i];
#pragma unroll
    for (int j = 0; j < i; ++j) s -= L[i][j] * x[j];
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) x[i] *= rdf[i];  // zero pivot -> 0, as Eigen's LDLT::solve
#pragma unroll
  for (int i = D - 1; i >= 0; --i) {
    double s = x[i];
#pragma unroll
    for (int j = D - 1; j > i; --j) s -= L[j][i] * x[j];  // x[i+1], the newest unknown, enters last: one dependent FMA per row
    x[i] = s;
  }
#pragma unroll
  for (int i = 0; i < D; ++i) dx[i] = x[i];
}

// T_cur_ref of every camera from the IMU-frame state: R = R_ci R(q) R_ic, t = R_ci (R(q) t_ic + t) + t_ci, one output
// element per lane (camera blocks hold the constant matrices, see kCamBlk).
SVO_D void refreshCameraTransforms(const SE3d& T, double* camblk, int n_cams, int lane) {
  const M3d R = quatToMatrix(T.q);
  const double tv[3] = {T.t.x, T.t.y, T.t.z};
  for (int idx = lane; idx < 12 * n_cams; idx += 32) {
    const int c = idx / 12, e = idx - 12 * c;
    double* cb = camblk + kCamBlk * c;
    const double* Rci = cb;
    const double* Ric = cb + 9;
    double v;
    if (e < 9) {
      const int r = e / 3, col = e - 3 * r;
      v = 0.0;
#pragma unroll
      for (int k = 0; k < 3; ++k) v += Rci[3 * r + k] * (R.m[k][0] * Ric[col] + R.m[k][1] * Ric[3 + col] + R.m[k][2] * Ric[6 + col]);
    } else {
      const int r = e - 9;
      v = cb[18 + r];
#pragma unroll
      for (int k = 0; k < 3; ++k) v += Rci[3 * r + k] * (R.m[k][0] * cb[21] + R.m[k][1] * cb[22] + R.m[k][2] * cb[23] + tv[k]);
    }
    cb[24 + e] = v;
  }
}

// index of patch element (X, Y) of the 6x6 interpolated reference patch among the 32 stored values
// (rows 0 and 5 keep only X = 1..4: the corners are never read)
SVO_HD constexpr int patchIdx(int X, int Y) { return Y == 0 ? X - 1 : (Y == 5 ? 28 + X - 1 : 4 + (Y - 1) * 6 + X); }

SVO_D int patchIdxRt(int X, int Y) { return Y == 0 ? X - 1 : (Y == 5 ? 27 + X : 4 + (Y - 1) * 6 + X); }  // patchIdx for run-time rows

struct PatchSums {  // weighted sums over the 16 pixels of one patch
  double sxx, sxy, syy;                         // H pose block
  double sx6, sy6, sx7, sy7, s66, s67, s77;     // H illumination blocks
};

// Sums that form H when every weight is 1: they depend on the reference patch only.
template <bool ILLUM>
SVO_D PatchSums unitWeightSums(const PatchT* patch, int stride, bool est_gain, bool est_off) {
  PatchSums p = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int y = 0; y < 4; ++y)
#pragma unroll
    for (int x = 0; x < 4; ++x) {
      const double ref = patchLoad(patch + patchIdx(x + 1, y + 1) * stride);
      const double dx = 0.5 * (patchLoad(patch + patchIdx(x + 2, y + 1) * stride) - patchLoad(patch + patchIdx(x, y + 1) * stride));
      const double dy = 0.5 * (patchLoad(patch + patchIdx(x + 1, y + 2) * stride) - patchLoad(patch + patchIdx(x + 1, y) * stride));
      p.sxx += dx * dx; p.sxy += dx * dy; p.syy += dy * dy;
      if (ILLUM) {
        const double a6 = est_gain ? -ref : 0.0, a7 = est_off ? -1.0 : 0.0;
        p.sx6 += dx * a6; p.sy6 += dy * a6; p.sx7 += dx * a7; p.sy7 += dy * a7;
        p.s66 += a6 * a6; p.s67 += a6 * a7; p.s77 += a7 * a7;
      }
    }
  return p;
}

// Rank-2 expansion of one patch's sums into the D(D+1)/2 upper-triangle entries of H, warp-reduced into red[0..NH).
template <int D>
SVO_D void reduceH(const PatchSums& p, const double* jp0, const double* jp1, double* red, int lane) {
  constexpr int NH = D * (D + 1) / 2;
  double ua[6], va[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    ua[k] = p.sxx * jp0[k] + p.sxy * jp1[k];
    va[k] = p.sxy * jp0[k] + p.syy * jp1[k];
  }
  double h[NH > 32 ? 40 : 32];
  int idx = 0;
#pragma unroll
  for (int a = 0; a < D; ++a) {
#pragma unroll
    for (int b = a; b < D; ++b) {
      double v;
      if (b < 6) v = ua[a] * jp0[b] + va[a] * jp1[b];
      else if (a < 6) v = (b == 6) ? (jp0[a] * p.sx6 + jp1[a] * p.sy6) : (jp0[a] * p.sx7 + jp1[a] * p.sy7);
      else v = (a == 6 && b == 6) ? p.s66 : (a == 6 ? p.s67 : p.s77);
      h[idx++] = v;
    }
  }
#pragma unroll
  for (int k = NH; k < (NH > 32 ? 40 : 32); ++k) h[k] = 0.0;
  const double t = warpSumMulti<32>(h, lane);       // lane k holds entry k
  if (lane < (NH < 32 ? NH : 32)) red[lane] += t;
  if (NH > 32) {                                     // 8-DoF: entries 32..35
    const double t2 = warpSumMulti<8>(h + 32, lane);  // lanes 4k hold entry 32 + k
    if ((lane & 3) == 0 && 32 + (lane >> 2) < NH) red[32 + (lane >> 2)] += t2;
  }
}

// 2x6 Jacobian rows of one patch, rebuilt from the caches (only needed when H is reduced):
// Jp = (mult * Jproj) * R_cam_imu * [I | -skew(p_imu)] * scale (sparse_img_align.cpp:297-316, :373-376).
template <bool DJ>
SVO_D void patchJacobian(const double* xyz, const double* aux, int stride, const double* cb, double mult, double scale,
                         double* jp0, double* jp1) {
  const double X = xyz[0], Y = xyz[stride], Z = xyz[2 * stride];
  double J[2][3];
  if (DJ) {
#pragma unroll
    for (int k = 0; k < 3; ++k) { J[0][k] = aux[k * stride]; J[1][k] = aux[(3 + k) * stride]; }
  } else {
    const double iz = aux[0], sI = -mult * iz;  // Frame::jacobian_xyz2uv_imu (frame.h:342-357) times the focal length
    J[0][0] = sI; J[0][1] = 0.0; J[0][2] = -sI * X * iz;
    J[1][0] = 0.0; J[1][1] = sI; J[1][2] = -sI * Y * iz;
  }
  const double* Rci = cb;
  const double* Ric = cb + 9;
  const double px = Ric[0] * X + Ric[1] * Y + Ric[2] * Z + cb[21];
  const double py = Ric[3] * X + Ric[4] * Y + Ric[5] * Z + cb[22];
  const double pz = Ric[6] * X + Ric[7] * Y + Ric[8] * Z + cb[23];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    double* jp = r == 0 ? jp0 : jp1;
    const double b0 = (J[r][0] * Rci[0] + J[r][1] * Rci[3] + J[r][2] * Rci[6]) * scale;
    const double b1 = (J[r][0] * Rci[1] + J[r][1] * Rci[4] + J[r][2] * Rci[7]) * scale;
    const double b2 = (J[r][0] * Rci[2] + J[r][1] * Rci[5] + J[r][2] * Rci[8]) * scale;
    jp[0] = b0; jp[1] = b1; jp[2] = b2;
    jp[3] = b2 * py - b1 * pz;
    jp[4] = b0 * pz - b2 * px;
    jp[5] = b1 * px - b0 * py;
  }
}

}  // namespace svo_align
