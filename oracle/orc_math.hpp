// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's small-matrix / SE3 / camera arithmetic for
// the direct front-end hot path. Nothing in the product path (svo_pro_universal_b200/,
// include/) may include, link or call anything in oracle/. Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
//
// Parity status: the reference's own tests hold no golden vectors for this path
// (SURVEY.md §4, §8c), and Eigen/OpenCV/glog are not installed, so these functions
// are line-faithful restatements ("parity unpinned" except for rows a2-a4, which are
// pinned against the reference's own fast_neon sources compiled into oracle/_ref).
//
// Each function cites the reference file:line it follows (paths relative to the
// reference root).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstddef>
#include <limits>
#include <algorithm>

namespace orc {

struct V2 { double x, y; };
struct V3 { double x, y, z; };

inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator-(const V3& a) { return {-a.x, -a.y, -a.z}; }
inline double dot(const V3& a, const V3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(const V3& a, const V3& b) {
  return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double norm(const V3& a) { return std::sqrt(dot(a, a)); }
// Eigen MatrixBase::normalized(): v / sqrt(squaredNorm) when squaredNorm > 0.
inline V3 normalized(const V3& a) {
  const double n2 = dot(a, a);
  if (n2 > 0.0) { const double n = std::sqrt(n2); return {a.x / n, a.y / n, a.z / n}; }
  return a;
}
inline V2 normalized(const V2& a) {
  const double n2 = a.x * a.x + a.y * a.y;
  if (n2 > 0.0) { const double n = std::sqrt(n2); return {a.x / n, a.y / n}; }
  return a;
}

struct M3 { double m[3][3]; };
inline V3 operator*(const M3& A, const V3& v) {
  return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z,
          A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
          A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}

// ---------------------------------------------------------------------------
// Rotation quaternion with minkindr semantics (w, x, y, z).
// ref: 3rd/minkindr/include/kindr/minimal/implementation/rotation-quaternion-inl.h
struct Quat {
  double w = 1, x = 0, y = 0, z = 0;
};

// Eigen::Quaternion product (Hamilton).
inline Quat quatMulRaw(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
inline double quatSqNorm(const Quat& q) { return q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z; }
inline void quatNormalize(Quat& q) {
  const double n = std::sqrt(quatSqNorm(q));
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
// ref: rotation-quaternion-inl.h:435-442 (operator* for double) + :580-589 (normalizationHelper)
inline Quat quatMul(const Quat& a, const Quat& b) {
  Quat r = quatMulRaw(a, b);
  if (std::abs(quatSqNorm(r) - 1.0) > 1.0e-4) quatNormalize(r);
  return r;
}
// ref: rotation-quaternion-inl.h:298-300 (inverse() == conjugated())
inline Quat quatConj(const Quat& q) { return {q.w, -q.x, -q.y, -q.z}; }

// Eigen QuaternionBase::_transformVector: v + 2w (q x v) + 2 q x (q x v)
// ref: rotation-quaternion-inl.h:323-326 (rotate -> q_A_B_*v)
inline V3 quatRotate(const Quat& q, const V3& v) {
  const V3 qv{q.x, q.y, q.z};
  V3 uv = cross(qv, v);
  uv = uv + uv;
  return v + uv * q.w + cross(qv, uv);
}
// ref: rotation-quaternion-inl.h:352-355 (inverseRotate -> q_A_B_.inverse()*v; Eigen inverse = conj / |q|^2)
inline V3 quatInverseRotate(const Quat& q, const V3& v) {
  const double n2 = quatSqNorm(q);
  Quat qi{q.w / n2, -q.x / n2, -q.y / n2, -q.z / n2};
  return quatRotate(qi, v);
}
// Eigen QuaternionBase::toRotationMatrix
// ref: rotation-quaternion-inl.h:461-464 (getRotationMatrix)
inline M3 quatToMatrix(const Quat& q) {
  M3 R;
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R.m[0][0] = 1.0 - (tyy + tzz); R.m[0][1] = txy - twz;         R.m[0][2] = txz + twy;
  R.m[1][0] = txy + twz;         R.m[1][1] = 1.0 - (txx + tzz); R.m[1][2] = tyz - twx;
  R.m[2][0] = txz - twy;         R.m[2][1] = tyz + twx;         R.m[2][2] = 1.0 - (txx + tyy);
  return R;
}

// ref: rotation-quaternion-inl.h:89-92
inline bool isLessThenEpsilons4thRoot(double x) {
  static const double epsilon4thRoot = std::pow(std::numeric_limits<double>::epsilon(), 1.0 / 4.0);
  return x < epsilon4thRoot;
}
// ref: rotation-quaternion-inl.h:95-100
inline double arcSinXOverX(double x) {
  if (isLessThenEpsilons4thRoot(std::fabs(x))) return 1.0 + x * x * (1.0 / 6.0);
  return std::asin(x) / x;
}
// ref: rotation-quaternion-inl.h:519-536 (exp)
inline Quat quatExp(const V3& dx) {
  const double theta = norm(dx);
  double na;
  if (isLessThenEpsilons4thRoot(theta)) {
    static const double one_over_48 = 1.0 / 48.0;
    na = 0.5 + (theta * theta) * one_over_48;
  } else {
    na = std::sin(theta * 0.5) / theta;
  }
  const double ct = std::cos(theta * 0.5);
  return {ct, dx.x * na, dx.y * na, dx.z * na};
}
// ref: rotation-quaternion-inl.h:478-516 (log)
inline V3 quatLog(const Quat& q) {
  const V3 a{q.x, q.y, q.z};
  const double na = norm(a);
  const double eta = q.w;
  double scale;
  if (std::fabs(eta) < na) {
    if (eta >= 0) scale = std::acos(eta) / na;
    else scale = -std::acos(-eta) / na;
  } else {
    if (eta > 0) scale = arcSinXOverX(na);
    else scale = -arcSinXOverX(na);
  }
  return a * (2.0 * scale);
}

// ---------------------------------------------------------------------------
// kindr::minimal::QuatTransformation
// ref: 3rd/minkindr/include/kindr/minimal/implementation/quat-transformation-inl.h
struct SE3 {
  Quat q;
  V3 t{0, 0, 0};
};
// ref: quat-transformation-inl.h:150-156
inline SE3 operator*(const SE3& a, const SE3& b) {
  SE3 r;
  r.q = quatMul(a.q, b.q);
  r.t = a.t + quatRotate(a.q, b.t);
  return r;
}
// ref: quat-transformation-inl.h:158-163 (transform)
inline V3 operator*(const SE3& T, const V3& p) { return quatRotate(T.q, p) + T.t; }
// ref: quat-transformation-inl.h:209-213 (inverse)
inline SE3 inverse(const SE3& T) {
  SE3 r;
  r.q = quatConj(T.q);
  r.t = -quatInverseRotate(T.q, T.t);
  return r;
}
// ref: quat-transformation-inl.h:229-232 + ctor :79-84 (exp: head = translation, tail = rotation vector)
inline SE3 se3Exp(const double v[6]) {
  SE3 r;
  r.q = quatExp({v[3], v[4], v[5]});
  r.t = {v[0], v[1], v[2]};
  return r;
}
// ref: quat-transformation-inl.h:234-239 (log)
inline void se3Log(const SE3& T, double out[6]) {
  const V3 l = quatLog(T.q);
  out[0] = T.t.x; out[1] = T.t.y; out[2] = T.t.z;
  out[3] = l.x; out[4] = l.y; out[5] = l.z;
}
inline SE3 se3FromArray(const double* a) {  // (qw qx qy qz tx ty tz)
  SE3 T;
  T.q = {a[0], a[1], a[2], a[3]};
  T.t = {a[4], a[5], a[6]};
  return T;
}
inline void se3ToArray(const SE3& T, double* a) {
  a[0] = T.q.w; a[1] = T.q.x; a[2] = T.q.y; a[3] = T.q.z;
  a[4] = T.t.x; a[5] = T.t.y; a[6] = T.t.z;
}

// ---------------------------------------------------------------------------
// Pinhole camera with optional radial-tangential distortion.
// ref: src/vikit/vikit_cameras/include/vikit/cameras/implementation/pinhole_projection.hpp:30-76
// ref: src/vikit/vikit_cameras/include/vikit/cameras/radial_tangential_distortion.h:34-95
// ref: src/vikit/vikit_cameras/include/vikit/cameras/no_distortion.h
struct Camera {
  double fx, fy, cx, cy;
  double k1, k2, p1, p2;
  int width, height;
  int distortion;  // 0 = none, 1 = radtan

  // radial_tangential_distortion.h:46-57 (Vector2d overload used by project3)
  inline void distort(double x, double y, double& xd, double& yd) const {
    if (distortion == 0) { xd = x; yd = y; return; }
    const double xx = x * x;
    const double yy = y * y;
    const double xy = x * y;
    const double xy2 = 2.0 * xy;
    const double r2 = xx + yy;
    const double cdist = (k1 + k2 * r2) * r2;
    xd = x + x * cdist + p1 * xy2 + p2 * (r2 + 2.0 * xx);
    yd = y + y * cdist + p2 * xy2 + p1 * (r2 + 2.0 * yy);
  }
  // radial_tangential_distortion.h:80-95
  inline void undistort(double& x, double& y) const {
    if (distortion == 0) return;
    const double x0 = x, y0 = y;
    for (int i = 0; i < 5; ++i) {
      const double xx = x * x;
      const double yy = y * y;
      const double xy = x * y;
      const double xy2 = 2 * xy;
      const double r2 = xx + yy;
      const double icdist = 1.0 / (1.0 + (k1 + k2 * r2) * r2);
      const double dx = p1 * xy2 + p2 * (r2 + 2.0 * xx);
      const double dy = p2 * xy2 + p1 * (r2 + 2.0 * yy);
      x = (x0 - dx) * icdist;
      y = (y0 - dy) * icdist;
    }
  }
  // radial_tangential_distortion.h:59-78
  inline void distJacobian(double px, double py, double J[2][2]) const {
    if (distortion == 0) { J[0][0] = 1; J[0][1] = 0; J[1][0] = 0; J[1][1] = 1; return; }
    const double xx = px * px;
    const double yy = py * py;
    const double xy = px * py;
    const double r2 = xx + yy;
    const double cdist = (k1 + k2 * r2) * r2;
    const double k2_r2_x4 = k2 * r2 * 4.0;
    const double cdist_p1 = cdist + 1.0;
    J[0][0] = cdist_p1 + k1 * 2.0 * xx + k2_r2_x4 * xx + 2.0 * p1 * py + 6.0 * p2 * px;
    J[1][1] = cdist_p1 + k1 * 2.0 * yy + k2_r2_x4 * yy + 2.0 * p2 * px + 6.0 * p1 * py;
    J[1][0] = 2.0 * k1 * xy + k2_r2_x4 * xy + 2.0 * p1 * px + 2.0 * p2 * py;
    J[0][1] = J[1][0];
  }
  // pinhole_projection.hpp:30-41
  inline V3 backProject3(const V2& px) const {
    double x = (px.x - cx) * (1.0 / fx);
    double y = (px.y - cy) * (1.0 / fy);
    undistort(x, y);
    return {x, y, 1.0};
  }
  // pinhole_projection.hpp:44-64
  inline V2 project3(const V3& p, double J[2][3] = nullptr) const {
    const double z_inv = 1 / p.z;
    const double u = p.x * z_inv, v = p.y * z_inv;
    double ud, vd;
    distort(u, v, ud, vd);
    V2 out{fx * ud + cx, fy * vd + cy};
    if (J) {
      double duv[2][3];
      duv[0][0] = z_inv; duv[0][1] = 0.0;   duv[0][2] = -p.x * z_inv * z_inv;
      duv[1][0] = 0.0;   duv[1][1] = z_inv; duv[1][2] = -p.y * z_inv * z_inv;
      double Jd[2][2];
      distJacobian(u, v, Jd);  // NB: the reference passes the *undistorted* uv here (:61)
      // (focal_matrix * Jd) * duv, Eigen left-to-right
      double FJ[2][2] = {{fx * Jd[0][0], fx * Jd[0][1]}, {fy * Jd[1][0], fy * Jd[1][1]}};
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) J[r][c] = FJ[r][0] * duv[0][c] + FJ[r][1] * duv[1][c];
    }
    return out;
  }
  // pinhole_projection.hpp:66-70
  inline double errorMultiplier() const { return std::abs(fx); }
  // pinhole_projection.hpp:72-76
  inline double getAngleError(double img_err) const {
    return std::atan(img_err / (2.0 * fx)) + std::atan(img_err / (2.0 * fy));
  }
  // src/vikit/vikit_cameras/include/vikit/cameras/implementation/camera_geometry_base.hpp:7-15
  inline bool isKeypointVisible(double x, double y) const {
    return x >= 0.0 && y >= 0.0 && x < static_cast<double>(width) && y < static_cast<double>(height);
  }
  // camera_geometry_base.hpp:17-29 (instantiated with Vector2i in depth_filter.cpp:414-418)
  inline bool isKeypointVisibleWithMarginInt(int x, int y, int margin) const {
    return x >= margin && y >= margin && x < (width - margin) && y < (height - margin);
  }
};

// A strided 8-bit image view (stands in for cv::Mat as used on this path: data/step/cols/rows).
struct Img {
  const uint8_t* data;
  int cols, rows, step;
};

}  // namespace orc
