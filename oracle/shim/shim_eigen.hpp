// TEST INFRASTRUCTURE ONLY (oracle/): a small fixed-size linear-algebra stand-in with Eigen's spelling, so that the
// reference's direct front-end sources can be compiled from /root/reference without Eigen (not installed here).
// It has value semantics and no expression templates: `H += J * J.transpose()` forms the rank-1 product and adds it
// coefficient by coefficient — the operations Eigen's lazy evaluation performs for these small fixed sizes, in the same
// order (products summed left to right over the inner index). inverse() uses the closed cofactor forms Eigen uses for
// sizes <= 4 (Eigen/src/LU/InverseImpl.h); this part is restated, not Eigen's code, and DESIGN.md says so.
#pragma once
#include <algorithm>  // Eigen/Core pulls <algorithm> in; reference headers rely on it (std::count, std::max)
#include <cmath>
#include <cstddef>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_STRONG_INLINE inline
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
typedef std::ptrdiff_t Index;

template <typename T> using aligned_allocator = std::allocator<T>;

template <typename T, int R, int C, int Opt = 0, int MR = R, int MC = C, typename Enable = void> class Matrix;

// CRTP base so that reference templates written against Eigen::MatrixBase<Derived> (occupancy_grid_2d.h:82-88) bind.
template <typename Derived> struct MatrixBase {
  const Derived& derived() const { return static_cast<const Derived&>(*this); }
  Derived& derived() { return static_cast<Derived&>(*this); }
  template <typename D2 = Derived> auto operator()(int i) const -> decltype(std::declval<const D2&>().coeffAt(i)) { return derived().coeffAt(i); }
  template <typename D2 = Derived> auto operator[](int i) const -> decltype(std::declval<const D2&>().coeffAt(i)) { return derived().coeffAt(i); }
};
#define EIGEN_STATIC_ASSERT_MATRIX_SPECIFIC_SIZE(TYPE, ROWS, COLS) \
  static_assert(TYPE::RowsAtCompileTime == ROWS && TYPE::ColsAtCompileTime == COLS, "matrix of the wrong size")
#define EIGEN_STATIC_ASSERT_VECTOR_SPECIFIC_SIZE(TYPE, SIZE) \
  static_assert(TYPE::RowsAtCompileTime * TYPE::ColsAtCompileTime == SIZE, "vector of the wrong size")

template <typename M> struct CommaInit {
  M& m; int i;
  CommaInit(M& m_, typename M::Scalar v) : m(m_), i(0) { put(v); }
  void put(typename M::Scalar v) { const int r = i / M::ColsAtCompileTime, c = i % M::ColsAtCompileTime; m(r, c) = v; ++i; }
  CommaInit& operator,(typename M::Scalar v) { put(v); return *this; }
};

// ---------------------------------------------------------------------------------------------- fixed size
template <typename T, int R, int C, int Opt, int MR, int MC>
class Matrix<T, R, C, Opt, MR, MC, typename std::enable_if<(R > 0 && C > 0)>::type>
    : public MatrixBase<Matrix<T, R, C, Opt, MR, MC>> {
 public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = R * C };
  T d[R * C];  // column-major

  Matrix() {}
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 2, int>::type = 0>
  Matrix(T x, T y) { d[0] = x; d[1] = y; }
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 3, int>::type = 0>
  Matrix(T x, T y, T z) { d[0] = x; d[1] = y; d[2] = z; }
  template <int RR = R, int CC = C, typename std::enable_if<RR * CC == 4 && (RR == 1 || CC == 1), int>::type = 0>
  Matrix(T x, T y, T z, T w) { d[0] = x; d[1] = y; d[2] = z; d[3] = w; }

  static Matrix Zero() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(0); return m; }
  static Matrix Zero(int, int) { return Zero(); }
  static Matrix Zero(int) { return Zero(); }
  static Matrix Ones() { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = T(1); return m; }
  static Matrix Constant(T v) { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = v; return m; }
  static Matrix Identity() { Matrix m = Zero(); for (int i = 0; i < (R < C ? R : C); ++i) m(i, i) = T(1); return m; }
  Matrix& setZero() { *this = Zero(); return *this; }
  Matrix& setIdentity() { *this = Identity(); return *this; }
  Matrix& setConstant(T v) { *this = Constant(v); return *this; }
  Matrix& setOnes() { *this = Ones(); return *this; }

  static constexpr int rows() { return R; }
  static constexpr int cols() { return C; }
  static constexpr int size() { return R * C; }
  T* data() { return d; }
  const T* data() const { return d; }
  T& operator()(int i, int j) { return d[j * R + i]; }
  const T& operator()(int i, int j) const { return d[j * R + i]; }
  T& operator()(int i) { return d[i]; }
  const T& operator()(int i) const { return d[i]; }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
  T& x() { return d[0]; } const T& x() const { return d[0]; }
  T& y() { return d[1]; } const T& y() const { return d[1]; }
  T& z() { return d[2]; } const T& z() const { return d[2]; }
  T& w() { return d[3]; } const T& w() const { return d[3]; }
  const T& coeffAt(int i) const { return d[i]; }
  T& coeffRef(int i, int j) { return (*this)(i, j); }
  const T& coeff(int i, int j) const { return (*this)(i, j); }

  CommaInit<Matrix> operator<<(T v) { return CommaInit<Matrix>(*this, v); }

  Matrix operator+(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] + o.d[i]; return m; }
  Matrix operator-(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] - o.d[i]; return m; }
  Matrix operator-() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = -d[i]; return m; }
  Matrix operator*(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * s; return m; }
  Matrix operator/(T s) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] / s; return m; }
  friend Matrix operator*(T s, const Matrix& a) { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = s * a.d[i]; return m; }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] += o.d[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < R * C; ++i) d[i] -= o.d[i]; return *this; }
  Matrix& operator*=(T s) { for (int i = 0; i < R * C; ++i) d[i] *= s; return *this; }
  Matrix& operator/=(T s) { for (int i = 0; i < R * C; ++i) d[i] /= s; return *this; }
  bool operator==(const Matrix& o) const { for (int i = 0; i < R * C; ++i) if (!(d[i] == o.d[i])) return false; return true; }
  bool operator!=(const Matrix& o) const { return !(*this == o); }

  template <int K>
  Matrix<T, R, K> operator*(const Matrix<T, C, K>& o) const {
    Matrix<T, R, K> m;
    for (int j = 0; j < K; ++j)
      for (int i = 0; i < R; ++i) {
        T s = (*this)(i, 0) * o(0, j);
        for (int k = 1; k < C; ++k) s += (*this)(i, k) * o(k, j);
        m(i, j) = s;
      }
    return m;
  }
  Matrix<T, C, R> transpose() const {
    Matrix<T, C, R> m;
    for (int i = 0; i < R; ++i) for (int j = 0; j < C; ++j) m(j, i) = (*this)(i, j);
    return m;
  }
  template <typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> m; for (int i = 0; i < R * C; ++i) m.d[i] = static_cast<U>(d[i]); return m; }
  T dot(const Matrix& o) const { T s = d[0] * o.d[0]; for (int i = 1; i < R * C; ++i) s += d[i] * o.d[i]; return s; }
  T squaredNorm() const { return dot(*this); }
  T norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const { const T n = norm(); return n > T(0) ? *this / n : *this; }
  void normalize() { *this = normalized(); }
  T sum() const { T s = d[0]; for (int i = 1; i < R * C; ++i) s += d[i]; return s; }
  T maxCoeff() const { T s = d[0]; for (int i = 1; i < R * C; ++i) if (d[i] > s) s = d[i]; return s; }
  T minCoeff() const { T s = d[0]; for (int i = 1; i < R * C; ++i) if (d[i] < s) s = d[i]; return s; }
  Matrix cwiseAbs() const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = std::abs(d[i]); return m; }
  Matrix cwiseProduct(const Matrix& o) const { Matrix m; for (int i = 0; i < R * C; ++i) m.d[i] = d[i] * o.d[i]; return m; }
  template <typename U = T> Matrix<T, 3, 1> cross(const Matrix<T, 3, 1>& o) const {
    return Matrix<T, 3, 1>(d[1] * o.d[2] - d[2] * o.d[1], d[2] * o.d[0] - d[0] * o.d[2], d[0] * o.d[1] - d[1] * o.d[0]);
  }
  template <int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> m; for (int i = 0; i < N; ++i) m.d[i] = d[i]; return m; }
  template <int N> Matrix<T, N, 1> tail() const { Matrix<T, N, 1> m; for (int i = 0; i < N; ++i) m.d[i] = d[R * C - N + i]; return m; }
  Matrix<T, R, 1> col(int j) const { Matrix<T, R, 1> m; for (int i = 0; i < R; ++i) m.d[i] = (*this)(i, j); return m; }
  Matrix<T, 1, C> row(int i) const { Matrix<T, 1, C> m; for (int j = 0; j < C; ++j) m.d[j] = (*this)(i, j); return m; }
  bool allFinite() const { for (int i = 0; i < R * C; ++i) if (!std::isfinite((double)d[i])) return false; return true; }

  T determinant() const {
    static_assert(R == C && R <= 3, "determinant: sizes 1..3 only");
    const Matrix& m = *this;
    if (R == 1) return m(0, 0);
    if (R == 2) return m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1);
    return m(0, 0) * (m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1)) - m(0, 1) * (m(1, 0) * m(2, 2) - m(1, 2) * m(2, 0)) +
           m(0, 2) * (m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0));
  }
  // Closed forms of Eigen/src/LU/InverseImpl.h (compute_inverse<MatrixType, ResultType, 2|3|4>), scalar path.
  Matrix inverse() const {
    static_assert(R == C && R >= 1 && R <= 4, "inverse: sizes 1..4 only");
    const Matrix& m = *this;
    Matrix r;
    if (R == 1) { r(0, 0) = T(1) / m(0, 0); return r; }
    if (R == 2) {
      const T invdet = T(1) / (m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1));
      r(0, 0) = m(1, 1) * invdet; r(1, 0) = -m(1, 0) * invdet; r(0, 1) = -m(0, 1) * invdet; r(1, 1) = m(0, 0) * invdet;
      return r;
    }
    if (R == 3) {
      auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
      };
      const T c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
      const T det = (c00 * m(0, 0) + c10 * m(1, 0)) + c20 * m(2, 0);
      const T invdet = T(1) / det;
      r(0, 0) = c00 * invdet; r(0, 1) = c10 * invdet; r(0, 2) = c20 * invdet;
      r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
      r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
      return r;
    }
    // 4x4: general_det3_helper / cofactor_4x4 scheme
    auto det3h = [&](int i1, int i2, int i3, int j1, int j2, int j3) {
      return m(i1, j1) * (m(i2, j2) * m(i3, j3) - m(i2, j3) * m(i3, j2));
    };
    auto cof4 = [&](int i, int j) {
      const int i1 = (i + 1) % 4, i2 = (i + 2) % 4, i3 = (i + 3) % 4, j1 = (j + 1) % 4, j2 = (j + 2) % 4, j3 = (j + 3) % 4;
      const T v = det3h(i1, i2, i3, j1, j2, j3) + det3h(i2, i3, i1, j1, j2, j3) + det3h(i3, i1, i2, j1, j2, j3);
      return ((i + j) & 1) ? -v : v;
    };
    T c[4][4];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) c[i][j] = cof4(i, j);
    // result(j, i) = cofactor(i, j); det = column 0 of m dotted with row 0 of the result
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r(j, i) = c[i][j];
    const T det = ((m(0, 0) * r(0, 0) + m(1, 0) * r(0, 1)) + m(2, 0) * r(0, 2)) + m(3, 0) * r(0, 3);
    for (int i = 0; i < 16; ++i) r.d[i] /= det;
    return r;
  }
};

// ---------------------------------------------------------------------------------------------- dynamic size
// Only what the reference's type aliases and containers need to exist (svo/common/types.h); column-major storage.
template <typename T, int R, int C, int Opt, int MR, int MC>
class Matrix<T, R, C, Opt, MR, MC, typename std::enable_if<(R < 0 || C < 0)>::type> {
 public:
  typedef T Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C };
  Matrix() : r_(R > 0 ? R : 0), c_(C > 0 ? C : 0) {}
  Matrix(int r, int c) { resize(r, c); }
  explicit Matrix(int n) { if (C == 1) resize(n, 1); else resize(1, n); }
  void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, T()); }
  void resize(int n) { if (C == 1) resize(n, 1); else resize(R > 0 ? R : 1, n); }
  void conservativeResize(int r, int c) {
    std::vector<T> n((size_t)r * c, T());
    for (int j = 0; j < (c < c_ ? c : c_); ++j) for (int i = 0; i < (r < r_ ? r : r_); ++i) n[(size_t)j * r + i] = v_[(size_t)j * r_ + i];
    v_.swap(n); r_ = r; c_ = c;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
  T& operator()(int i, int j) { return v_[(size_t)j * r_ + i]; }
  const T& operator()(int i, int j) const { return v_[(size_t)j * r_ + i]; }
  T& operator()(int i) { return v_[i]; }
  const T& operator()(int i) const { return v_[i]; }
  T& operator[](int i) { return v_[i]; }
  const T& operator[](int i) const { return v_[i]; }
  template <int RR = R> Matrix<T, (RR > 0 ? RR : 1), 1> col(int j) const {
    Matrix<T, (RR > 0 ? RR : 1), 1> m; for (int i = 0; i < r_; ++i) m.d[i] = (*this)(i, j); return m;
  }
  void setZero() { for (auto& x : v_) x = T(); }
 private:
  int r_, c_;
  std::vector<T> v_;
};

template <typename T> using Ref = T&;

typedef Matrix<float, 2, 1> Vector2f;  typedef Matrix<double, 2, 1> Vector2d;  typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, 3, 1> Vector3f;  typedef Matrix<double, 3, 1> Vector3d;  typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, 4, 1> Vector4f;  typedef Matrix<double, 4, 1> Vector4d;  typedef Matrix<int, 4, 1> Vector4i;
typedef Matrix<float, 2, 2> Matrix2f;  typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<float, 3, 3> Matrix3f;  typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;  typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;

}  // namespace Eigen
