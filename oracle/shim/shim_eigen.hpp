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
#include <cstdlib>
#include <functional>
#include <limits>
#include <ostream>
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
struct NoChange_t {};
static const NoChange_t NoChange = NoChange_t();
enum { ComputeFullU = 4, ComputeFullV = 16, ComputeThinU = 8, ComputeThinV = 32 };

template <typename T> using aligned_allocator = std::allocator<T>;

template <typename T, int R, int C, int Opt = 0, int MR = R, int MC = C, typename Enable = void> class Matrix;
// Stands for any Eigen expression the checker never evaluates (Point::triangulateLinear's row-wise / column-wise reductions and its
// rank-revealing QR, src/svo_common/src/point.cpp:169-213): it type-checks, and aborts when reached.
struct ShimAny {
  ShimAny transpose() const { std::abort(); }
  ShimAny sum() const { std::abort(); }
  ShimAny squaredNorm() const { std::abort(); }
  ShimAny asDiagonal() const { std::abort(); }
  ShimAny inverse() const { std::abort(); }
  ShimAny rowwise() const { std::abort(); }
  ShimAny colwise() const { std::abort(); }
  template <typename X> ShimAny cwiseProduct(const X&) const { std::abort(); }
  template <typename M> operator M() const { std::abort(); }
};
template <typename X> inline ShimAny operator*(const X&, const ShimAny&) { std::abort(); }
template <typename X> inline ShimAny operator-(const X&, const ShimAny&) { std::abort(); }
inline ShimAny operator-(const ShimAny&, const ShimAny&) { std::abort(); }
template <typename M> struct ColPivHouseholderQR {  // declared for type-checking only
  void setThreshold(double) { std::abort(); }
  size_t rank() const { std::abort(); }
  template <typename B> ShimAny solve(const B&) const { std::abort(); }
};

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

// Eigen's CommaInitializer: scalars and blocks are placed left to right; a row of blocks ends when the columns are used up
// and the next one starts below it (so `A << v1, v2` with 3-vectors fills the COLUMNS of a 3x2 matrix).
template <typename M> struct CommaInit {
  M& m; int row, col, block_rows;
  typedef typename M::Scalar S;
  CommaInit(M& m_, S v) : m(m_), row(0), col(0), block_rows(1) { put(v); }
  template <int P, int Q> CommaInit(M& m_, const Matrix<S, P, Q>& v) : m(m_), row(0), col(0), block_rows(P) { putBlock(v); }
  void advance(int rows_of_block) {
    if (col == M::ColsAtCompileTime) { row += block_rows; col = 0; block_rows = rows_of_block; }
  }
  void put(S v) { advance(1); m(row, col) = v; col += 1; }
  template <int P, int Q> void putBlock(const Matrix<S, P, Q>& v) {
    advance(P);
    for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(row + i, col + j) = v(i, j);
    col += Q;
  }
  CommaInit& operator,(S v) { put(v); return *this; }
  template <int P, int Q> CommaInit& operator,(const Matrix<S, P, Q>& v) { putBlock(v); return *this; }
  M& finished() { return m; }
};

// Writable P x Q window into a matrix (block<P,Q>(i,j), col(j), head<N>(), ...). Reads convert to a value.
template <typename M, int P, int Q> struct BlockRef;
template <typename M, int P, int Q> using Block = BlockRef<M, P, Q>;  // only named in declarations of the tracker headers
template <typename M, int P, int Q> struct BlockRef {
  typedef typename M::Scalar T;
  typedef Matrix<T, P, Q> Value;
  M& m; int i0, j0;
  BlockRef(M& m_, int i, int j) : m(m_), i0(i), j0(j) {}
  Value eval() const { Value v; for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) v(i, j) = m(i0 + i, j0 + j); return v; }
  operator Value() const { return eval(); }
  BlockRef& operator=(const Value& v) { for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i0 + i, j0 + j) = v(i, j); return *this; }
  template <int PP = P, int QQ = Q, typename std::enable_if<(PP != QQ) && (PP == 1 || QQ == 1), int>::type = 0>
  BlockRef& operator=(const Matrix<T, QQ, PP>& v) { return *this = Value(v.transpose()); }  // row <-> column vector
  BlockRef& operator=(const BlockRef& o) { return *this = o.eval(); }
  template <typename M2> BlockRef& operator=(const BlockRef<M2, P, Q>& o) { return *this = o.eval(); }
  BlockRef& operator+=(const Value& v) { for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i0 + i, j0 + j) += v(i, j); return *this; }
  BlockRef& operator-=(const Value& v) { for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i0 + i, j0 + j) -= v(i, j); return *this; }
  BlockRef& operator*=(T s) { for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i0 + i, j0 + j) *= s; return *this; }
  BlockRef& setZero() { return *this = Value::Zero(); }
  BlockRef& setIdentity() { return *this = Value::Identity(); }
  T& operator()(int i, int j) { return m(i0 + i, j0 + j); }
  T& operator()(int i) { return Q == 1 ? m(i0 + i, j0) : m(i0, j0 + i); }
  T& operator[](int i) { return (*this)(i); }
  Value operator-() const { return -eval(); }
  Value operator*(T s) const { return eval() * s; }
  Value operator/(T s) const { return eval() / s; }
  friend Value operator*(T s, const BlockRef& b) { return s * b.eval(); }
  Value operator+(const Value& o) const { return eval() + o; }
  Value operator-(const Value& o) const { return eval() - o; }
  template <int K> Matrix<T, P, K> operator*(const Matrix<T, Q, K>& o) const { return eval() * o; }
  template <typename U> Matrix<U, P, Q> cast() const { return eval().template cast<U>(); }
  Matrix<T, Q, P> transpose() const { return eval().transpose(); }
  T norm() const { return eval().norm(); }
  T squaredNorm() const { return eval().squaredNorm(); }
  T dot(const Value& o) const { return eval().dot(o); }
  Value normalized() const { return eval().normalized(); }
  T x() const { return eval()[0]; }
  T y() const { return eval()[1]; }
  T z() const { return eval()[2]; }
};

// Runtime-size window into a dynamic matrix (block(i,j,p,q), middleCols(j,n), segment(i,n)).
template <typename M> struct DynBlockRef {
  typedef typename M::Scalar T;
  M& m; int i0, j0, p, q;
  DynBlockRef(M& m_, int i, int j, int p_, int q_) : m(m_), i0(i), j0(j), p(p_), q(q_) {}
  template <typename M2> DynBlockRef& operator=(const M2& v) { for (int j = 0; j < q; ++j) for (int i = 0; i < p; ++i) m(i0 + i, j0 + j) = v(i, j); return *this; }
  DynBlockRef& setConstant(T v) { for (int j = 0; j < q; ++j) for (int i = 0; i < p; ++i) m(i0 + i, j0 + j) = v; return *this; }
  DynBlockRef& setZero() { return setConstant(T(0)); }
  T& operator()(int i, int j) { return m(i0 + i, j0 + j); }
  int rows() const { return p; }
  int cols() const { return q; }
};

// H.ldlt().solve(g): symmetric diagonal pivoting on max |A_kk|, in place on the lower triangle; zero pivots give a zero
// solution component. Restated from Eigen/src/Cholesky/LDLT.h (ldlt_inplace<Lower>::unblocked, LDLT::_solve_impl) — the
// same restatement the oracle uses (oracle/orc_sparse_align.hpp:ldltSolve); this is Eigen's arithmetic, not the
// reference's, and is NOT pinned by compiling the reference.
template <typename T, int D> struct LDLTSolver {
  Matrix<T, D, D> A;
  int transp[D];
  explicit LDLTSolver(const Matrix<T, D, D>& H) : A(H) {
    for (int k = 0; k < D; ++k) {
      int piv = k; T big = std::abs(A(k, k));
      for (int i = k + 1; i < D; ++i) if (std::abs(A(i, i)) > big) { big = std::abs(A(i, i)); piv = i; }
      transp[k] = piv;
      if (piv != k) {
        for (int j = 0; j < k; ++j) std::swap(A(k, j), A(piv, j));
        for (int i = piv + 1; i < D; ++i) std::swap(A(i, k), A(i, piv));
        std::swap(A(k, k), A(piv, piv));
        for (int i = k + 1; i < piv; ++i) { const T tmp = A(i, k); A(i, k) = A(piv, i); A(piv, i) = tmp; }
      }
      T temp[D];
      for (int j = 0; j < k; ++j) temp[j] = A(j, j) * A(k, j);
      for (int j = 0; j < k; ++j) A(k, k) -= A(k, j) * temp[j];
      for (int i = k + 1; i < D; ++i) for (int j = 0; j < k; ++j) A(i, k) -= A(i, j) * temp[j];
      const T akk = A(k, k);
      if (std::abs(akk) > T(0)) for (int i = k + 1; i < D; ++i) A(i, k) /= akk;
    }
  }
  Matrix<T, D, 1> solve(const Matrix<T, D, 1>& g) const {
    Matrix<T, D, 1> x = g;
    for (int k = 0; k < D; ++k) std::swap(x[k], x[transp[k]]);
    for (int i = 0; i < D; ++i) for (int j = 0; j < i; ++j) x[i] -= A(i, j) * x[j];
    const T tolerance = T(1) / std::numeric_limits<T>::max();
    for (int i = 0; i < D; ++i) { if (std::abs(A(i, i)) > tolerance) x[i] /= A(i, i); else x[i] = T(0); }
    for (int i = D - 1; i >= 0; --i) for (int j = i + 1; j < D; ++j) x[i] -= A(j, i) * x[j];
    for (int k = D - 1; k >= 0; --k) std::swap(x[k], x[transp[k]]);
    return x;
  }
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
  template <int RR = R, int CC = C, typename std::enable_if<(RR != CC) && (RR == 1 || CC == 1), int>::type = 0>
  Matrix(const Matrix<T, CC, RR>& v) { for (int i = 0; i < R * C; ++i) d[i] = v.d[i]; }  // row <-> column vector (Eigen allows it)
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
  template <int P, int Q> CommaInit<Matrix> operator<<(const Matrix<T, P, Q>& v) { return CommaInit<Matrix>(*this, v); }

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
  template <int N> BlockRef<Matrix, (C == 1 ? N : 1), (C == 1 ? 1 : N)> head() { return BlockRef<Matrix, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(*this, 0, 0); }
  template <int N> BlockRef<Matrix, (C == 1 ? N : 1), (C == 1 ? 1 : N)> tail() {
    return BlockRef<Matrix, (C == 1 ? N : 1), (C == 1 ? 1 : N)>(*this, C == 1 ? R - N : 0, C == 1 ? 0 : C - N);
  }
  template <int P, int Q> Matrix<T, P, Q> block(int i0, int j0) const { Matrix<T, P, Q> m; for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i, j) = (*this)(i0 + i, j0 + j); return m; }
  template <int P, int Q> BlockRef<Matrix, P, Q> block(int i0, int j0) { return BlockRef<Matrix, P, Q>(*this, i0, j0); }
  template <int P, int Q> Matrix<T, P, Q> topLeftCorner() const { return block<P, Q>(0, 0); }
  template <int P, int Q> BlockRef<Matrix, P, Q> topLeftCorner() { return BlockRef<Matrix, P, Q>(*this, 0, 0); }
  template <int P, int Q> Matrix<T, P, Q> topRightCorner() const { return block<P, Q>(0, C - Q); }
  template <int P, int Q> BlockRef<Matrix, P, Q> topRightCorner() { return BlockRef<Matrix, P, Q>(*this, 0, C - Q); }
  template <int P, int Q> Matrix<T, P, Q> bottomRightCorner() const { return block<P, Q>(R - P, C - Q); }
  template <int P, int Q> BlockRef<Matrix, P, Q> bottomRightCorner() { return BlockRef<Matrix, P, Q>(*this, R - P, C - Q); }
  template <int P, int Q> Matrix<T, P, Q> bottomLeftCorner() const { return block<P, Q>(R - P, 0); }
  template <int P, int Q> BlockRef<Matrix, P, Q> bottomLeftCorner() { return BlockRef<Matrix, P, Q>(*this, R - P, 0); }
  template <int Q> Matrix<T, R, Q> leftCols() const { return block<R, Q>(0, 0); }
  template <int Q> BlockRef<Matrix, R, Q> leftCols() { return BlockRef<Matrix, R, Q>(*this, 0, 0); }
  template <int Q> Matrix<T, R, Q> rightCols() const { return block<R, Q>(0, C - Q); }
  template <int Q> BlockRef<Matrix, R, Q> rightCols() { return BlockRef<Matrix, R, Q>(*this, 0, C - Q); }
  BlockRef<Matrix, R, 1> col(int j) { return BlockRef<Matrix, R, 1>(*this, 0, j); }
  BlockRef<Matrix, 1, C> row(int i) { return BlockRef<Matrix, 1, C>(*this, i, 0); }
  Matrix& noalias() { return *this; }
  const Matrix& eval() const { return *this; }
  Matrix<T, (R < C ? R : C), 1> diagonal() const { Matrix<T, (R < C ? R : C), 1> v; for (int i = 0; i < (R < C ? R : C); ++i) v[i] = (*this)(i, i); return v; }
  Matrix<T, R * C, R * C> asDiagonal() const { Matrix<T, R * C, R * C> m = Matrix<T, R * C, R * C>::Zero(); for (int i = 0; i < R * C; ++i) m(i, i) = d[i]; return m; }
  LDLTSolver<T, R> ldlt() const { static_assert(R == C, "ldlt: square only"); return LDLTSolver<T, R>(*this); }
  ColPivHouseholderQR<Matrix> colPivHouseholderQr() const { std::abort(); }
  bool isApprox(const Matrix& o, T prec = T(1e-12)) const { return (*this - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm()); }
  Matrix& setRandom() { for (int i = 0; i < R * C; ++i) d[i] = T(2) * T(std::rand()) / T(RAND_MAX) - T(1); return *this; }
  friend std::ostream& operator<<(std::ostream& os, const Matrix& m) {
    for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) os << m(i, j) << (j + 1 < C ? " " : ""); if (i + 1 < R) os << "\n"; }
    return os;
  }
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
  template <typename M2, int P, int Q> Matrix(const BlockRef<M2, P, Q>& b) : Matrix(b.eval()) {}
  template <int RR, int CC, typename std::enable_if<(RR > 0 && CC > 0), int>::type = 0>
  Matrix(const Matrix<T, RR, CC>& m) { resize(RR, CC); for (int j = 0; j < CC; ++j) for (int i = 0; i < RR; ++i) (*this)(i, j) = m(i, j); }
  void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, Store()); }
  void resize(int n) { if (C == 1) resize(n, 1); else resize(R > 0 ? R : 1, n); }
  void resize(NoChange_t, int c) { resize(r_, c); }
  DynBlockRef<Matrix> block(int i0, int j0, int p, int q) { return DynBlockRef<Matrix>(*this, i0, j0, p, q); }
  DynBlockRef<Matrix> middleCols(int j0, int n) { return DynBlockRef<Matrix>(*this, 0, j0, r_, n); }
  DynBlockRef<Matrix> segment(int i0, int n) { return C == 1 ? DynBlockRef<Matrix>(*this, i0, 0, n, 1) : DynBlockRef<Matrix>(*this, 0, i0, 1, n); }
  void setConstant(T v) { for (int i = 0; i < r_ * c_; ++i) data()[i] = v; }
  void setConstant(int n, T v) { resize(n); setConstant(v); }
  ShimAny transpose() const { std::abort(); }
  ShimAny rowwise() const { std::abort(); }
  template <typename X> ShimAny cwiseProduct(const X&) const { std::abort(); }
  // `m.array().rowwise() / m.colwise().norm().array()` (normalise every column: feature_tracker.cpp:156) is evaluated for real
  struct ColNorms { std::vector<T> n; const ColNorms& array() const { return *this; } };
  struct Colwise {
    const Matrix& m;
    ColNorms norm() const {
      ColNorms r;
      for (int j = 0; j < m.cols(); ++j) { T s = T(0); for (int i = 0; i < m.rows(); ++i) s += m(i, j) * m(i, j); r.n.push_back(std::sqrt(s)); }
      return r;
    }
    ShimAny squaredNorm() const { std::abort(); }
    ShimAny sum() const { std::abort(); }
  };
  struct Rowwise {
    const Matrix& m;
    Matrix operator/(const ColNorms& c) const {
      Matrix r = m;
      for (int j = 0; j < m.cols(); ++j) for (int i = 0; i < m.rows(); ++i) r(i, j) = m(i, j) / c.n[j];
      return r;
    }
  };
  struct ArrayView { const Matrix& m; Rowwise rowwise() const { return Rowwise{m}; } };
  Colwise colwise() const { return Colwise{*this}; }
  ArrayView array() const { return ArrayView{*this}; }
  Matrix head(int n) const { Matrix r; r.resize(n); for (int i = 0; i < n; ++i) r.data()[i] = data()[i]; return r; }           // vectors
  Matrix leftCols(int n) const { Matrix r; r.resize(r_, n); for (int j = 0; j < n; ++j) for (int i = 0; i < r_; ++i) r(i, j) = (*this)(i, j); return r; }
  void conservativeResize(int n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(r_, n); }
  void conservativeResize(NoChange_t, int c) { conservativeResize(r_, c); }
  void conservativeResize(int r, NoChange_t) { conservativeResize(r, c_); }
  void resize(int r, NoChange_t) { resize(r, c_); }
  template <int RR = R> BlockRef<Matrix, (RR > 0 ? RR : 1), 1> col(int j) { return BlockRef<Matrix, (RR > 0 ? RR : 1), 1>(*this, 0, j); }
  template <int P, int Q> BlockRef<Matrix, P, Q> block(int i0, int j0) { return BlockRef<Matrix, P, Q>(*this, i0, j0); }
  template <int P, int Q> Matrix<T, P, Q> block(int i0, int j0) const { Matrix<T, P, Q> m; for (int j = 0; j < Q; ++j) for (int i = 0; i < P; ++i) m(i, j) = (*this)(i0 + i, j0 + j); return m; }
  void conservativeResize(int r, int c) {
    std::vector<Store> n((size_t)r * c, Store());
    for (int j = 0; j < (c < c_ ? c : c_); ++j) for (int i = 0; i < (r < r_ ? r : r_); ++i) n[(size_t)j * r + i] = v_[(size_t)j * r_ + i];
    v_.swap(n); r_ = r; c_ = c;
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  T* data() { return reinterpret_cast<T*>(v_.data()); }
  const T* data() const { return reinterpret_cast<const T*>(v_.data()); }
  T& operator()(int i, int j) { return data()[(size_t)j * r_ + i]; }
  const T& operator()(int i, int j) const { return data()[(size_t)j * r_ + i]; }
  T& operator()(int i) { return data()[i]; }
  const T& operator()(int i) const { return data()[i]; }
  T& operator[](int i) { return data()[i]; }
  const T& operator[](int i) const { return data()[i]; }
  template <int RR = R> Matrix<T, (RR > 0 ? RR : 1), 1> col(int j) const {
    Matrix<T, (RR > 0 ? RR : 1), 1> m; for (int i = 0; i < r_; ++i) m.d[i] = (*this)(i, j); return m;
  }
  T norm() const { T s = T(0); for (int i = 0; i < r_ * c_; ++i) s += data()[i] * data()[i]; return std::sqrt(s); }
  void setZero() { for (auto& x : v_) x = Store(); }
 private:
  typedef typename std::conditional<std::is_same<T, bool>::value, unsigned char, T>::type Store;  // vector<bool> has no T&
  int r_, c_;
  std::vector<Store> v_;
};

// Eigen::Ref as a value with write-back: a mutable Ref copies its target in and stores the final value back when it
// dies; a const Ref is a plain copy. Equivalent to a view for the single-threaded, non-aliased uses in the reference.
template <typename M> class Ref : public M {
 public:
  Ref(M& m) : M(m), wb_([&m](const M& v) { m = v; }) {}
  template <typename MM, int P, int Q> Ref(BlockRef<MM, P, Q> b) : M(b.eval()), wb_([b](const M& v) mutable { b = v; }) {}
  Ref(const Ref& o) : M(static_cast<const M&>(o)), wb_(o.wb_) {}
  ~Ref() { if (wb_) wb_(*this); }
  Ref& operator=(const M& v) { M::operator=(v); return *this; }
  Ref& operator=(const Ref& v) { M::operator=(static_cast<const M&>(v)); return *this; }
 private:
  std::function<void(const M&)> wb_;
};
template <typename M> class Ref<const M> : public M {
 public:
  Ref(const M& m) : M(m) {}
  template <typename MM, int P, int Q> Ref(const BlockRef<MM, P, Q>& b) : M(b.eval()) {}
};

template <typename M> class Map {  // view on caller memory, column-major, fixed size
 public:
  typedef typename M::Scalar Scalar;
  explicit Map(Scalar* p) : p_(p) {}
  Scalar& operator()(int i, int j) { return p_[j * M::RowsAtCompileTime + i]; }
  Scalar& operator()(int i) { return p_[i]; }
  Map& setZero() { for (int i = 0; i < M::SizeAtCompileTime; ++i) p_[i] = Scalar(0); return *this; }
  Map& operator=(const M& m) { for (int i = 0; i < M::SizeAtCompileTime; ++i) p_[i] = m.d[i]; return *this; }
  operator M() const { M m; for (int i = 0; i < M::SizeAtCompileTime; ++i) m.d[i] = p_[i]; return m; }
 private:
  Scalar* p_;
};

template <typename T> class AngleAxis;
template <typename M> class JacobiSVD;  // declared only (nearest-orthonormal-matrix helper of minkindr, never instantiated)

// Eigen::Quaternion, scalar path of Eigen/src/Geometry/Quaternion.h: coefficients stored (x, y, z, w).
template <typename T> class Quaternion {
 public:
  typedef T Scalar;
  typedef Matrix<T, 3, 1> Vector3;
  typedef Matrix<T, 3, 3> Matrix3;
  Quaternion() {}
  Quaternion(T w, T x, T y, T z) { c_[0] = x; c_[1] = y; c_[2] = z; c_[3] = w; }
  explicit Quaternion(const Matrix<T, 4, 1>& coeffs) : c_(coeffs) {}
  explicit Quaternion(const AngleAxis<T>& aa);
  explicit Quaternion(const Matrix3& m) {  // quaternionbase_assign_impl<Other,3,3>
    T t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > T(0)) {
      t = std::sqrt(t + T(1.0));
      w() = T(0.5) * t;
      t = T(0.5) / t;
      x() = (m(2, 1) - m(1, 2)) * t; y() = (m(0, 2) - m(2, 0)) * t; z() = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + T(1.0));
      c_[i] = T(0.5) * t;
      t = T(0.5) / t;
      w() = (m(k, j) - m(j, k)) * t;
      c_[j] = (m(j, i) + m(i, j)) * t;
      c_[k] = (m(k, i) + m(i, k)) * t;
    }
  }
  static Quaternion Identity() { return Quaternion(T(1), T(0), T(0), T(0)); }
  Quaternion& setIdentity() { *this = Identity(); return *this; }
  T& x() { return c_[0]; } T& y() { return c_[1]; } T& z() { return c_[2]; } T& w() { return c_[3]; }
  const T& x() const { return c_[0]; } const T& y() const { return c_[1]; } const T& z() const { return c_[2]; } const T& w() const { return c_[3]; }
  Vector3 vec() const { return Vector3(c_[0], c_[1], c_[2]); }
  Matrix<T, 4, 1>& coeffs() { return c_; }
  const Matrix<T, 4, 1>& coeffs() const { return c_; }
  Quaternion operator*(const Quaternion& b) const {  // quat_product<Arch::Target, ...>
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
  Vector3 _transformVector(const Vector3& v) const {  // QuaternionBase::_transformVector
    Vector3 uv = this->vec().cross(v);
    uv += uv;
    return v + this->w() * uv + this->vec().cross(uv);
  }
  Vector3 operator*(const Vector3& v) const { return _transformVector(v); }
  template <typename M> Vector3 operator*(const BlockRef<M, 3, 1>& v) const { return _transformVector(v.eval()); }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  T squaredNorm() const { return c_.squaredNorm(); }
  T norm() const { return c_.norm(); }
  void normalize() { c_ /= c_.norm(); }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion inverse() const {  // conjugate / squaredNorm
    const T n2 = this->squaredNorm();
    if (n2 > T(0)) return Quaternion(Matrix<T, 4, 1>(conjugate().coeffs() / n2));
    return Quaternion(Matrix<T, 4, 1>::Zero());
  }
  T dot(const Quaternion& o) const { return c_.dot(o.c_); }
  bool isApprox(const Quaternion& o, T prec = T(1e-12)) const { return c_.isApprox(o.c_, prec); }
  Matrix3 toRotationMatrix() const {  // QuaternionBase::toRotationMatrix
    Matrix3 res;
    const T tx = T(2) * x(), ty = T(2) * y(), tz = T(2) * z();
    const T twx = tx * w(), twy = ty * w(), twz = tz * w();
    const T txx = tx * x(), txy = ty * x(), txz = tz * x();
    const T tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
    return res;
  }
  Matrix3 matrix() const { return toRotationMatrix(); }
  template <typename U> Quaternion<U> cast() const { return Quaternion<U>(U(w()), U(x()), U(y()), U(z())); }
 private:
  Matrix<T, 4, 1> c_;
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <typename T> class AngleAxis {
 public:
  typedef T Scalar;
  typedef Matrix<T, 3, 1> Vector3;
  AngleAxis() : angle_(T(0)), axis_(Vector3(T(1), T(0), T(0))) {}
  AngleAxis(T angle, const Vector3& axis) : angle_(angle), axis_(axis) {}
  static AngleAxis Identity() { return AngleAxis(); }
  explicit AngleAxis(const Quaternion<T>& q) {  // AngleAxis::operator=(QuaternionBase)
    T n = q.vec().norm();
    if (n < std::numeric_limits<T>::epsilon()) n = q.vec().norm();
    if (n != T(0)) { angle_ = T(2) * std::atan2(n, std::abs(q.w())); if (q.w() < T(0)) n = -n; axis_ = q.vec() / n; }
    else { angle_ = T(0); axis_ = Vector3(T(1), T(0), T(0)); }
  }
  explicit AngleAxis(const Matrix<T, 3, 3>& m) { *this = AngleAxis(Quaternion<T>(m)); }
  T& angle() { return angle_; } const T& angle() const { return angle_; }
  Vector3& axis() { return axis_; } const Vector3& axis() const { return axis_; }
  AngleAxis inverse() const { return AngleAxis(-angle_, axis_); }
  Matrix<T, 3, 3> toRotationMatrix() const { return Quaternion<T>(*this).toRotationMatrix(); }
  Vector3 operator*(const Vector3& v) const { return toRotationMatrix() * v; }
  bool isApprox(const AngleAxis& o, T prec = T(1e-12)) const { return axis_.isApprox(o.axis_, prec) && std::abs(angle_ - o.angle_) <= prec; }
 private:
  T angle_; Vector3 axis_;
};
template <typename T> Quaternion<T>::Quaternion(const AngleAxis<T>& aa) {  // QuaternionBase::operator=(AngleAxis)
  const T ha = T(0.5) * aa.angle();
  this->w() = std::cos(ha);
  const Vector3 v = std::sin(ha) * aa.axis();
  c_[0] = v[0]; c_[1] = v[1]; c_[2] = v[2];
}
typedef AngleAxis<double> AngleAxisd;
template <typename T, int N> class DiagonalMatrix {
 public:
  DiagonalMatrix(T a, T b) { static_assert(N == 2, "2 entries"); d_[0] = a; d_[1] = b; }
  DiagonalMatrix(T a, T b, T c) { static_assert(N == 3, "3 entries"); d_[0] = a; d_[1] = b; d_[2] = c; }
  template <int K> Matrix<T, N, K> operator*(const Matrix<T, N, K>& m) const {  // scales row i by d_i
    Matrix<T, N, K> r; for (int j = 0; j < K; ++j) for (int i = 0; i < N; ++i) r(i, j) = d_[i] * m(i, j); return r;
  }
 private:
  T d_[N];
};

typedef Matrix<float, 2, 1> Vector2f;  typedef Matrix<double, 2, 1> Vector2d;  typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<float, 3, 1> Vector3f;  typedef Matrix<double, 3, 1> Vector3d;  typedef Matrix<int, 3, 1> Vector3i;
typedef Matrix<float, 4, 1> Vector4f;  typedef Matrix<double, 4, 1> Vector4d;  typedef Matrix<int, 4, 1> Vector4i;
typedef Matrix<float, 2, 2> Matrix2f;  typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<float, 3, 3> Matrix3f;  typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 4, 4> Matrix4f;  typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<double, 6, 1> Vector6d;
typedef Matrix<double, 6, 6> Matrix6d;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, 2, Dynamic> Matrix2Xd;
typedef Matrix<double, 3, Dynamic> Matrix3Xd;
typedef Matrix<double, 2, 3> Matrix23d;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;

}  // namespace Eigen
