/*
 * ORACLE support (test infrastructure, NOT product code): the few fixed-size Eigen types the reference's CPU op kernels
 * use (warp2d.cc, depthtoflow.cc, rotation_format.h, helper.h), so that those sources compile unmodified for oracle/_ref.
 *
 * Eigen is THIRD PARTY and absent from /root/reference and from this image (the reference uses the copy bundled with the
 * TensorFlow 1.4.0 pip package, Eigen 3.3.x).  Everything here is eager evaluation on plain arrays (no expression
 * templates); where Eigen's evaluation ORDER decides the floating-point result, Eigen 3.3's published behaviour is
 * restated and marked [order]:
 *   - Matrix<T,4,1>::dot            4-lane packet product + predux: (p0 + p2) + (p1 + p3)
 *   - Matrix<T,3,3> * Matrix<T,3,1> coefficient-based product, left to right: (r0*v0 + r1*v1) + r2*v2
 *   - Matrix<T,3,1>::norm           sqrt of the left-to-right sum of squares
 *   - Quaternion::normalize         coeffs (x,y,z,w) /= sqrt(((x^2 + y^2) + z^2) + w^2)
 *   - Quaternion / AngleAxis::toRotationMatrix   the formulas of Eigen/src/Geometry/{Quaternion,AngleAxis}.h
 *   - Matrix<T,3,3>::inverse        cofactors times 1/det, Eigen/src/LU/InverseImpl.h (depthtonormals.cc)
 *   - Matrix<T,3,1>::cross / normalize   Eigen/src/Geometry/OrthoMethods.h, Eigen/src/Core/Dot.h
 * These are the same orders oracle/geometry_ops_impl.h uses, so a mismatch between _ref and the C oracle points at the
 * OP logic (index math, branches, loops), which is the part that comes from the reference's own source.
 */
#ifndef ORACLE_EIGEN_STUB_H
#define ORACLE_EIGEN_STUB_H
#include <cmath>
#include <cstddef>

#define EIGEN_STATIC_ASSERT_VECTOR_SPECIFIC_SIZE(TYPE, SIZE) \
  static_assert(TYPE::SizeAtCompileTime == SIZE && (TYPE::RowsAtCompileTime == 1 || TYPE::ColsAtCompileTime == 1), "vector size");
#define EIGEN_STATIC_ASSERT_MATRIX_SPECIFIC_SIZE(TYPE, ROWS, COLS) \
  static_assert(TYPE::RowsAtCompileTime == ROWS && TYPE::ColsAtCompileTime == COLS, "matrix size");

namespace Eigen {

template <class T, int R, int C> class Matrix;
template <class M> class Map;
template <class D> struct scalar_of;
template <class T, int R, int C> struct scalar_of<Matrix<T, R, C> > { typedef T type; };
template <class T, int R, int C> struct scalar_of<Map<const Matrix<T, R, C> > > { typedef T type; };

// CRTP base, as in Eigen: the reference passes `const Eigen::MatrixBase<Derived>&` around
template <class Derived>
class MatrixBase {
 public:
  const Derived& derived() const { return *static_cast<const Derived*>(this); }
  Derived& derived() { return *static_cast<Derived*>(this); }
  // coefficient access through the base, as the reference uses it on `const MatrixBase<VEC2T>&` parameters
  typedef typename scalar_of<Derived>::type BaseScalar;
  const BaseScalar& x() const { return derived()(0); }
  const BaseScalar& y() const { return derived()(1); }
  const BaseScalar& z() const { return derived()(2); }
  int rows() const { return derived().rows(); }
  int cols() const { return derived().cols(); }
  const BaseScalar& operator()(int i, int j) const { return derived()(i, j); }
};

template <class D> struct traits;

#define ORACLE_EIGEN_COMMON(Derived, T, R, C)                                                              \
  typedef T Scalar;                                                                                       \
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = R * C };                       \
  int rows() const { return R; }                                                                          \
  int cols() const { return C; }                                                                          \
  const T& operator()(int i, int j) const { return d_[j * R + i]; }   /* column major, like Eigen */     \
  T& operator()(int i, int j) { return d_[j * R + i]; }                                                   \
  const T& operator()(int i) const { return d_[i]; }                                                      \
  T& operator()(int i) { return d_[i]; }                                                                  \
  const T& operator[](int i) const { return d_[i]; }                                                      \
  T& operator[](int i) { return d_[i]; }                                                                  \
  const T& x() const { return d_[0]; }                                                                    \
  T& x() { return d_[0]; }                                                                                \
  const T& y() const { return d_[1]; }                                                                    \
  T& y() { return d_[1]; }                                                                                \
  const T& z() const { return d_[2]; }                                                                    \
  T& z() { return d_[2]; }                                                                                \
  const T& w() const { return d_[3]; }                                                                    \
  T& w() { return d_[3]; }

template <class T, int R, int C>
class Matrix : public MatrixBase<Matrix<T, R, C> > {
 public:
  ORACLE_EIGEN_COMMON(Matrix, T, R, C)
  Matrix() {}
  Matrix(const T& a, const T& b) { static_assert(R * C == 2, "size"); d_[0] = a; d_[1] = b; }
  Matrix(const T& a, const T& b, const T& c) { static_assert(R * C == 3, "size"); d_[0] = a; d_[1] = b; d_[2] = c; }
  Matrix(const T& a, const T& b, const T& c, const T& d) { static_assert(R * C == 4, "size"); d_[0] = a; d_[1] = b; d_[2] = c; d_[3] = d; }
  template <class D>
  Matrix(const MatrixBase<D>& o) { assign(o.derived()); }
  template <class D>
  Matrix& operator=(const MatrixBase<D>& o) { assign(o.derived()); return *this; }

  template <class U>
  Matrix<U, R, C> cast() const {   // static_cast per coefficient: (int)float is cvttss2si on x86-64
    Matrix<U, R, C> r;
    for (int i = 0; i < R * C; ++i) r(i) = static_cast<U>(d_[i]);
    return r;
  }
  template <class D>
  T dot(const MatrixBase<D>& o) const {
    const D& b = o.derived();
    if (R * C == 4) return (d_[0] * b(0) + d_[2] * b(2)) + (d_[1] * b(1) + d_[3] * b(3));   // [order] predux of a 4-lane packet
    T s = d_[0] * b(0);
    for (int i = 1; i < R * C; ++i) s = s + d_[i] * b(i);
    return s;
  }
  T squaredNorm() const { T s = d_[0] * d_[0]; for (int i = 1; i < R * C; ++i) s = s + d_[i] * d_[i]; return s; }   // [order]
  T norm() const { return std::sqrt(squaredNorm()); }
  Matrix& operator/=(const T& s) { for (int i = 0; i < R * C; ++i) d_[i] = d_[i] / s; return *this; }
  Matrix& operator*=(const T& s) { for (int i = 0; i < R * C; ++i) d_[i] = d_[i] * s; return *this; }
  template <class D>
  Matrix cwiseProduct(const MatrixBase<D>& o) const { Matrix r; for (int i = 0; i < R * C; ++i) r(i) = d_[i] * o.derived()(i); return r; }
  Matrix<T, R + 1, 1> homogeneous() const {
    static_assert(C == 1, "vector");
    Matrix<T, R + 1, 1> r;
    for (int i = 0; i < R; ++i) r(i) = d_[i];
    r(R) = T(1);
    return r;
  }
  template <int N>
  Matrix<T, N, C> topRows() const { Matrix<T, N, C> r; for (int j = 0; j < C; ++j) for (int i = 0; i < N; ++i) r(i, j) = (*this)(i, j); return r; }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> r; for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r(j, i) = (*this)(i, j); return r; }
  void setIdentity() { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) (*this)(i, j) = (i == j) ? T(1) : T(0); }
  static Matrix Identity() { Matrix r; r.setIdentity(); return r; }
  // 3x3 inverse, Eigen/src/LU/InverseImpl.h (compute_inverse<MatrixType, ResultType, 3> + compute_inverse_size3_helper) [order]:
  // cofactor_3x3<i,j>(m) = m(i1,j1)*m(i2,j2) - m(i1,j2)*m(i2,j1) with i1=(i+1)%3, i2=(i+2)%3, j1=(j+1)%3, j2=(j+2)%3;
  // det = (c00*m00 + c10*m10) + c20*m20; result(0,j) = cofactor<j,0> * invdet, result(1,j) = cofactor<j,1> * invdet, ...
  Matrix inverse() const {
    static_assert(R == 3 && C == 3, "only the fixed 3x3 inverse of depthtonormals.cc is restated");
    const Matrix& m = *this;
    auto cof = [&](int i, int j) { const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                                   return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1); };
    const T c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const T det = (c0 * m(0, 0) + c1 * m(1, 0)) + c2 * m(2, 0);
    const T invdet = T(1) / det;
    Matrix r;
    r(0, 0) = c0 * invdet; r(0, 1) = c1 * invdet; r(0, 2) = c2 * invdet;
    r(1, 0) = cof(0, 1) * invdet; r(1, 1) = cof(1, 1) * invdet; r(1, 2) = cof(2, 1) * invdet;
    r(2, 0) = cof(0, 2) * invdet; r(2, 1) = cof(1, 2) * invdet; r(2, 2) = cof(2, 2) * invdet;
    return r;
  }
  // comma initialiser (`v << a, b, c;`): coefficients in storage order for vectors, row by row for matrices, like Eigen
  struct CommaInit {
    Matrix& m; int k;
    CommaInit& operator,(const T& v) { m(k / C, k % C) = v; ++k; return *this; }
  };
  CommaInit operator<<(const T& v) { (*this)(0, 0) = v; return CommaInit{*this, 1}; }
  template <class D>
  Matrix cross(const MatrixBase<D>& o) const {   // MatrixBase::cross, Eigen/src/Geometry/OrthoMethods.h
    static_assert(R * C == 3, "vector of size 3");
    const D& b = o.derived();
    return Matrix(d_[1] * b(2) - d_[2] * b(1), d_[2] * b(0) - d_[0] * b(2), d_[0] * b(1) - d_[1] * b(0));
  }
  void normalize() {   // MatrixBase::normalize (Eigen 3.3): z = squaredNorm(); if (z > 0) derived() /= sqrt(z)
    const T z = squaredNorm();
    if (z > T(0)) *this /= std::sqrt(z);
  }
  void setZero() { for (int i = 0; i < R * C; ++i) d_[i] = T(0); }
  const T* data() const { return d_; }
  T* data() { return d_; }

 private:
  template <class D>
  void assign(const D& o) {
    static_assert((int)D::RowsAtCompileTime == R && (int)D::ColsAtCompileTime == C, "size");
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) (*this)(i, j) = o(i, j);
  }
  T d_[R * C];
};

// Map<const Matrix<T,R,C>>: a view in Eigen; the kernels only read through it, so a copy is equivalent
template <class T, int R, int C>
class Map<const Matrix<T, R, C> > : public MatrixBase<Map<const Matrix<T, R, C> > > {
 public:
  ORACLE_EIGEN_COMMON(Map, T, R, C)
  explicit Map(const T* p) { for (int i = 0; i < R * C; ++i) d_[i] = p[i]; }
  Matrix<T, C, R> transpose() const { Matrix<T, C, R> r; for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r(j, i) = (*this)(i, j); return r; }
 private:
  T d_[R * C];
};

// ---- free operators on anything derived from MatrixBase -----------------------------------------------------------------
template <class A, class B>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
  for (int j = 0; j < A::ColsAtCompileTime; ++j) for (int i = 0; i < A::RowsAtCompileTime; ++i) r(i, j) = a.derived()(i, j) + b.derived()(i, j);
  return r;
}
template <class A, class B>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
  for (int j = 0; j < A::ColsAtCompileTime; ++j) for (int i = 0; i < A::RowsAtCompileTime; ++i) r(i, j) = a.derived()(i, j) - b.derived()(i, j);
  return r;
}
template <class A>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator*(const typename A::Scalar& s, const MatrixBase<A>& a) {
  Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
  for (int j = 0; j < A::ColsAtCompileTime; ++j) for (int i = 0; i < A::RowsAtCompileTime; ++i) r(i, j) = s * a.derived()(i, j);
  return r;
}
template <class A>
Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> operator*(const MatrixBase<A>& a, const typename A::Scalar& s) {
  Matrix<typename A::Scalar, A::RowsAtCompileTime, A::ColsAtCompileTime> r;
  for (int j = 0; j < A::ColsAtCompileTime; ++j) for (int i = 0; i < A::RowsAtCompileTime; ++i) r(i, j) = a.derived()(i, j) * s;
  return r;
}
// matrix product, coefficient based, left to right [order]
template <class A, class B>
Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  static_assert((int)A::ColsAtCompileTime == (int)B::RowsAtCompileTime, "inner size");
  Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> r;
  for (int j = 0; j < B::ColsAtCompileTime; ++j)
    for (int i = 0; i < A::RowsAtCompileTime; ++i) {
      typename A::Scalar s = a.derived()(i, 0) * b.derived()(0, j);
      for (int k = 1; k < A::ColsAtCompileTime; ++k) s = s + a.derived()(i, k) * b.derived()(k, j);
      r(i, j) = s;
    }
  return r;
}

// ---- Geometry -----------------------------------------------------------------------------------------------------------
template <class T>
class Quaternion {
 public:
  Quaternion() {}
  T& w() { return w_; }
  T& x() { return x_; }
  T& y() { return y_; }
  T& z() { return z_; }
  void normalize() {   // coeffs() /= norm(), coefficient order (x, y, z, w) [order]
    const T n = std::sqrt(((x_ * x_ + y_ * y_) + z_ * z_) + w_ * w_);
    x_ = x_ / n; y_ = y_ / n; z_ = z_ / n; w_ = w_ / n;
  }
  Matrix<T, 3, 3> toRotationMatrix() const {   // Eigen/src/Geometry/Quaternion.h, QuaternionBase::toRotationMatrix
    Matrix<T, 3, 3> res;
    const T tx = T(2) * x_, ty = T(2) * y_, tz = T(2) * z_;
    const T twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const T txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const T tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    res(0, 0) = T(1) - (tyy + tzz); res(0, 1) = txy - twz; res(0, 2) = txz + twy;
    res(1, 0) = txy + twz; res(1, 1) = T(1) - (txx + tzz); res(1, 2) = tyz - twx;
    res(2, 0) = txz - twy; res(2, 1) = tyz + twx; res(2, 2) = T(1) - (txx + tyy);
    return res;
  }
 private:
  T x_, y_, z_, w_;
};

template <class T>
class AngleAxis {
 public:
  template <class D>
  AngleAxis(const T& angle, const MatrixBase<D>& axis) : angle_(angle), axis_(axis) {}
  Matrix<T, 3, 3> toRotationMatrix() const {   // Eigen/src/Geometry/AngleAxis.h, AngleAxis::toRotationMatrix
    Matrix<T, 3, 3> res;
    const T sin_axis_x = std::sin(angle_) * axis_.x(), sin_axis_y = std::sin(angle_) * axis_.y(), sin_axis_z = std::sin(angle_) * axis_.z();
    const T c = std::cos(angle_);
    const T cos1_axis_x = (T(1) - c) * axis_.x(), cos1_axis_y = (T(1) - c) * axis_.y(), cos1_axis_z = (T(1) - c) * axis_.z();
    T tmp;
    tmp = cos1_axis_x * axis_.y();
    res(0, 1) = tmp - sin_axis_z;
    res(1, 0) = tmp + sin_axis_z;
    tmp = cos1_axis_x * axis_.z();
    res(0, 2) = tmp + sin_axis_y;
    res(2, 0) = tmp - sin_axis_y;
    tmp = cos1_axis_y * axis_.z();
    res(1, 2) = tmp - sin_axis_x;
    res(2, 1) = tmp + sin_axis_x;
    res(0, 0) = cos1_axis_x * axis_.x() + c;
    res(1, 1) = cos1_axis_y * axis_.y() + c;
    res(2, 2) = cos1_axis_z * axis_.z() + c;
    return res;
  }
 private:
  T angle_;
  Matrix<T, 3, 1> axis_;
};

}  // namespace Eigen
#endif
