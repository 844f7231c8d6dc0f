/*
 * eigen_shim.h — the slice of Eigen 3.3 that the reference's hot-path sources touch.
 * TEST INFRASTRUCTURE ONLY (oracle/_ref): lets the UNMODIFIED reference sources under
 * /root/reference compile in a container that has no Eigen.  Not used by the product.
 *
 * What is restated here (from Eigen 3.3.7 as shipped with Ubuntu 20.04, x86-64, SSE2, no FMA —
 * the reference's CI build, .github/workflows/build_test.yml:63) is ONLY the floating-point
 * operation order of the handful of Eigen expressions the reference evaluates:
 *
 *  - Vector3d redux (squaredNorm): LinearVectorizedTraversal + CompleteUnrolling with Packet2d
 *    (EIGEN_UNALIGNED_VECTORIZE=1): predux(x², y²) then + z²  →  (x²+y²)+z²      [Core/Redux.h]
 *  - Vector4d redux (Quaternion::coeffs().squaredNorm(), storage order x,y,z,w):
 *    (x²,y²)+(z²,w²) then predux  →  (x²+z²)+(y²+w²)                              [Core/Redux.h]
 *  - normalized(): n / sqrt(squaredNorm) per component (division)                 [Core/Dot.h]
 *  - Quaterniond * Vector3d: uv = vec×v; uv += uv; (v + w*uv) + vec×uv   [Geometry/Quaternion.h]
 *  - Quaterniond * Quaterniond (double, SSE2 specialisation):
 *      x = (aw*bx + ay*bz) - (az*by - ax*bw)      y = (aw*by + ay*bw) + (az*bx - ax*bz)
 *      z = (aw*bz - ay*bx) + (az*bw + ax*by)      w = (aw*bw - ay*by) - (az*bz + ax*bx)
 *                                                             [Geometry/arch/Geometry_SSE.h]
 *  - Quaterniond = AngleAxisd: ha = 0.5*angle; w = cos(ha); vec = sin(ha)*axis
 *  - cross(): a.y*b.z - a.z*b.y, ...                                  [Geometry/OrthoMethods.h]
 *  - coefficient-wise a + s*b etc.: one multiply, one add per component, no contraction
 *    (compile with -ffp-contract=off).
 * Eigen itself is not in this container, so these are recalled, not diffed — see DESIGN.md §2.
 */
#ifndef A3D_ORACLE_EIGEN_SHIM_H_
#define A3D_ORACLE_EIGEN_SHIM_H_

#include <cmath>
#include <cstddef>
#include <memory>
#include <ostream>
#include <sstream>  // Eigen/Core pulls these in transitively; trajectory.h:110 relies on it
#include <iostream>
#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 7

namespace Eigen {

template <class T>
class aligned_allocator : public std::allocator<T> {
 public:
  template <class U>
  struct rebind {
    typedef aligned_allocator<U> other;
  };
  aligned_allocator() = default;
  aligned_allocator(const aligned_allocator&) = default;
  template <class U>
  aligned_allocator(const aligned_allocator<U>&) {}  // NOLINT
};

template <typename S, int R, int C>
class Matrix;

template <typename S>
class Quaternion;
template <typename S>
class AngleAxis;

// ---- fixed 3-vector ------------------------------------------------------------------------
template <typename S>
class Matrix<S, 3, 1> {
 public:
  typedef S Scalar;
  Matrix() : v_{S(0), S(0), S(0)} {}  // (Eigen leaves this uninitialised)
  template <typename A, typename B, typename D>
  Matrix(const A& x, const B& y, const D& z) : v_{static_cast<S>(x), static_cast<S>(y), static_cast<S>(z)} {}

  static Matrix Zero() { return Matrix(S(0), S(0), S(0)); }
  static Matrix UnitX() { return Matrix(S(1), S(0), S(0)); }
  static Matrix UnitY() { return Matrix(S(0), S(1), S(0)); }
  static Matrix UnitZ() { return Matrix(S(0), S(0), S(1)); }

  S& x() { return v_[0]; }
  S& y() { return v_[1]; }
  S& z() { return v_[2]; }
  const S& x() const { return v_[0]; }
  const S& y() const { return v_[1]; }
  const S& z() const { return v_[2]; }
  S& operator[](int i) { return v_[i]; }
  const S& operator[](int i) const { return v_[i]; }
  S& operator()(int i) { return v_[i]; }
  const S& operator()(int i) const { return v_[i]; }
  const S* data() const { return v_; }
  S* data() { return v_; }
  int rows() const { return 3; }
  int cols() const { return 1; }
  int size() const { return 3; }

  Matrix operator+(const Matrix& o) const { return Matrix(v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]); }
  Matrix operator-(const Matrix& o) const { return Matrix(v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]); }
  Matrix operator-() const { return Matrix(-v_[0], -v_[1], -v_[2]); }
  Matrix operator*(S s) const { return Matrix(v_[0] * s, v_[1] * s, v_[2] * s); }
  Matrix operator/(S s) const { return Matrix(v_[0] / s, v_[1] / s, v_[2] / s); }
  Matrix& operator+=(const Matrix& o) {
    v_[0] += o.v_[0];
    v_[1] += o.v_[1];
    v_[2] += o.v_[2];
    return *this;
  }
  Matrix& operator-=(const Matrix& o) {
    v_[0] -= o.v_[0];
    v_[1] -= o.v_[1];
    v_[2] -= o.v_[2];
    return *this;
  }
  Matrix& operator*=(S s) {
    v_[0] *= s;
    v_[1] *= s;
    v_[2] *= s;
    return *this;
  }
  // Eigen: (a.cwiseEqual(b)).all()
  bool operator==(const Matrix& o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2]; }
  bool operator!=(const Matrix& o) const { return !(*this == o); }

  S dot(const Matrix& o) const { return (v_[0] * o.v_[0] + v_[1] * o.v_[1]) + v_[2] * o.v_[2]; }
  Matrix cross(const Matrix& o) const {
    return Matrix(v_[1] * o.v_[2] - v_[2] * o.v_[1], v_[2] * o.v_[0] - v_[0] * o.v_[2],
                  v_[0] * o.v_[1] - v_[1] * o.v_[0]);
  }
  S squaredNorm() const { return (v_[0] * v_[0] + v_[1] * v_[1]) + v_[2] * v_[2]; }
  S norm() const { return std::sqrt(squaredNorm()); }
  Matrix normalized() const {
    const S z = squaredNorm();
    if (z > S(0)) {
      const S n = std::sqrt(z);
      return Matrix(v_[0] / n, v_[1] / n, v_[2] / n);
    }
    return *this;
  }
  void normalize() { *this = normalized(); }
  template <typename T>
  Matrix<T, 3, 1> cast() const {
    return Matrix<T, 3, 1>(static_cast<T>(v_[0]), static_cast<T>(v_[1]), static_cast<T>(v_[2]));
  }
  void setZero() { v_[0] = v_[1] = v_[2] = S(0); }

  // printable row view, only for EigenTrajectoryPoint::toString (trajectory.h:109-121)
  struct TransposeView {
    const Matrix& m;
  };
  TransposeView transpose() const { return TransposeView{*this}; }

 private:
  S v_[3];
};

template <typename S>
inline Matrix<S, 3, 1> operator*(S s, const Matrix<S, 3, 1>& v) {
  return Matrix<S, 3, 1>(s * v.x(), s * v.y(), s * v.z());
}
template <typename S>
inline std::ostream& operator<<(std::ostream& os, const typename Matrix<S, 3, 1>::TransposeView& t) {
  return os << t.m.x() << " " << t.m.y() << " " << t.m.z();
}
inline std::ostream& operator<<(std::ostream& os, const Matrix<double, 3, 1>::TransposeView& t) {
  return os << t.m.x() << " " << t.m.y() << " " << t.m.z();
}
template <typename S>
inline std::ostream& operator<<(std::ostream& os, const Matrix<S, 3, 1>& v) {
  return os << v.x() << "\n" << v.y() << "\n" << v.z();
}

typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<int, 3, 1> Vector3i;

// ---- 3x3 matrix: only what trajectory.h's Affine3d operator* needs to type-check -----------
template <typename S>
class Matrix<S, 3, 3> {
 public:
  Matrix() : m_{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  S& operator()(int r, int c) { return m_[3 * r + c]; }
  const S& operator()(int r, int c) const { return m_[3 * r + c]; }
  Matrix<S, 3, 1> operator*(const Matrix<S, 3, 1>& v) const {
    return Matrix<S, 3, 1>((m_[0] * v.x() + m_[1] * v.y()) + m_[2] * v.z(), (m_[3] * v.x() + m_[4] * v.y()) + m_[5] * v.z(),
                           (m_[6] * v.x() + m_[7] * v.y()) + m_[8] * v.z());
  }
  Matrix operator*(const Matrix& o) const {
    Matrix r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r(i, j) = ((*this)(i, 0) * o(0, j) + (*this)(i, 1) * o(1, j)) + (*this)(i, 2) * o(2, j);
    return r;
  }

 private:
  S m_[9];
};
typedef Matrix<double, 3, 3> Matrix3d;

// ---- AngleAxis --------------------------------------------------------------------------------
template <typename S>
class AngleAxis {
 public:
  AngleAxis() : angle_(0), axis_(S(1), S(0), S(0)) {}
  AngleAxis(S angle, const Matrix<S, 3, 1>& axis) : angle_(angle), axis_(axis) {}
  S angle() const { return angle_; }
  const Matrix<S, 3, 1>& axis() const { return axis_; }

 private:
  S angle_;
  Matrix<S, 3, 1> axis_;
};
typedef AngleAxis<double> AngleAxisd;

// ---- Quaternion -------------------------------------------------------------------------------
template <typename S>
class Quaternion {
 public:
  Quaternion() : x_(0), y_(0), z_(0), w_(1) {}  // (Eigen leaves this uninitialised)
  Quaternion(S w, S x, S y, S z) : x_(x), y_(y), z_(z), w_(w) {}
  Quaternion(const AngleAxis<S>& aa) { *this = aa; }  // NOLINT (implicit like Eigen's)
  explicit Quaternion(const Matrix<S, 3, 3>& m) { *this = m; }
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }

  // Geometry/Quaternion.h: operator=(AngleAxis)
  Quaternion& operator=(const AngleAxis<S>& aa) {
    const S ha = S(0.5) * aa.angle();
    w_ = std::cos(ha);
    const S s = std::sin(ha);
    x_ = s * aa.axis().x();
    y_ = s * aa.axis().y();
    z_ = s * aa.axis().z();
    return *this;
  }
  // Rotation matrix → quaternion (Shepperd); only reachable through trajectory.h's unused
  // Affine3d operator*.
  Quaternion& operator=(const Matrix<S, 3, 3>& m) {
    S t = m(0, 0) + m(1, 1) + m(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      w_ = S(0.5) * t;
      t = S(0.5) / t;
      x_ = (m(2, 1) - m(1, 2)) * t;
      y_ = (m(0, 2) - m(2, 0)) * t;
      z_ = (m(1, 0) - m(0, 1)) * t;
    } else {
      int i = 0;
      if (m(1, 1) > m(0, 0)) i = 1;
      if (m(2, 2) > m(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m(i, i) - m(j, j) - m(k, k) + S(1.0));
      S q[3];
      q[i] = S(0.5) * t;
      t = S(0.5) / t;
      w_ = (m(k, j) - m(j, k)) * t;
      q[j] = (m(j, i) + m(i, j)) * t;
      q[k] = (m(k, i) + m(i, k)) * t;
      x_ = q[0];
      y_ = q[1];
      z_ = q[2];
    }
    return *this;
  }

  S& x() { return x_; }
  S& y() { return y_; }
  S& z() { return z_; }
  S& w() { return w_; }
  const S& x() const { return x_; }
  const S& y() const { return y_; }
  const S& z() const { return z_; }
  const S& w() const { return w_; }
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(x_, y_, z_); }

  // coeffs() is a Vector4 stored x,y,z,w; redux with Packet2d: (x²+z²)+(y²+w²)
  S squaredNorm() const { return (x_ * x_ + z_ * z_) + (y_ * y_ + w_ * w_); }
  S norm() const { return std::sqrt(squaredNorm()); }
  Quaternion normalized() const {
    const S z = squaredNorm();
    if (z > S(0)) {
      const S n = std::sqrt(z);
      return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n);
    }
    return *this;
  }
  void normalize() { *this = normalized(); }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
  }

  // Geometry/arch/Geometry_SSE.h quat_product<SSE, ..., double>
  Quaternion operator*(const Quaternion& b) const {
    const Quaternion& a = *this;
    const S t1x = a.w_ * b.x_ + a.y_ * b.z_, t1y = a.w_ * b.y_ + a.y_ * b.w_;
    const S t2x = a.z_ * b.x_ - a.x_ * b.z_, t2y = a.z_ * b.y_ - a.x_ * b.w_;
    const S rx = t1x - t2y, ry = t1y + t2x;
    const S u1x = a.w_ * b.z_ - a.y_ * b.x_, u1y = a.w_ * b.w_ - a.y_ * b.y_;
    const S u2x = a.z_ * b.z_ + a.x_ * b.x_, u2y = a.z_ * b.w_ + a.x_ * b.y_;
    const S rz = u1x + u2y, rw = u1y - u2x;
    return Quaternion(rw, rx, ry, rz);
  }
  Quaternion& operator*=(const Quaternion& b) { return *this = *this * b; }

  // Geometry/Quaternion.h _transformVector
  Matrix<S, 3, 1> operator*(const Matrix<S, 3, 1>& v) const {
    const Matrix<S, 3, 1> qv = vec();
    Matrix<S, 3, 1> uv = qv.cross(v);
    uv += uv;
    const Matrix<S, 3, 1> c = qv.cross(uv);
    return Matrix<S, 3, 1>((v.x() + w_ * uv.x()) + c.x(), (v.y() + w_ * uv.y()) + c.y(), (v.z() + w_ * uv.z()) + c.z());
  }

  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> r;
    const S tx = S(2) * x_, ty = S(2) * y_, tz = S(2) * z_;
    const S twx = tx * w_, twy = ty * w_, twz = tz * w_;
    const S txx = tx * x_, txy = ty * x_, txz = tz * x_;
    const S tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    r(0, 0) = S(1) - (tyy + tzz);
    r(0, 1) = txy - twz;
    r(0, 2) = txz + twy;
    r(1, 0) = txy + twz;
    r(1, 1) = S(1) - (txx + tzz);
    r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy;
    r(2, 1) = tyz + twx;
    r(2, 2) = S(1) - (txx + tyy);
    return r;
  }

 private:
  S x_, y_, z_, w_;  // Eigen's storage order
};
typedef Quaternion<double> Quaterniond;

// rotation-matrix * quaternion → rotation matrix (RotationBase friend operator*)
template <typename S>
inline Matrix<S, 3, 3> operator*(const Matrix<S, 3, 3>& m, const Quaternion<S>& q) {
  return m * q.toRotationMatrix();
}

// ---- Affine3d: only so that trajectory.h:131-146 (never called on this path) compiles -------
class Affine3d {
 public:
  Affine3d() {}
  const Matrix3d& rotation() const { return r_; }
  Vector3d operator*(const Vector3d& v) const { return r_ * v + t_; }

 private:
  Matrix3d r_;
  Vector3d t_;
};

// ---- ArrayXXi: IterativeRayCaster::ray_table_ (iterative_ray_caster.h:44) -------------------
class ArrayXXi {
 public:
  ArrayXXi() : rows_(0), cols_(0) {}
  static ArrayXXi Zero(int rows, int cols) {
    ArrayXXi a;
    a.rows_ = rows;
    a.cols_ = cols;
    a.d_.assign(static_cast<size_t>(rows) * cols, 0);
    return a;
  }
  int& operator()(int i, int j) { return d_[static_cast<size_t>(j) * rows_ + i]; }  // column-major
  const int& operator()(int i, int j) const { return d_[static_cast<size_t>(j) * rows_ + i]; }
  int rows() const { return rows_; }
  int cols() const { return cols_; }

 private:
  int rows_, cols_;
  std::vector<int> d_;
};

}  // namespace Eigen
#endif  // A3D_ORACLE_EIGEN_SHIM_H_
