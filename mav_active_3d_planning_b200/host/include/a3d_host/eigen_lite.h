// eigen_lite.h — the few Eigen types the view-gain module API exposes (Vector3d, Quaterniond,
// AngleAxisd), for machines without Eigen.  If real Eigen is on the include path it is used instead.
#pragma once
#if defined(A3D_HOST_USE_EIGEN) || (defined(__has_include) && __has_include(<Eigen/Geometry>))
#include <Eigen/Core>
#include <Eigen/Geometry>
#else
#include <cmath>
namespace Eigen {

class Vector3d {
 public:
  Vector3d() : v_{0.0, 0.0, 0.0} {}
  Vector3d(double x, double y, double z) : v_{x, y, z} {}
  static Vector3d Zero() { return Vector3d(); }
  static Vector3d UnitX() { return Vector3d(1, 0, 0); }
  static Vector3d UnitY() { return Vector3d(0, 1, 0); }
  static Vector3d UnitZ() { return Vector3d(0, 0, 1); }
  double& x() { return v_[0]; }
  double& y() { return v_[1]; }
  double& z() { return v_[2]; }
  double x() const { return v_[0]; }
  double y() const { return v_[1]; }
  double z() const { return v_[2]; }
  double& operator[](int i) { return v_[i]; }
  double operator[](int i) const { return v_[i]; }
  const double* data() const { return v_; }
  Vector3d operator+(const Vector3d& o) const { return {v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]}; }
  Vector3d operator-(const Vector3d& o) const { return {v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]}; }
  Vector3d operator*(double s) const { return {v_[0] * s, v_[1] * s, v_[2] * s}; }
  Vector3d& operator+=(const Vector3d& o) { return *this = *this + o; }
  bool operator==(const Vector3d& o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2]; }
  bool operator!=(const Vector3d& o) const { return !(*this == o); }
  double dot(const Vector3d& o) const { return v_[0] * o.v_[0] + v_[1] * o.v_[1] + v_[2] * o.v_[2]; }
  Vector3d cross(const Vector3d& o) const {
    return {v_[1] * o.v_[2] - v_[2] * o.v_[1], v_[2] * o.v_[0] - v_[0] * o.v_[2],
            v_[0] * o.v_[1] - v_[1] * o.v_[0]};
  }
  double norm() const { return std::sqrt((v_[0] * v_[0] + v_[1] * v_[1]) + v_[2] * v_[2]); }
  Vector3d normalized() const {
    const double n = norm();
    return {v_[0] / n, v_[1] / n, v_[2] / n};
  }

 private:
  double v_[3];
};
inline Vector3d operator*(double s, const Vector3d& v) { return v * s; }

class AngleAxisd {
 public:
  AngleAxisd(double angle, const Vector3d& axis) : angle_(angle), axis_(axis) {}
  double angle() const { return angle_; }
  const Vector3d& axis() const { return axis_; }

 private:
  double angle_;
  Vector3d axis_;
};

class Quaterniond {
 public:
  Quaterniond() : w_(1.0), x_(0.0), y_(0.0), z_(0.0) {}
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  explicit Quaterniond(const AngleAxisd& aa) {
    const double h = 0.5 * aa.angle(), s = std::sin(h);
    w_ = std::cos(h);
    x_ = s * aa.axis().x();
    y_ = s * aa.axis().y();
    z_ = s * aa.axis().z();
  }
  Quaterniond& operator=(const AngleAxisd& aa) { return *this = Quaterniond(aa); }
  static Quaterniond Identity() { return Quaterniond(); }
  double w() const { return w_; }
  double x() const { return x_; }
  double y() const { return y_; }
  double z() const { return z_; }
  Vector3d vec() const { return {x_, y_, z_}; }
  Quaterniond normalized() const {
    // Eigen 3.3 (SSE2): coeffs x,y,z,w squared, reduced two lanes at a time
    const double n = std::sqrt((x_ * x_ + z_ * z_) + (y_ * y_ + w_ * w_));
    return {w_ / n, x_ / n, y_ / n, z_ / n};
  }
  // Hamilton product
  Quaterniond operator*(const Quaterniond& b) const {
    return {w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
            w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_};
  }
  // rotation of a vector, Eigen's operation order: uv = vec x v; uv += uv; v + w*uv + vec x uv
  Vector3d operator*(const Vector3d& v) const {
    Vector3d uv = vec().cross(v);
    uv += uv;
    return v + w_ * uv + vec().cross(uv);
  }

 private:
  double w_, x_, y_, z_;
};

}  // namespace Eigen
#endif
