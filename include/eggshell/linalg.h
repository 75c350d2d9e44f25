// Minimal Eigen-free stand-ins for the Eigen types that appear in eggshell's public headers
// (/root/reference/eggshell/utils.h:8-15: Vector3d, Matrix3d, Quaterniond, VectorXd, MatrixXd,
// ArrayXb).  Only what the step API needs: construction, element access, a few products.  If the
// including project already provides Eigen, define EGGSHELL_USE_EIGEN before including.
#ifndef EGGSHELL_LINALG_H_
#define EGGSHELL_LINALG_H_
#ifdef EGGSHELL_USE_EIGEN
#include "Eigen/Dense"
using Eigen::Matrix3d;
using Eigen::MatrixXd;
using Eigen::Quaterniond;
using Eigen::Vector3d;
using Eigen::VectorXd;
typedef Eigen::Array<bool, Eigen::Dynamic, 1> ArrayXb;
#else
#include <array>
#include <cmath>
#include <cstddef>
#include <vector>

struct Vector3d {
  double v[3];
  Vector3d() : v{0, 0, 0} {}
  Vector3d(double x, double y, double z) : v{x, y, z} {}
  static Vector3d Zero() { return Vector3d(); }
  static Vector3d UnitX() { return Vector3d(1, 0, 0); }
  static Vector3d UnitZ() { return Vector3d(0, 0, 1); }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  Vector3d operator*(double s) const { return Vector3d(v[0] * s, v[1] * s, v[2] * s); }
  Vector3d operator/(double s) const { return Vector3d(v[0] / s, v[1] / s, v[2] / s); }
  double dot(const Vector3d& o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
  Vector3d cross(const Vector3d& o) const {
    return Vector3d(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
  }
  double norm() const { return std::sqrt(dot(*this)); }
};
inline Vector3d operator*(double s, const Vector3d& a) { return a * s; }

struct Matrix3d {   // row-major
  double m[9];
  Matrix3d() : m{0, 0, 0, 0, 0, 0, 0, 0, 0} {}
  static Matrix3d Zero() { return Matrix3d(); }
  static Matrix3d Identity() { Matrix3d r; r.m[0] = r.m[4] = r.m[8] = 1; return r; }
  double& operator()(int i, int j) { return m[3 * i + j]; }
  double operator()(int i, int j) const { return m[3 * i + j]; }
  Vector3d col(int c) const { return Vector3d(m[c], m[3 + c], m[6 + c]); }
  Matrix3d transpose() const { Matrix3d r; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = (*this)(j, i); return r; }
  Matrix3d operator*(const Matrix3d& o) const {
    Matrix3d r;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r(i, j) = (*this)(i, 0) * o(0, j) + (*this)(i, 1) * o(1, j) + (*this)(i, 2) * o(2, j);
    return r;
  }
  Vector3d operator*(const Vector3d& a) const {
    return Vector3d(m[0] * a[0] + m[1] * a[1] + m[2] * a[2], m[3] * a[0] + m[4] * a[1] + m[5] * a[2], m[6] * a[0] + m[7] * a[1] + m[8] * a[2]);
  }
  Matrix3d operator*(double s) const { Matrix3d r; for (int k = 0; k < 9; k++) r.m[k] = m[k] * s; return r; }
};

struct Quaterniond {
  double w_, x_, y_, z_;
  Quaterniond() : w_(1), x_(0), y_(0), z_(0) {}
  Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
  static Quaterniond FromAngleAxis(double angle, const Vector3d& axis) {
    double h = 0.5 * angle, s = std::sin(h);
    return Quaterniond(std::cos(h), s * axis[0], s * axis[1], s * axis[2]);
  }
  Quaterniond operator*(const Quaterniond& b) const {
    return Quaterniond(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                       w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  Matrix3d matrix() const {
    double tx = 2 * x_, ty = 2 * y_, tz = 2 * z_, twx = tx * w_, twy = ty * w_, twz = tz * w_;
    double txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
    Matrix3d r;
    r(0, 0) = 1 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
    r(1, 0) = txy + twz; r(1, 1) = 1 - (txx + tzz); r(1, 2) = tyz - twx;
    r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1 - (txx + tyy);
    return r;
  }
};

struct VectorXd {
  std::vector<double> d;
  VectorXd() {}
  explicit VectorXd(int n) : d((size_t)n, 0.0) {}
  static VectorXd Zero(int n) { return VectorXd(n); }
  int size() const { return (int)d.size(); }
  int rows() const { return (int)d.size(); }
  double& operator()(int i) { return d[(size_t)i]; }
  double operator()(int i) const { return d[(size_t)i]; }
};
struct MatrixXd {   // row-major
  int r_ = 0, c_ = 0;
  std::vector<double> d;
  MatrixXd() {}
  MatrixXd(int r, int c) : r_(r), c_(c), d((size_t)r * c, 0.0) {}
  static MatrixXd Zero(int r, int c) { return MatrixXd(r, c); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int i, int j) { return d[(size_t)i * c_ + j]; }
  double operator()(int i, int j) const { return d[(size_t)i * c_ + j]; }
};
struct ArrayXb {
  std::vector<unsigned char> d;
  ArrayXb() {}
  explicit ArrayXb(int n) : d((size_t)n, 0) {}
  int size() const { return (int)d.size(); }
  int rows() const { return (int)d.size(); }
  unsigned char& operator()(int i) { return d[(size_t)i]; }
  bool operator()(int i) const { return d[(size_t)i] != 0; }
};
#endif  // EGGSHELL_USE_EIGEN
#endif
