// Source-compatible mirror of /root/reference/eggshell/body.h:13-96.  A Body is a host-side view of
// one rigid body; the Ensemble uploads it to the device batch and refreshes it after every Step.
#ifndef EGGSHELL_BODY_H_
#define EGGSHELL_BODY_H_
#include "linalg.h"

class Body {
 public:
  virtual ~Body() = default;
  Body() : m_(1.0), R_(Matrix3d::Identity()) { I_ = CalculateInertia(m_); }
  explicit Body(const Vector3d& p, const Matrix3d& R) : p_(p), m_(1.0), R_(R) { I_ = CalculateInertia(m_); }
  explicit Body(const Vector3d& p, const Vector3d& v, const Matrix3d& R, const Vector3d& w)
      : p_(p), v_(v), m_(1.0), R_(R), w_(w) { I_ = CalculateInertia(m_); }
  explicit Body(const Vector3d& p, const Vector3d& v, double m, const Matrix3d& R, const Vector3d& w, const Matrix3d& I)
      : p_(p), v_(v), m_(m), R_(R), w_(w), I_(I) {}
  explicit Body(const Vector3d& p, const Vector3d& v, const Quaterniond& q, const Vector3d& w)
      : p_(p), v_(v), m_(1.0), R_(q.matrix()), w_(w) { I_ = CalculateInertia(m_); }
  explicit Body(const Vector3d& p, const Vector3d& v, double m, const Quaterniond& q, const Vector3d& w, const Matrix3d& I)
      : p_(p), v_(v), m_(m), R_(q.matrix()), w_(w), I_(I) {}

  const Vector3d& p() const { return p_; }
  const Vector3d& v() const { return v_; }
  double m() const { return m_; }
  const Matrix3d& R() const { return R_; }
  const Vector3d w_b() const { return R_.transpose() * w_; }
  const Vector3d& w_g() const { return w_; }
  const Matrix3d& I_b() const { return I_; }
  const Matrix3d I_g() const { return R_ * I_ * R_.transpose(); }

  enum struct BodyType { Box = 0 };

  void SetP(const Vector3d& p) { p_ = p; }
  void SetV(const Vector3d& v) { v_ = v; }
  void SetM(double m) { m_ = m; }
  void SetR(const Matrix3d& R) { R_ = R; }
  void SetR(const Quaterniond& q) { R_ = q.matrix(); }
  void SetW_GlobalFrame(const Vector3d& w) { w_ = w; }
  void SetW_BodyFrame(const Vector3d& w) { w_ = R_ * w; }
  void SetI(const Matrix3d& I) { I_ = I; }
  void Rotate(const Matrix3d& R) { R_ = R * R_; }
  double GetRotationalKE() const { Vector3d wb = w_b(); return wb.dot(I_ * wb); }
  void Draw() const;
  const Vector3d GetSideLengths() const { return side_lengths_; }

 private:
  Vector3d p_, v_;
  double m_;
  Matrix3d R_;
  Vector3d w_;
  Matrix3d I_;
  const Vector3d side_lengths_ = Vector3d(0.3, 0.3, 0.3);   // body.h:91
  Matrix3d CalculateInertia(double m) const {                // body.cc:19-36
    double x = side_lengths_(0), y = side_lengths_(1), z = side_lengths_(2);
    Matrix3d I;
    I(0, 0) = m / 12 * (y * y + z * z);
    I(1, 1) = m / 12 * (x * x + z * z);
    I(2, 2) = m / 12 * (x * x + y * y);
    return I;
  }
};
#endif
