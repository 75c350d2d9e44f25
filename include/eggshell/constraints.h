// Mirror of /root/reference/eggshell/constraints.h:14-48, joints.h:12-50 and contact.h:11-45.
// Constraints are descriptors: their Jacobians are assembled on the device (egg_assemble_kernel).
#ifndef EGGSHELL_CONSTRAINTS_H_
#define EGGSHELL_CONSTRAINTS_H_
#include <memory>
#include <string>
#include "body.h"
#include "collision.h"

class Constraint {
 public:
  explicit Constraint(const std::shared_ptr<Body> b0, int i0, const std::shared_ptr<Body> b1, int i1)
      : i0_(i0), i1_(i1), b0_(b0), b1_(b1) {}
  virtual ~Constraint() = default;
  virtual VectorXd ComputeError() const = 0;
  virtual Vector3d GetConstraintPosition() const = 0;
  virtual std::string PrintInfo() const { return "Constraint"; }
  int i0_ = -1;   // indices into the ensemble's component list, -1 = world / ground
  int i1_ = -1;
 protected:
  const std::shared_ptr<Body> b0_, b1_;
};

class Joint : public Constraint {
 public:
  explicit Joint(const std::shared_ptr<Body> b0, int i0, const Vector3d& c0, const Vector3d& c1)
      : Constraint(b0, i0, nullptr, -1), c0_(c0), c1_(c1) {}
  explicit Joint(const std::shared_ptr<Body> b0, int i0, const Vector3d& c0, const std::shared_ptr<Body> b1, int i1, const Vector3d& c1)
      : Constraint(b0, i0, b1, i1), c0_(c0), c1_(c1) {}
  const Vector3d& c0() const { return c0_; }
  const Vector3d& c1() const { return c1_; }
 protected:
  Vector3d c0_, c1_;
};

class BallAndSocketJoint : public Joint {
 public:
  using Joint::Joint;
  VectorXd ComputeError() const override {           // joints.cc:3-11
    Vector3d e = b0_->p() + b0_->R() * c0_;
    e = (b1_ == nullptr) ? e - c1_ : e - b1_->p() - b1_->R() * c1_;
    VectorXd r(3);
    for (int k = 0; k < 3; k++) r(k) = e(k);
    return r;
  }
  Vector3d GetConstraintPosition() const override {  // joints.cc:56-75
    Vector3d p0 = b0_->p() + b0_->R() * c0_;
    if (b1_ == nullptr) return p0;
    return (p0 + (b1_->p() + b1_->R() * c1_)) / 2;
  }
};

class Contact : public Constraint {
 public:
  explicit Contact(const std::shared_ptr<Body> b, int index, const ContactGeometry& cg) : Constraint(nullptr, -1, b, index), cg_(cg) {}
  explicit Contact(const std::shared_ptr<Body> b0, int i0, const std::shared_ptr<Body> b1, int i1, const ContactGeometry& cg, const CollisionInfo& ci)
      : Constraint(b0, i0, b1, i1), cg_(cg), ci_(ci) {}
  enum struct FrictionModel { NO_FRICTION, INFINITE, BOX, COULOMB_PYRAMID };
  VectorXd ComputeError() const override { VectorXd e(3); e(2) = -cg_.depth; return e; }   // contact.cc:14-22
  Vector3d GetConstraintPosition() const override { return cg_.position; }
  const ContactGeometry& geometry() const { return cg_; }
  const CollisionInfo& info() const { return ci_; }
  double lambda[3] = {0, 0, 0};   // multipliers of the last Step (tangent, tangent, normal)
 private:
  const ContactGeometry cg_;
  const CollisionInfo ci_;
};
#endif
