// Source-compatible mirror of /root/reference/eggshell/ensembles.h:25-200.  An Ensemble owns one
// world of a device batch (egg_batch, include/egg_cuda.h); Init()/Step() forward to the C ABI, the
// Body / Contact objects are refreshed from the device after every call.  Where the reference
// aborts through Panic() (ensembles.cc:281,404,533) this mirror calls Panic() too, so callers see
// the same behaviour; the C ABI underneath reports status words instead.
#ifndef EGGSHELL_ENSEMBLES_H_
#define EGGSHELL_ENSEMBLES_H_
#include <array>
#include <limits>
#include <memory>
#include <vector>
#include "constraints.h"

typedef std::vector<std::shared_ptr<Body>> ComponentsList;
typedef std::vector<std::shared_ptr<Joint>> JointsList;
typedef std::vector<std::shared_ptr<Contact>> ContactsList;
typedef std::vector<std::shared_ptr<Constraint>> ConstraintsList;

struct egg_batch;

class Ensemble {
 public:
  Ensemble();
  virtual ~Ensemble();
  virtual void Init();                                            // ensembles.cc:24-29
  virtual MatrixXd ComputeJ() const;                              // ensembles.cc:31-36
  virtual MatrixXd ComputeJ(ArrayXb* C, VectorXd* x_lo, VectorXd* x_hi) const;   // ensembles.cc:38-87
  enum struct Integrator { EXPLICIT_EULER = 0, OPEN_DYNAMICS_ENGINE, IMPLICIT_MIDPOINT };
  virtual void Step(double dt, Integrator g = Integrator::OPEN_DYNAMICS_ENGINE);   // ensembles.cc:390-427
  virtual VectorXd ComputeJDotV() const;                          // ensembles.cc:89-98: Panics, as the reference does
  void InitStabilize();                                           // ensembles.cc:602-622
  void PostStabilize(int max_steps = 500);                        // ensembles.cc:624-645
  virtual void Draw() const;
  bool CheckConservationOfEnergy();                               // ensembles.cc:186-200
  const MatrixXd& M_inverse() const { return M_inverse_; }
  const ConstraintsList constraints() const;                      // joints then contacts, ensembles.cc:234-239
  const ComponentsList& components() const { return components_; }

  // Extensions of the mirror (not in the reference): choose the solver wired into ComputeVDot
  // before Init().  0 = dense Murty (reference default), 1 = matrix-free PGS.
  void SetSolver(int solver, int k_max = 500) { solver_ = solver; k_max_ = k_max; }
  int status() const { return status_; }                          // egg_status bits of the last call
  int last_sweeps() const { return sweeps_; }

 protected:
  int n_ = 0;
  ComponentsList components_;
  JointsList joints_;
  ContactsList contacts_;
  MatrixXd M_inverse_;
  VectorXd external_force_torque_;
  double total_rotational_ke_ = std::numeric_limits<double>::infinity();   // ensembles.h:92
  void UpdateContacts();                                          // ensembles.cc:445-480 (+ :241-329)

 private:
  egg_batch* batch_ = nullptr;
  int solver_ = 0, k_max_ = 500, status_ = 0, sweeps_ = 0;
  void Upload();
  void UploadState();                                             // Body p, R, v, w -> device (setters between steps take effect)
  void Download(bool with_contacts);
};

class Chain : public Ensemble {                                   // ensembles.cc:668-707
 public:
  Chain(int num_links, const Vector3d& anchor_position);
};

class Cairn : public Ensemble {                                   // ensembles.cc:709-728
 public:
  Cairn(int num_rocks, const std::array<double, 2>& x_bound, const std::array<double, 2>& y_bound,
        const std::array<double, 2>& z_bound);
 private:
  const double max_init_v_ = 1;
  const double max_init_w_ = 1;
};
#endif
