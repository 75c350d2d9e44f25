// Host side of the drop-in: Ensemble / Chain / Cairn (mirror of
// /root/reference/eggshell/ensembles.{h,cc}) forwarding to the C ABI of libeggshell_b200.so.
// No numerical work of the step happens here: Init/Step/UpdateContacts are device calls.
#include "eggshell/ensembles.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "egg_cuda.h"
#include "eggshell/model.h"

namespace {
void check(int rc, const char* what) {
  if (rc != EGG_OK) Panic("%s failed (%d): %s", what, rc, egg_last_error());
}
}  // namespace

Ensemble::Ensemble() {}
Ensemble::~Ensemble() {
  if (batch_) egg_destroy(batch_);
}

const ConstraintsList Ensemble::constraints() const {
  ConstraintsList c;
  c.insert(c.end(), joints_.begin(), joints_.end());
  c.insert(c.end(), contacts_.begin(), contacts_.end());
  return c;
}

void Ensemble::Upload() {
  const int n = n_;
  std::vector<double> p(3 * n), R(9 * n), v(3 * n), w(3 * n), m(n), I(9 * n);
  for (int i = 0; i < n; i++) {
    const Body& b = *components_.at(i);
    for (int k = 0; k < 3; k++) { p[3 * i + k] = b.p()(k); v[3 * i + k] = b.v()(k); w[3 * i + k] = b.w_g()(k); }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { R[9 * i + 3 * r + c] = b.R()(r, c); I[9 * i + 3 * r + c] = b.I_b()(r, c); }
    m[i] = b.m();
  }
  check(egg_set_bodies(batch_, p.data(), R.data(), v.data(), w.data(), m.data(), I.data(), nullptr), "egg_set_bodies");
  const int nj = (int)joints_.size();
  if (nj) {
    std::vector<int> i0(nj), i1(nj);
    std::vector<double> c0(3 * nj), c1(3 * nj);
    for (int k = 0; k < nj; k++) {
      i0[k] = joints_[k]->i0_; i1[k] = joints_[k]->i1_;
      for (int a = 0; a < 3; a++) { c0[3 * k + a] = joints_[k]->c0()(a); c1[3 * k + a] = joints_[k]->c1()(a); }
    }
    check(egg_set_joints(batch_, i0.data(), i1.data(), c0.data(), c1.data()), "egg_set_joints");
  }
}

void Ensemble::UploadState() {
  const int n = n_;
  std::vector<double> p(3 * n), R(9 * n), v(3 * n), w(3 * n);
  for (int i = 0; i < n; i++) {
    const Body& b = *components_.at(i);
    for (int k = 0; k < 3; k++) { p[3 * i + k] = b.p()(k); v[3 * i + k] = b.v()(k); w[3 * i + k] = b.w_g()(k); }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) R[9 * i + 3 * r + c] = b.R()(r, c);
  }
  check(egg_set_state(batch_, p.data(), R.data(), v.data(), w.data()), "egg_set_state");
}

void Ensemble::Download(bool with_contacts) {
  const int n = n_;
  std::vector<double> p(3 * n), R(9 * n), v(3 * n), w(3 * n);
  check(egg_get_bodies(batch_, p.data(), R.data(), v.data(), w.data()), "egg_get_bodies");
  for (int i = 0; i < n; i++) {
    Body& b = *components_.at(i);
    b.SetP(Vector3d(p[3 * i], p[3 * i + 1], p[3 * i + 2]));
    b.SetV(Vector3d(v[3 * i], v[3 * i + 1], v[3 * i + 2]));
    b.SetW_GlobalFrame(Vector3d(w[3 * i], w[3 * i + 1], w[3 * i + 2]));
    Matrix3d Rm;
    for (int k = 0; k < 9; k++) Rm.m[k] = R[9 * i + k];
    b.SetR(Rm);
  }
  int stats[8] = {0};
  check(egg_get_status(batch_, &status_, stats, nullptr), "egg_get_status");
  sweeps_ = stats[4];
  if (!with_contacts) return;
  const int mc = egg_capacity(batch_), nj = (int)joints_.size();
  int count = 0;
  std::vector<int> i0(mc), i1(mc), code(mc);
  std::vector<double> pos(3 * mc), nrm(3 * mc), depth(mc), lam(3 * (size_t)(nj + mc));
  check(egg_get_contacts(batch_, &count, i0.data(), i1.data(), pos.data(), nrm.data(), depth.data(), code.data(), lam.data(), nullptr),
        "egg_get_contacts");
  contacts_.clear();
  for (int k = 0; k < count; k++) {
    ContactGeometry cg(Vector3d(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]), Vector3d(nrm[3 * k], nrm[3 * k + 1], nrm[3 * k + 2]), depth[k]);
    std::shared_ptr<Contact> c;
    if (i0[k] < 0) {
      c = std::shared_ptr<Contact>(new Contact(components_.at(i1[k]), i1[k], cg));      // ensembles.cc:454
    } else {
      CollisionInfo ci;
      ci.code = code[k];
      ci.separating_axis = cg.normal;
      c = std::shared_ptr<Contact>(new Contact(components_.at(i0[k]), i0[k], components_.at(i1[k]), i1[k], cg, ci));   // :469-470
    }
    for (int a = 0; a < 3; a++) c->lambda[a] = lam[3 * (size_t)(nj + k) + a];
    contacts_.push_back(c);
  }
}

void Ensemble::Init() {
  n_ = (int)components_.size();
  if (batch_) { egg_destroy(batch_); batch_ = nullptr; }
  egg_desc d;
  check(egg_desc_default(&d, 1, n_, (int)joints_.size()), "egg_desc_default");
  d.solver = solver_;
  d.k_max = k_max_;
  d.taps = 0;
  check(egg_create(&d, &batch_), "egg_create");
  Upload();
  check(egg_init(batch_), "egg_init");
  // M_inverse_ / external_force_torque_ as frozen by Init (ensembles.cc:202-222)
  std::vector<double> ml(n_), ma(9 * n_), f(6 * n_);
  check(egg_get_static(batch_, ml.data(), ma.data(), f.data()), "egg_get_static");
  M_inverse_ = MatrixXd::Zero(6 * n_, 6 * n_);
  external_force_torque_ = VectorXd::Zero(6 * n_);
  for (int i = 0; i < n_; i++) {
    for (int r = 0; r < 3; r++) {
      M_inverse_(6 * i + r, 6 * i + r) = ml[i];
      for (int c = 0; c < 3; c++) M_inverse_(6 * i + 3 + r, 6 * i + 3 + c) = ma[9 * i + 3 * r + c];
    }
    for (int k = 0; k < 6; k++) external_force_torque_(6 * i + k) = f[6 * i + k];
  }
  Download(false);
  if (status_ & EGG_ST_BAD_INIT) Panic("Check initial conditions failed.");                       // ensembles.cc:27
  if (status_ & EGG_ST_JOINT_CONFLICT) Panic("Joint constraints conflict or cause overconstraint.");   // :281
}

void Ensemble::UpdateContacts() {
  check(egg_update_contacts(batch_), "egg_update_contacts");
  Download(true);
}

void Ensemble::Step(double dt, Integrator g) {
  if (!batch_) Panic("Ensemble::Step before Init");
  if (g == Integrator::IMPLICIT_MIDPOINT) Panic("Implicit midpoint integrator is not properly implemented and tested.");   // :403-405
  if (g == Integrator::EXPLICIT_EULER) Panic("EXPLICIT_EULER is not part of the accelerated path (ensembles.cc:397-402).");
  // The reference's Step reads the Body objects (GetVelocities, ensembles.cc:429-436), so SetP /
  // SetR / SetV / SetW_GlobalFrame between steps take effect: push the host state first (the
  // values are the ones the last Download wrote unless the caller changed them).
  UploadState();
  check(egg_step(batch_, dt, EGG_OPEN_DYNAMICS_ENGINE, 1), "egg_step");
  Download(true);
  if (status_ & EGG_ST_JOINT_CONFLICT) Panic("Joint constraints conflict or cause overconstraint.");
  if (status_ & EGG_ST_LCP_FAILED) Panic("Lcp::MixedConstraintsSolver exited without reaching a solution.");   // :531-534
}

void Ensemble::InitStabilize() {
  // ensembles.cc:602-622 on the device: relax positions while the squared constraint error exceeds
  // 1e-9 (at most 100 relaxations), then CheckAndCorrectEnsembleState.
  int steps = 0;
  double e2 = 0;
  check(egg_init_stabilize(batch_, 100, &steps, &e2), "egg_init_stabilize");
  Download(true);
  std::printf("Pre-stabilization steps count : %d\nFinal err_sq : %g\n", steps, e2);
}

void Ensemble::PostStabilize(int max_steps) {
  // ensembles.cc:624-645 on the device: StepPostStabilization(dt = 0.1) while the squared
  // constraint error exceeds 1e-9, at most max_steps times.
  if (!batch_) Panic("Ensemble::PostStabilize before Init");
  UploadState();
  int steps = 0;
  double e2 = 0;
  check(egg_post_stabilize(batch_, max_steps, &steps, &e2), "egg_post_stabilize");
  Download(false);
}

VectorXd Ensemble::ComputeJDotV() const {
  // ensembles.cc:95-98 / :123-129: both halves Panic in the reference (only the ODE stepper, which
  // never needs Jdot v, works with contacts).
  Panic("Jdot is hardcoded to have 3 rows. Should be decided by the Constraint.");
  return VectorXd::Zero(0);
}

bool Ensemble::CheckConservationOfEnergy() {   // ensembles.cc:186-200
  double energy = 0;
  for (const auto& b : components_) energy = energy + b->GetRotationalKE();
  if (std::fabs(energy - total_rotational_ke_) > 1e-9 && total_rotational_ke_ != std::numeric_limits<double>::infinity()) {
    std::printf("Total rotational KE was %g, now it's %g\n", total_rotational_ke_, energy);
    total_rotational_ke_ = energy;
    return false;
  }
  total_rotational_ke_ = energy;
  return true;
}

MatrixXd Ensemble::ComputeJ() const {
  ArrayXb C;
  VectorXd lo, hi;
  return ComputeJ(&C, &lo, &hi);
}

// Dense J in reference row order, rebuilt on the host from the constraint descriptors (a tap for
// callers that inspect J, e.g. the reference's solver tests; the step never forms it).
MatrixXd Ensemble::ComputeJ(ArrayXb* C, VectorXd* x_lo, VectorXd* x_hi) const {
  const int nc = (int)(joints_.size() + contacts_.size());
  MatrixXd J = MatrixXd::Zero(3 * nc, 6 * n_);
  *C = ArrayXb(3 * nc);
  *x_lo = VectorXd::Zero(3 * nc);
  *x_hi = VectorXd::Zero(3 * nc);
  auto put = [&](int row, int body, const Matrix3d& lin, const Matrix3d& ang) {
    if (body < 0) return;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { J(row + r, 6 * body + c) = lin(r, c); J(row + r, 6 * body + 3 + c) = ang(r, c); }
  };
  auto crossmat = [](const Vector3d& a) { Matrix3d m; m(0, 1) = -a(2); m(0, 2) = a(1); m(1, 0) = a(2); m(1, 2) = -a(0); m(2, 0) = -a(1); m(2, 1) = a(0); return m; };
  int row = 0;
  for (const auto& j : joints_) {                      // joints.cc:13-35
    const Body& b0 = *components_.at(j->i0_);
    put(row, j->i0_, Matrix3d::Identity(), crossmat(b0.R() * j->c0()) * -1.0);
    if (j->i1_ >= 0) put(row, j->i1_, Matrix3d::Identity() * -1.0, crossmat(components_.at(j->i1_)->R() * j->c1()));
    for (int k = 0; k < 3; k++) (*C)(row + k) = 1;
    row += 3;
  }
  for (const auto& c : contacts_) {                    // contact.cc:38-117, BOX friction
    const Vector3d n = c->geometry().normal, z(0, 0, 1);
    // FromTwoVectors(n, z) for n not anti-parallel to z (utils.cc:233-236)
    double cs = n.dot(z) / n.norm();
    Matrix3d Rc = Matrix3d::Identity();
    if (cs > -1.0 + 1e-12) {
      Vector3d v0 = n / n.norm(), ax = v0.cross(z);
      double s = std::sqrt((1 + cs) * 2);
      Rc = Quaterniond(s * 0.5, ax(0) / s, ax(1) / s, ax(2) / s).matrix();
    }
    if (c->i0_ >= 0) put(row, c->i0_, Rc * -1.0, Rc * crossmat(c->geometry().position - components_.at(c->i0_)->p()));
    put(row, c->i1_, Rc, Rc * (crossmat(c->geometry().position - components_.at(c->i1_)->p()) * -1.0));
    (*x_lo)(row) = -1; (*x_lo)(row + 1) = -1; (*x_lo)(row + 2) = 0;
    (*x_hi)(row) = 1; (*x_hi)(row + 1) = 1; (*x_hi)(row + 2) = INFINITY;
    row += 3;
  }
  return J;
}

void Ensemble::Draw() const {
  for (const auto& b : components_) DrawBox(b->p(), b->R(), b->GetSideLengths());
  for (const auto& j : joints_) DrawPoint(j->GetConstraintPosition());
  for (const auto& c : contacts_) { DrawPoint(c->geometry().position); DrawLine(c->geometry().position, c->geometry().position + c->geometry().normal * 0.1); }
}

Chain::Chain(int num_links, const Vector3d& anchor_position) {
  if (num_links <= 0) Panic("Chain needs at least one link");
  n_ = num_links;
  Quaterniond q = Quaterniond::FromAngleAxis(0.95531661812451, Vector3d::UnitZ()) * Quaterniond::FromAngleAxis(M_PI / 4, Vector3d::UnitX());
  Matrix3d R = q.matrix();
  for (int i = 0; i < n_; i++) {
    Vector3d p(std::sqrt(3.0) * 0.3 * i, 0, 0);
    components_.push_back(std::shared_ptr<Body>(new Body(p + anchor_position, Vector3d::Zero(), R, Vector3d::Zero())));
  }
  Vector3d c1(0.15, -0.15, 0.15), c2(-0.15, 0.15, -0.15);
  for (int i = 0; i < n_ - 1; i++)
    joints_.push_back(std::shared_ptr<Joint>(new BallAndSocketJoint(components_.at(i), i, c1, components_.at(i + 1), i + 1, c2)));
  joints_.push_back(std::shared_ptr<Joint>(new BallAndSocketJoint(components_.at(0), 0, Vector3d::Zero(), components_.at(0)->p())));
}

namespace {
// Eigen's Random()/UnitRandom() draw from std::rand(); the reference never seeds it.
double eigen_random(double lo, double hi) { return lo + (hi - lo) * double(std::rand()) / double(RAND_MAX); }
}  // namespace

Cairn::Cairn(int num_rocks, const std::array<double, 2>& xb, const std::array<double, 2>& yb, const std::array<double, 2>& zb) {
  n_ = num_rocks;
  Matrix3d I = Matrix3d::Identity() * 0.1;
  for (int i = 0; i < num_rocks; i++) {
    double r0 = eigen_random(-1, 1), r1 = eigen_random(-1, 1), r2 = eigen_random(-1, 1);
    Vector3d p((r0 + 1) / 2 * std::fabs(xb[1] - xb[0]) + std::fmin(xb[0], xb[1]), (r1 + 1) / 2 * std::fabs(yb[1] - yb[0]) + std::fmin(yb[0], yb[1]),
               (r2 + 1) / 2 * std::fabs(zb[1] - zb[0]) + std::fmin(zb[0], zb[1]));
    double u1 = eigen_random(0, 1), u2 = eigen_random(0, 2 * M_PI), u3 = eigen_random(0, 2 * M_PI);
    double a = std::sqrt(1 - u1), b = std::sqrt(u1);
    Quaterniond q(a * std::sin(u2), a * std::cos(u2), b * std::sin(u3), b * std::cos(u3));
    Vector3d v(eigen_random(-1, 1), eigen_random(-1, 1), eigen_random(-1, 1));
    Vector3d w(eigen_random(-1, 1), eigen_random(-1, 1), eigen_random(-1, 1));
    components_.push_back(std::shared_ptr<Body>(new Body(p, v * max_init_v_, 1.0, q, w * max_init_w_, I)));
  }
}
