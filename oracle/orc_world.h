// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of eggshell's ensemble stepper and of the matrix-free / dense iterative
// solvers.  Follows /root/reference/eggshell/ensembles.cc:24-87,156-171,202-591,602-728,
// joints.cc:3-75, contact.cc:14-117, body.cc:19-36, sparse_iterations.cc:35-226,
// sparse_iterations_utils.cc:12-720.
#pragma once
#include <cstdlib>
#include "orc_collision.h"
#include "orc_lcp.h"

namespace orc {

struct Body {                     // body.h:13-96
  Vec3 p, v, w;
  double m = 1;
  Mat3 R = Mat3::identity();
  Mat3 I;                         // body frame
  Vec3 side = Vec3(0.3, 0.3, 0.3);   // body.h:91 (box side lengths; sphere / capsule dims, orc_collision.h)
  int shape = 0;                  // 0 box (the reference's only collider), 1 sphere, 2 capsule
  Mat3 I_g() const { return R * I * transpose(R); }   // body.h:58, (R*I)*R^T
};
// body.cc:19-36
inline Mat3 box_inertia(double m, const Vec3& s) {
  Mat3 r;
  r.m[0][0] = m / 12 * (s.y * s.y + s.z * s.z);
  r.m[1][1] = m / 12 * (s.x * s.x + s.z * s.z);
  r.m[2][2] = m / 12 * (s.x * s.x + s.y * s.y);
  return r;
}

struct Joint { int i0 = -1, i1 = -1; Vec3 c0, c1; };   // joints.h:12-29 (ball and socket only)
struct Contact {                                       // contact.h:11-45
  int i0 = -1, i1 = -1;
  ContactGeometry cg;
  int code = 0;
};

// One constraint's 3x6 Jacobian pair + row metadata (constraints.h:21-26).
struct Rows3 {
  double j0[3][6], j1[3][6];
  unsigned char type[3];   // 1 = equality
  double lo[3], hi[3];
  int i0, i1;
};

enum Solver { SOLVER_DENSE_MURTY = 0, SOLVER_PGS = 1, SOLVER_JACOBI = 2, SOLVER_SOR = 3 };
enum CfmMode { CFM_AUTO = 0, CFM_ALWAYS = 1, CFM_NEVER = 2 };
enum Quirk {
  QUIRK_GS_BOUNDS_SHIFT = 1,   // q2: block i>0 is projected with the bounds of its neighbour
  QUIRK_DENSE_IGNORES_BOUNDS = 2,  // q1: Murty runs with [0,inf) on every inequality row
};
enum Status {
  ST_OK = 0,
  ST_LCP_FAILED = 1,        // reference would Panic (ensembles.cc:531-534)
  ST_JOINT_CONFLICT = 2,    // reference would Panic (ensembles.cc:280-285)
  ST_BAD_INIT = 4,          // CheckInitialConditions failed (ensembles.cc:27)
};

struct Params {
  double erp = 0.2;              // ensembles.h:166
  double cfm = 0.01;             // ensembles.cc:14
  double min_constraint_dist = 1e-6;   // ensembles.cc:15
  double tol = kAllowNumericalError;   // constants.h:5
  int k_max = 500;               // sparse_iterations.cc:19
  double gravity[3] = {0, 0, -9.8};    // constants.h:8
  int solver = SOLVER_DENSE_MURTY;
  int cfm_mode = CFM_AUTO;
  int quirks = QUIRK_GS_BOUNDS_SHIFT | QUIRK_DENSE_IGNORES_BOUNDS;
};

struct StepStats {
  int n_contacts_raw = 0, n_contacts = 0, n_rows = 0;
  int n_pair_tests = 0, n_pair_hits = 0;
  int sweeps = 0, pivots = 0, cfm_applied = 0, status = 0;
  double residual = 0;
};

struct World {
  int n = 0;
  std::vector<Body> bodies;
  std::vector<Joint> joints;
  std::vector<Contact> contacts;
  std::vector<Mat3> Minv_ang;       // ensembles.cc:202-212, frozen at Init (q4)
  std::vector<double> Minv_lin;
  std::vector<double> f_ext;        // 6n, ensembles.cc:214-222, frozen at Init (q4)
  Params prm;
  StepStats stats;
  // parity taps of the last step
  std::vector<int> pair_hit_i, pair_hit_j, pair_hit_code, pair_hit_count;
  std::vector<int> ground_count;
  Vec lambda, rhs;
  std::vector<int> row_state;       // PGS: 0 free, 1 at lo, 2 at hi, 3 equality.  Murty: S mask
};

// ---------------------------------------------------------------------------------------------
// Constraint rows.

inline Vec3 joint_error(const World& W, const Joint& j) {        // joints.cc:3-11
  const Body& b0 = W.bodies[j.i0];
  if (j.i1 < 0) return b0.p + b0.R * j.c0 - j.c1;
  const Body& b1 = W.bodies[j.i1];
  return b0.p + b0.R * j.c0 - b1.p - b1.R * j.c1;
}
inline Vec3 joint_position(const World& W, const Joint& j) {     // joints.cc:56-75
  const Body& b0 = W.bodies[j.i0];
  Vec3 p0 = b0.p + b0.R * j.c0;
  if (j.i1 < 0) return p0;
  const Body& b1 = W.bodies[j.i1];
  Vec3 p1 = b1.p + b1.R * j.c1;
  return (p0 + p1) / 2;
}
inline void put_block(double dst[3][6], const Mat3& lin, const Mat3& ang) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { dst[r][c] = lin.m[r][c]; dst[r][3 + c] = ang.m[r][c]; }
}
inline void joint_rows(const World& W, const Joint& j, Rows3* o) {   // joints.cc:13-35
  const Body& b0 = W.bodies[j.i0];
  Mat3 zero;
  put_block(o->j0, Mat3::identity(), cross_mat(b0.R * j.c0) * -1.0);
  if (j.i1 < 0) put_block(o->j1, zero, zero);
  else put_block(o->j1, Mat3::identity() * -1.0, cross_mat(W.bodies[j.i1].R * j.c1));
  for (int k = 0; k < 3; k++) { o->type[k] = 1; o->lo[k] = 0; o->hi[k] = 0; }
  o->i0 = j.i0; o->i1 = j.i1;
}
inline void contact_rows(const World& W, const Contact& c, Rows3* o) {   // contact.cc:38-117 (BOX)
  Mat3 R = align_vectors(c.cg.normal, Vec3(0, 0, 1));
  Mat3 zero;
  if (c.i0 < 0) put_block(o->j0, zero, zero);
  else put_block(o->j0, R * (Mat3::identity() * -1.0), R * cross_mat(c.cg.position - W.bodies[c.i0].p));
  if (c.i1 < 0) put_block(o->j1, zero, zero);
  else put_block(o->j1, R * Mat3::identity(), R * (cross_mat(c.cg.position - W.bodies[c.i1].p) * -1.0));
  const double inf = std::numeric_limits<double>::infinity();
  o->type[0] = o->type[1] = o->type[2] = 0;
  o->lo[0] = -1; o->lo[1] = -1; o->lo[2] = 0;      // kBoxFrictionBound = 1 (contact.cc:11)
  o->hi[0] = 1;  o->hi[1] = 1;  o->hi[2] = inf;
  o->i0 = c.i0; o->i1 = c.i1;
}
// ensembles.cc:234-239 + :38-87: joints (list order) then contacts.
inline void all_rows(const World& W, std::vector<Rows3>* rows) {
  rows->resize(W.joints.size() + W.contacts.size());
  size_t k = 0;
  for (const auto& j : W.joints) joint_rows(W, j, &(*rows)[k++]);
  for (const auto& c : W.contacts) contact_rows(W, c, &(*rows)[k++]);
}
// ensembles.cc:156-171
inline Vec position_error(const World& W) {
  Vec e;
  for (const auto& j : W.joints) { Vec3 v = joint_error(W, j); e.push_back(v.x); e.push_back(v.y); e.push_back(v.z); }
  for (const auto& c : W.contacts) { e.push_back(0); e.push_back(0); e.push_back(-c.cg.depth); }   // contact.cc:14-22
  return e;
}
inline Mat dense_J(const World& W, const std::vector<Rows3>& rows) {
  Mat J((int)rows.size() * 3, 6 * W.n);
  for (size_t k = 0; k < rows.size(); k++)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 6; c++) {
        if (rows[k].i0 >= 0) J((int)k * 3 + r, rows[k].i0 * 6 + c) = rows[k].j0[r][c];
        if (rows[k].i1 >= 0) J((int)k * 3 + r, rows[k].i1 * 6 + c) = rows[k].j1[r][c];
      }
  return J;
}
inline Mat dense_Minv(const World& W) {
  Mat M(6 * W.n, 6 * W.n);
  for (int i = 0; i < W.n; i++)
    for (int r = 0; r < 3; r++) {
      M(6 * i + r, 6 * i + r) = W.Minv_lin[i];
      for (int c = 0; c < 3; c++) M(6 * i + 3 + r, 6 * i + 3 + c) = W.Minv_ang[i].m[r][c];
    }
  return M;
}

// ---------------------------------------------------------------------------------------------
// Init (ensembles.cc:24-29, 202-232).

inline void check_and_correct(World& W);

inline int world_init(World& W) {
  W.n = (int)W.bodies.size();
  W.Minv_ang.resize(W.n);
  W.Minv_lin.resize(W.n);
  W.f_ext.assign(6 * W.n, 0.0);
  for (int i = 0; i < W.n; i++) {
    const Body& b = W.bodies[i];
    W.Minv_lin[i] = 1.0 / b.m;
    W.Minv_ang[i] = inverse3(b.I_g());
    for (int k = 0; k < 3; k++) W.f_ext[6 * i + k] = b.m * W.prm.gravity[k];
    Vec3 t = (cross_mat(b.w) * -1.0) * b.I_g() * b.w;       // ((-1*CrossMat(w)) * I_g) * w
    W.f_ext[6 * i + 3] = t.x; W.f_ext[6 * i + 4] = t.y; W.f_ext[6 * i + 5] = t.z;
  }
  W.stats = StepStats();
  Vec e = position_error(W);                                // CheckInitialConditions: isZero(1e-9)
  for (double v : e) if (std::fabs(v) > kAllowNumericalError) W.stats.status |= ST_BAD_INIT;
  check_and_correct(W);
  return W.stats.status;
}

// ---------------------------------------------------------------------------------------------
// Contacts (ensembles.cc:445-480) and de-duplication (ensembles.cc:241-388).

inline void update_contacts(World& W) {
  W.contacts.clear();
  W.pair_hit_i.clear(); W.pair_hit_j.clear(); W.pair_hit_code.clear(); W.pair_hit_count.clear();
  W.ground_count.assign(W.n, 0);
  std::vector<ContactGeometry> cgs;
  for (int i = 0; i < W.n; i++) {
    cgs.clear();
    const Body& b = W.bodies[i];
    if (b.shape == 0) collide_box_and_ground(b.p, b.R, b.side, &cgs);
    else collide_round_and_ground(b.shape, b.p, b.R, b.side, &cgs);
    W.ground_count[i] = (int)cgs.size();
    for (const auto& cg : cgs) { Contact c; c.i0 = -1; c.i1 = i; c.cg = cg; W.contacts.push_back(c); }
  }
  int tests = 0;
  for (int i = 0; i < W.n; i++)
    for (int j = i + 1; j < W.n; j++) {
      cgs.clear();
      CollisionInfo ci;
      Box b1{W.bodies[i].p, W.bodies[i].R, W.bodies[i].side * 0.5};
      Box b2{W.bodies[j].p, W.bodies[j].R, W.bodies[j].side * 0.5};
      tests++;
      const int s1 = W.bodies[i].shape, s2 = W.bodies[j].shape;
      bool hit = false;
      if (s1 == 0 && s2 == 0) hit = collide_boxes(b1, b2, &ci, &cgs);
      else if (s1 == 1 && s2 == 1) hit = collide_spheres(W.bodies[i].p, W.bodies[i].side[0], W.bodies[j].p, W.bodies[j].side[0], &ci, &cgs);
      // other shape pairs: no pairwise narrowphase (documented limitation)
      if (hit) {
        W.pair_hit_i.push_back(i); W.pair_hit_j.push_back(j);
        W.pair_hit_code.push_back(ci.code); W.pair_hit_count.push_back((int)cgs.size());
      }
      for (const auto& cg : cgs) { Contact c; c.i0 = i; c.i1 = j; c.cg = cg; c.code = ci.code; W.contacts.push_back(c); }
    }
  W.stats.n_pair_tests = tests;
  W.stats.n_pair_hits = (int)W.pair_hit_i.size();
  W.stats.n_contacts_raw = (int)W.contacts.size();
}

inline void check_and_correct(World& W) {
  const double dmin = W.prm.min_constraint_dist;
  const int nc = (int)W.contacts.size(), nj = (int)W.joints.size();
  std::vector<unsigned char> del(nc, 0);
  auto key_lo = [](int a, int b) { return a < b ? a : b; };
  auto key_hi = [](int a, int b) { return a < b ? b : a; };
  for (int i = 0; i + 1 < W.n; i++)
    for (int j = i + 1; j < W.n; j++) {
      std::vector<int> pj, pc;
      for (int k = 0; k < nj; k++)
        if (key_lo(W.joints[k].i0, W.joints[k].i1) == i && key_hi(W.joints[k].i0, W.joints[k].i1) == j) pj.push_back(k);
      for (int k = 0; k < nc; k++)
        if (key_lo(W.contacts[k].i0, W.contacts[k].i1) == i && key_hi(W.contacts[k].i0, W.contacts[k].i1) == j) pc.push_back(k);
      for (size_t a = 0; a < pj.size(); a++)
        for (size_t b = a + 1; b < pj.size(); b++)
          if (norm(joint_position(W, W.joints[pj[a]]) - joint_position(W, W.joints[pj[b]])) < dmin)
            W.stats.status |= ST_JOINT_CONFLICT;
      for (size_t a = 0; a < pj.size(); a++)
        for (size_t b = 0; b < pc.size(); b++)
          if (norm(joint_position(W, W.joints[pj[a]]) - W.contacts[pc[b]].cg.position) < dmin) del[pc[b]] = 1;
      for (size_t a = 0; a < pc.size(); a++)
        for (size_t b = a + 1; b < pc.size(); b++)
          if (norm(W.contacts[pc[a]].cg.position - W.contacts[pc[b]].cg.position) < dmin) del[pc[b]] = 1;
    }
  std::vector<Contact> keep;
  for (int k = 0; k < nc; k++) if (!del[k]) keep.push_back(W.contacts[k]);
  W.contacts.swap(keep);
  W.stats.n_contacts = (int)W.contacts.size();
}

// ---------------------------------------------------------------------------------------------
// Matrix-free block operators on A = J M^-1 J^T (sparse_iterations_utils.cc).  The reference
// recomputes every 3x3 block A_ij = (J_i,b * M^-1_b) * J_j,b^T on each visit; the oracle caches
// the blocks of one constraint list (identical arithmetic, evaluated once) and visits them in
// the reference's j order.

struct BlockSystem {
  int nc = 0;
  std::vector<Rows3> rows;
  // CSR over constraint pairs (i, j != i) that share a body, j ascending; 3x3 blocks.
  std::vector<int> start, col;
  std::vector<double> blk;      // 9 per entry
  std::vector<double> diag;     // 9 per constraint, no cfm
  std::vector<unsigned char> type;
  Vec lo, hi;
};

inline void jm_block(const World& W, const double ji[3][6], int body, const double jj[3][6], double out[3][3], bool accumulate) {
  // (ji * Minv_body) * jj^T with Minv_body = blockdiag(1/m I, Iinv)
  double t[3][6];
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) t[r][c] = ji[r][c] * W.Minv_lin[body];
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += ji[r][3 + k] * W.Minv_ang[body].m[k][c];
      t[r][3 + c] = s;
    }
  }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      double s = 0;
      for (int k = 0; k < 6; k++) s += t[r][k] * jj[c][k];
      out[r][c] = accumulate ? out[r][c] + s : s;
    }
}

inline void build_block_system(const World& W, BlockSystem* B) {
  all_rows(W, &B->rows);
  const int nc = B->nc = (int)B->rows.size();
  B->start.assign(nc + 1, 0);
  B->col.clear(); B->blk.clear();
  B->diag.assign((size_t)nc * 9, 0.0);
  B->type.resize(nc * 3); B->lo.resize(nc * 3); B->hi.resize(nc * 3);
  for (int i = 0; i < nc; i++) {
    const Rows3& ri = B->rows[i];
    for (int k = 0; k < 3; k++) { B->type[3 * i + k] = ri.type[k]; B->lo[3 * i + k] = ri.lo[k]; B->hi[3 * i + k] = ri.hi[k]; }
    double d[3][3] = {{0}};
    if (ri.i0 >= 0) jm_block(W, ri.j0, ri.i0, ri.j0, d, true);
    if (ri.i1 >= 0) jm_block(W, ri.j1, ri.i1, ri.j1, d, true);
    for (int k = 0; k < 9; k++) B->diag[(size_t)i * 9 + k] = d[k / 3][k % 3];
    for (int j = 0; j < nc; j++) {
      if (j == i) continue;
      const Rows3& rj = B->rows[j];
      // sparse_iterations_utils.cc:186-199: if / else-if per body slot of constraint i
      bool any = false;
      double s[3][3] = {{0}};
      if (ri.i0 >= 0 && ri.i0 == rj.i0) { jm_block(W, ri.j0, ri.i0, rj.j0, s, true); any = true; }
      else if (ri.i0 >= 0 && ri.i0 == rj.i1) { jm_block(W, ri.j0, ri.i0, rj.j1, s, true); any = true; }
      if (ri.i1 >= 0 && ri.i1 == rj.i0) { jm_block(W, ri.j1, ri.i1, rj.j0, s, true); any = true; }
      else if (ri.i1 >= 0 && ri.i1 == rj.i1) { jm_block(W, ri.j1, ri.i1, rj.j1, s, true); any = true; }
      if (!any) continue;
      B->col.push_back(j);
      for (int k = 0; k < 9; k++) B->blk.push_back(s[k / 3][k % 3]);
    }
    B->start[i + 1] = (int)B->col.size();
  }
}

// sparse_iterations_utils.cc:624-695 CalculateSparseJMJtX.
inline Vec sparse_Ax(const BlockSystem& B, const Vec& x, double cfm) {
  Vec y(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++) {
    int e = B.start[i];
    const int e_end = B.start[i + 1];
    bool diag_done = false;
    auto add_diag = [&]() {
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += (B.diag[(size_t)i * 9 + r * 3 + c] + (r == c ? cfm : 0.0)) * x[3 * i + c];
        y[3 * i + r] += s;
      }
      diag_done = true;
    };
    for (; e < e_end; e++) {
      int j = B.col[e];
      if (!diag_done && j > i) add_diag();
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += B.blk[(size_t)e * 9 + r * 3 + c] * x[3 * j + c];
        y[3 * i + r] += s;
      }
    }
    if (!diag_done) add_diag();
  }
  return y;
}
// sparse_iterations_utils.cc:495-561 CalculateSparseUx (strict upper of the diagonal block, then
// blocks j > i).
inline Vec sparse_Ux(const BlockSystem& B, const Vec& x) {
  Vec y(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++) {
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int c = r + 1; c < 3; c++) s += B.diag[(size_t)i * 9 + r * 3 + c] * x[3 * i + c];
      y[3 * i + r] += s;
    }
    for (int e = B.start[i]; e < B.start[i + 1]; e++) {
      int j = B.col[e];
      if (j < i) continue;
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += B.blk[(size_t)e * 9 + r * 3 + c] * x[3 * j + c];
        y[3 * i + r] += s;
      }
    }
  }
  return y;
}
// sparse_iterations_utils.cc:427-493 CalculateSparseLx.
inline Vec sparse_Lx(const BlockSystem& B, const Vec& x) {
  Vec y(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++) {
    for (int e = B.start[i]; e < B.start[i + 1]; e++) {
      int j = B.col[e];
      if (j > i) break;
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += B.blk[(size_t)e * 9 + r * 3 + c] * x[3 * j + c];
        y[3 * i + r] += s;
      }
    }
    for (int r = 0; r < 3; r++) {
      double s = 0;
      for (int c = 0; c < r; c++) s += B.diag[(size_t)i * 9 + r * 3 + c] * x[3 * i + c];
      y[3 * i + r] += s;
    }
  }
  return y;
}
// sparse_iterations_utils.cc:571-603 CalculateSparseDx: ((d + eps) * scale) .* x
inline Vec sparse_Dx(const BlockSystem& B, const Vec& x, double eps, double scale) {
  Vec y(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++)
    for (int r = 0; r < 3; r++) y[3 * i + r] = ((B.diag[(size_t)i * 9 + r * 4] + eps) * scale) * x[3 * i + r];
  return y;
}
inline double apply_projection(double x, bool C, double lo, double hi) {   // sparse_iterations_utils.cc:12-21
  if (!C) {
    if (x < lo) return lo;
    else if (x > hi) return hi;
  }
  return x;
}
// Which constraint's (type, lo, hi) the reference ends up projecting block i with (q2).
inline int bounds_source(int i, int nc, bool lower, bool quirk) {
  if (!quirk) return i;
  if (lower) return i > 0 ? i - 1 : i;          // sparse_iterations_utils.cc:169,180,229-235
  return i < nc - 1 ? i + 1 : i;                // :302,315,362-368
}
// sparse_iterations_utils.cc:159-243 MatrixSolveSparseLowerTriangle.
inline Vec sparse_solve_lower(const BlockSystem& B, const Vec& rhs, double eps, double scale, bool quirk) {
  Vec x(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++) {
    double sub[3] = {0, 0, 0};
    for (int e = B.start[i]; e < B.start[i + 1]; e++) {
      int j = B.col[e];
      if (j > i) break;
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += B.blk[(size_t)e * 9 + r * 3 + c] * x[3 * j + c];
        sub[r] += s;
      }
    }
    double d[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) d[r][c] = B.diag[(size_t)i * 9 + r * 3 + c] + (r == c ? eps : 0.0);
    for (int r = 0; r < 3; r++) d[r][r] = d[r][r] * scale;
    const int src = bounds_source(i, B.nc, true, quirk);
    for (int k = 0; k < 3; k++) {
      for (int l = 0; l < k; l++) sub[k] += d[k][l] * x[3 * i + l];
      x[3 * i + k] = apply_projection((rhs[3 * i + k] - sub[k]) / d[k][k], B.type[3 * src + k], B.lo[3 * src + k], B.hi[3 * src + k]);
    }
  }
  return x;
}
// sparse_iterations_utils.cc:292-373 MatrixSolveSparseUpperTriangle.
inline Vec sparse_solve_upper(const BlockSystem& B, const Vec& rhs, double eps, double scale, bool quirk) {
  Vec x(B.nc * 3, 0.0);
  for (int i = B.nc - 1; i >= 0; i--) {
    double sub[3] = {0, 0, 0};
    for (int e = B.start[i + 1] - 1; e >= B.start[i]; e--) {     // j descending, j > i
      int j = B.col[e];
      if (j < i) break;
      for (int r = 0; r < 3; r++) {
        double s = 0;
        for (int c = 0; c < 3; c++) s += B.blk[(size_t)e * 9 + r * 3 + c] * x[3 * j + c];
        sub[r] += s;
      }
    }
    double d[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) d[r][c] = B.diag[(size_t)i * 9 + r * 3 + c] + (r == c ? eps : 0.0);
    for (int r = 0; r < 3; r++) d[r][r] = d[r][r] * scale;
    const int src = bounds_source(i, B.nc, false, quirk);
    for (int k = 2; k >= 0; k--) {
      for (int l = k + 1; l < 3; l++) sub[k] += d[k][l] * x[3 * i + l];
      x[3 * i + k] = apply_projection((rhs[3 * i + k] - sub[k]) / d[k][k], B.type[3 * src + k], B.lo[3 * src + k], B.hi[3 * src + k]);
    }
  }
  return x;
}
// sparse_iterations_utils.cc:67-108 MatrixSolveSparseDiagonal (own bounds: no inner ComputeJ).
inline Vec sparse_solve_diag(const BlockSystem& B, const Vec& rhs, double eps, double scale) {
  Vec x(B.nc * 3, 0.0);
  for (int i = 0; i < B.nc; i++)
    for (int k = 0; k < 3; k++) {
      double d = (B.diag[(size_t)i * 9 + k * 4] + eps) * scale;
      x[3 * i + k] = apply_projection(1.0 / d * rhs[3 * i + k], B.type[3 * i + k], B.lo[3 * i + k], B.hi[3 * i + k]);
    }
  return x;
}

// sparse_iterations.cc:35-69 GetResidualError (both overloads share this tail).
inline double residual_error(const Vec& w, const Vec& x, const std::vector<unsigned char>& C, const Vec& lo, const Vec& hi) {
  double se = 0, s1 = 0, s2 = 0, s3 = 0;
  for (size_t i = 0; i < w.size(); i++) {
    if (C[i]) { se += w[i] * w[i]; continue; }
    if (x[i] == lo[i] && w[i] < 0) s1 += w[i] * w[i];
    if (x[i] == hi[i] && w[i] > 0) s2 += w[i] * w[i];
    if (x[i] > lo[i] && x[i] < hi[i]) s3 += w[i] * w[i];
  }
  return std::sqrt(se) + (std::sqrt(s1) + std::sqrt(s2) + std::sqrt(s3));
}

enum IterType { IT_JACOBI = 0, IT_GS = 1, IT_SOR = 2 };
constexpr double kSorOmega = 1.5;            // sparse_iterations.cc:15
constexpr double kSOR = 1 / kSorOmega;       // sparse_iterations.cc:16

// sparse_iterations.cc:148-226 BaseIteration(constraints, ...).
inline Vec sparse_iteration(const BlockSystem& B, const Vec& rhs, int type, double cfm, int k_max,
                            double tol, bool quirk, int* sweeps_out, double* resid_out) {
  Vec x = rhs;
  int it = 0;
  if (B.nc == 0) { if (sweeps_out) *sweeps_out = 0; if (resid_out) *resid_out = 0; return Vec(); }
  auto resid = [&](const Vec& xx) {
    Vec w = sparse_Ax(B, xx, cfm);
    for (size_t i = 0; i < w.size(); i++) w[i] -= rhs[i];
    return residual_error(w, xx, B.type, B.lo, B.hi);
  };
  double err = resid(x);
  while (err > tol && it < k_max) {
    Vec nx;
    if (type == IT_JACOBI) { Vec a = sparse_Lx(B, x), b = sparse_Ux(B, x); nx.resize(a.size()); for (size_t i = 0; i < a.size(); i++) nx[i] = a[i] + b[i]; }
    else if (type == IT_GS) nx = sparse_Ux(B, x);
    else { Vec a = sparse_Lx(B, x), b = sparse_Dx(B, x, cfm, 1 - kSOR); nx.resize(a.size()); for (size_t i = 0; i < a.size(); i++) nx[i] = a[i] + b[i]; }
    Vec irhs(nx.size());
    for (size_t i = 0; i < nx.size(); i++) irhs[i] = -1 * nx[i] + rhs[i];
    if (type == IT_JACOBI) x = sparse_solve_diag(B, irhs, cfm, 1.0);
    else if (type == IT_GS) x = sparse_solve_lower(B, irhs, cfm, 1.0, quirk);
    else x = sparse_solve_upper(B, irhs, cfm, kSOR, quirk);
    err = resid(x);
    ++it;
  }
  if (sweeps_out) *sweeps_out = it;
  if (resid_out) *resid_out = err;
  return x;
}

// Dense-matrix variants (sparse_iterations.cc:72-144, sparse_iterations_utils.cc:24-36,110-127,
// 245-262).  The reference's spectral-radius CHECK (sparse_iterations.cc:116-121) is a
// diagnostic that never alters x and is not restated.
inline Vec dense_iteration(const Mat& A, const Vec& b, int type, const std::vector<unsigned char>& C,
                           const Vec& lo, const Vec& hi, int k_max, double tol, int* sweeps_out) {
  const int dim = A.r;
  if (dim == 0) { if (sweeps_out) *sweeps_out = 0; return Vec(); }
  Mat M(dim, dim), N(dim, dim);
  for (int i = 0; i < dim; i++)
    for (int j = 0; j < dim; j++) {
      double a = A(i, j);
      if (type == IT_JACOBI) { if (i == j) M(i, j) = a; else N(i, j) = -1 * a; }
      else if (type == IT_GS) { if (j <= i) M(i, j) = a; else N(i, j) = -1 * a; }
      else {
        if (j > i) M(i, j) = a; else if (j == i) M(i, j) = kSOR * a;
        if (j < i) N(i, j) = -1 * a; else if (j == i) N(i, j) = (kSOR - 1) * a;
      }
    }
  auto resid = [&](const Vec& xx) {
    Vec w = matvec(A, xx);
    for (int i = 0; i < dim; i++) w[i] -= b[i];
    return residual_error(w, xx, C, lo, hi);
  };
  Vec x = b;
  int it = 0;
  double err = resid(x);
  while (err > tol && it < k_max) {
    Vec rhs = matvec(N, x);
    for (int i = 0; i < dim; i++) rhs[i] += b[i];
    Vec nx(dim, 0.0);
    if (type == IT_JACOBI) {
      for (int i = 0; i < dim; i++) nx[i] = apply_projection(1.0 / M(i, i) * rhs[i], C[i], lo[i], hi[i]);
    } else if (type == IT_GS) {
      for (int i = 0; i < dim; i++) {
        double s = 0;
        for (int j = 0; j < i; j++) s += M(i, j) * nx[j];
        nx[i] = apply_projection((rhs[i] - s) / M(i, i), C[i], lo[i], hi[i]);
      }
    } else {
      for (int i = dim - 1; i >= 0; i--) {
        double s = 0;
        for (int j = i + 1; j < dim; j++) s += M(i, j) * nx[j];
        nx[i] = apply_projection((rhs[i] - s) / M(i, i), C[i], lo[i], hi[i]);
      }
    }
    x = nx;
    err = resid(x);
    ++it;
  }
  if (sweeps_out) *sweeps_out = it;
  return x;
}

// ---------------------------------------------------------------------------------------------
// Step (ensembles.cc:390-427, 498-538, 563-591).

inline void world_step(World& W, double dt) {
  const int n = W.n;
  W.stats.status &= ~ST_LCP_FAILED;
  Vec v(6 * n);                                             // GetVelocities, ensembles.cc:429-436
  for (int i = 0; i < n; i++) {
    v[6 * i + 0] = W.bodies[i].v.x; v[6 * i + 1] = W.bodies[i].v.y; v[6 * i + 2] = W.bodies[i].v.z;
    v[6 * i + 3] = W.bodies[i].w.x; v[6 * i + 4] = W.bodies[i].w.y; v[6 * i + 5] = W.bodies[i].w.z;
  }
  update_contacts(W);
  check_and_correct(W);

  // StepVelocities_ODE (ensembles.cc:563-575)
  std::vector<Rows3> rows;
  all_rows(W, &rows);
  const int nr = (int)rows.size() * 3;
  W.stats.n_rows = nr;
  W.stats.sweeps = W.stats.pivots = W.stats.cfm_applied = 0;
  W.stats.residual = 0;
  Vec err = position_error(W);
  Vec u(6 * n);                                             // v/dt + M^-1 f
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) u[6 * i + k] = v[6 * i + k] / dt + W.Minv_lin[i] * W.f_ext[6 * i + k];
    Vec3 t = W.Minv_ang[i] * Vec3(W.f_ext[6 * i + 3], W.f_ext[6 * i + 4], W.f_ext[6 * i + 5]);
    for (int k = 0; k < 3; k++) u[6 * i + 3 + k] = v[6 * i + 3 + k] / dt + t[k];
  }
  Vec rhs(nr);
  for (size_t c = 0; c < rows.size(); c++)
    for (int r = 0; r < 3; r++) {
      double s = 0;
      if (rows[c].i0 >= 0) for (int k = 0; k < 6; k++) s += rows[c].j0[r][k] * u[6 * rows[c].i0 + k];
      if (rows[c].i1 >= 0) for (int k = 0; k < 6; k++) s += rows[c].j1[r][k] * u[6 * rows[c].i1 + k];
      rhs[3 * c + r] = -W.prm.erp / dt / dt * err[3 * c + r] - s;
    }
  W.rhs = rhs;

  // ComputeVDot (ensembles.cc:498-538)
  Vec lambda(nr, 0.0);
  W.row_state.assign(nr, 0);
  if (nr > 0) {
    if (W.prm.solver == SOLVER_DENSE_MURTY) {
      Mat J = dense_J(W, rows);
      Mat A = matmul(matmul(J, dense_Minv(W)), transpose(J));
      bool good;
      if (W.prm.cfm_mode == CFM_ALWAYS) good = false;
      else if (W.prm.cfm_mode == CFM_NEVER) good = true;
      else good = condition_number(A) < kGoodConditionNumber;      // utils.cc:273-287
      if (!good) { for (int i = 0; i < nr; i++) A(i, i) += W.prm.cfm; W.stats.cfm_applied = 1; }
      Mask C(nr);
      Vec lo(nr), hi(nr);
      for (size_t c = 0; c < rows.size(); c++)
        for (int r = 0; r < 3; r++) { C[3 * c + r] = rows[c].type[r]; lo[3 * c + r] = rows[c].lo[r]; hi[3 * c + r] = rows[c].hi[r]; }
      Vec wv;
      MurtyStats ms;
      bool ok = mixed_constraints_solver(A, rhs, C, lo, hi, lambda, wv, !(W.prm.quirks & QUIRK_DENSE_IGNORES_BOUNDS), &ms);
      if (!ok) W.stats.status |= ST_LCP_FAILED;
      W.stats.pivots = ms.iterations;
      int k = 0;
      for (int i = 0; i < nr; i++) W.row_state[i] = C[i] ? 3 : (ms.S[k++] ? 0 : 1);
    } else {
      BlockSystem B;
      build_block_system(W, &B);
      int type = W.prm.solver == SOLVER_PGS ? IT_GS : (W.prm.solver == SOLVER_JACOBI ? IT_JACOBI : IT_SOR);
      // The matrix-free path always carries cfm on the diagonal (it has no condition test).
      lambda = sparse_iteration(B, rhs, type, W.prm.cfm, W.prm.k_max, W.prm.tol,
                                (W.prm.quirks & QUIRK_GS_BOUNDS_SHIFT) != 0, &W.stats.sweeps, &W.stats.residual);
      W.stats.cfm_applied = W.prm.cfm != 0;
      for (int i = 0; i < nr; i++) {
        if (B.type[i]) W.row_state[i] = 3;
        else if (lambda[i] == B.lo[i]) W.row_state[i] = 1;
        else if (lambda[i] == B.hi[i]) W.row_state[i] = 2;
        else W.row_state[i] = 0;
      }
    }
  }
  W.lambda = lambda;

  // v_dot = M^-1 (f + J^T lambda); v_new = v + dt v_dot  (ensembles.cc:535, 572)
  Vec g = W.f_ext;
  for (size_t c = 0; c < rows.size(); c++)
    for (int r = 0; r < 3; r++) {
      double l = lambda[3 * c + r];
      if (rows[c].i0 >= 0) for (int k = 0; k < 6; k++) g[6 * rows[c].i0 + k] += rows[c].j0[r][k] * l;
      if (rows[c].i1 >= 0) for (int k = 0; k < 6; k++) g[6 * rows[c].i1 + k] += rows[c].j1[r][k] * l;
    }
  Vec vn(6 * n);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) vn[6 * i + k] = v[6 * i + k] + dt * (W.Minv_lin[i] * g[6 * i + k]);
    Vec3 t = W.Minv_ang[i] * Vec3(g[6 * i + 3], g[6 * i + 4], g[6 * i + 5]);
    for (int k = 0; k < 3; k++) vn[6 * i + 3 + k] = v[6 * i + 3 + k] + dt * t[k];
  }
  // StepPositions_ODE (ensembles.cc:577-591)
  for (int i = 0; i < n; i++) {
    Body& b = W.bodies[i];
    b.v = Vec3(vn[6 * i], vn[6 * i + 1], vn[6 * i + 2]);
    b.w = Vec3(vn[6 * i + 3], vn[6 * i + 4], vn[6 * i + 5]);
    Vec3 vmid = (Vec3(v[6 * i], v[6 * i + 1], v[6 * i + 2]) + b.v) / 2.0;
    b.p = b.p + dt * vmid;
    Vec3 wmid = (Vec3(v[6 * i + 3], v[6 * i + 4], v[6 * i + 5]) + b.w) / 2.0;
    b.R = quat_to_mat(w_to_q(wmid, dt)) * b.R;
  }
}

// ---------------------------------------------------------------------------------------------
// Stabilisation (ensembles.cc:602-666).

inline Vec velocity_relaxation(const World& W, double step_scale) {   // ensembles.cc:659-666
  std::vector<Rows3> rows;
  all_rows(W, &rows);
  Mat J = dense_J(W, rows);
  Vec err = position_error(W);
  if (J.r == 0) return Vec(6 * W.n, 0.0);
  LDLT f;
  f.compute(matmul(J, transpose(J)));
  Vec y = f.solve(err);
  Vec out(6 * W.n, 0.0);
  for (int i = 0; i < J.r; i++)
    for (int c = 0; c < J.c; c++) out[c] += (-1.0 * step_scale * J(i, c)) * y[i];
  return out;
}
inline void step_positions_explicit_euler(World& W, double dt, const Vec& v) {   // ensembles.cc:553-561
  for (int i = 0; i < W.n; i++) {
    Body& b = W.bodies[i];
    b.p = b.p + dt * Vec3(v[6 * i], v[6 * i + 1], v[6 * i + 2]);
    b.R = quat_to_mat(w_to_q(Vec3(v[6 * i + 3], v[6 * i + 4], v[6 * i + 5]), dt)) * b.R;
  }
}
inline int init_stabilize(World& W, double* final_err_sq, int max_steps = 100) {   // ensembles.cc:602-622
  update_contacts(W);
  Vec err = position_error(W);
  double e2 = 0;
  for (double x : err) e2 += x * x;
  int steps = 0;
  while (e2 > kAllowNumericalError && steps < max_steps) {
    step_positions_explicit_euler(W, 0.001 * 500, velocity_relaxation(W, 0.2));
    update_contacts(W);
    err = position_error(W);
    e2 = 0;
    for (double x : err) e2 += x * x;
    ++steps;
  }
  check_and_correct(W);
  if (final_err_sq) *final_err_sq = e2;
  return steps;
}

// ensembles.cc:624-645 PostStabilize + :652-657 StepPostStabilization(dt = kSimTimeStep * 100):
// positions by explicit Euler with the relaxation "velocity", then v += relaxation.  The contact
// list is whatever the last UpdateContacts / Step left (it is not refreshed in the loop, so a
// contact's error stays its stored depth).
inline int post_stabilize(World& W, double* final_err_sq, int max_steps = 500) {
  Vec err = position_error(W);
  double e2 = 0;
  for (double x : err) e2 += x * x;
  int steps = 0;
  while (e2 > kAllowNumericalError && steps < max_steps) {
    Vec c = velocity_relaxation(W, 0.2);
    step_positions_explicit_euler(W, 0.001 * 100, c);
    for (int i = 0; i < W.n; i++) {                       // UpdateComponentsVelocities(v + c)
      Body& b = W.bodies[i];
      b.v = Vec3(b.v.x + c[6 * i], b.v.y + c[6 * i + 1], b.v.z + c[6 * i + 2]);
      b.w = Vec3(b.w.x + c[6 * i + 3], b.w.y + c[6 * i + 4], b.w.z + c[6 * i + 5]);
    }
    err = position_error(W);
    e2 = 0;
    for (double x : err) e2 += x * x;
    ++steps;
  }
  if (final_err_sq) *final_err_sq = e2;
  return steps;
}

// ---------------------------------------------------------------------------------------------
// Built-in scenes (ensembles.cc:668-728).

inline void build_chain(World& W, int num_links, const Vec3& anchor) {
  W.bodies.clear(); W.joints.clear(); W.contacts.clear();
  Quat q = quat_mul(angle_axis_to_quat(0.95531661812451, Vec3(0, 0, 1)), angle_axis_to_quat(M_PI / 4, Vec3(1, 0, 0)));
  Mat3 R = quat_to_mat(q);
  for (int i = 0; i < num_links; i++) {
    Body b;
    b.p = Vec3(std::sqrt(3.0) * 0.3 * i, 0, 0) + anchor;
    b.R = R;
    b.m = 1.0;
    b.I = box_inertia(b.m, b.side);
    W.bodies.push_back(b);
  }
  for (int i = 0; i < num_links - 1; i++) {
    Joint j; j.i0 = i; j.i1 = i + 1; j.c0 = Vec3(0.15, -0.15, 0.15); j.c1 = Vec3(-0.15, 0.15, -0.15);
    W.joints.push_back(j);
  }
  Joint a; a.i0 = 0; a.i1 = -1; a.c0 = Vec3(0, 0, 0); a.c1 = W.bodies[0].p;
  W.joints.push_back(a);
  W.n = num_links;
}
// Eigen's Random()/UnitRandom() draw from std::rand(); the reference never seeds it.
inline double eigen_random(double lo, double hi) { return lo + (hi - lo) * double(std::rand()) / double(RAND_MAX); }
inline void build_cairn(World& W, int num_rocks, const double xb[2], const double yb[2], const double zb[2]) {
  W.bodies.clear(); W.joints.clear(); W.contacts.clear();
  for (int i = 0; i < num_rocks; i++) {
    Body b;
    double r0 = eigen_random(-1, 1), r1 = eigen_random(-1, 1), r2 = eigen_random(-1, 1);
    Vec3 u((r0 + 1) / 2, (r1 + 1) / 2, (r2 + 1) / 2);        // utils.cc:26-39
    b.p = Vec3(u.x * std::fabs(xb[1] - xb[0]) + std::min(xb[0], xb[1]),
               u.y * std::fabs(yb[1] - yb[0]) + std::min(yb[0], yb[1]),
               u.z * std::fabs(zb[1] - zb[0]) + std::min(zb[0], zb[1]));
    double u1 = eigen_random(0, 1), u2 = eigen_random(0, 2 * M_PI), u3 = eigen_random(0, 2 * M_PI);
    double a = std::sqrt(1 - u1), bb = std::sqrt(u1);
    Quat q; q.w = a * std::sin(u2); q.x = a * std::cos(u2); q.y = bb * std::sin(u3); q.z = bb * std::cos(u3);
    b.R = quat_to_mat(q);
    b.v = Vec3(eigen_random(-1, 1), eigen_random(-1, 1), eigen_random(-1, 1)) * 1.0;
    b.w = Vec3(eigen_random(-1, 1), eigen_random(-1, 1), eigen_random(-1, 1)) * 1.0;
    b.m = 1.0;
    b.I = Mat3::identity() * 0.1;
    W.bodies.push_back(b);
  }
  W.n = num_rocks;
}

}  // namespace orc
