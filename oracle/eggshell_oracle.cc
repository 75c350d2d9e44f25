// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C ABI (for ctypes) over the Eigen-free CPU restatement of eggshell's step path
// (/root/reference/eggshell/{collision,contact,joints,ensembles,lcp,sparse_iterations,
// sparse_iterations_utils,utils,body}.cc; per-function file:line citations are in the orc_*.h
// headers).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library; the product (eggshell_b200/) never links or imports it.
//
// Parity pins: lcp.cc:348-389 literal 5x5 LCP vectors, utils.cc:398-497 integer gather/scatter
// vectors, and the property tests of collision.cc:527-809, lcp.cc:412-528,
// sparse_iterations.cc:355-748, sparse_iterations_utils.cc:938-1248 restated in tests/.
// PARITY UNPINNED for Eigen 3.3.8 internals the reference has no test for (see orc_linalg.h).
#include <thread>
#include <chrono>
#include "orc_world.h"

using namespace orc;

static Vec3 v3(const double* p) { return Vec3(p[0], p[1], p[2]); }
static Mat3 m3(const double* p) {
  Mat3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = p[i * 3 + j];
  return r;
}
static void put3(double* o, const Vec3& v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static void putm(double* o, const Mat3& m) {
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) o[i * 3 + j] = m.m[i][j];
}
static Mat mat_from(const double* a, int r, int c) {
  Mat m(r, c);
  std::copy(a, a + (size_t)r * c, m.a.begin());
  return m;
}
static Mask mask_from(const unsigned char* a, int n) { return Mask(a, a + n); }

extern "C" {

// ---- geometry -------------------------------------------------------------------------------
int orc_collide_box_ground(const double* c, const double* R, const double* side, double* out7x8) {
  std::vector<ContactGeometry> cg;
  collide_box_and_ground(v3(c), m3(R), v3(side), &cg);
  for (size_t i = 0; i < cg.size(); i++) {
    put3(out7x8 + 7 * i, cg[i].position);
    put3(out7x8 + 7 * i + 3, cg[i].normal);
    out7x8[7 * i + 6] = cg[i].depth;
  }
  return (int)cg.size();
}
// halfsides are passed directly (the Box overload, collision.cc:166).
int orc_collide_boxes(const double* c1, const double* R1, const double* h1, const double* c2,
                      const double* R2, const double* h2, double* info4, int* code, double* out7xN,
                      int max_out) {
  Box b1{v3(c1), m3(R1), v3(h1)}, b2{v3(c2), m3(R2), v3(h2)};
  CollisionInfo ci;
  std::vector<ContactGeometry> cg;
  bool hit = collide_boxes(b1, b2, &ci, &cg);
  info4[0] = ci.depth;
  put3(info4 + 1, ci.separating_axis);
  *code = hit ? ci.code : 0;
  for (size_t i = 0; i < cg.size() && (int)i < max_out; i++) {
    put3(out7xN + 7 * i, cg[i].position);
    put3(out7xN + 7 * i + 3, cg[i].normal);
    out7xN[7 * i + 6] = cg[i].depth;
  }
  return hit ? (int)cg.size() : 0;
}
int orc_boxes_separated(const double* c1, const double* R1, const double* h1, const double* c2,
                        const double* R2, const double* h2) {
  Box b1{v3(c1), m3(R1), v3(h1)}, b2{v3(c2), m3(R2), v3(h2)};
  return boxes_separated(b1, b2) ? 1 : 0;
}
void orc_line_closest_approach(const double* pa, const double* ua, const double* pb, const double* ub,
                               double* alpha, double* beta) {
  line_closest_approach(v3(pa), v3(ua), v3(pb), v3(ub), alpha, beta);
}
int orc_intersect_line_segment_and_line(const double* p1, const double* p2, const double* n, double d, double* p) {
  Vec2 o;
  bool r = intersect_line_segment_and_line(Vec2(p1[0], p1[1]), Vec2(p2[0], p2[1]), Vec2(n[0], n[1]), d, &o);
  p[0] = o.x; p[1] = o.y;
  return r;
}
int orc_clip_polygon(const double* poly, int np, const double* n, double d, double* out, int max_out) {
  std::vector<Vec2> in, o;
  for (int i = 0; i < np; i++) in.push_back(Vec2(poly[2 * i], poly[2 * i + 1]));
  clip_polygon_by_half_space(in, Vec2(n[0], n[1]), d, &o);
  for (size_t i = 0; i < o.size() && (int)i < max_out; i++) { out[2 * i] = o[i].x; out[2 * i + 1] = o[i].y; }
  return (int)o.size();
}
int orc_intersect_box_rect(const double* cB, const double* RB, const double* hB, const double* cR,
                           const double* RR, const double* hR, double* out, int max_out) {
  Box B{v3(cB), m3(RB), v3(hB)}, R{v3(cR), m3(RR), v3(hR)};
  std::vector<Vec2> o;
  intersect_box_and_rectangle(B, R, &o);
  for (size_t i = 0; i < o.size() && (int)i < max_out; i++) { out[2 * i] = o[i].x; out[2 * i + 1] = o[i].y; }
  return (int)o.size();
}

// ---- utils ----------------------------------------------------------------------------------
void orc_cross_mat(const double* a, double* out9) { putm(out9, cross_mat(v3(a))); }
void orc_w_to_q_matrix(const double* w, double dt, double* out9) { putm(out9, quat_to_mat(w_to_q(v3(w), dt))); }
void orc_align_vectors(const double* a, const double* b, double* out9) { putm(out9, align_vectors(v3(a), v3(b))); }
void orc_select_submatrix(const double* A, int n, const unsigned char* ri, const unsigned char* ci, double* out) {
  Mat S = select_submatrix(mat_from(A, n, n), mask_from(ri, n), mask_from(ci, n));
  std::copy(S.a.begin(), S.a.end(), out);
}
void orc_update_submatrix(double* A, int n, const unsigned char* ri, const unsigned char* ci, const double* m, int mr, int mc) {
  Mat M = mat_from(A, n, n);
  update_submatrix(M, mask_from(ri, n), mask_from(ci, n), mat_from(m, mr, mc));
  std::copy(M.a.begin(), M.a.end(), A);
}
int orc_select_subvector(const double* v, int n, const unsigned char* ind, double* out) {
  Vec s = select_subvector(Vec(v, v + n), mask_from(ind, n));
  std::copy(s.begin(), s.end(), out);
  return (int)s.size();
}
void orc_update_subvector(double* v, int n, const unsigned char* ind, const double* nv, int nn) {
  Vec x(v, v + n);
  update_subvector(x, mask_from(ind, n), Vec(nv, nv + nn));
  std::copy(x.begin(), x.end(), v);
}
void orc_update_subvector_scalar(double* v, int n, const unsigned char* ind, double d) {
  Vec x(v, v + n);
  update_subvector(x, mask_from(ind, n), d);
  std::copy(x.begin(), x.end(), v);
}
void orc_ldlt_solve(const double* A, int n, const double* b, double* x) {
  LDLT f;
  f.compute(mat_from(A, n, n));
  Vec s = f.solve(Vec(b, b + n));
  std::copy(s.begin(), s.end(), x);
}
void orc_lu_inverse(const double* A, int n, double* out) {
  Mat I = lu_inverse(mat_from(A, n, n));
  std::copy(I.a.begin(), I.a.end(), out);
}
double orc_condition_number(const double* A, int r, int c) { return condition_number(mat_from(A, r, c)); }

// ---- dense LCP ------------------------------------------------------------------------------
int orc_check_murty_solution(const double* A, const double* b, const double* x, const double* w, int n,
                             unsigned char* S, double err) {
  Mask s = mask_from(S, n);
  Vec lo(n, 0.0), hi(n, std::numeric_limits<double>::infinity()), Cx(n, 0.0);
  bool r = check_murty_solution(mat_from(A, n, n), Vec(b, b + n), Vec(x, x + n), Vec(w, w + n), s, Cx, lo, hi, err);
  std::copy(s.begin(), s.end(), S);
  return r;
}
int orc_murty(const double* A, const double* b, int n, const double* lo, const double* hi, double* x,
              double* w, int* iters, unsigned char* S) {
  Vec xx, ww;
  MurtyStats st;
  bool ok = murty_principal_pivot(mat_from(A, n, n), Vec(b, b + n), xx, ww, Vec(lo, lo + n), Vec(hi, hi + n), &st);
  std::copy(xx.begin(), xx.end(), x);
  std::copy(ww.begin(), ww.end(), w);
  if (iters) *iters = st.iterations;
  if (S) std::copy(st.S.begin(), st.S.end(), S);
  return ok;
}
int orc_mixed_solver(const double* A, const double* b, int n, const unsigned char* C, const double* lo,
                     const double* hi, int honour_bounds, double* x, double* w, int* iters) {
  Vec xx, ww;
  MurtyStats st;
  bool ok = mixed_constraints_solver(mat_from(A, n, n), Vec(b, b + n), mask_from(C, n), Vec(lo, lo + n),
                                     Vec(hi, hi + n), xx, ww, honour_bounds != 0, &st);
  std::copy(xx.begin(), xx.end(), x);
  std::copy(ww.begin(), ww.end(), w);
  if (iters) *iters = st.iterations;
  return ok;
}
int orc_dense_iteration(const double* A, const double* b, int n, int type, const unsigned char* C,
                        const double* lo, const double* hi, int k_max, double tol, double* x) {
  int sweeps = 0;
  Vec r = dense_iteration(mat_from(A, n, n), Vec(b, b + n), type, std::vector<unsigned char>(C, C + n),
                          Vec(lo, lo + n), Vec(hi, hi + n), k_max, tol, &sweeps);
  std::copy(r.begin(), r.end(), x);
  return sweeps;
}

// ---- worlds ---------------------------------------------------------------------------------
void* orc_world_create() { return new World(); }
void orc_world_destroy(void* w) { delete (World*)w; }
void orc_world_set_params(void* wp, double erp, double cfm, double min_dist, double tol, int k_max,
                          const double* gravity, int solver, int cfm_mode, int quirks) {
  World& W = *(World*)wp;
  W.prm.erp = erp; W.prm.cfm = cfm; W.prm.min_constraint_dist = min_dist; W.prm.tol = tol;
  W.prm.k_max = k_max; W.prm.solver = solver; W.prm.cfm_mode = cfm_mode; W.prm.quirks = quirks;
  for (int k = 0; k < 3; k++) W.prm.gravity[k] = gravity[k];
}
// Arrays are [n][3] / [n][9] row-major per body.
void orc_world_set_bodies(void* wp, int n, const double* p, const double* R, const double* v,
                          const double* w, const double* m, const double* I, const double* side) {
  World& W = *(World*)wp;
  W.bodies.resize(n);
  W.n = n;
  for (int i = 0; i < n; i++) {
    Body& b = W.bodies[i];
    b.p = v3(p + 3 * i); b.R = m3(R + 9 * i); b.v = v3(v + 3 * i); b.w = v3(w + 3 * i);
    b.m = m[i]; b.I = m3(I + 9 * i);
    if (side) b.side = v3(side + 3 * i);
  }
}
void orc_world_set_shapes(void* wp, const int* shape, const double* dims) {
  World& W = *(World*)wp;
  for (int i = 0; i < W.n; i++) { W.bodies[i].shape = shape[i]; W.bodies[i].side = v3(dims + 3 * i); }
}
void orc_world_set_state(void* wp, const double* p, const double* R, const double* v, const double* w) {
  World& W = *(World*)wp;
  for (int i = 0; i < W.n; i++) {
    Body& b = W.bodies[i];
    b.p = v3(p + 3 * i); b.R = m3(R + 9 * i); b.v = v3(v + 3 * i); b.w = v3(w + 3 * i);
  }
}
void orc_world_set_joints(void* wp, int nj, const int* i0, const int* i1, const double* c0, const double* c1) {
  World& W = *(World*)wp;
  W.joints.resize(nj);
  for (int k = 0; k < nj; k++) { W.joints[k].i0 = i0[k]; W.joints[k].i1 = i1[k]; W.joints[k].c0 = v3(c0 + 3 * k); W.joints[k].c1 = v3(c1 + 3 * k); }
}
// f_ext override (6 per body) after init; the reference's external_force_torque_ slot.
void orc_world_set_fext(void* wp, const double* f) { World& W = *(World*)wp; W.f_ext.assign(f, f + 6 * W.n); }
void orc_world_build_chain(void* wp, int links, const double* anchor) { build_chain(*(World*)wp, links, v3(anchor)); }
void orc_world_build_cairn(void* wp, int rocks, const double* xb, const double* yb, const double* zb) { build_cairn(*(World*)wp, rocks, xb, yb, zb); }
int orc_world_init(void* wp) { return world_init(*(World*)wp); }
int orc_world_init_stabilize(void* wp, double* final_err_sq) { return init_stabilize(*(World*)wp, final_err_sq); }
int orc_world_init_stabilize_n(void* wp, int max_steps, double* final_err_sq) { return init_stabilize(*(World*)wp, final_err_sq, max_steps); }
int orc_world_post_stabilize(void* wp, int max_steps, double* final_err_sq) { return post_stabilize(*(World*)wp, final_err_sq, max_steps); }
int orc_world_step(void* wp, double dt) { World& W = *(World*)wp; world_step(W, dt); return W.stats.status; }
int orc_world_n(void* wp) { return ((World*)wp)->n; }
int orc_world_n_joints(void* wp) { return (int)((World*)wp)->joints.size(); }
int orc_world_n_contacts(void* wp) { return (int)((World*)wp)->contacts.size(); }
void orc_world_update_contacts(void* wp, int dedupe) { World& W = *(World*)wp; update_contacts(W); if (dedupe) check_and_correct(W); }
void orc_world_get_bodies(void* wp, double* p, double* R, double* v, double* w) {
  World& W = *(World*)wp;
  for (int i = 0; i < W.n; i++) {
    put3(p + 3 * i, W.bodies[i].p); putm(R + 9 * i, W.bodies[i].R); put3(v + 3 * i, W.bodies[i].v); put3(w + 3 * i, W.bodies[i].w);
  }
}
void orc_world_get_static(void* wp, double* m, double* I, double* minv_lin, double* minv_ang, double* fext) {
  World& W = *(World*)wp;
  for (int i = 0; i < W.n; i++) {
    if (m) m[i] = W.bodies[i].m;
    if (I) putm(I + 9 * i, W.bodies[i].I);
    if (minv_lin) minv_lin[i] = W.Minv_lin[i];
    if (minv_ang) putm(minv_ang + 9 * i, W.Minv_ang[i]);
  }
  if (fext) std::copy(W.f_ext.begin(), W.f_ext.end(), fext);
}
void orc_world_get_joints(void* wp, int* i0, int* i1, double* c0, double* c1) {
  World& W = *(World*)wp;
  for (size_t k = 0; k < W.joints.size(); k++) { i0[k] = W.joints[k].i0; i1[k] = W.joints[k].i1; put3(c0 + 3 * k, W.joints[k].c0); put3(c1 + 3 * k, W.joints[k].c1); }
}
void orc_world_get_contacts(void* wp, int* i0, int* i1, double* pos, double* nrm, double* depth, int* code) {
  World& W = *(World*)wp;
  for (size_t k = 0; k < W.contacts.size(); k++) {
    const Contact& c = W.contacts[k];
    i0[k] = c.i0; i1[k] = c.i1; put3(pos + 3 * k, c.cg.position); put3(nrm + 3 * k, c.cg.normal); depth[k] = c.cg.depth; code[k] = c.code;
  }
}
int orc_world_get_pair_hits(void* wp, int* pi, int* pj, int* code, int* count, int max_out) {
  World& W = *(World*)wp;
  int n = (int)W.pair_hit_i.size();
  for (int k = 0; k < n && k < max_out; k++) { pi[k] = W.pair_hit_i[k]; pj[k] = W.pair_hit_j[k]; code[k] = W.pair_hit_code[k]; count[k] = W.pair_hit_count[k]; }
  return n;
}
void orc_world_get_ground_counts(void* wp, int* out) { World& W = *(World*)wp; std::copy(W.ground_count.begin(), W.ground_count.end(), out); }
// stats: [n_contacts_raw, n_contacts, n_rows, n_pair_tests, n_pair_hits, sweeps, pivots, cfm_applied, status]
void orc_world_get_stats(void* wp, int* out9, double* residual) {
  const StepStats& s = ((World*)wp)->stats;
  out9[0] = s.n_contacts_raw; out9[1] = s.n_contacts; out9[2] = s.n_rows; out9[3] = s.n_pair_tests; out9[4] = s.n_pair_hits;
  out9[5] = s.sweeps; out9[6] = s.pivots; out9[7] = s.cfm_applied; out9[8] = s.status;
  if (residual) *residual = s.residual;
}
int orc_world_get_solution(void* wp, double* lambda, double* rhs, int* row_state) {
  World& W = *(World*)wp;
  int nr = (int)W.lambda.size();
  if (lambda) std::copy(W.lambda.begin(), W.lambda.end(), lambda);
  if (rhs) std::copy(W.rhs.begin(), W.rhs.end(), rhs);
  if (row_state) std::copy(W.row_state.begin(), W.row_state.end(), row_state);
  return nr;
}
// Rows of the current constraint list (joints then contacts): J0,J1 [nc][3][6], type/lo/hi [3nc], i0/i1 [nc], err [3nc].
int orc_world_get_rows(void* wp, double* J0, double* J1, unsigned char* type, double* lo, double* hi, int* i0, int* i1, double* err) {
  World& W = *(World*)wp;
  std::vector<Rows3> rows;
  all_rows(W, &rows);
  Vec e = position_error(W);
  for (size_t c = 0; c < rows.size(); c++) {
    for (int r = 0; r < 3; r++) {
      for (int k = 0; k < 6; k++) { J0[(c * 3 + r) * 6 + k] = rows[c].j0[r][k]; J1[(c * 3 + r) * 6 + k] = rows[c].j1[r][k]; }
      type[3 * c + r] = rows[c].type[r]; lo[3 * c + r] = rows[c].lo[r]; hi[3 * c + r] = rows[c].hi[r];
      if (err) err[3 * c + r] = e[3 * c + r];
    }
    i0[c] = rows[c].i0; i1[c] = rows[c].i1;
  }
  return (int)rows.size();
}
// Dense A = J M^-1 J^T + cfm I of the current constraint list (rows x rows, row-major).
int orc_world_dense_A(void* wp, double cfm, double* out) {
  World& W = *(World*)wp;
  std::vector<Rows3> rows;
  all_rows(W, &rows);
  Mat J = dense_J(W, rows);
  Mat A = matmul(matmul(J, dense_Minv(W)), transpose(J));
  for (int i = 0; i < A.r; i++) A(i, i) += cfm;
  if (out) std::copy(A.a.begin(), A.a.end(), out);
  return A.r;
}
// Matrix-free products: op 0 = A x (cfm on diag), 1 = Ux, 2 = Lx, 3 = Dx(eps=cfm, scale), 4 = LxUx.
void orc_world_sparse_product(void* wp, int op, const double* x, double cfm, double scale, double* out) {
  World& W = *(World*)wp;
  BlockSystem B;
  build_block_system(W, &B);
  Vec xx(x, x + 3 * B.nc), y;
  if (op == 0) y = sparse_Ax(B, xx, cfm);
  else if (op == 1) y = sparse_Ux(B, xx);
  else if (op == 2) y = sparse_Lx(B, xx);
  else if (op == 3) y = sparse_Dx(B, xx, cfm, scale);
  else { Vec a = sparse_Lx(B, xx), b = sparse_Ux(B, xx); y.resize(a.size()); for (size_t i = 0; i < a.size(); i++) y[i] = a[i] + b[i]; }
  std::copy(y.begin(), y.end(), out);
}
// Matrix-free triangular / diagonal solves: which 0 = lower, 1 = upper, 2 = diagonal.
void orc_world_sparse_solve(void* wp, int which, const double* rhs, double cfm, double scale, int quirk, double* out) {
  World& W = *(World*)wp;
  BlockSystem B;
  build_block_system(W, &B);
  Vec r(rhs, rhs + 3 * B.nc), y;
  if (which == 0) y = sparse_solve_lower(B, r, cfm, scale, quirk != 0);
  else if (which == 1) y = sparse_solve_upper(B, r, cfm, scale, quirk != 0);
  else y = sparse_solve_diag(B, r, cfm, scale);
  std::copy(y.begin(), y.end(), out);
}
// sparse::{Jacobi,GaussSeidel,SOR}Iteration(constraints, M_inverse, rhs, cfm): type 0/1/2.
int orc_world_sparse_iteration(void* wp, int type, const double* rhs, double cfm, int k_max, double tol, int quirk, double* x) {
  World& W = *(World*)wp;
  BlockSystem B;
  build_block_system(W, &B);
  int sweeps = 0;
  Vec r = sparse_iteration(B, Vec(rhs, rhs + 3 * B.nc), type, cfm, k_max, tol, quirk != 0, &sweeps, nullptr);
  std::copy(r.begin(), r.end(), x);
  return sweeps;
}

// CPU baseline: step `count` worlds `steps` times each, worlds split over `nthreads` std::threads.
// Returns wall seconds.  totals[0] += sum over world-steps of rows*sweeps, totals[1] += sum of rows.
double orc_batch_step(void** worlds, int count, double dt, int steps, int nthreads, double* totals) {
  if (nthreads < 1) nthreads = 1;
  std::vector<double> rs((size_t)nthreads, 0.0), rr((size_t)nthreads, 0.0);
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](int t) {
    for (int k = t; k < count; k += nthreads) {
      World& W = *(World*)worlds[k];
      for (int s = 0; s < steps; s++) {
        world_step(W, dt);
        rs[t] += (double)W.stats.n_rows * (double)W.stats.sweeps;
        rr[t] += (double)W.stats.n_rows;
      }
    }
  };
  if (nthreads == 1) work(0);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
    for (auto& t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  if (totals) for (int t = 0; t < nthreads; t++) { totals[0] += rs[t]; totals[1] += rr[t]; }
  return std::chrono::duration<double>(t1 - t0).count();
}
int orc_hardware_concurrency() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
