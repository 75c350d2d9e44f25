// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of eggshell's dense mixed-LCP path: /root/reference/eggshell/lcp.cc:20-336.
#pragma once
#include "orc_linalg.h"

namespace orc {

constexpr double kAllowNumericalError = 1e-9;       // constants.h:5
constexpr double kLcpLooserAllowedError = 1e-8;     // lcp.cc:16
constexpr double kGoodConditionNumber = 1e7;        // constants.h:12

struct MurtyStats {
  int iterations = 0;       // value of `iter` when the loop exits
  int hit_cap = 0;          // iter >= max_iterations
  Mask S;                   // active-set mask at return (lcp.cc:176)
};

// lcp.cc:20-94.  May flip exactly one entry of S (least index first) and returns false, or runs
// the goodness checks.  Cx holds, for i not in S, which bound x(i) sits at.
inline bool check_murty_solution(const Mat& A, const Vec& b, const Vec& x, const Vec& w, Mask& S,
                                 Vec& Cx, const Vec& lo, const Vec& hi, double err = 0) {
  const double check_err = std::fabs(err) > kAllowNumericalError ? std::fabs(err) : kAllowNumericalError;
  const int dim = (int)S.size();
  for (int i = 0; i < dim; i++) {
    if (S[i]) {
      if (x[i] < lo[i]) { S[i] = 0; Cx[i] = lo[i]; return false; }
      else if (x[i] > hi[i]) { S[i] = 0; Cx[i] = hi[i]; return false; }
    } else {
      if (Cx[i] == lo[i] && w[i] < 0) { S[i] = 1; return false; }
      else if (Cx[i] == hi[i] && w[i] > 0) { S[i] = 1; return false; }
    }
  }
  for (int i = 0; i < dim; i++) if (x[i] < lo[i] || x[i] > hi[i]) return false;
  for (int i = 0; i < dim; i++) {
    if (x[i] == lo[i] && w[i] < 0) return false;
    if (x[i] == hi[i] && w[i] > 0) return false;
  }
  Vec lhs = matvec(A, x);
  double s = 0;
  for (int i = 0; i < dim; i++) { double d = lhs[i] - (b[i] + w[i]); s += d * d; }
  if (std::sqrt(s) > check_err) return false;
  return true;
}

// lcp.cc:98-104: sum of the non-positive parts of x and w.
inline double solution_goodness(const Vec& x, const Vec& w) {
  double gx = 0, gw = 0;
  for (double v : x) gx += (v > 0) ? 0.0 : v;
  for (double v : w) gw += (v > 0) ? 0.0 : v;
  return gx + gw;
}
// lcp.cc:127-137
inline bool update_previous_best(const Vec& nx, const Vec& nw, Vec& px, Vec& pw) {
  if (nx == px && nw == pw) return false;
  if (solution_goodness(nx, nw) > solution_goodness(px, pw)) { px = nx; pw = nw; }
  return true;
}

// lcp.cc:157-274.  Returns the reference's bool; x,w are the best-goodness iterate.
inline bool murty_principal_pivot(const Mat& A, const Vec& b, Vec& x, Vec& w, const Vec& lo,
                                  const Vec& hi, MurtyStats* st = nullptr) {
  const int dim = (int)b.size();
  const double p2 = std::pow(2.0, dim);
  const int max_iterations = p2 > 1000 ? 1000 : (int)p2;
  int iter = 0;
  Mask S(dim, 1);
  x.assign(dim, 0.0);
  w.resize(dim);
  for (int i = 0; i < dim; i++) w[i] = -b[i];
  Vec Cx(dim);
  for (int i = 0; i < dim; i++) Cx[i] = 1.0 * lo[i];
  Vec bx = x, bw = w;
  while (iter < max_iterations) {
    if (!check_murty_solution(A, b, x, w, S, Cx, lo, hi)) {
      Mat Ass = select_submatrix(A, S, S);
      LDLT f;
      f.compute(Ass);
      Vec xs = f.solve(select_subvector(b, S));
      update_subvector(x, S, xs);
      for (int i = 0; i < dim; i++) {
        if (!S[i] && Cx[i] == lo[i]) x[i] = lo[i];
      }
      for (int i = 0; i < dim; i++) {
        if (!S[i] && Cx[i] == hi[i]) x[i] = hi[i];
      }
      Mask nS(dim);
      for (int i = 0; i < dim; i++) nS[i] = !S[i];
      Mat Ans = select_submatrix(A, nS, S);
      Vec wn = matvec(Ans, select_subvector(x, S));
      Vec bn = select_subvector(b, nS);
      for (size_t i = 0; i < wn.size(); i++) wn[i] -= bn[i];
      update_subvector(w, nS, wn);
      update_subvector(w, S, 0.0);
      update_previous_best(x, w, bx, bw);
    } else {
      break;
    }
    ++iter;
  }
  x = bx;
  w = bw;
  bool ok;
  if (iter >= max_iterations) ok = check_murty_solution(A, b, x, w, S, Cx, lo, hi, kLcpLooserAllowedError);
  else ok = check_murty_solution(A, b, x, w, S, Cx, lo, hi);
  if (st) { st->iterations = iter; st->hit_cap = iter >= max_iterations; st->S = S; }
  return ok;
}

// lcp.cc:276-336.  q1: the per-row bounds passed in are IGNORED by the reference (4-argument
// Murty => [0, inf) on every inequality row); honour_bounds=true is the non-reference variant.
inline bool mixed_constraints_solver(const Mat& A, const Vec& b, const Mask& C, const Vec& x_lo,
                                     const Vec& x_hi, Vec& x, Vec& w, bool honour_bounds = false,
                                     MurtyStats* st = nullptr) {
  const int dim = A.r;
  Mask nC(dim);
  int dim_eq = 0;
  for (int i = 0; i < dim; i++) { nC[i] = !C[i]; dim_eq += C[i] != 0; }
  Mat A_ee = select_submatrix(A, C, C), A_ei = select_submatrix(A, C, nC);
  Mat A_ie = select_submatrix(A, nC, C), A_ii = select_submatrix(A, nC, nC);
  Vec b_e = select_subvector(b, C), b_i = select_subvector(b, nC);
  Mat lhs = A_ii;
  Vec rhs = b_i;
  if (dim_eq > 0 && dim - dim_eq > 0) {
    Mat Aee_inv = lu_inverse(A_ee);
    Mat T = matmul(matmul(A_ie, Aee_inv), A_ei);           // (A_ie * A_ee^-1) * A_ei
    for (size_t k = 0; k < lhs.a.size(); k++) lhs.a[k] -= T.a[k];
    Vec t = matvec(matmul(A_ie, Aee_inv), b_e);
    for (size_t k = 0; k < rhs.size(); k++) rhs[k] -= t[k];
  }
  const int ni = dim - dim_eq;
  Vec lo(ni, 0.0), hi(ni, std::numeric_limits<double>::infinity());
  if (honour_bounds) { lo = select_subvector(x_lo, nC); hi = select_subvector(x_hi, nC); }
  Vec x_i, w_i;
  bool ok = murty_principal_pivot(lhs, rhs, x_i, w_i, lo, hi, st);
  Vec x_e;
  if (dim_eq > 0) {
    Vec r = b_e;
    if (ni > 0) {
      Vec t = matvec(A_ei, x_i);
      for (int k = 0; k < dim_eq; k++) r[k] -= t[k];
    }
    LDLT f;
    f.compute(A_ee);
    x_e = f.solve(r);
  }
  x.assign(dim, std::numeric_limits<double>::infinity());
  update_subvector(x, C, x_e);
  update_subvector(x, nC, x_i);
  w.assign(dim, 0.0);
  update_subvector(w, nC, w_i);
  return ok;
}

}  // namespace orc
