// Dense mixed-LCP path: what the reference actually ships in ComputeVDot.
//
// Replaces (per world, one CTA per world, working set in an L2-resident global scratch):
//   Ensemble::ComputeVDot            ensembles.cc:498-538   A = J M^-1 J^T (+ cfm I), solve, v_dot
//   CheckMatrixCondition             utils.cc:256-287       "cond >= 1e7 => add cfm" decision
//   Lcp::MixedConstraintsSolver      lcp.cc:276-336         Schur complement on the equality rows
//   Lcp::MurtyPrincipalPivot         lcp.cc:157-274         least-index principal pivoting
//   CheckMurtySolution / best-so-far lcp.cc:20-94,98-137
//   Eigen LDLT<>::compute / solve    (lcp.cc:203,317)       restated as in oracle/orc_linalg.h
//   + v' = v + dt M^-1 (f + J^T x) and StepPositions_ODE    ensembles.cc:535,572-591
//
// The pivot sequence and the active set are discrete outputs compared bit-exactly with the oracle,
// so every reduction whose result feeds a comparison is summed in the oracle's order (ascending
// index, one accumulator) and this file is compiled with -fmad=false.
//
// Deviation (documented in DESIGN.md): the reference decides "add cfm" from a JacobiSVD condition
// number (an O(rows^3) SVD per step).  Here the decision uses the pivot range of a diagonally
// pivoted LDL^T of A (max |d| / min |d| >= 1e7, or a non-positive pivot).  The two agree whenever
// the condition number is not within a small factor of 1e7; on the built-in scenes it is either
// < 1e4 or > 1e15.
#include "egg_internal.cuh"

namespace {

constexpr int DT = 128;   // threads per CTA

struct DenseScratch {
  double* A;      // [R][R]
  double* L;      // [I][I]   Schur complement (Murty's A)
  double* F;      // [R][R]   factorisation workspace (>= max(E, I)^2)
  double* X;      // [E][I+1] A_ee^-1 [A_ei | b_e]
  double* Jb;     // [nc][2][18] Jacobian blocks, then [nc][2][18] J M^-1
  double* vec;    // 12 vectors of length R
  int* ivec;      // 4 int vectors of length R
};

// ---- Eigen-style LDLT on a k x k matrix M (row-major, leading dimension ld), lower part used ----
// tr[k] transpositions; tmp[k] scratch.  Mirrors orc::LDLT::compute.
__device__ void ldlt_compute(double* M, int ld, int k, int* tr, double* tmp) {
  const int tid = threadIdx.x;
  __shared__ int s_big;
  for (int kk = 0; kk < k; kk++) {
    if (tid == 0) {
      int big = kk;
      double best = fabs(M[(size_t)kk * ld + kk]);
      for (int i = kk + 1; i < k; i++) {
        double v = fabs(M[(size_t)i * ld + i]);
        if (v > best) { best = v; big = i; }
      }
      tr[kk] = big;
      s_big = big;
    }
    __syncthreads();
    const int big = s_big;
    if (big != kk) {
      for (int j = tid; j < kk; j += DT) { double t = M[(size_t)kk * ld + j]; M[(size_t)kk * ld + j] = M[(size_t)big * ld + j]; M[(size_t)big * ld + j] = t; }
      for (int i = big + 1 + tid; i < k; i += DT) { double t = M[(size_t)i * ld + kk]; M[(size_t)i * ld + kk] = M[(size_t)i * ld + big]; M[(size_t)i * ld + big] = t; }
      for (int i = kk + 1 + tid; i < big; i += DT) { double t = M[(size_t)i * ld + kk]; M[(size_t)i * ld + kk] = M[(size_t)big * ld + i]; M[(size_t)big * ld + i] = t; }
      if (tid == 0) { double t = M[(size_t)kk * ld + kk]; M[(size_t)kk * ld + kk] = M[(size_t)big * ld + big]; M[(size_t)big * ld + big] = t; }
      __syncthreads();
    }
    if (kk > 0) {
      for (int j = tid; j < kk; j += DT) tmp[j] = M[(size_t)j * ld + j] * M[(size_t)kk * ld + j];
      __syncthreads();
      if (tid == 0) {
        double s = 0;
        for (int j = 0; j < kk; j++) s += M[(size_t)kk * ld + j] * tmp[j];
        M[(size_t)kk * ld + kk] -= s;
      }
      for (int i = kk + 1 + tid; i < k; i += DT) {
        double t = 0;
        for (int j = 0; j < kk; j++) t += M[(size_t)i * ld + j] * tmp[j];
        M[(size_t)i * ld + kk] -= t;
      }
      __syncthreads();
    }
    const double akk = M[(size_t)kk * ld + kk];
    const bool valid = fabs(akk) > 0;
    if (kk == 0 && !valid) {
      for (int j = tid; j < k; j += DT) tr[j] = j;
      __syncthreads();
      return;
    }
    if (valid)
      for (int i = kk + 1 + tid; i < k; i += DT) M[(size_t)i * ld + kk] /= akk;
    __syncthreads();
  }
}

// x <- solve(M factor, x) in place (x length k).  Mirrors orc::LDLT::solve.
__device__ void ldlt_solve(const double* M, int ld, int k, const int* tr, double* x) {
  const int tid = threadIdx.x;
  if (tid == 0)
    for (int i = 0; i < k; i++) if (tr[i] != i) { double t = x[i]; x[i] = x[tr[i]]; x[tr[i]] = t; }
  __syncthreads();
  for (int i = 0; i < k; i++) {            // L^-1, column-oriented: same per-entry operation order
    const double xi = x[i];
    for (int r = i + 1 + tid; r < k; r += DT) x[r] -= M[(size_t)r * ld + i] * xi;
    __syncthreads();
  }
  const double tol = 1.0 / 1.7976931348623157e308;
  for (int i = tid; i < k; i += DT) {
    const double dd = M[(size_t)i * ld + i];
    x[i] = (fabs(dd) > tol) ? x[i] / dd : 0.0;
  }
  __syncthreads();
  for (int i = k - 1; i >= 0; i--) {       // L^-T
    const double xi = x[i];
    for (int r = tid; r < i; r += DT) x[r] -= M[(size_t)i * ld + r] * xi;
    __syncthreads();
  }
  if (tid == 0)
    for (int i = k - 1; i >= 0; i--) if (tr[i] != i) { double t = x[i]; x[i] = x[tr[i]]; x[tr[i]] = t; }
  __syncthreads();
}

// X[:, j] <- solve(M factor, X[:, j]) for ncols right-hand sides stored as columns of X (row
// stride ldx).  One thread per column runs the whole substitution for it: the per-entry operation
// order is that of ldlt_solve (so the results are the same bits), the threads read M as a
// broadcast and X coalesced, and there is no barrier inside (the one-RHS-at-a-time form cost
// 2k barriers per right-hand side: 37 000 for the Schur complement of a 32-link chain).
__device__ void ldlt_solve_columns(const double* M, int ld, int k, const int* tr, double* X, int ldx, int ncols) {
  const double tol = 1.0 / 1.7976931348623157e308;
  for (int j = threadIdx.x; j < ncols; j += DT) {
    double* x = X + j;
    for (int i = 0; i < k; i++) if (tr[i] != i) { double t = x[(size_t)i * ldx]; x[(size_t)i * ldx] = x[(size_t)tr[i] * ldx]; x[(size_t)tr[i] * ldx] = t; }
    for (int i = 0; i < k; i++) {            // L^-1
      const double xi = x[(size_t)i * ldx];
      for (int r = i + 1; r < k; r++) x[(size_t)r * ldx] -= M[(size_t)r * ld + i] * xi;
    }
    for (int i = 0; i < k; i++) {
      const double dd = M[(size_t)i * ld + i];
      x[(size_t)i * ldx] = (fabs(dd) > tol) ? x[(size_t)i * ldx] / dd : 0.0;
    }
    for (int i = k - 1; i >= 0; i--) {       // L^-T
      const double xi = x[(size_t)i * ldx];
      for (int r = 0; r < i; r++) x[(size_t)r * ldx] -= M[(size_t)i * ld + r] * xi;
    }
    for (int i = k - 1; i >= 0; i--) if (tr[i] != i) { double t = x[(size_t)i * ldx]; x[(size_t)i * ldx] = x[(size_t)tr[i] * ldx]; x[(size_t)tr[i] * ldx] = t; }
  }
  __syncthreads();
}

// Pivot range of a complete-diagonal-pivoted LDL^T (right-looking, updated diagonal): the
// "is A ill conditioned" proxy.  Destroys M.  Returns true if well conditioned (ratio < 1e7).
__device__ bool well_conditioned(double* M, int ld, int k) {
  const int tid = threadIdx.x;
  __shared__ int s_piv;
  __shared__ double s_dmax, s_dmin;
  __shared__ int s_bad;
  if (tid == 0) { s_dmax = 0; s_dmin = 1.7976931348623157e308; s_bad = 0; }
  __syncthreads();
  for (int kk = 0; kk < k; kk++) {
    if (tid == 0) {
      int big = kk;
      double best = M[(size_t)kk * ld + kk];
      for (int i = kk + 1; i < k; i++) { double v = M[(size_t)i * ld + i]; if (v > best) { best = v; big = i; } }
      s_piv = big;
      if (!(best > 0)) s_bad = 1;
      else { if (best > s_dmax) s_dmax = best; if (best < s_dmin) s_dmin = best; }
    }
    __syncthreads();
    if (s_bad) return false;
    if (s_dmax >= 1e7 * s_dmin) return false;
    const int big = s_piv;
    if (big != kk) {   // full symmetric swap (both triangles kept)
      for (int j = tid; j < k; j += DT) { double t = M[(size_t)kk * ld + j]; M[(size_t)kk * ld + j] = M[(size_t)big * ld + j]; M[(size_t)big * ld + j] = t; }
      __syncthreads();
      for (int i = tid; i < k; i += DT) { double t = M[(size_t)i * ld + kk]; M[(size_t)i * ld + kk] = M[(size_t)i * ld + big]; M[(size_t)i * ld + big] = t; }
      __syncthreads();
    }
    // trailing update, one column j per thread (coalesced across the threads, the multiplier
    // column M[i][kk] is a broadcast): no integer or floating-point division in the k^3/3 loop.
    // Only the pivot RANGE of this factorisation is used (a yes/no decision), so the rounding of
    // a * b / d versus a * (b / d) is immaterial.
    const double invd = 1.0 / M[(size_t)kk * ld + kk];
    for (int j = kk + 1 + tid; j < k; j += DT) {
      const double ukj = M[(size_t)kk * ld + j] * invd;
      for (int i = kk + 1; i < k; i++) M[(size_t)i * ld + j] -= M[(size_t)i * ld + kk] * ukj;
    }
    __syncthreads();
  }
  return true;
}

// lcp.cc:20-94 with x_lo / x_hi per row; Cx[i] = bound at which a non-basic x(i) sits.
// Returns 1 = solution, 0 = not a solution (S possibly flipped at the least offending index).
__device__ int check_murty(const double* A, int ld, int dim, const double* b, const double* x, const double* w, unsigned char* S,
                           double* Cx, const double* lo, const double* hi, double err, double* tmp) {
  const int tid = threadIdx.x;
  __shared__ int s_first, s_result;
  if (tid == 0) { s_first = dim; s_result = -1; }
  __syncthreads();
  int mine = dim;
  for (int i = tid; i < dim; i += DT) {
    bool viol;
    if (S[i]) viol = (x[i] < lo[i]) || (x[i] > hi[i]);
    else viol = (Cx[i] == lo[i] && w[i] < 0) || (Cx[i] == hi[i] && w[i] > 0);
    if (viol) { mine = i; break; }
  }
  if (mine < dim) atomicMin(&s_first, mine);
  __syncthreads();
  if (s_first < dim) {
    if (tid == 0) {
      const int i = s_first;
      if (S[i]) { S[i] = 0; Cx[i] = (x[i] < lo[i]) ? lo[i] : hi[i]; }
      else S[i] = 1;
    }
    __syncthreads();
    return 0;
  }
  // goodness checks
  int bad = 0;
  for (int i = tid; i < dim; i += DT) {
    if (x[i] < lo[i] || x[i] > hi[i]) bad = 1;
    if (x[i] == lo[i] && w[i] < 0) bad = 1;
    if (x[i] == hi[i] && w[i] > 0) bad = 1;
  }
  if (__syncthreads_or(bad)) return 0;
  for (int i = tid; i < dim; i += DT) {
    double s = 0;
    for (int j = 0; j < dim; j++) s += A[(size_t)i * ld + j] * x[j];
    tmp[i] = s - (b[i] + w[i]);
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int i = 0; i < dim; i++) s += tmp[i] * tmp[i];
    const double chk = fabs(err) > 1e-9 ? fabs(err) : 1e-9;
    s_result = (sqrt(s) > chk) ? 0 : 1;
  }
  __syncthreads();
  return s_result;
}

__global__ void __launch_bounds__(DT) egg_dense_kernel(EggDev d, double dt, double* scratch, size_t per_world, int Rcap) {
  const int w = blockIdx.x, tid = threadIdx.x, n = d.n, nj = d.nj;
  const int nc = nj + d.c_count[w];
  int R = 3 * nc;
  const int E = 3 * nj;
  __shared__ int s_flag;
  extern __shared__ double sa[];   // [6][n] accumulator a = M^-1 J^T lambda
  for (int i = tid; i < 6 * n; i += DT) sa[i] = 0.0;
  const double* st = d.stat + (size_t)w * EGG_STAT * n;
  double* lam_out = d.lam_out + (size_t)w * 3 * d.nrec;
  int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
  int* stt = d.stats + (size_t)w * 8;
  bool overflow = false;
  if (R > Rcap) { overflow = true; R = 0; }
  const int I = R - (R ? E : 0);
  int pivots = 0, cfm_applied = 0, lcp_failed = 0;

  if (R > 0) {
    double* base = scratch + (size_t)w * per_world;
    double* A = base;
    double* L = A + (size_t)Rcap * Rcap;
    double* F = L + (size_t)Rcap * Rcap;
    double* X = F + (size_t)Rcap * Rcap;
    double* Jb = X + (size_t)Rcap * (Rcap + 1);
    double* vec = Jb + (size_t)(Rcap / 3 + 1) * 72;
    double* b = vec;                 // rhs [R]
    double* x = vec + Rcap;          // Murty x [I]
    double* wv = vec + 2 * Rcap;     // Murty w [I]
    double* bx = vec + 3 * Rcap;
    double* bw = vec + 4 * Rcap;
    double* lo = vec + 5 * Rcap;
    double* hi = vec + 6 * Rcap;
    double* Cx = vec + 7 * Rcap;
    double* tmp = vec + 8 * Rcap;
    double* xs = vec + 9 * Rcap;     // sub-system rhs / solution
    double* rhsL = vec + 10 * Rcap;  // Schur rhs [I]
    double* lamv = vec + 11 * Rcap;  // final lambda [R], reference row order
    int* ivec = reinterpret_cast<int*>(base + per_world) - 4 * Rcap;
    int* tr = ivec;
    int* sidx = ivec + Rcap;
    int* ci0 = ivec + 2 * Rcap;      // body indices per constraint (reference order)
    int* ci1 = ivec + 3 * Rcap;
    unsigned char* S = reinterpret_cast<unsigned char*>(ivec) - Rcap;   // basic-set mask, just below the int vectors in the tail

    // ---- 1. Jacobian blocks per constraint in REFERENCE order (records are in level order) ----
    const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
    for (int s = tid; s < nc; s += DT) {
      const double* r = recs + (size_t)s * EGG_REC;
      const int i0 = __double2loint(r[REC_IDX]), i1 = __double2hiint(r[REC_IDX]);
      const int c = __double2loint(r[REC_META]);
      ci0[c] = i0; ci1[c] = i1;
      d3 r0 = mk3(r[REC_R0], r[REC_R0 + 1], r[REC_R0 + 2]), r1 = mk3(r[REC_R1], r[REC_R1 + 1], r[REC_R1 + 2]);
      double* J0 = Jb + (size_t)c * 72;
      double* J1 = J0 + 18;
      double* B0 = J0 + 36;
      double* B1 = J0 + 54;
      for (int k = 0; k < 3; k++) {
        d3 rc = mk3(r[REC_RC + 3 * k], r[REC_RC + 3 * k + 1], r[REC_RC + 3 * k + 2]);
        d3 a0 = cross3(rc, r0), a1 = cross3(r1, rc);
        J0[6 * k] = -rc.x; J0[6 * k + 1] = -rc.y; J0[6 * k + 2] = -rc.z; J0[6 * k + 3] = a0.x; J0[6 * k + 4] = a0.y; J0[6 * k + 5] = a0.z;
        J1[6 * k] = rc.x; J1[6 * k + 1] = rc.y; J1[6 * k + 2] = rc.z; J1[6 * k + 3] = a1.x; J1[6 * k + 4] = a1.y; J1[6 * k + 5] = a1.z;
        b[3 * c + k] = r[REC_RHS + k];
      }
      for (int side = 0; side < 2; side++) {
        const int bd = side ? i1 : i0;
        const double* Jx = side ? J1 : J0;
        double* Bx = side ? B1 : B0;
        if (bd < 0) { for (int q = 0; q < 18; q++) Bx[q] = 0.0; continue; }
        const double mi = st[bd];
        double Ii[9];
        for (int q = 0; q < 9; q++) Ii[q] = st[(1 + q) * n + bd];
        for (int k = 0; k < 3; k++) {       // (J * M^-1) row k
          for (int q = 0; q < 3; q++) Bx[6 * k + q] = Jx[6 * k + q] * mi;
          for (int q = 0; q < 3; q++) {
            double s2 = 0;
            for (int t = 0; t < 3; t++) s2 += Jx[6 * k + 3 + t] * Ii[3 * t + q];
            Bx[6 * k + 3 + q] = s2;
          }
        }
      }
      // bounds of the rows of c: joints are equalities; contacts [0,inf) under q1, BOX otherwise
      const bool q1 = (d.prm.quirks & 2) != 0;
      for (int k = 0; k < 3; k++) {
        const bool contact = c >= nj;
        lo[3 * c + k] = (contact && !q1 && k < 2) ? -1.0 : 0.0;
        hi[3 * c + k] = (contact && !q1 && k < 2) ? 1.0 : __longlong_as_double(0x7ff0000000000000LL);   // +inf
      }
    }
    __syncthreads();

    // ---- 2. A = J M^-1 J^T (ensembles.cc:510): block (c, c') = sum over shared bodies ----
    for (int e = tid; e < nc * nc; e += DT) {
      const int c = e / nc, c2 = e % nc;
      const int a0 = ci0[c], a1 = ci1[c], e0 = ci0[c2], e1 = ci1[c2];
      double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      // visit the bodies of c in ascending column order, as the dense product does
      for (int pass = 0; pass < 2; pass++) {
        const int side = ((a0 <= a1) == (pass == 0)) ? 0 : 1;
        const int bd = side ? a1 : a0;
        if (bd < 0) continue;
        int side2 = -1;
        if (bd == e0) side2 = 0; else if (bd == e1) side2 = 1;
        if (side2 < 0) continue;
        const double* Bx = Jb + (size_t)c * 72 + 36 + 18 * side;
        const double* Jy = Jb + (size_t)c2 * 72 + 18 * side2;
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) {
            double s2 = blk[3 * k + l];
            for (int q = 0; q < 6; q++) s2 += Bx[6 * k + q] * Jy[6 * l + q];
            blk[3 * k + l] = s2;
          }
      }
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) A[(size_t)(3 * c + k) * R + 3 * c2 + l] = blk[3 * k + l];
    }
    __syncthreads();

    // ---- 3. cfm decision (ensembles.cc:513-521) ----
    bool good;
    if (d.prm.cfm_mode == 1) good = false;
    else if (d.prm.cfm_mode == 2) good = true;
    else {
      for (int e = tid; e < R * R; e += DT) F[e] = A[e];
      __syncthreads();
      good = well_conditioned(F, R, R);
      __syncthreads();
    }
    if (!good) {
      for (int i = tid; i < R; i += DT) A[(size_t)i * R + i] += d.prm.cfm;
      cfm_applied = 1;
      __syncthreads();
    }

    // ---- 4. Schur complement on the equality rows (lcp.cc:286-294); rows 0..E-1 are the joints ----
    if (I > 0) {
      if (E > 0) {
        for (int e = tid; e < E * E; e += DT) F[(size_t)(e / E) * E + e % E] = A[(size_t)(e / E) * R + e % E];
        __syncthreads();
        ldlt_compute(F, E, E, tr, tmp);
        // X[:, j] = A_ee^-1 A_ei[:, j] (j < I), X[:, I] = A_ee^-1 b_e ; all right-hand sides at once
        for (int e = tid; e < E * (I + 1); e += DT) {
          const int i = e / (I + 1), j = e % (I + 1);
          X[e] = (j < I) ? A[(size_t)i * R + E + j] : b[i];
        }
        __syncthreads();
        ldlt_solve_columns(F, E, E, tr, X, I + 1, I + 1);
      }
      for (int e = tid; e < I * (I + 1); e += DT) {
        const int i = e / (I + 1), j = e % (I + 1);
        double s2 = 0;
        for (int k = 0; k < E; k++) s2 += A[(size_t)(E + i) * R + k] * X[(size_t)k * (I + 1) + j];
        if (j < I) L[(size_t)i * I + j] = A[(size_t)(E + i) * R + E + j] - s2;
        else rhsL[i] = b[E + i] - s2;
      }
      __syncthreads();

      // ---- 5. Murty principal pivoting on (L, rhsL) with bounds lo/hi of the inequality rows ----
      const double* loI = lo + E;
      const double* hiI = hi + E;
      for (int i = tid; i < I; i += DT) { S[i] = 1; x[i] = 0.0; wv[i] = -rhsL[i]; Cx[i] = loI[i]; bx[i] = 0.0; bw[i] = -rhsL[i]; }
      __syncthreads();
      const int max_it = (I >= 10) ? 1000 : (1 << I);
      int iter = 0;
      while (iter < max_it) {
        if (check_murty(L, I, I, rhsL, x, wv, S, Cx, loI, hiI, 0.0, tmp)) break;
        // gather the basic set
        if (tid == 0) {
          int ks = 0;
          for (int i = 0; i < I; i++) if (S[i]) sidx[ks++] = i;
          s_flag = ks;
        }
        __syncthreads();
        const int ks = s_flag;
        for (int e = tid; e < ks * ks; e += DT) F[e] = L[(size_t)sidx[e / ks] * I + sidx[e % ks]];
        for (int i = tid; i < ks; i += DT) xs[i] = rhsL[sidx[i]];
        __syncthreads();
        ldlt_compute(F, ks, ks, tr, tmp);
        ldlt_solve(F, ks, ks, tr, xs);
        for (int i = tid; i < ks; i += DT) x[sidx[i]] = xs[i];
        for (int i = tid; i < I; i += DT)
          if (!S[i]) x[i] = (Cx[i] == loI[i]) ? loI[i] : hiI[i];
        __syncthreads();
        for (int i = tid; i < I; i += DT) {
          if (S[i]) { wv[i] = 0.0; continue; }
          double s2 = 0;
          for (int k = 0; k < ks; k++) s2 += L[(size_t)i * I + sidx[k]] * xs[k];
          wv[i] = s2 - rhsL[i];
        }
        __syncthreads();
        // UpdatePreviousBestSolution (lcp.cc:127-137)
        if (tid == 0) {
          bool same = true;
          double gn = 0, gp = 0, gnw = 0, gpw = 0;
          for (int i = 0; i < I; i++) {
            if (x[i] != bx[i] || wv[i] != bw[i]) same = false;
            gn += (x[i] > 0) ? 0.0 : x[i];
            gp += (bx[i] > 0) ? 0.0 : bx[i];
          }
          for (int i = 0; i < I; i++) { gnw += (wv[i] > 0) ? 0.0 : wv[i]; gpw += (bw[i] > 0) ? 0.0 : bw[i]; }
          s_flag = (!same && (gn + gnw) > (gp + gpw)) ? 1 : 0;
        }
        __syncthreads();
        if (s_flag)
          for (int i = tid; i < I; i += DT) { bx[i] = x[i]; bw[i] = wv[i]; }
        __syncthreads();
        ++iter;
      }
      pivots = iter;
      for (int i = tid; i < I; i += DT) { x[i] = bx[i]; wv[i] = bw[i]; }
      __syncthreads();
      const int ok = check_murty(L, I, I, rhsL, x, wv, S, Cx, loI, hiI, (iter >= max_it) ? 1e-8 : 0.0, tmp);
      if (!ok) lcp_failed = 1;
    }

    // ---- 6. x_e = A_ee.ldlt().solve(b_e - A_ei x_i)  (lcp.cc:317) ----
    if (E > 0) {
      for (int i = tid; i < E; i += DT) {
        double s2 = 0;
        for (int j = 0; j < I; j++) s2 += A[(size_t)i * R + E + j] * x[j];
        xs[i] = b[i] - s2;
      }
      __syncthreads();
      // (F was reused by Murty: refactor A_ee)
      for (int e = tid; e < E * E; e += DT) F[(size_t)(e / E) * E + e % E] = A[(size_t)(e / E) * R + e % E];
      __syncthreads();
      ldlt_compute(F, E, E, tr, tmp);
      ldlt_solve(F, E, E, tr, xs);
      for (int i = tid; i < E; i += DT) lamv[i] = xs[i];
    }
    for (int i = tid; i < I; i += DT) lamv[E + i] = x[i];
    __syncthreads();

    // ---- 7. outputs + a = M^-1 J^T lambda (per body, constraints in reference order) ----
    for (int i = tid; i < R; i += DT) {
      lam_out[i] = lamv[i];
      rs_out[i] = (i < E) ? 3 : (S[i - E] ? 0 : 1);
    }
    for (int bd = tid; bd < n; bd += DT) {
      double g[6] = {0, 0, 0, 0, 0, 0};
      for (int c = 0; c < nc; c++) {
        for (int side = 0; side < 2; side++) {
          if ((side ? ci1[c] : ci0[c]) != bd) continue;
          const double* Jx = Jb + (size_t)c * 72 + 18 * side;
          for (int k = 0; k < 3; k++) {
            const double l = lamv[3 * c + k];
            for (int q = 0; q < 6; q++) g[q] += Jx[6 * k + q] * l;
          }
        }
      }
      const double mi = st[bd];
      double Ii[9];
      for (int q = 0; q < 9; q++) Ii[q] = st[(1 + q) * n + bd];
      d3 al = mk3(g[0], g[1], g[2]) * mi;
      d3 aa = mmulv(Ii, mk3(g[3], g[4], g[5]));
      sa[bd] = al.x; sa[n + bd] = al.y; sa[2 * n + bd] = al.z;
      sa[3 * n + bd] = aa.x; sa[4 * n + bd] = aa.y; sa[5 * n + bd] = aa.z;
    }
  }
  __syncthreads();

  if (tid == 0) {
    stt[4] = 0;
    stt[5] = pivots;
    stt[6] = cfm_applied;
    stt[7] = 0;
    d.resid[w] = 0.0;
    int flags = d.status[w] & ~1;
    if (lcp_failed) flags |= 1;
    if (overflow) flags |= 32;
    d.status[w] = flags;
  }

  // ---- 8. integrate (same arithmetic as the PGS path) ----
  double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  for (int bd = tid; bd < n; bd += DT) {
    const double mi = st[bd];
    double Ii[9];
    for (int k = 0; k < 9; k++) Ii[k] = st[(1 + k) * n + bd];
    d3 fl = mk3(st[10 * n + bd], st[11 * n + bd], st[12 * n + bd]);
    d3 ft = mk3(st[13 * n + bd], st[14 * n + bd], st[15 * n + bd]);
    d3 v = mk3(dyn[12 * n + bd], dyn[13 * n + bd], dyn[14 * n + bd]);
    d3 wv3 = mk3(dyn[15 * n + bd], dyn[16 * n + bd], dyn[17 * n + bd]);
    d3 al = mk3(sa[bd], sa[n + bd], sa[2 * n + bd]);
    d3 aa = mk3(sa[3 * n + bd], sa[4 * n + bd], sa[5 * n + bd]);
    d3 vn = v + dt * (fl * mi + al);
    d3 wn = wv3 + dt * (mmulv(Ii, ft) + aa);
    d3 vmid = (v + vn) / 2.0, wmid = (wv3 + wn) / 2.0;
    d3 p = mk3(dyn[bd], dyn[n + bd], dyn[2 * n + bd]) + dt * vmid;
    double wnorm = norm3(wmid);
    double z2 = dot3(wmid, wmid);
    d3 axis = (z2 > 0) ? wmid / sqrt(z2) : wmid;
    double ha = 0.5 * (wnorm * dt);
    double qw = cos(ha), sn = sin(ha);
    double qx = sn * axis.x, qy = sn * axis.y, qz = sn * axis.z;
    double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
    double twx = tx * qw, twy = ty * qw, twz = tz * qw;
    double txx = tx * qx, txy = ty * qx, txz = tz * qx;
    double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    double Q[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx,
                   txz - twy, tyz + twx, 1 - (txx + tyy)};
    double Rm[9], Rn[9];
    for (int k = 0; k < 9; k++) Rm[k] = dyn[(3 + k) * n + bd];
    mmulm(Q, Rm, Rn);
    dyn[bd] = p.x; dyn[n + bd] = p.y; dyn[2 * n + bd] = p.z;
    for (int k = 0; k < 9; k++) dyn[(3 + k) * n + bd] = Rn[k];
    dyn[12 * n + bd] = vn.x; dyn[13 * n + bd] = vn.y; dyn[14 * n + bd] = vn.z;
    dyn[15 * n + bd] = wn.x; dyn[16 * n + bd] = wn.y; dyn[17 * n + bd] = wn.z;
    double chk = p.x + p.y + p.z + vn.x + vn.y + vn.z + wn.x + wn.y + wn.z;
    if (!(fabs(chk) < 1e300)) atomicOr(&d.status[w], 16);
  }
}

// ---------------------------------------------------------------------------------------------
// Position relaxation: one iteration of Ensemble::InitStabilize's loop body
// (/root/reference/eggshell/ensembles.cc:602-622): StepPositionRelaxation(dt = 0.5) =
// StepPositions_ExplicitEuler(dt, -0.2 J^T (J J^T)^-1 err)  (:647-650, 659-666, 553-561) for every
// world whose squared position error still exceeds 1e-9 and whose step counter is below
// max_steps.  Contacts must have been refreshed WITHOUT de-duplication (the reference only runs
// CheckAndCorrectEnsembleState after the loop) and the records assembled for the current state.
// prog[w] = {steps taken, active flag}; err2[w] = squared error seen by this call.
__global__ void __launch_bounds__(DT) egg_relax_kernel(EggDev d, double dt, double step_scale, int max_steps, double* scratch,
                                                        size_t per_world, int Rcap, int* prog, double* err2, int* any_active) {
  const int w = blockIdx.x, tid = threadIdx.x, n = d.n, nj = d.nj;
  const int nc = nj + d.c_count[w];
  const int R = 3 * nc;
  __shared__ double s_e2;
  __shared__ int s_go;
  double* base = scratch + (size_t)w * per_world;
  double* A = base;
  double* F = A + (size_t)Rcap * Rcap * 2;
  double* Jb = F + (size_t)Rcap * Rcap + (size_t)Rcap * (Rcap + 1);
  double* vec = Jb + (size_t)(Rcap / 3 + 1) * 72;
  double* e = vec;                 // err [R], reference row order
  double* y = vec + Rcap;
  double* tmp = vec + 2 * Rcap;
  int* ivec = reinterpret_cast<int*>(base + per_world) - 4 * Rcap;
  int* tr = ivec;
  int* ci0 = ivec + 2 * Rcap;
  int* ci1 = ivec + 3 * Rcap;
  const double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
  const double* geom = d.c_geom + (size_t)w * 7 * d.maxc;
  const bool fits = R <= Rcap;

  // J blocks + position error per constraint, reference order
  for (int s = tid; fits && s < nc; s += DT) {
    const double* r = recs + (size_t)s * EGG_REC;
    const int i0 = __double2loint(r[REC_IDX]), i1 = __double2hiint(r[REC_IDX]);
    const int c = __double2loint(r[REC_META]);
    ci0[c] = i0; ci1[c] = i1;
    d3 r0 = mk3(r[REC_R0], r[REC_R0 + 1], r[REC_R0 + 2]), r1 = mk3(r[REC_R1], r[REC_R1 + 1], r[REC_R1 + 2]);
    double* J0 = Jb + (size_t)c * 72;
    double* J1 = J0 + 18;
    for (int k = 0; k < 3; k++) {
      d3 rc = mk3(r[REC_RC + 3 * k], r[REC_RC + 3 * k + 1], r[REC_RC + 3 * k + 2]);
      d3 a0 = cross3(rc, r0), a1 = cross3(r1, rc);
      const double z0 = (i0 >= 0) ? 1.0 : 0.0, z1 = (i1 >= 0) ? 1.0 : 0.0;
      J0[6 * k] = -rc.x * z0; J0[6 * k + 1] = -rc.y * z0; J0[6 * k + 2] = -rc.z * z0; J0[6 * k + 3] = a0.x * z0; J0[6 * k + 4] = a0.y * z0; J0[6 * k + 5] = a0.z * z0;
      J1[6 * k] = rc.x * z1; J1[6 * k + 1] = rc.y * z1; J1[6 * k + 2] = rc.z * z1; J1[6 * k + 3] = a1.x * z1; J1[6 * k + 4] = a1.y * z1; J1[6 * k + 5] = a1.z * z1;
    }
    if (c < nj) {                                            // joints.cc:3-11
      d3 p0 = mk3(dyn[i0], dyn[n + i0], dyn[2 * n + i0]);
      d3 er;
      if (i1 < 0) {
        const double* jc = d.jc + (size_t)w * 6 * nj;
        er = p0 + r0 - mk3(jc[3 * nj + c], jc[4 * nj + c], jc[5 * nj + c]);
      } else {
        er = p0 + r0 - mk3(dyn[i1], dyn[n + i1], dyn[2 * n + i1]) - r1;
      }
      e[3 * c] = er.x; e[3 * c + 1] = er.y; e[3 * c + 2] = er.z;
    } else {                                                 // contact.cc:14-22
      e[3 * c] = 0.0; e[3 * c + 1] = 0.0; e[3 * c + 2] = -geom[6 * d.maxc + (c - nj)];
    }
  }
  __syncthreads();
  if (tid == 0) {
    double s2 = 0;
    for (int i = 0; fits && i < R; i++) s2 += e[i] * e[i];
    s_e2 = s2;
    const int steps = prog[2 * w];
    s_go = (fits && s2 > 1e-9 && steps < max_steps) ? 1 : 0;
    err2[w] = s2;
    prog[2 * w + 1] = s_go;
    if (s_go) { prog[2 * w] = steps + 1; atomicOr(any_active, 1); }
    if (!fits) atomicOr(&d.status[w], 32 /*EGG_ST_DENSE_OVERFLOW*/);
  }
  __syncthreads();
  if (!s_go) return;

  // A = J J^T (ensembles.cc:664), blocks over shared bodies in ascending body order
  for (int q = tid; q < nc * nc; q += DT) {
    const int c = q / nc, c2 = q % nc;
    const int a0 = ci0[c], a1 = ci1[c], e0 = ci0[c2], e1 = ci1[c2];
    double blk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int pass = 0; pass < 2; pass++) {
      const int side = ((a0 <= a1) == (pass == 0)) ? 0 : 1;
      const int bd = side ? a1 : a0;
      if (bd < 0) continue;
      int side2 = -1;
      if (bd == e0) side2 = 0; else if (bd == e1) side2 = 1;
      if (side2 < 0) continue;
      const double* Jx = Jb + (size_t)c * 72 + 18 * side;
      const double* Jy = Jb + (size_t)c2 * 72 + 18 * side2;
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          double s2 = blk[3 * k + l];
          for (int t = 0; t < 6; t++) s2 += Jx[6 * k + t] * Jy[6 * l + t];
          blk[3 * k + l] = s2;
        }
    }
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) A[(size_t)(3 * c + k) * R + 3 * c2 + l] = blk[3 * k + l];
  }
  for (int i = tid; i < R; i += DT) y[i] = e[i];
  __syncthreads();
  ldlt_compute(A, R, R, tr, tmp);
  ldlt_solve(A, R, R, tr, y);

  // velocity_correction = -step_scale J^T y ; StepPositions_ExplicitEuler(dt, .)
  double* dynw = d.dyn + (size_t)w * EGG_DYN * n;
  for (int bd = tid; bd < n; bd += DT) {
    double g[6] = {0, 0, 0, 0, 0, 0};
    for (int c = 0; c < nc; c++)
      for (int side = 0; side < 2; side++) {
        if ((side ? ci1[c] : ci0[c]) != bd) continue;
        const double* Jx = Jb + (size_t)c * 72 + 18 * side;
        for (int k = 0; k < 3; k++)
          for (int t = 0; t < 6; t++) g[t] += (-1.0 * step_scale * Jx[6 * k + t]) * y[3 * c + k];
      }
    d3 p = mk3(dynw[bd], dynw[n + bd], dynw[2 * n + bd]) + dt * mk3(g[0], g[1], g[2]);
    d3 wv = mk3(g[3], g[4], g[5]);
    double z2 = dot3(wv, wv);
    d3 axis = (z2 > 0) ? wv / sqrt(z2) : wv;
    double ha = 0.5 * (norm3(wv) * dt);
    double qw = cos(ha), sn = sin(ha);
    double qx = sn * axis.x, qy = sn * axis.y, qz = sn * axis.z;
    double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw;
    double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
    double Q[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    double Rm[9], Rn[9];
    for (int k = 0; k < 9; k++) Rm[k] = dynw[(3 + k) * n + bd];
    mmulm(Q, Rm, Rn);
    dynw[bd] = p.x; dynw[n + bd] = p.y; dynw[2 * n + bd] = p.z;
    for (int k = 0; k < 9; k++) dynw[(3 + k) * n + bd] = Rn[k];
  }
}

}  // namespace

// Rows the dense path is provisioned for: EGG_DENSE_ROWS or min(3 nrec, 336), a multiple of 3.
int egg_dense_row_cap(const EggDev& d) {
  const char* e = getenv("EGG_DENSE_ROWS");
  int cap = e ? atoi(e) : 336;
  if (cap > 3 * d.nrec) cap = 3 * d.nrec;
  if (cap < 3) cap = 3;
  return (cap / 3) * 3;
}

static size_t per_world_doubles(int Rcap) {
  // A, L, F [R^2 each], X [R (R+1)], Jb [(R/3+1) 72], 12 vectors + (S bytes + 4 int vectors) rounded up
  size_t dbl = (size_t)Rcap * Rcap * 3 + (size_t)Rcap * (Rcap + 1) + (size_t)(Rcap / 3 + 1) * 72 + 12 * (size_t)Rcap;
  size_t tail_bytes = (size_t)Rcap * (4 * sizeof(int) + 1) + 16;
  return dbl + (tail_bytes + 7) / 8;
}

size_t egg_dense_scratch_bytes(const EggDev& d) { return per_world_doubles(egg_dense_row_cap(d)) * sizeof(double) * (size_t)d.W; }

void egg_launch_solve_dense(const EggDev& d, double dt, cudaStream_t s, void* scratch, size_t scratch_bytes) {
  const int Rcap = egg_dense_row_cap(d);
  const size_t pw = per_world_doubles(Rcap);
  size_t smem = (size_t)6 * d.n * sizeof(double);
  egg_dense_kernel<<<d.W, DT, smem, s>>>(d, dt, reinterpret_cast<double*>(scratch), pw, Rcap);
}

void egg_launch_relax(const EggDev& d, double dt, double step_scale, int max_steps, void* scratch, int* prog, double* err2, int* any_active,
                      cudaStream_t s) {
  const int Rcap = egg_dense_row_cap(d);
  const size_t pw = per_world_doubles(Rcap);
  egg_relax_kernel<<<d.W, DT, 0, s>>>(d, dt, step_scale, max_steps, reinterpret_cast<double*>(scratch), pw, Rcap, prog, err2, any_active);
}
