// Dense mixed-LCP path: what the reference actually ships in ComputeVDot.
//
// Replaces (per world):
//   Ensemble::ComputeVDot            ensembles.cc:498-538   A = J M^-1 J^T (+ cfm I), solve, v_dot
//   CheckMatrixCondition             utils.cc:256-287       "cond >= 1e7 => add cfm" decision
//   Lcp::MixedConstraintsSolver      lcp.cc:276-336         Schur complement on the equality rows
//   Lcp::MurtyPrincipalPivot         lcp.cc:157-274         least-index principal pivoting
//   CheckMurtySolution / best-so-far lcp.cc:20-94,98-137
//   Eigen LDLT<>::compute / solve    (lcp.cc:203,317)       restated as in oracle/orc_linalg.h
//   dynamic MatrixXd::inverse()      (lcp.cc:293-294)       partial-pivot LU, as orc_linalg.h
//   + v' = v + dt M^-1 (f + J^T x) and StepPositions_ODE    ensembles.cc:535,572-591
//   Ensemble::CalculateVelocityRelaxation / StepPositionRelaxation / StepPostStabilization
//                                    ensembles.cc:647-666   (egg_relax_kernel)
//
// Execution model.  A persistent grid of 256-thread CTAs (two per SM) pulls worlds from an atomic
// queue: the Murty loop takes 1 ... 1000 pivots depending on the world, so a static world -> CTA
// map would leave most of the GPU idle behind the slowest worlds.  Every CTA owns one scratch slab
// in global memory (A, the Schur complement, LU workspace: written once per step, L2 resident) and
// keeps what the pivot loop touches hundreds of times in shared memory:
//   * the LDL^T factor of the basic block A_SS as a packed lower triangle (k (k+1)/2 doubles;
//     k <= ~145 rows fit next to the vectors, larger blocks fall back to the CTA's global slab),
//   * x, w, the best-so-far pair, the right-hand side, the basic-set mask and index lists.
//
// Exactness.  The pivot sequence and the active set are discrete outputs that are compared
// bit-exactly with the oracle, so every value that feeds a comparison is computed with the
// oracle's operation order and this file is compiled with -fmad=false:
//   * Eigen's LDLT pivots on the largest |diagonal| of the NOT-YET-UPDATED trailing part
//     (orc_linalg.h LDLT::compute), so the whole pivot order is a function of the diagonal of the
//     input alone: it is computed up front (order_by_diag: a parallel rank when all |d| differ, the
//     swap-by-swap simulation otherwise), the lower triangle is gathered in pivoted order and the
//     factorisation runs without data movement.  Every entry is the same left-looking dot product
//     (ascending index, one accumulator) the oracle evaluates; rows run in parallel.
//   * sums that decide something (goodness of an iterate) skip only exact zeros, which cannot
//     change an IEEE sum that starts at +0.
//   * the Schur complement follows lcp.cc:293-294 literally: LU inverse of A_ee, (A_ie A_ee^-1) A_ei
//     with ascending-index dot products; structurally zero blocks of A_ie / A_ei are skipped (adding
//     +-0 products is a no-op).
//
// Deviation (documented in DESIGN.md): the reference decides "add cfm" from a JacobiSVD condition
// number (an O(rows^3) SVD with a large constant, every step).  Here the decision comes from a
// diagonally pivoted LDL^T of A: its pivot range is a lower bound of cond_2(A) for a PSD matrix
// (pivots are diagonal entries of Schur complements, which lie inside [lambda_min, lambda_max]),
// so a range >= 1e7 means ill conditioned for certain; below that, lambda_max and lambda_min are
// refined by power / inverse iteration with the same factor (cond_refine) until the estimate has
// converged or crossed 1e7.  tests/test_gpu_parity.py::test_cfm_decision_sweep sweeps cond(A)
// through 1e7 and reports where the two decisions part.
#include "egg_internal.cuh"
#include <cstdlib>

namespace {

constexpr int DT = 256;   // threads per CTA
constexpr int NWARP = DT / 32;
constexpr unsigned FULL = 0xffffffffu;
#define EGG_INF __longlong_as_double(0x7ff0000000000000LL)

struct DenseCfg {
  int Rcap;                // rows the path is provisioned for (multiple of 3)
  int ncap;                // constraints = Rcap / 3
  int Fcap;                // doubles of packed-factor storage in shared memory
  size_t off_A, off_Wk, off_Lm, off_T1, off_Fg, off_Jb, off_vec, off_int;   // offsets into the CTA slab (doubles)
  size_t per_cta;          // doubles per CTA slab
  size_t smem;             // dynamic shared memory bytes
};

constexpr int PB = 4;     // columns per panel of the factorisation

struct Sm {
  double *x, *w, *bx, *bw, *rhs, *xs;          // [Rcap] each
  double* tmp;                                 // [PB Rcap] scratch (panel temporaries of the factorisation; [Rcap] elsewhere)
  double* F;                                   // [Fcap] packed factor
  double* sa;                                  // [6 n] a = M^-1 J^T lambda
  unsigned short *sidx, *perm;                 // [Rcap]
  unsigned char *S, *athi;                     // [Rcap]
};

__device__ __forceinline__ size_t tri(int i) { return (size_t)i * (size_t)(i + 1) / 2; }

// Phase timing for development builds (EGG_DENSE_TIMING=1 python -m eggshell_b200.build --force):
// TICK(slot) charges the cycles since the previous TICK of this CTA to `slot`; the totals of all
// CTAs land in EggDev::dbg (egg_get_debug_counters).  Compiled out otherwise.
#ifdef EGG_DENSE_TIMING
__shared__ long long s_prof_t0;
__shared__ unsigned long long s_prof[32];
#define TICK(slot) do { __syncthreads(); if (threadIdx.x == 0) { const long long t__ = clock64(); s_prof[slot] += (unsigned long long)(t__ - s_prof_t0); s_prof_t0 = t__; } } while (0)
#else
#define TICK(slot) do { } while (0)
#endif
enum { T_ROWS = 0, T_CFM, T_SCHUR, T_CHECK, T_INDEX, T_ORDER, T_GATHER, T_FACTOR, T_SOLVE, T_W, T_BEST, T_XE, T_OUT, T_FA, T_FB, T_FC1, T_FC2, T_SFWD, T_SBWD, T_COUNT };

__device__ Sm carve(unsigned char* raw, const DenseCfg& c, int n) {
  Sm s;
  double* p = reinterpret_cast<double*>(raw);
  s.x = p; p += c.Rcap; s.w = p; p += c.Rcap; s.bx = p; p += c.Rcap; s.bw = p; p += c.Rcap;
  s.rhs = p; p += c.Rcap; s.xs = p; p += c.Rcap; s.tmp = p; p += (size_t)PB * c.Rcap;
  s.sa = p; p += 6 * n;
  s.F = p; p += c.Fcap;
  s.sidx = reinterpret_cast<unsigned short*>(p);
  s.perm = s.sidx + c.Rcap;
  s.S = reinterpret_cast<unsigned char*>(s.perm + c.Rcap);
  s.athi = s.S + c.Rcap;
  return s;
}

// ---- ordered compaction of the flags equal to `want` -> ascending index list; returns the count
__device__ int build_index_list(const unsigned char* flag, int want, int len, unsigned short* out) {
  __shared__ int s_cnt[NWARP];
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  int base = 0;
  for (int c0 = 0; c0 < len; c0 += DT) {
    const int i = c0 + tid;
    const bool f = i < len && (flag[i] != 0) == (want != 0);
    const unsigned m = __ballot_sync(FULL, f);
    if (lane == 0) s_cnt[wp] = __popc(m);
    __syncthreads();
    int off = base, tot = 0;
#pragma unroll
    for (int q = 0; q < NWARP; q++) { const int c = s_cnt[q]; if (q < wp) off += c; tot += c; }
    if (f) out[off + __popc(m & ((1u << lane) - 1u))] = (unsigned short)i;
    base += tot;
    __syncthreads();
  }
  return base;
}

// ---- pivot order of Eigen's LDLT from the |diagonal| alone -----------------------------------
// dabs[k] (shared) = |diagonal| by element; perm[pos] = element chosen at step pos.  Mirrors the
// selection of orc::LDLT::compute: at step kk the FIRST position >= kk holding the largest value
// wins and is swapped with position kk.  With all values distinct that is the descending sort;
// ties take the swap-by-swap simulation (warp 0; dabs is permuted along).
__device__ void order_by_diag(double* dabs, int k, unsigned short* perm) {
  const int tid = threadIdx.x;
  int tie = 0;
  for (int e = tid; e < k; e += DT) {
    const double v = dabs[e];
    int gt = 0;
    for (int f = 0; f < k; f++) {
      const double u = dabs[f];
      gt += (u > v);
      tie |= (u == v) && (f != e);
    }
    perm[gt] = (unsigned short)e;    // a permutation iff there are no ties
  }
  if (!__syncthreads_or(tie)) return;
  if (tid < 32) {
    const int lane = tid;
    for (int e = lane; e < k; e += 32) perm[e] = (unsigned short)e;
    __syncwarp();
    for (int kk = 0; kk < k; kk++) {
      double best = -1.0;
      int bp = k;
      for (int p = kk + lane; p < k; p += 32) {
        const double v = dabs[p];
        if (v > best) { best = v; bp = p; }      // ascending p per lane: first max of the lane
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(FULL, best, o);
        const int op = __shfl_xor_sync(FULL, bp, o);
        if (ob > best || (ob == best && op < bp)) { best = ob; bp = op; }
      }
      if (lane == 0 && bp != kk && bp < k) {
        const double tv = dabs[kk]; dabs[kk] = dabs[bp]; dabs[bp] = tv;
        const unsigned short te = perm[kk]; perm[kk] = perm[bp]; perm[bp] = te;
      }
      __syncwarp();
    }
  }
  __syncthreads();
}

// ---- packed LDL^T: gather, factor, solve ------------------------------------------------------
// The factor lives in shared memory when it fits and in the CTA's global slab otherwise; the
// routines are templates over a tiny accessor so that the shared-memory case compiles to ld.shared /
// st.shared (through a generic pointer every access would be a generic LD/ST and count as a
// long-scoreboard wait: profiles/r2c_dense_lines.txt).
struct ShMem {
  unsigned base;          // shared-space byte address of element 0
  __device__ __forceinline__ double ld(int i) const {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(base + 8u * (unsigned)i) : "memory");
    return v;
  }
  __device__ __forceinline__ void st(int i, double v) const { asm volatile("st.shared.f64 [%0], %1;" ::"r"(base + 8u * (unsigned)i), "d"(v) : "memory"); }
  __device__ __forceinline__ double2 ld2(int i) const {   // elements i, i + 1 (i even, base 16-byte aligned)
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(base + 8u * (unsigned)i) : "memory");
    return v;
  }
};
struct GlMem {
  double* p;
  __device__ __forceinline__ double ld(int i) const { return p[i]; }
  __device__ __forceinline__ void st(int i, double v) const { p[i] = v; }
};
__device__ __forceinline__ ShMem shmem(const double* p) { ShMem m; m.base = (unsigned)__cvta_generic_to_shared(p); return m; }
__device__ __forceinline__ int itri(int i) { return i * (i + 1) / 2; }

// Source matrix M (row-major, leading dimension ld, global); element e of the block is row/column
// idx[e] of M (idx == nullptr: e itself).  The factor of the symmetrically permuted block is built
// from M's LOWER triangle, exactly as Eigen's in-place swaps do (orc::LDLT::compute).
template <class FM>
__device__ void ldlt_gather(const double* __restrict__ M, int ld, const unsigned short* idx, const unsigned short* perm, int k, FM F) {
  const int total = itri(k);
  // entries are independent: four loads in flight per thread (the source is L2-resident, ~700 cycles away)
#pragma unroll 4
  for (int e = threadIdx.x; e < total; e += DT) {
    int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    i += (itri(i + 1) <= e) ? 1 : 0;          // the float estimate is off by at most one either way
    i -= (itri(i) > e) ? 1 : 0;
    const int j = e - itri(i);
    int a = perm[i], b = perm[j];
    if (idx) { a = idx[a]; b = idx[b]; }
    const int r = a > b ? a : b, c = a > b ? b : a;
    F.st(e, M[(size_t)r * ld + c]);
  }
  __syncthreads();
}

// In-place left-looking factorisation of the packed lower triangle F (k x k, k <= 2 DT): on exit F
// holds L strictly below the diagonal and D on it.  Per entry: t = sum_j fma(L[i][j], D[j] L[kk][j], t)
// for ascending j from 0 in one accumulator, then A[i][kk] - t, then / D[kk]: orc::LDLT::compute.
//
// The dot product of an entry is a serial chain, so the columns are processed in panels of PB:
//   A. temporaries D[j] L[c0+q][j] of the finished columns j < c0 for the PB panel columns;
//   B. the partial sums over j < c0 of every panel entry (i >= c0, q < PB): all chains of a panel
//      have the same length c0, so they are dealt out evenly over ALL threads, two chains per
//      thread at a time (rows do not belong to threads: with a row per thread the last warp would
//      carry most of the work and the other schedulers would idle);
//   C. the terms inside the panel: warp 0 finishes the PB x PB diagonal block with shuffles and
//      publishes D and the in-panel temporaries; one thread per row below finishes its PB entries.
// Four barriers per panel.  T = shared scratch [PB k]: temporaries [PB c0] then partial sums.
template <class FM>
__device__ void ldlt_factor(FM F, int k, ShMem T) {
  const int tid = threadIdx.x, lane = tid & 31;
  __shared__ double s_D[PB];              // pivots of the panel columns
  __shared__ double s_tin[PB][PB];        // in-panel temporaries: s_tin[q][q2] = D[c0+q2] L[c0+q][c0+q2], q2 < q
  __shared__ int s_stop;
  constexpr int RPT = 2;                  // rows per thread in phase C2: k <= RPT * DT
  if (tid == 0) s_stop = 0;
  for (int c0 = 0; c0 < k; c0 += PB) {
    const int nb = (k - c0 < PB) ? k - c0 : PB;
    const int accb = PB * c0;             // partial sums start here in T
    // A. temporaries
    for (int e = tid; e < PB * c0; e += DT) {
      const int jj = e / PB, q = e - jj * PB;
      T.st(e, (q < nb) ? F.ld(itri(jj) + jj) * F.ld(itri(c0 + q) + jj) : 0.0);
    }
    __syncthreads();
    TICK(T_FA);
    // B. partial sums: one row per thread with its PB chains in registers, so that one load of
    // L[i][j] feeds PB multiply-adds and the PB temporaries of column j come as two 128-bit
    // broadcast loads (two unrelated chains per thread needed two shared-memory loads per
    // multiply-add; C4 +6 %.  Requesting the next trip's operands before this trip's multiply-adds
    // was measured slower: 8.6 k against 11.0 k world-steps/s).  Every chain is still its own
    // accumulator over ascending j: the same bits.
    static_assert(PB == 4, "phase B reads the PB temporaries of a column as two double2");
    for (int i = c0 + tid; i < k; i += DT) {
      const int ri = itri(i);
      double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
      int jj = 0;
      for (; jj + 2 <= c0; jj += 2) {
        const double a0 = F.ld(ri + jj), a1 = F.ld(ri + jj + 1);
        const double2 t0 = T.ld2(jj * PB), t1 = T.ld2(jj * PB + 2), u0 = T.ld2((jj + 1) * PB), u1 = T.ld2((jj + 1) * PB + 2);
        s0 = __fma_rn(a0, t0.x, s0); s1 = __fma_rn(a0, t0.y, s1); s2 = __fma_rn(a0, t1.x, s2); s3 = __fma_rn(a0, t1.y, s3);
        s0 = __fma_rn(a1, u0.x, s0); s1 = __fma_rn(a1, u0.y, s1); s2 = __fma_rn(a1, u1.x, s2); s3 = __fma_rn(a1, u1.y, s3);
      }
      for (; jj < c0; jj++) {
        const double a0 = F.ld(ri + jj);
        const double2 t0 = T.ld2(jj * PB), t1 = T.ld2(jj * PB + 2);
        s0 = __fma_rn(a0, t0.x, s0); s1 = __fma_rn(a0, t0.y, s1); s2 = __fma_rn(a0, t1.x, s2); s3 = __fma_rn(a0, t1.y, s3);
      }
      const int o = accb + (i - c0) * PB;
      T.st(o, s0); T.st(o + 1, s1); T.st(o + 2, s2); T.st(o + 3, s3);
    }
    __syncthreads();
    TICK(T_FB);
    // C1. the diagonal block: rows c0 .. c0+nb-1 on lanes 0 .. nb-1 of warp 0
    if (tid < 32) {
      double myf[PB];                     // this lane's finished entries F[c0+lane][c0+q], q < lane
      double Dq[PB];
#pragma unroll
      for (int q = 0; q < PB; q++) { myf[q] = 0.0; Dq[q] = 0.0; }
#pragma unroll
      for (int q = 0; q < PB; q++) {
        if (q < nb) {
          // finish column c0+q for the block rows l >= q: terms q2 < q use row (c0+q)'s entries, held by lane q
          double t = (lane < nb) ? T.ld(accb + lane * PB + q) : 0.0;
#pragma unroll
          for (int q2 = 0; q2 < PB; q2++) {
            if (q2 < q) {
              const double tin = __shfl_sync(FULL, Dq[q2] * myf[q2], q);     // D[c0+q2] L[c0+q][c0+q2]
              t = __fma_rn(myf[q2], tin, t);
              if (lane == q) s_tin[q][q2] = tin;
            }
          }
          double val = 0.0;
          if (lane >= q && lane < nb) val = F.ld(itri(c0 + lane) + c0 + q) - t;
          const double akk = __shfl_sync(FULL, val, q);
          const bool valid = fabs(akk) > 0;
          Dq[q] = akk;
          if (lane == q) { F.st(itri(c0 + q) + c0 + q, val); s_D[q] = akk; if (c0 + q == 0 && !valid) s_stop = 1; }
          if (lane > q && lane < nb) {
            if (valid) val /= akk;
            F.st(itri(c0 + lane) + c0 + q, val);
            myf[q] = val;
          }
        }
      }
    }
    __syncthreads();
    TICK(T_FC1);
    if (s_stop) break;                     // orc::LDLT::compute: first pivot zero, matrix left as it is
    // C2. the rows below the block finish their panel entries on their own
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int i = c0 + nb + tid + r * DT;
      if (i < k) {
        const int rb = itri(i) + c0;
        double myf[PB];
#pragma unroll
        for (int q = 0; q < PB; q++) {
          if (q < nb) {
            double t = T.ld(accb + (i - c0) * PB + q);
#pragma unroll
            for (int q2 = 0; q2 < PB; q2++)
              if (q2 < q) t = __fma_rn(myf[q2], s_tin[q][q2], t);
            double val = F.ld(rb + q) - t;
            const double akk = s_D[q];
            if (fabs(akk) > 0) val /= akk;
            F.st(rb + q, val);
            myf[q] = val;
          }
        }
      }
    }
    __syncthreads();
    TICK(T_FC2);
  }
  __syncthreads();
}

// The plain column-by-column form of the same factorisation (any k; two barriers per column).
template <class FM>
__device__ void ldlt_factor_columns(FM F, int k, double* tmp) {
  const int tid = threadIdx.x;
  for (int kk = 0; kk < k; kk++) {
    const int rk = itri(kk);
    if (kk > 0) {
      // temp[j] = D[j] L[kk][j]; entry kk-1 was written by the thread that closed column kk-1 (below)
      for (int j = tid; j < kk - 1; j += DT) tmp[j] = F.ld(itri(j) + j) * F.ld(rk + j);
      __syncthreads();
      for (int i = kk + tid; i < k; i += DT) {
        const int ri = itri(i);
        double t = 0;
        for (int j = 0; j < kk; j++) t = __fma_rn(F.ld(ri + j), tmp[j], t);
        F.st(ri + kk, F.ld(ri + kk) - t);
      }
      __syncthreads();
    }
    const double akk = F.ld(rk + kk);
    const bool valid = fabs(akk) > 0;
    if (kk == 0 && !valid) break;                      // orc::LDLT::compute: matrix left as it is
    for (int i = kk + 1 + tid; i < k; i += DT) {
      double v = F.ld(itri(i) + kk);
      if (valid) { v /= akk; F.st(itri(i) + kk, v); }
      if (i == kk + 1) tmp[kk] = akk * v;              // the temp entry of the next column that needs this division
    }
    // the barrier at the top of the next column orders these writes before their use
  }
  __syncthreads();
}

// x <- L^-T D^+ L^-1 x for x in PIVOTED order (x[pos] belongs to element perm[pos]); the caller
// applies the transpositions by gathering / scattering through perm.  orc::LDLT::solve: forward,
// row i receives its terms for ascending j; the backward sweep is the column form (descending j);
// every term is one fma(-L, x_j, x_i).  Every thread keeps its rows' running values in registers;
// the unknowns are resolved in panels of PB: the warp that owns the panel's rows solves the
// PB x PB triangle with shuffles and publishes the values, everybody else applies the PB terms in
// order.  One barrier per panel.
template <class FM>
__device__ void ldlt_solve(FM F, int k, double* x) {
  const int tid = threadIdx.x, lane = tid & 31;
  constexpr int RPT = 2;                      // k <= RPT * DT, as in ldlt_factor
  if (k > RPT * DT) {                         // plain form: one barrier per unknown
    for (int j = 0; j + 1 < k; j++) {
      const double xj = x[j];
      for (int i = j + 1 + tid; i < k; i += DT) x[i] = __fma_rn(-F.ld(itri(i) + j), xj, x[i]);
      __syncthreads();
    }
    const double tol0 = 1.0 / 1.7976931348623157e308;
    for (int i = tid; i < k; i += DT) { const double dd = F.ld(itri(i) + i); x[i] = (fabs(dd) > tol0) ? x[i] / dd : 0.0; }
    __syncthreads();
    for (int j = k - 1; j >= 1; j--) {
      const double xj = x[j];
      const int rj = itri(j);
      for (int i = tid; i < j; i += DT) x[i] = __fma_rn(-F.ld(rj + i), xj, x[i]);
      __syncthreads();
    }
    return;
  }
  double xr[RPT];
#pragma unroll
  for (int r = 0; r < RPT; r++) { const int i = tid + r * DT; xr[r] = (i < k) ? x[i] : 0.0; }
  __syncthreads();
  // ---- L^-1 ----
  for (int c0 = 0; c0 < k; c0 += PB) {
    const int nb = (k - c0 < PB) ? k - c0 : PB;
    const int ro = c0 / DT;                   // which of this thread's rows the panel rows are (all PB rows: same r, same warp)
    const int t0 = c0 - ro * DT;              // thread that owns row c0
    if ((tid >> 5) == (t0 >> 5)) {
      const int l0 = t0 & 31;
#pragma unroll
      for (int r = 0; r < RPT; r++) {
        if (r == ro) {
#pragma unroll
          for (int q = 0; q < PB; q++) {
            if (q < nb) {
              const double xq = __shfl_sync(FULL, xr[r], l0 + q);
              const int me = lane - l0;       // my row is c0 + me
              if (me > q && me < nb) xr[r] = __fma_rn(-F.ld(itri(c0 + me) + c0 + q), xq, xr[r]);
              if (me == q) x[c0 + q] = xq;
            }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int i = tid + r * DT;
      if (i >= c0 + nb && i < k) {
        const int ri = itri(i) + c0;
#pragma unroll
        for (int q = 0; q < PB; q++)
          if (q < nb) xr[r] = __fma_rn(-F.ld(ri + q), x[c0 + q], xr[r]);
      }
    }
  }
  TICK(T_SFWD);
  // ---- D^+ ----
  const double tol = 1.0 / 1.7976931348623157e308;
#pragma unroll
  for (int r = 0; r < RPT; r++) {
    const int i = tid + r * DT;
    if (i < k) { const double dd = F.ld(itri(i) + i); xr[r] = (fabs(dd) > tol) ? xr[r] / dd : 0.0; }
  }
  // ---- L^-T, panels from the bottom ----
  const int last = ((k - 1) / PB) * PB;
  for (int c0 = last; c0 >= 0; c0 -= PB) {
    const int nb = (k - c0 < PB) ? k - c0 : PB;
    const int ro = c0 / DT;
    const int t0 = c0 - ro * DT;
    if ((tid >> 5) == (t0 >> 5)) {
      const int l0 = t0 & 31;
#pragma unroll
      for (int r = 0; r < RPT; r++) {
        if (r == ro) {
#pragma unroll
          for (int q = PB - 1; q >= 0; q--) {
            if (q < nb) {
              const double xq = __shfl_sync(FULL, xr[r], l0 + q);
              const int me = lane - l0;
              if (me >= 0 && me < q) xr[r] = __fma_rn(-F.ld(itri(c0 + q) + c0 + me), xq, xr[r]);
              if (me == q) x[c0 + q] = xq;
            }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; r++) {
      const int i = tid + r * DT;
      if (i < c0) {
#pragma unroll
        for (int q = PB - 1; q >= 0; q--)
          if (q < nb) xr[r] = __fma_rn(-F.ld(itri(c0 + q) + i), x[c0 + q], xr[r]);
      }
    }
  }
  __syncthreads();
  TICK(T_SBWD);
}

// Factor the k x k block of M selected by idx (ascending element list, or nullptr) -- in shared
// memory if it fits, else in the CTA's global slab -- and solve for the right-hand side held in
// pivoted order in sm.xs by the caller's `load` (called once the pivot order sm.perm is known).
template <class Load>
__device__ void ldlt_block_solve(const double* M, int ld, const unsigned short* idx, int k, const Sm& sm, const DenseCfg& cfg, double* Fg, Load load) {
  for (int e = threadIdx.x; e < k; e += DT) {
    const int g = idx ? idx[e] : e;
    sm.tmp[e] = fabs(M[(size_t)g * ld + g]);
  }
  __syncthreads();
  order_by_diag(sm.tmp, k, sm.perm);
  __syncthreads();
  TICK(T_ORDER);
  load();
  const ShMem T = shmem(sm.tmp);
  if ((size_t)tri(k) <= (size_t)cfg.Fcap) {
    const ShMem F = shmem(sm.F);
    ldlt_gather(M, ld, idx, sm.perm, k, F);
    TICK(T_GATHER);
    if (k <= 2 * DT) ldlt_factor(F, k, T); else ldlt_factor_columns(F, k, sm.tmp);
    TICK(T_FACTOR);
    ldlt_solve(F, k, sm.xs);
  } else {
    GlMem F; F.p = Fg;
    ldlt_gather(M, ld, idx, sm.perm, k, F);
    TICK(T_GATHER);
    if (k <= 2 * DT) ldlt_factor(F, k, T); else ldlt_factor_columns(F, k, sm.tmp);
    TICK(T_FACTOR);
    ldlt_solve(F, k, sm.xs);
  }
  TICK(T_SOLVE);
}

// ---- cfm decision ------------------------------------------------------------------------------
// Complete-diagonal-pivoted LDL^T of the symmetric PSD matrix M (R x R, global, both triangles
// kept up to date), right-looking.  Returns 0 = ill conditioned for certain (pivot range >= 1e7 or
// a non-positive pivot), 1 = factorisation completed; then *lb = pivot range, piv[] = pivot
// sequence, M holds the factor (column kk below the diagonal = L[:,kk] * d_kk, row kk right of the
// diagonal likewise, d on the diagonal) in pivoted coordinates.
__device__ int pivoted_ldlt_range(double* M, int R, unsigned short* piv, double* lb) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  __shared__ double s_val[NWARP];
  __shared__ int s_idx[NWARP];
  __shared__ double s_dmax, s_dmin;
  __shared__ int s_big;
  if (tid == 0) { s_dmax = 0; s_dmin = 1.7976931348623157e308; }
  __syncthreads();
  for (int kk = 0; kk < R; kk++) {
    double best = -1.7976931348623157e308;
    int bi = R;
    for (int i = kk + tid; i < R; i += DT) { const double v = M[(size_t)i * R + i]; if (v > best) { best = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(FULL, best, o);
      const int oi = __shfl_xor_sync(FULL, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) { s_val[wp] = best; s_idx[wp] = bi; }
    __syncthreads();
    if (tid == 0) {
      double b = s_val[0]; int ix = s_idx[0];
      for (int q = 1; q < NWARP; q++) if (s_val[q] > b || (s_val[q] == b && s_idx[q] < ix)) { b = s_val[q]; ix = s_idx[q]; }
      s_big = ix;
      piv[kk] = (unsigned short)ix;
      if (!(b > 0)) s_big = -1;
      else { if (b > s_dmax) s_dmax = b; if (b < s_dmin) s_dmin = b; }
    }
    __syncthreads();
    const int big = s_big;
    if (big < 0 || s_dmax >= 1e7 * s_dmin) return 0;
    if (big != kk) {   // full symmetric swap (both triangles kept)
      for (int j = tid; j < R; j += DT) { double t = M[(size_t)kk * R + j]; M[(size_t)kk * R + j] = M[(size_t)big * R + j]; M[(size_t)big * R + j] = t; }
      __syncthreads();
      for (int i = tid; i < R; i += DT) { double t = M[(size_t)i * R + kk]; M[(size_t)i * R + kk] = M[(size_t)i * R + big]; M[(size_t)i * R + big] = t; }
      __syncthreads();
    }
    // trailing update, one column j per thread (coalesced across the threads, the multiplier
    // column M[i][kk] is a broadcast)
    const double invd = 1.0 / M[(size_t)kk * R + kk];
    for (int j = kk + 1 + tid; j < R; j += DT) {
      const double ukj = M[(size_t)kk * R + j] * invd;
      for (int i = kk + 1; i < R; i++) M[(size_t)i * R + j] -= M[(size_t)i * R + kk] * ukj;
    }
    __syncthreads();
  }
  *lb = s_dmax / s_dmin;
  return 1;
}

// y = A x for the symmetric R x R matrix A in global memory (thread per row), x and y shared.
__device__ void sym_matvec(const double* __restrict__ A, int R, const double* x, double* y) {
  for (int i = threadIdx.x; i < R; i += DT) {
    double s = 0;
    for (int j = 0; j < R; j++) s += A[(size_t)j * R + i] * x[j];   // column i = row i (symmetric): coalesced over the threads
    y[i] = s;
  }
  __syncthreads();
}
__device__ double block_norm2(const double* v, int R, double* red) {
  const int tid = threadIdx.x, lane = tid & 31, wp = tid >> 5;
  double s = 0;
  for (int i = tid; i < R; i += DT) s += v[i] * v[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
  __syncthreads();
  if (lane == 0) red[wp] = s;
  __syncthreads();
  double t = 0;
  for (int q = 0; q < NWARP; q++) t += red[q];
  __syncthreads();
  return t;
}
// z <- (P^T L D L^T P)^-1 z with the factor left in M by pivoted_ldlt_range (pivoted coordinates:
// the swaps were applied to the matrix, so z is swapped along).
__device__ void pivoted_solve(const double* M, int R, const unsigned short* piv, double* z) {
  const int tid = threadIdx.x;
  if (tid == 0)
    for (int k = 0; k < R; k++) { const int p = piv[k]; if (p != k) { const double t = z[k]; z[k] = z[p]; z[p] = t; } }
  __syncthreads();
  for (int k = 0; k + 1 < R; k++) {                     // L^-1 (L[:,k] = M[:,k] / d_k)
    const double zk = z[k] / M[(size_t)k * R + k];
    for (int i = k + 1 + tid; i < R; i += DT) z[i] -= M[(size_t)k * R + i] * zk;   // row k right of the diagonal = column k below it
    __syncthreads();
  }
  for (int i = tid; i < R; i += DT) z[i] /= M[(size_t)i * R + i];
  __syncthreads();
  for (int k = R - 1; k >= 1; k--) {                    // L^-T
    const double zk = z[k];
    for (int i = tid; i < k; i += DT) z[i] -= (M[(size_t)i * R + k] / M[(size_t)i * R + i]) * zk;
    __syncthreads();
  }
  if (tid == 0)
    for (int k = R - 1; k >= 0; k--) { const int p = piv[k]; if (p != k) { const double t = z[k]; z[k] = z[p]; z[p] = t; } }
  __syncthreads();
}
// cond_2(A) = lambda_max / lambda_min by power iteration on A and inverse iteration on its factor.
// Both Rayleigh quotients approach their eigenvalue from inside the spectrum, so the running
// estimate grows towards cond_2(A): crossing 1e7 decides "ill conditioned"; convergence of both
// quotients (relative change < 1e-10 twice in a row) below 1e7 decides "well conditioned".
__device__ bool cond_refine_is_good(const double* A, const double* Mf, int R, const unsigned short* piv, double* u, double* v, double* y, double* red) {
  for (int i = threadIdx.x; i < R; i += DT) { u[i] = 1.0 + 0.37 * (double)((i * 7919) % 13); v[i] = 1.0 + 0.29 * (double)((i * 104729) % 11); }
  __syncthreads();
  double lmax = 0, lmin_inv = 0;
  int calm = 0;
  for (int it = 0; it < 400; it++) {
    // power step on A
    const double nu = sqrt(block_norm2(u, R, red));
    for (int i = threadIdx.x; i < R; i += DT) u[i] /= nu;
    __syncthreads();
    sym_matvec(A, R, u, y);
    double q = 0;
    { double s = 0; for (int i = threadIdx.x; i < R; i += DT) s += u[i] * y[i];
      const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      __syncthreads();
      if (lane == 0) red[wp] = s;
      __syncthreads();
      for (int k = 0; k < NWARP; k++) q += red[k];
      __syncthreads(); }
    for (int i = threadIdx.x; i < R; i += DT) u[i] = y[i];
    __syncthreads();
    // inverse step on the factor: Rayleigh quotient of A^-1
    const double nv = sqrt(block_norm2(v, R, red));
    for (int i = threadIdx.x; i < R; i += DT) { v[i] /= nv; y[i] = v[i]; }
    __syncthreads();
    pivoted_solve(Mf, R, piv, y);
    double qi = 0;
    { double s = 0; for (int i = threadIdx.x; i < R; i += DT) s += v[i] * y[i];
      const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
      __syncthreads();
      if (lane == 0) red[wp] = s;
      __syncthreads();
      for (int k = 0; k < NWARP; k++) qi += red[k];
      __syncthreads(); }
    for (int i = threadIdx.x; i < R; i += DT) v[i] = y[i];
    __syncthreads();
    const bool settled = fabs(q - lmax) <= 1e-10 * fabs(q) && fabs(qi - lmin_inv) <= 1e-10 * fabs(qi);
    lmax = q; lmin_inv = qi;
    if (lmax * lmin_inv >= 1e7) return false;
    calm = settled ? calm + 1 : 0;
    if (calm >= 2) return true;
  }
  return lmax * lmin_inv < 1e7;
}

// ---- lcp.cc:20-94 -------------------------------------------------------------------------------
// Bounds of inequality row i: [0, inf) under quirk q1 (and on every normal row), [-1, 1] on the
// two tangential rows of a contact otherwise.
__device__ __forceinline__ double row_lo(int i, bool boxed) { return (boxed && (i % 3) < 2) ? -1.0 : 0.0; }
__device__ __forceinline__ double row_hi(int i, bool boxed) { return (boxed && (i % 3) < 2) ? 1.0 : EGG_INF; }

// Returns 1 = solution, 0 = not a solution (S flipped at the least offending index, if any).
__device__ int check_murty(const double* __restrict__ A, int dim, const Sm& sm, const double* x, const double* w, bool boxed, double err) {
  const int tid = threadIdx.x;
  __shared__ int s_first, s_result;
  if (tid == 0) { s_first = dim; s_result = -1; }
  __syncthreads();
  int mine = dim;
  for (int i = tid; i < dim; i += DT) {
    const double lo = row_lo(i, boxed), hi = row_hi(i, boxed);
    bool viol;
    if (sm.S[i]) viol = (x[i] < lo) || (x[i] > hi);
    else viol = (!sm.athi[i] && w[i] < 0) || (sm.athi[i] && w[i] > 0);
    if (viol) { mine = i; break; }
  }
  if (mine < dim) atomicMin(&s_first, mine);
  __syncthreads();
  const int first = s_first;
  if (first < dim) {
    if (tid == 0) {
      if (sm.S[first]) { sm.S[first] = 0; sm.athi[first] = (x[first] < row_lo(first, boxed)) ? 0 : 1; }
      else sm.S[first] = 1;
    }
    __syncthreads();
    return 0;
  }
  int bad = 0;
  for (int i = tid; i < dim; i += DT) {
    const double lo = row_lo(i, boxed), hi = row_hi(i, boxed);
    if (x[i] < lo || x[i] > hi) bad = 1;
    if (x[i] == lo && w[i] < 0) bad = 1;
    if (x[i] == hi && w[i] > 0) bad = 1;
  }
  if (__syncthreads_or(bad)) return 0;
  for (int i = tid; i < dim; i += DT) {
    double s = 0;
    for (int j = 0; j < dim; j++) s += A[(size_t)i * dim + j] * x[j];
    sm.tmp[i] = s - (sm.rhs[i] + w[i]);
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0;
    for (int i = 0; i < dim; i++) s += sm.tmp[i] * sm.tmp[i];
    const double chk = fabs(err) > 1e-9 ? fabs(err) : 1e-9;
    s_result = (sqrt(s) > chk) ? 0 : 1;
  }
  __syncthreads();
  return s_result;
}

// lcp.cc:98-104: sum of the non-positive parts of v in index order; exact zeros and positive
// entries contribute +0 and are skipped (the sum starts at +0 and can never become -0).
__device__ double ordered_nonpositive_sum(const double* v, int len, unsigned short* list, unsigned char* flag) {
  for (int i = threadIdx.x; i < len; i += DT) flag[i] = (v[i] < 0) ? 1 : 0;
  __syncthreads();
  const int cnt = build_index_list(flag, 1, len, list);
  __shared__ double s_sum;
  if (threadIdx.x == 0) {
    double s = 0;
    for (int q = 0; q < cnt; q++) s += v[list[q]];
    s_sum = s;
  }
  __syncthreads();
  const double r = s_sum;
  __syncthreads();
  return r;
}

// ---- integrate (same arithmetic as the PGS path): v' = v + dt (M^-1 f + a), midpoint position ---
__device__ void integrate_world(const EggDev& d, int w, double dt, const double* sa) {
  const int n = d.n;
  const double* st = d.stat + (size_t)w * EGG_STAT * n;
  double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  for (int bd = threadIdx.x; bd < n; bd += DT) {
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

// Jacobian blocks J0, J1 (3x6 each) of every constraint in REFERENCE order from the per-world
// records (which are in level order); zero_world: write zeros on the world / ground side.
__device__ void jacobian_blocks(const EggDev& d, int w, int nc, double* Jb, int* ci0, int* ci1, double* rhs_out, bool zero_world) {
  const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
  for (int s = threadIdx.x; s < nc; s += DT) {
    const double* r = recs + (size_t)s * EGG_REC;
    const int i0 = __double2loint(r[REC_IDX]), i1 = __double2hiint(r[REC_IDX]);
    const int c = __double2loint(r[REC_META]);
    ci0[c] = i0; ci1[c] = i1;
    d3 r0 = mk3(r[REC_R0], r[REC_R0 + 1], r[REC_R0 + 2]), r1 = mk3(r[REC_R1], r[REC_R1 + 1], r[REC_R1 + 2]);
    double* J0 = Jb + (size_t)c * 72;
    double* J1 = J0 + 18;
    const double z0 = (zero_world && i0 < 0) ? 0.0 : 1.0, z1 = (zero_world && i1 < 0) ? 0.0 : 1.0;
    for (int k = 0; k < 3; k++) {
      d3 rc = mk3(r[REC_RC + 3 * k], r[REC_RC + 3 * k + 1], r[REC_RC + 3 * k + 2]);
      d3 a0 = cross3(rc, r0), a1 = cross3(r1, rc);
      J0[6 * k] = -rc.x * z0; J0[6 * k + 1] = -rc.y * z0; J0[6 * k + 2] = -rc.z * z0; J0[6 * k + 3] = a0.x * z0; J0[6 * k + 4] = a0.y * z0; J0[6 * k + 5] = a0.z * z0;
      J1[6 * k] = rc.x * z1; J1[6 * k + 1] = rc.y * z1; J1[6 * k + 2] = rc.z * z1; J1[6 * k + 3] = a1.x * z1; J1[6 * k + 4] = a1.y * z1; J1[6 * k + 5] = a1.z * z1;
      if (rhs_out) rhs_out[3 * c + k] = r[REC_RHS + k];
    }
  }
}

// Block (c, c2) of G = X J^T with X = J M^-1 (use_minv) or X = J: sum over the shared bodies in
// ascending column (= body) order, as the dense products of the oracle do.
__device__ __forceinline__ void pair_block(const double* Jb, const int* ci0, const int* ci1, int c, int c2, bool use_minv, double* blk) {
  const int a0 = ci0[c], a1 = ci1[c], e0 = ci0[c2], e1 = ci1[c2];
  for (int q = 0; q < 9; q++) blk[q] = 0.0;
  for (int pass = 0; pass < 2; pass++) {
    const int side = ((a0 <= a1) == (pass == 0)) ? 0 : 1;
    const int bd = side ? a1 : a0;
    if (bd < 0) continue;
    int side2 = -1;
    if (bd == e0) side2 = 0; else if (bd == e1) side2 = 1;
    if (side2 < 0) continue;
    const double* Bx = Jb + (size_t)c * 72 + (use_minv ? 36 : 0) + 18 * side;
    const double* Jy = Jb + (size_t)c2 * 72 + 18 * side2;
    for (int k = 0; k < 3; k++)
      for (int l = 0; l < 3; l++) {
        double s2 = blk[3 * k + l];
        for (int q = 0; q < 6; q++) s2 += Bx[6 * k + q] * Jy[6 * l + q];
        blk[3 * k + l] = s2;
      }
  }
}
__device__ __forceinline__ bool shares_body(const int* ci0, const int* ci1, int c, int c2) {
  const int a0 = ci0[c], a1 = ci1[c], e0 = ci0[c2], e1 = ci1[c2];
  return (a0 >= 0 && (a0 == e0 || a0 == e1)) || (a1 >= 0 && (a1 == e0 || a1 == e1));
}

// ---- the step kernel ----------------------------------------------------------------------------
__global__ void __launch_bounds__(DT, 2) egg_dense_kernel(EggDev d, double dt, double* scratch, DenseCfg cfg) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int tid = threadIdx.x, n = d.n, nj = d.nj;
  const Sm sm = carve(sm_raw, cfg, n);
  __shared__ int s_world, s_flag;
  __shared__ double s_red[NWARP];
  double* slab = scratch + (size_t)blockIdx.x * cfg.per_cta;
  double* A = slab + cfg.off_A;
  double* Wk = slab + cfg.off_Wk;
  double* Lm = slab + cfg.off_Lm;
  double* T1 = slab + cfg.off_T1;
  double* Fg = slab + cfg.off_Fg;
  double* Jb = slab + cfg.off_Jb;
  double* b = slab + cfg.off_vec;                 // rhs [R]
  double* lamv = b + cfg.Rcap;                    // lambda [R], reference row order
  int* ci0 = reinterpret_cast<int*>(slab + cfg.off_int);
  int* ci1 = ci0 + cfg.ncap;
  unsigned short* piv = reinterpret_cast<unsigned short*>(ci1 + cfg.ncap);   // [Rcap]
  const bool boxed = (d.prm.quirks & 2) == 0;     // q1 off: BOX friction bounds honoured

  while (true) {
    if (tid == 0) s_world = atomicAdd(d.work_ctr + 1, 1);
    __syncthreads();
    const int w = s_world;
    __syncthreads();
    if (w >= d.W) break;
#ifdef EGG_DENSE_TIMING
    if (tid == 0) { s_prof_t0 = clock64(); for (int q = 0; q < 32; q++) s_prof[q] = 0; }
    __syncthreads();
#endif

    const int nc = nj + d.c_count[w];
    int R = 3 * nc;
    const int E = 3 * nj;
    const double* st = d.stat + (size_t)w * EGG_STAT * n;
    double* lam_out = d.lam_out + (size_t)w * 3 * d.nrec;
    int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
    int* stt = d.stats + (size_t)w * 8;
    for (int i = tid; i < 6 * n; i += DT) sm.sa[i] = 0.0;
    bool overflow = false;
    if (R > cfg.Rcap) { overflow = true; R = 0; }
    const int I = R - (R ? E : 0);
    int pivots = 0, cfm_applied = 0, lcp_failed = 0;
    // FP64 operations of the reference algorithm on this world (mul and add counted separately):
    // A = J M^-1 J^T is not counted (sparse, O(rows)); cfm decision = one factorisation R^3/3;
    // A_ee^-1 by LU 2 E^3; Schur products 2 (I E^2 + I^2 E) as dense GEMMs; per pivot LDL^T k^3/3 +
    // two triangular solves 2 k^2 + w = A_NS x_S 2 k (I - k); final A_ee LDL^T + solve.
    double work = 0.0;

    if (R > 0) {
      // ---- 1. Jacobian blocks, J M^-1 blocks and rhs in reference row order ----
      jacobian_blocks(d, w, nc, Jb, ci0, ci1, b, false);
      __syncthreads();
      for (int c = tid; c < nc; c += DT) {
        for (int side = 0; side < 2; side++) {
          const int bd = side ? ci1[c] : ci0[c];
          const double* Jx = Jb + (size_t)c * 72 + 18 * side;
          double* Bx = Jb + (size_t)c * 72 + 36 + 18 * side;
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
      }
      __syncthreads();

      // ---- 2. A = J M^-1 J^T (ensembles.cc:510) ----
      for (int e = tid; e < nc * nc; e += DT) {
        const int c = e / nc, c2 = e % nc;
        double blk[9];
        pair_block(Jb, ci0, ci1, c, c2, true, blk);
        for (int k = 0; k < 3; k++)
          for (int l = 0; l < 3; l++) A[(size_t)(3 * c + k) * R + 3 * c2 + l] = blk[3 * k + l];
      }
      __syncthreads();

      TICK(T_ROWS);
      // ---- 3. cfm decision (ensembles.cc:513-521) ----
      bool good;
      if (d.prm.cfm_mode == 1) good = false;
      else if (d.prm.cfm_mode == 2) good = true;
      else {
        // Two contacts between the same two bodies (or of one body with the ground) make J rank
        // deficient whatever else is in the world: the 6 rows only see the relative twist, and the
        // relative velocities of two points of a rigid motion differ by w x (p2 - p1), which has no
        // component along p2 - p1.  Then A = J M^-1 J^T is singular, the SVD condition number is
        // ~1e16 or inf (far beyond 1e7) and cfm is added: no factorisation needed.  Contacts of one
        // pair are consecutive in the list (ensembles.cc:449-473).
        int twin = 0;
        for (int c = nj + tid; c + 1 < nc; c += DT) twin |= (ci0[c] == ci0[c + 1]) && (ci1[c] == ci1[c + 1]);
        work += (double)R * R * R / 3.0;               // the reference's decision costs this much (and more) either way
        if (__syncthreads_or(twin)) good = false;
        else {
          for (int e = tid; e < R * R; e += DT) Wk[e] = A[e];
          __syncthreads();
          double lb = 0;
          good = pivoted_ldlt_range(Wk, R, piv, &lb) != 0;
          __syncthreads();
          if (good && lb >= 1e4) good = cond_refine_is_good(A, Wk, R, piv, sm.x, sm.w, sm.xs, s_red);
          __syncthreads();
        }
      }
      if (!good) {
        for (int i = tid; i < R; i += DT) A[(size_t)i * R + i] += d.prm.cfm;
        cfm_applied = 1;
        __syncthreads();
      }

      TICK(T_CFM);
      // ---- 4. Schur complement on the equality rows (lcp.cc:286-294); rows 0..E-1 are the joints ----
      if (I > 0) {
        if (E > 0) {
          // 4a. A_ee^-1 by partial-pivot LU (orc::lu_inverse): lu in Wk[0 .. E^2), inverse in Wk[E^2 .. 2 E^2)
          work += 2.0 * E * E * E + 2.0 * ((double)I * E * E + (double)I * I * E);
          double* lu = ((size_t)E * E <= (size_t)cfg.Fcap) ? sm.F : Wk;     // the LU lives in shared memory when it fits
          double* Ainv = Wk + (size_t)E * E;
          unsigned short* lperm = sm.sidx;           // row permutation of the LU (free until the Murty loop)
          for (int e = tid; e < E * E; e += DT) lu[e] = A[(size_t)(e / E) * R + e % E];
          for (int i = tid; i < E; i += DT) lperm[i] = (unsigned short)i;
          __syncthreads();
          for (int k = 0; k < E; k++) {
            // pivot: first row i >= k with the largest |lu(i,k)|
            __shared__ double s_pv[NWARP];
            __shared__ int s_pi[NWARP];
            double best = -1.0;
            int bi = E;
            for (int i = k + tid; i < E; i += DT) { const double v = fabs(lu[(size_t)i * E + k]); if (v > best) { best = v; bi = i; } }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const double ob = __shfl_xor_sync(FULL, best, o);
              const int oi = __shfl_xor_sync(FULL, bi, o);
              if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if ((tid & 31) == 0) { s_pv[tid >> 5] = best; s_pi[tid >> 5] = bi; }
            __syncthreads();
            if (tid == 0) {
              double bb = s_pv[0]; int ix = s_pi[0];
              for (int q = 1; q < NWARP; q++) if (s_pv[q] > bb || (s_pv[q] == bb && s_pi[q] < ix)) { bb = s_pv[q]; ix = s_pi[q]; }
              s_flag = ix;
            }
            __syncthreads();
            const int pv = s_flag;
            if (pv != k) {
              for (int j = tid; j < E; j += DT) { const double t = lu[(size_t)k * E + j]; lu[(size_t)k * E + j] = lu[(size_t)pv * E + j]; lu[(size_t)pv * E + j] = t; }
              if (tid == 0) { const unsigned short t = lperm[k]; lperm[k] = lperm[pv]; lperm[pv] = t; }
              __syncthreads();
            }
            const double dd = lu[(size_t)k * E + k];
            if (dd != 0) {
              for (int i = k + 1 + tid; i < E; i += DT) lu[(size_t)i * E + k] /= dd;
              __syncthreads();
            }
            // trailing update: thread per column j (coalesced), multiplier column broadcast
            for (int j = k + 1 + tid; j < E; j += DT) {
              const double ukj = lu[(size_t)k * E + j];
              for (int i = k + 1; i < E; i++) {
                const double l = lu[(size_t)i * E + k];
                if (l == 0) continue;
                lu[(size_t)i * E + j] -= l * ukj;
              }
            }
            __syncthreads();
          }
          // inverse, one thread per column c; the column lives in Ainv(:, c) (coalesced over the threads)
          for (int c = tid; c < E; c += DT) {
            for (int i = 0; i < E; i++) {
              double s = (lperm[i] == c) ? 1.0 : 0.0;
              for (int j = 0; j < i; j++) s -= lu[(size_t)i * E + j] * Ainv[(size_t)j * E + c];
              Ainv[(size_t)i * E + c] = s;
            }
            for (int i = E - 1; i >= 0; i--) {
              double s = Ainv[(size_t)i * E + c];
              for (int j = i + 1; j < E; j++) s -= lu[(size_t)i * E + j] * Ainv[(size_t)j * E + c];
              Ainv[(size_t)i * E + c] = s / lu[(size_t)i * E + i];
            }
          }
          __syncthreads();
          // 4b. T1 = A_ie A_ee^-1  (I x E); only joints that share a body with the contact of row i
          //     have non-zero entries in A_ie (zero entries are skipped by the oracle's matmul as well)
          for (int e = tid; e < I * E; e += DT) {
            const int i = e / E, j = e % E;
            const int c = nj + i / 3;
            double s = 0;
            for (int q = 0; q < nj; q++) {
              if (!shares_body(ci0, ci1, c, q)) continue;
              for (int t = 0; t < 3; t++) {
                const int k = 3 * q + t;
                const double aik = A[(size_t)(E + i) * R + k];
                if (aik == 0) continue;
                s += aik * Ainv[(size_t)k * E + j];
              }
            }
            T1[e] = s;
          }
          __syncthreads();
          // 4c. lhs = A_ii - T1 A_ei ; rhs = b_i - T1 b_e
          for (int e = tid; e < I * (I + 1); e += DT) {
            const int i = e / (I + 1), j = e % (I + 1);
            double s = 0;
            if (j < I) {
              const int c2 = nj + j / 3;
              for (int q = 0; q < nj; q++) {
                if (!shares_body(ci0, ci1, c2, q)) continue;       // A_ei(3q.., j) structurally zero otherwise
                for (int t = 0; t < 3; t++) {
                  const int k = 3 * q + t;
                  const double aik = T1[(size_t)i * E + k];
                  if (aik == 0) continue;
                  s += aik * A[(size_t)k * R + E + j];
                }
              }
              Lm[(size_t)i * I + j] = A[(size_t)(E + i) * R + E + j] - s;
            } else {
              for (int k = 0; k < E; k++) s += T1[(size_t)i * E + k] * b[k];
              sm.rhs[i] = b[E + i] - s;
            }
          }
        } else {
          for (int e = tid; e < I * I; e += DT) Lm[e] = A[e];
          for (int i = tid; i < I; i += DT) sm.rhs[i] = b[i];
        }
        __syncthreads();

        TICK(T_SCHUR);
        // ---- 5. Murty principal pivoting on (Lm, rhs) ----
        for (int i = tid; i < I; i += DT) { sm.S[i] = 1; sm.athi[i] = 0; sm.x[i] = 0.0; sm.w[i] = -sm.rhs[i]; sm.bx[i] = 0.0; sm.bw[i] = -sm.rhs[i]; }
        __syncthreads();
        // goodness of the best-so-far pair (lcp.cc:98-104), kept instead of recomputed
        double gbest = ordered_nonpositive_sum(sm.bx, I, sm.perm, reinterpret_cast<unsigned char*>(sm.tmp));
        gbest = gbest + ordered_nonpositive_sum(sm.bw, I, sm.perm, reinterpret_cast<unsigned char*>(sm.tmp));
        const int max_it = (I >= 10) ? 1000 : (1 << I);
        int iter = 0;
        while (iter < max_it) {
          if (check_murty(Lm, I, sm, sm.x, sm.w, boxed, 0.0)) break;
          TICK(T_CHECK);
          const int ks = build_index_list(sm.S, 1, I, sm.sidx);
          TICK(T_INDEX);
          work += (double)ks * ks * ks / 3.0 + 2.0 * ks * ks + 2.0 * ks * (double)(I - ks);
          if (ks > 0) {
            ldlt_block_solve(Lm, I, sm.sidx, ks, sm, cfg, Fg, [&]() { for (int p = tid; p < ks; p += DT) sm.xs[p] = sm.rhs[sm.sidx[sm.perm[p]]]; });
            for (int p = tid; p < ks; p += DT) sm.x[sm.sidx[sm.perm[p]]] = sm.xs[p];
          }
          for (int i = tid; i < I; i += DT)
            if (!sm.S[i]) sm.x[i] = sm.athi[i] ? row_hi(i, boxed) : row_lo(i, boxed);
          __syncthreads();
          // w_N = A_NS x_S - b_N (lcp.cc:219-226): row i sums A(i, j) x(j) over the basic j in ascending
          // order.  The products are formed by all threads (independent, coalesced L2 loads: nothing
          // waits on the sum) into the free tail of the factor storage, chunk by chunk; one thread
          // per non-basic row then adds its chunk in order.
          {
            const int nn = build_index_list(sm.S, 0, I, sm.perm);      // non-basic rows, ascending (perm is free again)
            double* P = sm.F + (((size_t)tri(ks) <= (size_t)cfg.Fcap) ? tri(ks) : 0);
            const int room = cfg.Fcap - (int)(P - sm.F);
            int cw = nn > 0 ? room / nn : ks;                           // basic columns per chunk
            if (cw > ks) cw = ks;
            for (int q = tid; q < nn; q += DT) sm.tmp[q] = 0.0;        // running sums
            for (int i = tid; i < I; i += DT) if (sm.S[i]) sm.w[i] = 0.0;
            __syncthreads();
            if (cw >= 8) {
              for (int k0 = 0; k0 < ks; k0 += cw) {
                const int kc = (ks - k0 < cw) ? ks - k0 : cw;
#pragma unroll 4
                for (int e = tid; e < nn * kc; e += DT) {
                  const int q = e / kc, k = e - q * kc;
                  const int g = sm.sidx[k0 + k];
                  P[e] = Lm[(size_t)sm.perm[q] * I + g] * sm.x[g];
                }
                __syncthreads();
                for (int q = tid; q < nn; q += DT) {
                  double s2 = sm.tmp[q];
                  const double* pq = P + (size_t)q * kc;
                  for (int k = 0; k < kc; k++) s2 += pq[k];
                  sm.tmp[q] = s2;
                }
                __syncthreads();
              }
            } else {                                                    // no room to stage products: thread per row
              for (int q = tid; q < nn; q += DT) {
                const double* Li = Lm + (size_t)sm.perm[q] * I;
                double s2 = 0;
                for (int k = 0; k < ks; k++) { const int g = sm.sidx[k]; s2 += Li[g] * sm.x[g]; }
                sm.tmp[q] = s2;
              }
              __syncthreads();
            }
            for (int q = tid; q < nn; q += DT) { const int i = sm.perm[q]; sm.w[i] = sm.tmp[q] - sm.rhs[i]; }
            __syncthreads();
          }
          TICK(T_W);
          // UpdatePreviousBestSolution (lcp.cc:127-137)
          int differs = 0;
          for (int i = tid; i < I; i += DT) differs |= (sm.x[i] != sm.bx[i]) || (sm.w[i] != sm.bw[i]);
          if (__syncthreads_or(differs)) {
            double g = ordered_nonpositive_sum(sm.x, I, sm.perm, reinterpret_cast<unsigned char*>(sm.tmp));
            g = g + ordered_nonpositive_sum(sm.w, I, sm.perm, reinterpret_cast<unsigned char*>(sm.tmp));
            if (g > gbest) {
              gbest = g;
              for (int i = tid; i < I; i += DT) { sm.bx[i] = sm.x[i]; sm.bw[i] = sm.w[i]; }
            }
            __syncthreads();
          }
          TICK(T_BEST);
          ++iter;
        }
        pivots = iter;
        for (int i = tid; i < I; i += DT) { sm.x[i] = sm.bx[i]; sm.w[i] = sm.bw[i]; }
        __syncthreads();
        if (!check_murty(Lm, I, sm, sm.x, sm.w, boxed, (iter >= max_it) ? 1e-8 : 0.0)) lcp_failed = 1;
      }

      TICK(T_CHECK);
      // ---- 6. x_e = A_ee.ldlt().solve(b_e - A_ei x_i)  (lcp.cc:317) ----
      if (E > 0) {
        for (int i = tid; i < E; i += DT) {
          double s2 = 0;
          for (int j = 0; j < I; j++) s2 += A[(size_t)i * R + E + j] * sm.x[j];
          sm.w[i] = b[i] - s2;                 // (w is free now)
        }
        __syncthreads();
        work += (double)E * E * E / 3.0 + 2.0 * E * E + 2.0 * E * I;
        ldlt_block_solve(A, R, nullptr, E, sm, cfg, Fg, [&]() { for (int p = tid; p < E; p += DT) sm.xs[p] = sm.w[sm.perm[p]]; });
        for (int p = tid; p < E; p += DT) lamv[sm.perm[p]] = sm.xs[p];
      }
      for (int i = tid; i < I; i += DT) lamv[E + i] = sm.x[i];
      __syncthreads();

      TICK(T_XE);
      // ---- 7. outputs + a = M^-1 J^T lambda (per body, constraints in reference order) ----
      for (int i = tid; i < R; i += DT) {
        lam_out[i] = lamv[i];
        rs_out[i] = (i < E) ? 3 : (sm.S[i - E] ? 0 : 1);
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
        sm.sa[bd] = al.x; sm.sa[n + bd] = al.y; sm.sa[2 * n + bd] = al.z;
        sm.sa[3 * n + bd] = aa.x; sm.sa[4 * n + bd] = aa.y; sm.sa[5 * n + bd] = aa.z;
      }
    }
    __syncthreads();

    if (tid == 0) {
      stt[4] = 0;
      stt[5] = pivots;
      stt[6] = cfm_applied;
      stt[7] = 0;
      d.resid[w] = 0.0;
      if (d.work) d.work[w] = work;
      int flags = d.status[w] & ~1;
      if (lcp_failed) flags |= 1;
      if (overflow) flags |= 32;
      d.status[w] = flags;
    }
    __syncthreads();
    integrate_world(d, w, dt, sm.sa);
    __syncthreads();
#ifdef EGG_DENSE_TIMING
    TICK(T_OUT);
    if (tid < T_COUNT) atomicAdd(d.dbg + tid, s_prof[tid]);
    if (tid == 0) atomicAdd(d.dbg + 31, 1ull);
    __syncthreads();
#endif
  }
}

// ---------------------------------------------------------------------------------------------
// Position relaxation.  One iteration of the loop body of Ensemble::InitStabilize
// (ensembles.cc:602-622; mode 0: StepPositionRelaxation = StepPositions_ExplicitEuler(dt, c)) or of
// Ensemble::PostStabilize (ensembles.cc:624-645; mode 1: StepPostStabilization = the same position
// step plus v += c) with c = -step_scale J^T (J J^T)^-1 err (ensembles.cc:659-666), for every world
// whose squared position error still exceeds 1e-9 and whose step counter is below max_steps.
// The records must have been assembled for the current state (and, for mode 0, the contacts
// refreshed WITHOUT de-duplication: the reference only runs CheckAndCorrectEnsembleState after the
// loop).  prog[w] = {steps taken, active flag}; err2[w] = squared error seen by this call.
__global__ void __launch_bounds__(DT, 2) egg_relax_kernel(EggDev d, double dt, double step_scale, int max_steps, int mode, double* scratch,
                                                           DenseCfg cfg, int* prog, double* err2, int* any_active) {
  extern __shared__ __align__(16) unsigned char sm_raw[];
  const int tid = threadIdx.x, n = d.n, nj = d.nj;
  const Sm sm = carve(sm_raw, cfg, n);
  __shared__ int s_world, s_go;
  double* slab = scratch + (size_t)blockIdx.x * cfg.per_cta;
  double* A = slab + cfg.off_A;
  double* Fg = slab + cfg.off_Fg;
  double* Jb = slab + cfg.off_Jb;
  int* ci0 = reinterpret_cast<int*>(slab + cfg.off_int);
  int* ci1 = ci0 + cfg.ncap;
  double* e = sm.x;                // err [R], reference row order

  while (true) {
    if (tid == 0) s_world = atomicAdd(d.work_ctr + 1, 1);
    __syncthreads();
    const int w = s_world;
    __syncthreads();
    if (w >= d.W) break;
    const int nc = nj + d.c_count[w];
    const int R = 3 * nc;
    const bool fits = R <= cfg.Rcap;
    const double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
    const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
    const double* geom = d.c_geom + (size_t)w * 7 * d.maxc;

    if (fits) jacobian_blocks(d, w, nc, Jb, ci0, ci1, nullptr, true);
    __syncthreads();
    // position error per constraint, reference order (joints.cc:3-11, contact.cc:14-22)
    for (int s = tid; fits && s < nc; s += DT) {
      const double* r = recs + (size_t)s * EGG_REC;
      const int i0 = __double2loint(r[REC_IDX]), i1 = __double2hiint(r[REC_IDX]);
      const int c = __double2loint(r[REC_META]);
      if (c < nj) {
        d3 r0 = mk3(r[REC_R0], r[REC_R0 + 1], r[REC_R0 + 2]), r1 = mk3(r[REC_R1], r[REC_R1 + 1], r[REC_R1 + 2]);
        d3 p0 = mk3(dyn[i0], dyn[n + i0], dyn[2 * n + i0]);
        d3 er;
        if (i1 < 0) {
          const double* jc = d.jc + (size_t)w * 6 * nj;
          er = p0 + r0 - mk3(jc[3 * nj + c], jc[4 * nj + c], jc[5 * nj + c]);
        } else {
          er = p0 + r0 - mk3(dyn[i1], dyn[n + i1], dyn[2 * n + i1]) - r1;
        }
        e[3 * c] = er.x; e[3 * c + 1] = er.y; e[3 * c + 2] = er.z;
      } else {
        e[3 * c] = 0.0; e[3 * c + 1] = 0.0; e[3 * c + 2] = -geom[6 * d.maxc + (c - nj)];
      }
    }
    __syncthreads();
    if (tid == 0) {
      double s2 = 0;
      for (int i = 0; fits && i < R; i++) s2 += e[i] * e[i];
      const int steps = prog[2 * w];
      s_go = (fits && s2 > 1e-9 && steps < max_steps) ? 1 : 0;
      err2[w] = s2;
      prog[2 * w + 1] = s_go;
      if (s_go) { prog[2 * w] = steps + 1; atomicOr(any_active, 1); }
      if (!fits) atomicOr(&d.status[w], 32 /*EGG_ST_DENSE_OVERFLOW*/);
    }
    __syncthreads();
    if (!s_go) continue;

    // A = J J^T (ensembles.cc:664), blocks over shared bodies in ascending body order
    for (int q = tid; q < nc * nc; q += DT) {
      const int c = q / nc, c2 = q % nc;
      double blk[9];
      pair_block(Jb, ci0, ci1, c, c2, false, blk);
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) A[(size_t)(3 * c + k) * R + 3 * c2 + l] = blk[3 * k + l];
    }
    __syncthreads();
    ldlt_block_solve(A, R, nullptr, R, sm, cfg, Fg, [&]() { for (int p = tid; p < R; p += DT) sm.xs[p] = e[sm.perm[p]]; });
    double* y = sm.w;
    for (int p = tid; p < R; p += DT) y[sm.perm[p]] = sm.xs[p];
    __syncthreads();

    // c = -step_scale J^T y ; StepPositions_ExplicitEuler(dt, c) ; mode 1: v += c
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
      if (mode == 1) {                          // UpdateComponentsVelocities(v + c), ensembles.cc:655-656
        for (int t = 0; t < 6; t++) dynw[(12 + t) * n + bd] = dynw[(12 + t) * n + bd] + g[t];
      }
    }
    __syncthreads();
  }
}

int g_sms = 0;
int sm_count() {
  if (!g_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_sms <= 0) g_sms = 148;
  }
  return g_sms;
}

}  // namespace

// Rows the dense path is provisioned for: EGG_DENSE_ROWS, else 3 (nj + 4 n) but at least 336 --
// never more than the batch can produce (3 nrec).  A world with more rows than this is flagged
// EGG_ST_DENSE_OVERFLOW and stepped with lambda = 0.
int egg_dense_row_cap(const EggDev& d) {
  const char* e = getenv("EGG_DENSE_ROWS");
  int cap = e ? atoi(e) : 3 * (d.nj + 4 * d.n);
  if (!e && cap < 336) cap = 336;
  if (cap > 3 * d.nrec) cap = 3 * d.nrec;
  if (cap < 3) cap = 3;
  return (cap / 3) * 3;
}

static DenseCfg dense_cfg(const EggDev& d) {
  DenseCfg c;
  const size_t R = (size_t)egg_dense_row_cap(d);
  c.Rcap = (int)R;
  c.ncap = (int)(R / 3);
  size_t o = 0;
  auto take = [&](size_t cnt) { const size_t at = o; o += (cnt + 1) & ~(size_t)1; return at; };
  c.off_A = take(R * R);
  c.off_Wk = take(2 * R * R);
  c.off_Lm = take(R * R);
  c.off_T1 = take(R * R);
  c.off_Fg = take(R * (R + 1) / 2);
  c.off_Jb = take((size_t)(c.ncap + 1) * 72);
  c.off_vec = take(2 * R);
  c.off_int = take((size_t)c.ncap + (R + 3) / 4 + 2);      // 2 ncap ints + Rcap u16
  c.per_cta = o;
  // shared memory: 6 vectors + the panel scratch, 6 n accumulator, index lists and masks; the rest of the per-CTA
  // budget (two CTAs per SM) is the packed factor
  const size_t fixed = ((6 + PB) * R + 6 * (size_t)d.n) * sizeof(double) + R * (2 * sizeof(unsigned short) + 2) + 64;
  const char* e = getenv("EGG_DENSE_SMEM_KB");
  size_t budget = (size_t)(e ? atoi(e) : 111) * 1024;
  if (budget > 225 * 1024) budget = 225 * 1024;
  size_t f = budget > fixed + 1024 ? (budget - fixed) / sizeof(double) : 128;
  if (f > R * (R + 1) / 2) f = R * (R + 1) / 2;
  c.Fcap = (int)f;
  c.smem = fixed + f * sizeof(double);
  return c;
}

static int dense_grid(const EggDev& d, const void* kernel, size_t smem) {
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, DT, smem);
  if (per_sm < 1) per_sm = 1;
  const int g = sm_count() * per_sm;
  return d.W < g ? d.W : g;
}

// Slabs for the largest grid either kernel can be launched with (two CTAs per SM).
size_t egg_dense_scratch_bytes(const EggDev& d) {
  const DenseCfg c = dense_cfg(d);
  const int g = sm_count() * 2;
  return c.per_cta * sizeof(double) * (size_t)(d.W < g ? d.W : g);
}

int egg_dense_smem_fits(const EggDev& d, size_t limit) { return dense_cfg(d).smem <= limit; }

cudaError_t egg_launch_solve_dense(const EggDev& d, double dt, cudaStream_t s, void* scratch, size_t scratch_bytes) {
  const DenseCfg c = dense_cfg(d);
  cudaError_t e = cudaFuncSetAttribute(egg_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  int grid = dense_grid(d, (const void*)egg_dense_kernel, c.smem);
  const size_t slabs = scratch_bytes / (c.per_cta * sizeof(double));
  if ((size_t)grid > slabs) grid = (int)slabs;
  if (grid < 1) return cudaErrorInvalidValue;
  e = cudaMemsetAsync(d.work_ctr + 1, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  egg_dense_kernel<<<grid, DT, c.smem, s>>>(d, dt, reinterpret_cast<double*>(scratch), c);
  return cudaGetLastError();
}

cudaError_t egg_launch_relax(const EggDev& d, double dt, double step_scale, int max_steps, int mode, void* scratch, size_t scratch_bytes,
                             int* prog, double* err2, int* any_active, cudaStream_t s) {
  const DenseCfg c = dense_cfg(d);
  cudaError_t e = cudaFuncSetAttribute(egg_relax_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem);
  if (e != cudaSuccess) return e;
  int grid = dense_grid(d, (const void*)egg_relax_kernel, c.smem);
  const size_t slabs = scratch_bytes / (c.per_cta * sizeof(double));
  if ((size_t)grid > slabs) grid = (int)slabs;
  if (grid < 1) return cudaErrorInvalidValue;
  e = cudaMemsetAsync(d.work_ctr + 1, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  egg_relax_kernel<<<grid, DT, c.smem, s>>>(d, dt, step_scale, max_steps, mode, reinterpret_cast<double*>(scratch), c, prog, err2, any_active);
  return cudaGetLastError();
}
