// Constraint-row assembly and the projected Gauss-Seidel solve with fused integrate.
//
// Replaces (per world):
//   egg_assemble_kernel  Joint/Contact::ComputeJ + error  joints.cc:3-35, contact.cc:14-117,
//                        Ensemble::ComputeJ / rhs         ensembles.cc:38-87, 156-171, 563-570
//   egg_pgs_kernel       sparse::GaussSeidelIteration     sparse_iterations.cc:148-226, 51-69,
//                        matrix-free block ops            sparse_iterations_utils.cc:12-21,159-243,495-695
//                        + v' = v + dt M^-1 (f + J^T x)   ensembles.cc:535, 572-573
//                        + StepPositions_ODE / WtoQ       ensembles.cc:577-591, utils.cc:82-89
//
// Formulation.  Every constraint (joint or contact) is one 3-row block whose two 3x6 Jacobians
// are [-Rc, Rc [r0]x] and [Rc, -Rc [r1]x] (contact.cc:60-75; a ball joint is the same shape with
// Rc = -I, joints.cc:22-30).  Instead of streaming 2x3x6 Jacobian entries per block the kernels
// keep a compact 240-byte record (Rc, r0, r1, the 3x3 diagonal block D of J M^-1 J^T, rhs) and
// the body-space accumulator a = M^-1 J^T x (6 doubles per body, in shared memory).  One block
// update is  t = Rc (vel1(a) - vel0(a)),  row-by-row projected substitution inside the 3x3
// diagonal block exactly as sparse_iterations_utils.cc:229-236, and an impulse scatter into a.
//
// Schedule.  Gauss-Seidel is sequential in constraint order, but blocks that share no body
// commute exactly.  Blocks are grouped into dependency levels (level(c) = 1 + max level of any
// earlier block sharing a body) and levels are cut into stages of <= 32 blocks; running stage
// after stage with one lane per block is bit-identical to the sequential sweep.
//
// Execution.  One warp (one CTA) owns one world for all of its sweeps.  A stage's records are one
// contiguous piece-major chunk in HBM/L2; lane 0 streams it into a shared-memory ring with a
// single TMA bulk copy (cp.async.bulk + mbarrier complete_tx) NSTAGE-1 stages ahead of its use,
// so the dependent FP64 chain of a stage never waits on L2.  Stages are separated by
// __syncwarp() only.
#include "egg_internal.cuh"
#include "egg_record.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// Assembly: one CTA per world.  Thread 0 computes the dependency levels / stages from the body
// indices staged in shared memory; then one thread per constraint builds its record and writes
// it at its (stage, lane) position in piece-major order.
template <int NT>
__global__ void __launch_bounds__(NT) egg_assemble_kernel(EggDev d, double dt, int stage_cap) {
  extern __shared__ double sm[];
  const int n = d.n, nj = d.nj, w = blockIdx.x, tid = threadIdx.x;
  double* sdyn = sm;                         // [18][n]
  double* sst = sm + EGG_DYN * n;            // [16][n]
  int* si0 = (int*)(sm + (EGG_DYN + EGG_STAT) * n);   // [nrec]
  int* si1 = si0 + d.nrec;                   // [nrec]
  int* slot = si1 + d.nrec;                  // [nrec] record slot of c
  int* lev = slot + d.nrec;                  // [nrec] dependency level of c
  int* lstart = lev + d.nrec;                // [nrec + 1] level counts -> running slot cursor
  int* lstage = lstart + d.nrec + 1;         // [nrec + 1] first stage of each level
  int* lfirst = lstage + d.nrec + 1;         // [nrec + 1] first slot of each level
  int* blast = lfirst + d.nrec + 1;          // [n] last level that touched the body
  const double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  const double* st = d.stat + (size_t)w * EGG_STAT * n;
  for (int i = tid; i < EGG_DYN * n; i += NT) sdyn[i] = dyn[i];
  for (int i = tid; i < EGG_STAT * n; i += NT) sst[i] = st[i];
  const int ncon = d.c_count[w];
  const int nc = nj + ncon;
  const int* c_i0 = d.c_i0 + (size_t)w * d.maxc;
  const int* c_i1 = d.c_i1 + (size_t)w * d.maxc;
  const double* geom = d.c_geom + (size_t)w * 7 * d.maxc;
  const int maxc = d.maxc;
  for (int c = tid; c < nc; c += NT) {
    if (c < nj) { si0[c] = d.j_i0[(size_t)w * nj + c]; si1[c] = d.j_i1[(size_t)w * nj + c]; }
    else { si0[c] = c_i0[c - nj]; si1[c] = c_i1[c - nj]; }
  }
  for (int b = tid; b < n; b += NT) blast[b] = -1;
  for (int c = tid; c <= nc; c += NT) lstart[c] = 0;
  __syncthreads();

  int* gstage = d.level_start + (size_t)w * (d.nrec + 1);
  if (tid == 0) {
    int nl = 0;
    for (int c = 0; c < nc; c++) {           // reference order: joints, then contacts
      const int i0 = si0[c], i1 = si1[c];
      int l = -1;
      if (i0 >= 0) l = max(l, blast[i0]);
      if (i1 >= 0) l = max(l, blast[i1]);
      l += 1;
      if (i0 >= 0) blast[i0] = l;
      if (i1 >= 0) blast[i1] = l;
      lev[c] = l;
      lstart[l]++;
      nl = max(nl, l + 1);
    }
    int ns = 0, run = 0;
    for (int l = 0; l < nl; l++) {
      const int cnt = lstart[l];
      lfirst[l] = run;
      lstart[l] = run;                       // running cursor of the level
      lstage[l] = ns;
      for (int k = 0; k < cnt; k += stage_cap) gstage[ns++] = run + k;
      run += cnt;
    }
    gstage[ns] = nc;
    d.n_levels[w] = ns;
    for (int c = 0; c < nc; c++) slot[c] = lstart[lev[c]]++;   // stable inside a level
  }
  __syncthreads();

  double* recw = d.rec + (size_t)w * d.nrec * EGG_REC;
  for (int c = tid; c < nc; c += NT) {
    const int i0 = si0[c], i1 = si1[c];
    double v[EGG_REC];
    egg_build_record(d, w, c, i0, i1, sdyn, sst, geom, dt, false, v);
    if (d.slot_of) d.slot_of[(size_t)w * d.nrec + c] = slot[c];
    if (d.rec_minv) {   // per-block copy of M^-1 of body i0 then i1 (zeros for the world / ground side)
      double* mv = d.rec_minv + ((size_t)w * d.nrec + slot[c]) * 20;
      for (int k = 0; k < 10; k++) {
        mv[k] = (i0 >= 0) ? sst[k * n + i0] : 0.0;
        mv[10 + k] = (i1 >= 0) ? sst[k * n + i1] : 0.0;
      }
    }
    double2* out = reinterpret_cast<double2*>(recw + (size_t)slot[c] * EGG_REC);
#pragma unroll
    for (int p = 0; p < EGG_PIECES; p++) out[p] = make_double2(v[2 * p], v[2 * p + 1]);
  }
}

// ---------------------------------------------------------------------------------------------
// Shared block math (both solver variants).

struct BlockRec {
  double Rc[9];
  d3 r0, r1;
  double doff[3], ddiag[3], inva[3], rhs[3];
  int i0, i1, orig, kind;
};

__device__ __forceinline__ void unpack_rec(const double2* v, BlockRec& r) {
  // pieces: [0..4] Rc0..8 r0.x | [5..7] r0.y r0.z r1 | ... straight double order, see REC_*
  double f[EGG_REC];
#pragma unroll
  for (int p = 0; p < EGG_PIECES; p++) { f[2 * p] = v[p].x; f[2 * p + 1] = v[p].y; }
#pragma unroll
  for (int k = 0; k < 9; k++) r.Rc[k] = f[REC_RC + k];
  r.r0 = mk3(f[REC_R0], f[REC_R0 + 1], f[REC_R0 + 2]);
  r.r1 = mk3(f[REC_R1], f[REC_R1 + 1], f[REC_R1 + 2]);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.doff[k] = f[REC_DOFF + k]; r.ddiag[k] = f[REC_DDIAG + k]; r.inva[k] = f[REC_INVA + k]; r.rhs[k] = f[REC_RHS + k];
  }
  r.i0 = __double2loint(f[REC_IDX]); r.i1 = __double2hiint(f[REC_IDX]);
  r.orig = __double2loint(f[REC_META]); r.kind = __double2hiint(f[REC_META]);
}

// Per-body shared-memory struct: [0..5] a = M^-1 J^T x (lin, ang), [6] 1/m, [7..15] I^-1 (row-major).
// The stride is odd (17 or 7 doubles) so that lanes touching different bodies hit different banks,
// and every field is at a compile-time offset from one base address (no per-field index math).
template <bool MS> struct BodyStride { static constexpr int value = MS ? 17 : 7; };

// t = J a for the block: Rc (vel1 - vel0), vel_b = a_lin + a_ang x r_b.
template <int BS, int OFF = 0>
__device__ __forceinline__ d3 block_Ja(const BlockRec& r, const double* sb) {
  d3 u = mk3(0, 0, 0);
  if (r.i1 >= 0) {
    const double* q = sb + r.i1 * BS + OFF;
    d3 al = mk3(q[0], q[1], q[2]);
    d3 aa = mk3(q[3], q[4], q[5]);
    u = al + cross3(aa, r.r1);
  }
  if (r.i0 >= 0) {
    const double* q = sb + r.i0 * BS + OFF;
    d3 al = mk3(q[0], q[1], q[2]);
    d3 aa = mk3(q[3], q[4], q[5]);
    u = u - (al + cross3(aa, r.r0));
  }
  return mmulv(r.Rc, u);
}

// a += M^-1 J^T delta for the block.  MS: M^-1 from the body struct, else from the read-only
// global array st ([10][n] = 1/m, Iinv).
template <bool MS, int BS = BodyStride<MS>::value, int MO = 6>
__device__ __forceinline__ void block_scatter(const BlockRec& r, d3 delta, double* sb, const double* st, int n) {
  d3 imp = mtmulv(r.Rc, delta);
  if (r.i1 >= 0) {
    const int b = r.i1;
    double* q = sb + b * BS;
    double Ii[9];
    const double mi = MS ? q[MO] : __ldg(st + b);
#pragma unroll
    for (int k = 0; k < 9; k++) Ii[k] = MS ? q[MO + 1 + k] : __ldg(st + (1 + k) * n + b);
    d3 da = mmulv(Ii, cross3(r.r1, imp));
    q[0] += mi * imp.x; q[1] += mi * imp.y; q[2] += mi * imp.z;
    q[3] += da.x; q[4] += da.y; q[5] += da.z;
  }
  if (r.i0 >= 0) {
    const int b = r.i0;
    double* q = sb + b * BS;
    double Ii[9];
    const double mi = MS ? q[MO] : __ldg(st + b);
#pragma unroll
    for (int k = 0; k < 9; k++) Ii[k] = MS ? q[MO + 1 + k] : __ldg(st + (1 + k) * n + b);
    d3 da = mmulv(Ii, cross3(r.r0, imp));
    q[0] -= mi * imp.x; q[1] -= mi * imp.y; q[2] -= mi * imp.z;
    q[3] -= da.x; q[4] -= da.y; q[5] -= da.z;
  }
}

// Zero the accumulators of a world and (MS) load its M^-1 into the body structs.
template <bool MS, int BS = BodyStride<MS>::value, int MO = 6>
__device__ __forceinline__ void init_bodies(double* sb, const double* st, int n, int lane, int nl) {
  for (int i = lane; i < n * BS; i += nl) {
    const int b = i / BS, f = i - b * BS;
    sb[i] = (MS && f >= MO && f < MO + 10) ? st[(f - MO) * n + b] : 0.0;
  }
}

__device__ __forceinline__ double project(double x, int kind, int row) {   // sparse_iterations_utils.cc:12-21
  if (kind == KIND_CONTACT) {
    if (row < 2) { if (x < -1.0) return -1.0; else if (x > 1.0) return 1.0; }
    else { if (x < 0.0) return 0.0; }
  }
  return x;
}

// One Gauss-Seidel block update: row-by-row substitution inside the 3x3 diagonal block
// (sparse_iterations_utils.cc:229-236).  x is updated in place; returns delta.
__device__ __forceinline__ d3 gs_rows(const BlockRec& r, d3 t, double& x0, double& x1, double& x2) {
  double n0 = project((r.rhs[0] - t.x + r.ddiag[0] * x0) * r.inva[0], r.kind, 0);
  double d0 = n0 - x0;
  double n1 = project((r.rhs[1] - (t.y + r.doff[0] * d0) + r.ddiag[1] * x1) * r.inva[1], r.kind, 1);
  double d1 = n1 - x1;
  double n2 = project((r.rhs[2] - (t.z + r.doff[1] * d0 + r.doff[2] * d1) + r.ddiag[2] * x2) * r.inva[2], r.kind, 2);
  double d2 = n2 - x2;
  x0 = n0; x1 = n1; x2 = n2;
  return mk3(d0, d1, d2);
}

// GetResidualError partial sums for one block (sparse_iterations.cc:51-69); the reference
// classifies with each block's OWN bounds here (ConstructMixedConstraints).
__device__ __forceinline__ void residual_rows(const BlockRec& r, d3 t, double x0, double x1, double x2, bool eq, double cfm,
                                              double& se, double& s1, double& s2, double& s3) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const double x = (k == 0) ? x0 : (k == 1 ? x1 : x2);
    const double wv = get3(t, k) + cfm * x - r.rhs[k];
    if (eq) { se += wv * wv; continue; }
    const double lo = (k < 2) ? -1.0 : 0.0;
    const bool has_hi = (k < 2);
    if (x == lo && wv < 0) s1 += wv * wv;
    if (has_hi && x == 1.0 && wv > 0) s2 += wv * wv;
    if (x > lo && (!has_hi || x < 1.0)) s3 += wv * wv;
  }
}

// v' = v + dt (M^-1 f + a); p += dt (v+v')/2; R <- WtoQ((w+w')/2, dt) R  for body b of a world.
template <int BS, int OFF = 0>
__device__ __forceinline__ bool integrate_body(double* dyn, const double* st, const double* sb, int n, int b, double dt) {
  const double mi = __ldg(st + b);
  double Ii[9];
#pragma unroll
  for (int k = 0; k < 9; k++) Ii[k] = __ldg(st + (1 + k) * n + b);
  d3 fl = mk3(st[10 * n + b], st[11 * n + b], st[12 * n + b]);
  d3 ft = mk3(st[13 * n + b], st[14 * n + b], st[15 * n + b]);
  d3 v = mk3(dyn[12 * n + b], dyn[13 * n + b], dyn[14 * n + b]);
  d3 wv = mk3(dyn[15 * n + b], dyn[16 * n + b], dyn[17 * n + b]);
  d3 al = mk3(sb[b * BS + OFF], sb[b * BS + OFF + 1], sb[b * BS + OFF + 2]);
  d3 aa = mk3(sb[b * BS + OFF + 3], sb[b * BS + OFF + 4], sb[b * BS + OFF + 5]);
  d3 vn = v + dt * (fl * mi + al);
  d3 wn = wv + dt * (mmulv(Ii, ft) + aa);
  d3 vmid = (v + vn) / 2.0, wmid = (wv + wn) / 2.0;
  d3 p = mk3(dyn[b], dyn[n + b], dyn[2 * n + b]) + dt * vmid;
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
  double R[9], Rn[9];
#pragma unroll
  for (int k = 0; k < 9; k++) R[k] = dyn[(3 + k) * n + b];
  mmulm(Q, R, Rn);
  dyn[b] = p.x; dyn[n + b] = p.y; dyn[2 * n + b] = p.z;
#pragma unroll
  for (int k = 0; k < 9; k++) dyn[(3 + k) * n + b] = Rn[k];
  dyn[12 * n + b] = vn.x; dyn[13 * n + b] = vn.y; dyn[14 * n + b] = vn.z;
  dyn[15 * n + b] = wn.x; dyn[16 * n + b] = wn.y; dyn[17 * n + b] = wn.z;
  double chk = p.x + p.y + p.z + vn.x + vn.y + vn.z + wn.x + wn.y + wn.z;
  return !(fabs(chk) < 1e300);
}

// Multipliers / row state of slot `s` in reference row order.
__device__ __forceinline__ void write_solution(const EggDev& d, int w, const double* recs, const double* lam, int s) {
  const int nj = d.nj;
  const double meta = recs[(size_t)s * EGG_REC + REC_META];
  const int orig = __double2loint(meta);
  const bool eq = orig < nj;
  double* lo_out = d.lam_out + (size_t)w * 3 * d.nrec;
  int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double x = __ldcg(lam + 3 * (size_t)s + k);
    lo_out[3 * orig + k] = x;
    int state = 0;
    if (eq) state = 3;
    else if (x == ((k < 2) ? -1.0 : 0.0)) state = 1;
    else if (k < 2 && x == 1.0) state = 2;
    rs_out[3 * orig + k] = state;
  }
}

// ---------------------------------------------------------------------------------------------
// Variant "mw" (default): a warp steps G = 32/LPW worlds in lock-step, LPW lanes per world.
//
// ncu on the one-world-per-warp kernels (profiles/r1a_*) showed the solve is bound by the
// dependent-issue latency of a ~350-instruction stage with ~4 of 32 lanes active (FP64 pipe 11 %,
// issue slots 30 %, 1.5 warps per scheduler because shared memory capped the CTAs per SM).  This
// variant packs several worlds into one warp so an instruction serves 4-8x more blocks and
// relies on many resident warps (not on a prefetch ring) to cover the L2 latency of the record
// loads; the accumulator a (and optionally M^-1) lives in shared memory.
template <int LPW, bool MINV_SMEM>
__global__ void __launch_bounds__(32) egg_pgs_mw_kernel(EggDev d, double dt, int tabcap, int flags) {
  constexpr int G = 32 / LPW;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x, sub = lane / LPW, sl = lane % LPW;
  constexpr int BS = BodyStride<MINV_SMEM>::value;   // doubles per body kept in shared memory
  double* sa = reinterpret_cast<double*>(smraw) + (size_t)sub * BS * n;
  unsigned char* tab = smraw + (size_t)G * BS * n * 8 + (size_t)sub * tabcap;   // stage sizes (<= LPW each)
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max, nj = d.nj;

  for (int wbase = blockIdx.x * G; wbase < d.W; wbase += gridDim.x * G) {
    const int w = wbase + sub;
    const bool valid = w < d.W;
    const int wc = valid ? w : d.W - 1;
    const double* st = d.stat + (size_t)wc * EGG_STAT * n;
    const int nc = valid ? nj + d.c_count[wc] : 0;
    const int ns = valid ? d.n_levels[wc] : 0;
    const int* gls = d.level_start + (size_t)wc * (d.nrec + 1);
    const double* recs = d.rec + (size_t)wc * d.nrec * EGG_REC;
    double* lam = d.lam + (size_t)wc * d.nrec * 3;
    init_bodies<MINV_SMEM>(sa, st, n, sl, LPW);
    for (int i = sl; i < ns && i < tabcap; i += LPW) tab[i] = (unsigned char)(gls[i + 1] - gls[i]);
    __syncwarp();
    auto stage_cnt = [&](int s) -> int { return (s < tabcap) ? (int)tab[s] : (__ldg(gls + s + 1) - __ldg(gls + s)); };

    int ns_max = ns, nc_max = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ns_max = max(ns_max, __shfl_xor_sync(0xffffffffu, ns_max, o));
      nc_max = max(nc_max, __shfl_xor_sync(0xffffffffu, nc_max, o));
    }
    const int nchunk = (nc_max + LPW - 1) / LPW;

    bool active = nc > 0;
    double err = 0.0;
    int it = 0;
    int pass_kind = 0;   // 0: x0 = rhs scatter (sparse_iterations.cc:202), 1: GS update, 2: residual
    if (__any_sync(0xffffffffu, active)) {
      while (true) {
        const int nsteps = (pass_kind == 2) ? nchunk : ns_max;
        double se = 0, s1 = 0, s2 = 0, s3 = 0;
        int s0 = 0;                      // first slot of the current stage (update passes)
        for (int t = 0; t < nsteps; t++) {
          int slot = -1;
          if (pass_kind == 2) {
            const int f = t * LPW + sl;
            if (active && f < nc) slot = f;
          } else if (active && t < ns) {
            const int cnt = stage_cnt(t);
            if (sl < cnt) slot = s0 + sl;
            s0 += cnt;
          }
          // L1 prefetch of the next step's record (no registers held across the step)
          if (flags & 1) {
            int nslot = -1;
            if (pass_kind == 2) {
              const int f = (t + 1) * LPW + sl;
              if (active && f < nc) nslot = f;
            } else if (active && t + 1 < ns) {
              if (sl < stage_cnt(t + 1)) nslot = s0 + sl;
            }
            if (nslot >= 0) {
              const char* np = reinterpret_cast<const char*>(recs + (size_t)nslot * EGG_REC);
              asm volatile("prefetch.global.L1 [%0];" ::"l"(np));
              asm volatile("prefetch.global.L1 [%0];" ::"l"(np + 128));
              asm volatile("prefetch.global.L2 [%0];" ::"l"(lam + 3 * (size_t)nslot));
            }
          }
          if (slot >= 0) {
            const double2* rp = reinterpret_cast<const double2*>(recs + (size_t)slot * EGG_REC);
            double2 v[EGG_PIECES];
#pragma unroll
            for (int p = 0; p < EGG_PIECES; p++) v[p] = __ldg(rp + p);
            double* lp = lam + 3 * (size_t)slot;
            BlockRec r;
            unpack_rec(v, r);
            if (pass_kind == 0) {
              lp[0] = r.rhs[0]; lp[1] = r.rhs[1]; lp[2] = r.rhs[2];
              block_scatter<MINV_SMEM>(r, mk3(r.rhs[0], r.rhs[1], r.rhs[2]), sa, st, n);
            } else {
              double c0, c1, c2;
              if (flags & 2) { c0 = __ldcg(lp); c1 = __ldcg(lp + 1); c2 = __ldcg(lp + 2); }
              else { c0 = lp[0]; c1 = lp[1]; c2 = lp[2]; }
              d3 t3 = block_Ja<BS>(r, sa);
              if (pass_kind == 1) {
                d3 dl = gs_rows(r, t3, c0, c1, c2);
                lp[0] = c0; lp[1] = c1; lp[2] = c2;
                block_scatter<MINV_SMEM>(r, dl, sa, st, n);
              } else {
                residual_rows(r, t3, c0, c1, c2, r.orig < nj, cfm, se, s1, s2, s3);
              }
            }
          }
          if (pass_kind != 2) __syncwarp();
        }
        if (pass_kind == 2) {
#pragma unroll
          for (int o = LPW / 2; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
          }
          if (active) {
            err = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));
            if (!(err > tol && it < k_max)) active = false;
          }
          if (!__any_sync(0xffffffffu, active)) break;
          pass_kind = 1;
        } else {
          if (pass_kind == 1 && active) ++it;
          pass_kind = 2;
          __syncwarp();
        }
      }
    }
    __syncwarp();

    if (valid) {
      for (int s = sl; s < nc; s += LPW) write_solution(d, w, recs, lam, s);
      if (sl == 0) {
        int* stt = d.stats + (size_t)w * 8;
        stt[4] = it;
        stt[5] = 0;
        stt[6] = (cfm != 0.0);
        stt[7] = ns;
        d.resid[w] = err;
      }
      double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
      bool bad = false;
      for (int b = sl; b < n; b += LPW) bad |= integrate_body<BS>(dyn, st, sa, n, b, dt);
      if (bad) atomicOr(&d.status[w], 16 /*EGG_ST_NONFINITE*/);
    }
    __syncwarp();
  }
}

// Variant "mwpf": as "mw", plus a register double buffer: the next step's record and multipliers
// are loaded while the current step computes (costs ~100 registers => 9 warps per SM), M^-1 comes
// through the read-only L1 path.  Fastest measured variant for wide worlds (n = 64).
template <int LPW, bool MINV_SMEM>
__global__ void __launch_bounds__(32) egg_pgs_mwpf_kernel(EggDev d, double dt, int tabcap) {
  constexpr int G = 32 / LPW;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x, sub = lane / LPW, sl = lane % LPW;
  constexpr int BS = BodyStride<MINV_SMEM>::value;
  double* sa = reinterpret_cast<double*>(smraw) + (size_t)sub * BS * n;
  unsigned char* tab = smraw + (size_t)G * BS * n * 8 + (size_t)sub * tabcap;   // stage sizes (<= LPW each)
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max, nj = d.nj;

  for (int wbase = blockIdx.x * G; wbase < d.W; wbase += gridDim.x * G) {
    const int w = wbase + sub;
    const bool valid = w < d.W;
    const int wc = valid ? w : d.W - 1;
    const double* st = d.stat + (size_t)wc * EGG_STAT * n;
    const int nc = valid ? nj + d.c_count[wc] : 0;
    const int ns = valid ? d.n_levels[wc] : 0;
    const int* gls = d.level_start + (size_t)wc * (d.nrec + 1);
    const double* recs = d.rec + (size_t)wc * d.nrec * EGG_REC;
    double* lam = d.lam + (size_t)wc * d.nrec * 3;
    init_bodies<MINV_SMEM>(sa, st, n, sl, LPW);
    for (int i = sl; i < ns && i < tabcap; i += LPW) tab[i] = (unsigned char)(gls[i + 1] - gls[i]);
    __syncwarp();
    auto stage_cnt = [&](int s) -> int { return (s < tabcap) ? (int)tab[s] : (__ldg(gls + s + 1) - __ldg(gls + s)); };

    int ns_max = ns, nc_max = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ns_max = max(ns_max, __shfl_xor_sync(0xffffffffu, ns_max, o));
      nc_max = max(nc_max, __shfl_xor_sync(0xffffffffu, nc_max, o));
    }
    const int nchunk = (nc_max + LPW - 1) / LPW;

    bool active = nc > 0;
    double err = 0.0;
    int it = 0;
    int pass_kind = 0;   // 0: x0 = rhs scatter (sparse_iterations.cc:202), 1: GS update, 2: residual
    // record (+ multipliers) of the current step and of the next one (register double buffer)
    double2 cur[EGG_PIECES], nxt[EGG_PIECES];
    double c0 = 0, c1 = 0, c2 = 0, p0 = 0, p1 = 0, p2 = 0;
    int cur_slot = -1, nxt_slot = -1;
    // stage cursor of the update passes
    int s0_next = 0;     // first slot of the stage that will be fetched next
    auto fetch = [&](int slot, bool with_lam) {
      nxt_slot = slot;
      if (slot >= 0) {
        const double2* rp = reinterpret_cast<const double2*>(recs + (size_t)slot * EGG_REC);
#pragma unroll
        for (int p = 0; p < EGG_PIECES; p++) nxt[p] = __ldg(rp + p);
        if (with_lam) {
          p0 = __ldcg(lam + 3 * (size_t)slot); p1 = __ldcg(lam + 3 * (size_t)slot + 1); p2 = __ldcg(lam + 3 * (size_t)slot + 2);
        }
      }
    };
    auto stage_slot = [&](int s) -> int {     // slot of this lane in stage s (call with s in order)
      if (!active || s >= ns) return -1;
      const int cnt = stage_cnt(s);
      const int mine = (sl < cnt) ? s0_next + sl : -1;
      s0_next += cnt;
      return mine;
    };
    auto chunk_slot = [&](int k) -> int {
      const int f = k * LPW + sl;
      return (active && f < nc) ? f : -1;
    };

    if (__any_sync(0xffffffffu, active)) {
      s0_next = 0;
      fetch(stage_slot(0), false);
      while (true) {
        const int nsteps = (pass_kind == 2) ? nchunk : ns_max;
        double se = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int t = 0; t < nsteps; t++) {
          // rotate the double buffer, then prefetch step t+1 (or step 0 of the next pass)
#pragma unroll
          for (int p = 0; p < EGG_PIECES; p++) cur[p] = nxt[p];
          cur_slot = nxt_slot; c0 = p0; c1 = p1; c2 = p2;
          const bool last = (t + 1 == nsteps);
          if (!last) {
            fetch(pass_kind == 2 ? chunk_slot(t + 1) : stage_slot(t + 1), pass_kind != 0);
          } else if (pass_kind == 2) {
            s0_next = 0;
            fetch(stage_slot(0), true);       // residual never writes lam: safe to prefetch
          } else {
            fetch(chunk_slot(0), false);      // lam of chunk 0 may still be in flight: read it later
          }
          if (cur_slot >= 0) {
            BlockRec r;
            unpack_rec(cur, r);
            double* lp = lam + 3 * (size_t)cur_slot;
            if (pass_kind == 0) {
              lp[0] = r.rhs[0]; lp[1] = r.rhs[1]; lp[2] = r.rhs[2];
              block_scatter<MINV_SMEM>(r, mk3(r.rhs[0], r.rhs[1], r.rhs[2]), sa, st, n);
            } else if (pass_kind == 1) {
              d3 t3 = block_Ja<BS>(r, sa);
              d3 dl = gs_rows(r, t3, c0, c1, c2);
              lp[0] = c0; lp[1] = c1; lp[2] = c2;
              block_scatter<MINV_SMEM>(r, dl, sa, st, n);
            } else {
              d3 t3 = block_Ja<BS>(r, sa);
              residual_rows(r, t3, c0, c1, c2, r.orig < nj, cfm, se, s1, s2, s3);
            }
          }
          __syncwarp();
          if (last && pass_kind != 2 && nxt_slot >= 0) {
            // first residual chunk: its multipliers were written during the pass that just ended
            p0 = __ldcg(lam + 3 * (size_t)nxt_slot); p1 = __ldcg(lam + 3 * (size_t)nxt_slot + 1); p2 = __ldcg(lam + 3 * (size_t)nxt_slot + 2);
          }
        }
        if (pass_kind == 2) {
#pragma unroll
          for (int o = LPW / 2; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
          }
          if (active) {
            err = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));
            if (!(err > tol && it < k_max)) active = false;
          }
          if (!__any_sync(0xffffffffu, active)) break;
          if (!active) nxt_slot = -1;          // this world dropped out: discard its prefetch
          pass_kind = 1;
        } else {
          if (pass_kind == 1 && active) ++it;
          pass_kind = 2;
        }
      }
    }
    __syncwarp();

    if (valid) {
      for (int s = sl; s < nc; s += LPW) write_solution(d, w, recs, lam, s);
      if (sl == 0) {
        int* stt = d.stats + (size_t)w * 8;
        stt[4] = it;
        stt[5] = 0;
        stt[6] = (cfm != 0.0);
        stt[7] = ns;
        d.resid[w] = err;
      }
      double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
      bool bad = false;
      for (int b = sl; b < n; b += LPW) bad |= integrate_body<BS>(dyn, st, sa, n, b, dt);
      if (bad) atomicOr(&d.status[w], 16 /*EGG_ST_NONFINITE*/);
    }
    __syncwarp();
  }
}

// Variant "fused" (default for wide worlds): as "mwpf", but the termination residual of sweep k is
// evaluated INSIDE sweep k+1 (from a frozen copy a_prev of the accumulator and the multipliers of
// sweep k), so the records are streamed once per sweep instead of twice and a sweep has one
// pass of stages instead of two.  Sweep k+1 is therefore speculative: if the residual of sweep k
// turns out to be <= tol the kernel returns x_k / a_prev and discards x_{k+1} (the multipliers
// ping-pong between two buffers).  Results, sweep counts and clamp states are identical to the
// unfused order of operations (sparse_iterations.cc:204-222).
// Body struct: [0..5] a, [6..11] a_prev, [12..21] 1/m, I^-1 (MS), odd stride.
template <bool MS> struct FusedStride { static constexpr int value = MS ? 23 : 13; };

template <int LPW, bool MINV_SMEM>
__global__ void __launch_bounds__(32) egg_pgs_fused_kernel(EggDev d, double dt, int tabcap) {
  constexpr int G = 32 / LPW;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x, sub = lane / LPW, sl = lane % LPW;
  constexpr int BS = FusedStride<MINV_SMEM>::value;
  constexpr int MO = 12;
  double* sa = reinterpret_cast<double*>(smraw) + (size_t)sub * BS * n;
  unsigned char* tab = smraw + (size_t)G * BS * n * 8 + (size_t)sub * tabcap;
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max, nj = d.nj;

  for (int wbase = blockIdx.x * G; wbase < d.W; wbase += gridDim.x * G) {
    const int w = wbase + sub;
    const bool valid = w < d.W;
    const int wc = valid ? w : d.W - 1;
    const double* st = d.stat + (size_t)wc * EGG_STAT * n;
    const int nc = valid ? nj + d.c_count[wc] : 0;
    const int ns = valid ? d.n_levels[wc] : 0;
    const int* gls = d.level_start + (size_t)wc * (d.nrec + 1);
    const double* recs = d.rec + (size_t)wc * d.nrec * EGG_REC;
    double* lamb[2] = {d.lam + (size_t)wc * d.nrec * 3, d.lam2 + (size_t)wc * d.nrec * 3};
    init_bodies<MINV_SMEM, BS, MO>(sa, st, n, sl, LPW);
    for (int i = sl; i < ns && i < tabcap; i += LPW) tab[i] = (unsigned char)(gls[i + 1] - gls[i]);
    __syncwarp();
    auto stage_cnt = [&](int s) -> int { return (s < tabcap) ? (int)tab[s] : (__ldg(gls + s + 1) - __ldg(gls + s)); };

    int ns_max = ns, nc_max = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ns_max = max(ns_max, __shfl_xor_sync(0xffffffffu, ns_max, o));
      nc_max = max(nc_max, __shfl_xor_sync(0xffffffffu, nc_max, o));
    }
    const int nchunk = (nc_max + LPW - 1) / LPW;

    bool active = nc > 0, use_prev = false;
    double err = 0.0;
    int it = 0;          // accepted sweeps of this world
    int k = 0;           // accepted sweeps of the still-active worlds (warp-uniform)
    int rd = 0;          // buffer holding x_k
    int pass_kind = 0;   // 0: x0 = rhs scatter, 3: fused residual(k) + update(k -> k+1), 2: residual only
    double2 cur[EGG_PIECES], nxt[EGG_PIECES];
    double c0 = 0, c1 = 0, c2 = 0, p0 = 0, p1 = 0, p2 = 0;
    int cur_slot = -1, nxt_slot = -1;
    int s0_next = 0;
    auto fetch = [&](int slot, bool with_lam) {
      nxt_slot = slot;
      if (slot >= 0) {
        const double2* rp = reinterpret_cast<const double2*>(recs + (size_t)slot * EGG_REC);
#pragma unroll
        for (int p = 0; p < EGG_PIECES; p++) nxt[p] = __ldg(rp + p);
        if (with_lam) {
          const double* lq = lamb[rd] + 3 * (size_t)slot;
          p0 = __ldcg(lq); p1 = __ldcg(lq + 1); p2 = __ldcg(lq + 2);
        }
      }
    };
    auto stage_slot = [&](int s) -> int {
      if (!active || s >= ns) return -1;
      const int cnt = stage_cnt(s);
      const int mine = (sl < cnt) ? s0_next + sl : -1;
      s0_next += cnt;
      return mine;
    };
    auto chunk_slot = [&](int c) -> int {
      const int f = c * LPW + sl;
      return (active && f < nc) ? f : -1;
    };
    auto load_lam_for_next = [&]() {   // multipliers of the prefetched first step of the next pass
      if (nxt_slot >= 0) {
        const double* lq = lamb[rd] + 3 * (size_t)nxt_slot;
        p0 = __ldcg(lq); p1 = __ldcg(lq + 1); p2 = __ldcg(lq + 2);
      }
    };

    if (__any_sync(0xffffffffu, active)) {
      s0_next = 0;
      fetch(stage_slot(0), false);
      while (true) {
        const int nsteps = (pass_kind == 2) ? nchunk : ns_max;
        double se = 0, s1 = 0, s2 = 0, s3 = 0;
        if (pass_kind == 3) {   // freeze a_k: the residual of sweep k is taken against it
          for (int b = sl; active && b < n; b += LPW) {
            double* q = sa + b * BS;
#pragma unroll
            for (int f = 0; f < 6; f++) q[6 + f] = q[f];
          }
          __syncwarp();
        }
        for (int t = 0; t < nsteps; t++) {
#pragma unroll
          for (int p = 0; p < EGG_PIECES; p++) cur[p] = nxt[p];
          cur_slot = nxt_slot; c0 = p0; c1 = p1; c2 = p2;
          const bool last = (t + 1 == nsteps);
          // within a pass the read buffer is never written, so next-step multipliers can be
          // prefetched; across a pass boundary only the record is (its buffer is chosen later)
          if (!last) fetch(pass_kind == 2 ? chunk_slot(t + 1) : stage_slot(t + 1), pass_kind != 0);
          if (cur_slot >= 0) {
            BlockRec r;
            unpack_rec(cur, r);
            if (pass_kind == 0) {
              double* lp = lamb[0] + 3 * (size_t)cur_slot;
              lp[0] = r.rhs[0]; lp[1] = r.rhs[1]; lp[2] = r.rhs[2];
              block_scatter<MINV_SMEM, BS, MO>(r, mk3(r.rhs[0], r.rhs[1], r.rhs[2]), sa, st, n);
            } else if (pass_kind == 3) {
              d3 tp = block_Ja<BS, 6>(r, sa);
              residual_rows(r, tp, c0, c1, c2, r.orig < nj, cfm, se, s1, s2, s3);
              d3 t3 = block_Ja<BS, 0>(r, sa);
              d3 dl = gs_rows(r, t3, c0, c1, c2);
              double* lp = lamb[rd ^ 1] + 3 * (size_t)cur_slot;
              lp[0] = c0; lp[1] = c1; lp[2] = c2;
              block_scatter<MINV_SMEM, BS, MO>(r, dl, sa, st, n);
            } else {
              d3 t3 = block_Ja<BS, 0>(r, sa);
              residual_rows(r, t3, c0, c1, c2, r.orig < nj, cfm, se, s1, s2, s3);
            }
          }
          __syncwarp();
        }
        if (pass_kind == 0) {
          pass_kind = (k_max > 0) ? 3 : 2;
        } else {
#pragma unroll
          for (int o = LPW / 2; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
          }
          if (active) {
            err = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));   // residual of x_k
            if (pass_kind == 2) {
              active = false;                                    // k == k_max: x_k is final, a is final
            } else if (err > tol) {
              ++it; rd ^= 1;                                     // accept x_{k+1}
            } else {
              active = false; use_prev = true;                   // x_k had converged: drop x_{k+1}
            }
          }
          if (!__any_sync(0xffffffffu, active)) break;
          ++k;
          pass_kind = (k >= k_max) ? 2 : 3;
        }
        // first step of the next pass
        s0_next = 0;
        fetch(pass_kind == 2 ? chunk_slot(0) : stage_slot(0), false);
        load_lam_for_next();
      }
    }
    __syncwarp();

    if (valid) {
      const double* lamf = lamb[rd];
      for (int s = sl; s < nc; s += LPW) write_solution(d, w, recs, lamf, s);
      if (sl == 0) {
        int* stt = d.stats + (size_t)w * 8;
        stt[4] = it;
        stt[5] = 0;
        stt[6] = (cfm != 0.0);
        stt[7] = ns;
        d.resid[w] = err;
      }
      double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
      bool bad = false;
      for (int b = sl; b < n; b += LPW)
        bad |= use_prev ? integrate_body<BS, 6>(dyn, st, sa, n, b, dt) : integrate_body<BS, 0>(dyn, st, sa, n, b, dt);
      if (bad) atomicOr(&d.status[w], 16 /*EGG_ST_NONFINITE*/);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Variant "tma": one world per warp, a stage's records staged through a shared-memory ring by
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Kept selectable (EGG_PGS_VARIANT=tma)
// as the measured alternative; see profiles/r1a_pgs_tma_ring_summary.txt.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

constexpr int TMA_NSTAGE = 3;                       // ring depth: NSTAGE-1 stages of lookahead
constexpr int TMA_CAP = 8;                          // blocks per stage (240-byte stride is conflict-free for <= 8 lanes)
constexpr int TMA_SLOT_BYTES = TMA_CAP * EGG_REC * 8;
constexpr int TMA_LS_CAP = 1023;

// Shared memory per CTA (one warp): a[6n] | minv[10n] | ring | mbar | stage starts
__global__ void __launch_bounds__(32) egg_pgs_tma_kernel(EggDev d, double dt) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x;
  double* sa = reinterpret_cast<double*>(smraw);
  constexpr int BS = BodyStride<true>::value;
  unsigned char* ring = smraw + (((size_t)BS * n * 8 + 127) & ~(size_t)127);
  unsigned long long* mbar = reinterpret_cast<unsigned long long*>(ring + (size_t)TMA_NSTAGE * TMA_SLOT_BYTES);
  int* sls = reinterpret_cast<int*>(mbar + TMA_NSTAGE);
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max;
  const int nj = d.nj;

  if (lane == 0) {
    for (int k = 0; k < TMA_NSTAGE; k++) mbar_init(&mbar[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  unsigned issued = 0, consumed = 0;   // warp-uniform monotonic stage counters (slot / parity)

  for (int w = blockIdx.x; w < d.W; w += gridDim.x) {
    const double* st = d.stat + (size_t)w * EGG_STAT * n;
    init_bodies<true>(sa, st, n, lane, 32);
    const int nc = nj + d.c_count[w];
    const int ns = d.n_levels[w];
    const int* gls = d.level_start + (size_t)w * (d.nrec + 1);
    for (int i = lane; i <= ns && i <= TMA_LS_CAP; i += 32) sls[i] = gls[i];
    const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
    double* lam = d.lam + (size_t)w * d.nrec * 3;
    __syncwarp();
    auto stage_start = [&](int s) -> int { return (s <= TMA_LS_CAP) ? sls[s] : __ldg(gls + s); };

    unsigned q_issue = 0;
    auto issue = [&]() {                        // counters are warp-uniform; lane 0 talks to the TMA
      if (lane == 0) {
        const int s = (int)(q_issue % (unsigned)ns);
        const int s0 = stage_start(s), cnt = stage_start(s + 1) - s0;
        const unsigned slot = issued % TMA_NSTAGE;
        const unsigned bytes = (unsigned)cnt * EGG_REC * 8;
        mbar_expect_tx(&mbar[slot], bytes);
        tma_bulk_g2s(ring + (size_t)slot * TMA_SLOT_BYTES, recs + (size_t)s0 * EGG_REC, bytes, &mbar[slot]);
      }
      issued++;
      q_issue++;
    };
    double err = 0.0;
    int it = 0;
    if (nc > 0) {
      for (int k = 0; k < TMA_NSTAGE - 1; k++) issue();
      __syncwarp();
      int pass_kind = 0;
      bool done = false;
      double c0 = 0, c1 = 0, c2 = 0;
      while (!done) {
        double se = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int s = 0; s < ns; s++) {
          issue();   // refills the slot consumed by the previous stage (all lanes passed its __syncwarp)
          const unsigned slot = consumed % TMA_NSTAGE, parity = (consumed / TMA_NSTAGE) & 1u;
          while (!mbar_try_wait(&mbar[slot], parity)) {}
          consumed++;
          const int s0 = stage_start(s), cnt = stage_start(s + 1) - s0;
          const int sn = (s + 1 == ns) ? 0 : s + 1;
          const int s0n = stage_start(sn), cntn = stage_start(sn + 1) - s0n;
          double p0 = 0, p1 = 0, p2 = 0;
          if (ns > 1 && lane < cntn) {
            const double* pq = lam + 3 * (size_t)(s0n + lane);
            p0 = __ldcg(pq); p1 = __ldcg(pq + 1); p2 = __ldcg(pq + 2);
          }
          if (lane < cnt) {
            BlockRec r;
            unpack_rec(reinterpret_cast<const double2*>(ring + (size_t)slot * TMA_SLOT_BYTES) + lane * EGG_PIECES, r);
            double* lp = lam + 3 * (size_t)(s0 + lane);
            if (ns == 1 && pass_kind != 0) { c0 = __ldcg(lp); c1 = __ldcg(lp + 1); c2 = __ldcg(lp + 2); }
            if (pass_kind == 0) {
              lp[0] = r.rhs[0]; lp[1] = r.rhs[1]; lp[2] = r.rhs[2];
              block_scatter<true>(r, mk3(r.rhs[0], r.rhs[1], r.rhs[2]), sa, st, n);
            } else if (pass_kind == 1) {
              d3 t = block_Ja<BS>(r, sa);
              d3 dl = gs_rows(r, t, c0, c1, c2);
              lp[0] = c0; lp[1] = c1; lp[2] = c2;
              block_scatter<true>(r, dl, sa, st, n);
            } else {
              d3 t = block_Ja<BS>(r, sa);
              residual_rows(r, t, c0, c1, c2, r.orig < nj, cfm, se, s1, s2, s3);
            }
          }
          c0 = p0; c1 = p1; c2 = p2;
          __syncwarp();
        }
        if (pass_kind == 2) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            s3 += __shfl_xor_sync(0xffffffffu, s3, o);
          }
          err = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));
          if (err > tol && it < k_max) pass_kind = 1; else done = true;
        } else {
          if (pass_kind == 1) ++it;
          pass_kind = 2;
        }
      }
      while (consumed != issued) {   // drain the lookahead before the ring is reused
        const unsigned slot = consumed % TMA_NSTAGE, parity = (consumed / TMA_NSTAGE) & 1u;
        while (!mbar_try_wait(&mbar[slot], parity)) {}
        consumed++;
      }
      __syncwarp();
    }
    for (int s = lane; s < nc; s += 32) write_solution(d, w, recs, lam, s);
    if (lane == 0) {
      int* stt = d.stats + (size_t)w * 8;
      stt[4] = it;
      stt[5] = 0;
      stt[6] = (cfm != 0.0);
      stt[7] = ns;
      d.resid[w] = err;
    }
    double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
    bool bad = false;
    for (int b = lane; b < n; b += 32) bad |= integrate_body<BS>(dyn, st, sa, n, b, dt);
    if (__any_sync(0xffffffffu, bad) && lane == 0) d.status[w] |= 16 /*EGG_ST_NONFINITE*/;
    __syncwarp();
  }
}

int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}
int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
bool use_tma_variant() {
  const char* e = getenv("EGG_PGS_VARIANT");
  return e && e[0] == 't';
}
// Variant selection: EGG_PGS_VARIANT = mw | mwpf | fused | fast | tma; default fast
// (egg_pgs_fast.cu: fused residual, cp.async-staged records).
int pgs_variant(const EggDev& d) {   // 0 mw, 1 mwpf, 2 fused, 3 fast, 4 stream
  if (d.rec_fmt) return 4;
  const char* e = getenv("EGG_PGS_VARIANT");
  if (e && e[0] == 'f' && e[1] == 'a') return 3;
  if (e && e[0] == 'f') return 2;
  if (e && e[0] == 'm') return (e[1] == 'w' && e[2] == 'p') ? 1 : 0;
  return 3;
}
bool use_pf_variant(const EggDev& d) { return pgs_variant(d) != 0; }

template <int LPW, bool MINV_SMEM>
void launch_mw2(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int G = 32 / LPW;
  int tabcap = d.nrec + 1;
  if (tabcap > 512) tabcap = 512;
  tabcap = (tabcap + 15) & ~15;
  size_t smem = (size_t)G * (MINV_SMEM ? 17 : 7) * d.n * 8 + (size_t)G * tabcap;
  cudaFuncSetAttribute(egg_pgs_mw_kernel<LPW, MINV_SMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int per_sm = env_int("EGG_PGS_CTAS_PER_SM", 16);
  int groups = (d.W + G - 1) / G;
  int grid = groups < num_sms() * per_sm ? groups : num_sms() * per_sm;
  egg_pgs_mw_kernel<LPW, MINV_SMEM><<<grid, 32, smem, s>>>(d, dt, tabcap, env_int("EGG_PGS_FLAGS", 0));
}
template <int LPW>
void launch_fused(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int G = 32 / LPW;
  int tabcap = d.nrec + 1;
  if (tabcap > 512) tabcap = 512;
  tabcap = (tabcap + 15) & ~15;
  const bool ms = env_int("EGG_PGS_MINV_SMEM", 1) != 0;
  size_t smem = (size_t)G * (ms ? 23 : 13) * d.n * 8 + (size_t)G * tabcap;
  int per_sm = env_int("EGG_PGS_CTAS_PER_SM", ms ? 6 : 9);
  int groups = (d.W + G - 1) / G;
  int grid = groups < num_sms() * per_sm ? groups : num_sms() * per_sm;
  if (ms) {
    cudaFuncSetAttribute(egg_pgs_fused_kernel<LPW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_pgs_fused_kernel<LPW, true><<<grid, 32, smem, s>>>(d, dt, tabcap);
  } else {
    cudaFuncSetAttribute(egg_pgs_fused_kernel<LPW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_pgs_fused_kernel<LPW, false><<<grid, 32, smem, s>>>(d, dt, tabcap);
  }
}
template <int LPW>
void launch_mwpf(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int G = 32 / LPW;
  int tabcap = d.nrec + 1;
  if (tabcap > 512) tabcap = 512;
  tabcap = (tabcap + 15) & ~15;
  const bool ms = env_int("EGG_PGS_MINV_SMEM", 1) != 0;
  size_t smem = (size_t)G * (ms ? 17 : 7) * d.n * 8 + (size_t)G * tabcap;
  int per_sm = env_int("EGG_PGS_CTAS_PER_SM", 9);
  int groups = (d.W + G - 1) / G;
  int grid = groups < num_sms() * per_sm ? groups : num_sms() * per_sm;
  if (ms) {
    cudaFuncSetAttribute(egg_pgs_mwpf_kernel<LPW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_pgs_mwpf_kernel<LPW, true><<<grid, 32, smem, s>>>(d, dt, tabcap);
  } else {
    cudaFuncSetAttribute(egg_pgs_mwpf_kernel<LPW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_pgs_mwpf_kernel<LPW, false><<<grid, 32, smem, s>>>(d, dt, tabcap);
  }
}
template <int LPW>
void launch_mw(const EggDev& d, double dt, cudaStream_t s) {
  // M^-1 in shared memory when the per-warp footprint still allows >= 8 warps per SM.
  constexpr int G = 32 / LPW;
  const bool minv_smem = env_int("EGG_PGS_MINV_SMEM", 1) != 0;
  if (minv_smem) launch_mw2<LPW, true>(d, dt, s);
  else launch_mw2<LPW, false>(d, dt, s);
}

}  // namespace

// Lanes per world of the default solver = maximum blocks per stage the assembly may emit.
int egg_stage_cap(const EggDev& d) {
  if (use_tma_variant()) return TMA_CAP;
  int lpw = env_int("EGG_PGS_LPW", 0);
  if (pgs_variant(d) == 4) {
    // narrow worlds expose little parallelism per level: fewer lanes, more worlds per warp
    // (measured: stack10 87 ms at 1 / 101 at 2 / 133 at 4; legged20 95 ms at 4 / 115 at 2 / 123 at 8)
    // -- unless the batch is too small to give every SM a few warps that way
    if (lpw != 1 && lpw != 2 && lpw != 4 && lpw != 8 && lpw != 16) {
      lpw = (d.n <= 12) ? 1 : (d.n <= 24 ? 4 : 8);
      const long long want = (long long)num_sms() * env_int("EGG_PGS_MIN_WARPS_PER_SM", 4);
      while (lpw < 8 && (long long)d.W * lpw / 32 < want) lpw *= 2;
    }
    return lpw;
  }
  if (lpw != 1 && lpw != 2 && lpw != 4 && lpw != 8 && lpw != 16 && lpw != 32) lpw = 8;
  if (pgs_variant(d) == 3) { if (lpw != 4 && lpw != 8 && lpw != 16) lpw = (d.n <= 12) ? 4 : 8; }
  else if (use_pf_variant(d) && lpw != 4) lpw = 8;
  return lpw;
}

void egg_launch_assemble(const EggDev& d, double dt, cudaStream_t s) {
  if (d.rec_fmt) { egg_launch_assemble_stream(d, dt, s); return; }
  size_t smem = (size_t)(EGG_DYN + EGG_STAT) * d.n * sizeof(double) + (size_t)(7 * d.nrec + d.n + 8) * sizeof(int);
  const int cap = egg_stage_cap(d);
  if (d.nrec <= 128) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(egg_assemble_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_assemble_kernel<64><<<d.W, 64, smem, s>>>(d, dt, cap);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(egg_assemble_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_assemble_kernel<256><<<d.W, 256, smem, s>>>(d, dt, cap);
  }
}

void egg_launch_solve_pgs(const EggDev& d, double dt, cudaStream_t s) {
  if (use_tma_variant()) {
    size_t smem = (((size_t)17 * d.n * 8 + 127) & ~(size_t)127) + (size_t)TMA_NSTAGE * TMA_SLOT_BYTES + TMA_NSTAGE * 8 + (size_t)(TMA_LS_CAP + 1) * 4;
    cudaFuncSetAttribute(egg_pgs_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = env_int("EGG_PGS_CTAS_PER_SM", 12);
    int grid = d.W < num_sms() * per_sm ? d.W : num_sms() * per_sm;
    egg_pgs_tma_kernel<<<grid, 32, smem, s>>>(d, dt);
    return;
  }
  if (pgs_variant(d) == 4) {
    egg_launch_solve_pgs_stream(d, dt, s);
    return;
  }
  if (pgs_variant(d) == 3) {
    egg_launch_solve_pgs_fast(d, dt, egg_stage_cap(d), s);
    return;
  }
  if (pgs_variant(d) == 2) {
    if (egg_stage_cap(d) == 4) launch_fused<4>(d, dt, s);
    else launch_fused<8>(d, dt, s);
    return;
  }
  if (pgs_variant(d) == 1) {
    if (egg_stage_cap(d) == 4) launch_mwpf<4>(d, dt, s);
    else launch_mwpf<8>(d, dt, s);
    return;
  }
  switch (egg_stage_cap(d)) {
    case 1: launch_mw<1>(d, dt, s); break;
    case 2: launch_mw<2>(d, dt, s); break;
    case 4: launch_mw<4>(d, dt, s); break;
    case 8: launch_mw<8>(d, dt, s); break;
    case 16: launch_mw<16>(d, dt, s); break;
    default: launch_mw<32>(d, dt, s); break;
  }
}
