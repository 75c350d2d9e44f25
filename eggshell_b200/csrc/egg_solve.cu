// Init, constraint-row assembly, projected Gauss-Seidel solve + fused integrate, rollout cost.
//
// Replaces (per world):
//   egg_init_kernel      Ensemble::Init pieces            ensembles.cc:24-29, 202-232
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
// keep the compact record (Rc, r0, r1, the 3x3 diagonal block D of J M^-1 J^T, rhs) and the
// body-space accumulator a = M^-1 J^T x (6 doubles per body, in shared memory).  One block update
// is then  t = Rc (vel1(a) - vel0(a)),  row-by-row projected substitution inside the 3x3 diagonal
// block exactly as sparse_iterations_utils.cc:229-236, and an impulse scatter back into a.
//
// Gauss-Seidel is sequential in constraint order.  Blocks that share no body commute exactly, so
// blocks are grouped into dependency levels (level(c) = 1 + max level of any earlier block that
// shares a body); running level after level, lanes in parallel inside a level, is bit-identical
// to the sequential sweep.  One warp owns one world; levels are separated by __syncwarp().
#include "egg_internal.cuh"

namespace {

__device__ inline void inverse3(const double* m, double* r) {
  double c00 = m[4] * m[8] - m[5] * m[7];
  double c01 = m[5] * m[6] - m[3] * m[8];
  double c02 = m[3] * m[7] - m[4] * m[6];
  double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
  double id = 1.0 / det;
  r[0] = c00 * id; r[3] = c01 * id; r[6] = c02 * id;
  r[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  r[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  r[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  r[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  r[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  r[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// ---------------------------------------------------------------------------------------------
// Init: thread per (world, body).
__global__ void egg_init_kernel(EggDev d) {
  const int n = d.n;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)d.W * n) return;
  const int w = (int)(gid / n), b = (int)(gid % n);
  const double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  const double* bp = d.bpar + (size_t)w * EGG_BPAR * n;
  double* st = d.stat + (size_t)w * EGG_STAT * n;
  double R[9], I[9], RI[9], Ig[9], Rt[9], inv[9];
  for (int k = 0; k < 9; k++) { R[k] = dyn[(3 + k) * n + b]; I[k] = bp[(4 + k) * n + b]; }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Rt[3 * i + j] = R[3 * j + i];
  mmulm(R, I, RI);
  mmulm(RI, Rt, Ig);           // I_g = (R I) R^T, body.h:58
  inverse3(Ig, inv);           // ensembles.cc:210
  const double m = bp[3 * n + b];
  st[0 * n + b] = 1.0 / m;     // ensembles.cc:207
  for (int k = 0; k < 9; k++) st[(1 + k) * n + b] = inv[k];
  d3 wv = mk3(dyn[15 * n + b], dyn[16 * n + b], dyn[17 * n + b]);
  // f_ext = [m g ; ((-[w]x) I_g) w]   ensembles.cc:218-220
  double ncm[9] = {-0.0, wv.z, -wv.y, -wv.z, -0.0, wv.x, wv.y, -wv.x, -0.0};
  double t[9];
  mmulm(ncm, Ig, t);
  d3 tq = mmulv(t, wv);
  st[10 * n + b] = m * d.prm.g[0];
  st[11 * n + b] = m * d.prm.g[1];
  st[12 * n + b] = m * d.prm.g[2];
  st[13 * n + b] = tq.x; st[14 * n + b] = tq.y; st[15 * n + b] = tq.z;
  if (b == 0) {
    d.cost0[2 * w] = dyn[0 * n];       // x of body 0
    d.cost0[2 * w + 1] = dyn[2 * n];   // z of body 0
    d.status[w] = 0;
    for (int k = 0; k < 8; k++) d.stats[(size_t)w * 8 + k] = 0;
    d.resid[w] = 0;
    d.c_count[w] = 0;
  }
}

// CheckInitialConditions (ensembles.cc:224-232): every joint error component within 1e-9.
__global__ void egg_init_check_kernel(EggDev d) {
  const int n = d.n, nj = d.nj;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (long long)d.W * nj) return;
  const int w = (int)(gid / nj), k = (int)(gid % nj);
  const double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  const double* jc = d.jc + (size_t)w * 6 * nj;
  const int i0 = d.j_i0[(size_t)w * nj + k], i1 = d.j_i1[(size_t)w * nj + k];
  double R0[9];
  for (int c = 0; c < 9; c++) R0[c] = dyn[(3 + c) * n + i0];
  d3 e = mk3(dyn[i0], dyn[n + i0], dyn[2 * n + i0]) + mmulv(R0, mk3(jc[k], jc[nj + k], jc[2 * nj + k]));
  d3 c1 = mk3(jc[3 * nj + k], jc[4 * nj + k], jc[5 * nj + k]);
  if (i1 < 0) e = e - c1;
  else {
    double R1[9];
    for (int c = 0; c < 9; c++) R1[c] = dyn[(3 + c) * n + i1];
    e = e - mk3(dyn[i1], dyn[n + i1], dyn[2 * n + i1]) - mmulv(R1, c1);
  }
  if (fabs(e.x) > 1e-9 || fabs(e.y) > 1e-9 || fabs(e.z) > 1e-9) atomicOr(&d.status[w], 4 /*EGG_ST_BAD_INIT*/);
}

// ---------------------------------------------------------------------------------------------
// Eigen 3.3 Quaternion::FromTwoVectors(normal, z).toRotationMatrix()  (utils.cc:233-236).  The
// exactly anti-parallel case uses the same pinned rule as the oracle (orc_linalg.h).
__device__ inline void align_to_z(d3 nrm, double* R) {
  double z2 = dot3(nrm, nrm);
  d3 v0 = (z2 > 0) ? nrm / sqrt(z2) : nrm;
  double c = v0.z;   // dot(v1 = (0,0,1), v0)
  double qw, qx, qy, qz;
  if (c < -1.0 + 1e-12) {
    c = fmax(c, -1.0);
    int k = 0;
    if (fabs(v0.y) < fabs(get3(v0, k))) k = 1;
    if (fabs(v0.z) < fabs(get3(v0, k))) k = 2;
    d3 e = mk3(k == 0, k == 1, k == 2);
    d3 ax = cross3(v0, e);
    double a2 = dot3(ax, ax);
    if (a2 > 0) ax = ax / sqrt(a2);
    double w2 = (1.0 + c) * 0.5;
    qw = sqrt(w2);
    double s = sqrt(1.0 - w2);
    qx = ax.x * s; qy = ax.y * s; qz = ax.z * s;
  } else {
    d3 ax = cross3(v0, mk3(0, 0, 1));
    double s = sqrt((1.0 + c) * 2.0);
    double invs = 1.0 / s;
    qx = ax.x * invs; qy = ax.y * invs; qz = ax.z * invs;
    qw = s * 0.5;
  }
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}

// Assembly: one CTA per world, thread per constraint; then levels + ordered scatter.
template <int NT>
__global__ void __launch_bounds__(NT) egg_assemble_kernel(EggDev d, double dt) {
  extern __shared__ double sm[];
  const int n = d.n, nj = d.nj, w = blockIdx.x, tid = threadIdx.x;
  double* sdyn = sm;                         // [18][n]
  double* sst = sm + EGG_DYN * n;            // [16][n]
  int* lvl = (int*)(sm + (EGG_DYN + EGG_STAT) * n);   // [nrec] level of constraint c, then its slot
  int* blast = lvl + d.nrec;                 // [n] last level touching the body
  int* lcount = blast + n;                   // [nrec + 1]
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
  __syncthreads();

  // Levels: sequential over constraints in reference order (joints, then contacts).
  if (tid == 0) {
    for (int b = 0; b < n; b++) blast[b] = -1;
    for (int c = 0; c <= nc; c++) lcount[c] = 0;
    int nl = 0;
    for (int c = 0; c < nc; c++) {
      int i0, i1;
      if (c < nj) { i0 = d.j_i0[(size_t)w * nj + c]; i1 = d.j_i1[(size_t)w * nj + c]; }
      else { i0 = c_i0[c - nj]; i1 = c_i1[c - nj]; }
      int l = -1;
      if (i0 >= 0) l = max(l, blast[i0]);
      if (i1 >= 0) l = max(l, blast[i1]);
      l += 1;
      if (i0 >= 0) blast[i0] = l;
      if (i1 >= 0) blast[i1] = l;
      lvl[c] = l;
      lcount[l + 1]++;
      nl = max(nl, l + 1);
    }
    for (int l = 0; l < nl; l++) lcount[l + 1] += lcount[l];     // lcount[l] = start of level l
    int* ls = d.level_start + (size_t)w * (d.nrec + 1);
    for (int l = 0; l <= nl; l++) ls[l] = lcount[l];
    d.n_levels[w] = nl;
    for (int c = 0; c < nc; c++) { int l = lvl[c]; lvl[c] = lcount[l]++; }   // slot of c (stable)
  }
  __syncthreads();

  const double erp = d.prm.erp, cfm = d.prm.cfm;
  const bool shift = (d.prm.quirks & 1) != 0;
  for (int c = tid; c < nc; c += NT) {
    int i0, i1, kind;
    double Rc[9];
    d3 r0 = mk3(0, 0, 0), r1 = mk3(0, 0, 0), err;
    if (c < nj) {
      i0 = d.j_i0[(size_t)w * nj + c];
      i1 = d.j_i1[(size_t)w * nj + c];
      const double* jc = d.jc + (size_t)w * 6 * nj;
      d3 c0 = mk3(jc[c], jc[nj + c], jc[2 * nj + c]);
      d3 c1 = mk3(jc[3 * nj + c], jc[4 * nj + c], jc[5 * nj + c]);
      for (int k = 0; k < 9; k++) Rc[k] = 0;
      Rc[0] = Rc[4] = Rc[8] = -1.0;
      double R0[9];
      for (int k = 0; k < 9; k++) R0[k] = sdyn[(3 + k) * n + i0];
      r0 = mmulv(R0, c0);
      d3 p0 = mk3(sdyn[i0], sdyn[n + i0], sdyn[2 * n + i0]);
      if (i1 < 0) {
        err = p0 + r0 - c1;                       // joints.cc:6
      } else {
        double R1[9];
        for (int k = 0; k < 9; k++) R1[k] = sdyn[(3 + k) * n + i1];
        r1 = mmulv(R1, c1);
        d3 p1 = mk3(sdyn[i1], sdyn[n + i1], sdyn[2 * n + i1]);
        err = p0 + r0 - p1 - r1;                  // joints.cc:8
      }
      kind = KIND_EQUALITY;
    } else {
      const int k = c - nj;
      i0 = c_i0[k];
      i1 = c_i1[k];
      d3 pos = mk3(geom[0 * maxc + k], geom[1 * maxc + k], geom[2 * maxc + k]);
      d3 nrm = mk3(geom[3 * maxc + k], geom[4 * maxc + k], geom[5 * maxc + k]);
      align_to_z(nrm, Rc);
      if (i0 >= 0) r0 = pos - mk3(sdyn[i0], sdyn[n + i0], sdyn[2 * n + i0]);
      if (i1 >= 0) r1 = pos - mk3(sdyn[i1], sdyn[n + i1], sdyn[2 * n + i1]);
      err = mk3(0, 0, -geom[6 * maxc + k]);      // contact.cc:14-22
      kind = KIND_CONTACT;
    }
    // q2: the matrix-free lower-triangular solve projects block c > 0 with the (type, lo, hi) of
    // block c-1 (sparse_iterations_utils.cc:169,180,229-235).
    int ckind = kind;
    if (shift && c > 0) ckind = (c - 1 < nj) ? KIND_EQUALITY : KIND_CONTACT;

    // Jacobian rows: body0 lin = -Rc_k, ang = Rc_k x r0 ; body1 lin = Rc_k, ang = r1 x Rc_k.
    d3 jl[3], ja0[3], ja1[3];
    for (int k = 0; k < 3; k++) {
      jl[k] = mrow(Rc, k);
      ja0[k] = cross3(jl[k], r0);
      ja1[k] = cross3(r1, jl[k]);
    }
    // D = J0 M0^-1 J0^T + J1 M1^-1 J1^T and J u with u = v/dt + M^-1 f (ensembles.cc:569-570).
    double D[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double Ju[3] = {0, 0, 0};
    for (int side = 0; side < 2; side++) {
      const int b = side ? i1 : i0;
      if (b < 0) continue;
      const double mi = sst[b];
      double Ii[9];
      for (int k = 0; k < 9; k++) Ii[k] = sst[(1 + k) * n + b];
      const double sg = side ? 1.0 : -1.0;
      d3 v = mk3(sdyn[12 * n + b], sdyn[13 * n + b], sdyn[14 * n + b]);
      d3 wv = mk3(sdyn[15 * n + b], sdyn[16 * n + b], sdyn[17 * n + b]);
      d3 fl = mk3(sst[10 * n + b], sst[11 * n + b], sst[12 * n + b]);
      d3 ft = mk3(sst[13 * n + b], sst[14 * n + b], sst[15 * n + b]);
      d3 ul = v / dt + fl * mi;
      d3 ua = wv / dt + mmulv(Ii, ft);
      for (int k = 0; k < 3; k++) {
        d3 lin = jl[k] * sg;
        d3 ang = side ? ja1[k] : ja0[k];
        d3 Ia = mmulv(Ii, ang);
        for (int l = 0; l < 3; l++) {
          d3 lin2 = jl[l] * sg;
          d3 ang2 = side ? ja1[l] : ja0[l];
          D[3 * k + l] += mi * dot3(lin, lin2) + dot3(Ia, ang2);
        }
        Ju[k] += dot3(lin, ul) + dot3(ang, ua);
      }
    }
    const int slot = lvl[c];
    double* rec = d.rec + ((size_t)w * d.nrec + slot) * EGG_REC;
    for (int k = 0; k < 9; k++) rec[REC_RC + k] = Rc[k];
    rec[REC_R0] = r0.x; rec[REC_R0 + 1] = r0.y; rec[REC_R0 + 2] = r0.z;
    rec[REC_R1] = r1.x; rec[REC_R1 + 1] = r1.y; rec[REC_R1 + 2] = r1.z;
    rec[REC_DOFF] = D[3]; rec[REC_DOFF + 1] = D[6]; rec[REC_DOFF + 2] = D[7];
    for (int k = 0; k < 3; k++) {
      rec[REC_DDIAG + k] = D[4 * k];
      rec[REC_INVA + k] = 1.0 / (D[4 * k] + cfm);
      rec[REC_RHS + k] = -erp / dt / dt * get3(err, k) - Ju[k];
      rec[REC_ERR + k] = get3(err, k);
    }
    int2 idx = make_int2(i0, i1), meta = make_int2(c, ckind);
    reinterpret_cast<int2*>(rec)[REC_IDX] = idx;
    reinterpret_cast<int2*>(rec)[REC_META] = meta;
  }
}

// ---------------------------------------------------------------------------------------------
// PGS: one warp per world, WPB worlds per CTA, grid-stride over worlds.

struct BlockRec {
  double Rc[9];
  d3 r0, r1;
  double doff[3], ddiag[3], inva[3], rhs[3];
  int i0, i1, orig, kind;
};

__device__ inline void load_rec(const double* __restrict__ rec, BlockRec& r) {
  const double2* p = reinterpret_cast<const double2*>(rec);
  double v[30];
#pragma unroll
  for (int k = 0; k < 15; k++) { double2 t = __ldg(p + k); v[2 * k] = t.x; v[2 * k + 1] = t.y; }
#pragma unroll
  for (int k = 0; k < 9; k++) r.Rc[k] = v[REC_RC + k];
  r.r0 = mk3(v[REC_R0], v[REC_R0 + 1], v[REC_R0 + 2]);
  r.r1 = mk3(v[REC_R1], v[REC_R1 + 1], v[REC_R1 + 2]);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    r.doff[k] = v[REC_DOFF + k]; r.ddiag[k] = v[REC_DDIAG + k]; r.inva[k] = v[REC_INVA + k]; r.rhs[k] = v[REC_RHS + k];
  }
  r.i0 = __double2loint(v[REC_IDX]); r.i1 = __double2hiint(v[REC_IDX]);
  r.orig = __double2loint(v[REC_META]); r.kind = __double2hiint(v[REC_META]);
}

// t = J a for the block: Rc (vel1 - vel0), vel_b = a_lin + a_ang x r_b.
__device__ inline d3 block_Ja(const BlockRec& r, const double* sa, int n) {
  d3 u = mk3(0, 0, 0);
  if (r.i1 >= 0) {
    const int b = r.i1;
    d3 al = mk3(sa[b], sa[n + b], sa[2 * n + b]);
    d3 aa = mk3(sa[3 * n + b], sa[4 * n + b], sa[5 * n + b]);
    u = al + cross3(aa, r.r1);
  }
  if (r.i0 >= 0) {
    const int b = r.i0;
    d3 al = mk3(sa[b], sa[n + b], sa[2 * n + b]);
    d3 aa = mk3(sa[3 * n + b], sa[4 * n + b], sa[5 * n + b]);
    u = u - (al + cross3(aa, r.r0));
  }
  return mmulv(r.Rc, u);
}

// a += M^-1 J^T delta for the block.
__device__ inline void block_scatter(const BlockRec& r, d3 delta, double* sa, const double* sminv, int n) {
  d3 imp = mtmulv(r.Rc, delta);
  if (r.i1 >= 0) {
    const int b = r.i1;
    const double mi = sminv[b];
    double Ii[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Ii[k] = sminv[(1 + k) * n + b];
    d3 da = mmulv(Ii, cross3(r.r1, imp));
    sa[b] += mi * imp.x; sa[n + b] += mi * imp.y; sa[2 * n + b] += mi * imp.z;
    sa[3 * n + b] += da.x; sa[4 * n + b] += da.y; sa[5 * n + b] += da.z;
  }
  if (r.i0 >= 0) {
    const int b = r.i0;
    const double mi = sminv[b];
    double Ii[9];
#pragma unroll
    for (int k = 0; k < 9; k++) Ii[k] = sminv[(1 + k) * n + b];
    d3 da = mmulv(Ii, cross3(r.r0, imp));
    sa[b] -= mi * imp.x; sa[n + b] -= mi * imp.y; sa[2 * n + b] -= mi * imp.z;
    sa[3 * n + b] -= da.x; sa[4 * n + b] -= da.y; sa[5 * n + b] -= da.z;
  }
}

__device__ inline double project(double x, int kind, int row) {   // sparse_iterations_utils.cc:12-21
  if (kind == KIND_CONTACT) {
    if (row < 2) { if (x < -1.0) return -1.0; else if (x > 1.0) return 1.0; }
    else { if (x < 0.0) return 0.0; }
  }
  return x;
}

template <int WPB>
__global__ void __launch_bounds__(WPB * 32) egg_pgs_kernel(EggDev d, double dt) {
  extern __shared__ double sm[];
  const int n = d.n, lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double* sa = sm + (size_t)wib * 16 * n;   // [6][n] accumulator a = M^-1 J^T x
  double* sminv = sa + 6 * n;               // [10][n] 1/m, Iinv
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max;
  const int nj = d.nj;

  for (int w = blockIdx.x * WPB + wib; w < d.W; w += gridDim.x * WPB) {
    const double* st = d.stat + (size_t)w * EGG_STAT * n;
    for (int i = lane; i < 10 * n; i += 32) sminv[i] = st[i];
    for (int i = lane; i < 6 * n; i += 32) sa[i] = 0.0;
    const int nc = nj + d.c_count[w];
    const int nl = d.n_levels[w];
    const int* ls = d.level_start + (size_t)w * (d.nrec + 1);
    const double* recs = d.rec + (size_t)w * d.nrec * EGG_REC;
    double* lam = d.lam + (size_t)w * d.nrec * 3;
    __syncwarp();

    // x0 = rhs (sparse_iterations.cc:202); a = M^-1 J^T x0 accumulated in level order.
    for (int l = 0; l < nl; l++) {
      const int s0 = ls[l], s1 = ls[l + 1];
      for (int s = s0 + lane; s < s1; s += 32) {
        BlockRec r;
        load_rec(recs + (size_t)s * EGG_REC, r);
        lam[3 * s] = r.rhs[0]; lam[3 * s + 1] = r.rhs[1]; lam[3 * s + 2] = r.rhs[2];
        block_scatter(r, mk3(r.rhs[0], r.rhs[1], r.rhs[2]), sa, sminv, n);
      }
      __syncwarp();
    }

    // GetResidualError (sparse_iterations.cc:51-69): w = A x - rhs, four partial 2-norms.
    auto residual = [&]() -> double {
      double se = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int s = lane; s < nc; s += 32) {
        BlockRec r;
        load_rec(recs + (size_t)s * EGG_REC, r);
        d3 t = block_Ja(r, sa, n);
        // the reference classifies with each block's OWN bounds here (ConstructMixedConstraints)
        const bool eq = r.orig < nj;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          double x = lam[3 * s + k];
          double wv = get3(t, k) + cfm * x - r.rhs[k];
          if (eq) { se += wv * wv; continue; }
          double lo = (k < 2) ? -1.0 : 0.0;
          bool has_hi = (k < 2);
          if (x == lo && wv < 0) s1 += wv * wv;
          if (has_hi && x == 1.0 && wv > 0) s2 += wv * wv;
          if (x > lo && (!has_hi || x < 1.0)) s3 += wv * wv;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        se += __shfl_xor_sync(0xffffffffu, se, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      }
      return sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));
    };

    double err = (nc > 0) ? residual() : 0.0;
    int it = 0;
    while (err > tol && it < k_max) {
      for (int l = 0; l < nl; l++) {
        const int s0 = ls[l], s1 = ls[l + 1];
        for (int s = s0 + lane; s < s1; s += 32) {
          BlockRec r;
          load_rec(recs + (size_t)s * EGG_REC, r);
          d3 t = block_Ja(r, sa, n);
          double x0 = lam[3 * s], x1 = lam[3 * s + 1], x2 = lam[3 * s + 2];
          // row-by-row substitution inside the 3x3 diagonal block (sparse_iterations_utils.cc:229-236)
          double n0 = project((r.rhs[0] - t.x + r.ddiag[0] * x0) * r.inva[0], r.kind, 0);
          double d0 = n0 - x0;
          double n1 = project((r.rhs[1] - (t.y + r.doff[0] * d0) + r.ddiag[1] * x1) * r.inva[1], r.kind, 1);
          double d1 = n1 - x1;
          double n2 = project((r.rhs[2] - (t.z + r.doff[1] * d0 + r.doff[2] * d1) + r.ddiag[2] * x2) * r.inva[2], r.kind, 2);
          double d2 = n2 - x2;
          lam[3 * s] = n0; lam[3 * s + 1] = n1; lam[3 * s + 2] = n2;
          block_scatter(r, mk3(d0, d1, d2), sa, sminv, n);
        }
        __syncwarp();
      }
      err = residual();
      ++it;
    }

    // Multipliers / row state in reference row order.
    double* lo_out = d.lam_out + (size_t)w * 3 * d.nrec;
    int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
    for (int s = lane; s < nc; s += 32) {
      const double* rec = recs + (size_t)s * EGG_REC;
      const int orig = __double2loint(rec[REC_META]);
      const bool eq = orig < nj;
      for (int k = 0; k < 3; k++) {
        double x = lam[3 * s + k];
        lo_out[3 * orig + k] = x;
        int state = 0;
        if (eq) state = 3;
        else if (x == ((k < 2) ? -1.0 : 0.0)) state = 1;
        else if (k < 2 && x == 1.0) state = 2;
        rs_out[3 * orig + k] = state;
      }
    }
    if (lane == 0) {
      int* stt = d.stats + (size_t)w * 8;
      stt[4] = it;
      stt[5] = 0;
      stt[6] = (cfm != 0.0);
      d.resid[w] = err;
    }

    // v' = v + dt (M^-1 f + a); p += dt (v+v')/2; R <- WtoQ((w+w')/2, dt) R.
    double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
    bool bad = false;
    for (int b = lane; b < n; b += 32) {
      const double mi = sminv[b];
      double Ii[9];
      for (int k = 0; k < 9; k++) Ii[k] = sminv[(1 + k) * n + b];
      d3 fl = mk3(st[10 * n + b], st[11 * n + b], st[12 * n + b]);
      d3 ft = mk3(st[13 * n + b], st[14 * n + b], st[15 * n + b]);
      d3 v = mk3(dyn[12 * n + b], dyn[13 * n + b], dyn[14 * n + b]);
      d3 wv = mk3(dyn[15 * n + b], dyn[16 * n + b], dyn[17 * n + b]);
      d3 al = mk3(sa[b], sa[n + b], sa[2 * n + b]);
      d3 aa = mk3(sa[3 * n + b], sa[4 * n + b], sa[5 * n + b]);
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
      for (int k = 0; k < 9; k++) R[k] = dyn[(3 + k) * n + b];
      mmulm(Q, R, Rn);
      dyn[b] = p.x; dyn[n + b] = p.y; dyn[2 * n + b] = p.z;
      for (int k = 0; k < 9; k++) dyn[(3 + k) * n + b] = Rn[k];
      dyn[12 * n + b] = vn.x; dyn[13 * n + b] = vn.y; dyn[14 * n + b] = vn.z;
      dyn[15 * n + b] = wn.x; dyn[16 * n + b] = wn.y; dyn[17 * n + b] = wn.z;
      double chk = p.x + p.y + p.z + vn.x + vn.y + vn.z + wn.x + wn.y + wn.z;
      if (!(fabs(chk) < 1e300)) bad = true;
    }
    if (__any_sync(0xffffffffu, bad) && lane == 0) d.status[w] |= 16 /*EGG_ST_NONFINITE*/;
    __syncwarp();
  }
}

__global__ void egg_cost_kernel(EggDev d, double* cost) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.W) return;
  const double* dyn = d.dyn + (size_t)w * EGG_DYN * d.n;
  double dx = dyn[0] - d.cost0[2 * w];
  double dz = dyn[2 * d.n] - d.cost0[2 * w + 1];
  cost[w] = -dx + 10.0 * dz * dz;
}

// AoS [W][per_world][comps] (host order) -> SoA [W][comps_total][per_world] at component offset.
__global__ void egg_pack_kernel(const double* aos, int W, int per_world, int comps, double* soa, int soa_comps, int comp_off) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)W * per_world * comps;
  if (gid >= total) return;
  const int c = (int)(gid % comps);
  const long long t = gid / comps;
  const int b = (int)(t % per_world);
  const int w = (int)(t / per_world);
  soa[((size_t)w * soa_comps + comp_off + c) * per_world + b] = aos[gid];
}
__global__ void egg_unpack_kernel(double* aos, int W, int per_world, int comps, const double* soa, int soa_comps, int comp_off) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)W * per_world * comps;
  if (gid >= total) return;
  const int c = (int)(gid % comps);
  const long long t = gid / comps;
  const int b = (int)(t % per_world);
  const int w = (int)(t / per_world);
  aos[gid] = soa[((size_t)w * soa_comps + comp_off + c) * per_world + b];
}

}  // namespace

void egg_launch_init(const EggDev& d, cudaStream_t s) {
  long long t = (long long)d.W * d.n;
  egg_init_kernel<<<(unsigned)((t + 127) / 128), 128, 0, s>>>(d);
  if (d.nj > 0) {
    long long tj = (long long)d.W * d.nj;
    egg_init_check_kernel<<<(unsigned)((tj + 127) / 128), 128, 0, s>>>(d);
  }
}

void egg_launch_assemble(const EggDev& d, double dt, cudaStream_t s) {
  size_t smem = (size_t)(EGG_DYN + EGG_STAT) * d.n * sizeof(double) + (size_t)(2 * d.nrec + d.n + 2) * sizeof(int);
  if (d.nrec <= 128) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(egg_assemble_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_assemble_kernel<64><<<d.W, 64, smem, s>>>(d, dt);
  } else {
    if (smem > 48 * 1024) cudaFuncSetAttribute(egg_assemble_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    egg_assemble_kernel<256><<<d.W, 256, smem, s>>>(d, dt);
  }
}

static int g_num_sms = 0;
void egg_launch_solve_pgs(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int WPB = 4;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  size_t smem = (size_t)WPB * 16 * d.n * sizeof(double);
  if (smem > 48 * 1024) cudaFuncSetAttribute(egg_pgs_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int blocks_needed = (d.W + WPB - 1) / WPB;
  int resident = g_num_sms * 4;   // 4 CTAs x 4 warps per SM: 16 worlds in flight per SM
  int grid = blocks_needed < resident ? blocks_needed : resident;
  egg_pgs_kernel<WPB><<<grid, WPB * 32, smem, s>>>(d, dt);
}

void egg_launch_costs(const EggDev& d, double* cost_d, cudaStream_t s) {
  egg_cost_kernel<<<(d.W + 127) / 128, 128, 0, s>>>(d, cost_d);
}

void egg_launch_pack(int W, const double* aos, int per_world, int comps, double* soa, int soa_comps, int comp_off, cudaStream_t s) {
  long long total = (long long)W * per_world * comps;
  if (total == 0) return;
  egg_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(aos, W, per_world, comps, soa, soa_comps, comp_off);
}
void egg_launch_unpack(int W, double* aos, int per_world, int comps, const double* soa, int soa_comps, int comp_off, cudaStream_t s) {
  long long total = (long long)W * per_world * comps;
  if (total == 0) return;
  egg_unpack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(aos, W, per_world, comps, soa, soa_comps, comp_off);
}
