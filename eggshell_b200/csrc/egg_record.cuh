// Builds the compact 240-byte constraint record (REC_* of egg_internal.cuh) of one joint / contact.
// Shared by the per-world assembly (egg_pgs.cu) and the group-stream assembly (egg_pgs_stream.cu).
//
// Replaces: Joint/Contact::ComputeJ + error   joints.cc:3-35, contact.cc:14-117,
//           Ensemble::ComputeJ / rhs          ensembles.cc:38-87, 156-171, 563-570
#pragma once
#include "egg_internal.cuh"

// Eigen 3.3 Quaternion::FromTwoVectors(normal, z).toRotationMatrix()  (utils.cc:233-236).  The
// exactly anti-parallel case uses the same pinned rule as the oracle (orc_linalg.h).
// egg_align_quat: the quaternion (w, x, y, z); egg_quat_rot: its rotation matrix (the run-format
// records of egg_pgs_runs.cu store the quaternion and rebuild the matrix with the same expressions).
__device__ inline void egg_align_quat(d3 nrm, double* q) {
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
  q[0] = qw; q[1] = qx; q[2] = qy; q[3] = qz;
}
__device__ __forceinline__ void egg_quat_rot(double qw, double qx, double qy, double qz, double* R) {
  double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  double twx = tx * qw, twy = ty * qw, twz = tz * qw;
  double txx = tx * qx, txy = ty * qx, txz = tz * qx;
  double tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
  R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
__device__ inline void egg_align_to_z(d3 nrm, double* R) {
  double q[4];
  egg_align_quat(nrm, q);
  egg_quat_rot(q[0], q[1], q[2], q[3], R);
}

// u = v/dt + M^-1 f per body (the vector the rhs multiplies J with, ensembles.cc:569-570), computed
// once per body as the reference does (not once per constraint side): su [6][n] from the world's
// dyn [18][n] and stat [16][n].  Call with all threads of the CTA, then synchronise.
__device__ inline void egg_body_u(int n, const double* sdyn, const double* sst, double dt, double* su, int tid, int nthreads) {
  for (int b = tid; b < n; b += nthreads) {
    const double mi = sst[b];
    double Ii[9];
    for (int k = 0; k < 9; k++) Ii[k] = sst[(1 + k) * n + b];
    d3 v = mk3(sdyn[12 * n + b], sdyn[13 * n + b], sdyn[14 * n + b]);
    d3 wv = mk3(sdyn[15 * n + b], sdyn[16 * n + b], sdyn[17 * n + b]);
    d3 fl = mk3(sst[10 * n + b], sst[11 * n + b], sst[12 * n + b]);
    d3 ft = mk3(sst[13 * n + b], sst[14 * n + b], sst[15 * n + b]);
    d3 ul = v / dt + fl * mi;
    d3 ua = wv / dt + mmulv(Ii, ft);
    su[b] = ul.x; su[n + b] = ul.y; su[2 * n + b] = ul.z;
    su[3 * n + b] = ua.x; su[4 * n + b] = ua.y; su[5 * n + b] = ua.z;
  }
}

// Record of constraint c (joints first, then contacts) of world w whose bodies are (i0, i1).
// sdyn / sst = the world's dyn [18][n] and stat [16][n] arrays, su = egg_body_u's [6][n] (shared
// memory or global).  lam_in_rec: REC_DDIAG holds the multipliers (x0 = rhs) instead of the D diagonal.
__device__ inline void egg_build_record(const EggDev& d, int w, int c, int i0, int i1, const double* sdyn, const double* sst, const double* su,
                                        const double* geom, double dt, bool lam_in_rec, double* v) {
  const int n = d.n, nj = d.nj, maxc = d.maxc;
  const double erp = d.prm.erp, cfm = d.prm.cfm;
  const bool shift = (d.prm.quirks & 1) != 0;
  int kind;
  double Rc[9];
  d3 r0 = mk3(0, 0, 0), r1 = mk3(0, 0, 0), err;
  if (c < nj) {
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
    d3 pos = mk3(geom[0 * maxc + k], geom[1 * maxc + k], geom[2 * maxc + k]);
    d3 nrm = mk3(geom[3 * maxc + k], geom[4 * maxc + k], geom[5 * maxc + k]);
    egg_align_to_z(nrm, Rc);
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
    d3 ul = mk3(su[b], su[n + b], su[2 * n + b]);
    d3 ua = mk3(su[3 * n + b], su[4 * n + b], su[5 * n + b]);
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
  for (int k = 0; k < 9; k++) v[REC_RC + k] = Rc[k];
  v[REC_R0] = r0.x; v[REC_R0 + 1] = r0.y; v[REC_R0 + 2] = r0.z;
  v[REC_R1] = r1.x; v[REC_R1 + 1] = r1.y; v[REC_R1 + 2] = r1.z;
  v[REC_DOFF] = D[3]; v[REC_DOFF + 1] = D[6]; v[REC_DOFF + 2] = D[7];
  const double erp_dt2 = -erp / dt / dt;      // ensembles.cc:569
  for (int k = 0; k < 3; k++) {
    v[REC_INVA + k] = 1.0 / (D[4 * k] + cfm);
    v[REC_RHS + k] = erp_dt2 * get3(err, k) - Ju[k];
    v[REC_DDIAG + k] = lam_in_rec ? v[REC_RHS + k] : D[4 * k];
  }
  v[REC_IDX] = __hiloint2double(i1, i0);
  v[REC_META] = __hiloint2double(ckind, c);
  v[29] = 0.0;
}
