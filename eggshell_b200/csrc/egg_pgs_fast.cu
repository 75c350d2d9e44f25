// Variant "fast" of the projected Gauss-Seidel solve: the fused-residual algorithm of
// egg_pgs_fused_kernel (egg_pgs.cu) with the per-stage instruction stream cut down.
//
// ncu on the earlier variants (profiles/r1b_pgs_mwpf8_summary.txt) showed ~440 warp instructions
// per stage of which only 19 % were FP64 and 14 % memory: the rest was the register rotation of
// the prefetch double buffer (60 moves), 64-bit index arithmetic, scalar 8-byte shared-memory
// accesses, re-loads of the accumulator for the scatter, and divergence bookkeeping for the
// "body index -1" branches.  This kernel removes those:
//   * the stage loop is unrolled by two with ping-pong record buffers (no rotation moves);
//   * the ground / world-anchor side of a block points at a dummy body with M^-1 = 0, so gather
//     and scatter are branch-free;
//   * the per-body struct is 16-byte aligned (stride 22 / 14 doubles: conflict-free for 128-bit
//     accesses) and moved with LDS.128 / STS.128; the gathered accumulator is reused for the
//     scatter instead of being read again;
//   * projection and residual classification are select-based.
// Arithmetic per block is the same expression tree as egg_pgs.cu (same results).
//
// Replaces: sparse::GaussSeidelIteration + GetResidualError + the velocity/position update, i.e.
// /root/reference/eggshell/sparse_iterations.cc:148-226,51-69,
// sparse_iterations_utils.cc:12-21,159-243,495-695, ensembles.cc:535,572-591.
#include "egg_internal.cuh"
#include <math_constants.h>

namespace {

#define kInf CUDART_INF

// MV = where M^-1 of a block's two bodies comes from: 0 read-only global (scattered LDG), 1 the
// shared-memory body struct, 2 a per-block copy staged together with the record (coalesced).
template <int MV> struct FS { static constexpr int value = (MV == 1) ? 22 : 14; };   // body stride (doubles)

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }

// record pieces (double2 v[15]) -> fields
#define RC0 v[0].x
#define RC1 v[0].y
#define RC2 v[1].x
#define RC3 v[1].y
#define RC4 v[2].x
#define RC5 v[2].y
#define RC6 v[3].x
#define RC7 v[3].y
#define RC8 v[4].x
#define R0X v[4].y
#define R0Y v[5].x
#define R0Z v[5].y
#define R1X v[6].x
#define R1Y v[6].y
#define R1Z v[7].x
#define DO0 v[7].y
#define DO1 v[8].x
#define DO2 v[8].y
#define DD0 v[9].x
#define DD1 v[9].y
#define DD2 v[10].x
#define IA0 v[10].y
#define IA1 v[11].x
#define IA2 v[11].y
#define RH0 v[12].x
#define RH1 v[12].y
#define RH2 v[13].x
#define IDX v[13].y
#define MET v[14].x

// relative velocity-like vector u = (l1 + a1 x r1) - (l0 + a0 x r0), then t = Rc u
__device__ __forceinline__ V3 rel_t(const double2* v, const double2& p1a, const double2& p1b, const double2& p1c,
                                    const double2& p0a, const double2& p0b, const double2& p0c) {
  // body struct pieces: (l.x,l.y) (l.z,a.x) (a.y,a.z)
  const double l1x = p1a.x, l1y = p1a.y, l1z = p1b.x, a1x = p1b.y, a1y = p1c.x, a1z = p1c.y;
  const double l0x = p0a.x, l0y = p0a.y, l0z = p0b.x, a0x = p0b.y, a0y = p0c.x, a0z = p0c.y;
  const double u1x = l1x + (a1y * R1Z - a1z * R1Y), u1y = l1y + (a1z * R1X - a1x * R1Z), u1z = l1z + (a1x * R1Y - a1y * R1X);
  const double u0x = l0x + (a0y * R0Z - a0z * R0Y), u0y = l0y + (a0z * R0X - a0x * R0Z), u0z = l0z + (a0x * R0Y - a0y * R0X);
  const double ux = u1x - u0x, uy = u1y - u0y, uz = u1z - u0z;
  return v3(RC0 * ux + RC1 * uy + RC2 * uz, RC3 * ux + RC4 * uy + RC5 * uz, RC6 * ux + RC7 * uy + RC8 * uz);
}

template <int LPW, int MV, bool STAGED>
__global__ void __launch_bounds__(32) egg_pgs_fast_kernel(EggDev d, double dt, int tabcap) {
  constexpr int G = 32 / LPW;
  constexpr int BS = FS<MV>::value;
  constexpr bool MS = (MV == 1);
  static_assert(MV != 2 || STAGED, "per-block M^-1 needs the staged path");
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x, sub = lane / LPW, sl = lane % LPW;
  const int nb = n + 1;                                   // + dummy body (index n) with M^-1 = 0
  double* sb = reinterpret_cast<double*>(smraw) + (size_t)sub * BS * nb;
  // STAGED: one staging buffer of LPW records per world, filled by coalesced cp.async
  constexpr int STG_REC = LPW * EGG_REC * 8;                 // staged records
  constexpr int STG = STAGED ? STG_REC + (MV == 2 ? LPW * 160 : 0) : 0;   // + per-block M^-1 (20 doubles)
  unsigned char* stage = smraw + (size_t)G * BS * nb * 8 + (size_t)sub * STG;
  unsigned char* tab = smraw + (size_t)G * (BS * nb * 8 + STG) + (size_t)sub * tabcap;
  const double cfm = d.prm.cfm, tol = d.prm.tol;
  const int k_max = d.prm.k_max, nj = d.nj;

  for (int wbase = blockIdx.x * G; wbase < d.W; wbase += gridDim.x * G) {
    const int w = wbase + sub;
    const bool valid = w < d.W;
    const int wc = valid ? w : d.W - 1;
    const double* st = d.stat + (size_t)wc * EGG_STAT * n;
    const double* maos = d.minv_aos + (size_t)wc * nb * 10;   // [n+1][10], row n = 0
    const int nc = valid ? nj + d.c_count[wc] : 0;
    const int ns = valid ? d.n_levels[wc] : 0;
    const int* gls = d.level_start + (size_t)wc * (d.nrec + 1);
    const char* recs = reinterpret_cast<const char*>(d.rec + (size_t)wc * d.nrec * EGG_REC);
    double* const lam0 = d.lam + (size_t)wc * d.nrec * 3;
    double* const lam1 = d.lam2 + (size_t)wc * d.nrec * 3;
#define LAMB(which) ((which) ? lam1 : lam0)
    for (int i = sl; i < nb * BS; i += LPW) {
      const int b = i / BS, f = i - b * BS;
      sb[i] = (MS && f >= 12 && f < 22) ? maos[b * 10 + (f - 12)] : 0.0;
    }
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
    int it = 0, k = 0, rd = 0;
    int pass_kind = 0;   // 0: x0 = rhs scatter, 3: residual(k) + update(k -> k+1), 2: residual only
    double2 bufA[EGG_PIECES], bufB[EGG_PIECES];
    double lA0 = 0, lA1 = 0, lA2 = 0, lB0 = 0, lB1 = 0, lB2 = 0;
    int slotA = -1, slotB = -1;
    int s0_next = 0;
    int t_pf = 0, s0_pf = 0;   // L2 prefetch cursor, two stages ahead of the consumer
    double se = 0, s1 = 0, s2 = 0, s3 = 0;

    auto next_slot = [&](int kind, int t) -> int {          // slot of this lane in step t of a pass of `kind`
      if (kind == 2) {
        const int f = t * LPW + sl;
        return (active && f < nc) ? f : -1;
      }
      if (!active || t >= ns) return -1;
      const int cnt = stage_cnt(t);
      const int mine = (sl < cnt) ? s0_next + sl : -1;
      s0_next += cnt;
      return mine;
    };
#define FETCH(BUF, L0, L1, L2, SLOT, slot_expr, with_lam)                                          \
  {                                                                                                \
    SLOT = (slot_expr);                                                                            \
    if (SLOT >= 0) {                                                                               \
      const double2* rp = reinterpret_cast<const double2*>(recs + (unsigned)SLOT * (EGG_REC * 8)); \
      _Pragma("unroll") for (int p = 0; p < EGG_PIECES; p++) BUF[p] = __ldg(rp + p);               \
      if (with_lam) {                                                                              \
        const double* lq = LAMB(rd) + 3 * (unsigned)SLOT;                                          \
        L0 = __ldcg(lq); L1 = __ldcg(lq + 1); L2 = __ldcg(lq + 2);                                 \
      }                                                                                            \
    }                                                                                              \
  }

    // HBM -> L2 prefetch of the records of stage t_pf (the register fetch one stage ahead then
    // finds them in L2: one stage of compute does not cover an HBM round trip).  The cursor wraps
    // to stage 0 of the next sweep, which streams the same records again.
#define PF2()                                                                                  \
  if (active && ns > 0) {                                                                      \
    const int pcnt = stage_cnt(t_pf);                                                          \
    if (sl < pcnt) {                                                                           \
      const char* pa = recs + (unsigned)(s0_pf + sl) * (EGG_REC * 8);                          \
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pa));                                      \
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 128));                                \
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pa + 232));                                \
    }                                                                                          \
    s0_pf += pcnt;                                                                             \
    if (++t_pf >= ns) { t_pf = 0; s0_pf = 0; }                                                 \
  }

    // one block update / residual evaluation on record v with multipliers (x0,x1,x2)
    double2 mvr[MV == 2 ? 10 : 1];   // M^-1 of body i0 (5 pieces) then body i1 (5 pieces), MV == 2
    auto step = [&](const double2* v, double x0, double x1, double x2, int slot) {
      if (slot < 0) return;
      int i0 = __double2loint(IDX), i1 = __double2hiint(IDX);
      i0 = (i0 < 0) ? n : i0;
      i1 = (i1 < 0) ? n : i1;
      double2* q1 = reinterpret_cast<double2*>(sb + i1 * BS);
      double2* q0 = reinterpret_cast<double2*>(sb + i0 * BS);
      if (pass_kind == 0) {
        x0 = RH0; x1 = RH1; x2 = RH2;
      }
      double2 a1a = q1[0], a1b = q1[1], a1c = q1[2];
      double2 a0a = q0[0], a0b = q0[1], a0c = q0[2];
      const bool eq = __double2loint(MET) < nj;
      if (pass_kind == 3) {                                  // residual of sweep k against the frozen a_k
        V3 tp = rel_t(v, q1[3], q1[4], q1[5], q0[3], q0[4], q0[5]);
        const double w0 = tp.x + cfm * x0 - RH0, w1 = tp.y + cfm * x1 - RH1, w2 = tp.z + cfm * x2 - RH2;
        const double q0s = w0 * w0, q1s = w1 * w1, q2s = w2 * w2;
        if (eq) se += q0s + q1s + q2s;
        else {
          s1 += ((x0 == -1.0 && w0 < 0) ? q0s : 0.0) + ((x1 == -1.0 && w1 < 0) ? q1s : 0.0) + ((x2 == 0.0 && w2 < 0) ? q2s : 0.0);
          s2 += ((x0 == 1.0 && w0 > 0) ? q0s : 0.0) + ((x1 == 1.0 && w1 > 0) ? q1s : 0.0);
          s3 += ((x0 > -1.0 && x0 < 1.0) ? q0s : 0.0) + ((x1 > -1.0 && x1 < 1.0) ? q1s : 0.0) + ((x2 > 0.0) ? q2s : 0.0);
        }
      }
      V3 t = rel_t(v, a1a, a1b, a1c, a0a, a0b, a0c);
      if (pass_kind == 2) {
        const double w0 = t.x + cfm * x0 - RH0, w1 = t.y + cfm * x1 - RH1, w2 = t.z + cfm * x2 - RH2;
        const double q0s = w0 * w0, q1s = w1 * w1, q2s = w2 * w2;
        if (eq) se += q0s + q1s + q2s;
        else {
          s1 += ((x0 == -1.0 && w0 < 0) ? q0s : 0.0) + ((x1 == -1.0 && w1 < 0) ? q1s : 0.0) + ((x2 == 0.0 && w2 < 0) ? q2s : 0.0);
          s2 += ((x0 == 1.0 && w0 > 0) ? q0s : 0.0) + ((x1 == 1.0 && w1 > 0) ? q1s : 0.0);
          s3 += ((x0 > -1.0 && x0 < 1.0) ? q0s : 0.0) + ((x1 > -1.0 && x1 < 1.0) ? q1s : 0.0) + ((x2 > 0.0) ? q2s : 0.0);
        }
        return;
      }
      double d0, d1, d2;
      if (pass_kind == 0) {
        d0 = x0; d1 = x1; d2 = x2;                            // a = M^-1 J^T x0
        double* lp = lam0 + 3 * (unsigned)slot;
        lp[0] = x0; lp[1] = x1; lp[2] = x2;
      } else {
        // clamp kind of this block (q2 shift already folded in by the assembly kernel)
        const bool contact = __double2hiint(MET) == KIND_CONTACT;
        const double lo01 = contact ? -1.0 : -kInf, hi01 = contact ? 1.0 : kInf, lo2 = contact ? 0.0 : -kInf;
        double n0 = (RH0 - t.x + DD0 * x0) * IA0;
        n0 = fmin(fmax(n0, lo01), hi01);
        d0 = n0 - x0;
        double n1 = (RH1 - (t.y + DO0 * d0) + DD1 * x1) * IA1;
        n1 = fmin(fmax(n1, lo01), hi01);
        d1 = n1 - x1;
        double n2 = (RH2 - (t.z + DO1 * d0 + DO2 * d1) + DD2 * x2) * IA2;
        n2 = fmax(n2, lo2);
        d2 = n2 - x2;
        double* lp = LAMB(rd ^ 1) + 3 * (unsigned)slot;
        lp[0] = n0; lp[1] = n1; lp[2] = n2;
      }
      // impulse scatter: a += M^-1 J^T delta
      const double ix = RC0 * d0 + RC3 * d1 + RC6 * d2, iy = RC1 * d0 + RC4 * d1 + RC7 * d2, iz = RC2 * d0 + RC5 * d1 + RC8 * d2;
      {
        double m[10];
        if (MV == 2) { _Pragma("unroll") for (int p = 0; p < 5; p++) { m[2 * p] = mvr[(MV == 2 ? 5 : 0) + (MV == 2 ? p : 0)].x; m[2 * p + 1] = mvr[(MV == 2 ? 5 : 0) + (MV == 2 ? p : 0)].y; } }
        else if (MS) { const double2* mq = q1 + 6; _Pragma("unroll") for (int p = 0; p < 5; p++) { double2 tq = mq[p]; m[2 * p] = tq.x; m[2 * p + 1] = tq.y; } }
        else { const double2* mq = reinterpret_cast<const double2*>(maos + i1 * 10); _Pragma("unroll") for (int p = 0; p < 5; p++) { double2 tq = __ldg(mq + p); m[2 * p] = tq.x; m[2 * p + 1] = tq.y; } }
        const double cx = R1Y * iz - R1Z * iy, cy = R1Z * ix - R1X * iz, cz = R1X * iy - R1Y * ix;   // r1 x imp
        const double dax = m[1] * cx + m[2] * cy + m[3] * cz, day = m[4] * cx + m[5] * cy + m[6] * cz, daz = m[7] * cx + m[8] * cy + m[9] * cz;
        a1a.x += m[0] * ix; a1a.y += m[0] * iy; a1b.x += m[0] * iz;
        a1b.y += dax; a1c.x += day; a1c.y += daz;
        q1[0] = a1a; q1[1] = a1b; q1[2] = a1c;
      }
      {
        double m[10];
        if (MV == 2) { _Pragma("unroll") for (int p = 0; p < 5; p++) { m[2 * p] = mvr[MV == 2 ? p : 0].x; m[2 * p + 1] = mvr[MV == 2 ? p : 0].y; } }
        else if (MS) { const double2* mq = q0 + 6; _Pragma("unroll") for (int p = 0; p < 5; p++) { double2 tq = mq[p]; m[2 * p] = tq.x; m[2 * p + 1] = tq.y; } }
        else { const double2* mq = reinterpret_cast<const double2*>(maos + i0 * 10); _Pragma("unroll") for (int p = 0; p < 5; p++) { double2 tq = __ldg(mq + p); m[2 * p] = tq.x; m[2 * p + 1] = tq.y; } }
        const double cx = R0Y * iz - R0Z * iy, cy = R0Z * ix - R0X * iz, cz = R0X * iy - R0Y * ix;   // r0 x imp
        const double dax = m[1] * cx + m[2] * cy + m[3] * cz, day = m[4] * cx + m[5] * cy + m[6] * cz, daz = m[7] * cx + m[8] * cy + m[9] * cz;
        a0a.x -= m[0] * ix; a0a.y -= m[0] * iy; a0b.x -= m[0] * iz;
        a0b.y -= dax; a0c.x -= day; a0c.y -= daz;
        q0[0] = a0a; q0[1] = a0b; q0[2] = a0c;
      }
    };

    if (__any_sync(0xffffffffu, active)) {
      if (STAGED) {
        // ---- records staged through shared memory by coalesced cp.async ----
        // The scattered per-lane LDG.128 of the register-prefetch path costs one L1TEX wavefront
        // per 128-byte line touched (16 lines x 15 loads per stage: the per-SM L1TEX FIFO was
        // the limiter, profiles/r1c).  Here the lanes of a world copy the stage's contiguous
        // chunk (cnt x 240 B) with consecutive 16-byte cp.async, then each lane picks its record
        // out of shared memory (240-byte stride is conflict-free for 128-bit accesses).
        int c_s0 = 0, c_cnt = 0;                 // chunk being copied: first slot, blocks
        double n0l = 0, n1l = 0, n2l = 0;        // multipliers of the chunk being copied
        auto chunk_of = [&](int kind, int t) {   // sets (c_s0, c_cnt) for step t; call in order
          if (!active) { c_cnt = 0; return; }
          if (kind == 2) { c_s0 = t * LPW; c_cnt = min(LPW, nc - c_s0); if (c_cnt < 0) c_cnt = 0; return; }
          if (t >= ns) { c_cnt = 0; return; }
          c_s0 = s0_next; c_cnt = stage_cnt(t); s0_next += c_cnt;
        };
        auto issue_copy = [&](bool with_lam) {
          const unsigned pieces = (unsigned)c_cnt * EGG_PIECES;
          const char* src = recs + (unsigned)c_s0 * (EGG_REC * 8);
          const unsigned dst = (unsigned)__cvta_generic_to_shared(stage);
          for (unsigned p = sl; p < pieces; p += LPW)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + p * 16), "l"(src + p * 16) : "memory");
          if (MV == 2) {
            const unsigned mpieces = (unsigned)c_cnt * 10;
            const char* msrc = reinterpret_cast<const char*>(d.rec_minv + ((size_t)wc * d.nrec + c_s0) * 20);
            for (unsigned p = sl; p < mpieces; p += LPW)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + STG_REC + p * 16), "l"(msrc + p * 16) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (with_lam && sl < c_cnt) {
            const double* lq = LAMB(rd) + 3 * (unsigned)(c_s0 + sl);
            n0l = __ldcg(lq); n1l = __ldcg(lq + 1); n2l = __ldcg(lq + 2);
          }
        };
        s0_next = 0;
        chunk_of(0, 0);
        issue_copy(false);
        while (true) {
          const int nsteps = (pass_kind == 2) ? nchunk : ns_max;
          se = s1 = s2 = s3 = 0.0;
          if (pass_kind == 3) {   // freeze a_k (aprev <- a) for the residual of sweep k
            for (int b = sl; active && b < n; b += LPW) {
              double2* q = reinterpret_cast<double2*>(sb + b * BS);
              q[3] = q[0]; q[4] = q[1]; q[5] = q[2];
            }
          }
          const bool wl = pass_kind != 0;
          for (int t = 0; t < nsteps; t++) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncwarp();
            const int slot = (sl < c_cnt) ? c_s0 + sl : -1;
            const double x0 = n0l, x1 = n1l, x2 = n2l;
            if (slot >= 0) {
              const double2* sp = reinterpret_cast<const double2*>(stage) + sl * EGG_PIECES;
#pragma unroll
              for (int p = 0; p < EGG_PIECES; p++) bufA[p] = sp[p];
              if (MV == 2) {
                const double2* mp = reinterpret_cast<const double2*>(stage + STG_REC) + sl * 10;
#pragma unroll
                for (int p = 0; p < 10; p++) mvr[MV == 2 ? p : 0] = mp[p];
              }
            }
            __syncwarp();                          // staging buffer free again
            if (t + 1 < nsteps) { chunk_of(pass_kind, t + 1); issue_copy(wl); }
            step(bufA, x0, x1, x2, slot);
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
              if (pass_kind == 2) active = false;                  // k == k_max: x_k and a are final
              else if (err > tol) { ++it; rd ^= 1; }               // accept x_{k+1}
              else { active = false; use_prev = true; }            // x_k had converged: drop x_{k+1}
            }
            if (!__any_sync(0xffffffffu, active)) break;
            ++k;
            pass_kind = (k >= k_max) ? 2 : 3;
          }
          s0_next = 0;
          chunk_of(pass_kind, 0);
          issue_copy(true);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
      } else {
      s0_next = 0;
      FETCH(bufA, lA0, lA1, lA2, slotA, next_slot(0, 0), false)
      while (true) {
        const int nsteps = (pass_kind == 2) ? nchunk : ns_max;
        se = s1 = s2 = s3 = 0.0;
        if (pass_kind == 3) {   // freeze a_k (aprev <- a) for the residual of sweep k
          for (int b = sl; active && b < n; b += LPW) {
            double2* q = reinterpret_cast<double2*>(sb + b * BS);
            q[3] = q[0]; q[4] = q[1]; q[5] = q[2];
          }
          __syncwarp();
        }
        const bool wl = pass_kind != 0;
        if (pass_kind != 2 && active) {
          // cursor at stage 2 (mod ns): stages 0 and 1 are fetched directly below
          t_pf = 0; s0_pf = 0;
          for (int q = 0; q < 2 && ns > 0; q++) { s0_pf += stage_cnt(t_pf); if (++t_pf >= ns) { t_pf = 0; s0_pf = 0; } }
        }
        int t = 0;
        while (t < nsteps) {
          // step t on A, prefetch t+1 into B (within a pass the read buffer is never written)
          if (t + 1 < nsteps) FETCH(bufB, lB0, lB1, lB2, slotB, next_slot(pass_kind, t + 1), wl)
          if (pass_kind != 2) PF2()
          step(bufA, lA0, lA1, lA2, slotA);
          __syncwarp();
          ++t;
          if (t >= nsteps) break;
          if (t + 1 < nsteps) FETCH(bufA, lA0, lA1, lA2, slotA, next_slot(pass_kind, t + 1), wl)
          if (pass_kind != 2) PF2()
          step(bufB, lB0, lB1, lB2, slotB);
          __syncwarp();
          ++t;
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
            if (pass_kind == 2) active = false;                  // k == k_max: x_k and a are final
            else if (err > tol) { ++it; rd ^= 1; }               // accept x_{k+1}
            else { active = false; use_prev = true; }            // x_k had converged: drop x_{k+1}
          }
          if (!__any_sync(0xffffffffu, active)) break;
          ++k;
          pass_kind = (k >= k_max) ? 2 : 3;
        }
        // first step of the next pass always starts in buffer A (one exposed fetch per pass)
        s0_next = 0;
        FETCH(bufA, lA0, lA1, lA2, slotA, next_slot(pass_kind, 0), true)
      }
      }
    }
    __syncwarp();

    if (valid) {
      const double* lamf = LAMB(rd);
      double* lo_out = d.lam_out + (size_t)w * 3 * d.nrec;
      int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
      for (int s = sl; s < nc; s += LPW) {
        const double meta = *reinterpret_cast<const double*>(recs + (size_t)s * (EGG_REC * 8) + REC_META * 8);
        const int orig = __double2loint(meta);
        const bool eq = orig < nj;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double x = __ldcg(lamf + 3 * (size_t)s + c);
          lo_out[3 * orig + c] = x;
          int state = 0;
          if (eq) state = 3;
          else if (x == ((c < 2) ? -1.0 : 0.0)) state = 1;
          else if (c < 2 && x == 1.0) state = 2;
          rs_out[3 * orig + c] = state;
        }
      }
      if (sl == 0) {
        int* stt = d.stats + (size_t)w * 8;
        stt[4] = it;
        stt[5] = 0;
        stt[6] = (cfm != 0.0);
        stt[7] = ns;
        d.resid[w] = err;
      }
      // v' = v + dt (M^-1 f + a); p += dt (v+v')/2; R <- WtoQ((w+w')/2, dt) R  (ensembles.cc:535,572-591)
      double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
      const int off = use_prev ? 6 : 0;
      bool bad = false;
      for (int b = sl; b < n; b += LPW) {
        const double* q = sb + b * BS + off;
        const double mi = __ldg(st + b);
        double Ii[9];
#pragma unroll
        for (int c = 0; c < 9; c++) Ii[c] = __ldg(st + (1 + c) * n + b);
        d3 fl = mk3(st[10 * n + b], st[11 * n + b], st[12 * n + b]);
        d3 ft = mk3(st[13 * n + b], st[14 * n + b], st[15 * n + b]);
        d3 v = mk3(dyn[12 * n + b], dyn[13 * n + b], dyn[14 * n + b]);
        d3 wv = mk3(dyn[15 * n + b], dyn[16 * n + b], dyn[17 * n + b]);
        d3 vn = v + dt * (fl * mi + mk3(q[0], q[1], q[2]));
        d3 wn = wv + dt * (mmulv(Ii, ft) + mk3(q[3], q[4], q[5]));
        d3 vmid = (v + vn) / 2.0, wmid = (wv + wn) / 2.0;
        d3 p = mk3(dyn[b], dyn[n + b], dyn[2 * n + b]) + dt * vmid;
        double z2 = dot3(wmid, wmid);
        d3 axis = (z2 > 0) ? wmid / sqrt(z2) : wmid;
        double ha = 0.5 * (norm3(wmid) * dt);
        double qw = cos(ha), sn = sin(ha);
        double qx = sn * axis.x, qy = sn * axis.y, qz = sn * axis.z;
        double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz, twx = tx * qw, twy = ty * qw, twz = tz * qw;
        double txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy, tyz = tz * qy, tzz = tz * qz;
        double Q[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
        double R[9], Rn[9];
#pragma unroll
        for (int c = 0; c < 9; c++) R[c] = dyn[(3 + c) * n + b];
        mmulm(Q, R, Rn);
        dyn[b] = p.x; dyn[n + b] = p.y; dyn[2 * n + b] = p.z;
#pragma unroll
        for (int c = 0; c < 9; c++) dyn[(3 + c) * n + b] = Rn[c];
        dyn[12 * n + b] = vn.x; dyn[13 * n + b] = vn.y; dyn[14 * n + b] = vn.z;
        dyn[15 * n + b] = wn.x; dyn[16 * n + b] = wn.y; dyn[17 * n + b] = wn.z;
        double chk = p.x + p.y + p.z + vn.x + vn.y + vn.z + wn.x + wn.y + wn.z;
        bad |= !(fabs(chk) < 1e300);
      }
      if (bad) atomicOr(&d.status[w], 16 /*EGG_ST_NONFINITE*/);
    }
    __syncwarp();
  }
}

int env_i(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

template <int LPW, int MV, bool STAGED>
void launch(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int G = 32 / LPW;
  int tabcap = d.nrec + 1;
  // stage sizes beyond the cap are read through L1; 128 entries keep six warps of 4 x 64-body
  // worlds resident (2042 vs 2524 ms per C3 step at 65536 worlds with a 512-entry table)
  const int tabmax = env_i("EGG_PGS_TABCAP", 128);
  if (tabcap > tabmax) tabcap = tabmax;
  tabcap = (tabcap + 15) & ~15;
  size_t smem = (size_t)G * (FS<MV>::value * (d.n + 1) * 8 + (STAGED ? LPW * EGG_REC * 8 + (MV == 2 ? LPW * 160 : 0) : 0)) + (size_t)G * tabcap;
  cudaFuncSetAttribute(egg_pgs_fast_kernel<LPW, MV, STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = env_i("EGG_PGS_CTAS_PER_SM", 12);
  int groups = (d.W + G - 1) / G;
  int grid = groups < sms * per_sm ? groups : sms * per_sm;
  egg_pgs_fast_kernel<LPW, MV, STAGED><<<grid, 32, smem, s>>>(d, dt, tabcap);
}

template <int LPW>
void launch2(const EggDev& d, double dt, cudaStream_t s) {
  // M^-1 source (EGG_PGS_MINV: 0 global, 1 shared body struct, 2 staged per block): the body
  // struct while four warps' worth still fits comfortably, else the per-block copy.
  constexpr int G = 32 / LPW;
  const size_t with_ms = (size_t)G * (FS<1>::value * (d.n + 1) * 8 + LPW * EGG_REC * 8);
  int mv = env_i("EGG_PGS_MINV", with_ms <= 28 * 1024 ? 1 : 0);   // 2 measured no faster than 0 on C3 (12.1 vs 12.2 ms)
  const bool staged = env_i("EGG_PGS_STAGED", 1) != 0;
  if (!d.rec_minv && mv == 2) mv = 0;
  if (!staged) { if (mv == 1) launch<LPW, 1, false>(d, dt, s); else launch<LPW, 0, false>(d, dt, s); return; }
  if (mv == 1) launch<LPW, 1, true>(d, dt, s);
  else if (mv == 2) launch<LPW, 2, true>(d, dt, s);
  else launch<LPW, 0, true>(d, dt, s);
}

}  // namespace

void egg_launch_solve_pgs_fast(const EggDev& d, double dt, int lpw, cudaStream_t s) {
  if (lpw == 4) launch2<4>(d, dt, s);
  else if (lpw == 16) launch2<16>(d, dt, s);
  else launch2<8>(d, dt, s);
}
