// Variant "stream" of the projected Gauss-Seidel solve (default).
//
// What the profiles of the "fast" variant said (profiles/r1d_pgs_fast_staged_summary.txt): the
// solve is latency-bound (FP64 pipe 13 %, issue 21 %, 1.5 warps per scheduler) and the number of
// worlds in flight per SM is capped by shared memory: 112 B per body (accumulator + the frozen
// copy the fused residual needs) + a 1920-byte staging buffer + a stage table, and by 227
// registers per thread.  This variant attacks exactly those:
//
//   * No frozen accumulator.  The reference evaluates the residual of x_k after every sweep only to
//     decide "stop / continue" (sparse_iterations.cc:196-215).  The residual is a sum of norms of
//     non-negative per-row terms (sparse_iterations.cc:51-69), so the same formula over ANY subset
//     of blocks is a lower bound.  Before each sweep one chunk of <= LPW blocks (the "probe") is
//     evaluated against a = a_k (nothing has been updated yet, so no copy is needed); if that
//     partial residual already exceeds 2 tol the decision "continue" is certain and the sweep
//     runs without any residual work.  Otherwise the full residual is evaluated exactly (one
//     read-only pass) and the reference's decision is taken on it; the probe then moves to the
//     chunk that contributed most.  Sweep counts, final multipliers and the reported residual
//     are those of the reference algorithm; 48 B per body instead of 112.
//   * The multipliers live inside the record (the slot that used to hold the D diagonal: the row
//     update is written in increment form  x' = x + (rhs - J a - cfm x) / (D + cfm),  so only
//     1/(D+cfm) is needed), are staged with it and written back in place.
//   * A stage's records (cnt x 240 B, contiguous) are staged by ONE cp.async.bulk per world
//     (TMA bulk copy, mbarrier complete_tx) instead of a ~40-instruction cp.async loop; the
//     count of the next stage rides in the spare slot of the stage's first record, so no stage
//     table is kept in shared memory.
//   * World groups are handed out by an atomic counter (worlds differ 2x in work).
//
// Per world: n x 48 B accumulator + LPW x 240 B staging.  64-body worlds: 4992 B -> 11 resident
// warps of 4 worlds per SM (was 6).
//
// Replaces: sparse::GaussSeidelIteration + GetResidualError + the velocity/position update, i.e.
// /root/reference/eggshell/sparse_iterations.cc:148-226,51-69,
// sparse_iterations_utils.cc:12-21,159-243,495-695, ensembles.cc:535,572-591.
#include "egg_internal.cuh"
#include <math_constants.h>
#include <cstdlib>

namespace {

#define kInf CUDART_INF
constexpr int RECB = EGG_REC * 8;   // 240

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// records stream through L2 once per sweep: evict-first, so that the small per-body arrays stay
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar, unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(bar), "l"(pol)
               : "memory");
}
__device__ __forceinline__ unsigned long long policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ unsigned long long policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ double2 ldg_keep(const double2* p, unsigned long long pol) {   // read-only, L2 evict-last
  double2 r;
  asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(r.x), "=d"(r.y) : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra LAB_DONE;\n\t"
      "bra LAB_WAIT;\n"
      "LAB_DONE:\n\t}" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// record pieces (double2 v[15]) -> fields; layout REC_* of egg_internal.cuh with the multipliers
// in the REC_DDIAG slot and the next stage's block count in the low word of slot 29
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
#define LM0 v[9].x
#define LM1 v[9].y
#define LM2 v[10].x
#define IA0 v[10].y
#define IA1 v[11].x
#define IA2 v[11].y
#define RH0 v[12].x
#define RH1 v[12].y
#define RH2 v[13].x
#define IDX v[13].y
#define MET v[14].x

enum { MODE_INIT = 0, MODE_UPDATE = 1, MODE_RESID = 2 };
constexpr int PF_SPAN = 11;   // log2 bytes of an L2 prefetch span (0 = off)
enum { PH_INIT = 0, PH_PROBE = 1, PH_EXACT = 2, PH_UPDATE = 3 };

// ISO >= 1: every body's M^-1 is (1/m) I3, (1/c) I3 exactly (egg_init snaps numerically isotropic
// inverse inertias, see egg_solve.cu) and comes as one 16-byte load per body.  ISO == 2: all
// bodies of the batch share the same (1/m, 1/c) (every body of the reference is the same cube,
// body.h:91) and the pair is a kernel constant.
template <int LPW, int MINB, int ISO>
__global__ void __launch_bounds__(32, MINB) egg_pgs_stream_kernel(EggDev d, double dt, int pf_span, int dbg) {
  constexpr int G = 32 / LPW;
  extern __shared__ __align__(128) unsigned char smraw[];
  const int n = d.n, lane = threadIdx.x, sub = lane / LPW, sl = lane % LPW;
  // shared memory: [mbarrier 16 B][dummy body 48 B: the ground / world anchor, always zero]
  //                [G x n x 6 doubles accumulators][G x LPW x 240 B staging]
  const unsigned bar = s32(smraw);
  double* sb = reinterpret_cast<double*>(smraw + 64) + (size_t)sub * 6 * n;
  const int dummy = -(sub * n) - 1;            // body index of the dummy relative to this world's sb
  unsigned char* stage = smraw + 64 + (size_t)G * 48 * n + (size_t)sub * LPW * RECB;
  const unsigned stage_s = s32(stage);
  const double cfm = d.prm.cfm;
  const int nj = d.nj;
  const unsigned FULL = 0xffffffffu;

  if (lane == 0) {
    mbar_init(bar, G);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (lane < 6) reinterpret_cast<double*>(smraw + 16)[lane] = 0.0;
  const unsigned long long pol = policy_evict_first();
  const unsigned long long pol_keep = policy_evict_last();
  double um = 0.0, uc = 0.0;                   // ISO == 2: the batch-wide (1/m, 1/c)
  if (ISO == 2) { um = d.minv_iso[0]; uc = d.minv_iso[1]; }
  __syncwarp();
  unsigned parity = 0;
  const int ngroups = (d.W + G - 1) / G;

  while (true) {
    int grp = 0;
    if (lane == 0) grp = atomicAdd(d.work_ctr, 1);
    grp = __shfl_sync(FULL, grp, 0);
    if (grp >= ngroups) break;
    const int w = grp * G + sub;
    const bool valid = w < d.W;
    const int wc = valid ? w : d.W - 1;
    const double* maos = ISO ? d.minv_iso + (size_t)wc * (n + 1) * 2      // [n+1][2], row n = 0 (ground / world anchor)
                             : d.minv_aos + (size_t)wc * (n + 1) * 10;    // [n+1][10]
    const int nc = valid ? nj + d.c_count[wc] : 0;
    const int ns = valid ? d.n_levels[wc] : 0;
    char* recs = reinterpret_cast<char*>(d.rec + (size_t)wc * d.nrec * EGG_REC);
    int cnt0 = 0;                                                 // blocks in stage 0
    if (ns > 0) { const int* gls = d.level_start + (size_t)wc * (d.nrec + 1); cnt0 = gls[1] - gls[0]; }
    for (int i = sl; i < 6 * n; i += LPW) sb[i] = 0.0;

    int ns_max = ns, nc_max = nc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ns_max = max(ns_max, __shfl_xor_sync(FULL, ns_max, o));
      nc_max = max(nc_max, __shfl_xor_sync(FULL, nc_max, o));
    }
    const int nchunk_max = (nc_max + LPW - 1) / LPW;

    bool active = nc > 0;
    double err = 0.0;
    int it = 0;
    int probe_s0 = 0, probe_cnt = cnt0;        // the chunk whose residual is the cheap lower bound
    double se = 0, s1 = 0, s2 = 0, s3 = 0;
    double best = -1.0;                        // exact pass: largest per-block contribution seen by this lane
    int best_s0 = 0;

    // chunk in flight (being copied / just landed) for this world
    int c_s0 = 0, c_cnt = 0;
    auto issue = [&](int s0, int cnt) {        // one round: every world leader arrives exactly once
      c_s0 = s0; c_cnt = cnt;
      if (sl == 0) {
        if (cnt > 0) {
          const unsigned bytes = (unsigned)cnt * RECB;
          mbar_arrive_tx(bar, bytes);
          bulk_g2s(stage_s, recs + (size_t)((dbg & 2) ? 0 : s0) * RECB, bytes, bar, pol);
        } else {
          mbar_arrive(bar);
        }
      }
    };

    // One block on record v.  MODE_INIT: a = M^-1 J^T x0 (x0 = rhs, already in the record);
    // MODE_UPDATE: projected row-by-row update + impulse scatter; MODE_RESID: residual terms only.
    auto step = [&](const double2* v, int mode, int slot, int chunk_s0) {
      if (slot < 0) return;
      const int i0 = __double2loint(IDX), i1 = __double2hiint(IDX);
      const int j0 = (i0 < 0) ? n : i0, j1 = (i1 < 0) ? n : i1;
      // accumulator pieces: (l.x,l.y) (l.z,a.x) (a.y,a.z); the ground / world side reads the
      // all-zero dummy body (never written: its stores are predicated off below)
      double2* q1 = reinterpret_cast<double2*>(sb + (i1 < 0 ? dummy : i1) * 6);
      double2* q0 = reinterpret_cast<double2*>(sb + (i0 < 0 ? dummy : i0) * 6);
      double2 a1a = q1[0], a1b = q1[1], a1c = q1[2];
      double2 a0a = q0[0], a0b = q0[1], a0c = q0[2];
      constexpr int MN = ISO ? 2 : 10;
      double m1[MN], m0[MN];
      if (ISO == 2) {
        m1[0] = m0[0] = um; m1[1] = m0[1] = uc;
      } else if (mode != MODE_RESID) {          // M^-1 of both bodies: issued early, used in the scatter
        const double2* mq1 = reinterpret_cast<const double2*>(maos + j1 * MN);
        const double2* mq0 = reinterpret_cast<const double2*>(maos + j0 * MN);
#pragma unroll
        for (int p = 0; p < MN / 2; p++) {
          const double2 t1 = ldg_keep(mq1 + p, pol_keep), t0 = ldg_keep(mq0 + p, pol_keep);
          m1[2 * p] = t1.x; m1[2 * p + 1] = t1.y; m0[2 * p] = t0.x; m0[2 * p + 1] = t0.y;
        }
      }
      const double x0 = LM0, x1 = LM1, x2 = LM2;
      double d0, d1, d2;
      if (mode == MODE_INIT) {
        d0 = x0; d1 = x1; d2 = x2;
      } else {
        // t = J a = Rc ((l1 + a1 x r1) - (l0 + a0 x r0))
        const double u1x = a1a.x + (a1c.x * R1Z - a1c.y * R1Y), u1y = a1a.y + (a1c.y * R1X - a1b.y * R1Z), u1z = a1b.x + (a1b.y * R1Y - a1c.x * R1X);
        const double u0x = a0a.x + (a0c.x * R0Z - a0c.y * R0Y), u0y = a0a.y + (a0c.y * R0X - a0b.y * R0Z), u0z = a0b.x + (a0b.y * R0Y - a0c.x * R0X);
        const double ux = u1x - u0x, uy = u1y - u0y, uz = u1z - u0z;
        const double tx = RC0 * ux + RC1 * uy + RC2 * uz, ty = RC3 * ux + RC4 * uy + RC5 * uz, tz = RC6 * ux + RC7 * uy + RC8 * uz;
        // e = rhs - (J a + cfm x) = -(A x - b) of the block's rows
        const double e0 = (RH0 - cfm * x0) - tx, e1 = (RH1 - cfm * x1) - ty, e2 = (RH2 - cfm * x2) - tz;
        if (mode == MODE_RESID) {
          const bool eq = __double2loint(MET) < nj;
          const double q0s = e0 * e0, q1s = e1 * e1, q2s = e2 * e2;
          double c;
          if (eq) { c = q0s + q1s + q2s; se += c; }
          else {
            // w = -e: at lo with w < 0, at hi with w > 0, strictly inside (sparse_iterations.cc:51-69)
            const double c1 = ((x0 == -1.0 && e0 > 0) ? q0s : 0.0) + ((x1 == -1.0 && e1 > 0) ? q1s : 0.0) + ((x2 == 0.0 && e2 > 0) ? q2s : 0.0);
            const double c2 = ((x0 == 1.0 && e0 < 0) ? q0s : 0.0) + ((x1 == 1.0 && e1 < 0) ? q1s : 0.0);
            const double c3 = ((x0 > -1.0 && x0 < 1.0) ? q0s : 0.0) + ((x1 > -1.0 && x1 < 1.0) ? q1s : 0.0) + ((x2 > 0.0) ? q2s : 0.0);
            s1 += c1; s2 += c2; s3 += c3;
            c = c1 + c2 + c3;
          }
          if (c > best) { best = c; best_s0 = chunk_s0; }
          return;
        }
        // clamp kind of this block (q2 shift already folded in by the assembly kernel)
        const bool contact = __double2hiint(MET) == KIND_CONTACT;
        const double lo01 = contact ? -1.0 : -kInf, hi01 = contact ? 1.0 : kInf, lo2 = contact ? 0.0 : -kInf;
        double n0 = x0 + e0 * IA0;
        n0 = fmin(fmax(n0, lo01), hi01);
        d0 = n0 - x0;
        double n1 = x1 + (e1 - DO0 * d0) * IA1;
        n1 = fmin(fmax(n1, lo01), hi01);
        d1 = n1 - x1;
        double n2 = x2 + ((e2 - DO1 * d0) - DO2 * d1) * IA2;
        n2 = fmax(n2, lo2);
        d2 = n2 - x2;
        double* lp = reinterpret_cast<double*>(recs + (size_t)slot * RECB) + REC_DDIAG;
        if (!(dbg & 1)) { lp[0] = n0; lp[1] = n1; lp[2] = n2; }
      }
      // impulse scatter: a += M^-1 J^T delta
      const double ix = RC0 * d0 + RC3 * d1 + RC6 * d2, iy = RC1 * d0 + RC4 * d1 + RC7 * d2, iz = RC2 * d0 + RC5 * d1 + RC8 * d2;
      {
        const double cx = R1Y * iz - R1Z * iy, cy = R1Z * ix - R1X * iz, cz = R1X * iy - R1Y * ix;   // r1 x imp
        double dax, day, daz;
        if (ISO) { dax = m1[1] * cx; day = m1[1] * cy; daz = m1[1] * cz; }
        else { dax = m1[1] * cx + m1[2] * cy + m1[3] * cz; day = m1[4] * cx + m1[5] * cy + m1[6] * cz; daz = m1[7] * cx + m1[8] * cy + m1[9] * cz; }
        a1a.x += m1[0] * ix; a1a.y += m1[0] * iy; a1b.x += m1[0] * iz;
        a1b.y += dax; a1c.x += day; a1c.y += daz;
        if (i1 >= 0) { q1[0] = a1a; q1[1] = a1b; q1[2] = a1c; }
      }
      {
        const double cx = R0Y * iz - R0Z * iy, cy = R0Z * ix - R0X * iz, cz = R0X * iy - R0Y * ix;   // r0 x imp
        double dax, day, daz;
        if (ISO) { dax = m0[1] * cx; day = m0[1] * cy; daz = m0[1] * cz; }
        else { dax = m0[1] * cx + m0[2] * cy + m0[3] * cz; day = m0[4] * cx + m0[5] * cy + m0[6] * cz; daz = m0[7] * cx + m0[8] * cy + m0[9] * cz; }
        a0a.x -= m0[0] * ix; a0a.y -= m0[0] * iy; a0b.x -= m0[0] * iz;
        a0b.y -= dax; a0c.x -= day; a0c.y -= daz;
        if (i0 >= 0) { q0[0] = a0a; q0[1] = a0b; q0[2] = a0c; }
      }
    };
    auto reduce4 = [&]() {
#pragma unroll
      for (int o = LPW / 2; o > 0; o >>= 1) {
        se += __shfl_xor_sync(FULL, se, o);
        s1 += __shfl_xor_sync(FULL, s1, o);
        s2 += __shfl_xor_sync(FULL, s2, o);
        s3 += __shfl_xor_sync(FULL, s3, o);
      }
    };

    if (__any_sync(FULL, active)) {
      const double tol = d.prm.tol;
      const int k_max = d.prm.k_max;
      // One loop runs every phase (a single copy of the block code):
      //   PH_INIT   x0 = rhs scattered into the accumulator, stage by stage
      //   PH_PROBE  residual terms of one chunk: the cheap lower bound of the residual of x_k
      //   PH_EXACT  residual of x_k over consecutive LPW-slices of all blocks (read-only)
      //   PH_UPDATE sweep k -> k+1, stage by stage (next count from the staged record)
      // The first chunk of a phase is in flight when the phase starts; nothing is in flight when
      // PH_EXACT / PH_UPDATE end.
      int phase = PH_INIT, k = 0;
      bool need_exact = false;
      issue(0, active ? cnt0 : 0);
      while (true) {
        const int mode = (phase == PH_INIT) ? MODE_INIT : (phase == PH_UPDATE ? MODE_UPDATE : MODE_RESID);
        const int nsteps = (phase == PH_PROBE) ? 1 : (phase == PH_EXACT ? nchunk_max : ns_max);
        const bool on = (phase == PH_EXACT) ? need_exact : active;
        const bool staged = (phase == PH_INIT || phase == PH_UPDATE);
        double2 buf[EGG_PIECES];
        for (int t = 0; t < nsteps; t++) {
          mbar_wait(bar, parity);
          parity ^= 1u;
          const int s0 = c_s0, cnt = c_cnt;
          const int slot = (sl < cnt) ? s0 + sl : -1;
          int nxt = 0;
          if (cnt > 0) nxt = *reinterpret_cast<const int*>(stage + 29 * 8);   // broadcast read of record 0's spare slot
          if (slot >= 0) {
            const double2* sp = reinterpret_cast<const double2*>(stage) + sl * EGG_PIECES;
#pragma unroll
            for (int p = 0; p < EGG_PIECES; p++) buf[p] = sp[p];
          }
          __syncwarp();                            // staging buffer free again
          if (t + 1 < nsteps) {
            if (staged) issue(s0 + cnt, (on && t + 1 < ns) ? nxt : 0);
            else { const int b0 = (t + 1) * LPW; issue(b0, on ? max(0, min(LPW, nc - b0)) : 0); }
          }
          if (cnt > 0 && phase != PH_PROBE && pf_span > 0) {
            // HBM -> L2 prefetch in spans of 2^pf_span bytes.  Random ~1 KB reads reach only about
            // half of the HBM bandwidth (tools/micro/dram_chunk_bench: 3.4 TB/s at 960 B, 5.4 at
            // 3840 B, 7.0 at 7680 B), so DRAM is asked for whole spans: when the consumer enters
            // span i the lanes of the world prefetch span i+2 (cyclically: the next sweep streams
            // the same records again); the 1 KB stage copies then hit L2.
            const int sp_now = ((s0 + cnt) * RECB - 1) >> pf_span;
            const int sp_prev = (s0 * RECB - 1) >> pf_span;          // -1 >> k = -1: the first chunk enters span 0
            if (sp_now != sp_prev) {
              const int nspan = ((nc * RECB - 1) >> pf_span) + 1;
              int tgt = sp_now + 2;
              if (tgt >= nspan) tgt -= nspan;
              if (tgt >= nspan) tgt = nspan - 1;
              const int lines = 1 << (pf_span - 7);
              const char* base = recs + ((size_t)tgt << pf_span);
              const int lim = nc * RECB - (tgt << pf_span);          // bytes left in the world's records
              for (int q = sl; q < lines && q * 128 < lim; q += LPW) prefetch_l2(base + q * 128);
            }
          }
          step(buf, mode, slot, s0);
          __syncwarp();                            // accumulator writes visible to the next stage
        }
        // ---- phase transitions (warp-uniform) ----
        if (phase == PH_PROBE) {
          // speculate "continue": stage 0 of the sweep streams in while the bound is reduced
          issue(0, active ? cnt0 : 0);
          reduce4();
          const double lb = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));
          need_exact = active && !(lb > 2.0 * tol);
          if (!__any_sync(FULL, need_exact)) { phase = PH_UPDATE; if (active) ++it; continue; }
          mbar_wait(bar, parity);                  // rare: drain the speculative copy, evaluate exactly
          parity ^= 1u;
          __syncwarp();
          phase = PH_EXACT;
          se = s1 = s2 = s3 = 0.0;
          best = -1.0;
          issue(0, need_exact ? min(LPW, nc) : 0);
          continue;
        }
        if (phase == PH_EXACT) {
          reduce4();
          if (need_exact) {
            err = sqrt(se) + (sqrt(s1) + sqrt(s2) + sqrt(s3));   // residual of x_k
            if (k >= k_max || !(err > tol)) active = false;      // x_k is final
          }
          // move the probe to the chunk with the largest contribution
          double bb = best;
          int bs = best_s0;
#pragma unroll
          for (int o = LPW / 2; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, bb, o);
            const int os = __shfl_xor_sync(FULL, bs, o);
            if (ob > bb || (ob == bb && os < bs)) { bb = ob; bs = os; }
          }
          if (need_exact && bb > 0.0) { probe_s0 = bs; probe_cnt = min(LPW, nc - bs); }
          if (!__any_sync(FULL, active)) break;
          phase = PH_UPDATE;
          if (active) ++it;
          issue(0, active ? cnt0 : 0);
          continue;
        }
        if (phase == PH_UPDATE) {
          // multipliers were written with generic stores; the async proxy reads them back next sweep
          asm volatile("fence.proxy.async;" ::: "memory");
          __syncwarp();
          ++k;
        }
        // after PH_INIT (k = 0) or PH_UPDATE: start the check of x_k
        se = s1 = s2 = s3 = 0.0;
        best = -1.0;
        if (k < k_max) {
          phase = PH_PROBE;
          issue(probe_s0, active ? probe_cnt : 0);
        } else {
          phase = PH_EXACT;
          need_exact = active;
          issue(0, need_exact ? min(LPW, nc) : 0);
        }
      }
    }
    __syncwarp();

    if (valid) {
      double* lo_out = d.lam_out + (size_t)w * 3 * d.nrec;
      int* rs_out = d.row_state + (size_t)w * 3 * d.nrec;
      for (int s = sl; s < nc; s += LPW) {
        const double* rp = reinterpret_cast<const double*>(recs + (size_t)s * RECB);
        const int orig = __double2loint(rp[REC_META]);
        const bool eq = orig < nj;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double x = __ldcg(rp + REC_DDIAG + c);
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
      const double* st = d.stat + (size_t)w * EGG_STAT * n;
      bool bad = false;
      for (int b = sl; b < n; b += LPW) {
        const double* q = sb + b * 6;
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

template <int LPW, int MINB, int ISO>
void launch(const EggDev& d, double dt, cudaStream_t s) {
  constexpr int G = 32 / LPW;
  const size_t smem = 64 + (size_t)G * (48 * d.n + LPW * RECB);
  cudaFuncSetAttribute(egg_pgs_stream_kernel<LPW, MINB, ISO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, egg_pgs_stream_kernel<LPW, MINB, ISO>, 32, smem);
  if (per_sm < 1) per_sm = 1;
  const int cap = env_i("EGG_PGS_CTAS_PER_SM", 0);
  if (cap > 0 && cap < per_sm) per_sm = cap;
  const int groups = (d.W + G - 1) / G;
  const int grid = groups < sms * per_sm ? groups : sms * per_sm;
  cudaMemsetAsync(d.work_ctr, 0, sizeof(int), s);
  egg_pgs_stream_kernel<LPW, MINB, ISO><<<grid, 32, smem, s>>>(d, dt, env_i("EGG_PGS_PF_SPAN", PF_SPAN), env_i("EGG_PGS_DBG", 0));
}

// LPW = lanes per world = maximum blocks per stage the assembly emitted.  Registers are allocated
// per scheduler (16384 each): 12 one-warp CTAs per SM = 3 per scheduler = 168 registers.
template <int MINB, int ISO>
void launch_lpw(const EggDev& d, double dt, int lpw, cudaStream_t s) {
  switch (lpw) {
    case 1: launch<1, MINB, ISO>(d, dt, s); break;
    case 2: launch<2, MINB, ISO>(d, dt, s); break;
    case 4: launch<4, MINB, ISO>(d, dt, s); break;
    case 16: launch<16, MINB, ISO>(d, dt, s); break;
    default: launch<8, MINB, ISO>(d, dt, s); break;
  }
}
}  // namespace

void egg_launch_solve_pgs_stream(const EggDev& d, double dt, int lpw, cudaStream_t s) {
  if (d.iso == 2) launch_lpw<12, 2>(d, dt, lpw, s);
  else if (d.iso == 1) launch_lpw<12, 1>(d, dt, lpw, s);
  else launch_lpw<8, 0>(d, dt, lpw, s);
}
