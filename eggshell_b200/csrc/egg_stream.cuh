// Shared by the group-stream PGS kernels (egg_pgs_stream.cu, egg_pgs_runs.cu): record / round
// layout of the stream, the PTX wrappers for bulk copies and mbarriers.
#pragma once
#include "egg_internal.cuh"
#include <math_constants.h>

namespace {

#define kInf CUDART_INF
// Two record formats (layouts below): FP64 records of 22 doubles = 176 B = 11 x 16 B, and the opt-in
// precision = 32 format of 24 floats + the packed word = 112 B = 7 x 16 B (both odd multiples of
// 16 B: conflict-free 128-bit shared-memory reads at these strides).
constexpr int SREC = 22;            // doubles per FP64 stream record
constexpr int RECB64 = SREC * 8;    // 176
constexpr int RECB32 = 112;
constexpr int LAMB = 32;            // bytes of one block's multipliers (3 doubles + pad = one sector)
constexpr int BLKB_MAX = RECB64 + LAMB;   // stream bytes per block, FP64 records (allocation bound)

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

// Stream record (22 doubles): [0..5] rows 0 and 1 of Rc, [6..8] r0, [9..11] r1, [12..14] D
// off-diagonal, [15..17] 1/(D+cfm), [18..20] rhs, [21] spare.  Row 2 of the frame is NOT stored:
// Rc is a rotation (or -I for a joint), so row 2 = +-(row 0 x row 1), three multiply-subtracts off
// the dependent chain instead of 24 bytes per block pass.  The multipliers are NOT in the record
// either: a round keeps them in one contiguous array of 32-byte sectors in front of its records
// (x0 x1 x2 | packed word: i0+1 | (i1+1) << 10 | kind << 20 | original constraint index << 21), so
// that the write-back of a warp-stage is a run of consecutive full sectors.  (With the multipliers
// inside each record the scattered 32-byte stores alone cost 40 % of the stream:
// tools/micro/stream_bench measures 4.1 TB/s with them against 6.2 read-only and 5.6 with the
// compact array; and partial-sector stores additionally made L2 fetch every sector it merged:
// +8 GB reads per launch, profiles/r1f.)
// precision = 32 record (112 B): all 24 numbers (the full frame) as floats (96 B), then the packed
// word and 8 spare bytes.  The kernel widens them to double as it reads; multipliers, accumulators and all
// arithmetic stay FP64.
// Inside the kernel the 24 numbers of either format are fld[0..23]:
#define RC0 fld[0]
#define RC1 fld[1]
#define RC2 fld[2]
#define RC3 fld[3]
#define RC4 fld[4]
#define RC5 fld[5]
#define RC6 fld[6]
#define RC7 fld[7]
#define RC8 fld[8]
#define R0X fld[9]
#define R0Y fld[10]
#define R0Z fld[11]
#define R1X fld[12]
#define R1Y fld[13]
#define R1Z fld[14]
#define DO0 fld[15]
#define DO1 fld[16]
#define DO2 fld[17]
#define IA0 fld[18]
#define IA1 fld[19]
#define IA2 fld[20]
#define RH0 fld[21]
#define RH1 fld[22]
#define RH2 fld[23]

// Projection onto [lo, hi] with plain compare-selects: fmax / fmin on doubles compile to DSETP.MAX +
// NaN fix-up (6 instructions, 4 deep) in the middle of the dependent row chain; same result for
// every non-NaN input (a NaN stays a NaN and the world is flagged EGG_ST_NONFINITE by the integrate).
__device__ __forceinline__ double clamp_sel(double v, double lo, double hi) {
  v = (v < lo) ? lo : v;
  return (v > hi) ? hi : v;
}
__device__ __forceinline__ void st_sector(double* p, double a, double b, double c, double e) {   // one aligned 32-byte store
  asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(e) : "memory");
}


constexpr int HDRB = 64;      // round header bytes (keeps the multiplier sectors 32-byte aligned)
// bytes of a round with `total` blocks: header, multipliers, records, padded to a sector
__host__ __device__ inline unsigned round_bytes(int total, int blkb) { return (unsigned)(HDRB + blkb * total + 31) & ~31u; }
constexpr int HDR_NEXT = 40;  // byte 40: total blocks of the next round (cyclic); byte 41: of the one after (unused)

__host__ __device__ inline size_t group_stride_bytes(int nrec, int G) { return ((size_t)nrec * ((size_t)BLKB_MAX * G + HDRB + 32) + 255) & ~(size_t)255; }


// ---- run format (d.rmax > 1, egg_pgs_runs.cu) ------------------------------------------------
// A lane carries a RUN of up to RUN_MAX consecutive contacts of one manifold (same ordered body
// pair, same normal) through a stage.  Round t of a group =
//   [64-byte header][32-byte multiplier sector of every block, k-major][run records, 4 columns of
//   16 bytes][block records, 5 columns of 16 bytes, k-major]       (padded to a multiple of 32 B)
// k-major: first the k = 0 blocks of all runs in lane order, then the k = 1 blocks, ...; a column is
// the same 16-byte piece of all items, so that lanes read consecutive 16-byte words.
//   header: bytes 0..15 run length of every lane (4 bits each, lane = world slot * LPW + lane of the
//           world), u32 @16 bytes of this round, u32 @20 bytes of the next round (cyclic),
//           u16 @24 blocks, u16 @26 runs
//   multiplier sector: x0 x1 x2 | packed: reference constraint index, clamp kind << 20, equality << 21
//   run record (64 B): contact-frame quaternion w x y z | r1 - r0 (3) | packed: i0 + 1, (i1 + 1) << 10, joint << 20
//   block record (80 B): r0 (3) | 1 / (D + cfm) (3) | rhs (3) | spare
// Not stored: the frame matrix (rebuilt from the quaternion once per run), r1 (= r0 + the run's offset),
// the off-diagonal of the 3x3 block D (for isotropic bodies D = s I - c0 q0 q0^T - c1 q1 q1^T, q = Rc r).
constexpr int RUN_MAX = 2;
constexpr int RUNB = 64, RCOLS = 4;
constexpr int RBLKB = 80, BCOLS = 5;
constexpr int RH_BYTES = 16, RH_NEXT = 20, RH_TOTAL = 24, RH_NRUNS = 26;
__host__ __device__ inline unsigned runs_round_bytes(int total, int nruns) { return (unsigned)(HDRB + total * (LAMB + RBLKB) + nruns * RUNB + 31) & ~31u; }
// one-bit-per-nibble mask of the nibbles of x that are >= v (v = 1..4)
__host__ __device__ inline unsigned nib_ge(unsigned x, int v) {
  const unsigned b0 = x, b1 = x >> 1, b2 = x >> 2, b3 = x >> 3;
  unsigned m;
  if (v <= 1) m = b0 | b1 | b2 | b3;
  else if (v == 2) m = b1 | b2 | b3;
  else if (v == 3) m = (b0 & b1) | b2 | b3;
  else m = b2 | b3;
  return m & 0x11111111u;
}
// Fused integrate of body b of one world (dyn [18][n], stat [16][n]) with its accumulator
// a = M^-1 J^T x = (al, aa):  v' = v + dt (M^-1 f + a); p += dt (v+v')/2; R <- WtoQ((w+w')/2, dt) R
// (ensembles.cc:535,572-591).  Returns true when the new state is not finite.
__device__ __forceinline__ bool stream_integrate_body(double* dyn, const double* st, int n, int b, double alx, double aly, double alz, double aax, double aay, double aaz, double dt) {
  const double mi = __ldg(st + b);
  double Ii[9];
#pragma unroll
  for (int c = 0; c < 9; c++) Ii[c] = __ldg(st + (1 + c) * n + b);
  d3 fl = mk3(st[10 * n + b], st[11 * n + b], st[12 * n + b]);
  d3 ft = mk3(st[13 * n + b], st[14 * n + b], st[15 * n + b]);
  d3 v = mk3(dyn[12 * n + b], dyn[13 * n + b], dyn[14 * n + b]);
  d3 wv = mk3(dyn[15 * n + b], dyn[16 * n + b], dyn[17 * n + b]);
  d3 vn = v + dt * (fl * mi + mk3(alx, aly, alz));
  d3 wn = wv + dt * (mmulv(Ii, ft) + mk3(aax, aay, aaz));
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
  return !(fabs(chk) < 1e300);
}
// The same for world w from its accumulator sb [n][6]; `lanes` lanes of the warp starting at lane
// `sl` stride over the bodies.
__device__ __forceinline__ void stream_integrate_world(const EggDev& d, int w, const double* sb, int sl, int lanes, double dt) {
  const int n = d.n;
  double* dyn = d.dyn + (size_t)w * EGG_DYN * n;
  const double* st = d.stat + (size_t)w * EGG_STAT * n;
  bool bad = false;
  for (int b = sl; b < n; b += lanes) {
    const double* q = sb + b * 6;
    bad |= stream_integrate_body(dyn, st, n, b, q[0], q[1], q[2], q[3], q[4], q[5], dt);
  }
  if (bad) atomicOr(&d.status[w], 16 /*EGG_ST_NONFINITE*/);
}

}  // namespace
