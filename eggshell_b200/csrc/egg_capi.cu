// C ABI of libeggshell_b200.so (declared in include/egg_cuda.h).  Host-side plumbing only:
// device allocation, AoS<->SoA staging, kernel sequencing for Ensemble::Step
// (/root/reference/eggshell/ensembles.cc:390-427).  There is no CPU fallback: without a CUDA
// device every entry point fails with EGG_ERR_NO_DEVICE.
#include "../../include/egg_cuda.h"
#include "egg_internal.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>

cudaError_t egg_launch_solve_dense(const EggDev& d, double dt, cudaStream_t s, void* scratch, size_t scratch_bytes);
size_t egg_dense_scratch_bytes(const EggDev& d);
int egg_dense_smem_fits(const EggDev& d, size_t limit);
double egg_measure_fp64_tflops();
cudaError_t egg_launch_relax(const EggDev& d, double dt, double step_scale, int max_steps, int mode, void* scratch, size_t scratch_bytes,
                             int* prog, double* err2, int* any_active, cudaStream_t s);

static thread_local std::string g_err;
static void set_err(const char* what, cudaError_t e) {
  char buf[512];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  g_err = buf;
}
#define CK(call)                                \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) {                   \
      set_err(#call, e__);                      \
      return EGG_ERR_CUDA;                      \
    }                                           \
  } while (0)
// CK for egg_create after the batch exists: releases it on failure
#define CKB(call)                               \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) {                   \
      set_err(#call, e__);                      \
      egg_destroy(b);                           \
      return EGG_ERR_CUDA;                      \
    }                                           \
  } while (0)

struct egg_batch {
  egg_desc desc;
  EggDev dev;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  bool initialised = false;
  bool iso_known = false;     // dev.iso read back from the device after egg_init
  std::vector<void*> allocs;
  long long bytes = 0;
  long long launches = 0;
  double* stage = nullptr;        // device staging for AoS <-> SoA conversion
  size_t stage_bytes = 0;
  void* dense_scratch = nullptr;
  size_t dense_scratch_bytes = 0;
  int* relax_prog = nullptr;     // [2 W] relaxation progress, allocated by the first stabilisation call
  double* relax_err2 = nullptr;  // [W]
  int* relax_any = nullptr;      // [1]
  double* rec_aux = nullptr;     // per-world records for the relaxation when dev.rec holds the PGS group stream
  int* level_aux = nullptr;
  int device = 0;
  double* snap = nullptr;        // egg_snapshot copy of dev.dyn
  bool profiling = false;
  std::vector<cudaEvent_t> ev;   // 4 per profiled step
  std::vector<cudaEvent_t> ev_free;
};

// LK(wrapper call): a launch wrapper's cudaError_t -> EGG_ERR_CUDA with the wrapper named
#define LK(call)                                \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) {                   \
      set_err(#call, e__);                      \
      return EGG_ERR_CUDA;                      \
    }                                           \
  } while (0)

template <class T>
static int dalloc(egg_batch* b, T** p, size_t count) {
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, bytes);
  if (e != cudaSuccess) { set_err("cudaMalloc", e); return EGG_ERR_CUDA; }
  e = cudaMemsetAsync(q, 0, bytes, b->stream);
  if (e != cudaSuccess) { set_err("cudaMemset", e); return EGG_ERR_CUDA; }
  b->allocs.push_back(q);
  b->bytes += (long long)bytes;
  *p = (T*)q;
  return EGG_OK;
}
#define DA(ptr, count)                                  \
  do {                                                  \
    int r__ = dalloc(b, &(ptr), (size_t)(count));       \
    if (r__ != EGG_OK) { egg_destroy(b); return r__; }  \
  } while (0)

extern "C" {

const char* egg_last_error(void) { return g_err.c_str(); }
const char* egg_version(void) { return "eggshell_b200 0.1 (sm_100a)"; }

int egg_desc_default(egg_desc* d, int n_worlds, int n_bodies, int n_joints) {
  if (!d) return EGG_ERR_ARG;
  memset(d, 0, sizeof(*d));
  d->n_worlds = n_worlds;
  d->n_bodies = n_bodies;
  d->n_joints = n_joints;
  d->max_contacts = 0;
  d->precision = 64;
  d->solver = EGG_SOLVER_DENSE_MURTY;   // what the reference ships (ensembles.cc:21)
  d->k_max = 500;                       // sparse_iterations.cc:19
  d->tol = 1e-9;                        // constants.h:5
  d->cfm = 0.01;                        // ensembles.cc:14
  d->erp = 0.2;                         // ensembles.h:166
  d->gravity[0] = 0; d->gravity[1] = 0; d->gravity[2] = -9.8;   // constants.h:8
  d->min_constraint_dist = 1e-6;        // ensembles.cc:15
  d->quirks = EGG_QUIRKS_REFERENCE;
  d->cfm_mode = EGG_CFM_AUTO;
  d->device = 0;
  d->taps = 0;
  return EGG_OK;
}

int egg_create(const egg_desc* dsc, egg_batch** out) {
  if (!dsc || !out) return EGG_ERR_ARG;
  *out = nullptr;
  if (dsc->n_worlds <= 0 || dsc->n_bodies <= 0 || dsc->n_joints < 0 || dsc->n_bodies > 360) { g_err = "bad shape"; return EGG_ERR_ARG; }
  if (dsc->precision != 64 && dsc->precision != 32) { g_err = "precision must be 64 or 32"; return EGG_ERR_ARG; }
  if (dsc->precision == 32 && dsc->solver != EGG_SOLVER_PGS) { g_err = "precision = 32 (FP32 constraint records) exists for the PGS solver only"; return EGG_ERR_UNSUPPORTED; }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    g_err = "no CUDA device: eggshell_b200 has no CPU fallback";
    return EGG_ERR_NO_DEVICE;
  }
  if (dsc->device < 0 || dsc->device >= ndev) { g_err = "bad device ordinal"; return EGG_ERR_ARG; }
  CK(cudaSetDevice(dsc->device));
  egg_batch* b = new egg_batch();
  b->desc = *dsc;
  b->device = dsc->device;
  e = cudaStreamCreateWithFlags(&b->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { set_err("cudaStreamCreate", e); delete b; return EGG_ERR_CUDA; }
  b->own_stream = true;

  EggDev& d = b->dev;
  memset(&d, 0, sizeof(d));
  const int W = dsc->n_worlds, n = dsc->n_bodies, nj = dsc->n_joints;
  d.W = W; d.n = n; d.nj = nj;
  d.P = n * (n - 1) / 2;
  // Capacity: 8 ground contacts per body is exact (collision.cc:412-414); pairs emit <= 10 each
  // (4-gon clipped by 6 half-spaces).  Automatic = 8 n + 4 contacts for ~4 colliding pairs/body.
  int maxc = dsc->max_contacts > 0 ? dsc->max_contacts : 8 * n + 16 * n;
  if (maxc > 8 * n + 10 * d.P) maxc = 8 * n + 10 * d.P;
  if (maxc < 8) maxc = 8;
  maxc = (maxc + 7) & ~7;
  d.maxc = maxc;
  d.nrec = nj + maxc;
  b->desc.max_contacts = maxc;
  d.prm.erp = dsc->erp; d.prm.cfm = dsc->cfm; d.prm.tol = dsc->tol; d.prm.min_dist = dsc->min_constraint_dist;
  for (int k = 0; k < 3; k++) d.prm.g[k] = dsc->gravity[k];
  d.prm.k_max = dsc->k_max; d.prm.solver = dsc->solver; d.prm.quirks = dsc->quirks; d.prm.cfm_mode = dsc->cfm_mode;

  DA(d.dyn, (size_t)W * EGG_DYN * n);
  DA(d.stat, (size_t)W * EGG_STAT * n);
  DA(d.bpar, (size_t)W * EGG_BPAR * n);
  DA(d.minv_aos, (size_t)W * (n + 1) * 10);
  DA(d.j_i0, (size_t)W * nj);
  DA(d.j_i1, (size_t)W * nj);
  DA(d.jc, (size_t)W * 6 * nj);
  DA(d.c_count, W);
  DA(d.c_i0, (size_t)W * maxc);
  DA(d.c_i1, (size_t)W * maxc);
  DA(d.c_code, (size_t)W * maxc);
  DA(d.c_geom, (size_t)W * 7 * maxc);
  if (dsc->taps) {
    DA(d.pair_code, (size_t)W * d.P);
    DA(d.pair_cnt, (size_t)W * d.P);
  }
  // The PGS solver streams group-interleaved records (format 1, egg_pgs_stream.cu); the dense path,
  // Jacobi / SOR and the relaxation read per-world 240-byte records (format 0, egg_pgs.cu).
  d.rec_fmt = (dsc->solver == EGG_SOLVER_PGS) ? 1 : 0;
  d.rmax = 0;
  // the group-stream assembly keeps u16 level tables of 2 (n + 4 nrec) bytes per world in shared memory
  if (d.rec_fmt && d.nrec > 24000) { g_err = "PGS: more than 24000 constraint slots per world (n_joints + max_contacts) are not supported"; egg_destroy(b); return EGG_ERR_UNSUPPORTED; }
  if (d.rec_fmt) {
    // group stream: G = 32 / lpw worlds share one interleaved record stream
    d.lpw = egg_stage_cap(d);
    d.blkb = egg_stream_blkb(dsc->precision);
    const int groups = (W + 32 / d.lpw - 1) / (32 / d.lpw);
    DA(d.rec, egg_stream_rec_bytes(W, d.nrec, d.lpw) / sizeof(double));
    DA(d.c_pos, (size_t)W * d.nrec);
    DA(d.st_cnt, (size_t)W * d.nrec);
    // run format (egg_pgs_runs.cu): FP64 records, at most 8 lanes per world; whether the bodies are
    // isotropic (its other condition) is known after egg_init, so the choice is made at the first step
    const bool runs_asked = (dsc->quirks & EGG_OPT_PGS_RUNS) != 0 || (getenv("EGG_PGS_RUNS") && atoi(getenv("EGG_PGS_RUNS")) != 0);
    if (runs_asked && dsc->precision == 64 && d.lpw <= 8) {   // [W][nrec] words: only for batches that asked for the run format
      DA(d.st_runs, (size_t)W * d.nrec);
      d.run_cap = egg_run_cap(d);
    }
    DA(d.round_off, (size_t)groups * (d.nrec + 1));
    DA(d.grp_info, (size_t)groups * 4);
  } else {
    DA(d.rec, (size_t)W * d.nrec * EGG_REC);
  }
  if (dsc->solver == EGG_SOLVER_JACOBI || dsc->solver == EGG_SOLVER_SOR) {
    DA(d.lam, (size_t)W * d.nrec * 3);
    DA(d.lam2, (size_t)W * d.nrec * 3);
    DA(d.slot_of, (size_t)W * d.nrec);
  }
  DA(d.lam_out, (size_t)W * d.nrec * 3);
  DA(d.row_state, (size_t)W * d.nrec * 3);
  DA(d.n_levels, W);
  DA(d.status, W);
  DA(d.stats, (size_t)W * 8);
  DA(d.resid, W);
  DA(d.cost0, (size_t)W * 2);
  DA(d.work_ctr, 4);
  DA(d.dbg, 32);
  DA(d.minv_iso, (size_t)W * (n + 1) * 2);
  DA(d.iso_flag, 1);
  b->stage_bytes = (size_t)W * n * 9 * sizeof(double);
  size_t jb = (size_t)W * (nj > 0 ? nj : 1) * 3 * sizeof(double);
  if (jb > b->stage_bytes) b->stage_bytes = jb;
  {   // the contact taps are transposed through the same buffer: room for the geometry of up to 256 worlds at a time
    const size_t cb = (size_t)(W < 256 ? W : 256) * maxc * 3 * sizeof(double);
    if (cb > b->stage_bytes) b->stage_bytes = cb;
  }
  DA(b->stage, b->stage_bytes / sizeof(double));
  // Shared-memory needs of every kernel this batch can launch, against the device limit: an
  // unsupported shape fails here, not as a launch error inside the first egg_step.
  {
    int lim = 0;
    CKB(cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dsc->device));
    struct { const char* what; size_t need; } req[4] = {
        {"narrowphase (egg_collide_kernel)", egg_collide_smem(d)},
        {"row assembly", d.rec_fmt ? (size_t)0 : egg_assemble_smem(d)},
        {"PGS group stream", d.rec_fmt ? (d.st_runs ? std::max(egg_stream_smem(d), egg_runs_smem(d)) : egg_stream_smem(d)) : (size_t)0},
        {"Jacobi / SOR", (dsc->solver == EGG_SOLVER_JACOBI || dsc->solver == EGG_SOLVER_SOR) ? egg_iter_smem(d) : (size_t)0}};
    for (auto& r : req)
      if (r.need > (size_t)lim) {
        char buf[256];
        snprintf(buf, sizeof(buf), "%s needs %zu bytes of shared memory per CTA for n_bodies = %d, max_contacts = %d; the device offers %d", r.what, r.need, n, maxc, lim);
        g_err = buf;
        egg_destroy(b);
        return EGG_ERR_UNSUPPORTED;
      }
    if (dsc->solver == EGG_SOLVER_DENSE_MURTY && !egg_dense_smem_fits(d, (size_t)lim)) {
      g_err = "dense solver: the row capacity (EGG_DENSE_ROWS / n_joints + 4 n_bodies constraints) does not fit the shared memory of one CTA";
      egg_destroy(b);
      return EGG_ERR_UNSUPPORTED;
    }
  }
  if (dsc->solver == EGG_SOLVER_DENSE_MURTY) {
    DA(d.work, W);
    b->dense_scratch_bytes = egg_dense_scratch_bytes(d);
    char* p = nullptr;
    DA(p, b->dense_scratch_bytes);
    b->dense_scratch = p;
  }
  CKB(cudaStreamSynchronize(b->stream));
  *out = b;
  return EGG_OK;
}

void egg_destroy(egg_batch* b) {
  if (!b) return;
  cudaSetDevice(b->device);
  if (b->stream) cudaStreamSynchronize(b->stream);
  for (void* p : b->allocs) cudaFree(p);
  for (cudaEvent_t e : b->ev) cudaEventDestroy(e);
  for (cudaEvent_t e : b->ev_free) cudaEventDestroy(e);
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  delete b;
}

int egg_set_stream(egg_batch* b, void* cuda_stream) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  if (b->own_stream && b->stream) cudaStreamDestroy(b->stream);
  b->stream = (cudaStream_t)cuda_stream;
  b->own_stream = false;
  return EGG_OK;
}

int egg_sync(egg_batch* b) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  CK(cudaGetLastError());
  return EGG_OK;
}

int egg_set_profiling(egg_batch* b, int on) {
  if (!b) return EGG_ERR_ARG;
  b->profiling = on != 0;
  return EGG_OK;
}
int egg_get_kernel_ms(egg_batch* b, double* out4) {
  if (!b || !out4) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  CK(cudaStreamSynchronize(b->stream));
  for (int k = 0; k < 4; k++) out4[k] = 0;
  for (size_t s = 0; s + 3 < b->ev.size(); s += 4) {
    for (int k = 0; k < 3; k++) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, b->ev[s + k], b->ev[s + k + 1]));
      out4[k] += ms;
    }
    out4[3] += 1;
  }
  for (cudaEvent_t e : b->ev) b->ev_free.push_back(e);
  b->ev.clear();
  return EGG_OK;
}
double egg_fp64_peak_tflops(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device"; return (double)EGG_ERR_NO_DEVICE; }
  if (cudaSetDevice(device) != cudaSuccess) return (double)EGG_ERR_ARG;
  return egg_measure_fp64_tflops();
}
int egg_capacity(const egg_batch* b) { return b ? b->dev.maxc : 0; }
long long egg_device_bytes(const egg_batch* b) { return b ? b->bytes : 0; }
long long egg_launch_count(const egg_batch* b) { return b ? b->launches : 0; }

void* egg_host_alloc(long long bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void egg_host_free(void* p) { if (p) cudaFreeHost(p); }

// Host AoS [W][per_world][comps] -> device SoA component block.
static int upload(egg_batch* b, const double* host, int per_world, int comps, double* soa, int soa_comps, int comp_off) {
  size_t bytes = (size_t)b->dev.W * per_world * comps * sizeof(double);
  if (bytes == 0) return EGG_OK;
  CK(cudaMemcpyAsync(b->stage, host, bytes, cudaMemcpyHostToDevice, b->stream));
  LK(egg_launch_pack(b->dev.W, b->stage, per_world, comps, soa, soa_comps, comp_off, b->stream));
  b->launches++;
  return EGG_OK;
}
static int download(egg_batch* b, double* host, int per_world, int comps, const double* soa, int soa_comps, int comp_off) {
  size_t bytes = (size_t)b->dev.W * per_world * comps * sizeof(double);
  if (bytes == 0) return EGG_OK;
  LK(egg_launch_unpack(b->dev.W, b->stage, per_world, comps, soa, soa_comps, comp_off, b->stream));
  b->launches++;
  CK(cudaMemcpyAsync(host, b->stage, bytes, cudaMemcpyDeviceToHost, b->stream));
  // the staging buffer is reused by the next call
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}
#define RET(x) do { int r__ = (x); if (r__ != EGG_OK) return r__; } while (0)

int egg_set_state(egg_batch* b, const double* p, const double* R, const double* v, const double* w) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const int n = b->dev.n;
  if (p) RET(upload(b, p, n, 3, b->dev.dyn, EGG_DYN, 0));
  if (R) RET(upload(b, R, n, 9, b->dev.dyn, EGG_DYN, 3));
  if (v) RET(upload(b, v, n, 3, b->dev.dyn, EGG_DYN, 12));
  if (w) RET(upload(b, w, n, 3, b->dev.dyn, EGG_DYN, 15));
  return EGG_OK;
}

int egg_set_bodies(egg_batch* b, const double* p, const double* R, const double* v, const double* w,
                   const double* m, const double* I_body, const double* side) {
  if (!b || !p || !R || !v || !w || !m || !I_body) return EGG_ERR_ARG;
  RET(egg_set_state(b, p, R, v, w));
  const int n = b->dev.n;
  if (side) RET(upload(b, side, n, 3, b->dev.bpar, EGG_BPAR, 0));
  else {
    std::vector<double> s((size_t)b->dev.W * n * 3, 0.3);    // body.h:91
    RET(upload(b, s.data(), n, 3, b->dev.bpar, EGG_BPAR, 0));
    CK(cudaStreamSynchronize(b->stream));
  }
  RET(upload(b, m, n, 1, b->dev.bpar, EGG_BPAR, 3));
  RET(upload(b, I_body, n, 9, b->dev.bpar, EGG_BPAR, 4));
  b->initialised = false;
  return EGG_OK;
}

int egg_set_shapes(egg_batch* b, const int* shape, const double* dims) {
  if (!b || !shape || !dims) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const size_t cnt = (size_t)b->dev.W * b->dev.n;
  std::vector<double> t(cnt);
  for (size_t k = 0; k < cnt; k++) {
    if (shape[k] < 0 || shape[k] > 2) { g_err = "collider shape must be 0 (box), 1 (sphere) or 2 (capsule)"; return EGG_ERR_ARG; }
    t[k] = (double)shape[k];
  }
  RET(upload(b, t.data(), b->dev.n, 1, b->dev.bpar, EGG_BPAR, 13));
  CK(cudaStreamSynchronize(b->stream));      // t is a temporary
  RET(upload(b, dims, b->dev.n, 3, b->dev.bpar, EGG_BPAR, 0));
  return EGG_OK;
}

int egg_set_joints(egg_batch* b, const int* i0, const int* i1, const double* c0, const double* c1) {
  if (!b) return EGG_ERR_ARG;
  const int nj = b->dev.nj;
  if (nj == 0) return EGG_OK;
  if (!i0 || !i1 || !c0 || !c1) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const size_t cnt = (size_t)b->dev.W * nj;
  for (size_t k = 0; k < cnt; k++) {
    if (i0[k] < 0 || i0[k] >= b->dev.n || i1[k] < -1 || i1[k] >= b->dev.n) { g_err = "joint body index out of range"; return EGG_ERR_ARG; }
  }
  CK(cudaMemcpyAsync(b->dev.j_i0, i0, cnt * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  CK(cudaMemcpyAsync(b->dev.j_i1, i1, cnt * sizeof(int), cudaMemcpyHostToDevice, b->stream));
  RET(upload(b, c0, nj, 3, b->dev.jc, 6, 0));
  RET(upload(b, c1, nj, 3, b->dev.jc, 6, 3));
  b->initialised = false;
  return EGG_OK;
}

int egg_set_external(egg_batch* b, const double* f_ext) {
  if (!b || !f_ext) return EGG_ERR_ARG;
  if (!b->initialised) { g_err = "egg_set_external must follow egg_init"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  return upload(b, f_ext, b->dev.n, 6, b->dev.stat, EGG_STAT, 10);
}

int egg_init(egg_batch* b) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  LK(egg_launch_init(b->dev, b->stream));
  b->launches += (b->dev.nj > 0) ? 3 : 2;
  // Ensemble::Init ends with CheckAndCorrectEnsembleState (ensembles.cc:28): the joint-joint
  // conflict scan lives in the narrowphase kernel, so run it once here -- EGG_ST_JOINT_CONFLICT is
  // then visible right after egg_init, as the reference's Panic would be.
  LK(egg_launch_collide(b->dev, b->stream));
  LK(egg_launch_clear_contacts(b->dev, b->stream));   // the contact list stays empty until the first Step / UpdateContacts, as in the reference
  b->launches += 2;
  b->initialised = true;
  b->iso_known = false;
  return EGG_OK;
}

// Shared by egg_init_stabilize / egg_post_stabilize: scratch and progress buffers, allocated once.
static int relax_prepare(egg_batch* b) {
  const int W = b->dev.W;
  if (!b->dense_scratch) {
    if (!egg_dense_smem_fits(b->dev, 227 * 1024)) { g_err = "stabilisation: the row capacity does not fit the shared memory of one CTA"; return EGG_ERR_UNSUPPORTED; }
    b->dense_scratch_bytes = egg_dense_scratch_bytes(b->dev);
    char* p = nullptr;
    int r = dalloc(b, &p, b->dense_scratch_bytes);
    if (r != EGG_OK) return r;
    b->dense_scratch = p;
  }
  if (!b->relax_prog) {
    int r = dalloc(b, &b->relax_prog, (size_t)2 * W); if (r != EGG_OK) return r;
    r = dalloc(b, &b->relax_err2, (size_t)W); if (r != EGG_OK) return r;
    r = dalloc(b, &b->relax_any, (size_t)1); if (r != EGG_OK) return r;
  }
  CK(cudaMemsetAsync(b->relax_prog, 0, (size_t)2 * W * sizeof(int), b->stream));
  return EGG_OK;
}
static int relax_report(egg_batch* b, int* steps_out, double* err_sq_out) {
  const int W = b->dev.W;
  if (steps_out) {
    std::vector<int> h((size_t)2 * W);
    CK(cudaMemcpyAsync(h.data(), b->relax_prog, h.size() * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    for (int w = 0; w < W; w++) steps_out[w] = h[2 * w];
  }
  if (err_sq_out) CK(cudaMemcpyAsync(err_sq_out, b->relax_err2, (size_t)W * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  CK(cudaGetLastError());
  return EGG_OK;
}

int egg_init_stabilize(egg_batch* b, int max_steps, int* steps_out, double* err_sq_out) {
  if (!b || max_steps < 0) return EGG_ERR_ARG;
  if (!b->initialised) { g_err = "egg_init_stabilize before egg_init"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  RET(relax_prepare(b));
  EggDev nodedup = b->dev;
  nodedup.prm.min_dist = -1.0;   // UpdateContacts only: the loop of ensembles.cc:610-617 never de-duplicates
  nodedup.rec_fmt = 0;           // the relaxation kernel reads per-world records (dev.rec is large enough for either format)
  nodedup.level_start = nullptr;
  const double dt = 0.001 * 500;   // kSimTimeStep * 500 (ensembles.cc:611)
  for (int it = 0; it <= max_steps; it++) {
    LK(egg_launch_collide(nodedup, b->stream));
    LK(egg_launch_assemble(nodedup, dt, b->stream));
    CK(cudaMemsetAsync(b->relax_any, 0, sizeof(int), b->stream));
    LK(egg_launch_relax(nodedup, dt, 0.2, max_steps, 0, b->dense_scratch, b->dense_scratch_bytes, b->relax_prog, b->relax_err2, b->relax_any, b->stream));
    b->launches += 3;
    int h_any = 0;
    CK(cudaMemcpyAsync(&h_any, b->relax_any, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    if (!h_any) break;
  }
  LK(egg_launch_collide(b->dev, b->stream));   // CheckAndCorrectEnsembleState (ensembles.cc:618)
  b->launches++;
  return relax_report(b, steps_out, err_sq_out);
}

int egg_post_stabilize(egg_batch* b, int max_steps, int* steps_out, double* err_sq_out) {
  if (!b || max_steps < 0) return EGG_ERR_ARG;
  if (!b->initialised) { g_err = "egg_post_stabilize before egg_init"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  RET(relax_prepare(b));
  EggDev pw = b->dev;
  pw.rec_fmt = 0;
  pw.level_start = nullptr;
  const double dt = 0.001 * 100;   // kSimTimeStep * 100 (ensembles.cc:634)
  // ensembles.cc:624-645: the contact list is NOT refreshed inside this loop; the rows are rebuilt
  // from the current body state and the stored contact geometry every iteration
  for (int it = 0; it <= max_steps; it++) {
    LK(egg_launch_assemble(pw, dt, b->stream));
    CK(cudaMemsetAsync(b->relax_any, 0, sizeof(int), b->stream));
    LK(egg_launch_relax(pw, dt, 0.2, max_steps, 1, b->dense_scratch, b->dense_scratch_bytes, b->relax_prog, b->relax_err2, b->relax_any, b->stream));
    b->launches += 2;
    int h_any = 0;
    CK(cudaMemcpyAsync(&h_any, b->relax_any, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    if (!h_any) break;
  }
  return relax_report(b, steps_out, err_sq_out);
}

int egg_step(egg_batch* b, double dt, int integrator, int n_steps) {
  if (!b || n_steps < 0 || !(dt > 0)) return EGG_ERR_ARG;
  if (!b->initialised) { g_err = "egg_step before egg_init"; return EGG_ERR_STATE; }
  if (integrator != EGG_OPEN_DYNAMICS_ENGINE) {
    // ensembles.cc:398-405: EXPLICIT_EULER refuses contacts, IMPLICIT_MIDPOINT panics.
    g_err = "only OPEN_DYNAMICS_ENGINE is supported (as in the reference once contacts exist)";
    return EGG_ERR_UNSUPPORTED;
  }
  const int solver = b->dev.prm.solver;
  if (solver < EGG_SOLVER_DENSE_MURTY || solver > EGG_SOLVER_SOR) { g_err = "unknown solver"; return EGG_ERR_ARG; }
  CK(cudaSetDevice(b->device));
  if (!b->iso_known) {   // one 4-byte read-back per egg_init: which M^-1 layout the PGS kernel may use
    int flag = 0;
    CK(cudaMemcpyAsync(&flag, b->dev.iso_flag, sizeof(int), cudaMemcpyDeviceToHost, b->stream));
    CK(cudaStreamSynchronize(b->stream));
    b->dev.iso = !(flag & 1) ? 0 : ((flag & 2) ? 2 : 1);   // 0 general, 1 isotropic, 2 isotropic and uniform
    if (getenv("EGG_PGS_ISO") && atoi(getenv("EGG_PGS_ISO")) < b->dev.iso) b->dev.iso = atoi(getenv("EGG_PGS_ISO"));
    b->iso_known = true;
    b->dev.rmax = egg_runs_rmax(b->dev);   // run format of the PGS stream for isotropic bodies, else one block per lane and stage
  }
  for (int s = 0; s < n_steps; s++) {
    cudaEvent_t e[4] = {nullptr, nullptr, nullptr, nullptr};
    if (b->profiling) {
      for (int k = 0; k < 4; k++) {
        if (!b->ev_free.empty()) { e[k] = b->ev_free.back(); b->ev_free.pop_back(); }
        else CK(cudaEventCreate(&e[k]));
        b->ev.push_back(e[k]);
      }
      CK(cudaEventRecord(e[0], b->stream));
    }
    // EGG_SYNC_DEBUG=1: synchronise after every kernel and name the one that failed
    static const bool dbg_sync = getenv("EGG_SYNC_DEBUG") && atoi(getenv("EGG_SYNC_DEBUG")) != 0;
#define DBG_SYNC(what)                                                                       \
  if (dbg_sync) {                                                                            \
    cudaError_t e__ = cudaStreamSynchronize(b->stream);                                      \
    if (e__ != cudaSuccess) { set_err(what, e__); return EGG_ERR_CUDA; }                      \
  }
    LK(egg_launch_collide(b->dev, b->stream));
    DBG_SYNC("egg_collide_kernel")
    if (b->profiling) CK(cudaEventRecord(e[1], b->stream));
    LK(egg_launch_assemble(b->dev, dt, b->stream));
    DBG_SYNC("assembly kernels")
    if (b->profiling) CK(cudaEventRecord(e[2], b->stream));
    if (solver == EGG_SOLVER_PGS) LK(egg_launch_solve_pgs(b->dev, dt, b->stream));
    else if (solver == EGG_SOLVER_DENSE_MURTY) LK(egg_launch_solve_dense(b->dev, dt, b->stream, b->dense_scratch, b->dense_scratch_bytes));
    else LK(egg_launch_solve_iter(b->dev, dt, solver, b->stream));
    DBG_SYNC("solve kernel")
#undef DBG_SYNC
    if (b->profiling) CK(cudaEventRecord(e[3], b->stream));
    b->launches += (solver == EGG_SOLVER_PGS && b->dev.rec_fmt) ? 5 : 3;   // collide, assemble (1 or 3 kernels), solve
  }
  CK(cudaGetLastError());
  return EGG_OK;
}

int egg_update_contacts(egg_batch* b) {
  if (!b) return EGG_ERR_ARG;
  if (!b->initialised) { g_err = "egg_update_contacts before egg_init"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  LK(egg_launch_collide(b->dev, b->stream));
  b->launches++;
  return EGG_OK;
}

int egg_get_static(egg_batch* b, double* minv_lin, double* minv_ang, double* f_ext) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const int n = b->dev.n;
  if (minv_lin) RET(download(b, minv_lin, n, 1, b->dev.stat, EGG_STAT, 0));
  if (minv_ang) RET(download(b, minv_ang, n, 9, b->dev.stat, EGG_STAT, 1));
  if (f_ext) RET(download(b, f_ext, n, 6, b->dev.stat, EGG_STAT, 10));
  return EGG_OK;
}

int egg_snapshot(egg_batch* b) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const size_t cnt = (size_t)b->dev.W * EGG_DYN * b->dev.n;
  if (!b->snap) {
    int r = dalloc(b, &b->snap, cnt);
    if (r != EGG_OK) return r;
  }
  CK(cudaMemcpyAsync(b->snap, b->dev.dyn, cnt * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  return EGG_OK;
}
int egg_restore(egg_batch* b) {
  if (!b) return EGG_ERR_ARG;
  if (!b->snap) { g_err = "egg_restore before egg_snapshot"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  const size_t cnt = (size_t)b->dev.W * EGG_DYN * b->dev.n;
  CK(cudaMemcpyAsync(b->dev.dyn, b->snap, cnt * sizeof(double), cudaMemcpyDeviceToDevice, b->stream));
  return EGG_OK;
}

int egg_get_bodies(egg_batch* b, double* p, double* R, double* v, double* w) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const int n = b->dev.n;
  if (p) RET(download(b, p, n, 3, b->dev.dyn, EGG_DYN, 0));
  if (R) RET(download(b, R, n, 9, b->dev.dyn, EGG_DYN, 3));
  if (v) RET(download(b, v, n, 3, b->dev.dyn, EGG_DYN, 12));
  if (w) RET(download(b, w, n, 3, b->dev.dyn, EGG_DYN, 15));
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}

// Contact taps of worlds [first, first + count): the geometry is transposed from the device's
// [world][7][max_contacts] layout to the host's [world][max_contacts][3] by the unpack kernel, in
// chunks of as many worlds as fit the staging buffer (no host-side pass over the data).
int egg_get_contacts_range(egg_batch* b, int first, int nworlds, int* count, int* i0, int* i1, double* pos, double* nrm,
                           double* depth, int* code, double* lambda, int* row_state) {
  if (!b || first < 0 || nworlds < 0 || first + nworlds > b->dev.W) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const EggDev& d = b->dev;
  const size_t W = (size_t)nworlds, mc = d.maxc, f = (size_t)first;
  if (W == 0) return EGG_OK;
  if (count) CK(cudaMemcpyAsync(count, d.c_count + f, W * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (i0) CK(cudaMemcpyAsync(i0, d.c_i0 + f * mc, W * mc * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (i1) CK(cudaMemcpyAsync(i1, d.c_i1 + f * mc, W * mc * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (code) CK(cudaMemcpyAsync(code, d.c_code + f * mc, W * mc * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (lambda) CK(cudaMemcpyAsync(lambda, d.lam_out + f * 3 * d.nrec, W * 3 * d.nrec * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  if (row_state) CK(cudaMemcpyAsync(row_state, d.row_state + f * 3 * d.nrec, W * 3 * d.nrec * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (pos || nrm || depth) {
    size_t chunk = b->stage_bytes / (mc * 3 * sizeof(double));
    if (chunk == 0) { g_err = "staging buffer smaller than one world's contact geometry"; return EGG_ERR_UNSUPPORTED; }
    struct { double* host; int comps, off; } part[3] = {{pos, 3, 0}, {nrm, 3, 3}, {depth, 1, 6}};
    for (size_t w0 = 0; w0 < W; w0 += chunk) {
      const size_t nw = (W - w0 < chunk) ? W - w0 : chunk;
      for (auto& pt : part) {
        if (!pt.host) continue;
        LK(egg_launch_unpack((int)nw, b->stage, (int)mc, pt.comps, d.c_geom + (f + w0) * 7 * mc, 7, pt.off, b->stream));
        b->launches++;
        CK(cudaMemcpyAsync(pt.host + w0 * mc * pt.comps, b->stage, nw * mc * pt.comps * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
        CK(cudaStreamSynchronize(b->stream));      // the staging buffer is reused by the next part
      }
    }
  }
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}

int egg_get_contacts(egg_batch* b, int* count, int* i0, int* i1, double* pos, double* nrm,
                     double* depth, int* code, double* lambda, int* row_state) {
  if (!b) return EGG_ERR_ARG;
  return egg_get_contacts_range(b, 0, b->dev.W, count, i0, i1, pos, nrm, depth, code, lambda, row_state);
}

int egg_get_pair_hits_range(egg_batch* b, int first, int nworlds, int* n_hits, int* pi, int* pj, int* code, int* count, int max_pairs) {
  if (!b || !n_hits || first < 0 || nworlds < 0 || first + nworlds > b->dev.W) return EGG_ERR_ARG;
  const EggDev& d = b->dev;
  if (!d.pair_code) { g_err = "egg_get_pair_hits needs desc.taps = 1"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  const size_t W = (size_t)nworlds, P = d.P, f = (size_t)first;
  if (W == 0 || P == 0) { for (size_t w = 0; w < W; w++) n_hits[w] = 0; return EGG_OK; }
  std::vector<unsigned char> pc(W * P), pn(W * P);
  CK(cudaMemcpyAsync(pc.data(), d.pair_code + f * P, W * P, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaMemcpyAsync(pn.data(), d.pair_cnt + f * P, W * P, cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  for (size_t w = 0; w < W; w++) {
    int h = 0;
    size_t q = 0;
    for (int i = 0; i < d.n; i++)
      for (int j = i + 1; j < d.n; j++, q++) {
        if (!pc[w * P + q]) continue;
        if (h < max_pairs) {
          if (pi) pi[w * max_pairs + h] = i;
          if (pj) pj[w * max_pairs + h] = j;
          if (code) code[w * max_pairs + h] = pc[w * P + q];
          if (count) count[w * max_pairs + h] = pn[w * P + q];
        }
        h++;
      }
    n_hits[w] = h;
  }
  return EGG_OK;
}

int egg_get_pair_hits(egg_batch* b, int* n_hits, int* pi, int* pj, int* code, int* count, int max_pairs) {
  if (!b) return EGG_ERR_ARG;
  return egg_get_pair_hits_range(b, 0, b->dev.W, n_hits, pi, pj, code, count, max_pairs);
}

int egg_get_status(egg_batch* b, int* status, int* stats, double* residual) {
  if (!b) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  const size_t W = b->dev.W;
  if (status) CK(cudaMemcpyAsync(status, b->dev.status, W * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (stats) CK(cudaMemcpyAsync(stats, b->dev.stats, W * 8 * sizeof(int), cudaMemcpyDeviceToHost, b->stream));
  if (residual) CK(cudaMemcpyAsync(residual, b->dev.resid, W * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}

int egg_get_debug_counters(egg_batch* b, unsigned long long* out32, int reset) {
  if (!b || !out32) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  CK(cudaMemcpyAsync(out32, b->dev.dbg, 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, b->stream));
  if (reset) CK(cudaMemsetAsync(b->dev.dbg, 0, 32 * sizeof(unsigned long long), b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}

int egg_get_dense_work(egg_batch* b, double* flops) {
  if (!b || !flops) return EGG_ERR_ARG;
  if (!b->dev.work) { g_err = "egg_get_dense_work: the batch was not created with the dense solver"; return EGG_ERR_STATE; }
  CK(cudaSetDevice(b->device));
  CK(cudaMemcpyAsync(flops, b->dev.work, (size_t)b->dev.W * sizeof(double), cudaMemcpyDeviceToHost, b->stream));
  CK(cudaStreamSynchronize(b->stream));
  return EGG_OK;
}

int egg_rollout_costs(egg_batch* b, double* cost_d) {
  if (!b || !cost_d) return EGG_ERR_ARG;
  CK(cudaSetDevice(b->device));
  LK(egg_launch_costs(b->dev, cost_d, b->stream));
  b->launches++;
  return EGG_OK;
}

}  // extern "C"
