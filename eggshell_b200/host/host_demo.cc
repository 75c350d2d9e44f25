// Headless stand-in for the viewer's run loop (/root/reference/eggshell/eggshell_view.cc:540-554)
// and its draw-list (eggshell_view.cc:365-422): the reference has no on-disk format, so the state
// of a W-world batch is inspected / replayed through a small binary dump of our own.
//
//   host_demo [steps]
//       SimulationInitialization(), N x SimulationStep(), then a state dump of the hanging chain
//       (one line per body: index p.x p.y p.z v.x v.y v.z) and of the cairn.
//   host_demo batch <worlds> <links> <steps> <file>
//       W hanging chains (Chain(links, anchor), ensembles.cc:668-707, anchors and initial
//       velocities perturbed per world by a fixed LCG) on one device batch through the C ABI:
//       <steps> steps, write <file>, <steps> more steps, print "checksum <hex> <sum>".
//   host_demo replay <file> <steps>
//       rebuild the batch from <file>, take <steps> steps, print the same checksum line: a replay
//       from the dump must reproduce the original run bit for bit.
//   host_demo show <file> [world]
//       print one world of a dump in the text format of the first mode.
//
// Dump format (little endian): char magic[8] = "EGGSTAT1"; int32 W, n, nj, solver, k_max, step;
// double dt; then m[W n], I[W n 9], joints i0[W nj], i1[W nj] (int32), c0[W nj 3], c1[W nj 3],
// f_ext[W n 6], p[W n 3], R[W n 9], v[W n 3], w[W n 3] (doubles, world-major as in
// include/egg_cuda.h).  f_ext is part of the state: the reference freezes M^-1 and the external
// force / gyroscopic torque at Init (ensembles.cc:202-222 are never refreshed by Step), so a replay
// that re-initialised from the dumped velocities would see a different torque.  (M^-1 is
// recomputed by egg_init from the dumped R: identical for bodies with isotropic inertia, which is
// every body of the reference's scenes.)
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "egg_cuda.h"
#include "eggshell/ensembles.h"
#include "eggshell/model.h"

const Ensemble& EggshellHangingChain();
const Ensemble& EggshellCairn();

namespace {

struct Dump {
  int32_t W = 0, n = 0, nj = 0, solver = 1, k_max = 500, step = 0;
  double dt = 0.001;
  std::vector<double> m, I, c0, c1, fext, p, R, v, w;
  std::vector<int32_t> i0, i1;
};

void check(int rc, const char* what) {
  if (rc != EGG_OK) Panic("%s failed (%d): %s", what, rc, egg_last_error());
}

template <class T>
void put(FILE* f, const std::vector<T>& a) {
  if (!a.empty() && std::fwrite(a.data(), sizeof(T), a.size(), f) != a.size()) Panic("short write");
}
template <class T>
void get(FILE* f, std::vector<T>& a, size_t count) {
  a.resize(count);
  if (count && std::fread(a.data(), sizeof(T), count, f) != count) Panic("short read");
}

void save(const Dump& d, const char* path) {
  FILE* f = std::fopen(path, "wb");
  if (!f) Panic("cannot write %s", path);
  std::fwrite("EGGSTAT1", 1, 8, f);
  const int32_t hdr[6] = {d.W, d.n, d.nj, d.solver, d.k_max, d.step};
  std::fwrite(hdr, sizeof(int32_t), 6, f);
  std::fwrite(&d.dt, sizeof(double), 1, f);
  put(f, d.m); put(f, d.I); put(f, d.i0); put(f, d.i1); put(f, d.c0); put(f, d.c1); put(f, d.fext);
  put(f, d.p); put(f, d.R); put(f, d.v); put(f, d.w);
  std::fclose(f);
}

Dump load(const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) Panic("cannot read %s", path);
  char magic[8];
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "EGGSTAT1", 8) != 0) Panic("%s is not an eggshell state dump", path);
  Dump d;
  int32_t hdr[6];
  if (std::fread(hdr, sizeof(int32_t), 6, f) != 6 || std::fread(&d.dt, sizeof(double), 1, f) != 1) Panic("short read");
  d.W = hdr[0]; d.n = hdr[1]; d.nj = hdr[2]; d.solver = hdr[3]; d.k_max = hdr[4]; d.step = hdr[5];
  const size_t W = d.W, n = d.n, nj = d.nj;
  get(f, d.m, W * n); get(f, d.I, W * n * 9); get(f, d.i0, W * nj); get(f, d.i1, W * nj); get(f, d.c0, W * nj * 3); get(f, d.c1, W * nj * 3); get(f, d.fext, W * n * 6);
  get(f, d.p, W * n * 3); get(f, d.R, W * n * 9); get(f, d.v, W * n * 3); get(f, d.w, W * n * 3);
  std::fclose(f);
  return d;
}

egg_batch* make_batch(const Dump& d) {
  egg_desc desc;
  check(egg_desc_default(&desc, d.W, d.n, d.nj), "egg_desc_default");
  desc.solver = d.solver;
  desc.k_max = d.k_max;
  egg_batch* b = nullptr;
  check(egg_create(&desc, &b), "egg_create");
  check(egg_set_bodies(b, d.p.data(), d.R.data(), d.v.data(), d.w.data(), d.m.data(), d.I.data(), nullptr), "egg_set_bodies");
  if (d.nj) check(egg_set_joints(b, d.i0.data(), d.i1.data(), d.c0.data(), d.c1.data()), "egg_set_joints");
  check(egg_init(b), "egg_init");
  if (!d.fext.empty()) check(egg_set_external(b, d.fext.data()), "egg_set_external");   // a replay: the torque frozen at the original Init
  return b;
}

void fetch(egg_batch* b, Dump& d) { check(egg_get_bodies(b, d.p.data(), d.R.data(), d.v.data(), d.w.data()), "egg_get_bodies"); }

void print_checksum(const Dump& d) {
  uint64_t h = 1469598103934665603ull;                 // FNV-1a over the raw state bytes
  double sum = 0;
  for (const std::vector<double>* a : {&d.p, &d.R, &d.v, &d.w})
    for (double x : *a) {
      uint64_t u;
      std::memcpy(&u, &x, 8);
      for (int k = 0; k < 8; k++) { h ^= (u >> (8 * k)) & 0xff; h *= 1099511628211ull; }
      sum += x;
    }
  std::printf("checksum %016llx %.17g\n", (unsigned long long)h, sum);
}

// W copies of Chain(links, anchor) with per-world anchor / velocity perturbations.
Dump chain_batch(int W, int links) {
  Dump d;
  d.W = W; d.n = links; d.nj = links;
  const size_t n = links;
  d.m.assign((size_t)W * n, 1.0);
  d.I.assign((size_t)W * n * 9, 0.0);
  d.p.assign((size_t)W * n * 3, 0.0); d.R.assign((size_t)W * n * 9, 0.0); d.v.assign((size_t)W * n * 3, 0.0); d.w.assign((size_t)W * n * 3, 0.0);
  d.i0.resize((size_t)W * n); d.i1.resize((size_t)W * n); d.c0.assign((size_t)W * n * 3, 0.0); d.c1.assign((size_t)W * n * 3, 0.0);
  Chain proto(links, Vector3d(0, 0, 0));                // geometry of one chain from the mirror's own builder
  uint64_t s = 0x9e3779b97f4a7c15ull;
  auto rnd = [&]() { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)(s >> 11) / 9007199254740992.0 * 2.0 - 1.0; };
  const double I0 = 1.0 / 12 * (0.3 * 0.3 + 0.3 * 0.3);  // body.cc:19-36
  for (int w = 0; w < W; w++) {
    const double ax = 0.05 * rnd(), ay = 0.05 * rnd(), az = 0.5 + 0.2 * rnd();
    for (int i = 0; i < links; i++) {
      const Body& b = *proto.components()[i];
      const size_t k = (size_t)w * n + i;
      d.p[3 * k] = b.p()(0) + ax; d.p[3 * k + 1] = b.p()(1) + ay; d.p[3 * k + 2] = b.p()(2) + az;
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) d.R[9 * k + 3 * r + c] = b.R()(r, c);
      d.I[9 * k] = d.I[9 * k + 4] = d.I[9 * k + 8] = I0;
      if (i > 0) for (int a = 0; a < 3; a++) d.v[3 * k + a] = 0.1 * rnd();
      // joints in the reference's order: links-1 chain joints, then the world anchor of link 0
      if (i < links - 1) {
        d.i0[k] = i; d.i1[k] = i + 1;
        const double c0[3] = {0.15, -0.15, 0.15}, c1[3] = {-0.15, 0.15, -0.15};
        for (int a = 0; a < 3; a++) { d.c0[3 * k + a] = c0[a]; d.c1[3 * k + a] = c1[a]; }
      } else {
        d.i0[k] = 0; d.i1[k] = -1;
        const size_t k0 = (size_t)w * n;
        for (int a = 0; a < 3; a++) { d.c0[3 * k + a] = 0.0; d.c1[3 * k + a] = d.p[3 * k0 + a]; }
      }
    }
  }
  return d;
}

void show_world(const Dump& d, int w) {
  std::printf("world %d of %d, step %d, dt %g\n", w, d.W, d.step, d.dt);
  for (int i = 0; i < d.n; i++) {
    const size_t k = (size_t)w * d.n + i;
    std::printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", i, d.p[3 * k], d.p[3 * k + 1], d.p[3 * k + 2], d.v[3 * k], d.v[3 * k + 1], d.v[3 * k + 2]);
  }
}

}  // namespace

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "";
  if (mode == "batch" && argc >= 6) {
    const int W = std::atoi(argv[2]), links = std::atoi(argv[3]), steps = std::atoi(argv[4]);
    Dump d = chain_batch(W, links);
    egg_batch* b = make_batch(d);
    check(egg_step(b, d.dt, EGG_OPEN_DYNAMICS_ENGINE, steps), "egg_step");
    fetch(b, d);
    d.step = steps;
    d.fext.resize((size_t)W * links * 6);
    check(egg_get_static(b, nullptr, nullptr, d.fext.data()), "egg_get_static");
    save(d, argv[5]);
    check(egg_step(b, d.dt, EGG_OPEN_DYNAMICS_ENGINE, steps), "egg_step");
    fetch(b, d);
    std::vector<int> status(W);
    check(egg_get_status(b, status.data(), nullptr, nullptr), "egg_get_status");
    int st_or = 0;
    for (int x : status) st_or |= x;
    std::printf("batch %d worlds x %d links, %d + %d steps, status_or %d\n", W, links, steps, steps, st_or);
    print_checksum(d);
    egg_destroy(b);
    return 0;
  }
  if (mode == "replay" && argc >= 4) {
    Dump d = load(argv[2]);
    const int steps = std::atoi(argv[3]);
    egg_batch* b = make_batch(d);
    check(egg_step(b, d.dt, EGG_OPEN_DYNAMICS_ENGINE, steps), "egg_step");
    fetch(b, d);
    std::printf("replay of %d worlds from step %d, %d steps\n", d.W, d.step, steps);
    print_checksum(d);
    egg_destroy(b);
    return 0;
  }
  if (mode == "show" && argc >= 3) {
    Dump d = load(argv[2]);
    show_world(d, argc > 3 ? std::atoi(argv[3]) : 0);
    return 0;
  }
  int steps = argc > 1 ? std::atoi(argv[1]) : 10;
  SimulationInitialization();
  for (int s = 0; s < steps; s++)
    if (!SimulationStep()) break;
  const Ensemble& ch = EggshellHangingChain();
  std::printf("chain %d\n", (int)ch.components().size());
  int i = 0;
  for (const auto& b : ch.components()) {
    std::printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", i++, b->p()(0), b->p()(1), b->p()(2), b->v()(0), b->v()(1), b->v()(2));
  }
  std::printf("chain_constraints %d status %d\n", (int)ch.constraints().size(), ch.status());
  const Ensemble& ca = EggshellCairn();
  std::printf("cairn %d\n", (int)ca.components().size());
  i = 0;
  for (const auto& b : ca.components())
    std::printf("%d %.17g %.17g %.17g %.17g %.17g %.17g\n", i++, b->p()(0), b->p()(1), b->p()(2), b->v()(0), b->v()(1), b->v()(2));
  return 0;
}
