// CollideBoxes / CollideBoxAndGround (mirror of /root/reference/eggshell/collision.h:54-68) on the
// device narrowphase kernel via a throw-away one-world batch.
#include "eggshell/collision.h"

#include <cmath>
#include <vector>

#include "egg_cuda.h"
#include "eggshell/model.h"

namespace {
void fill(double* p, double* R, int i, const Vector3d& c, const Matrix3d& rot) {
  for (int k = 0; k < 3; k++) p[3 * i + k] = c(k);
  for (int r = 0; r < 3; r++) for (int q = 0; q < 3; q++) R[9 * i + 3 * r + q] = rot(r, q);
}
int run(int n, const double* p, const double* R, const double* side, std::vector<ContactGeometry>* out, int want_i0, int* code) {
  egg_desc d;
  egg_desc_default(&d, 1, n, 0);
  d.solver = EGG_SOLVER_PGS;
  d.min_constraint_dist = 0.0;   // the free functions do not de-duplicate (collision.cc)
  egg_batch* b = nullptr;
  if (egg_create(&d, &b) != EGG_OK) Panic("CollideBoxes: %s", egg_last_error());
  std::vector<double> z(3 * n, 0.0), m(n, 1.0), I(9 * n, 0.0);
  for (int i = 0; i < n; i++) I[9 * i] = I[9 * i + 4] = I[9 * i + 8] = 0.015;
  egg_set_bodies(b, p, R, z.data(), z.data(), m.data(), I.data(), side);
  egg_init(b);
  egg_update_contacts(b);
  const int mc = egg_capacity(b);
  int count = 0;
  std::vector<int> i0(mc), i1(mc), cd(mc);
  std::vector<double> pos(3 * mc), nrm(3 * mc), depth(mc);
  egg_get_contacts(b, &count, i0.data(), i1.data(), pos.data(), nrm.data(), depth.data(), cd.data(), nullptr, nullptr);
  int hits = 0;
  for (int k = 0; k < count; k++) {
    if ((want_i0 < 0) != (i0[k] < 0)) continue;
    out->push_back(ContactGeometry(Vector3d(pos[3 * k], pos[3 * k + 1], pos[3 * k + 2]), Vector3d(nrm[3 * k], nrm[3 * k + 1], nrm[3 * k + 2]), depth[k]));
    if (code) *code = cd[k];
    hits++;
  }
  egg_destroy(b);
  return hits;
}
}  // namespace

bool CollideBoxAndGround(const Vector3d& center, const Matrix3d& rotation, const Vector3d& side_lengths, std::vector<ContactGeometry>* contacts) {
  double p[3], R[9], s[3] = {side_lengths(0), side_lengths(1), side_lengths(2)};
  fill(p, R, 0, center, rotation);
  return run(1, p, R, s, contacts, -1, nullptr) > 0;
}

bool CollideBoxes(const Vector3d& c1, const Matrix3d& r1, const Vector3d& s1, const Vector3d& c2, const Matrix3d& r2, const Vector3d& s2,
                  CollisionInfo* info, std::vector<ContactGeometry>* contacts) {
  double p[6], R[18], s[6] = {s1(0), s1(1), s1(2), s2(0), s2(1), s2(2)};
  fill(p, R, 0, c1, r1);
  fill(p, R, 1, c2, r2);
  // lift both boxes far above the ground plane so that only the pair test can fire
  const double lift = 1e3;
  p[2] += lift; p[5] += lift;
  std::vector<ContactGeometry> tmp;
  int code = 0;
  int hits = run(2, p, R, s, &tmp, 0, &code);
  for (auto& c : tmp) { c.position(2) -= lift; contacts->push_back(c); }
  if (info) {
    info->code = hits ? code : 0;
    if (hits) { info->separating_axis = tmp[0].normal; info->depth = tmp[0].depth; }
  }
  return hits > 0;
}
