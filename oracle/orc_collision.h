// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_linalg.h header).
//
// CPU restatement of eggshell's narrowphase: /root/reference/eggshell/collision.cc:14-432.
// Every threshold, comparison direction and emission order follows the reference because the
// discrete outputs (hit / code / contact count / order) are compared bit-exactly.
#pragma once
#include "orc_linalg.h"

namespace orc {

struct Box {            // collision.cc:14-18
  Vec3 center;
  Mat3 R;
  Vec3 halfside;
};
struct ContactGeometry { // collision.h:11-27
  Vec3 position, normal;
  double depth = 0;
};
struct CollisionInfo {   // collision.h:33-50
  double depth = 0;
  Vec3 separating_axis;
  int code = 0;
};
struct Vec2 { double x = 0, y = 0; Vec2() {} Vec2(double a, double b) : x(a), y(b) {} };

inline double sign1(double a) { return (a >= 0) ? 1.0 : -1.0; }   // collision.cc:26-28, Sign(0)=+1

// collision.cc:47-62
inline void line_closest_approach(const Vec3& pa, const Vec3& ua, const Vec3& pb, const Vec3& ub,
                                  double* alpha, double* beta) {
  Vec3 p = pb - pa;
  double uaub = dot(ua, ub);
  double q1 = dot(ua, p);
  double q2 = -dot(ub, p);
  double d = 1 - uaub * uaub;
  if (d == 0) {
    *alpha = 0;
    *beta = 0;
  } else {
    *alpha = (q1 + uaub * q2) / d;
    *beta = (uaub * q1 + q2) / d;
  }
}

// collision.cc:70-80
inline bool intersect_line_segment_and_line(const Vec2& p1, const Vec2& p2, const Vec2& normal,
                                            double d, Vec2* p) {
  double k1 = normal.x * p1.x + normal.y * p1.y + d;
  double k2 = normal.x * p2.x + normal.y * p2.y + d;
  if (k1 * k2 < 0) {
    double t = k1 / (k2 - k1);
    p->x = p1.x - t * (p2.x - p1.x);
    p->y = p1.y - t * (p2.y - p1.y);
    return true;
  }
  return false;
}

// collision.cc:84-99
inline void clip_polygon_by_half_space(const std::vector<Vec2>& poly, const Vec2& normal, double d,
                                       std::vector<Vec2>* newpoly) {
  newpoly->clear();
  for (size_t i = 0; i < poly.size(); i++) {
    if (normal.x * poly[i].x + normal.y * poly[i].y + d >= 0) newpoly->push_back(poly[i]);
    Vec2 newp;
    if (intersect_line_segment_and_line(poly[i], poly[(i + 1) % poly.size()], normal, d, &newp))
      newpoly->push_back(newp);
  }
}

// collision.cc:105-158
inline void intersect_box_and_rectangle(const Box& B, const Box& R, std::vector<Vec2>* poly) {
  const double kTolerance = 1e-9;
  Vec3 Bc = B.center - R.center;
  poly->clear();
  poly->push_back(Vec2(-R.halfside[0], -R.halfside[1]));
  poly->push_back(Vec2(-R.halfside[0], R.halfside[1]));
  poly->push_back(Vec2(R.halfside[0], R.halfside[1]));
  poly->push_back(Vec2(R.halfside[0], -R.halfside[1]));
  std::vector<Vec2> newpoly;
  Vec3 Rnormal = R.R.col(2);
  for (int i = 0; i < 3; i++) {
    Vec3 Bnormal = B.R.col(i);
    double BnBc = dot(Bnormal, Bc);
    double crs = norm(cross(Bnormal, Rnormal));
    for (int j = -1; j <= 1; j += 2) {
      double Bd = -j * BnBc - B.halfside[i];
      if (crs < kTolerance) {
        if (Bd <= 0) continue;
        poly->clear();
        return;
      }
      Vec2 Hn(dot(R.R.col(0), Bnormal), dot(R.R.col(1), Bnormal));
      clip_polygon_by_half_space(*poly, Vec2(-j * Hn.x, -j * Hn.y), -Bd, &newpoly);
      newpoly.swap(*poly);
      if (poly->empty()) return;
    }
  }
}

// collision.cc:166-388
inline bool collide_boxes(const Box& box1, const Box& box2, CollisionInfo* info,
                          std::vector<ContactGeometry>* contacts) {
  const double kAlignmentTolerance = 0.9962;
  const double kTolerance = 1e-9;
  const Mat3& R1 = box1.R;
  const Mat3& R2 = box2.R;
  Mat3 R = transpose(R1) * R2;
  Vec3 p = tmul(R1, box2.center - box1.center);
  Mat3 Q;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Q.m[i][j] = std::fabs(R.m[i][j]);

  int aacount = 0;
  for (int i = 0; i < 3; i++) {
    double mx = std::max(Q.m[0][i], std::max(Q.m[1][i], Q.m[2][i]));
    aacount += (mx > kAlignmentTolerance);
  }

  double min_depth_FN = -std::numeric_limits<double>::max();
  Vec3 sepaxis_FN;
  int code_FN = 0;
  const Vec3& H1 = box1.halfside;
  const Vec3& H2 = box2.halfside;
  // Face-normal axes: box1 faces (codes 1-3) then box2 faces (codes 4-6); collision.cc:210-230.
  auto sep_fn = [&](double e1, double extent, const Vec3& nrm, int thecode) -> bool {
    double separation = std::fabs(e1) - extent;
    if (separation > 0) return false;
    if (separation > min_depth_FN) {       // strict >: earlier axis wins ties
      min_depth_FN = separation;
      sepaxis_FN = sign1(e1) * nrm;
      code_FN = thecode;
    }
    return true;
  };
  for (int i = 0; i < 3; i++)
    if (!sep_fn(p[i], H1[i] + dot(H2, Q.row(i)), R1.col(i), 1 + i)) return false;
  for (int i = 0; i < 3; i++)
    if (!sep_fn(dot(R.col(i), p), dot(H1, Q.col(i)) + H2[i], R2.col(i), 4 + i)) return false;

  // Edge x edge axes e_i(box1) x e_j(box2), codes 7 + 3 i + j, expressed in box1's frame;
  // collision.cc:237-271.  With (a,b) the two indices other than i (resp. j) in ascending
  // order, i1=(i+1)%3, i2=(i+2)%3:  n[i]=0, n[i1]=-R(i2,j), n[i2]=R(i1,j);
  // e1 = p[i2] R(i1,j) - p[i1] R(i2,j);
  // extent = H1[a] Q(b,j) + H1[b] Q(a,j) + H2[a'] Q(i,b') + H2[b'] Q(i,a')  (summed left to right).
  double min_depth_EE = -std::numeric_limits<double>::max();
  Vec3 sepaxis_EE;
  int code_EE = 0;
  for (int i = 0; i < 3; i++) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    const int ia = (i == 0) ? 1 : 0, ib = (i == 2) ? 1 : 2;
    for (int j = 0; j < 3; j++) {
      const int ja = (j == 0) ? 1 : 0, jb = (j == 2) ? 1 : 2;
      Vec3 n;
      n[i1] = -R.m[i2][j];
      n[i2] = R.m[i1][j];
      double len = norm(n);
      if (!(len > kTolerance)) continue;
      double e1 = p[i2] * R.m[i1][j] - p[i1] * R.m[i2][j];
      double extent = H1[ia] * Q.m[ib][j] + H1[ib] * Q.m[ia][j] + H2[ja] * Q.m[i][jb] + H2[jb] * Q.m[i][ja];
      double separation = std::fabs(e1) - extent;
      if (separation > 0) return false;
      separation /= len;
      if (separation > min_depth_EE) {
        min_depth_EE = separation;
        sepaxis_EE = n / (sign1(e1) * len);
        code_EE = 7 + 3 * i + j;
      }
    }
  }

  // collision.cc:279 CHECK(code_FN != 0 && code_EE != 0): with all nine edge axes degenerate
  // (fully aligned boxes) code_EE stays 0 and the reference would abort.  The oracle keeps going
  // with min_depth_EE = -DBL_MAX, which makes the face normal the best axis (documented quirk).
  sepaxis_EE = R1 * sepaxis_EE;
  bool best_FN = (min_depth_FN > min_depth_EE);
  if (info) {
    if (best_FN) {
      info->depth = -min_depth_FN;
      info->separating_axis = sepaxis_FN;
    } else {
      info->depth = -min_depth_EE;
      info->separating_axis = sepaxis_EE;
    }
  }

  if (aacount == 0 && !best_FN) {
    if (info) info->code = code_EE;
    Vec3 pa = box1.center, pb = box2.center;
    for (int j = 0; j < 3; j++) {
      pa = pa + (sign1(dot(sepaxis_EE, R1.col(j))) * H1[j]) * R1.col(j);
      pb = pb - (sign1(dot(sepaxis_EE, R2.col(j))) * H2[j]) * R2.col(j);
    }
    Vec3 ua = R1.col((code_EE - 7) / 3);
    Vec3 ub = R2.col((code_EE - 7) % 3);
    double alpha, beta;
    line_closest_approach(pa, ua, pb, ub, &alpha, &beta);
    ContactGeometry c;
    c.position = (pa + ua * alpha + pb + ub * beta) * 0.5;
    c.normal = sepaxis_EE;
    c.depth = -min_depth_EE;
    contacts->push_back(c);
    return true;
  }

  if (info) info->code = code_FN;
  const Box& A = (code_FN <= 3) ? box1 : box2;
  Box B = (code_FN <= 3) ? box2 : box1;
  Vec3 Aface_normal = sepaxis_FN * ((code_FN <= 3) ? 1.0 : -1.0);

  Vec3 nf = tmul(B.R, Aface_normal);
  int nf_index = 0;   // Eigen maxCoeff(&idx): first maximum wins
  {
    double best = std::fabs(nf[0]);
    for (int i = 1; i < 3; i++)
      if (std::fabs(nf[i]) > best) { best = std::fabs(nf[i]); nf_index = i; }
  }
  Vec3 Bface_normal = (-sign1(nf[nf_index])) * B.R.col(nf_index);
  {
    B.center = B.center + Bface_normal * B.halfside[nf_index];
    Mat3 BR;
    BR.set_col(0, B.R.col((nf_index + 1) % 3));
    BR.set_col(1, B.R.col((nf_index + 2) % 3));
    BR.set_col(2, B.R.col(nf_index));
    Vec3 Bh(B.halfside[(nf_index + 1) % 3], B.halfside[(nf_index + 2) % 3], 0);
    B.R = BR;
    B.halfside = Bh;
  }
  Vec3 AfaceCenter = A.center + Aface_normal * A.halfside[(code_FN - 1) % 3];
  double Ad = -dot(Aface_normal, AfaceCenter);

  std::vector<Vec2> poly;
  intersect_box_and_rectangle(A, B, &poly);
  for (size_t i = 0; i < poly.size(); i++) {
    Vec3 pos = B.center + B.R.col(0) * poly[i].x + B.R.col(1) * poly[i].y;
    double depth = -(dot(Aface_normal, pos) + Ad);
    if (std::fabs(depth) > kTolerance || aacount >= 2) {
      ContactGeometry c;
      c.position = pos;
      c.normal = sepaxis_FN;
      c.depth = depth;
      contacts->push_back(c);
    }
  }
  if (contacts->empty()) {   // q7: tests the caller's vector (fresh per pair in UpdateContacts)
    ContactGeometry c;
    c.position = box2.center;
    c.normal = sepaxis_FN;
    c.depth = -min_depth_FN;
    contacts->push_back(c);
    if (info) info->code = 16;
  }
  return true;
}

// collision.cc:408-432
inline bool collide_box_and_ground(const Vec3& center, const Mat3& rotation, const Vec3& side,
                                   std::vector<ContactGeometry>* contacts) {
  bool retval = false;
  for (int x = -1; x <= 1; x += 2)
    for (int y = -1; y <= 1; y += 2)
      for (int z = -1; z <= 1; z += 2) {
        Vec3 v = center + rotation.col(0) * side[0] * 0.5 * x + rotation.col(1) * side[1] * 0.5 * y +
                 rotation.col(2) * side[2] * 0.5 * z;
        if (v[2] < 0) {
          ContactGeometry c;
          c.position = v;
          c.normal = Vec3(0, 0, 1);
          c.depth = -v[2];
          contacts->push_back(c);
          retval = true;
        }
      }
  return retval;
}

// ---- Spheres and capsules (BASELINE.json north_star (a); NOT in the reference, whose only collider
// is the box, body.h:90-91 -- the reference merely has DrawSphere / DrawCapsule, model.h:16-25).
// Defined here by analogy with CollideBoxAndGround: "parity unpinned", the oracle is the definition.
// Shape codes: 0 box (dims = side lengths), 1 sphere (dims[0] = radius), 2 capsule (dims[0] = radius,
// dims[1] = length of the axis segment, along the body's z axis).
// Ground: the lowest point of every sphere (a capsule = two end spheres, -z end first) is a contact
// when it is below z = 0: position = that point, normal +z, depth = -z, as for a box vertex.
inline bool collide_round_and_ground(int shape, const Vec3& center, const Mat3& rotation, const Vec3& dims,
                                     std::vector<ContactGeometry>* contacts) {
  bool retval = false;
  const double r = dims[0];
  const int ends = (shape == 2) ? 2 : 1;
  for (int e = 0; e < ends; e++) {
    Vec3 c = center;
    if (shape == 2) c = center + rotation.col(2) * (dims[1] * 0.5 * (e == 0 ? -1.0 : 1.0));
    Vec3 v(c[0], c[1], c[2] - r);
    if (v[2] < 0) {
      ContactGeometry g;
      g.position = v;
      g.normal = Vec3(0, 0, 1);
      g.depth = -v[2];
      contacts->push_back(g);
      retval = true;
    }
  }
  return retval;
}
// Sphere - sphere: one contact at the middle of the overlap, normal out of sphere 1 (the box-box
// convention, collision.cc:325-387), depth = r1 + r2 - distance > 0; code 17.
inline bool collide_spheres(const Vec3& c1, double r1, const Vec3& c2, double r2, CollisionInfo* info,
                            std::vector<ContactGeometry>* contacts) {
  Vec3 d = c2 - c1;
  const double dist = norm(d);
  const double depth = (r1 + r2) - dist;
  if (!(depth > 0)) return false;
  Vec3 n = (dist > 0) ? d / dist : Vec3(0, 0, 1);
  ContactGeometry g;
  g.position = c1 + n * (r1 - depth * 0.5);
  g.normal = n;
  g.depth = depth;
  contacts->push_back(g);
  info->depth = depth; info->separating_axis = n; info->code = 17;
  return true;
}

// collision.cc:443-473 (test helpers): slow-but-sure 15-axis separation test.
inline bool boxes_separated_by_axis(const Box& b1, const Box& b2, const Vec3& axis) {
  double span1 = b1.halfside[0] * std::fabs(dot(axis, b1.R.col(0))) +
                 b1.halfside[1] * std::fabs(dot(axis, b1.R.col(1))) +
                 b1.halfside[2] * std::fabs(dot(axis, b1.R.col(2)));
  double span2 = b2.halfside[0] * std::fabs(dot(axis, b2.R.col(0))) +
                 b2.halfside[1] * std::fabs(dot(axis, b2.R.col(1))) +
                 b2.halfside[2] * std::fabs(dot(axis, b2.R.col(2)));
  return std::fabs(dot(axis, b1.center) - dot(axis, b2.center)) > (span1 + span2);
}
inline bool boxes_separated(const Box& b1, const Box& b2) {
  for (int i = 0; i < 3; i++) if (boxes_separated_by_axis(b1, b2, b1.R.col(i))) return true;
  for (int i = 0; i < 3; i++) if (boxes_separated_by_axis(b1, b2, b2.R.col(i))) return true;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      if (boxes_separated_by_axis(b1, b2, cross(b1.R.col(i), b2.R.col(j)))) return true;
  return false;
}

}  // namespace orc
