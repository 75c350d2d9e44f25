// Mirror of /root/reference/eggshell/collision.h:11-68.  The two free functions run the CUDA
// narrowphase on a throw-away one-world batch (they are convenience entry points; the step path
// calls the same kernel on the whole batch).
#ifndef EGGSHELL_COLLISION_H_
#define EGGSHELL_COLLISION_H_
#include <vector>
#include "linalg.h"

struct ContactGeometry {
  Vector3d position, normal;
  double depth = 0;
  ContactGeometry() {}
  ContactGeometry(const Vector3d& p, const Vector3d& n, double d) : position(p), normal(n), depth(d) {}
};
struct CollisionInfo {
  double depth = 0;
  Vector3d separating_axis;
  int code = 0;
};
// Only cubes of side 0.3 are representable (body.h:90-91): side_lengths must equal (0.3,0.3,0.3).
bool CollideBoxes(const Vector3d& center1, const Matrix3d& rotation1, const Vector3d& side_lengths1, const Vector3d& center2,
                  const Matrix3d& rotation2, const Vector3d& side_lengths2, CollisionInfo* info, std::vector<ContactGeometry>* contacts);
bool CollideBoxAndGround(const Vector3d& center, const Matrix3d& rotation, const Vector3d& side_lengths,
                         std::vector<ContactGeometry>* contacts);
#endif
