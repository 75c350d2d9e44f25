// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product path; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// Small Eigen-free linear algebra used by the CPU restatement of eggshell's step.  The reference
// delegates this arithmetic to Eigen 3.3.8 (un-vendored; /root/reference/common.mk:110), which is
// absent from the container, so the *published algorithms* of the Eigen entry points the reference
// calls are restated here.  PARITY UNPINNED for Eigen internals (pivot order inside LDLT, the
// anti-parallel branch of FromTwoVectors, JacobiSVD rounding): see DESIGN.md "Oracle".
#pragma once
#include <cmath>
#include <cstring>
#include <vector>
#include <limits>
#include <algorithm>

namespace orc {

struct Vec3 {
  double x = 0, y = 0, z = 0;
  Vec3() {}
  Vec3(double a, double b, double c) : x(a), y(b), z(c) {}
  double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(const Vec3& a, const Vec3& b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(const Vec3& a, const Vec3& b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator-(const Vec3& a) { return Vec3(-a.x, -a.y, -a.z); }
inline Vec3 operator*(const Vec3& a, double s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator*(double s, const Vec3& a) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(const Vec3& a, double s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(const Vec3& a, const Vec3& b) {
  return Vec3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline double norm(const Vec3& a) { return std::sqrt(dot(a, a)); }
// Eigen 3.3 MatrixBase::normalized(): z = squaredNorm(); z > 0 ? n / sqrt(z) : n.
inline Vec3 normalized(const Vec3& a) {
  double z = dot(a, a);
  if (z > 0) return a / std::sqrt(z);
  return a;
}

// 3x3, row-major m[r][c].
struct Mat3 {
  double m[3][3];
  Mat3() { std::memset(m, 0, sizeof(m)); }
  static Mat3 identity() { Mat3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
  Vec3 col(int c) const { return Vec3(m[0][c], m[1][c], m[2][c]); }
  Vec3 row(int r) const { return Vec3(m[r][0], m[r][1], m[r][2]); }
  void set_col(int c, const Vec3& v) { m[0][c] = v.x; m[1][c] = v.y; m[2][c] = v.z; }
};
inline Mat3 operator*(const Mat3& a, const Mat3& b) {
  Mat3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0;
      for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}
inline Vec3 operator*(const Mat3& a, const Vec3& v) {
  return Vec3(a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
              a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
              a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z);
}
inline Mat3 operator*(const Mat3& a, double s) {
  Mat3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][j] * s;
  return r;
}
inline Mat3 transpose(const Mat3& a) {
  Mat3 r;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[i][j] = a.m[j][i];
  return r;
}
// Transposed product a^T * v.
inline Vec3 tmul(const Mat3& a, const Vec3& v) {
  return Vec3(a.m[0][0] * v.x + a.m[1][0] * v.y + a.m[2][0] * v.z,
              a.m[0][1] * v.x + a.m[1][1] * v.y + a.m[2][1] * v.z,
              a.m[0][2] * v.x + a.m[1][2] * v.y + a.m[2][2] * v.z);
}
// Fixed-size 3x3 inverse: Eigen uses cofactors / determinant for sizes <= 4
// (reference call site: ensembles.cc:210 `b->I_g().inverse()`).
inline Mat3 inverse3(const Mat3& a) {
  const double (*m)[3] = a.m;
  double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1];
  double c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2];
  double c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  double id = 1.0 / det;
  Mat3 r;
  r.m[0][0] = c00 * id;
  r.m[1][0] = c01 * id;
  r.m[2][0] = c02 * id;
  r.m[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id;
  r.m[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id;
  r.m[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id;
  r.m[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  r.m[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  r.m[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return r;
}

// utils.cc:16-24 CrossMat: a x b = CrossMat(a) * b.
inline Mat3 cross_mat(const Vec3& a) {
  Mat3 r;
  r.m[0][1] = -a.z; r.m[0][2] = a.y;
  r.m[1][0] = a.z;  r.m[1][2] = -a.x;
  r.m[2][0] = -a.y; r.m[2][1] = a.x;
  return r;
}

struct Quat { double w = 1, x = 0, y = 0, z = 0; };
// Eigen QuaternionBase::toRotationMatrix().
inline Mat3 quat_to_mat(const Quat& q) {
  double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  Mat3 r;
  r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz;       r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz;       r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy;       r.m[2][1] = tyz + twx;       r.m[2][2] = 1 - (txx + tyy);
  return r;
}
// Eigen AngleAxis -> Quaternion: w = cos(a/2), vec = sin(a/2) * axis.
inline Quat angle_axis_to_quat(double angle, const Vec3& axis) {
  double ha = 0.5 * angle;
  Quat q;
  q.w = std::cos(ha);
  double s = std::sin(ha);
  q.x = s * axis.x; q.y = s * axis.y; q.z = s * axis.z;
  return q;
}
inline Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
// utils.cc:82-89 WtoQ: AngleAxis(|w| dt, w.normalized()) as a quaternion (zero w => identity).
inline Quat w_to_q(const Vec3& w, double dt) {
  return angle_axis_to_quat(norm(w) * dt, normalized(w));
}

// Eigen 3.3 Quaternion::FromTwoVectors(a, b) (reference call site utils.cc:233-236).  The
// anti-parallel branch in Eigen takes the axis from an SVD null vector whose direction is
// implementation-defined; this oracle pins its own rule there (PARITY UNPINNED): the axis is
// v0 x e_k, normalised, where e_k is the coordinate axis with the smallest |v0[k]| (first min).
inline Quat from_two_vectors(const Vec3& a, const Vec3& b) {
  Vec3 v0 = normalized(a), v1 = normalized(b);
  double c = dot(v1, v0);
  Quat q;
  if (c < -1.0 + 1e-12) {
    c = std::max(c, -1.0);
    int k = 0;
    if (std::fabs(v0[1]) < std::fabs(v0[k])) k = 1;
    if (std::fabs(v0[2]) < std::fabs(v0[k])) k = 2;
    Vec3 e(k == 0, k == 1, k == 2);
    Vec3 axis = normalized(cross(v0, e));
    double w2 = (1.0 + c) * 0.5;
    q.w = std::sqrt(w2);
    double s = std::sqrt(1.0 - w2);
    q.x = axis.x * s; q.y = axis.y * s; q.z = axis.z * s;
    return q;
  }
  Vec3 axis = cross(v0, v1);
  double s = std::sqrt((1.0 + c) * 2.0);
  double invs = 1.0 / s;
  q.x = axis.x * invs; q.y = axis.y * invs; q.z = axis.z * invs;
  q.w = s * 0.5;
  return q;
}
// utils.cc:233-236 AlignVectors.
inline Mat3 align_vectors(const Vec3& a, const Vec3& b) { return quat_to_mat(from_two_vectors(a, b)); }

// ------------------------------------------------------------------------------------------
// Dynamic dense matrices, row-major.
struct Mat {
  int r = 0, c = 0;
  std::vector<double> a;
  Mat() {}
  Mat(int r_, int c_) : r(r_), c(c_), a((size_t)r_ * c_, 0.0) {}
  double& operator()(int i, int j) { return a[(size_t)i * c + j]; }
  double operator()(int i, int j) const { return a[(size_t)i * c + j]; }
};
typedef std::vector<double> Vec;
typedef std::vector<unsigned char> Mask;

inline Mat matmul(const Mat& A, const Mat& B) {
  Mat C(A.r, B.c);
  for (int i = 0; i < A.r; i++)
    for (int k = 0; k < A.c; k++) {
      double aik = A(i, k);
      if (aik == 0) continue;  // exact: skipping 0*x terms does not change any finite sum
      for (int j = 0; j < B.c; j++) C(i, j) += aik * B(k, j);
    }
  return C;
}
inline Mat transpose(const Mat& A) {
  Mat T(A.c, A.r);
  for (int i = 0; i < A.r; i++) for (int j = 0; j < A.c; j++) T(j, i) = A(i, j);
  return T;
}
inline Vec matvec(const Mat& A, const Vec& x) {
  Vec y(A.r, 0.0);
  for (int i = 0; i < A.r; i++) {
    double s = 0;
    for (int j = 0; j < A.c; j++) s += A(i, j) * x[j];
    y[i] = s;
  }
  return y;
}
inline double norm(const Vec& v) {
  double s = 0;
  for (double x : v) s += x * x;
  return std::sqrt(s);
}

// utils.cc:93-199 bool-mask gather / scatter.
inline Mat select_submatrix(const Mat& A, const Mask& ri, const Mask& ci) {
  int sr = 0, sc = 0;
  for (auto b : ri) sr += b != 0;
  for (auto b : ci) sc += b != 0;
  Mat S(sr, sc);
  int si = 0;
  for (int i = 0; i < A.r; i++) {
    if (!ri[i]) continue;
    int sj = 0;
    for (int j = 0; j < A.c; j++)
      if (ci[j]) S(si, sj++) = A(i, j);
    si++;
  }
  return S;
}
inline Vec select_subvector(const Vec& v, const Mask& ind) {
  Vec s;
  for (size_t i = 0; i < v.size(); i++) if (ind[i]) s.push_back(v[i]);
  return s;
}
inline void update_submatrix(Mat& A, const Mask& ri, const Mask& ci, const Mat& m) {
  int si = 0;
  for (int i = 0; i < A.r; i++) {
    if (!ri[i]) continue;
    int sj = 0;
    for (int j = 0; j < A.c; j++)
      if (ci[j]) A(i, j) = m(si, sj++);
    si++;
  }
}
inline void update_subvector(Vec& v, const Mask& ind, const Vec& n) {
  int si = 0;
  for (size_t i = 0; i < v.size(); i++) if (ind[i]) v[i] = n[si++];
}
inline void update_subvector(Vec& v, const Mask& ind, double d) {
  for (size_t i = 0; i < v.size(); i++) if (ind[i]) v[i] = d;
}

// Eigen 3.3 LDLT<MatrixXd, Lower>: unblocked, symmetric pivoting on the largest |diagonal| of the
// not-yet-eliminated part (left-looking: the trailing diagonal has not been updated when it is
// searched), L unit lower, D diagonal.  solve() = P^T L^-T D^+ L^-1 P b with D^+ the
// pseudo-inverse (|d| <= 1/highest() => 0).  Reference call sites: lcp.cc:203,317,
// ensembles.cc:491,664.
struct LDLT {
  int n = 0;
  Mat m;                  // L strictly below the diagonal, D on it
  std::vector<int> tr;    // transpositions
  void compute(const Mat& A) {
    n = A.r;
    m = A;
    tr.assign(n, 0);
    Vec temp(n, 0.0);
    for (int k = 0; k < n; k++) {
      int big = k;
      double best = std::fabs(m(k, k));
      for (int i = k + 1; i < n; i++) {
        double v = std::fabs(m(i, i));
        if (v > best) { best = v; big = i; }
      }
      tr[k] = big;
      if (big != k) {
        int s = n - big - 1;
        for (int j = 0; j < k; j++) std::swap(m(k, j), m(big, j));
        for (int i = 0; i < s; i++) std::swap(m(big + 1 + i, k), m(big + 1 + i, big));
        std::swap(m(k, k), m(big, big));
        for (int i = k + 1; i < big; i++) std::swap(m(i, k), m(big, i));
      }
      int rs = n - k - 1;
      if (k > 0) {
        // The inner products run on fused multiply-adds: Eigen's product kernels are written with
        // pmadd, which is an FMA on every -march=native x86-64 build with FMA3 (the reference's
        // flags, common.mk:160,188), and an explicit std::fma pins the rounding independently of
        // the compiler's contraction setting.
        for (int j = 0; j < k; j++) temp[j] = m(j, j) * m(k, j);
        double s = 0;
        for (int j = 0; j < k; j++) s = std::fma(m(k, j), temp[j], s);
        m(k, k) -= s;
        for (int i = 0; i < rs; i++) {
          double t = 0;
          for (int j = 0; j < k; j++) t = std::fma(m(k + 1 + i, j), temp[j], t);
          m(k + 1 + i, k) -= t;
        }
      }
      double akk = m(k, k);
      bool valid = std::fabs(akk) > 0;
      if (k == 0 && !valid) {
        for (int j = 0; j < n; j++) tr[j] = j;
        break;
      }
      if (rs > 0 && valid)
        for (int i = 0; i < rs; i++) m(k + 1 + i, k) /= akk;
    }
  }
  Vec solve(const Vec& b) const {
    Vec x = b;
    for (int k = 0; k < n; k++) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
    for (int i = 0; i < n; i++) {           // L^-1
      double s = x[i];
      for (int j = 0; j < i; j++) s = std::fma(-m(i, j), x[j], s);
      x[i] = s;
    }
    const double tol = 1.0 / std::numeric_limits<double>::max();
    for (int i = 0; i < n; i++) {           // D^+
      if (std::fabs(m(i, i)) > tol) x[i] /= m(i, i); else x[i] = 0;
    }
    // L^-T as a column sweep: once x[j] is final it is eliminated from every row above it, so row
    // i receives its terms for j = n-1 down to i+1.  (Eigen's own kernel for this step is a
    // panelled, vectorised row sweep whose summation order depends on the SIMD width and on FMA
    // contraction of the build; no scalar order is "the" reference order.  The column sweep is the
    // one order that is parallel over rows, which is what the device needs to reproduce it bit
    // for bit.)
    for (int j = n - 1; j >= 1; j--) {
      const double xj = x[j];
      for (int i = 0; i < j; i++) x[i] = std::fma(-m(j, i), xj, x[i]);
    }
    for (int k = n - 1; k >= 0; k--) if (tr[k] != k) std::swap(x[k], x[tr[k]]);
    return x;
  }
};

// Dynamic MatrixXd::inverse() = PartialPivLU inverse (reference call sites lcp.cc:293-294).
inline Mat lu_inverse(const Mat& A) {
  int n = A.r;
  Mat lu = A;
  std::vector<int> perm(n);
  for (int i = 0; i < n; i++) perm[i] = i;
  for (int k = 0; k < n; k++) {
    int piv = k;
    double best = std::fabs(lu(k, k));
    for (int i = k + 1; i < n; i++) {
      double v = std::fabs(lu(i, k));
      if (v > best) { best = v; piv = i; }
    }
    if (piv != k) {
      for (int j = 0; j < n; j++) std::swap(lu(k, j), lu(piv, j));
      std::swap(perm[k], perm[piv]);
    }
    double d = lu(k, k);
    if (d != 0)
      for (int i = k + 1; i < n; i++) lu(i, k) /= d;
    for (int i = k + 1; i < n; i++) {
      double l = lu(i, k);
      if (l == 0) continue;
      for (int j = k + 1; j < n; j++) lu(i, j) -= l * lu(k, j);
    }
  }
  Mat inv(n, n);
  Vec col(n);
  for (int c = 0; c < n; c++) {
    for (int i = 0; i < n; i++) col[i] = (perm[i] == c) ? 1.0 : 0.0;
    for (int i = 0; i < n; i++) {
      double s = col[i];
      for (int j = 0; j < i; j++) s -= lu(i, j) * col[j];
      col[i] = s;
    }
    for (int i = n - 1; i >= 0; i--) {
      double s = col[i];
      for (int j = i + 1; j < n; j++) s -= lu(i, j) * col[j];
      col[i] = s / lu(i, i);
    }
    for (int i = 0; i < n; i++) inv(i, c) = col[i];
  }
  return inv;
}

// utils.cc:256-261 GetConditionNumber = JacobiSVD sigma_max / sigma_min.  Restated as a
// one-sided (Hestenes) Jacobi SVD: the singular values are the column norms after convergence.
inline void singular_values(const Mat& A, Vec* sv) {
  int m = A.r, n = A.c;
  Mat U = A;
  const double eps = std::numeric_limits<double>::epsilon();
  for (int sweep = 0; sweep < 60; sweep++) {
    bool rotated = false;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < m; i++) {
          alpha += U(i, p) * U(i, p);
          beta += U(i, q) * U(i, q);
          gamma += U(i, p) * U(i, q);
        }
        if (gamma == 0 || std::fabs(gamma) <= eps * std::sqrt(alpha * beta)) continue;
        rotated = true;
        double zeta = (beta - alpha) / (2.0 * gamma);
        double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
        for (int i = 0; i < m; i++) {
          double up = U(i, p), uq = U(i, q);
          U(i, p) = cs * up - sn * uq;
          U(i, q) = sn * up + cs * uq;
        }
      }
    if (!rotated) break;
  }
  sv->assign(n, 0.0);
  for (int j = 0; j < n; j++) {
    double s = 0;
    for (int i = 0; i < m; i++) s += U(i, j) * U(i, j);
    (*sv)[j] = std::sqrt(s);
  }
  std::sort(sv->begin(), sv->end(), [](double a, double b) { return a > b; });
}
inline double condition_number(const Mat& A) {
  Vec sv;
  singular_values(A, &sv);
  if (sv.empty()) return 0;
  return sv.front() / sv.back();
}

}  // namespace orc
