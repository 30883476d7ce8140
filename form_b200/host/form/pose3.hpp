// Minimal SE(3) pose standing in for gtsam::Pose3 on the host side.
//
// The reference passes gtsam::Pose3 through its API (form.hpp:79,
// matcher.hpp:48, constraints.hpp:118-159).  GTSAM is not available to this
// build, so this header restates the small part of the published GTSAM 4.3
// Pose3/Rot3 semantics the FORM host code relies on (SURVEY Appendix B,
// [external]):
//   * tangent order xi = [omega(3), v(3)] (rotation first),
//   * retract(xi) = T * Expmap(xi) with the full SE(3) exponential,
//   * localCoordinates(T2) = Logmap(T^-1 * T2),
//   * T * p = R p + t, with the dot-product order the oracle fixes
//     (SURVEY A.2): ((r0*x + r1*y) + r2*z) + t.
// Storage is row-major R[9] + t[3] = 96 bytes, identical to formgpu_pose in
// include/formgpu.h so a Pose3* can be handed to the C-ABI unchanged.
#pragma once

#include <array>
#include <cmath>
#include <cstddef>

namespace form {

using Vec3 = std::array<double, 3>;
using Vec6 = std::array<double, 6>;
using Mat3 = std::array<double, 9>;  // row-major
using Mat6 = std::array<double, 36>; // row-major

inline Mat3 mat3_identity() { return {1, 0, 0, 0, 1, 0, 0, 0, 1}; }

inline Mat3 mat3_mul(const Mat3 &A, const Mat3 &B) {
  Mat3 C{};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C[3 * r + c] = (A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
  return C;
}

inline Mat3 mat3_transpose(const Mat3 &A) {
  return {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
}

inline Vec3 mat3_vec(const Mat3 &A, const Vec3 &v) {
  return {(A[0] * v[0] + A[1] * v[1]) + A[2] * v[2],
          (A[3] * v[0] + A[4] * v[1]) + A[5] * v[2],
          (A[6] * v[0] + A[7] * v[1]) + A[8] * v[2]};
}

inline Mat3 skew(const Vec3 &w) {
  return {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
}

inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
  return {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2],
          a[0] * b[1] - a[1] * b[0]};
}

inline double dot(const Vec3 &a, const Vec3 &b) {
  return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2];
}

/// SO(3) exponential (Rodrigues).
inline Mat3 so3_expmap(const Vec3 &w) {
  const double th2 = dot(w, w);
  const Mat3 W = skew(w);
  const Mat3 WW = mat3_mul(W, W);
  double a, b;
  if (th2 > 1e-16) {
    const double th = std::sqrt(th2);
    a = std::sin(th) / th;
    b = (1.0 - std::cos(th)) / th2;
  } else {
    a = 1.0 - th2 / 6.0;
    b = 0.5 - th2 / 24.0;
  }
  Mat3 R = mat3_identity();
  for (int k = 0; k < 9; ++k) R[k] += a * W[k] + b * WW[k];
  return R;
}

/// SO(3) logarithm, robust near 0 and near pi.
inline Vec3 so3_logmap(const Mat3 &R) {
  const double tr = R[0] + R[4] + R[8];
  Vec3 w;
  if (tr + 1.0 < 1e-3) {
    // angle close to pi: pick the largest diagonal element
    // (same construction GTSAM's Rot3::Logmap uses, [external])
    const double R11 = R[0], R12 = R[1], R13 = R[2];
    const double R21 = R[3], R22 = R[4], R23 = R[5];
    const double R31 = R[6], R32 = R[7], R33 = R[8];
    const double PI = 3.14159265358979323846;
    if (R33 > R22 && R33 > R11) {
      const double W = R21 - R12, Q1 = 2.0 + 2.0 * R33, Q2 = R31 + R13, Q3 = R23 + R32;
      const double r = std::sqrt(Q1);
      const double one_over_r = 1 / r;
      const double norm = std::sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w = {sgn_w * scale * Q2, sgn_w * scale * Q3, sgn_w * scale * Q1};
    } else if (R22 > R11) {
      const double W = R13 - R31, Q1 = 2.0 + 2.0 * R22, Q2 = R23 + R32, Q3 = R12 + R21;
      const double r = std::sqrt(Q1);
      const double one_over_r = 1 / r;
      const double norm = std::sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w = {sgn_w * scale * Q3, sgn_w * scale * Q1, sgn_w * scale * Q2};
    } else {
      const double W = R32 - R23, Q1 = 2.0 + 2.0 * R11, Q2 = R12 + R21, Q3 = R31 + R13;
      const double r = std::sqrt(Q1);
      const double one_over_r = 1 / r;
      const double norm = std::sqrt(Q1 * Q1 + Q2 * Q2 + Q3 * Q3 + W * W);
      const double sgn_w = W < 0 ? -1.0 : 1.0;
      const double mag = PI - (2 * sgn_w * W) / norm;
      const double scale = 0.5 * one_over_r * mag;
      w = {sgn_w * scale * Q1, sgn_w * scale * Q2, sgn_w * scale * Q3};
    }
    return w;
  }
  double magnitude;
  const double tr_3 = tr - 3.0;
  if (tr_3 < -1e-6) {
    const double c = std::fmin(1.0, std::fmax(-1.0, (tr - 1.0) / 2.0));
    const double theta = std::acos(c);
    magnitude = theta / (2.0 * std::sin(theta));
  } else {
    // theta close to 0: Taylor expansion of theta / (2 sin theta)
    magnitude = 0.5 - tr_3 / 12.0 + tr_3 * tr_3 / 60.0;
  }
  w = {magnitude * (R[7] - R[5]), magnitude * (R[2] - R[6]), magnitude * (R[3] - R[1])};
  return w;
}

/// Inverse right Jacobian of SO(3) at omega (GTSAM Rot3::LogmapDerivative).
inline Mat3 so3_logmap_derivative(const Vec3 &w) {
  const double th2 = dot(w, w);
  const Mat3 W = skew(w);
  const Mat3 WW = mat3_mul(W, W);
  Mat3 J = mat3_identity();
  double c;
  if (th2 > 1e-10) {
    const double th = std::sqrt(th2);
    c = 1.0 / th2 - (1.0 + std::cos(th)) / (2.0 * th * std::sin(th));
  } else {
    c = 1.0 / 12.0 + th2 / 720.0;
  }
  for (int k = 0; k < 9; ++k) J[k] += 0.5 * W[k] + c * WW[k];
  return J;
}

struct Pose3 {
  Mat3 R{1, 0, 0, 0, 1, 0, 0, 0, 1};
  Vec3 t{0, 0, 0};

  Pose3() = default;
  Pose3(const Mat3 &R_, const Vec3 &t_) : R(R_), t(t_) {}

  static Pose3 Identity() { return Pose3(); }

  /// Rot3::RzRyRx(x, y, z) * translation, used by the reference's test seeds
  /// (tests/test_SeparateFactor.cpp:26-27).
  static Pose3 RzRyRx(double rx, double ry, double rz, const Vec3 &t_) {
    const double cx = std::cos(rx), sx = std::sin(rx);
    const double cy = std::cos(ry), sy = std::sin(ry);
    const double cz = std::cos(rz), sz = std::sin(rz);
    const Mat3 Rx{1, 0, 0, 0, cx, -sx, 0, sx, cx};
    const Mat3 Ry{cy, 0, sy, 0, 1, 0, -sy, 0, cy};
    const Mat3 Rz{cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    return Pose3(mat3_mul(Rz, mat3_mul(Ry, Rx)), t_);
  }

  const Mat3 &rotation() const { return R; }
  const Vec3 &translation() const { return t; }

  /// R p + t (gtsam::Pose3::transformFrom).
  Vec3 transformFrom(const Vec3 &p) const {
    const Vec3 q = mat3_vec(R, p);
    return {q[0] + t[0], q[1] + t[1], q[2] + t[2]};
  }
  Vec3 rotate(const Vec3 &n) const { return mat3_vec(R, n); }
  Vec3 operator*(const Vec3 &p) const { return transformFrom(p); }

  Pose3 operator*(const Pose3 &o) const {
    const Vec3 rt = mat3_vec(R, o.t);
    return Pose3(mat3_mul(R, o.R), {rt[0] + t[0], rt[1] + t[1], rt[2] + t[2]});
  }

  Pose3 inverse() const {
    const Mat3 Rt = mat3_transpose(R);
    const Vec3 v = mat3_vec(Rt, t);
    return Pose3(Rt, {-v[0], -v[1], -v[2]});
  }

  /// Full SE(3) exponential, xi = [omega, v].
  static Pose3 Expmap(const Vec6 &xi) {
    const Vec3 w{xi[0], xi[1], xi[2]}, v{xi[3], xi[4], xi[5]};
    const Mat3 R = so3_expmap(w);
    const double th2 = dot(w, w);
    if (th2 > 1e-20) {
      const double wv = dot(w, v);
      const Vec3 wxv = cross(w, v);
      const Vec3 Rwxv = mat3_vec(R, wxv);
      Vec3 t;
      for (int k = 0; k < 3; ++k) t[k] = (wxv[k] - Rwxv[k] + w[k] * wv) / th2;
      return Pose3(R, t);
    }
    return Pose3(R, v);
  }

  /// Full SE(3) logarithm, returns [omega, v].
  static Vec6 Logmap(const Pose3 &p) {
    const Vec3 w = so3_logmap(p.R);
    const double th = std::sqrt(dot(w, w));
    if (th < 1e-10) return {w[0], w[1], w[2], p.t[0], p.t[1], p.t[2]};
    const Vec3 k{w[0] / th, w[1] / th, w[2] / th};
    const Vec3 WT = cross(k, p.t);
    const Vec3 WWT = cross(k, WT);
    const double Tan = std::tan(0.5 * th);
    const double a = 0.5 * th, b = 1.0 - th / (2.0 * Tan);
    return {w[0], w[1], w[2], p.t[0] - a * WT[0] + b * WWT[0],
            p.t[1] - a * WT[1] + b * WWT[1], p.t[2] - a * WT[2] + b * WWT[2]};
  }

  Pose3 retract(const Vec6 &xi) const { return (*this) * Expmap(xi); }
  Vec6 localCoordinates(const Pose3 &o) const { return Logmap(inverse() * o); }

  /// d Logmap(T Exp(xi)) / d xi at xi = 0 (inverse right Jacobian of SE(3)),
  /// row-major 6x6.  Barfoot's Q-matrix form, as in GTSAM's
  /// Pose3::LogmapDerivative [external]; checked against finite differences in
  /// tests/test_host_math.py.
  static Mat6 LogmapDerivative(const Pose3 &p) {
    const Vec6 xi = Logmap(p);
    const Vec3 w{xi[0], xi[1], xi[2]}, v{xi[3], xi[4], xi[5]};
    const Mat3 Jw = so3_logmap_derivative(w);
    const Mat3 V = skew(v), W = skew(w);
    const Mat3 WV = mat3_mul(W, V), VW = mat3_mul(V, W);
    const Mat3 WVW = mat3_mul(WV, W);
    const Mat3 WWV = mat3_mul(W, WV), VWW = mat3_mul(VW, W);
    const Mat3 WVWW = mat3_mul(WVW, W), WWVW = mat3_mul(W, WVW);
    const double phi2 = dot(w, w), phi = std::sqrt(phi2);
    double c1, c2, c3;
    if (phi > 1e-5) {
      const double s = std::sin(phi), c = std::cos(phi);
      const double phi3 = phi2 * phi, phi4 = phi2 * phi2, phi5 = phi4 * phi;
      c1 = (phi - s) / phi3;
      c2 = (1.0 - phi2 / 2.0 - c) / phi4;
      c3 = 0.5 * (c2 - 3.0 * (phi - s - phi3 / 6.0) / phi5);
    } else {
      c1 = 1.0 / 6.0;
      c2 = -1.0 / 24.0;
      c3 = -1.0 / 120.0;
    }
    Mat3 Q;
    for (int k = 0; k < 9; ++k)
      Q[k] = -0.5 * V[k] + c1 * (WV[k] + VW[k] - WVW[k]) +
             c2 * (WWV[k] + VWW[k] - 3.0 * WVW[k]) - c3 * (WVWW[k] + WWVW[k]);
    const Mat3 JQJ = mat3_mul(Jw, mat3_mul(Q, Jw));
    Mat6 J{};
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) {
        J[6 * r + c] = Jw[3 * r + c];
        J[6 * (r + 3) + (c + 3)] = Jw[3 * r + c];
        J[6 * (r + 3) + c] = -JQJ[3 * r + c];
      }
    return J;
  }

  /// Re-orthonormalise the rotation the way gtsam::Rot3::normalized() does
  /// (first-order symmetric orthogonalisation, [external]); used by the
  /// constant-velocity prediction (constraints.cpp:88).
  Pose3 normalized() const {
    const Vec3 x{R[0], R[3], R[6]}, y{R[1], R[4], R[7]};
    const double err = dot(x, y);
    Vec3 xo, yo;
    for (int k = 0; k < 3; ++k) {
      xo[k] = x[k] - (err / 2) * y[k];
      yo[k] = y[k] - (err / 2) * x[k];
    }
    const Vec3 zo = cross(xo, yo);
    const double sx = 0.5 * (3 - dot(xo, xo)), sy = 0.5 * (3 - dot(yo, yo)),
                 sz = 0.5 * (3 - dot(zo, zo));
    Mat3 Rn;
    for (int k = 0; k < 3; ++k) {
      Rn[3 * k + 0] = sx * xo[k];
      Rn[3 * k + 1] = sy * yo[k];
      Rn[3 * k + 2] = sz * zo[k];
    }
    return Pose3(Rn, t);
  }
};
static_assert(sizeof(Pose3) == 96, "Pose3 must match formgpu_pose (96 bytes)");

inline double norm6(const Vec6 &v) {
  double s = 0;
  for (double e : v) s += e * e;
  return std::sqrt(s);
}

} // namespace form
