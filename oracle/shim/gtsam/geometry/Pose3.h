// TEST INFRASTRUCTURE - not GTSAM.  gtsam::Pose3 / Rot3 as far as FORM's sources use them
// (transform of a point, rotation of a normal, inverse, rotation().matrix() / transpose(),
// translation()).  Arithmetic [external]: T * p = R p + t with row-wise dot products
// ((r0 x + r1 y) + r2 z) + t; inverse = (R^T, -(R^T t)), same order.
// For FORM's smoother (constraints.cpp, form.cpp) also compose, retract / localCoordinates,
// Logmap and its derivative, Rot3::normalized: the SE(3) formulas of GTSAM [external] as
// restated in form_b200/host/form/pose3.hpp (checked against scipy's expm / logm and finite
// differences in tests/test_host_math.py), included here under another namespace name because
// the reference's own code lives in namespace form.
#pragma once

#include <Eigen/Dense>

#include <array>
#include <cmath>
#include <cstddef>
#define form formhostmath
#include "form/pose3.hpp" // form_b200/host/form/pose3.hpp (the reference tree has no file of this name)
#undef form

namespace gtsam {

using Vector3 = Eigen::Vector3d;
using Point3 = Eigen::Vector3d;
using Key = unsigned long long;
using Vector = Eigen::VectorXd;
using Matrix = Eigen::MatrixXd;
using Vector6 = Eigen::Matrix<double, 6, 1>;

class Rot3 {
public:
  Rot3() : m_r{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  explicit Rot3(const double r[9]) {
    for (int i = 0; i < 9; ++i) m_r[i] = r[i];
  }
  const double *data() const { return m_r; } // row-major
  /// Rot3::normalized(): first-order re-orthonormalisation
  Rot3 normalized() const {
    formhostmath::Pose3 p;
    for (int i = 0; i < 9; ++i) p.R[i] = m_r[i];
    const formhostmath::Pose3 n = p.normalized();
    return Rot3(n.R.data());
  }
  Eigen::Matrix3d matrix() const {
    Eigen::Matrix3d m;
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) m(r, c) = m_r[3 * r + c];
    return m;
  }
  Eigen::Matrix3d transpose() const { return matrix().transpose(); }
  template <typename V> Eigen::Vector3d operator*(const V &v) const {
    const double x = v(0), y = v(1), z = v(2);
    return Eigen::Vector3d((m_r[0] * x + m_r[1] * y) + m_r[2] * z, (m_r[3] * x + m_r[4] * y) + m_r[5] * z,
                           (m_r[6] * x + m_r[7] * y) + m_r[8] * z);
  }

private:
  double m_r[9];
};

class Pose3 {
public:
  Pose3() : m_t{0, 0, 0} {}
  Pose3(const double r[9], const double t[3]) : m_R(r), m_t{t[0], t[1], t[2]} {}
  Pose3(const Rot3 &R, const Point3 &t) : m_R(R), m_t{t(0), t(1), t(2)} {}
  static Pose3 Identity() { return Pose3(); }
  formhostmath::Pose3 host() const {
    formhostmath::Pose3 p;
    for (int i = 0; i < 9; ++i) p.R[i] = m_R.data()[i];
    for (int i = 0; i < 3; ++i) p.t[i] = m_t[i];
    return p;
  }
  static Pose3 from_host(const formhostmath::Pose3 &p) { return Pose3(p.R.data(), p.t.data()); }
  Pose3 operator*(const Pose3 &o) const { return from_host(host() * o.host()); }
  Pose3 retract(const Vector6 &xi) const {
    return from_host(host().retract({xi(0), xi(1), xi(2), xi(3), xi(4), xi(5)}));
  }
  Vector6 localCoordinates(const Pose3 &o) const {
    const formhostmath::Vec6 l = host().localCoordinates(o.host());
    Vector6 out;
    for (int a = 0; a < 6; ++a) out(a) = l[a];
    return out;
  }
  const Rot3 &rotation() const { return m_R; }
  Eigen::Vector3d translation() const { return Eigen::Vector3d(m_t[0], m_t[1], m_t[2]); }
  template <typename V> Eigen::Vector3d operator*(const V &v) const {
    const Eigen::Vector3d r = m_R * v;
    return Eigen::Vector3d(r(0) + m_t[0], r(1) + m_t[1], r(2) + m_t[2]);
  }
  Pose3 inverse() const {
    const double *R = m_R.data();
    const double rt[9] = {R[0], R[3], R[6], R[1], R[4], R[7], R[2], R[5], R[8]};
    const double t[3] = {-((rt[0] * m_t[0] + rt[1] * m_t[1]) + rt[2] * m_t[2]),
                         -((rt[3] * m_t[0] + rt[4] * m_t[1]) + rt[5] * m_t[2]),
                         -((rt[6] * m_t[0] + rt[7] * m_t[1]) + rt[8] * m_t[2])};
    return Pose3(rt, t);
  }

private:
  Rot3 m_R;
  double m_t[3];
};

} // namespace gtsam
