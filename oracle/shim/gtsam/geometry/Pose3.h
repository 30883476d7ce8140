// TEST INFRASTRUCTURE - not GTSAM.  gtsam::Pose3 / Rot3 as far as FORM's sources use them
// (transform of a point, rotation of a normal, inverse, rotation().matrix() / transpose(),
// translation()).  Arithmetic [external]: T * p = R p + t with row-wise dot products
// ((r0 x + r1 y) + r2 z) + t; inverse = (R^T, -(R^T t)), same order.
#pragma once

#include <Eigen/Dense>

namespace gtsam {

using Vector3 = Eigen::Vector3d;
using Point3 = Eigen::Vector3d;
using Key = unsigned long long;
using Vector = Eigen::VectorXd;
using Matrix = Eigen::MatrixXd;

class Rot3 {
public:
  Rot3() : m_r{1, 0, 0, 0, 1, 0, 0, 0, 1} {}
  explicit Rot3(const double r[9]) {
    for (int i = 0; i < 9; ++i) m_r[i] = r[i];
  }
  const double *data() const { return m_r; } // row-major
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
