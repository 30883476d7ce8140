// TEST INFRASTRUCTURE - see oracle.hpp. Stage 3 oracle: residuals, Jacobians and
// the 13x13 augmented information block per scan pair.
// Follows /root/reference/form/feature/factor.cpp:30-186 and
// /root/reference/form/optimization/gtsam.hpp:59-140.
#include "oracle.hpp"

#include <vector>

namespace form_oracle {

namespace {
inline void rot(const form::Mat3 &R, const double *v, double *o) {
  o[0] = (R[0] * v[0] + R[1] * v[1]) + R[2] * v[2];
  o[1] = (R[3] * v[0] + R[4] * v[1]) + R[5] * v[2];
  o[2] = (R[6] * v[0] + R[7] * v[1]) + R[8] * v[2];
}
inline void rot_t(const form::Mat3 &R, const double *v, double *o) {
  o[0] = (R[0] * v[0] + R[3] * v[1]) + R[6] * v[2];
  o[1] = (R[1] * v[0] + R[4] * v[1]) + R[7] * v[2];
  o[2] = (R[2] * v[0] + R[5] * v[1]) + R[8] * v[2];
}
} // namespace

// factor.cpp:30-80
void PlanePoint::evaluate(const Pose3 &Ti, const Pose3 &Tj, double *r, double *H1,
                          double *H2) const {
  const size_t n = num_constraints();
  for (size_t c = 0; c < n; ++c) {
    const double *pi = &p_i[3 * c], *ni = &n_i[3 * c], *pj = &p_j[3 * c];
    double w_ni[3], w_pi[3], w_pj[3], v[3];
    rot(Ti.R, ni, w_ni);                                  // :39
    rot(Ti.R, pi, w_pi);                                  // :40-41
    rot(Tj.R, pj, w_pj);                                  // :42-43
    for (int k = 0; k < 3; ++k) {
      w_pi[k] += Ti.t[k];
      w_pj[k] += Tj.t[k];
      v[k] = w_pj[k] - w_pi[k];                           // :44
    }
    r[c] = (w_ni[0] * v[0] + w_ni[1] * v[1]) + w_ni[2] * v[2]; // :45
    if (H1) {
      double RT_n[3], RT_v[3];
      rot_t(Ti.R, w_ni, RT_n);                            // :51
      rot_t(Ti.R, v, RT_v);                               // :52
      double *h = &H1[6 * c];
      h[0] = RT_n[1] * pi[2] - RT_n[2] * pi[1] - RT_v[1] * ni[2] + RT_v[2] * ni[1]; // :53-55
      h[1] = RT_n[2] * pi[0] - RT_n[0] * pi[2] - RT_v[2] * ni[0] + RT_v[0] * ni[2]; // :56-58
      h[2] = RT_n[0] * pi[1] - RT_n[1] * pi[0] - RT_v[0] * ni[1] + RT_v[1] * ni[0]; // :59-61
      h[3] = -RT_n[0];                                    // :62
      h[4] = -RT_n[1];
      h[5] = -RT_n[2];
    }
    if (H2) {
      double RT_n[3];
      rot_t(Tj.R, w_ni, RT_n);                            // :69
      double *h = &H2[6 * c];
      h[0] = -RT_n[1] * pj[2] + RT_n[2] * pj[1];          // :70-71
      h[1] = -RT_n[2] * pj[0] + RT_n[0] * pj[2];          // :72-73
      h[2] = -RT_n[0] * pj[1] + RT_n[1] * pj[0];          // :74-75
      h[3] = RT_n[0];                                     // :76
      h[4] = RT_n[1];
      h[5] = RT_n[2];
    }
  }
}

// factor.cpp:82-128
void PointPoint::evaluate(const Pose3 &Ti, const Pose3 &Tj, double *r, double *H1,
                          double *H2) const {
  const size_t m = num_constraints();
  for (size_t c = 0; c < m; ++c) {
    const double *pi = &p_i[3 * c], *pj = &p_j[3 * c];
    double w_pi[3], w_pj[3];
    rot(Ti.R, pi, w_pi);                                  // :90-91
    rot(Tj.R, pj, w_pj);                                  // :92-93
    for (int k = 0; k < 3; ++k) {
      w_pi[k] += Ti.t[k];
      w_pj[k] += Tj.t[k];
      r[3 * c + k] = w_pj[k] - w_pi[k];                   // :94-95 (column-major resize)
    }
    if (H1) {
      for (int row = 0; row < 3; ++row) {
        // Ri = -R_i (:101); temp = Ri.col(a)*p.row(b) - Ri.col(c)*p.row(d) (:103-108)
        const double R0 = Ti.R[3 * row + 0] * -1.0, R1 = Ti.R[3 * row + 1] * -1.0,
                     R2 = Ti.R[3 * row + 2] * -1.0;
        double *h = &H1[6 * (3 * c + row)];
        h[0] = R2 * pi[1] - R1 * pi[2];
        h[1] = R0 * pi[2] - R2 * pi[0];
        h[2] = R1 * pi[0] - R0 * pi[1];
        h[3] = R0;                                        // :109
        h[4] = R1;
        h[5] = R2;
      }
    }
    if (H2) {
      for (int row = 0; row < 3; ++row) {
        const double R0 = Tj.R[3 * row + 0], R1 = Tj.R[3 * row + 1], R2 = Tj.R[3 * row + 2];
        double *h = &H2[6 * (3 * c + row)];
        h[0] = R2 * pj[1] - R1 * pj[2];                   // :118-119
        h[1] = R0 * pj[2] - R2 * pj[0];                   // :120-121
        h[2] = R1 * pj[0] - R0 * pj[1];                   // :122-123
        h[3] = R0;                                        // :124
        h[4] = R1;
        h[5] = R2;
      }
    }
  }
}

// FeatureFactor::evaluateError (factor.cpp:141-186: planar rows then point rows),
// DenseFactor::linearize (gtsam.hpp:67-86), FastIsotropic (gtsam.hpp:121-139).
void linearize_pair(const PairConstraints &c, const Pose3 &Ti, const Pose3 &Tj, double sigma,
                    double out91[91]) {
  const size_t n = c.plane.num_residuals(), m3 = c.point.num_residuals();
  const size_t rows = n + m3;
  std::vector<double> r(rows), H1(6 * rows), H2(6 * rows);
  if (n) c.plane.evaluate(Ti, Tj, r.data(), H1.data(), H2.data());
  if (m3) c.point.evaluate(Ti, Tj, r.data() + n, H1.data() + 6 * n, H2.data() + 6 * n);
  const double invsigma = 1.0 / sigma; // gtsam.hpp:97
  double G[13][13] = {};
  for (size_t row = 0; row < rows; ++row) {
    double a[13];
    for (int k = 0; k < 6; ++k) {
      a[k] = H1[6 * row + k] * invsigma;     // WhitenSystem: A *= invsigma
      a[6 + k] = H2[6 * row + k] * invsigma;
    }
    a[12] = (-r[row]) * invsigma;            // b = -unwhitenedError, whitened
    for (int p = 0; p < 13; ++p)
      for (int q = p; q < 13; ++q) G[p][q] += a[p] * a[q]; // HessianFactor(JacobianFactor)
  }
  size_t e = 0;
  for (int p = 0; p < 13; ++p)
    for (int q = p; q < 13; ++q) out91[e++] = G[p][q];
}

double error_pair(const PairConstraints &c, const Pose3 &Ti, const Pose3 &Tj, double sigma) {
  const size_t n = c.plane.num_residuals(), m3 = c.point.num_residuals();
  std::vector<double> r(n + m3);
  if (n) c.plane.evaluate(Ti, Tj, r.data(), nullptr, nullptr);
  if (m3) c.point.evaluate(Ti, Tj, r.data() + n, nullptr, nullptr);
  const double invsigma = 1.0 / sigma;
  double s = 0.0;
  for (double v : r) {
    const double w = v * invsigma;
    s += w * w;
  }
  return 0.5 * s;
}

} // namespace form_oracle
