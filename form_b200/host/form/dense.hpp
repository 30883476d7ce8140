// Small dense linear algebra for the host-side fixed-lag smoother
// (<= 61 poses => <= 366 x 366 systems).  Row-major std::vector<double>.
//
// Stands in for the GTSAM pieces the reference's smoother uses
// (/root/reference/form/optimization/gtsam.hpp:40-54 optimizeDensely,
// constraints.cpp:163-168 eliminatePartialMultifrontal); GTSAM is not
// available to this build (SURVEY 8c, Appendix B [external]).
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

namespace form {
namespace dense {

/// In-place Cholesky A = L L^T of the n x n SPD matrix stored row-major in `a`
/// (lower triangle written).  Returns false if a pivot is not positive.
inline bool cholesky(std::vector<double> &a, size_t n) {
  for (size_t j = 0; j < n; ++j) {
    double *aj = &a[j * n];
    double d = aj[j];
    for (size_t k = 0; k < j; ++k) d -= aj[k] * aj[k];
    if (!(d > 0.0) || !std::isfinite(d)) return false;
    const double ljj = std::sqrt(d);
    aj[j] = ljj;
    const double inv = 1.0 / ljj;
    for (size_t i = j + 1; i < n; ++i) {
      double *ai = &a[i * n];
      double s = ai[j];
      for (size_t k = 0; k < j; ++k) s -= ai[k] * aj[k];
      ai[j] = s * inv;
    }
  }
  return true;
}

/// Solve L L^T x = b in place (b -> x) given the factor from cholesky().
inline void cholesky_solve(const std::vector<double> &l, size_t n, double *b) {
  for (size_t i = 0; i < n; ++i) {
    double s = b[i];
    const double *li = &l[i * n];
    for (size_t k = 0; k < i; ++k) s -= li[k] * b[k];
    b[i] = s / li[i];
  }
  for (size_t ii = n; ii-- > 0;) {
    double s = b[ii];
    for (size_t k = ii + 1; k < n; ++k) s -= l[k * n + ii] * b[k];
    b[ii] = s / l[ii * n + ii];
  }
}

/// Dense quadratic 0.5 * (f - 2 g^T d + d^T G d) in information form.
struct Quadratic {
  size_t n = 0;
  std::vector<double> G; // n x n, symmetric, row-major
  std::vector<double> g; // n
  double f = 0.0;

  void resize(size_t n_) {
    n = n_;
    G.assign(n * n, 0.0);
    g.assign(n, 0.0);
    f = 0.0;
  }
  /// 0.5 * (f - 2 g.d + d.G.d)
  double error(const double *d) const {
    double gd = 0.0, dGd = 0.0;
    for (size_t r = 0; r < n; ++r) {
      gd += g[r] * d[r];
      double row = 0.0;
      const double *Gr = &G[r * n];
      for (size_t c = 0; c < n; ++c) row += Gr[c] * d[c];
      dGd += d[r] * row;
    }
    return 0.5 * (f - 2.0 * gd + dGd);
  }
};

} // namespace dense
} // namespace form
