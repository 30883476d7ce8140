// TEST INFRASTRUCTURE - not GTSAM.
#pragma once
#include <gtsam/geometry/Pose3.h>
#include <cmath>
namespace gtsam {
inline bool fpEqual(double a, double b, double tol) { return std::fabs(a - b) <= tol; }
} // namespace gtsam
