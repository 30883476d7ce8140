// TEST INFRASTRUCTURE - not GTSAM.  form/form.cpp names gtsam::Velocity3 in a using-declaration.
#pragma once
#include <gtsam/geometry/Pose3.h>
namespace gtsam {
using Velocity3 = Vector3;
} // namespace gtsam
