// TEST INFRASTRUCTURE - not GTSAM.
#pragma once
#include <gtsam/nonlinear/LevenbergMarquardtOptimizer.h>
