// TEST INFRASTRUCTURE - not GTSAM.  Only what makes form/optimization/gtsam.hpp's
// DenseLMOptimizer declaration compile; the smoother is not part of the reference build.
#pragma once
#include <gtsam/linear/NoiseModel.h>
namespace gtsam {
class VectorValues {};
class NonlinearFactorGraph {};
class NonlinearOptimizerParams {};
class LevenbergMarquardtParams : public NonlinearOptimizerParams {};
class GaussianFactorGraph {
public:
  VectorValues optimizeDensely() const { return VectorValues(); }
};
class LevenbergMarquardtOptimizer {
public:
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph &, const Values &, const LevenbergMarquardtParams &) {}
  virtual ~LevenbergMarquardtOptimizer() = default;
  virtual VectorValues solve(const GaussianFactorGraph &gfg, const NonlinearOptimizerParams &params) const = 0;
};
} // namespace gtsam
