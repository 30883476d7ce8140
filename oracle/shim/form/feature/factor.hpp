// TEST INFRASTRUCTURE.  Shadows FORM's form/feature/factor.hpp on the include path of the
// reference build in oracle/ref/: Matcher::match (form/optimization/matcher.hpp) only needs
// the correspondence containers' push_back / clear, while the real header pulls in GTSAM's
// factor machinery (stage 3, tolerance class, pinned by finite differences instead).
#pragma once

#include "form/feature/features.hpp"

#include <memory>
#include <vector>

namespace form {

struct PlanePoint {
  typedef std::shared_ptr<PlanePoint> Ptr;
  std::vector<PlanarFeat> map_points, keypoints; // (point of scan i, keypoint of scan j) per match
  void push_back(const PlanarFeat &p_i, const PlanarFeat &p_j) {
    map_points.push_back(p_i);
    keypoints.push_back(p_j);
  }
  void clear() noexcept {
    map_points.clear();
    keypoints.clear();
  }
};

struct PointPoint {
  typedef std::shared_ptr<PointPoint> Ptr;
  std::vector<PointFeat> map_points, keypoints;
  void push_back(const PointFeat &p_i, const PointFeat &p_j) {
    map_points.push_back(p_i);
    keypoints.push_back(p_j);
  }
  void clear() noexcept {
    map_points.clear();
    keypoints.clear();
  }
};

} // namespace form
