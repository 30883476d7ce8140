// TEST INFRASTRUCTURE - not GTSAM.  gtsam::Values as a key -> Pose3 table.
#pragma once
#include <gtsam/geometry/Pose3.h>
#include <map>
namespace gtsam {
class Values {
public:
  void insert(unsigned long long key, const Pose3 &pose) { m_poses[key] = pose; }
  template <typename T> const T &at(unsigned long long key) const { return m_poses.at(key); }

private:
  std::map<unsigned long long, Pose3> m_poses;
};
} // namespace gtsam
