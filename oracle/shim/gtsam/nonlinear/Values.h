// TEST INFRASTRUCTURE - not GTSAM.  gtsam::Values as an ordered key -> Pose3 table (every
// variable of FORM's graphs is a Pose3), gtsam::VectorValues as key -> 6-vector.
#pragma once
#include <gtsam/geometry/Pose3.h>
#include <array>
#include <map>
#include <vector>
namespace gtsam {
class VectorValues {
public:
  std::map<Key, std::array<double, 6>> v;
};
class Values {
public:
  void insert(Key key, const Pose3 &pose) { m_poses[key] = pose; }
  void update(Key key, const Pose3 &pose) { m_poses.at(key) = pose; }
  void update(const Values &o) {
    for (const auto &kv : o.m_poses) m_poses.at(kv.first) = kv.second;
  }
  void erase(Key key) { m_poses.erase(key); }
  bool exists(Key key) const { return m_poses.count(key) != 0; }
  size_t size() const { return m_poses.size(); }
  template <typename T> const T &at(Key key) const { return m_poses.at(key); }
  std::vector<Key> keys() const {
    std::vector<Key> k;
    for (const auto &kv : m_poses) k.push_back(kv.first);
    return k;
  }
  /// every pose moved by its tangent vector: T . Expmap(delta)
  Values retract(const VectorValues &delta) const {
    Values out = *this;
    for (const auto &kv : delta.v) {
      Vector6 d;
      for (int a = 0; a < 6; ++a) d(a) = kv.second[a];
      out.m_poses.at(kv.first) = m_poses.at(kv.first).retract(d);
    }
    return out;
  }
  std::map<Key, Pose3>::const_iterator begin() const { return m_poses.begin(); }
  std::map<Key, Pose3>::const_iterator end() const { return m_poses.end(); }

private:
  std::map<Key, Pose3> m_poses;
};
} // namespace gtsam
