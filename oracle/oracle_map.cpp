// TEST INFRASTRUCTURE - see oracle.hpp. Stage 2 oracle: voxel map + NN search.
// Follows /root/reference/form/mapping/map.tpp:34-165 and map.hpp:37-61.
#include "oracle.hpp"

#include <cmath>

namespace form_oracle {

// map.tpp:54-68, same order
const int kVoxelShifts[27][3] = {
    {0, 0, 0},   {1, 0, 0},   {-1, 0, 0},  {0, 1, 0},   {0, -1, 0},  {0, 0, 1},   {0, 0, -1},
    {1, 1, 0},   {1, -1, 0},  {-1, 1, 0},  {-1, -1, 0}, {1, 0, 1},   {1, 0, -1},  {-1, 0, 1},
    {-1, 0, -1}, {0, 1, 1},   {0, 1, -1},  {0, -1, 1},  {0, -1, -1}, {1, 1, 1},   {1, 1, -1},
    {1, -1, 1},  {1, -1, -1}, {-1, 1, 1},  {-1, 1, -1}, {-1, -1, 1}, {-1, -1, -1}};

// map.tpp:34-38: (p.array() / width).floor().cast<int>()
VoxelKey compute_coords(double x, double y, double z, double w) {
  return {(int32_t)std::floor(x / w), (int32_t)std::floor(y / w), (int32_t)std::floor(z / w)};
}

// features.hpp:63-65: vec3() = pose * vec3()  (R p + t, A.2 order)
PointFeat transform(const PointFeat &p, const Pose3 &T) {
  const form::Vec3 q = T.transformFrom({p.x, p.y, p.z});
  PointFeat o = p;
  o.x = q[0]; o.y = q[1]; o.z = q[2];
  return o;
}

// features.hpp:137-140: point as above, normal rotated only
PlanarFeat transform(const PlanarFeat &p, const Pose3 &T) {
  const form::Vec3 q = T.transformFrom({p.x, p.y, p.z});
  const form::Vec3 n = T.rotate({p.nx, p.ny, p.nz});
  PlanarFeat o = p;
  o.x = q[0]; o.y = q[1]; o.z = q[2];
  o.nx = n[0]; o.ny = n[1]; o.nz = n[2];
  return o;
}

template <typename Feat>
void KeypointMap<Feat>::to_voxel_map(const std::map<uint64_t, Pose3> &poses, double voxel_width) {
  voxel_width_ = voxel_width;
  voxels_.clear();
  for (const auto &[scan, kps] : scans_) { // rule R4: scan ascending
    const Pose3 &T = poses.at(scan);       // map.tpp:137
    for (size_t k = 0; k < kps.size(); ++k) {
      MapPoint<Feat> mp{transform(kps[k], T), scan, (uint32_t)k}; // :141
      const VoxelKey key = compute_coords(mp.world.x, mp.world.y, mp.world.z, voxel_width_);
      voxels_[key].push_back(mp); // push_back, :40-52
    }
  }
}

template <typename Feat>
MatchResult<Feat> KeypointMap<Feat>::find_closest(const Feat &q) const {
  const VoxelKey c = compute_coords(q.x, q.y, q.z, voxel_width_); // :73
  MatchResult<Feat> res;
  Feat best_world{};
  for (const auto &s : kVoxelShifts) { // :77
    const VoxelKey key{c.x + s[0], c.y + s[1], c.z + s[2]};
    auto it = voxels_.find(key);
    if (it == voxels_.end()) continue;
    for (const auto &mp : it->second) { // :81-87
      // 4-lane double squared norm, lane 3 is the zero padding (A.2)
      const double d0 = mp.world.x - q.x, d1 = mp.world.y - q.y, d2 = mp.world.z - q.z;
      const double d3 = mp.world._ - q._;
      const double dist = (d0 * d0 + d2 * d2) + (d1 * d1 + d3 * d3);
      if (dist < res.dist_sqrd) { // strict: first in (shift, scan, k) order wins = R5
        res.dist_sqrd = dist;
        res.found = true;
        res.scan = mp.scan;
        res.k = mp.k;
        best_world = mp.world;
      }
    }
  }
  res.point_local = best_world; // caller moves it back to the scan frame
  return res;
}

template <typename Feat>
size_t KeypointMap<Feat>::insert_matches(uint64_t scan, const std::vector<Feat> &queries,
                                         const std::vector<MatchResult<Feat>> &matches) {
  if (matches.empty()) return 0; // map.tpp:152-154
  auto &kps = get(scan);         // :156
  const double thr = min_dist_map * min_dist_map; // :158
  size_t added = 0;
  for (size_t j = 0; j < matches.size(); ++j) {
    if (matches[j].dist_sqrd > thr) { // :161 strict; unfound = DBL_MAX inserts
      kps.push_back(queries[j]);
      ++added;
    }
  }
  return added;
}

template class KeypointMap<PlanarFeat>;
template class KeypointMap<PointFeat>;

} // namespace form_oracle
