// TEST INFRASTRUCTURE - see oracle.hpp.  The oracle behind the HotPath seam, so
// tests can run the shared host logic (Estimator, smoother) over the CPU
// restatement and compare with the CUDA path call by call.
#pragma once

#include "form/hotpath.hpp"
#include "oracle.hpp"

#include <map>

namespace form_oracle {

class OracleHotPath : public form::HotPath {
public:
  explicit OracleHotPath(const form::HotPathParams &p);

  void extract(const PointXYZf *scan, size_t n, uint64_t scan_idx,
               std::vector<PlanarFeat> &planar, std::vector<PointFeat> &point) override;
  void map_rebuild(const form::ScanPose *poses, size_t n_poses) override;
  void associate(const Pose3 &pose_k, std::vector<form::PairCount> &counts) override;
  void linearize(const form::PairKey *pairs, size_t n_pairs, const form::ScanPose *poses,
                 size_t n_poses, double *out91) override;
  void error(const form::PairKey *pairs, size_t n_pairs, const form::ScanPose *poses,
             size_t n_poses, double *out) override;
  void commit_scan(size_t &n_planar_added, size_t &n_point_added) override;
  void remove_scans(const uint64_t *scans, size_t n) override;
  void world_keypoints(const form::ScanPose *poses, size_t n_poses,
                       std::vector<PlanarFeat> &planar, std::vector<PointFeat> &point) override;

  // ---- introspection for parity tests ----
  form::HotPathParams params;
  ExtractParams extract_params;
  ExtractResult last_extract;
  uint64_t current_scan = 0;
  std::vector<PlanarFeat> cur_planar;
  std::vector<PointFeat> cur_point;
  KeypointMap<PlanarFeat> planar_map;
  KeypointMap<PointFeat> point_map;
  std::map<uint64_t, Pose3> map_poses;
  // Matcher::matches (matcher.hpp:52): persists across calls, see A.3-12
  std::vector<PlanarFeat> planar_match_queries;
  std::vector<PointFeat> point_match_queries;
  std::vector<MatchResult<PlanarFeat>> planar_matches;
  std::vector<MatchResult<PointFeat>> point_matches;
  // m_constraints[j][i] (constraints.hpp:91-99), ordered => rule R7
  std::map<uint64_t, std::map<uint64_t, PairConstraints>> constraints;

private:
  int threads() const;
};

} // namespace form_oracle
