// =============================================================================
// TEST INFRASTRUCTURE - NOT PRODUCT CODE.
//
// CPU oracle for the FORM per-scan hot path: a dependency-free C++17
// restatement of the reference's algorithm (feature extraction, voxel-hash
// association, residual/Jacobian linearisation).  Only tests/, bench.py's
// cpu_baseline / --impl reference legs and __graft_entry__.smoke() may link or
// load anything under oracle/.  The product (form_b200/) never does.
//
// PARITY UNPINNED: the reference ships no golden vectors or runnable tests for
// this path (its only test, tests/test_SeparateFactor.cpp, is stale and does
// not compile against the current headers) and cannot be built here (Eigen3,
// GTSAM, oneTBB, tsl::robin_map absent; no network).  The oracle is therefore
// pinned by (i) an independent numpy restatement, (ii) finite-difference
// Jacobian checks at the stale test's seeds, (iii) hand-computed voxel keys,
// (iv) explicit A^T A in numpy, and (v) frozen golden files under
// tests/golden/ (see tests/test_oracle_*.py).
//
// Every function cites the reference file:line it follows.  Where the
// reference is non-deterministic (unstable sort, concurrent_vector order,
// robin_map iteration) the canonical rules R1-R7 of SURVEY.md Appendix A.1
// apply; floating-point evaluation order follows Appendix A.2.  Compile with
// -ffp-contract=off so no FMA contraction changes index-determining results.
// =============================================================================
#pragma once

#include "form/pose3.hpp"
#include "form/types.hpp"

#include <cstddef>
#include <cstdint>
#include <limits>
#include <map>
#include <unordered_map>
#include <vector>

namespace form_oracle {

using form::PlanarFeat;
using form::PointFeat;
using form::PointXYZf;
using form::Pose3;

// ---------------------------------------------------------------------------
// Stage 1: feature extraction  (form/feature/extraction.{hpp,tpp})
// ---------------------------------------------------------------------------

/// FeatureExtractor::Params, extraction.hpp:59-88 (same defaults).
struct ExtractParams {
  size_t neighbor_points = 5;
  size_t num_sectors = 6;
  double planar_threshold = 1.0;
  size_t planar_feats_per_sector = 50;
  size_t point_feats_per_sector = 3;
  double radius = 1.0;
  size_t min_points = 5;
  double min_norm_squared = 1.0;
  double max_norm_squared = 100.0 * 100.0;
  int num_columns = 1024;
  int num_rows = 64;
};

/// Everything extract() computes, including the intermediates parity tests
/// compare bit-for-bit.
struct ExtractResult {
  std::vector<uint8_t> valid_mask;       // compute_valid_points
  std::vector<uint8_t> point_valid_mask; // compute_point_valid_points
  std::vector<float> curvature;          // per point, FLT_MAX when invalid
  std::vector<uint32_t> planar_indices;  // selection order, before normal drop
  std::vector<uint8_t> planar_keep;      // normal succeeded (same order)
  std::vector<int32_t> closest_prev;     // find_closest result (-1 none), same order
  std::vector<int32_t> closest_next;
  std::vector<uint32_t> point_indices;   // selection order
  std::vector<PlanarFeat> planar;        // final, rule R3 order
  std::vector<PointFeat> point;          // final, rule R3 order
};

/// FeatureExtractor::extract, extraction.tpp:29-132.  Returns false when the
/// scan size does not match rows*cols (the reference throws, :141-145).
bool extract(const ExtractParams &params, const PointXYZf *scan, size_t n,
             size_t scan_idx, int num_threads, ExtractResult &out);

/// Eigen::SelfAdjointEigenSolver<Matrix3f> restated (extraction.tpp:323-326):
/// eigen-decomposition of the symmetric 3x3 `cov` (row-major, lower triangle
/// read), eigenvalues ascending in `evals`, eigenvectors in the columns of
/// `evecs` (row-major).
void self_adjoint_eigen3f(const float cov[9], float evals[3], float evecs[9]);

// ---------------------------------------------------------------------------
// Stage 2: voxel map + association (form/mapping/map.{hpp,tpp},
//          form/optimization/matcher.hpp)
// ---------------------------------------------------------------------------

struct VoxelKey {
  int32_t x, y, z;
  bool operator==(const VoxelKey &o) const { return x == o.x && y == o.y && z == o.z; }
};
struct VoxelKeyHash {
  size_t operator()(const VoxelKey &v) const {
    // map.hpp:37-42 (value not observable; kept for flavour)
    return (size_t)((uint32_t)v.x * 73856093u ^ (uint32_t)v.y * 19349669u ^
                    (uint32_t)v.z * 83492791u);
  }
};

/// VoxelMap::computeCoords, map.tpp:34-38: floor(p / voxel_width) per axis.
VoxelKey compute_coords(double x, double y, double z, double voxel_width);

/// The 27 neighbour shifts in the reference's order, map.tpp:54-68.
extern const int kVoxelShifts[27][3];

/// A world-frame map point with its stable id (rule R4: (scan, k)).
template <typename Feat> struct MapPoint {
  Feat world;  // transformed into the world frame (features.hpp:63-65,137-140)
  uint64_t scan;
  uint32_t k;  // intra-scan insertion index
};

template <typename Feat> struct MatchResult {
  bool found = false;
  uint64_t scan = 0;  // id of the matched map point
  uint32_t k = 0;
  double dist_sqrd = std::numeric_limits<double>::max();
  Feat point_local{}; // matched map point moved back to its own scan frame
};

/// Per-type keypoint store + world voxel map (KeypointMap / VoxelMap).
template <typename Feat> class KeypointMap {
public:
  double min_dist_map = 0.1; // KeypointMapParams, map.hpp:97-100

  /// KeypointMap::get, map.tpp:98-110
  std::vector<Feat> &get(uint64_t scan) { return scans_[scan]; }
  const std::map<uint64_t, std::vector<Feat>> &scans() const { return scans_; }
  /// KeypointMap::remove, map.tpp:112-126
  void remove(uint64_t scan) { scans_.erase(scan); }

  /// KeypointMap::to_voxel_map, map.tpp:128-146 with rule R4 enumeration.
  void to_voxel_map(const std::map<uint64_t, Pose3> &poses, double voxel_width);

  /// VoxelMap::find_closest, map.tpp:70-91 with rule R5 tie-break.
  MatchResult<Feat> find_closest(const Feat &query_world) const;

  /// KeypointMap::insert_matches, map.tpp:148-165.
  size_t insert_matches(uint64_t scan, const std::vector<Feat> &queries,
                        const std::vector<MatchResult<Feat>> &matches);

  size_t num_voxels() const { return voxels_.size(); }
  double voxel_width() const { return voxel_width_; }

private:
  std::map<uint64_t, std::vector<Feat>> scans_; // ordered => rule R4
  double voxel_width_ = 0.5;
  std::unordered_map<VoxelKey, std::vector<MapPoint<Feat>>, VoxelKeyHash> voxels_;
};

/// Transform a keypoint by a pose (features.hpp:63-65 / :137-140).
PointFeat transform(const PointFeat &p, const Pose3 &T);
PlanarFeat transform(const PlanarFeat &p, const Pose3 &T);

// ---------------------------------------------------------------------------
// Stage 3: residuals, Jacobians, Hessian blocks (form/feature/factor.{hpp,cpp},
//          form/optimization/gtsam.hpp:59-140)
// ---------------------------------------------------------------------------

/// PlanePoint, factor.hpp:43-85: SoA of xyz triples.
struct PlanePoint {
  std::vector<double> p_i, n_i, p_j;
  void push_back(const PlanarFeat &pi, const PlanarFeat &pj) {
    p_i.insert(p_i.end(), {pi.x, pi.y, pi.z});
    n_i.insert(n_i.end(), {pi.nx, pi.ny, pi.nz});
    p_j.insert(p_j.end(), {pj.x, pj.y, pj.z});
  }
  void clear() { p_i.clear(); n_i.clear(); p_j.clear(); }
  size_t num_constraints() const { return p_i.size() / 3; }
  size_t num_residuals() const { return p_i.size() / 3; }
  /// PlanePoint::evaluateError, factor.cpp:30-80. r has n entries; H1/H2 are
  /// n x 6 row-major when non-null.
  void evaluate(const Pose3 &Ti, const Pose3 &Tj, double *r, double *H1, double *H2) const;
};

/// PointPoint, factor.hpp:91-130.
struct PointPoint {
  std::vector<double> p_i, p_j;
  void push_back(const PointFeat &pi, const PointFeat &pj) {
    p_i.insert(p_i.end(), {pi.x, pi.y, pi.z});
    p_j.insert(p_j.end(), {pj.x, pj.y, pj.z});
  }
  void clear() { p_i.clear(); p_j.clear(); }
  size_t num_constraints() const { return p_i.size() / 3; }
  size_t num_residuals() const { return p_i.size(); }
  /// PointPoint::evaluateError, factor.cpp:82-128. r has 3m entries; H1/H2 are
  /// 3m x 6 row-major when non-null.
  void evaluate(const Pose3 &Ti, const Pose3 &Tj, double *r, double *H1, double *H2) const;
};

struct PairConstraints {
  PlanePoint plane;
  PointPoint point;
  bool empty() const { return plane.num_constraints() == 0 && point.num_constraints() == 0; }
};

/// FeatureFactor::evaluateError (factor.cpp:141-186) + DenseFactor::linearize
/// (gtsam.hpp:67-86) + FastIsotropic whitening (gtsam.hpp:121-139): writes the
/// upper triangle of the 13x13 augmented information matrix
/// [[A^T A, A^T b],[b^T A, b^T b]], A = [J_i J_j]/sigma, b = -r/sigma,
/// row-major packed (91 doubles).
void linearize_pair(const PairConstraints &c, const Pose3 &Ti, const Pose3 &Tj, double sigma,
                    double out91[91]);

/// NoiseModelFactor::error: 0.5 * sum (r/sigma)^2.
double error_pair(const PairConstraints &c, const Pose3 &Ti, const Pose3 &Tj, double sigma);

} // namespace form_oracle
