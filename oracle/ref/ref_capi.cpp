// TEST INFRASTRUCTURE.  C entry points over FORM's own hot-path translation units, compiled
// UNMODIFIED from /root/reference (form/feature/extraction.{hpp,tpp}, features.hpp,
// form/utils.hpp, form/mapping/map.{hpp,tpp}, form/optimization/matcher.hpp,
// form/feature/factor.{hpp,cpp}, form/optimization/gtsam.hpp) against the API stand-ins in
// oracle/shim/ (Eigen, GTSAM, oneTBB and tsl::robin_map are not in this image).  The result, oracle/_ref/libformref.so,
// is used by tests/test_reference_pins.py to pin the oracle restatement (and through it
// the CUDA path) to the reference's real control flow.  It is never linked or loaded by
// the product.  The structs exchanged are the C-ABI PODs of include/formgpu.h, which are
// byte-identical to FORM's PointXYZf / PointFeat / PlanarFeat.
#include "form/feature/extraction.hpp"
#include "form/mapping/map.hpp"
#include "form/optimization/matcher.hpp"

#include "formgpu.h"

#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <tuple>

static_assert(sizeof(form::PointXYZf) == sizeof(formgpu_point4f), "PointXYZf layout");
static_assert(sizeof(form::PointFeat) == sizeof(formgpu_point_feat), "PointFeat layout");
static_assert(sizeof(form::PlanarFeat) == sizeof(formgpu_planar_feat), "PlanarFeat layout");

namespace {

form::FeatureExtractor::Params extractor_params(const formgpu_params &p) {
  form::FeatureExtractor::Params e;
  e.neighbor_points = (size_t)p.neighbor_points;
  e.num_sectors = (size_t)p.num_sectors;
  e.planar_threshold = p.planar_threshold;
  e.planar_feats_per_sector = (size_t)p.planar_feats_per_sector;
  e.point_feats_per_sector = (size_t)p.point_feats_per_sector;
  e.radius = p.radius;
  e.min_points = (size_t)p.min_points;
  e.min_norm_squared = p.min_norm_squared;
  e.max_norm_squared = p.max_norm_squared;
  e.num_columns = p.num_columns;
  e.num_rows = p.num_rows;
  return e;
}

gtsam::Pose3 to_pose(const formgpu_pose &p) { return gtsam::Pose3(p.R, p.t); }

using Constraints = tsl::robin_map<size_t, std::tuple<form::PlanePoint::Ptr, form::PointPoint::Ptr>>;

// The slice of Estimator state that stages 2a-2c touch (form/form.hpp:59-70).
struct World {
  formgpu_params P;
  form::KeypointMap<form::PlanarFeat> planar_map;
  form::KeypointMap<form::PointFeat> point_map;
  form::Matcher<form::PlanarFeat> planar_matcher;
  form::Matcher<form::PointFeat> point_matcher;
  Constraints constraints; // scan_constraints of the current scan: map scan -> (planar, point)
  explicit World(const formgpu_params &p)
      : P(p), planar_map(form::KeypointMapParams{p.min_dist_map}),
        point_map(form::KeypointMapParams{p.min_dist_map}),
        planar_matcher(form::MatcherParams{p.max_dist_matching, 1e-4, 30}),
        point_matcher(form::MatcherParams{p.max_dist_matching, 1e-4, 30}) {}
};

template <typename Feat, typename Out>
void copy_matches(const tbb::concurrent_vector<form::Match<Feat>> &m, Out *query, Out *point, double *dist,
                  uint8_t *found) {
  for (size_t i = 0; i < m.size(); ++i) {
    std::memcpy(&query[i], &m[i].query, sizeof(Feat));
    std::memcpy(&point[i], &m[i].point, sizeof(Feat));
    dist[i] = m[i].dist_sqrd;
    found[i] = m[i].found() ? 1 : 0;
  }
}

} // namespace

extern "C" {

/// FeatureExtractor::extract on a host scan.  Returns 0, 2 when the reference throws on a
/// wrong scan size (extraction.tpp:141-145), 3 when an output buffer is too small.
int formref_extract(const formgpu_params *p, const formgpu_point4f *scan, size_t n, uint64_t scan_idx,
                    formgpu_planar_feat *planar, size_t planar_cap, size_t *n_planar,
                    formgpu_point_feat *point, size_t point_cap, size_t *n_point) {
  form::FeatureExtractor ex(extractor_params(*p), 1);
  std::vector<form::PointXYZf> pts;
  pts.reserve(n);
  for (size_t i = 0; i < n; ++i) { // bindings.cpp:150-156: x, y, z, padding 0
    pts.emplace_back(scan[i].x, scan[i].y, scan[i].z);
    pts.back()._ = scan[i].w;
  }
  try {
    auto [pl, pt] = ex.extract(pts, (size_t)scan_idx);
    *n_planar = pl.size();
    *n_point = pt.size();
    if (pl.size() > planar_cap || pt.size() > point_cap) return 3;
    if (!pl.empty()) std::memcpy(planar, pl.data(), pl.size() * sizeof(form::PlanarFeat));
    if (!pt.empty()) std::memcpy(point, pt.data(), pt.size() * sizeof(form::PointFeat));
  } catch (const std::runtime_error &) {
    return 2;
  }
  return 0;
}

void *formref_world_create(const formgpu_params *p) { return new World(*p); }
void formref_world_destroy(void *w) { delete static_cast<World *>(w); }

/// VoxelMap::computeCoords through a one-point map (map.tpp:34-52): the voxel of `xyz`.
void formref_compute_coords(double x, double y, double z, double voxel_width, int32_t out[3]) {
  form::VoxelMap<form::PointFeat> map(voxel_width);
  map.push_back(form::PointFeat(x, y, z, 0));
  const auto &key = map.begin()->first;
  out[0] = key(0);
  out[1] = key(1);
  out[2] = key(2);
}

/// Matcher::match<0> + match<1> of the given current-scan keypoints against the stored
/// keypoints at `poses` (form/form.cpp:61-79: to_voxel_map with max_dist_matching, then
/// the two matches).  Outputs are sized by the keypoint counts.
int formref_world_associate(void *wv, const formgpu_scan_pose *poses, size_t n_poses, uint64_t cur_scan,
                            const formgpu_planar_feat *planar, size_t n_planar,
                            const formgpu_point_feat *point, size_t n_point,
                            formgpu_planar_feat *pl_query, formgpu_planar_feat *pl_point, double *pl_dist,
                            uint8_t *pl_found, formgpu_point_feat *pt_query, formgpu_point_feat *pt_point,
                            double *pt_dist, uint8_t *pt_found) {
  World &w = *static_cast<World *>(wv);
  gtsam::Values values;
  std::map<size_t, gtsam::Pose3> pose_of;
  for (size_t i = 0; i < n_poses; ++i) {
    values.insert(X(poses[i].scan), to_pose(poses[i].pose));
    pose_of[(size_t)poses[i].scan] = to_pose(poses[i].pose);
    if (poses[i].scan != cur_scan && w.constraints.find((size_t)poses[i].scan) == w.constraints.end())
      w.constraints.insert(std::make_pair(
          (size_t)poses[i].scan,
          std::make_tuple(std::make_shared<form::PlanePoint>(), std::make_shared<form::PointPoint>())));
  }
  std::vector<form::PlanarFeat> kp_planar(n_planar, form::PlanarFeat(0, 0, 0, 0, 0, 0, 0));
  std::vector<form::PointFeat> kp_point(n_point, form::PointFeat(0, 0, 0, 0));
  if (n_planar) std::memcpy(kp_planar.data(), planar, n_planar * sizeof(form::PlanarFeat));
  if (n_point) std::memcpy(kp_point.data(), point, n_point * sizeof(form::PointFeat));
  const auto world_planar = w.planar_map.to_voxel_map(values, w.P.max_dist_matching);
  const auto world_point = w.point_map.to_voxel_map(values, w.P.max_dist_matching);
  const std::function<gtsam::Pose3(size_t)> estimates = [&](size_t s) { return pose_of.at(s); };
  w.planar_matcher.match<0>(world_planar, kp_planar, estimates, w.constraints);
  w.point_matcher.match<1>(world_point, kp_point, estimates, w.constraints);
  if (w.planar_matcher.matches.size() != n_planar && n_planar) return 1;
  if (w.point_matcher.matches.size() != n_point && n_point) return 1;
  if (n_planar) copy_matches(w.planar_matcher.matches, pl_query, pl_point, pl_dist, pl_found);
  if (n_point) copy_matches(w.point_matcher.matches, pt_query, pt_point, pt_dist, pt_found);
  return 0;
}

/// Correspondences the last association appended for map scan `scan` (matcher.hpp:103-111).
void formref_world_constraint_counts(void *wv, uint64_t scan, size_t *n_planar, size_t *n_point) {
  World &w = *static_cast<World *>(wv);
  *n_planar = *n_point = 0;
  auto it = w.constraints.find((size_t)scan);
  if (it == w.constraints.end()) return;
  *n_planar = std::get<0>(it.value())->num_constraints();
  *n_point = std::get<1>(it.value())->num_constraints();
}

/// DenseFactor::linearize of FeatureFactor(X(scan_i), X(cur_scan)) over the correspondences
/// the last association appended for map scan `scan_i` (form/optimization/gtsam.hpp:67-86,
/// form/feature/factor.cpp:131-186, constraints.cpp:259-265): the packed upper triangle of
/// the 13x13 augmented information matrix.  Returns 1 when the pair has no correspondences.
int formref_world_linearize(void *wv, uint64_t scan_i, uint64_t cur_scan, const formgpu_pose *Ti,
                            const formgpu_pose *Tj, double sigma, double *out91) {
  World &w = *static_cast<World *>(wv);
  auto it = w.constraints.find((size_t)scan_i);
  if (it == w.constraints.end()) return 1;
  const auto &c = it.value();
  if (std::get<0>(c)->num_constraints() + std::get<1>(c)->num_constraints() == 0) return 1;
  form::FeatureFactor factor(X(scan_i), X(cur_scan), c, sigma);
  gtsam::Values values;
  values.insert(X(scan_i), to_pose(*Ti));
  values.insert(X(cur_scan), to_pose(*Tj));
  const auto gf = factor.linearize(values);
  const auto *hf = dynamic_cast<const gtsam::HessianFactor *>(gf.get());
  if (!hf || hf->info.rows() != 13) return 2;
  size_t e = 0;
  for (int r = 0; r < 13; ++r)
    for (int col = r; col < 13; ++col) out91[e++] = hf->info(r, col);
  return 0;
}

/// KeypointMap::insert_matches for both types (form/form.cpp:99-101).
void formref_world_commit(void *wv) {
  World &w = *static_cast<World *>(wv);
  w.planar_map.insert_matches(w.planar_matcher.matches);
  w.point_map.insert_matches(w.point_matcher.matches);
}

/// KeypointMap::remove (form/form.cpp:111) and the erase of the scan's constraints.
void formref_world_remove(void *wv, uint64_t scan) {
  World &w = *static_cast<World *>(wv);
  w.planar_map.remove((size_t)scan);
  w.point_map.remove((size_t)scan);
  w.constraints.erase((size_t)scan);
}

/// Stored (scan-local) keypoints of a scan; out may be NULL to query the count.
size_t formref_world_keypoints(void *wv, int type, uint64_t scan, void *out, size_t cap) {
  World &w = *static_cast<World *>(wv);
  if (type == 0) {
    const auto &v = w.planar_map.get((size_t)scan);
    if (out && v.size() <= cap && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(form::PlanarFeat));
    return v.size();
  }
  const auto &v = w.point_map.get((size_t)scan);
  if (out && v.size() <= cap && !v.empty()) std::memcpy(out, v.data(), v.size() * sizeof(form::PointFeat));
  return v.size();
}

} // extern "C"

namespace {
void copy_rows(const Eigen::MatrixXd &H, double *out) { // rows x 6, row-major out
  for (size_t r = 0; r < H.rows(); ++r)
    for (size_t c = 0; c < 6; ++c) out[6 * r + c] = H(r, c);
}
} // namespace

extern "C" {

/// PlanePoint::evaluateError (form/feature/factor.cpp:30-80) on n correspondences given as
/// n x 3 row-major arrays: residual[n], H1 / H2 [n][6].
void formref_plane_point(const double *p_i, const double *n_i, const double *p_j, size_t n,
                         const formgpu_pose *Ti, const formgpu_pose *Tj, double *residual, double *H1,
                         double *H2) {
  form::PlanePoint pp;
  pp.p_i.assign(p_i, p_i + 3 * n);
  pp.n_i.assign(n_i, n_i + 3 * n);
  pp.p_j.assign(p_j, p_j + 3 * n);
  Eigen::MatrixXd A, B;
  const gtsam::Vector r = pp.evaluateError(to_pose(*Ti), to_pose(*Tj), &A, &B);
  for (size_t k = 0; k < n; ++k) residual[k] = r(k);
  copy_rows(A, H1);
  copy_rows(B, H2);
}

/// PointPoint::evaluateError (factor.cpp:82-128): residual[3m], H1 / H2 [3m][6].
void formref_point_point(const double *p_i, const double *p_j, size_t m, const formgpu_pose *Ti,
                         const formgpu_pose *Tj, double *residual, double *H1, double *H2) {
  form::PointPoint pp;
  pp.p_i.assign(p_i, p_i + 3 * m);
  pp.p_j.assign(p_j, p_j + 3 * m);
  Eigen::MatrixXd A, B;
  const gtsam::Vector r = pp.evaluateError(to_pose(*Ti), to_pose(*Tj), &A, &B);
  for (size_t k = 0; k < 3 * m; ++k) residual[k] = r(k);
  copy_rows(A, H1);
  copy_rows(B, H2);
}

/// FeatureFactor + DenseFactor::linearize (factor.cpp:131-186, gtsam.hpp:67-86) on raw
/// correspondences: the packed upper triangle of the 13x13 augmented information matrix.
int formref_linearize_raw(const double *p_i, const double *n_i, const double *p_j, size_t n,
                          const double *q_i, const double *q_j, size_t m, const formgpu_pose *Ti,
                          const formgpu_pose *Tj, double sigma, double *out91) {
  auto planar = std::make_shared<form::PlanePoint>();
  planar->p_i.assign(p_i, p_i + 3 * n);
  planar->n_i.assign(n_i, n_i + 3 * n);
  planar->p_j.assign(p_j, p_j + 3 * n);
  auto point = std::make_shared<form::PointPoint>();
  point->p_i.assign(q_i, q_i + 3 * m);
  point->p_j.assign(q_j, q_j + 3 * m);
  form::FeatureFactor factor(X(0), X(1), std::make_tuple(planar, point), sigma);
  gtsam::Values values;
  values.insert(X(0), to_pose(*Ti));
  values.insert(X(1), to_pose(*Tj));
  const auto gf = factor.linearize(values);
  const auto *hf = dynamic_cast<const gtsam::HessianFactor *>(gf.get());
  if (!hf || hf->info.rows() != 13) return 2;
  size_t e = 0;
  for (int r = 0; r < 13; ++r)
    for (int col = r; col < 13; ++col) out91[e++] = hf->info(r, col);
  return 0;
}

} // extern "C"

// ---- FORM's own KeyScanner (form/mapping/keyscanner.{hpp,cpp}, no external dependency):
// pins form_b200/host/form/keyscanner.hpp in tests/test_reference_pins.py ----
#include "form/mapping/keyscanner.hpp"

extern "C" {

typedef size_t (*ref_connections_fn)(uint64_t scan, void *user);

void *ref_keyscanner_create(int64_t max_num_keyscans, int64_t max_steps_unused_keyscan,
                            size_t max_num_recent_scans, double keyscan_match_ratio) {
  form::KeyScanner::Params p;
  p.max_num_keyscans = max_num_keyscans;
  p.max_steps_unused_keyscan = max_steps_unused_keyscan;
  p.max_num_recent_scans = max_num_recent_scans;
  p.keyscan_match_ratio = keyscan_match_ratio;
  return new form::KeyScanner(p);
}
void ref_keyscanner_destroy(void *h) { delete static_cast<form::KeyScanner *>(h); }

/// KeyScanner::step (keyscanner.cpp:29-91); returns the number of scans to marginalise
size_t ref_keyscanner_step(void *h, uint64_t idx, size_t size, ref_connections_fn fn, void *user, uint64_t *marg,
                           size_t cap) {
  const std::vector<form::ScanIndex> out =
      static_cast<form::KeyScanner *>(h)->step(idx, size, [=](form::ScanIndex s) { return fn(s, user); });
  for (size_t k = 0; k < out.size() && k < cap; ++k) marg[k] = out[k];
  return out.size();
}
size_t ref_keyscanner_size(void *h) { return static_cast<form::KeyScanner *>(h)->size(); }

} // extern "C"
