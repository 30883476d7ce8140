// TEST INFRASTRUCTURE.  C entry points over FORM's own Estimator (see the section comment below);
// a translation unit of its own because form/mapping/keyscanner.hpp and
// form/optimization/matcher.hpp carry no include guard and ref_capi.cpp already includes them.
#include "form/form.hpp"

#include "formgpu.h"

#include <cstring>
#include <vector>

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
} // namespace

// ---- FORM's own Estimator: form/form.{hpp,cpp} + form/optimization/constraints.{hpp,cpp}, compiled
// unmodified over the smoother stand-ins of oracle/shim/gtsam/shim_smoother.h.  The control flow of
// register_scan (form.cpp:40-114), of the pair policy (constraints.cpp:252-308) and of
// marginalize (constraints.cpp:120-195) is the reference's own; GTSAM's optimiser is a stand-in
// with its published behaviour.  Pins form_b200/host/form/{form,constraints}.hpp end to end
// (tests/test_reference_pipeline.py). ----

extern "C" {

void *formref_est_create(const formgpu_params *p, double new_pose_threshold, double keyscan_match_ratio,
                         int max_num_rematches, int disable_smoothing, int64_t max_num_keyscans,
                         size_t max_num_recent_scans, int64_t max_steps_unused_keyscan) {
  form::Estimator::Params e;
  e.extraction = extractor_params(*p);
  e.matcher.max_dist_matching = p->max_dist_matching;
  e.matcher.new_pose_threshold = new_pose_threshold;
  e.matcher.max_num_rematches = (size_t)max_num_rematches;
  e.constraints.disable_smoothing = disable_smoothing != 0;
  e.constraints.planar_constraint_sigma = p->sigma;
  e.scans.max_num_keyscans = max_num_keyscans;
  e.scans.max_num_recent_scans = max_num_recent_scans;
  e.scans.max_steps_unused_keyscan = max_steps_unused_keyscan;
  e.scans.keyscan_match_ratio = keyscan_match_ratio;
  e.map.min_dist_map = p->min_dist_map;
  e.num_threads = 1;
  return new form::Estimator(e);
}
void formref_est_destroy(void *h) { delete static_cast<form::Estimator *>(h); }

/// Estimator::register_scan.  Returns 0, or 3 when an output buffer is too small.
int formref_est_register_scan(void *h, const formgpu_point4f *scan, size_t n, formgpu_planar_feat *planar,
                              size_t planar_cap, size_t *n_planar, formgpu_point_feat *point, size_t point_cap,
                              size_t *n_point) {
  std::vector<form::PointXYZf> s(n, form::PointXYZf(0.f, 0.f, 0.f));
  std::memcpy(s.data(), scan, n * sizeof(form::PointXYZf));
  const auto kp = static_cast<form::Estimator *>(h)->register_scan(s);
  const auto &pl = std::get<0>(kp);
  const auto &pt = std::get<1>(kp);
  *n_planar = pl.size();
  *n_point = pt.size();
  if (pl.size() > planar_cap || pt.size() > point_cap) return 3;
  if (!pl.empty()) std::memcpy(planar, pl.data(), pl.size() * sizeof(form::PlanarFeat));
  if (!pt.empty()) std::memcpy(point, pt.data(), pt.size() * sizeof(form::PointFeat)); // data() may be null when empty
  return 0;
}

static void copy_pose(const gtsam::Pose3 &T, formgpu_pose *out) {
  for (int i = 0; i < 9; ++i) out->R[i] = T.rotation().data()[i];
  const auto t = T.translation();
  for (int i = 0; i < 3; ++i) out->t[i] = t(i);
}
void formref_est_pose(void *h, formgpu_pose *out) {
  copy_pose(static_cast<form::Estimator *>(h)->current_lidar_estimate(), out);
}
/// the poses of the fixed-lag window (ConstraintManager::get_values), ascending scan index
size_t formref_est_window(void *h, formgpu_scan_pose *out, size_t cap) {
  const gtsam::Values &v = static_cast<form::Estimator *>(h)->m_constraints.get_values();
  size_t k = 0;
  for (const auto &kv : v) {
    if (k < cap) {
      out[k].scan = kv.first;
      copy_pose(kv.second, &out[k].pose);
    }
    ++k;
  }
  return k;
}

} // extern "C"
