// Flat C API over form::Estimator and the trace replayer, generated twice from
// this one header: libformhost.so instantiates it with the CUDA hot path
// (prefix formhost_), the test-only oracle library with the CPU oracle (prefix
// oracle_).  Used by the Python tests, bench.py and the evalio-style pipeline.
#pragma once

#include "form/form.hpp"
#include "form/trace.hpp"
#include "formgpu.h"

#include <chrono>
#include <cstring>
#include <memory>

extern "C" {
/// Parameters of Estimator::Params that are not part of formgpu_params
/// (python/bindings.cpp:76-87 of FORM).
typedef struct formhost_est_params {
  formgpu_params hot;
  double new_pose_threshold;       /* 1e-4 */
  double keyscan_match_ratio;      /* 0.1 */
  int32_t max_num_rematches;       /* 30 */
  int32_t disable_smoothing;       /* 0 */
  int32_t max_num_keyscans;        /* 50 */
  int32_t max_num_recent_scans;    /* 10 */
  int32_t max_steps_unused_keyscan;/* 10 */
  int32_t num_threads;             /* 0 */
  int32_t device;                  /* 0 */
  int32_t record_trace;            /* 0 */
  int32_t gtsam_lm_schedule;       /* 0: fused trial linearisation; 1: GTSAM's linearise + error calls */
  int32_t reserved;
} formhost_est_params;
}

namespace form {
namespace capi {

inline Estimator::Params to_estimator_params(const formhost_est_params &p) {
  Estimator::Params e;
  const formgpu_params &h = p.hot;
  e.extraction.neighbor_points = (size_t)h.neighbor_points;
  e.extraction.num_sectors = (size_t)h.num_sectors;
  e.extraction.planar_threshold = h.planar_threshold;
  e.extraction.planar_feats_per_sector = (size_t)h.planar_feats_per_sector;
  e.extraction.point_feats_per_sector = (size_t)h.point_feats_per_sector;
  e.extraction.radius = h.radius;
  e.extraction.min_points = (size_t)h.min_points;
  e.extraction.min_norm_squared = h.min_norm_squared;
  e.extraction.max_norm_squared = h.max_norm_squared;
  e.extraction.num_columns = h.num_columns;
  e.extraction.num_rows = h.num_rows;
  e.matcher.max_dist_matching = h.max_dist_matching;
  e.matcher.new_pose_threshold = p.new_pose_threshold;
  e.matcher.max_num_rematches = (size_t)p.max_num_rematches;
  e.constraints.disable_smoothing = p.disable_smoothing != 0;
  e.constraints.planar_constraint_sigma = h.sigma;
  e.constraints.fused_trial_linearization = p.gtsam_lm_schedule == 0;
  e.scans.max_num_keyscans = p.max_num_keyscans;
  e.scans.max_num_recent_scans = (size_t)p.max_num_recent_scans;
  e.scans.max_steps_unused_keyscan = p.max_steps_unused_keyscan;
  e.scans.keyscan_match_ratio = p.keyscan_match_ratio;
  e.map.min_dist_map = h.min_dist_map;
  e.num_threads = (size_t)p.num_threads;
  e.device = p.device;
  return e;
}

struct EstimatorHandle {
  std::shared_ptr<HotPath> backend;
  std::shared_ptr<HotPath> recorder;
  Trace trace;
  std::unique_ptr<Estimator> est;
  std::string error;
};

struct ReplayHandle {
  const Trace *trace = nullptr;
  std::shared_ptr<HotPath> backend;
  ReplayStats stats;
  size_t points_per_scan = 0;
  std::vector<PlanarFeat> planar;
  std::vector<PointFeat> point;
};

template <typename Factory>
EstimatorHandle *est_create(const formhost_est_params *p, Factory make_backend, std::string &err) {
  auto h = std::make_unique<EstimatorHandle>();
  try {
    const Estimator::Params ep = to_estimator_params(*p);
    h->backend = make_backend(Estimator::hotpath_params(ep), *p);
    std::shared_ptr<HotPath> use = h->backend;
    if (p->record_trace) {
      h->recorder = std::make_shared<RecordingHotPath>(*h->backend, h->trace);
      use = h->recorder;
    }
    h->est = std::make_unique<Estimator>(ep, use);
  } catch (const std::exception &e) {
    err = e.what();
    return nullptr;
  }
  return h.release();
}

inline int est_register_scan(EstimatorHandle *h, const formgpu_point4f *scan, size_t n,
                             formgpu_planar_feat *planar, size_t planar_cap, size_t *n_planar,
                             formgpu_point_feat *point, size_t point_cap, size_t *n_point) {
  const size_t before = h->est->last_error().size();
  (void)before;
  auto kp = h->est->register_scan(reinterpret_cast<const PointXYZf *>(scan), n);
  if (!h->est->last_error().empty() && std::get<0>(kp).empty() && std::get<1>(kp).empty() &&
      h->est->last_error() != h->error) {
    h->error = h->est->last_error();
    return FORMGPU_ERR_STATE;
  }
  const auto &pl = std::get<0>(kp);
  const auto &pt = std::get<1>(kp);
  if (n_planar) *n_planar = pl.size();
  if (n_point) *n_point = pt.size();
  if (pl.size() > planar_cap || pt.size() > point_cap) return FORMGPU_ERR_CAPACITY;
  if (planar && !pl.empty()) std::memcpy(planar, pl.data(), pl.size() * sizeof(PlanarFeat));
  if (point && !pt.empty()) std::memcpy(point, pt.data(), pt.size() * sizeof(PointFeat));
  return FORMGPU_OK;
}

inline void est_pose(EstimatorHandle *h, formgpu_pose *out) {
  const Pose3 T = h->est->current_lidar_estimate();
  std::memcpy(out, &T, sizeof(Pose3));
}

inline int est_window(EstimatorHandle *h, formgpu_scan_pose *out, size_t cap, size_t *n) {
  const Values &v = h->est->m_constraints.get_values();
  *n = v.size();
  if (v.size() > cap) return FORMGPU_ERR_CAPACITY;
  size_t k = 0;
  for (const auto &kv : v) {
    out[k].scan = kv.first;
    std::memcpy(&out[k].pose, &kv.second, sizeof(Pose3));
    ++k;
  }
  return FORMGPU_OK;
}

inline void est_stats(EstimatorHandle *h, uint64_t out[8]) {
  const SmootherStats &s = h->est->m_constraints.stats();
  out[0] = s.optimize_calls;
  out[1] = s.lm_iterations;
  out[2] = s.linearize_calls;
  out[3] = s.error_calls;
  out[4] = s.linearized_pairs;
  out[5] = s.error_pairs;
  out[6] = h->est->icp_iterations();
  out[7] = h->est->m_constraints.get_values().size();
}

inline int est_map(EstimatorHandle *h, formgpu_planar_feat *planar, size_t planar_cap,
                   size_t *n_planar, formgpu_point_feat *point, size_t point_cap, size_t *n_point) {
  std::vector<PlanarFeat> pl;
  std::vector<PointFeat> pt;
  try {
    h->est->m_keypoint_map.world_keypoints(h->est->m_constraints.get_values(), pl, pt);
  } catch (const std::exception &e) {
    h->error = e.what();
    return FORMGPU_ERR_STATE;
  }
  *n_planar = pl.size();
  *n_point = pt.size();
  if (pl.size() > planar_cap || pt.size() > point_cap) return FORMGPU_ERR_CAPACITY;
  if (!pl.empty()) std::memcpy(planar, pl.data(), pl.size() * sizeof(PlanarFeat));
  if (!pt.empty()) std::memcpy(point, pt.data(), pt.size() * sizeof(PointFeat));
  return FORMGPU_OK;
}

/// Replay scans [first, last); scans[s] points at scan s (host memory).  Returns
/// the wall time in seconds (steady_clock around the calls, as form::Timer does).
inline double replay_run_host(ReplayHandle *r, size_t first, size_t last,
                              const formgpu_point4f *const *scans) {
  const auto t0 = std::chrono::steady_clock::now();
  replay(*r->trace, *r->backend, first, last,
         [&](uint64_t scan_idx, size_t &np, size_t &nq) {
           // position of this scan in the trace == its index in `scans`
           r->backend->extract(reinterpret_cast<const PointXYZf *>(scans[scan_idx]),
                               r->points_per_scan, scan_idx, r->planar, r->point);
           np = r->planar.size();
           nq = r->point.size();
         },
         r->points_per_scan, r->stats);
  return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

inline void replay_stats(const ReplayHandle *r, uint64_t out[20], double *checksum) {
  const ReplayStats &s = r->stats;
  const uint64_t v[20] = {s.scans,       s.points,      s.planar_kp,   s.point_kp,  s.assoc_calls,
                          s.assoc_queries, s.map_rebuilds, s.map_points, s.lin_calls, s.lin_pairs,
                          s.lin_planar,  s.lin_point,   s.err_calls,   s.err_pairs, s.err_planar,
                          s.err_point,   s.novel_planar, s.novel_point, s.assoc_planar, s.assoc_point};
  std::memcpy(out, v, sizeof(v));
  if (checksum) *checksum = s.checksum;
}

} // namespace capi
} // namespace form
