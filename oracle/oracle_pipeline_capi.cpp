// TEST INFRASTRUCTURE - see oracle.hpp.  The shared host logic (form::Estimator,
// trace replay) instantiated over the CPU oracle: the reference pipeline the
// CUDA pipeline is compared with, and bench.py's CPU baseline.
#include "form/capi_impl.hpp"
#include "oracle_hotpath.hpp"

using namespace form;
using namespace form::capi;

namespace {
std::string g_error;
std::shared_ptr<HotPath> make_oracle(const HotPathParams &hp, const formhost_est_params &p) {
  HotPathParams h = hp;
  h.num_threads = (size_t)p.num_threads;
  return std::make_shared<form_oracle::OracleHotPath>(h);
}
} // namespace

extern "C" {

const char *oracle_est_last_error(void) { return g_error.c_str(); }
void *oracle_est_create(const formhost_est_params *p) { return est_create(p, make_oracle, g_error); }
void oracle_est_destroy(void *h) { delete static_cast<EstimatorHandle *>(h); }
int oracle_est_register_scan(void *h, const formgpu_point4f *scan, size_t n,
                             formgpu_planar_feat *planar, size_t planar_cap, size_t *n_planar,
                             formgpu_point_feat *point, size_t point_cap, size_t *n_point) {
  return est_register_scan(static_cast<EstimatorHandle *>(h), scan, n, planar, planar_cap, n_planar,
                           point, point_cap, n_point);
}
void oracle_est_pose(void *h, formgpu_pose *out) { est_pose(static_cast<EstimatorHandle *>(h), out); }
int oracle_est_window(void *h, formgpu_scan_pose *out, size_t cap, size_t *n) {
  return est_window(static_cast<EstimatorHandle *>(h), out, cap, n);
}
void oracle_est_stats(void *h, uint64_t out[8]) { est_stats(static_cast<EstimatorHandle *>(h), out); }
int oracle_est_map(void *h, formgpu_planar_feat *planar, size_t planar_cap, size_t *n_planar,
                   formgpu_point_feat *point, size_t point_cap, size_t *n_point) {
  return est_map(static_cast<EstimatorHandle *>(h), planar, planar_cap, n_planar, point, point_cap,
                 n_point);
}
const void *oracle_est_trace(void *h) { return &static_cast<EstimatorHandle *>(h)->trace; }
size_t oracle_trace_num_scans(const void *t) { return static_cast<const Trace *>(t)->num_scans(); }

void *oracle_replay_create(const void *trace, const formhost_est_params *p) {
  auto r = std::make_unique<ReplayHandle>();
  r->trace = static_cast<const Trace *>(trace);
  const Estimator::Params ep = to_estimator_params(*p);
  HotPathParams hp = Estimator::hotpath_params(ep);
  hp.num_threads = (size_t)p->num_threads;
  r->backend = std::make_shared<form_oracle::OracleHotPath>(hp);
  r->points_per_scan = (size_t)p->hot.num_rows * p->hot.num_columns;
  return r.release();
}
void oracle_replay_destroy(void *r) { delete static_cast<ReplayHandle *>(r); }
double oracle_replay_run_host(void *r, size_t first, size_t last,
                              const formgpu_point4f *const *scans) {
  try {
    return replay_run_host(static_cast<ReplayHandle *>(r), first, last, scans);
  } catch (const std::exception &e) {
    g_error = e.what();
    return -1.0;
  }
}
void oracle_replay_stats(void *r, uint64_t out[20], double *checksum) {
  replay_stats(static_cast<ReplayHandle *>(r), out, checksum);
}
void oracle_replay_reset_stats(void *r) {
  auto *h = static_cast<ReplayHandle *>(r);
  auto table = std::move(h->stats.table);
  h->stats = ReplayStats();
  h->stats.table = std::move(table);
}

} // extern "C"
