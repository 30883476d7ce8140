// TEST INFRASTRUCTURE - see oracle.hpp.  form::ConstraintManager (the host-side fixed-lag
// smoother of form_b200/host/form/constraints.hpp, a restatement of
// /root/reference/form/optimization/constraints.cpp:39-336 over GTSAM's published semantics)
// driven over HAND-MADE correspondences, so that tests/test_smoother_independent.py can check
// its Levenberg-Marquardt optimum and its Schur-complement marginals against an independent
// numpy / scipy formulation of the same least-squares problem.  Nothing of the product links this.
#include "form/constraints.hpp"
#include "formgpu.h"
#include "oracle.hpp"

#include <cstring>
#include <map>
#include <memory>

using namespace form;

namespace {

/// A hot path whose pairs carry the correspondences the test put there: stage 3 only
/// (linearize_pair / error_pair of oracle_factor.cpp), every other stage is out of the test.
class RawHotPath : public HotPath {
public:
  explicit RawHotPath(double sigma) : m_sigma(sigma) {}
  std::map<std::pair<uint64_t, uint64_t>, form_oracle::PairConstraints> pairs; // (i, j), j > i

  void extract(const PointXYZf *, size_t, uint64_t, std::vector<PlanarFeat> &, std::vector<PointFeat> &) override {
    throw HotPathError("RawHotPath: stage 1 is not part of the smoother test");
  }
  void map_rebuild(const ScanPose *, size_t) override {}
  void associate(const Pose3 &, std::vector<PairCount> &) override {
    throw HotPathError("RawHotPath: stage 2 is not part of the smoother test");
  }
  void linearize(const PairKey *keys, size_t n_pairs, const ScanPose *poses, size_t n_poses, double *out91) override {
    for (size_t p = 0; p < n_pairs; ++p)
      form_oracle::linearize_pair(pairs.at({keys[p].i, keys[p].j}), find(poses, n_poses, keys[p].i),
                                  find(poses, n_poses, keys[p].j), m_sigma, out91 + 91 * p);
  }
  void error(const PairKey *keys, size_t n_pairs, const ScanPose *poses, size_t n_poses, double *out) override {
    for (size_t p = 0; p < n_pairs; ++p)
      out[p] = form_oracle::error_pair(pairs.at({keys[p].i, keys[p].j}), find(poses, n_poses, keys[p].i),
                                       find(poses, n_poses, keys[p].j), m_sigma);
  }
  void commit_scan(size_t &a, size_t &b) override { a = b = 0; }
  void remove_scans(const uint64_t *scans, size_t n) override {
    for (auto it = pairs.begin(); it != pairs.end();) {
      bool hit = false;
      for (size_t k = 0; k < n; ++k) hit = hit || it->first.first == scans[k] || it->first.second == scans[k];
      it = hit ? pairs.erase(it) : std::next(it);
    }
  }
  void world_keypoints(const ScanPose *, size_t, std::vector<PlanarFeat> &, std::vector<PointFeat> &) override {}

private:
  static const Pose3 &find(const ScanPose *poses, size_t n, uint64_t scan) {
    for (size_t p = 0; p < n; ++p)
      if (poses[p].scan == scan) return poses[p].pose;
    throw HotPathError("RawHotPath: no pose for a scan of the pair");
  }
  double m_sigma;
};

struct SmootherHandle {
  RawHotPath hp;
  ConstraintManager cm;
  std::string error;
  SmootherHandle(double sigma, const ConstraintManager::Params &p) : hp(sigma), cm(p) { cm.set_hotpath(&hp); }
};

} // namespace

extern "C" {

/// tolerances <= 0 keep GTSAM's defaults (1e-5 / 1e-5)
void *oracle_smoother_create(double sigma, double pose_sigma, int fused, int disable_smoothing,
                             double rel_tol, double abs_tol) {
  ConstraintManager::Params p;
  p.planar_constraint_sigma = sigma;
  p.pose_sigma = pose_sigma;
  p.fused_trial_linearization = fused != 0;
  p.disable_smoothing = disable_smoothing != 0;
  if (rel_tol > 0.0) p.opt_params.relativeErrorTol = rel_tol;
  if (abs_tol > 0.0) p.opt_params.absoluteErrorTol = abs_tol;
  return new SmootherHandle(sigma, p);
}
void oracle_smoother_destroy(void *h) { delete static_cast<SmootherHandle *>(h); }
const char *oracle_smoother_error(void *h) { return static_cast<SmootherHandle *>(h)->error.c_str(); }

/// correspondences of pair (i, j), j > i: n plane rows (p_i, n_i, p_j) and m point rows (p_i, p_j),
/// xyz triples in the frames of scan i / scan j (factor.hpp:43-130)
void oracle_smoother_set_pair(void *h, uint64_t i, uint64_t j, const double *pl_pi, const double *pl_ni,
                              const double *pl_pj, size_t n, const double *pt_pi, const double *pt_pj, size_t m) {
  form_oracle::PairConstraints c;
  c.plane.p_i.assign(pl_pi, pl_pi + 3 * n);
  c.plane.n_i.assign(pl_ni, pl_ni + 3 * n);
  c.plane.p_j.assign(pl_pj, pl_pj + 3 * n);
  c.point.p_i.assign(pt_pi, pt_pi + 3 * m);
  c.point.p_j.assign(pt_pj, pt_pj + 3 * m);
  static_cast<SmootherHandle *>(h)->hp.pairs[{i, j}] = std::move(c);
}

/// ConstraintManager::step + the counts Matcher::match would have left for the new scan
uint64_t oracle_smoother_step(void *h, const formgpu_pose *pose) {
  auto *s = static_cast<SmootherHandle *>(h);
  const uint64_t scan = s->cm.step(*reinterpret_cast<const Pose3 *>(pose));
  std::vector<PairCount> counts;
  for (const auto &kv : s->hp.pairs)
    if (kv.first.second == scan)
      counts.push_back({kv.first.first, (uint32_t)kv.second.plane.num_constraints(),
                        (uint32_t)kv.second.point.num_constraints()});
  s->cm.set_current_counts(counts);
  return scan;
}

/// optimize(fast) followed by update_values, as register_scan does (form.cpp:92-93)
int oracle_smoother_optimize(void *h, int fast) {
  auto *s = static_cast<SmootherHandle *>(h);
  try {
    s->cm.update_values(s->cm.optimize(fast != 0));
    return 0;
  } catch (const std::exception &e) {
    s->error = e.what();
    return -1;
  }
}

int oracle_smoother_marginalize(void *h, const uint64_t *scans, size_t n) {
  auto *s = static_cast<SmootherHandle *>(h);
  try {
    s->cm.marginalize(std::vector<ScanIndex>(scans, scans + n));
    return 0;
  } catch (const std::exception &e) {
    s->error = e.what();
    return -1;
  }
}

void oracle_smoother_set_pose(void *h, uint64_t scan, const formgpu_pose *pose) {
  static_cast<SmootherHandle *>(h)->cm.update_pose(scan, *reinterpret_cast<const Pose3 *>(pose));
}

size_t oracle_smoother_values(void *h, formgpu_scan_pose *out, size_t cap) {
  const Values &v = static_cast<SmootherHandle *>(h)->cm.get_values();
  size_t k = 0;
  for (const auto &kv : v) {
    if (k < cap) {
      out[k].scan = kv.first;
      std::memcpy(&out[k].pose, &kv.second, sizeof(formgpu_pose));
    }
    ++k;
  }
  return k;
}

/// optimize_calls, lm_iterations, linearize_calls, error_calls, linearized_pairs, error_pairs
void oracle_smoother_stats(void *h, uint64_t out[6]) {
  const SmootherStats &st = static_cast<SmootherHandle *>(h)->cm.stats();
  out[0] = st.optimize_calls;
  out[1] = st.lm_iterations;
  out[2] = st.linearize_calls;
  out[3] = st.error_calls;
  out[4] = st.linearized_pairs;
  out[5] = st.error_pairs;
}

/// The live marginal factors (LinearContainerFactor semantics): number of them, and for the
/// idx-th its keys, linearisation points and the quadratic 0.5 (f - 2 g.d + d.G.d).
size_t oracle_smoother_num_marginals(void *h) { return static_cast<SmootherHandle *>(h)->cm.marginals().size(); }
size_t oracle_smoother_marginal_keys(void *h, size_t idx, uint64_t *keys, formgpu_pose *lin_points, size_t cap) {
  const LinearContainer &c = *static_cast<SmootherHandle *>(h)->cm.marginals().at(idx);
  for (size_t k = 0; k < c.keys.size() && k < cap; ++k) {
    keys[k] = c.keys[k];
    std::memcpy(&lin_points[k], &c.lin_points[k], sizeof(formgpu_pose));
  }
  return c.keys.size();
}
void oracle_smoother_marginal_quadratic(void *h, size_t idx, double *G, double *g, double *f) {
  const LinearContainer &c = *static_cast<SmootherHandle *>(h)->cm.marginals().at(idx);
  std::copy(c.q.G.begin(), c.q.G.end(), G);
  std::copy(c.q.g.begin(), c.q.g.end(), g);
  *f = c.q.f;
}

} // extern "C"

// ---- form::KeyScanner (host/form/keyscanner.hpp; /root/reference/form/mapping/keyscanner.cpp:29-91)
// behind a C callback for the connection counts, for tests/test_keyscanner_model.py ----
#include "form/keyscanner.hpp"

extern "C" {

typedef size_t (*oracle_connections_fn)(uint64_t scan, void *user);

void *oracle_keyscanner_create(int64_t max_num_keyscans, int64_t max_steps_unused_keyscan,
                               size_t max_num_recent_scans, double keyscan_match_ratio) {
  KeyScanner::Params p;
  p.max_num_keyscans = max_num_keyscans;
  p.max_steps_unused_keyscan = max_steps_unused_keyscan;
  p.max_num_recent_scans = max_num_recent_scans;
  p.keyscan_match_ratio = keyscan_match_ratio;
  return new KeyScanner(p);
}
void oracle_keyscanner_destroy(void *h) { delete static_cast<KeyScanner *>(h); }

/// KeyScanner::step; returns the number of scans to marginalise (written to marg, in order)
size_t oracle_keyscanner_step(void *h, uint64_t idx, size_t size, oracle_connections_fn fn, void *user,
                              uint64_t *marg, size_t cap) {
  const std::vector<ScanIndex> out =
      static_cast<KeyScanner *>(h)->step(idx, size, [&](ScanIndex s) { return fn(s, user); });
  for (size_t k = 0; k < out.size() && k < cap; ++k) marg[k] = out[k];
  return out.size();
}

/// key scans then recent scans, oldest first; returns their numbers through n_key / n_recent
void oracle_keyscanner_state(void *h, uint64_t *key, size_t *n_key, uint64_t *recent, size_t *n_recent, size_t cap) {
  const KeyScanner &ks = *static_cast<KeyScanner *>(h);
  size_t k = 0;
  for (const Scan &s : ks.keyscans())
    if (k < cap) key[k++] = s.idx;
  *n_key = ks.keyscans().size();
  k = 0;
  for (const Scan &s : ks.recent_scans())
    if (k < cap) recent[k++] = s.idx;
  *n_recent = ks.recent_scans().size();
}
size_t oracle_keyscanner_oldest_rf(void *h) { return static_cast<KeyScanner *>(h)->oldest_rf(); }

} // extern "C"

extern "C" size_t oracle_keyscanner_size(void *h) { return static_cast<KeyScanner *>(h)->size(); }
