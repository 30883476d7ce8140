// TEST INFRASTRUCTURE - see oracle.hpp.
#include "oracle_hotpath.hpp"

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

namespace form_oracle {

namespace {

// Persistent worker pool standing in for TBB's parallel_for
// (extraction.tpp:103, matcher.hpp:87; GTSAM linearises factors in parallel).
class Pool {
public:
  static Pool &get() {
    // leaked on purpose: the detached workers wait on its condition variable for the
    // whole process lifetime, and destroying a condvar with waiters blocks at exit
    static Pool *p = new Pool();
    return *p;
  }
  void parallel_for(size_t n, int nthreads, const std::function<void(size_t, size_t)> &fn) {
    if (nthreads <= 1 || n < 2) {
      fn(0, n);
      return;
    }
    std::unique_lock<std::mutex> run_lock(run_mu_); // one parallel_for at a time
    ensure(nthreads - 1);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn;
      n_ = n;
      parts_ = nthreads;
      next_part_.store(1);
      pending_ = nthreads - 1;
      ++epoch_;
    }
    cv_.notify_all();
    fn(0, n / nthreads); // part 0 on the caller
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

private:
  void ensure(int workers) {
    while ((int)threads_.size() < workers) {
      threads_.emplace_back([this, id = (int)threads_.size()] { loop(id); });
      threads_.back().detach();
    }
  }
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(size_t, size_t)> *fn;
      size_t n;
      int parts;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (id >= parts_ - 1) continue; // not needed this round
        fn = fn_;
        n = n_;
        parts = parts_;
      }
      const int part = next_part_.fetch_add(1);
      if (part < parts) (*fn)(n * part / parts, n * (part + 1) / parts);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      done_cv_.notify_one();
    }
  }
  std::mutex mu_, run_mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> threads_;
  const std::function<void(size_t, size_t)> *fn_ = nullptr;
  size_t n_ = 0;
  int parts_ = 0, pending_ = 0;
  std::atomic<int> next_part_{0};
  uint64_t epoch_ = 0;
};

ExtractParams to_extract_params(const form::HotPathParams &p) {
  ExtractParams e;
  e.neighbor_points = p.neighbor_points;
  e.num_sectors = p.num_sectors;
  e.planar_threshold = p.planar_threshold;
  e.planar_feats_per_sector = p.planar_feats_per_sector;
  e.point_feats_per_sector = p.point_feats_per_sector;
  e.radius = p.radius;
  e.min_points = p.min_points;
  e.min_norm_squared = p.min_norm_squared;
  e.max_norm_squared = p.max_norm_squared;
  e.num_columns = p.num_columns;
  e.num_rows = p.num_rows;
  return e;
}

std::map<uint64_t, Pose3> pose_map(const form::ScanPose *poses, size_t n) {
  std::map<uint64_t, Pose3> m;
  for (size_t i = 0; i < n; ++i) m[poses[i].scan] = poses[i].pose;
  return m;
}

// Matcher::match<I>, matcher.hpp:67-112, rule R6 (match j <-> keypoint j).
template <typename Feat, typename Push>
void match_type(const KeypointMap<Feat> &map, const std::vector<Feat> &keypoints,
                const Pose3 &init, const std::map<uint64_t, Pose3> &poses, double max_dist,
                int nthreads, std::vector<Feat> &queries, std::vector<MatchResult<Feat>> &matches,
                const std::function<void()> &clear_constraints, const Push &push) {
  if (keypoints.empty()) return; // :72-74 - leaves stale matches/constraints
  matches.assign(keypoints.size(), MatchResult<Feat>()); // :77
  queries = keypoints;
  clear_constraints();                                     // :78-80
  const double max_d2 = max_dist * max_dist;               // :82
  Pool::get().parallel_for(keypoints.size(), nthreads, [&](size_t a, size_t b) {
    for (size_t j = a; j < b; ++j) {
      MatchResult<Feat> m = map.find_closest(transform(keypoints[j], init)); // :89
      if (m.found) {
        // :93-96 move the world-frame map point back into its own scan frame
        m.point_local = transform(m.point_local, poses.at(m.scan).inverse());
      }
      matches[j] = m;
    }
  });
  for (size_t j = 0; j < matches.size(); ++j) { // :103-111
    if (matches[j].dist_sqrd < max_d2) push(matches[j], keypoints[j]);
  }
}

} // namespace

OracleHotPath::OracleHotPath(const form::HotPathParams &p)
    : params(p), extract_params(to_extract_params(p)) {
  planar_map.min_dist_map = p.min_dist_map;
  point_map.min_dist_map = p.min_dist_map;
}

int OracleHotPath::threads() const {
  int nt = params.num_threads > 0 ? (int)params.num_threads
                                  : (int)std::thread::hardware_concurrency();
  return nt < 1 ? 1 : nt;
}

void OracleHotPath::extract(const PointXYZf *scan, size_t n, uint64_t scan_idx,
                            std::vector<PlanarFeat> &planar, std::vector<PointFeat> &point) {
  if (!form_oracle::extract(extract_params, scan, n, (size_t)scan_idx, threads(), last_extract))
    throw form::HotPathError("Provided scan does not match the expected size");
  current_scan = scan_idx;
  cur_planar = last_extract.planar;
  cur_point = last_extract.point;
  planar = cur_planar;
  point = cur_point;
}

void OracleHotPath::map_rebuild(const form::ScanPose *poses, size_t n_poses) {
  map_poses = pose_map(poses, n_poses);
  planar_map.to_voxel_map(map_poses, params.max_dist_matching); // form.cpp:61-65
  point_map.to_voxel_map(map_poses, params.max_dist_matching);
}

void OracleHotPath::associate(const Pose3 &pose_k, std::vector<form::PairCount> &counts) {
  auto &mine = constraints[current_scan];
  match_type<PlanarFeat>(
      planar_map, cur_planar, pose_k, map_poses, params.max_dist_matching, threads(),
      planar_match_queries, planar_matches,
      [&] { for (auto &kv : mine) kv.second.plane.clear(); },
      [&](const MatchResult<PlanarFeat> &m, const PlanarFeat &kp) {
        mine[m.scan].plane.push_back(m.point_local, kp);
      });
  match_type<PointFeat>(
      point_map, cur_point, pose_k, map_poses, params.max_dist_matching, threads(),
      point_match_queries, point_matches,
      [&] { for (auto &kv : mine) kv.second.point.clear(); },
      [&](const MatchResult<PointFeat> &m, const PointFeat &kp) {
        mine[m.scan].point.push_back(m.point_local, kp);
      });
  counts.clear();
  for (const auto &[i, c] : mine)
    if (!c.empty())
      counts.push_back({i, (uint32_t)c.plane.num_constraints(), (uint32_t)c.point.num_constraints()});
}

void OracleHotPath::linearize(const form::PairKey *pairs, size_t n_pairs,
                              const form::ScanPose *poses, size_t n_poses, double *out91) {
  const auto pm = pose_map(poses, n_poses);
  static const PairConstraints kEmpty;
  Pool::get().parallel_for(n_pairs, threads(), [&](size_t a, size_t b) {
    for (size_t p = a; p < b; ++p) {
      const PairConstraints *c = &kEmpty;
      auto jt = constraints.find(pairs[p].j);
      if (jt != constraints.end()) {
        auto it = jt->second.find(pairs[p].i);
        if (it != jt->second.end()) c = &it->second;
      }
      linearize_pair(*c, pm.at(pairs[p].i), pm.at(pairs[p].j), params.sigma, out91 + 91 * p);
    }
  });
}

void OracleHotPath::error(const form::PairKey *pairs, size_t n_pairs, const form::ScanPose *poses,
                          size_t n_poses, double *out) {
  const auto pm = pose_map(poses, n_poses);
  static const PairConstraints kEmpty;
  Pool::get().parallel_for(n_pairs, threads(), [&](size_t a, size_t b) {
    for (size_t p = a; p < b; ++p) {
      const PairConstraints *c = &kEmpty;
      auto jt = constraints.find(pairs[p].j);
      if (jt != constraints.end()) {
        auto it = jt->second.find(pairs[p].i);
        if (it != jt->second.end()) c = &it->second;
      }
      out[p] = error_pair(*c, pm.at(pairs[p].i), pm.at(pairs[p].j), params.sigma);
    }
  });
}

void OracleHotPath::commit_scan(size_t &n_planar_added, size_t &n_point_added) {
  // insert_matches infers the scan from matches.front().query.scan (map.tpp:155)
  n_planar_added = n_point_added = 0;
  if (!planar_matches.empty())
    n_planar_added = planar_map.insert_matches(planar_match_queries.front().scan,
                                               planar_match_queries, planar_matches);
  if (!point_matches.empty())
    n_point_added = point_map.insert_matches(point_match_queries.front().scan,
                                             point_match_queries, point_matches);
}

void OracleHotPath::remove_scans(const uint64_t *scans, size_t n) {
  for (size_t s = 0; s < n; ++s) {
    planar_map.remove(scans[s]); // form.cpp:111
    point_map.remove(scans[s]);
    constraints.erase(scans[s]); // constraints.cpp:186-194
    for (auto &kv : constraints) kv.second.erase(scans[s]);
  }
}

void OracleHotPath::world_keypoints(const form::ScanPose *poses, size_t n_poses,
                                    std::vector<PlanarFeat> &planar,
                                    std::vector<PointFeat> &point) {
  const auto pm = pose_map(poses, n_poses);
  planar.clear();
  point.clear();
  for (const auto &[scan, kps] : planar_map.scans())
    for (const auto &kp : kps) planar.push_back(transform(kp, pm.at(scan)));
  for (const auto &[scan, kps] : point_map.scans())
    for (const auto &kp : kps) point.push_back(transform(kp, pm.at(scan)));
}

} // namespace form_oracle
