// Recording and replay of the hot-path call sequence of a run.
//
// bench.py measures "feature + association + linearisation" without the host
// smoother in the timed region: one untimed run of the real pipeline records
// every HotPath call with its pose / pair arguments, and the timed region
// replays exactly those calls (on the CUDA path, and on the CPU baseline).
#pragma once

#include "form/hotpath.hpp"

#include <algorithm>
#include <functional>
#include <utility>
#include <vector>

namespace form {

struct TraceOp {
  enum Kind : int { EXTRACT = 0, MAP_REBUILD, ASSOCIATE, LINEARIZE, ERROR, COMMIT, REMOVE, ASSOC_LIN };
  Kind kind;
  uint64_t scan = 0;            // EXTRACT: scan index
  Pose3 pose;                   // ASSOCIATE
  std::vector<ScanPose> poses;  // MAP_REBUILD, LINEARIZE, ERROR
  std::vector<PairKey> pairs;   // LINEARIZE, ERROR
  std::vector<uint64_t> ids;    // REMOVE
};

struct Trace {
  std::vector<TraceOp> ops;
  std::vector<size_t> scan_begin; // ops index where each EXTRACT (= register_scan) starts
  size_t num_scans() const { return scan_begin.size(); }
  size_t op_end(size_t scan_pos) const {
    return scan_pos + 1 < scan_begin.size() ? scan_begin[scan_pos + 1] : ops.size();
  }
};

/// Decorator that forwards to `inner` and appends every call to `trace`.
class RecordingHotPath : public HotPath {
public:
  RecordingHotPath(HotPath &inner, Trace &trace) : m_inner(inner), m_trace(trace) {}

  void begin_scan() override { m_inner.begin_scan(); }
  void end_scan() override { m_inner.end_scan(); }
  void extract(const PointXYZf *scan, size_t n, uint64_t scan_idx, std::vector<PlanarFeat> &planar,
               std::vector<PointFeat> &point) override {
    m_trace.scan_begin.push_back(m_trace.ops.size());
    TraceOp op;
    op.kind = TraceOp::EXTRACT;
    op.scan = scan_idx;
    m_trace.ops.push_back(std::move(op));
    m_inner.extract(scan, n, scan_idx, planar, point);
  }
  void map_rebuild(const ScanPose *poses, size_t n_poses) override {
    TraceOp op;
    op.kind = TraceOp::MAP_REBUILD;
    op.poses.assign(poses, poses + n_poses);
    m_trace.ops.push_back(std::move(op));
    m_inner.map_rebuild(poses, n_poses);
  }
  void associate(const Pose3 &pose_k, std::vector<PairCount> &counts) override {
    TraceOp op;
    op.kind = TraceOp::ASSOCIATE;
    op.pose = pose_k;
    m_trace.ops.push_back(std::move(op));
    m_inner.associate(pose_k, counts);
  }
  void associate_linearize(uint64_t current_scan, const ScanPose *poses, size_t n_poses,
                           std::vector<PairCount> &counts, std::vector<double> &blocks) override {
    TraceOp op;
    op.kind = TraceOp::ASSOC_LIN;
    op.scan = current_scan;
    op.poses.assign(poses, poses + n_poses);
    m_trace.ops.push_back(std::move(op));
    m_inner.associate_linearize(current_scan, poses, n_poses, counts, blocks);
  }
  void linearize(const PairKey *pairs, size_t n_pairs, const ScanPose *poses, size_t n_poses,
                 double *out91) override {
    TraceOp op;
    op.kind = TraceOp::LINEARIZE;
    op.pairs.assign(pairs, pairs + n_pairs);
    op.poses.assign(poses, poses + n_poses);
    m_trace.ops.push_back(std::move(op));
    m_inner.linearize(pairs, n_pairs, poses, n_poses, out91);
  }
  void error(const PairKey *pairs, size_t n_pairs, const ScanPose *poses, size_t n_poses,
             double *out) override {
    TraceOp op;
    op.kind = TraceOp::ERROR;
    op.pairs.assign(pairs, pairs + n_pairs);
    op.poses.assign(poses, poses + n_poses);
    m_trace.ops.push_back(std::move(op));
    m_inner.error(pairs, n_pairs, poses, n_poses, out);
  }
  void commit_scan(size_t &a, size_t &b) override {
    TraceOp op;
    op.kind = TraceOp::COMMIT;
    m_trace.ops.push_back(std::move(op));
    m_inner.commit_scan(a, b);
  }
  void remove_scans(const uint64_t *scans, size_t n) override {
    TraceOp op;
    op.kind = TraceOp::REMOVE;
    op.ids.assign(scans, scans + n);
    m_trace.ops.push_back(std::move(op));
    m_inner.remove_scans(scans, n);
  }
  void world_keypoints(const ScanPose *poses, size_t n_poses, std::vector<PlanarFeat> &planar,
                       std::vector<PointFeat> &point) override {
    m_inner.world_keypoints(poses, n_poses, planar, point);
  }

private:
  HotPath &m_inner;
  Trace &m_trace;
};

/// Work counters of a replay (inputs of the algorithmic-bytes formula).
struct ReplayStats {
  uint64_t scans = 0, points = 0;
  uint64_t planar_kp = 0, point_kp = 0;
  uint64_t assoc_calls = 0, assoc_queries = 0;
  uint64_t map_rebuilds = 0, map_points = 0;
  uint64_t lin_calls = 0, lin_pairs = 0, lin_planar = 0, lin_point = 0;
  uint64_t err_calls = 0, err_pairs = 0, err_planar = 0, err_point = 0;
  uint64_t novel_planar = 0, novel_point = 0;
  uint64_t assoc_planar = 0, assoc_point = 0; // correspondences accepted by the associations
  double checksum = 0.0; // sum of every returned block / error, to keep the work observable
  // correspondence counts per live pair (carried across replay() calls)
  std::vector<std::pair<PairKey, std::pair<uint32_t, uint32_t>>> table;
};

/// Replays ops of scans [first, last) of `trace` on `hp`.  `do_extract` performs
/// the EXTRACT op (host or device-resident scan) and returns the keypoint counts.
inline void replay(const Trace &trace, HotPath &hp, size_t first, size_t last,
                   const std::function<void(uint64_t, size_t &, size_t &)> &do_extract,
                   size_t points_per_scan, ReplayStats &st) {
  std::vector<PairCount> counts;
  std::vector<double> out;
  auto &table = st.table; // correspondence counts per pair, to account linearisation work
  auto lookup = [&](const PairKey &p) -> std::pair<uint32_t, uint32_t> {
    for (const auto &e : table)
      if (e.first.i == p.i && e.first.j == p.j) return e.second;
    return {0, 0};
  };
  uint64_t cur_scan = 0;
  size_t cur_np = 0, cur_nq = 0;
  for (size_t s = first; s < last && s < trace.num_scans(); ++s) {
    for (size_t o = trace.scan_begin[s]; o < trace.op_end(s); ++o) {
      const TraceOp &op = trace.ops[o];
      switch (op.kind) {
      case TraceOp::EXTRACT:
        cur_scan = op.scan;
        do_extract(op.scan, cur_np, cur_nq);
        st.scans += 1;
        st.points += points_per_scan;
        st.planar_kp += cur_np;
        st.point_kp += cur_nq;
        break;
      case TraceOp::MAP_REBUILD:
        hp.map_rebuild(op.poses.data(), op.poses.size());
        st.map_rebuilds += 1;
        break;
      case TraceOp::ASSOCIATE: {
        hp.associate(op.pose, counts);
        st.assoc_calls += 1;
        st.assoc_queries += cur_np + cur_nq;
        // refresh the rows of the current scan
        table.erase(std::remove_if(table.begin(), table.end(),
                                   [&](const auto &e) { return e.first.j == cur_scan; }),
                    table.end());
        for (const auto &c : counts) {
          table.push_back({{c.i, cur_scan}, {c.n_planar, c.n_point}});
          st.assoc_planar += c.n_planar;
          st.assoc_point += c.n_point;
        }
        break;
      }
      case TraceOp::ASSOC_LIN: {
        hp.associate_linearize(op.scan, op.poses.data(), op.poses.size(), counts, out);
        st.assoc_calls += 1;
        st.assoc_queries += cur_np + cur_nq;
        table.erase(std::remove_if(table.begin(), table.end(),
                                   [&](const auto &e) { return e.first.j == cur_scan; }),
                    table.end());
        for (const auto &c : counts) {
          table.push_back({{c.i, cur_scan}, {c.n_planar, c.n_point}});
          st.assoc_planar += c.n_planar;
          st.assoc_point += c.n_point;
        }
        st.lin_calls += 1;
        st.lin_pairs += counts.size();
        for (size_t p = 0; p < counts.size(); ++p) {
          st.lin_planar += counts[p].n_planar;
          st.lin_point += counts[p].n_point;
          st.checksum += out[91 * p + 90];
        }
        break;
      }
      case TraceOp::LINEARIZE: {
        out.resize(91 * op.pairs.size());
        hp.linearize(op.pairs.data(), op.pairs.size(), op.poses.data(), op.poses.size(), out.data());
        st.lin_calls += 1;
        st.lin_pairs += op.pairs.size();
        for (const auto &p : op.pairs) {
          const auto c = lookup(p);
          st.lin_planar += c.first;
          st.lin_point += c.second;
        }
        for (size_t p = 0; p < op.pairs.size(); ++p) st.checksum += out[91 * p + 90];
        break;
      }
      case TraceOp::ERROR: {
        out.resize(op.pairs.size());
        hp.error(op.pairs.data(), op.pairs.size(), op.poses.data(), op.poses.size(), out.data());
        st.err_calls += 1;
        st.err_pairs += op.pairs.size();
        for (const auto &p : op.pairs) {
          const auto c = lookup(p);
          st.err_planar += c.first;
          st.err_point += c.second;
        }
        for (size_t p = 0; p < op.pairs.size(); ++p) st.checksum += out[p];
        break;
      }
      case TraceOp::COMMIT: {
        size_t a = 0, b = 0;
        hp.commit_scan(a, b);
        st.novel_planar += a;
        st.novel_point += b;
        st.map_points += a + b;
        break;
      }
      case TraceOp::REMOVE:
        hp.remove_scans(op.ids.data(), op.ids.size());
        table.erase(std::remove_if(table.begin(), table.end(),
                                   [&](const auto &e) {
                                     for (uint64_t id : op.ids)
                                       if (e.first.i == id || e.first.j == id) return true;
                                     return false;
                                   }),
                    table.end());
        break;
      }
    }
  }
}

} // namespace form
