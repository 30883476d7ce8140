// Host-side fixed-lag smoother: owns the pose estimates, the prior / marginal
// factors and the bookkeeping of which scan pairs carry correspondences, and
// runs dense Levenberg-Marquardt over the window.
//
// Mirrors form::ConstraintManager (/root/reference/form/optimization/
// constraints.hpp:50-168, constraints.cpp:39-336) with the same public method
// names.  The reference delegates to GTSAM; GTSAM is not available to this
// build, so the pieces it uses are restated here with GTSAM 4.3's published
// semantics ([external]):
//   * LevenbergMarquardtParams defaults: lambda0 1e-5, factor 10 (fixed),
//     lambda in [0, 1e5], additive damping lambda*I, minModelFidelity 1e-3,
//     maxIterations 100, relativeErrorTol 1e-5, absoluteErrorTol 1e-5, errorTol 0;
//   * dense solve of the damped normal equations (gtsam.hpp:49-53);
//   * HessianFactor error 0.5 (f - 2 g.d + d.G.d);
//   * LinearContainerFactor: re-centre a stored quadratic on the current values;
//   * partial elimination = Schur complement onto the kept poses.
// What is NOT here is the hot path: every FeatureFactor linearisation and error
// evaluation goes through HotPath (the CUDA library in the product), exactly at
// the places the reference calls DenseFactor::linearize / NoiseModelFactor::error
// under the pair-set policy of get_graph (constraints.cpp:252-308).
#pragma once

#include "form/dense.hpp"
#include "form/hotpath.hpp"

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <vector>

namespace form {

using Values = std::map<uint64_t, Pose3>; // key order = gtsam::Ordering::NATURAL

/// gtsam::LevenbergMarquardtParams (defaults, constraints.hpp:66-68).
struct LMParams {
  double lambdaInitial = 1e-5;
  double lambdaFactor = 10.0;
  double lambdaUpperBound = 1e5;
  double lambdaLowerBound = 0.0;
  double minModelFidelity = 1e-3;
  size_t maxIterations = 100;
  double relativeErrorTol = 1e-5;
  double absoluteErrorTol = 1e-5;
  double errorTol = 0.0;
};

/// HessianFactor over a list of pose keys wrapped as a LinearContainerFactor:
/// a quadratic in the tangent space of `lin_points`.
struct LinearContainer {
  std::vector<uint64_t> keys;
  std::vector<Pose3> lin_points;
  dense::Quadratic q; // n = 6 * keys.size()

  /// delta0 = lin_points.localCoordinates(values)
  std::vector<double> delta0(const Values &values) const {
    std::vector<double> d(6 * keys.size());
    for (size_t k = 0; k < keys.size(); ++k) {
      const Vec6 l = lin_points[k].localCoordinates(values.at(keys[k]));
      for (int a = 0; a < 6; ++a) d[6 * k + a] = l[a];
    }
    return d;
  }
  double error(const Values &values) const {
    const std::vector<double> d = delta0(values);
    return q.error(d.data());
  }
  /// quadratic re-centred on `values`: G' = G, g' = g - G d0, f' = f - 2 g.d0 + d0.G.d0
  dense::Quadratic linearize(const Values &values) const {
    const std::vector<double> d = delta0(values);
    dense::Quadratic o;
    o.n = q.n;
    o.G = q.G;
    o.g = q.g;
    for (size_t r = 0; r < q.n; ++r) {
      double s = 0.0;
      for (size_t c = 0; c < q.n; ++c) s += q.G[r * q.n + c] * d[c];
      o.g[r] -= s;
    }
    o.f = 2.0 * q.error(d.data());
    return o;
  }
  bool involves(uint64_t key) const { return std::find(keys.begin(), keys.end(), key) != keys.end(); }
};

/// gtsam::PriorFactor<Pose3> with an isotropic sigma (constraints.cpp:218-220).
struct PosePrior {
  uint64_t key;
  Pose3 prior;
  double sigma;
};

/// Counters for reporting / trace analysis.
struct SmootherStats {
  size_t optimize_calls = 0;
  size_t lm_iterations = 0;
  size_t linearize_calls = 0;
  size_t error_calls = 0;
  size_t linearized_pairs = 0;
  size_t error_pairs = 0;
};

class ConstraintManager {
public:
  struct Params {
    bool disable_smoothing = false;
    double planar_constraint_sigma = 0.1;
    double pose_sigma = 1e-3; // Isotropic::Sigma(6, 1e-3), constraints.hpp:65
    LMParams opt_params;
    /// Host scheduling of the hot-path calls inside LM.  false: GTSAM's schedule - one
    /// linearisation per outer iteration plus one error evaluation per trial step.
    /// true: every trial step is LINEARISED instead (its error is 0.5 * f of the returned
    /// blocks, the same quantity), and an accepted trial's blocks are reused as the next
    /// iteration's linearisation - one device round trip per LM step instead of two.
    /// Same iterates either way (up to the rounding of 0.5 * f vs the error kernel).
    bool fused_trial_linearization = true;
  };
  using PairCounts = std::map<uint64_t, std::pair<uint32_t, uint32_t>>; // i -> (planar, point)

  ConstraintManager() = default;
  explicit ConstraintManager(const Params &params) : m_params(params) {}

  void set_hotpath(HotPath *hp) { m_hotpath = hp; }

  // ------------------------- Doers ------------------------- //
  /// constraints.cpp:71-101
  Pose3 predict_next() const noexcept {
    if (!initialized()) return Pose3::Identity();
    const size_t scan = m_scan + 1;
    const bool prev = scan > 0 && m_values.count(scan - 1);
    const bool prev_prev = scan > 1 && m_values.count(scan - 2);
    if (prev && prev_prev) {
      const Pose3 &p1 = m_values.at(scan - 1), &p2 = m_values.at(scan - 2);
      const Pose3 pred = p1 * (p2.inverse() * p1);
      return pred.normalized();
    }
    if (prev) return m_values.at(scan - 1);
    return Pose3::Identity();
  }

  /// constraints.cpp:206-223: returns the new scan index.
  size_t step(const Pose3 &pose) noexcept {
    if (initialized()) ++m_scan;
    m_values[m_scan] = pose;
    m_fast_linear.reset();
    if (m_scan == 0) m_priors.push_back(PosePrior{0, pose, m_params.pose_sigma});
    m_counts[m_scan]; // get_current_constraints()
    return m_scan;
  }

  /// Result of Matcher::match for the current scan (matcher.hpp:103-111 fills
  /// m_constraints[m_scan][i]).
  void set_current_counts(const std::vector<PairCount> &counts) {
    PairCounts &mine = m_counts[m_scan];
    mine.clear();
    for (const auto &c : counts) mine[c.i] = {c.n_planar, c.n_point};
  }

  bool fused_schedule() const noexcept { return m_params.fused_trial_linearization; }

  /// constraints.cpp:103-118.  `first_blocks` (fused schedule only): the blocks of the
  /// current scan's pairs at the current values, already obtained together with the
  /// association (HotPath::associate_linearize), in ascending scan order.
  Values optimize(bool fast = false, const std::vector<double> *first_blocks = nullptr) {
    ++m_stats.optimize_calls;
    Graph g = m_params.disable_smoothing ? get_single_graph() : get_graph(fast);
    Values init;
    if (m_params.disable_smoothing) init[m_scan] = m_values.at(m_scan);
    else init = m_values;
    const bool usable = first_blocks && (fast || m_params.disable_smoothing) &&
                        first_blocks->size() == 91 * g.pairs.size();
    return levenberg_marquardt(g, init, usable ? first_blocks->data() : nullptr);
  }

  /// constraints.cpp:120-195
  void marginalize(const std::vector<ScanIndex> &scans) {
    if (scans.empty()) return;
    auto is_marg = [&](uint64_t k) { return std::find(scans.begin(), scans.end(), k) != scans.end(); };

    // collect the factors touching the marginalised poses
    std::vector<const PosePrior *> priors;
    std::vector<std::shared_ptr<LinearContainer>> containers;
    for (auto it = m_priors.begin(); it != m_priors.end();) {
      if (is_marg(it->key)) {
        m_dropped_priors.push_back(*it);
        it = m_priors.erase(it);
      } else ++it;
    }
    for (const auto &p : m_dropped_priors) priors.push_back(&p);
    for (size_t s = 0; s < m_marginals.size(); ++s) {
      if (!m_marginals[s]) continue;
      bool hit = false;
      for (uint64_t k : m_marginals[s]->keys) hit = hit || is_marg(k);
      if (hit) {
        containers.push_back(m_marginals[s]);
        m_marginals[s].reset();
        m_empty_slots.push_back(s);
      }
    }
    std::vector<PairKey> pairs;
    for (const auto &[j, row] : m_counts)
      for (const auto &[i, c] : row)
        if ((c.first || c.second) && (is_marg(i) || is_marg(j))) pairs.push_back({i, j});

    // involved keys: marginalised first, then the rest, each ascending
    std::vector<uint64_t> keys;
    auto add_key = [&](uint64_t k) {
      if (std::find(keys.begin(), keys.end(), k) == keys.end()) keys.push_back(k);
    };
    for (const auto *p : priors) add_key(p->key);
    for (const auto &c : containers)
      for (uint64_t k : c->keys) add_key(k);
    for (const auto &p : pairs) {
      add_key(p.i);
      add_key(p.j);
    }
    std::vector<uint64_t> mk, rk;
    for (uint64_t k : keys) (is_marg(k) ? mk : rk).push_back(k);
    std::sort(mk.begin(), mk.end());
    std::sort(rk.begin(), rk.end());
    std::vector<uint64_t> order = mk;
    order.insert(order.end(), rk.begin(), rk.end());

    if (!order.empty() && !rk.empty()) {
      // linearise the dropped factors at the current values (constraints.cpp:164)
      Graph g;
      g.priors = priors;
      for (const auto &c : containers) g.containers.push_back(c.get());
      g.pairs = pairs;
      dense::Quadratic sys = linearize_graph(g, m_values, order);
      // Schur complement onto the kept poses (eliminatePartialMultifrontal, :165-166)
      const size_t na = 6 * mk.size(), nc = 6 * rk.size(), n = na + nc;
      auto lc = std::make_shared<LinearContainer>();
      lc->keys = rk;
      for (uint64_t k : rk) lc->lin_points.push_back(m_values.at(k));
      lc->q.resize(nc);
      if (na == 0) {
        lc->q = sys;
      } else {
        std::vector<double> A(na * na);
        for (size_t r = 0; r < na; ++r)
          for (size_t c = 0; c < na; ++c) A[r * na + c] = sys.G[r * n + c];
        bool ok = dense::cholesky(A, na);
        for (double jitter = 1e-9; !ok && jitter < 1.0; jitter *= 100.0) {
          // rank-deficient marginal (e.g. a pose seen only through planes): regularise
          for (size_t r = 0; r < na; ++r)
            for (size_t c = 0; c < na; ++c) A[r * na + c] = sys.G[r * n + c] + (r == c ? jitter : 0.0);
          ok = dense::cholesky(A, na);
        }
        // X = A^-1 [B | ga]
        std::vector<double> X((nc + 1) * na);
        for (size_t c = 0; c < nc; ++c) {
          double *col = &X[c * na];
          for (size_t r = 0; r < na; ++r) col[r] = sys.G[r * n + (na + c)];
          if (ok) dense::cholesky_solve(A, na, col);
        }
        double *xa = &X[nc * na];
        for (size_t r = 0; r < na; ++r) xa[r] = sys.g[r];
        if (ok) dense::cholesky_solve(A, na, xa);
        for (size_t r = 0; r < nc; ++r) {
          for (size_t c = 0; c < nc; ++c) {
            double s = sys.G[(na + r) * n + (na + c)];
            for (size_t k = 0; k < na; ++k) s -= sys.G[k * n + (na + r)] * X[c * na + k];
            lc->q.G[r * nc + c] = s;
          }
          double s = sys.g[na + r];
          for (size_t k = 0; k < na; ++k) s -= sys.G[k * n + (na + r)] * xa[k];
          lc->q.g[r] = s;
        }
        double f = sys.f;
        for (size_t k = 0; k < na; ++k) f -= sys.g[k] * xa[k];
        lc->q.f = f;
        // symmetrise against round-off
        for (size_t r = 0; r < nc; ++r)
          for (size_t c = r + 1; c < nc; ++c) {
            const double v = 0.5 * (lc->q.G[r * nc + c] + lc->q.G[c * nc + r]);
            lc->q.G[r * nc + c] = lc->q.G[c * nc + r] = v;
          }
      }
      // put the marginal factor back (constraints.cpp:171-178)
      if (m_empty_slots.empty()) {
        m_marginals.push_back(lc);
      } else {
        m_marginals[m_empty_slots.back()] = lc;
        m_empty_slots.pop_back();
      }
    }
    m_dropped_priors.clear();

    for (uint64_t k : scans) m_values.erase(k); // :181-183
    for (uint64_t f : scans) {                  // :186-194
      m_counts.erase(f);
      for (auto &kv : m_counts) kv.second.erase(f);
    }
    if (m_hotpath) m_hotpath->remove_scans(scans.data(), scans.size());
  }

  // ------------------------- Setters ------------------------- //
  void update_values(const Values &values) noexcept { // constraints.cpp:198-204
    for (const auto &kv : values) m_values[kv.first] = kv.second;
  }
  void update_pose(const ScanIndex &scan, const Pose3 &pose) noexcept { m_values[scan] = pose; }
  void update_current_pose(const Pose3 &pose) noexcept { update_pose(m_scan, pose); }

  // ------------------------- Getters ------------------------- //
  bool initialized() const noexcept { return !m_values.empty(); }
  Pose3 get_pose(const ScanIndex &scan) const { return m_values.at(scan); }
  Pose3 get_current_pose() const { return m_values.at(m_scan); }
  const Values &get_values() const noexcept { return m_values; }
  size_t current_scan() const noexcept { return m_scan; }
  const SmootherStats &stats() const noexcept { return m_stats; }
  const std::map<uint64_t, PairCounts> &pair_counts() const noexcept { return m_counts; }
  /// The live marginal factors (the LinearContainerFactors of m_other_factors).
  std::vector<std::shared_ptr<const LinearContainer>> marginals() const {
    std::vector<std::shared_ptr<const LinearContainer>> out;
    for (const auto &m : m_marginals)
      if (m) out.push_back(m);
    return out;
  }

  /// constraints.cpp:319-336
  size_t num_recent_connections(const ScanIndex &scan, const ScanIndex &oldest) const noexcept {
    size_t count = 0;
    for (const auto &[j, row] : m_counts) {
      if (j < oldest) continue;
      auto it = row.find(scan);
      if (it != row.end()) count += it->second.first + it->second.second;
    }
    return count;
  }

  std::vector<ScanPose> scan_poses(const Values &values) const {
    std::vector<ScanPose> out;
    out.reserve(values.size());
    for (const auto &kv : values) out.push_back({kv.first, kv.second});
    return out;
  }

private:
  /// What one optimisation sees: prior(s), linear containers, and the FeatureFactor
  /// pairs the hot path evaluates.
  struct Graph {
    std::vector<const PosePrior *> priors;
    std::vector<const LinearContainer *> containers;
    std::vector<PairKey> pairs;
    bool single = false; // ablation: only X(m_scan) is a variable (BinaryFactorWrapper)
  };

  /// constraints.cpp:252-308
  Graph get_graph(bool fast) {
    Graph g;
    for (const auto &p : m_priors) g.priors.push_back(&p);
    for (const auto &m : m_marginals)
      if (m) g.containers.push_back(m.get());
    if (fast) {
      for (const auto &[i, c] : m_counts.at(m_scan))
        if (c.first || c.second) g.pairs.push_back({i, m_scan});
      if (!m_fast_linear) {
        // one-off linearisation of every previous pair at the current values (:268-288)
        std::vector<PairKey> prev;
        for (const auto &[j, row] : m_counts) {
          if (j == m_scan) continue;
          for (const auto &[i, c] : row)
            if (c.first || c.second) prev.push_back({i, j});
        }
        Graph pg;
        pg.pairs = prev;
        std::vector<uint64_t> order;
        for (const auto &kv : m_values) order.push_back(kv.first);
        LinearContainer lc;
        lc.keys = order;
        for (uint64_t k : order) lc.lin_points.push_back(m_values.at(k));
        lc.q = linearize_graph(pg, m_values, order);
        m_fast_linear = std::move(lc);
      }
      g.containers.push_back(&*m_fast_linear);
    } else {
      for (const auto &[j, row] : m_counts)
        for (const auto &[i, c] : row)
          if (c.first || c.second) g.pairs.push_back({i, j});
    }
    return g;
  }

  /// constraints.cpp:235-250
  Graph get_single_graph() {
    Graph g;
    g.single = true;
    for (const auto &[i, c] : m_counts.at(m_scan))
      if (c.first || c.second) g.pairs.push_back({i, m_scan});
    return g;
  }

  static size_t index_of(const std::vector<uint64_t> &order, uint64_t key) {
    return (size_t)(std::find(order.begin(), order.end(), key) - order.begin());
  }

  /// All poses the hot path may need: optimisation values over the stored ones.
  std::vector<ScanPose> merged_poses(const Values &values) const {
    Values all = m_values;
    for (const auto &kv : values) all[kv.first] = kv.second;
    return scan_poses(all);
  }

  /// NonlinearFactorGraph::linearize + dense assembly in `order`.
  dense::Quadratic linearize_graph(const Graph &g, const Values &values,
                                   const std::vector<uint64_t> &order,
                                   const double *preloaded_blocks = nullptr) {
    dense::Quadratic sys;
    const size_t n = 6 * order.size();
    sys.resize(n);
    for (const PosePrior *p : g.priors) {
      const size_t o = 6 * index_of(order, p->key);
      const Pose3 between = p->prior.inverse() * values.at(p->key);
      const Vec6 e = Pose3::Logmap(between);
      const Mat6 J = Pose3::LogmapDerivative(between);
      const double w = 1.0 / (p->sigma * p->sigma);
      for (int r = 0; r < 6; ++r) {
        double gr = 0.0;
        for (int k = 0; k < 6; ++k) gr += J[6 * k + r] * e[k];
        sys.g[o + r] -= w * gr;
        for (int c = 0; c < 6; ++c) {
          double s = 0.0;
          for (int k = 0; k < 6; ++k) s += J[6 * k + r] * J[6 * k + c];
          sys.G[(o + r) * n + (o + c)] += w * s;
        }
      }
      double ee = 0.0;
      for (int k = 0; k < 6; ++k) ee += e[k] * e[k];
      sys.f += w * ee;
    }
    for (const LinearContainer *c : g.containers) {
      const dense::Quadratic q = c->linearize(values);
      std::vector<size_t> off(c->keys.size());
      for (size_t k = 0; k < c->keys.size(); ++k) off[k] = 6 * index_of(order, c->keys[k]);
      for (size_t a = 0; a < c->keys.size(); ++a)
        for (int r = 0; r < 6; ++r) {
          sys.g[off[a] + r] += q.g[6 * a + r];
          for (size_t b = 0; b < c->keys.size(); ++b)
            for (int cc = 0; cc < 6; ++cc)
              sys.G[(off[a] + r) * n + (off[b] + cc)] += q.G[(6 * a + r) * q.n + (6 * b + cc)];
        }
      sys.f += q.f;
    }
    if (!g.pairs.empty()) {
      std::vector<double> blocks(91 * g.pairs.size());
      if (preloaded_blocks) {
        std::copy(preloaded_blocks, preloaded_blocks + blocks.size(), blocks.begin());
      } else {
        const std::vector<ScanPose> poses = merged_poses(values);
        m_hotpath->linearize(g.pairs.data(), g.pairs.size(), poses.data(), poses.size(), blocks.data());
      }
      ++m_stats.linearize_calls;
      m_stats.linearized_pairs += g.pairs.size();
      for (size_t p = 0; p < g.pairs.size(); ++p) {
        const double *b = &blocks[91 * p];
        // unpack the 13x13 upper triangle
        double M[13][13];
        size_t e = 0;
        for (int r = 0; r < 13; ++r)
          for (int c = r; c < 13; ++c) M[r][c] = M[c][r] = b[e++];
        const bool has_i = !g.single;
        const size_t oi = has_i ? 6 * index_of(order, g.pairs[p].i) : 0;
        const size_t oj = 6 * index_of(order, g.pairs[p].j);
        for (int r = 0; r < 6; ++r) {
          for (int c = 0; c < 6; ++c) {
            if (has_i) {
              sys.G[(oi + r) * n + (oi + c)] += M[r][c];
              sys.G[(oi + r) * n + (oj + c)] += M[r][6 + c];
              sys.G[(oj + r) * n + (oi + c)] += M[6 + r][c];
            }
            sys.G[(oj + r) * n + (oj + c)] += M[6 + r][6 + c];
          }
          if (has_i) sys.g[oi + r] += M[r][12];
          sys.g[oj + r] += M[6 + r][12];
        }
        sys.f += M[12][12];
      }
    }
    return sys;
  }

  /// NonlinearFactorGraph::error
  double graph_error(const Graph &g, const Values &values) {
    double err = 0.0;
    for (const PosePrior *p : g.priors) {
      const Vec6 e = p->prior.localCoordinates(values.at(p->key));
      double ee = 0.0;
      for (int k = 0; k < 6; ++k) ee += e[k] * e[k];
      err += 0.5 * ee / (p->sigma * p->sigma);
    }
    for (const LinearContainer *c : g.containers) err += c->error(values);
    if (!g.pairs.empty()) {
      std::vector<double> e(g.pairs.size());
      const std::vector<ScanPose> poses = merged_poses(values);
      m_hotpath->error(g.pairs.data(), g.pairs.size(), poses.data(), poses.size(), e.data());
      ++m_stats.error_calls;
      m_stats.error_pairs += g.pairs.size();
      for (double v : e) err += v;
    }
    return err;
  }

  /// DenseLMOptimizer::optimize (gtsam.hpp:40-54) with GTSAM's LM schedule.
  Values levenberg_marquardt(const Graph &g, const Values &initial,
                             const double *first_blocks = nullptr) {
    const LMParams &P = m_params.opt_params;
    const bool fused = m_params.fused_trial_linearization;
    Values values = initial;
    std::vector<uint64_t> order;
    for (const auto &kv : values) order.push_back(kv.first);
    const size_t n = 6 * order.size();
    dense::Quadratic lin;      // linearisation at `values` (valid when have_lin)
    bool have_lin = false;
    double error;
    if (fused) {
      lin = linearize_graph(g, values, order, first_blocks);
      have_lin = true;
      error = 0.5 * lin.f;
    } else {
      error = graph_error(g, values);
    }
    double lambda = P.lambdaInitial;
    size_t iterations = 0;

    auto converged = [&](double cur, double next) {
      if (next <= P.errorTol) return true;
      const double abs_dec = cur - next;
      const double rel_dec = abs_dec / cur;
      return (P.relativeErrorTol != 0.0 && rel_dec <= P.relativeErrorTol) || abs_dec <= P.absoluteErrorTol;
    };

    if (error <= P.errorTol || iterations >= P.maxIterations) return values;
    double new_error = error, current_error;
    do {
      current_error = new_error;
      // ---- iterate(): linearise once, then search lambda ----
      if (!have_lin) lin = linearize_graph(g, values, order);
      have_lin = true;
      for (;;) {
        // buildDampedSystem: + lambda * I (diagonalDamping = false)
        std::vector<double> A = lin.G;
        for (size_t r = 0; r < n; ++r) A[r * n + r] += lambda;
        std::vector<double> delta = lin.g;
        bool step_ok = false, stop_search = false;
        double trial_error = std::numeric_limits<double>::infinity();
        Values trial;
        dense::Quadratic trial_lin;
        double fidelity = 0.0;
        if (dense::cholesky(A, n)) {
          dense::cholesky_solve(A, n, delta.data());
          const double old_lin = 0.5 * lin.f;
          const double new_lin = lin.error(delta.data());
          const double lin_change = old_lin - new_lin;
          if (lin_change >= 0.0) {
            trial = values;
            for (size_t k = 0; k < order.size(); ++k) {
              Vec6 d;
              for (int a = 0; a < 6; ++a) d[a] = delta[6 * k + a];
              trial[order[k]] = values.at(order[k]).retract(d);
            }
            if (fused) {
              trial_lin = linearize_graph(g, trial, order);
              trial_error = 0.5 * trial_lin.f;
            } else {
              trial_error = graph_error(g, trial);
            }
            const double cost_change = error - trial_error;
            if (lin_change > std::numeric_limits<double>::epsilon() * old_lin) {
              fidelity = cost_change / lin_change;
              step_ok = fidelity > P.minModelFidelity;
            }
            // GTSAM's tryLambda tests this after (not instead of) the fidelity block: a
            // REJECTED step whose cost change is tiny ends the lambda search
            if (std::fabs(cost_change) < P.relativeErrorTol * error) stop_search = true;
          }
        }
        if (step_ok) {
          values = std::move(trial);
          error = trial_error;
          if (fused) lin = std::move(trial_lin); // already the linearisation at the new values
          else have_lin = false;
          lambda = std::max(P.lambdaLowerBound, lambda / P.lambdaFactor);
          ++iterations;
          ++m_stats.lm_iterations;
          break;
        } else if (!stop_search) {
          lambda *= P.lambdaFactor;
          if (lambda >= P.lambdaUpperBound) break; // give up at maximum lambda
        } else {
          break;
        }
      }
      new_error = error;
    } while (iterations < P.maxIterations && !converged(current_error, new_error) &&
             std::isfinite(current_error));
    return values;
  }

  Params m_params;
  HotPath *m_hotpath = nullptr;
  Values m_values;
  ScanIndex m_scan = 0;
  std::vector<PosePrior> m_priors;                          // m_other_factors: priors
  std::vector<PosePrior> m_dropped_priors;
  std::vector<std::shared_ptr<LinearContainer>> m_marginals; // m_other_factors: marginals
  std::vector<size_t> m_empty_slots;
  std::optional<LinearContainer> m_fast_linear;
  std::map<uint64_t, PairCounts> m_counts; // m_constraints[j][i] sizes
  SmootherStats m_stats;
};

} // namespace form
