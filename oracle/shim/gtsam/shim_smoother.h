// TEST INFRASTRUCTURE - not GTSAM.  The slice of GTSAM's nonlinear-optimisation machinery that
// form/optimization/constraints.{hpp,cpp} and form/form.cpp touch, so that FORM's OWN smoother and
// Estimator::register_scan compile unmodified and run with their own control flow (which factors
// enter which graph, what is marginalised when, the ICP loop and its exit test, key-scan coupling).
// What stands in for GTSAM here is its PUBLISHED behaviour [external, SURVEY App. B]:
//   * NonlinearFactorGraph: a list of shared factors (null slots allowed); linearize() = every
//     factor's linearize(); error() = sum of the factors' errors;
//   * PriorFactor<Pose3> (addPrior): error vector Logmap(prior^-1 x), Jacobian LogmapDerivative;
//   * HessianFactor(GaussianFactorGraph) / optimizeDensely(): the factors' augmented information
//     matrices summed into one dense system over the keys in ascending order, solved by Cholesky;
//   * eliminatePartialMultifrontal(keys): the Schur complement onto the remaining keys, returned
//     as ONE factor (GTSAM may return it as several; their sum is the same quadratic);
//   * LinearContainerFactor: a stored quadratic with its linearisation points, re-centred on the
//     current values in linearize() / error();
//   * LevenbergMarquardtOptimizer: defaults lambda0 1e-5, factor 10, bounds [0, 1e5], additive
//     damping, minModelFidelity 1e-3, 100 iterations, rel/abs tolerance 1e-5; iterate() =
//     linearise once, tryLambda() until it says stop; virtual solve() (DenseLMOptimizer overrides it).
// Every variable is a Pose3 (dimension 6).  The dense algebra is form_b200/host/form/dense.hpp.
#pragma once

#include <gtsam/inference/Ordering.h>
#include <gtsam/inference/Symbol.h>
#include <gtsam/linear/NoiseModel.h>

#include "form/dense.hpp" // form_b200/host/form/dense.hpp (namespace form::dense; the reference has no such name)

#include <algorithm>
#include <limits>
#include <set>
#include <stdexcept>

namespace gtsam {

class IndeterminantLinearSystemException : public std::runtime_error {
public:
  IndeterminantLinearSystemException() : std::runtime_error("indeterminant linear system") {}
};

namespace shim {
inline size_t index_of(const KeyVector &order, Key k) {
  return (size_t)(std::find(order.begin(), order.end(), k) - order.begin());
}
/// adds a factor's augmented information matrix into a dense system over `order`
inline void add_factor(const HessianFactor &f, const KeyVector &order, form::dense::Quadratic &sys) {
  const size_t n = sys.n, m = 6 * f.keys.size();
  std::vector<size_t> off(f.keys.size());
  for (size_t a = 0; a < f.keys.size(); ++a) off[a] = 6 * index_of(order, f.keys[a]);
  for (size_t a = 0; a < f.keys.size(); ++a)
    for (int r = 0; r < 6; ++r) {
      sys.g[off[a] + r] += f.info(6 * a + r, m);
      for (size_t b = 0; b < f.keys.size(); ++b)
        for (int c = 0; c < 6; ++c) sys.G[(off[a] + r) * n + off[b] + c] += f.info(6 * a + r, 6 * b + c);
    }
  sys.f += f.info(m, m);
}
inline HessianFactor to_factor(const KeyVector &keys, const form::dense::Quadratic &q) {
  HessianFactor f;
  f.keys = keys;
  const size_t n = q.n;
  f.info = Matrix(n + 1, n + 1);
  for (size_t r = 0; r < n; ++r) {
    for (size_t c = 0; c < n; ++c) f.info(r, c) = q.G[r * n + c];
    f.info(r, n) = q.g[r];
    f.info(n, r) = q.g[r];
  }
  f.info(n, n) = q.f;
  return f;
}
} // namespace shim

class BayesTreeStub {};

class GaussianFactorGraph {
public:
  std::vector<std::shared_ptr<GaussianFactor>> factors;
  void push_back(const std::shared_ptr<GaussianFactor> &f) { factors.push_back(f); }
  size_t size() const { return factors.size(); }
  auto begin() const { return factors.begin(); }
  auto end() const { return factors.end(); }

  KeyVector keys() const { // ascending
    std::set<Key> s;
    for (const auto &f : factors)
      for (Key k : hessian(f).keys) s.insert(k);
    return KeyVector(s.begin(), s.end());
  }
  /// the whole graph as one dense quadratic over `order`
  form::dense::Quadratic dense(const KeyVector &order) const {
    form::dense::Quadratic sys;
    sys.resize(6 * order.size());
    for (const auto &f : factors) shim::add_factor(hessian(f), order, sys);
    return sys;
  }
  double error(const VectorValues &x) const {
    const KeyVector order = keys();
    std::vector<double> d(6 * order.size(), 0.0);
    for (size_t k = 0; k < order.size(); ++k) {
      auto it = x.v.find(order[k]);
      if (it != x.v.end())
        for (int a = 0; a < 6; ++a) d[6 * k + a] = it->second[a];
    }
    return dense(order).error(d.data());
  }
  VectorValues optimizeDensely() const {
    const KeyVector order = keys();
    const form::dense::Quadratic sys = dense(order);
    std::vector<double> A = sys.G, x = sys.g;
    if (!form::dense::cholesky(A, sys.n)) throw IndeterminantLinearSystemException();
    form::dense::cholesky_solve(A, sys.n, x.data());
    VectorValues out;
    for (size_t k = 0; k < order.size(); ++k)
      for (int a = 0; a < 6; ++a) out.v[order[k]][a] = x[6 * k + a];
    return out;
  }
  /// Schur complement of `marg` out of the graph; second = the marginal on the remaining keys
  std::pair<std::shared_ptr<BayesTreeStub>, std::shared_ptr<GaussianFactorGraph>>
  eliminatePartialMultifrontal(const KeyVector &marg) const {
    auto remaining = std::make_shared<GaussianFactorGraph>();
    const KeyVector all = keys();
    KeyVector mk, rk;
    for (Key k : all) (std::find(marg.begin(), marg.end(), k) != marg.end() ? mk : rk).push_back(k);
    if (!rk.empty()) {
      KeyVector order = mk;
      order.insert(order.end(), rk.begin(), rk.end());
      const form::dense::Quadratic sys = dense(order);
      const size_t na = 6 * mk.size(), nc = 6 * rk.size(), n = na + nc;
      form::dense::Quadratic q;
      q.resize(nc);
      if (na == 0) {
        q = sys;
      } else {
        std::vector<double> A(na * na);
        for (size_t r = 0; r < na; ++r)
          for (size_t c = 0; c < na; ++c) A[r * na + c] = sys.G[r * n + c];
        bool ok = form::dense::cholesky(A, na);
        for (double jitter = 1e-9; !ok && jitter < 1.0; jitter *= 100.0) { // rank-deficient block: regularise
          for (size_t r = 0; r < na; ++r)
            for (size_t c = 0; c < na; ++c) A[r * na + c] = sys.G[r * n + c] + (r == c ? jitter : 0.0);
          ok = form::dense::cholesky(A, na);
        }
        std::vector<double> X((nc + 1) * na);
        for (size_t c = 0; c < nc; ++c) {
          double *col = &X[c * na];
          for (size_t r = 0; r < na; ++r) col[r] = sys.G[r * n + (na + c)];
          if (ok) form::dense::cholesky_solve(A, na, col);
        }
        double *xa = &X[nc * na];
        for (size_t r = 0; r < na; ++r) xa[r] = sys.g[r];
        if (ok) form::dense::cholesky_solve(A, na, xa);
        for (size_t r = 0; r < nc; ++r) {
          for (size_t c = 0; c < nc; ++c) {
            double s = sys.G[(na + r) * n + (na + c)];
            for (size_t k = 0; k < na; ++k) s -= sys.G[k * n + (na + r)] * X[c * na + k];
            q.G[r * nc + c] = s;
          }
          double s = sys.g[na + r];
          for (size_t k = 0; k < na; ++k) s -= sys.G[k * n + (na + r)] * xa[k];
          q.g[r] = s;
        }
        double f = sys.f;
        for (size_t k = 0; k < na; ++k) f -= sys.g[k] * xa[k];
        q.f = f;
        for (size_t r = 0; r < nc; ++r)
          for (size_t c = r + 1; c < nc; ++c) {
            const double v = 0.5 * (q.G[r * nc + c] + q.G[c * nc + r]);
            q.G[r * nc + c] = q.G[c * nc + r] = v;
          }
      }
      remaining->push_back(std::make_shared<HessianFactor>(shim::to_factor(rk, q)));
    }
    return {std::make_shared<BayesTreeStub>(), remaining};
  }

private:
  static const HessianFactor &hessian(const std::shared_ptr<GaussianFactor> &f) {
    const HessianFactor *h = dynamic_cast<const HessianFactor *>(f.get());
    if (!h) throw std::runtime_error("shim: only Hessian factors are produced by the stand-ins");
    return *h;
  }
};

inline HessianFactor::HessianFactor(const GaussianFactorGraph &graph) {
  const KeyVector order = graph.keys();
  *this = shim::to_factor(order, graph.dense(order));
}

/// gtsam::LinearContainerFactor over a HessianFactor
class LinearContainerFactor : public NonlinearFactor {
public:
  LinearContainerFactor(const HessianFactor &factor, const Values &linearizationPoint) : m_factor(factor) {
    for (Key k : factor.keys) m_lin.push_back(linearizationPoint.at<Pose3>(k));
  }
  const KeyVector &keys() const override { return m_factor.keys; }
  double error(const Values &c) const override {
    const std::vector<double> d = delta(c);
    return quadratic().error(d.data());
  }
  /// the stored quadratic re-centred on c: G' = G, g' = g - G d, f' = f - 2 g.d + d.G.d
  std::shared_ptr<GaussianFactor> linearize(const Values &c) const override {
    const std::vector<double> d = delta(c);
    const form::dense::Quadratic q = quadratic();
    form::dense::Quadratic o = q;
    for (size_t r = 0; r < q.n; ++r) {
      double s = 0.0;
      for (size_t col = 0; col < q.n; ++col) s += q.G[r * q.n + col] * d[col];
      o.g[r] -= s;
    }
    o.f = 2.0 * q.error(d.data());
    return std::make_shared<HessianFactor>(shim::to_factor(m_factor.keys, o));
  }
  /// every factor of a linear graph wrapped with the given linearisation points
  static class NonlinearFactorGraph ConvertLinearGraph(const GaussianFactorGraph &graph, const Values &lin);

private:
  std::vector<double> delta(const Values &c) const {
    std::vector<double> d(6 * m_lin.size());
    for (size_t k = 0; k < m_lin.size(); ++k) {
      const Vector6 l = m_lin[k].localCoordinates(c.at<Pose3>(m_factor.keys[k]));
      for (int a = 0; a < 6; ++a) d[6 * k + a] = l(a);
    }
    return d;
  }
  form::dense::Quadratic quadratic() const {
    form::dense::Quadratic q;
    const size_t n = 6 * m_factor.keys.size();
    q.resize(n);
    for (size_t r = 0; r < n; ++r) {
      for (size_t c = 0; c < n; ++c) q.G[r * n + c] = m_factor.info(r, c);
      q.g[r] = m_factor.info(r, n);
    }
    q.f = m_factor.info(n, n);
    return q;
  }
  HessianFactor m_factor;
  std::vector<Pose3> m_lin;
};

/// gtsam::PriorFactor<Pose3>
class PosePriorFactor : public NonlinearFactor {
public:
  PosePriorFactor(Key key, const Pose3 &prior, const SharedNoiseModel &model) : m_keys{key}, m_prior(prior), m_model(model) {}
  const KeyVector &keys() const override { return m_keys; }
  double error(const Values &x) const override {
    const Vector6 e = m_prior.localCoordinates(x.at<Pose3>(m_keys[0]));
    Vector v(6);
    for (int a = 0; a < 6; ++a) v(a) = e(a);
    return shim::half_whitened_norm(m_model, v);
  }
  std::shared_ptr<GaussianFactor> linearize(const Values &x) const override {
    const formhostmath::Pose3 between = m_prior.host().inverse() * x.at<Pose3>(m_keys[0]).host();
    const formhostmath::Vec6 e = formhostmath::Pose3::Logmap(between);
    const formhostmath::Mat6 J = formhostmath::Pose3::LogmapDerivative(between);
    std::vector<Matrix> A(1, Matrix(6, 6));
    Vector b(6);
    for (int r = 0; r < 6; ++r) {
      b(r) = -e[r];
      for (int c = 0; c < 6; ++c) A[0](r, c) = J[6 * r + c];
    }
    static_cast<const noiseModel::Gaussian *>(m_model.get())->WhitenSystem(A, b);
    std::vector<std::pair<Key, Matrix>> terms(1);
    terms[0].first = m_keys[0];
    terms[0].second.swap(A[0]);
    return std::make_shared<HessianFactor>(JacobianFactor(terms, b));
  }

private:
  KeyVector m_keys;
  Pose3 m_prior;
  SharedNoiseModel m_model;
};

class NonlinearFactorGraph {
public:
  typedef std::vector<NonlinearFactor::shared_ptr>::iterator iterator;
  typedef std::vector<NonlinearFactor::shared_ptr>::const_iterator const_iterator;
  template <typename F, typename = std::enable_if_t<std::is_base_of<NonlinearFactor, F>::value>>
  void push_back(const F &factor) {
    m_factors.push_back(std::make_shared<F>(factor));
  }
  void push_back(const NonlinearFactor::shared_ptr &factor) { m_factors.push_back(factor); }
  void replace(size_t index, const NonlinearFactor::shared_ptr &factor) { m_factors.at(index) = factor; }
  void addPrior(Key key, const Pose3 &prior, const SharedNoiseModel &model) {
    m_factors.push_back(std::make_shared<PosePriorFactor>(key, prior, model));
  }
  size_t size() const { return m_factors.size(); }
  iterator begin() { return m_factors.begin(); }
  iterator end() { return m_factors.end(); }
  const_iterator begin() const { return m_factors.begin(); }
  const_iterator end() const { return m_factors.end(); }
  std::shared_ptr<GaussianFactorGraph> linearize(const Values &x) const {
    auto g = std::make_shared<GaussianFactorGraph>();
    for (const auto &f : m_factors)
      if (f) g->push_back(f->linearize(x));
    return g;
  }
  double error(const Values &x) const {
    double e = 0.0;
    for (const auto &f : m_factors)
      if (f) e += f->error(x);
    return e;
  }

private:
  std::vector<NonlinearFactor::shared_ptr> m_factors;
};

inline NonlinearFactorGraph LinearContainerFactor::ConvertLinearGraph(const GaussianFactorGraph &graph, const Values &lin) {
  NonlinearFactorGraph out;
  for (const auto &f : graph) {
    const HessianFactor *h = dynamic_cast<const HessianFactor *>(f.get());
    if (h) out.push_back(std::make_shared<LinearContainerFactor>(*h, lin));
  }
  return out;
}

class NonlinearOptimizerParams {
public:
  virtual ~NonlinearOptimizerParams() = default;
  size_t maxIterations = 100;
  double relativeErrorTol = 1e-5;
  double absoluteErrorTol = 1e-5;
  double errorTol = 0.0;
  Ordering::OrderingType orderingType = Ordering::COLAMD;
};

class LevenbergMarquardtParams : public NonlinearOptimizerParams {
public:
  double lambdaInitial = 1e-5;
  double lambdaFactor = 10.0;
  double lambdaUpperBound = 1e5;
  double lambdaLowerBound = 0.0;
  double minModelFidelity = 1e-3;
  bool diagonalDamping = false;
  bool useFixedLambdaFactor = true;
};

class LevenbergMarquardtOptimizer {
public:
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph &graph, const Values &initial,
                              const LevenbergMarquardtParams &params)
      : m_graph(graph), m_values(initial), m_params(params), m_lambda(params.lambdaInitial) {}
  virtual ~LevenbergMarquardtOptimizer() = default;
  virtual VectorValues solve(const GaussianFactorGraph &gfg, const NonlinearOptimizerParams &) const {
    return gfg.optimizeDensely();
  }
  size_t iterations() const { return m_iterations; }
  size_t linearizations() const { return m_linearizations; }
  size_t error_evaluations() const { return m_error_evaluations; }

  /// NonlinearOptimizer::defaultOptimize
  const Values &optimize() {
    m_error = graph_error(m_values);
    double currentError = m_error;
    if (currentError <= m_params.errorTol || m_iterations >= m_params.maxIterations) return m_values;
    double newError = currentError;
    do {
      currentError = newError;
      iterate();
      newError = m_error;
    } while (m_iterations < m_params.maxIterations && !converged(currentError, newError) &&
             std::isfinite(currentError));
    return m_values;
  }

private:
  double graph_error(const Values &v) {
    ++m_error_evaluations;
    return m_graph.error(v);
  }
  bool converged(double currentError, double newError) const {
    if (newError <= m_params.errorTol) return true;
    const double absoluteDecrease = currentError - newError;
    const double relativeDecrease = absoluteDecrease / currentError;
    return (m_params.relativeErrorTol != 0.0 && relativeDecrease <= m_params.relativeErrorTol) ||
           absoluteDecrease <= m_params.absoluteErrorTol;
  }
  void iterate() {
    const std::shared_ptr<GaussianFactorGraph> linear = m_graph.linearize(m_values);
    ++m_linearizations;
    while (!try_lambda(*linear)) {
    }
  }
  bool try_lambda(const GaussianFactorGraph &linear) {
    // buildDampedSystem: one prior of information lambda * I per variable (diagonalDamping = false)
    GaussianFactorGraph damped = linear;
    for (Key k : linear.keys()) {
      form::dense::Quadratic q;
      q.resize(6);
      for (int a = 0; a < 6; ++a) q.G[a * 6 + a] = m_lambda;
      damped.push_back(std::make_shared<HessianFactor>(shim::to_factor(KeyVector{k}, q)));
    }
    double modelFidelity = 0.0, newError = std::numeric_limits<double>::infinity(), costChange = 0.0;
    bool step_is_successful = false, stopSearchingLambda = false, solved = true;
    Values newValues;
    VectorValues delta;
    try {
      delta = solve(damped, m_params);
    } catch (const IndeterminantLinearSystemException &) {
      solved = false;
    }
    if (solved) {
      const double oldLinearizedError = linear.error(VectorValues());
      const double newlinearizedError = linear.error(delta);
      const double linearizedCostChange = oldLinearizedError - newlinearizedError;
      if (linearizedCostChange >= 0) {
        newValues = m_values.retract(delta);
        newError = graph_error(newValues);
        costChange = m_error - newError;
        if (linearizedCostChange > std::numeric_limits<double>::epsilon() * oldLinearizedError) {
          modelFidelity = costChange / linearizedCostChange;
          step_is_successful = modelFidelity > m_params.minModelFidelity;
        }
        const double minAbsoluteTolerance = m_params.relativeErrorTol * m_error;
        if (std::abs(costChange) < minAbsoluteTolerance) stopSearchingLambda = true;
      }
    }
    if (step_is_successful) {
      m_values = newValues;
      m_error = newError;
      m_lambda = std::max(m_params.lambdaLowerBound, m_lambda / m_params.lambdaFactor);
      ++m_iterations;
      return true;
    } else if (!stopSearchingLambda) {
      m_lambda *= m_params.lambdaFactor;
      return m_lambda >= m_params.lambdaUpperBound; // give up at the maximum lambda
    }
    return true;
  }

  NonlinearFactorGraph m_graph;
  Values m_values;
  LevenbergMarquardtParams m_params;
  double m_lambda;
  double m_error = 0.0;
  size_t m_iterations = 0, m_linearizations = 0, m_error_evaluations = 0;
};

} // namespace gtsam
