// TEST INFRASTRUCTURE - not GTSAM.  The slice of GTSAM's factor machinery that
// form/optimization/gtsam.hpp and form/feature/factor.{hpp,cpp} touch, with GTSAM's
// published semantics [external, SURVEY App. B]:
//   * noiseModel::Gaussian::WhitenSystem(A, b) whitens every Jacobian block and the rhs
//     through the model's virtual WhitenInPlace / whitenInPlace;
//   * NoiseModelFactor2::unwhitenedError(x, H) = evaluateError(x[key1], x[key2], &H[0], &H[1]);
//   * JacobianFactor(terms, b) keeps the blocks; HessianFactor(JacobianFactor) is the
//     augmented information matrix [A b]^T [A b], variables in term order;
//   * NonlinearFactor as the common base (keys, error = 0.5 |whitened residual|^2, linearize),
//     noiseModel::Isotropic, NoiseModelFactor1's default linearisation - what
//     form/optimization/constraints.cpp needs on top (the smoother stand-ins are in
//     gtsam/shim_smoother.h).
#pragma once

#include <gtsam/base/Vector.h>
#include <gtsam/nonlinear/Values.h>

#include <cmath>
#include <iostream>
#include <memory>
#include <string>
#include <typeinfo>
#include <utility>
#include <vector>

namespace gtsam {

template <int R, int C> class OptionalJacobian { // OptionalJacobian<Dynamic, Dynamic>
public:
  OptionalJacobian() = default;
  OptionalJacobian(std::nullptr_t) {}
  OptionalJacobian(Matrix *m) : m_m(m) {}
  explicit operator bool() const { return m_m != nullptr; }
  Matrix *operator->() const { return m_m; }
  Matrix &operator*() const { return *m_m; }

private:
  Matrix *m_m = nullptr;
};

namespace noiseModel {
class Base {
public:
  typedef std::shared_ptr<Base> shared_ptr;
  explicit Base(size_t dim) : m_dim(dim) {}
  virtual ~Base() = default;
  size_t dim() const { return m_dim; }
  virtual void print(const std::string &s = "") const = 0;
  virtual bool equals(const Base &expected, double tol = 1e-9) const = 0;

private:
  size_t m_dim;
};
class Gaussian : public Base {
public:
  explicit Gaussian(size_t dim) : Base(dim) {}
  virtual Vector whiten(const Vector &v) const = 0;
  virtual Vector unwhiten(const Vector &v) const = 0;
  virtual Matrix Whiten(const Matrix &H) const = 0;
  virtual void WhitenInPlace(Matrix &H) const = 0;
  virtual void whitenInPlace(Vector &v) const = 0;
  virtual void WhitenInPlace(Eigen::Block<Matrix> H) const = 0;
  void WhitenSystem(std::vector<Matrix> &A, Vector &b) const {
    for (Matrix &Aj : A) WhitenInPlace(Aj);
    whitenInPlace(b);
  }
};
/// noiseModel::Isotropic::Sigma(dim, sigma): whitening = division by sigma
class Isotropic : public Gaussian {
public:
  typedef std::shared_ptr<Isotropic> shared_ptr;
  static shared_ptr Sigma(size_t dim, double sigma) { return shared_ptr(new Isotropic(dim, sigma)); }
  void print(const std::string &s = "") const override { std::cout << s << "Isotropic(" << m_sigma << ")\n"; }
  bool equals(const Base &expected, double tol = 1e-9) const override {
    const Isotropic *p = dynamic_cast<const Isotropic *>(&expected);
    return p && std::fabs(p->m_sigma - m_sigma) <= tol && p->dim() == dim();
  }
  Vector whiten(const Vector &v) const override { return v * m_inv; }
  Vector unwhiten(const Vector &v) const override { return v * m_sigma; }
  Matrix Whiten(const Matrix &H) const override { return H * m_inv; }
  void WhitenInPlace(Matrix &H) const override { H *= m_inv; }
  void whitenInPlace(Vector &v) const override { v *= m_inv; }
  void WhitenInPlace(Eigen::Block<Matrix> H) const override { H *= m_inv; }

private:
  Isotropic(size_t dim, double sigma) : Gaussian(dim), m_sigma(sigma), m_inv(1.0 / sigma) {}
  double m_sigma, m_inv;
};
} // namespace noiseModel
using SharedNoiseModel = noiseModel::Base::shared_ptr;
using KeyVector = std::vector<Key>;

class GaussianFactor {
public:
  virtual ~GaussianFactor() = default;
};
class GaussianFactorGraph;
class JacobianFactor : public GaussianFactor {
public:
  JacobianFactor(const std::vector<std::pair<Key, Matrix>> &terms, const Vector &b) : terms(terms), b(b) {}
  std::vector<std::pair<Key, Matrix>> terms;
  Vector b;
};
class HessianFactor : public GaussianFactor {
public:
  HessianFactor() = default;
  /// all factors of a linear graph combined into one dense factor (shim_smoother.h)
  explicit HessianFactor(const GaussianFactorGraph &graph);
  explicit HessianFactor(const JacobianFactor &jf) {
    size_t n = 0;
    for (const auto &t : jf.terms) {
      keys.push_back(t.first);
      n += t.second.cols();
    }
    const size_t m = jf.b.size();
    Matrix Ab(m, n + 1);
    size_t c0 = 0;
    for (const auto &t : jf.terms) {
      for (size_t c = 0; c < t.second.cols(); ++c)
        for (size_t r = 0; r < m; ++r) Ab(r, c0 + c) = t.second(r, c);
      c0 += t.second.cols();
    }
    for (size_t r = 0; r < m; ++r) Ab(r, n) = jf.b(r);
    info = Matrix(Ab.transpose() * Ab);
  }
  std::vector<Key> keys;
  Matrix info; // (n + 1) x (n + 1) augmented information matrix
};

/// gtsam::NonlinearFactor: what a NonlinearFactorGraph holds
class NonlinearFactor {
public:
  typedef std::shared_ptr<NonlinearFactor> shared_ptr;
  virtual ~NonlinearFactor() = default;
  virtual const KeyVector &keys() const = 0;
  virtual double error(const Values &x) const = 0;
  virtual std::shared_ptr<GaussianFactor> linearize(const Values &x) const = 0;
};

namespace shim {
/// 0.5 * |whiten(e)|^2 (NoiseModelFactor::error)
inline double half_whitened_norm(const SharedNoiseModel &model, Vector e) {
  static_cast<const noiseModel::Gaussian *>(model.get())->whitenInPlace(e);
  double s = 0.0;
  for (size_t r = 0; r < e.size(); ++r) s += e(r) * e(r);
  return 0.5 * s;
}
} // namespace shim

template <typename V1, typename V2> class NoiseModelFactor2 : public NonlinearFactor {
public:
  NoiseModelFactor2(const SharedNoiseModel &noiseModel, Key i, Key j) : noiseModel_(noiseModel), keys_{i, j} {}
  virtual Vector evaluateError(const V1 &, const V2 &, Matrix *H1 = nullptr, Matrix *H2 = nullptr) const = 0;
  size_t size() const { return 2; }
  const KeyVector &keys() const override { return keys_; }
  double error(const Values &x) const override {
    return shim::half_whitened_norm(noiseModel_, evaluateError(x.template at<V1>(keys_[0]), x.template at<V2>(keys_[1])));
  }
  const SharedNoiseModel &noiseModel() const { return noiseModel_; }
  Vector unwhitenedError(const Values &x, std::vector<Matrix> &H) const {
    return evaluateError(x.at<V1>(keys_[0]), x.at<V2>(keys_[1]), &H[0], &H[1]);
  }

protected:
  // the reference calls noiseModel_->WhitenSystem: GTSAM keeps the Gaussian interface here
  struct ModelPtr {
    SharedNoiseModel p;
    ModelPtr(const SharedNoiseModel &q) : p(q) {}
    explicit operator bool() const { return (bool)p; }
    const noiseModel::Gaussian *operator->() const { return static_cast<const noiseModel::Gaussian *>(p.get()); }
    operator const SharedNoiseModel &() const { return p; }
  } noiseModel_;
  std::vector<Key> keys_;
};

template <typename V1> class NoiseModelFactor1 : public NonlinearFactor {
public:
  NoiseModelFactor1(const SharedNoiseModel &noiseModel, Key i) : noiseModel_(noiseModel), keys_{i} {}
  virtual Vector evaluateError(const V1 &, Matrix *H = nullptr) const = 0;
  const KeyVector &keys() const override { return keys_; }
  double error(const Values &x) const override {
    return shim::half_whitened_norm(noiseModel_, evaluateError(x.template at<V1>(keys_[0])));
  }
  /// NoiseModelFactor::linearize: whitened Jacobian and rhs b = -error
  std::shared_ptr<GaussianFactor> linearize(const Values &x) const override {
    std::vector<Matrix> A(1);
    Vector b = -evaluateError(x.template at<V1>(keys_[0]), &A[0]);
    static_cast<const noiseModel::Gaussian *>(noiseModel_.get())->WhitenSystem(A, b);
    std::vector<std::pair<Key, Matrix>> terms(1);
    terms[0].first = keys_[0];
    terms[0].second.swap(A[0]);
    return std::make_shared<HessianFactor>(JacobianFactor(terms, b));
  }

protected:
  SharedNoiseModel noiseModel_;
  KeyVector keys_;
};

} // namespace gtsam
