// Stand-in for Ceres Solver (TEST INFRASTRUCTURE, oracle/_ref build only).  THIRD-PARTY SEMANTICS RESTATED, for exactly the
// configuration the reference uses (laserOdometry.cpp:331-334,413-418,487-492; laserMapping.cpp:363-367,468-475): one parameter
// block, residual blocks of dimension 1 sharing one LossFunction, TRUST_REGION / LEVENBERG_MARQUARDT, DENSE_QR, default options of
// Ceres 1.13/1.14 (trust_region_minimizer.cc, levenberg_marquardt_strategy.cc, corrector.cc, loss_function.cc — recalled from
// the published source, which is not in this container).  The cost functions whose Evaluate() this solver
// calls are the REFERENCE'S OWN classes (include/alego/utility.h:122-349), compiled unmodified.
#ifndef ALEGO_REF_SHIM_CERES_H
#define ALEGO_REF_SHIM_CERES_H
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <string>
#include <vector>

namespace ceres {

class CostFunction {
 public:
  virtual ~CostFunction() {}
  virtual bool Evaluate(double const *const *parameters, double *residuals, double **jacobians) const = 0;
  int num_residuals() const { return num_residuals_; }
  int parameter_size() const { return parameter_size_; }

 protected:
  int num_residuals_ = 0, parameter_size_ = 0;
};

template <int kNumResiduals, int N0>
class SizedCostFunction : public CostFunction {
 public:
  SizedCostFunction() { num_residuals_ = kNumResiduals; parameter_size_ = N0; }
};

class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

class HuberLoss : public LossFunction {  // loss_function.cc
 public:
  explicit HuberLoss(double a) : a_(a), b_(a * a) {}
  void Evaluate(double s, double rho[3]) const override {
    if (s > b_) {
      const double r = std::sqrt(s);
      rho[0] = 2.0 * a_ * r - b_;
      rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r);
      rho[2] = -rho[1] / (2.0 * s);
    } else {
      rho[0] = s; rho[1] = 1.0; rho[2] = 0.0;
    }
  }

 private:
  const double a_, b_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

class Problem {
 public:
  struct Options {};
  Problem() {}
  explicit Problem(const Options &) {}
  ~Problem() {  // Problem owns cost and loss functions by default (each loss deleted once)
    for (CostFunction *c : cost_) delete c;
    std::sort(loss_.begin(), loss_.end());
    loss_.erase(std::unique(loss_.begin(), loss_.end()), loss_.end());
    for (LossFunction *l : loss_) delete l;
  }
  void AddParameterBlock(double *values, int size) { x_ = values; n_ = size; }
  void AddResidualBlock(CostFunction *c, LossFunction *l, double *x) {
    x_ = x;
    n_ = c->parameter_size();
    cost_.push_back(c);
    loss_.push_back(l);
    block_loss_.push_back(l);
  }
  int NumResidualBlocks() const { return static_cast<int>(cost_.size()); }

  std::vector<CostFunction *> cost_;
  std::vector<LossFunction *> loss_, block_loss_;
  double *x_ = nullptr;
  int n_ = 0;
};

struct Solver {
  struct Options {
    LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY;
    int max_num_iterations = 50;
    bool minimizer_progress_to_stdout = false;
    bool check_gradients = false;
    double gradient_check_relative_precision = 1e-8;
    double initial_trust_region_radius = 1e4, max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    int max_num_consecutive_invalid_steps = 5;
    bool jacobi_scaling = true;
  };
  struct Summary {
    TerminationType termination_type = NO_CONVERGENCE;
    double initial_cost = 0, final_cost = 0;
    int num_successful_steps = 0, num_unsuccessful_steps = 0;
    int num_iterations = 0;  // step attempts (Ceres' iterations.size() - 1)
    std::string message;
    // one entry per iteration record (0 = the initial evaluation): cost and the parameter vector after it
    std::vector<double> trace_cost;
    std::vector<std::vector<double>> trace_x;
    std::string BriefReport() const {
      char buf[256];
      std::snprintf(buf, sizeof buf, "Ceres(shim) Solver Report: Iterations: %d, Initial cost: %e, Final cost: %e, Termination: %s",
                    num_iterations, initial_cost, final_cost, termination_type == CONVERGENCE ? "CONVERGENCE" : "NO_CONVERGENCE");
      return buf;
    }
  };
};

// every Solve() appends its Summary here (per thread) so that a driver can read iteration traces without touching the caller's code
inline std::vector<Solver::Summary> &solve_log() { static thread_local std::vector<Solver::Summary> log; return log; }

namespace shim_detail {

struct Eval {
  double cost = 0;
  std::vector<double> r, J;  // corrected residuals; corrected Jacobian, row-major m x n
  std::vector<double> g;     // J^T r
};

// ResidualBlock::Evaluate + Corrector (rho'' <= 0 for Huber: residual and Jacobian scaled by sqrt(rho'))
inline bool evaluate(const Problem &p, const double *x, Eval *out, double *cost_out = nullptr) {
  const int m = p.NumResidualBlocks(), n = p.n_;
  double cost = 0;
  if (out) { out->r.assign(m, 0.0); out->J.assign(static_cast<std::size_t>(m) * n, 0.0); out->g.assign(n, 0.0); }
  std::vector<double> jrow(n);
  for (int k = 0; k < m; ++k) {
    double r = 0;
    double *jac[1] = {jrow.data()};
    const double *params[1] = {x};
    if (!p.cost_[k]->Evaluate(params, &r, out ? jac : nullptr)) return false;
    const double s = r * r;
    double rho[3] = {s, 1.0, 0.0};
    if (p.block_loss_[k]) p.block_loss_[k]->Evaluate(s, rho);
    cost += 0.5 * rho[0];
    if (out) {
      double w = 1.0;
      if (p.block_loss_[k]) {
        // Corrector::Corrector: rho[2] <= 0 -> residual_scaling = sqrt(rho[1]), alpha_sq_norm = 0
        w = std::sqrt(rho[1]);
      }
      out->r[k] = w * r;
      for (int c = 0; c < n; ++c) {
        out->J[static_cast<std::size_t>(k) * n + c] = w * jrow[c];
        out->g[c] += out->J[static_cast<std::size_t>(k) * n + c] * out->r[k];
      }
    }
  }
  if (out) out->cost = cost;
  if (cost_out) *cost_out = cost;
  return true;
}

// DenseQRSolver: min || [J; diag(D)] s - [r; 0] || by Householder QR of the stacked matrix
inline bool dense_qr(const std::vector<double> &J, const std::vector<double> &r, const std::vector<double> &D, int m, int n,
                     std::vector<double> *s) {
  const int M = m + n;
  std::vector<double> A(static_cast<std::size_t>(M) * n, 0.0), b(M, 0.0);
  for (int i = 0; i < m; ++i) { for (int c = 0; c < n; ++c) A[static_cast<std::size_t>(i) * n + c] = J[static_cast<std::size_t>(i) * n + c]; b[i] = r[i]; }
  for (int c = 0; c < n; ++c) A[static_cast<std::size_t>(m + c) * n + c] = D[c];
  std::vector<double> v(M);
  for (int c = 0; c < n; ++c) {
    double nrm = 0;
    for (int i = c; i < M; ++i) nrm += A[static_cast<std::size_t>(i) * n + c] * A[static_cast<std::size_t>(i) * n + c];
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return false;
    const double alpha = A[static_cast<std::size_t>(c) * n + c] > 0 ? -nrm : nrm;
    for (int i = c; i < M; ++i) v[i] = A[static_cast<std::size_t>(i) * n + c];
    v[c] -= alpha;
    double vn = 0;
    for (int i = c; i < M; ++i) vn += v[i] * v[i];
    if (vn == 0.0) continue;
    for (int c2 = c; c2 < n; ++c2) {
      double d = 0;
      for (int i = c; i < M; ++i) d += v[i] * A[static_cast<std::size_t>(i) * n + c2];
      const double f = 2.0 * d / vn;
      for (int i = c; i < M; ++i) A[static_cast<std::size_t>(i) * n + c2] -= f * v[i];
    }
    double d = 0;
    for (int i = c; i < M; ++i) d += v[i] * b[i];
    const double f = 2.0 * d / vn;
    for (int i = c; i < M; ++i) b[i] -= f * v[i];
  }
  s->assign(n, 0.0);
  for (int c = n - 1; c >= 0; --c) {
    double acc = b[c];
    for (int c2 = c + 1; c2 < n; ++c2) acc -= A[static_cast<std::size_t>(c) * n + c2] * (*s)[c2];
    if (A[static_cast<std::size_t>(c) * n + c] == 0.0) return false;
    (*s)[c] = acc / A[static_cast<std::size_t>(c) * n + c];
  }
  for (int c = 0; c < n; ++c) if (!std::isfinite((*s)[c])) return false;
  return true;
}

}  // namespace shim_detail

inline void Solve(const Solver::Options &opt, Problem *problem, Solver::Summary *sum) {
  using namespace shim_detail;
  const int m = problem->NumResidualBlocks(), n = problem->n_;
  double *x = problem->x_;
  *sum = Solver::Summary();
  if (m == 0 || n == 0) { sum->termination_type = CONVERGENCE; return; }
  auto col_sq = [&](const Eval &e, int c) { double q = 0; for (int i = 0; i < m; ++i) q += e.J[static_cast<std::size_t>(i) * n + c] * e.J[static_cast<std::size_t>(i) * n + c]; return q; };
  auto norm = [&](const double *v) { double q = 0; for (int c = 0; c < n; ++c) q += v[c] * v[c]; return std::sqrt(q); };
  auto record = [&](double cost) { sum->trace_cost.push_back(cost); sum->trace_x.push_back(std::vector<double>(x, x + n)); };
  Eval E;
  if (!evaluate(*problem, x, &E)) { sum->termination_type = FAILURE; return; }
  double cost = E.cost;
  sum->initial_cost = cost;
  std::vector<double> scale(n, 1.0);
  if (opt.jacobi_scaling) for (int c = 0; c < n; ++c) scale[c] = 1.0 / (1.0 + std::sqrt(col_sq(E, c)));
  auto apply_scale = [&](Eval &e) { for (int i = 0; i < m; ++i) for (int c = 0; c < n; ++c) e.J[static_cast<std::size_t>(i) * n + c] *= scale[c]; };
  apply_scale(E);
  double radius = opt.initial_trust_region_radius, decrease_factor = 2.0;
  bool reuse_diagonal = false;
  std::vector<double> diag(n), lm(n), step(n), xc(n);
  double x_norm = norm(x);
  int iter = 0, invalid_run = 0;
  record(cost);
  {  // gradient check at the starting point
    double gmax = 0;
    for (int c = 0; c < n; ++c) gmax = std::max(gmax, std::fabs(E.g[c]));
    if (gmax <= opt.gradient_tolerance) { sum->termination_type = CONVERGENCE; sum->final_cost = cost; return; }
  }
  while (true) {
    if (iter >= opt.max_num_iterations) { sum->termination_type = NO_CONVERGENCE; break; }
    if (radius <= opt.min_trust_region_radius) { sum->termination_type = CONVERGENCE; break; }
    ++iter;
    if (!reuse_diagonal)
      for (int c = 0; c < n; ++c) diag[c] = std::min(std::max(col_sq(E, c), opt.min_lm_diagonal), opt.max_lm_diagonal);
    for (int c = 0; c < n; ++c) lm[c] = std::sqrt(diag[c] / radius);
    const bool ok = dense_qr(E.J, E.r, lm, m, n, &step);
    reuse_diagonal = true;
    double model_change = 0;
    if (ok) {
      for (int c = 0; c < n; ++c) step[c] = -step[c];
      for (int i = 0; i < m; ++i) {
        double mr = 0;
        for (int c = 0; c < n; ++c) mr += E.J[static_cast<std::size_t>(i) * n + c] * step[c];
        model_change -= mr * (E.r[i] + mr / 2.0);
      }
    }
    if (!ok || !(model_change > 0.0)) {
      ++sum->num_unsuccessful_steps;
      if (++invalid_run >= opt.max_num_consecutive_invalid_steps) { sum->termination_type = FAILURE; break; }
      radius *= 0.5;  // LevenbergMarquardtStrategy::StepIsInvalid
      record(cost);
      continue;
    }
    invalid_run = 0;
    double step_sq = 0;
    for (int c = 0; c < n; ++c) { xc[c] = x[c] + step[c] * scale[c]; step_sq += (x[c] - xc[c]) * (x[c] - xc[c]); }
    const double step_norm = std::sqrt(step_sq);
    double cand_cost = 0;
    evaluate(*problem, xc.data(), nullptr, &cand_cost);
    if (step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance)) { sum->termination_type = CONVERGENCE; break; }
    const double cost_change = cost - cand_cost;
    if (std::fabs(cost_change) <= opt.function_tolerance * cost) { sum->termination_type = CONVERGENCE; break; }
    const double rho = cost_change / model_change;
    if (rho > opt.min_relative_decrease) {
      for (int c = 0; c < n; ++c) x[c] = xc[c];
      x_norm = norm(x);
      evaluate(*problem, x, &E);
      cost = E.cost;
      apply_scale(E);
      radius = std::min(opt.max_trust_region_radius, radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * rho - 1.0, 3)));
      decrease_factor = 2.0;
      reuse_diagonal = false;
      ++sum->num_successful_steps;
      record(cost);
      double gmax = 0;
      for (int c = 0; c < n; ++c) gmax = std::max(gmax, std::fabs(E.g[c]));
      if (gmax <= opt.gradient_tolerance) { sum->termination_type = CONVERGENCE; break; }
    } else {
      ++sum->num_unsuccessful_steps;
      radius = radius / decrease_factor;
      decrease_factor *= 2.0;
      reuse_diagonal = true;
      record(cost);
    }
  }
  sum->num_iterations = iter;
  sum->final_cost = cost;
  solve_log().push_back(*sum);
}

}  // namespace ceres
#endif
