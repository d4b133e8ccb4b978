// qp_solver.hpp -- drop-in replacement for the reference's include/smooth/feedback/qp_solver.hpp on top of the sfb C ABI.
//
// It replaces that ONE header and nothing else: problem and solution types stay the reference's own
//   #include <smooth/feedback/qp.hpp>   QuadraticProgram<M,N,S>, QuadraticProgramSparse<S>, QPSolutionStatus, QPSolution<M,N,S>
// (asif_func.hpp and ocp_to_qp.hpp keep including qp.hpp for the problem types, so nothing may be redefined here), and this
// file provides exactly what qp_solver.hpp provides to its includers (mpc.hpp:10, asif.hpp:9):
//   QPSolverParams                       qp_solver.hpp:29-68
//   detail::qp_solution_t<Pbm>           qp_solver.hpp:72-76
//   QPSolver<Pbm>::{QPSolver, analyze, solve, sol}   qp_solver.hpp:242-757, with the reference's signatures:
//        const QPSolution<M,N,Scalar> & solve(const Pbm &, std::optional<std::reference_wrapper<const QPSolution<M,N,Scalar>>> = {})
//   solve_qp(pbm, prm, warmstart) -> detail::qp_solution_t<Pbm>          qp_solver.hpp:779-787
// so that the call sites compile unchanged:  `const auto & sol = qp_solver_.solve(qp_, warmstart_); warmstart_ = sol;`
// (mpc.hpp:491,513 with std::optional<QPSolution<-1,-1,double>> warmstart_, :635) and
// `auto sol = feedback::solve_qp(qp_, prm_.qp, warmstart_); warmstart_ = sol;` (asif.hpp:97-99,109).
// tests/cpp/replay_mpc_asif.cpp replays those declarations and calls verbatim against this header.
// Plus the one extension the GPU engine exists for: QPSolver::solve_batch.
//
// `Pbm` is any type with public members P, q, A, l, u, exactly as in the reference (qp_solver.hpp:245-251 only looks at
// decltype(Pbm::A)): dense problems go to sfb_qp_solve_dense_batch_*, sparse ones (A derives from Eigen::SparseMatrixBase,
// the reference's own test, qp_solver.hpp:247) to sfb_qp_solve_sparse_batch_*.  For sparse problems the pattern is analysed
// on the first solve after construction or copy, like the reference's analyzePattern (qp_solver.hpp:424, LDLTWrapper
// :209-231); all problems of one solve_batch call must share that pattern (compressed storage required).
//
// All numerics run on the GPU through libsfb.so; there is no CPU fallback: if no device is available the constructor of
// the solver throws std::runtime_error with the library's message.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <functional>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include <smooth/feedback/qp.hpp>

#include "../sfb.h"

namespace smooth::feedback {

/// Options (qp_solver.hpp:29-68) -- float members stay float on purpose
struct QPSolverParams
{
  bool verbose = false;
  float alpha = 1.6f;
  float rho = 0.1f;
  float sigma = 1e-6f;
  bool scaling = true;
  float eps_abs = 1e-3f;
  float eps_rel = 1e-3f;
  float eps_primal_inf = 1e-4f;
  float eps_dual_inf = 1e-4f;
  std::optional<uint32_t> max_iter = {};
  std::optional<std::chrono::nanoseconds> max_time = {};
  uint32_t stop_check_iter = 25;
  bool polish = true;
  uint32_t polish_iter = 5;
  float delta = 1e-6f;
};

namespace detail {

/// qp_solver.hpp:72-76
template<typename Pbm>
using qp_solution_t = QPSolution<
  decltype(Pbm::A)::RowsAtCompileTime,
  decltype(Pbm::A)::ColsAtCompileTime,
  typename decltype(Pbm::A)::Scalar>;

inline sfb_qp_params to_c(const QPSolverParams & p)
{
  sfb_qp_params c;
  sfb_qp_params_default(&c);
  c.verbose = p.verbose;
  c.alpha = p.alpha; c.rho = p.rho; c.sigma = p.sigma;
  c.scaling = p.scaling;
  c.eps_abs = p.eps_abs; c.eps_rel = p.eps_rel;
  c.eps_primal_inf = p.eps_primal_inf; c.eps_dual_inf = p.eps_dual_inf;
  c.has_max_iter = p.max_iter.has_value(); c.max_iter = p.max_iter.value_or(0);
  c.has_max_time = p.max_time.has_value(); c.max_time_ns = p.max_time ? p.max_time->count() : 0;
  c.stop_check_iter = p.stop_check_iter;
  c.polish = p.polish; c.polish_iter = p.polish_iter; c.delta = p.delta;
  return c;
}

/// RAII owner of an sfb handle; copies create a fresh handle (a solver copy is an independent solver, cf. LDLTWrapper)
class Handle
{
public:
  Handle() { open(); }
  Handle(const Handle &) { open(); }
  Handle(Handle && o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  Handle & operator=(const Handle &) { return *this; }
  Handle & operator=(Handle && o) noexcept { std::swap(h_, o.h_); return *this; }
  ~Handle() { if (h_) { sfb_destroy(h_); } }
  sfb_handle_t get() const { return h_; }

private:
  void open()
  {
    if (sfb_create(0, nullptr, &h_) != SFB_OK) {
      throw std::runtime_error(std::string("smooth::feedback (B200 engine): ") + sfb_last_error_message(nullptr));
    }
  }
  sfb_handle_t h_{nullptr};
};

/// Owner of an sfb_qp_sparse_pattern_t; a copy starts without a pattern (re-analysed on its first solve)
class Pattern
{
public:
  Pattern() = default;
  Pattern(const Pattern &) {}
  Pattern(Pattern && o) noexcept : p_(o.p_) { o.p_ = nullptr; }
  Pattern & operator=(const Pattern &) { reset(); return *this; }
  Pattern & operator=(Pattern && o) noexcept { std::swap(p_, o.p_); return *this; }
  ~Pattern() { reset(); }
  void reset() { if (p_) { sfb_qp_sparse_pattern_destroy(p_); p_ = nullptr; } }
  sfb_qp_sparse_pattern_t get() const { return p_; }
  sfb_qp_sparse_pattern_t * put() { reset(); return &p_; }

private:
  sfb_qp_sparse_pattern_t p_{nullptr};
};

}  // namespace detail

template<typename Pbm>
class QPSolver
{
  using AmatT                  = decltype(Pbm::A);
  using Scalar                 = typename AmatT::Scalar;
  static constexpr bool sparse = std::is_base_of_v<Eigen::SparseMatrixBase<AmatT>, AmatT>;  // qp_solver.hpp:247
  static constexpr Eigen::Index M = AmatT::RowsAtCompileTime;
  static constexpr Eigen::Index N = AmatT::ColsAtCompileTime;
  static_assert(std::is_same_v<Scalar, double> || std::is_same_v<Scalar, float>, "double or float problems only");

public:
  using Solution = QPSolution<M, N, Scalar>;  // == detail::qp_solution_t<Pbm>

  QPSolver(const QPSolverParams & prm = {}) : prm_(prm) {}
  QPSolver(const Pbm & pbm, const QPSolverParams & prm = {}) : prm_(prm) { analyze(pbm); }

  /// Access most recent QP solution (qp_solver.hpp:292)
  const Solution & sol() const { return sol_; }

  /// Prepare for solving problems (qp_solver.hpp:297-338): sizes the staging buffers, zeroes the solution
  void analyze(const Pbm & pbm)
  {
    n_ = static_cast<int>(pbm.A.cols());
    m_ = static_cast<int>(pbm.A.rows());
    if constexpr (sparse) { pattern_.reset(); }
    sol_.primal.resize(n_);
    sol_.dual.resize(m_);
    for (int i = 0; i < n_; ++i) { sol_.primal(i) = 0; }
    for (int i = 0; i < m_; ++i) { sol_.dual(i) = 0; }
  }

  /// Solve quadratic program (qp_solver.hpp:343-568)
  const Solution &
  solve(const Pbm & pbm, std::optional<std::reference_wrapper<const Solution>> warmstart = {})
  {
    std::vector<Solution> out(1);
    if (warmstart.has_value()) {
      const Solution & ws = warmstart.value().get();
      solve_batch(std::span<const Pbm>(&pbm, 1), std::span<Solution>(out), std::span<const Solution>(&ws, 1));
    } else {
      solve_batch(std::span<const Pbm>(&pbm, 1), std::span<Solution>(out));
    }
    sol_ = std::move(out[0]);
    return sol_;
  }

  /// EXTENSION: solve many independent problems of one shape in one GPU pass
  void solve_batch(std::span<const Pbm> pbms, std::span<Solution> sols, std::span<const Solution> warm = {})
  {
    if (pbms.empty()) { return; }
    const int64_t B = static_cast<int64_t>(pbms.size());
    const int n = static_cast<int>(pbms[0].A.cols()), m = static_cast<int>(pbms[0].A.rows());
    if constexpr (sparse) {
      solve_batch_sparse(pbms, sols, warm, B, n, m);
    } else {
      solve_batch_dense(pbms, sols, warm, B, n, m);
    }
  }

private:
  void solve_batch_dense(std::span<const Pbm> pbms, std::span<Solution> sols, std::span<const Solution> warm, int64_t B, int n, int m)
  {
    P_.resize(B * n * n); q_.resize(B * n); A_.resize(B * m * n); l_.resize(B * m); u_.resize(B * m);
    x_.resize(B * n); y_.resize(B * m); obj_.resize(B); st_.resize(B); it_.resize(B);
    const bool has_warm = !warm.empty();
    if (has_warm) { wx_.resize(B * n); wy_.resize(B * m); }
    for (int64_t b = 0; b < B; ++b) {  // marshal into the C ABI's column-major batch layout (sfb.h)
      const Pbm & p = pbms[b];
      for (int j = 0; j < n; ++j) {
        for (int i = 0; i < n; ++i) { P_[(b * n + j) * n + i] = p.P(i, j); }
        for (int i = 0; i < m; ++i) { A_[(b * n + j) * m + i] = p.A(i, j); }
        q_[b * n + j] = p.q(j);
        if (has_warm) { wx_[b * n + j] = warm[b].primal(j); }
      }
      for (int i = 0; i < m; ++i) {
        l_[b * m + i] = p.l(i);
        u_[b * m + i] = p.u(i);
        if (has_warm) { wy_[b * m + i] = warm[b].dual(i); }
      }
    }
    const sfb_qp_params c = detail::to_c(prm_);
    int rc;
    if constexpr (std::is_same_v<Scalar, double>) {
      rc = sfb_qp_solve_dense_batch_f64(handle_.get(), &c, B, n, m, P_.data(), q_.data(), A_.data(), l_.data(),
        u_.data(), has_warm ? wx_.data() : nullptr, has_warm ? wy_.data() : nullptr, x_.data(), y_.data(), obj_.data(),
        st_.data(), it_.data(), nullptr, nullptr);
    } else {
      rc = sfb_qp_solve_dense_batch_f32(handle_.get(), &c, B, n, m, P_.data(), q_.data(), A_.data(), l_.data(),
        u_.data(), has_warm ? wx_.data() : nullptr, has_warm ? wy_.data() : nullptr, x_.data(), y_.data(), obj_.data(),
        st_.data(), it_.data(), nullptr, nullptr);
    }
    if (rc != SFB_OK) {
      // API misuse / CUDA failure (never a per-instance numerical outcome): surface loudly
      throw std::runtime_error(std::string("sfb_qp_solve_dense_batch: ") + sfb_last_error_message(handle_.get()));
    }
    for (int64_t b = 0; b < B; ++b) {
      Solution & s = sols[b];
      s.code = static_cast<QPSolutionStatus>(st_[b]);
      s.iter = it_[b];
      s.objective = obj_[b];
      s.primal.resize(n);
      s.dual.resize(m);
      for (int j = 0; j < n; ++j) { s.primal(j) = x_[b * n + j]; }
      for (int i = 0; i < m; ++i) { s.dual(i) = y_[b * m + i]; }
    }
  }

  /// sparse problems sharing one pattern (QuadraticProgramSparse, the MPC call site mpc.hpp:491)
  void solve_batch_sparse(std::span<const Pbm> pbms, std::span<Solution> sols, std::span<const Solution> warm, int64_t B, int n, int m)
  {
    const Pbm & p0 = pbms[0];
    if (!p0.P.isCompressed() || !p0.A.isCompressed()) {
      throw std::invalid_argument("smooth::feedback (B200 engine): sparse problems must be in compressed storage (makeCompressed)");
    }
    const int64_t nnzP = static_cast<int64_t>(p0.P.nonZeros()), nnzA = static_cast<int64_t>(p0.A.nonZeros());
    if (!pattern_.get() || n != n_ || m != m_) {  // analyzePattern, once per solver object (qp_solver.hpp:424)
      n_ = n; m_ = m;
      std::vector<int32_t> pc(p0.P.outerIndexPtr(), p0.P.outerIndexPtr() + n + 1), pr(p0.P.innerIndexPtr(), p0.P.innerIndexPtr() + nnzP);
      std::vector<int32_t> ar(p0.A.outerIndexPtr(), p0.A.outerIndexPtr() + m + 1), ac(p0.A.innerIndexPtr(), p0.A.innerIndexPtr() + nnzA);
      if (sfb_qp_sparse_analyze(handle_.get(), n, m, pc.data(), pr.data(), ar.data(), ac.data(), pattern_.put()) != SFB_OK) {
        throw std::runtime_error(std::string("sfb_qp_sparse_analyze: ") + sfb_last_error_message(handle_.get()));
      }
      pat_P_ = std::move(pr); pat_A_ = std::move(ac); pat_Po_ = std::move(pc); pat_Ao_ = std::move(ar);
    }
    P_.resize(B * nnzP); q_.resize(B * n); A_.resize(B * nnzA); l_.resize(B * m); u_.resize(B * m);
    x_.resize(B * n); y_.resize(B * m); obj_.resize(B); st_.resize(B); it_.resize(B);
    const bool has_warm = !warm.empty();
    if (has_warm) { wx_.resize(B * n); wy_.resize(B * m); }
    for (int64_t b = 0; b < B; ++b) {
      const Pbm & p = pbms[b];
      if (!p.P.isCompressed() || !p.A.isCompressed()) {
        throw std::invalid_argument("smooth::feedback (B200 engine): sparse problems must be in compressed storage (makeCompressed)");
      }
      if (static_cast<int64_t>(p.P.nonZeros()) != nnzP || static_cast<int64_t>(p.A.nonZeros()) != nnzA ||
          !std::equal(pat_Po_.begin(), pat_Po_.end(), p.P.outerIndexPtr()) || !std::equal(pat_Ao_.begin(), pat_Ao_.end(), p.A.outerIndexPtr()) ||
          !std::equal(pat_P_.begin(), pat_P_.end(), p.P.innerIndexPtr()) || !std::equal(pat_A_.begin(), pat_A_.end(), p.A.innerIndexPtr())) {
        throw std::invalid_argument("smooth::feedback (B200 engine): problems of one batch must share the analysed sparsity pattern");
      }
      std::copy(p.P.valuePtr(), p.P.valuePtr() + nnzP, P_.begin() + b * nnzP);
      std::copy(p.A.valuePtr(), p.A.valuePtr() + nnzA, A_.begin() + b * nnzA);
      for (int j = 0; j < n; ++j) {
        q_[b * n + j] = p.q(j);
        if (has_warm) { wx_[b * n + j] = warm[b].primal(j); }
      }
      for (int i = 0; i < m; ++i) {
        l_[b * m + i] = p.l(i);
        u_[b * m + i] = p.u(i);
        if (has_warm) { wy_[b * m + i] = warm[b].dual(i); }
      }
    }
    const sfb_qp_params c = detail::to_c(prm_);
    int rc;
    if constexpr (std::is_same_v<Scalar, double>) {
      rc = sfb_qp_solve_sparse_batch_f64(handle_.get(), pattern_.get(), &c, B, P_.data(), q_.data(), A_.data(), l_.data(), u_.data(),
        has_warm ? wx_.data() : nullptr, has_warm ? wy_.data() : nullptr, x_.data(), y_.data(), obj_.data(), st_.data(), it_.data(),
        nullptr, nullptr);
    } else {
      rc = sfb_qp_solve_sparse_batch_f32(handle_.get(), pattern_.get(), &c, B, P_.data(), q_.data(), A_.data(), l_.data(), u_.data(),
        has_warm ? wx_.data() : nullptr, has_warm ? wy_.data() : nullptr, x_.data(), y_.data(), obj_.data(), st_.data(), it_.data(),
        nullptr, nullptr);
    }
    if (rc != SFB_OK) { throw std::runtime_error(std::string("sfb_qp_solve_sparse_batch: ") + sfb_last_error_message(handle_.get())); }
    for (int64_t b = 0; b < B; ++b) {
      Solution & s = sols[b];
      s.code = static_cast<QPSolutionStatus>(st_[b]);
      s.iter = it_[b];
      s.objective = obj_[b];
      s.primal.resize(n);
      s.dual.resize(m);
      for (int j = 0; j < n; ++j) { s.primal(j) = x_[b * n + j]; }
      for (int i = 0; i < m; ++i) { s.dual(i) = y_[b * m + i]; }
    }
  }

  QPSolverParams prm_{};
  Solution sol_{};
  int n_{0}, m_{0};
  detail::Handle handle_{};
  detail::Pattern pattern_{};
  std::vector<int32_t> pat_P_, pat_A_, pat_Po_, pat_Ao_;
  std::vector<Scalar> P_, q_, A_, l_, u_, wx_, wy_, x_, y_, obj_;
  std::vector<int32_t> st_;
  std::vector<uint32_t> it_;
};

/// One-off solve (qp_solver.hpp:779-787): fresh solver per call
template<typename Pbm>
detail::qp_solution_t<Pbm> solve_qp(
  const Pbm & pbm,
  const QPSolverParams & prm,
  std::optional<std::reference_wrapper<const detail::qp_solution_t<Pbm>>> warmstart = {})
{
  QPSolver<Pbm> solver(pbm, prm);
  return solver.solve(pbm, warmstart);
}

}  // namespace smooth::feedback
