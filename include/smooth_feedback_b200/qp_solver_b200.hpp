// qp_solver_b200.hpp -- C++20 host-side overlay: smooth_feedback's QP solver surface on top of the sfb C ABI.
//
// Same names, argument meaning and error behaviour as the reference headers
//   include/smooth/feedback/qp.hpp:82-108          QPSolutionStatus, QPSolution
//   include/smooth/feedback/qp_solver.hpp:29-68    QPSolverParams
//   include/smooth/feedback/qp_solver.hpp:242-757  QPSolver<Pbm>::{QPSolver, analyze, solve, sol}
//   include/smooth/feedback/qp_solver.hpp:779-787  solve_qp
// so that `qp_solver_.solve(qp_, warmstart_)` (mpc.hpp:491) and `solve_qp(qp_, prm_.qp, warmstart_)` (asif.hpp:97)
// compile unchanged; plus the one extension the GPU engine exists for: QPSolver::solve_batch.
//
// `Pbm` is any type with public dense members P, q, A, l, u that offer rows(), cols(), operator()(i,j) / operator()(i)
// (Eigen matrices do; tests/cpp/mock_eigen.hpp is the stand-in used where Eigen is not installed).  Dense problems only:
// the sparse MPC-sized path is not part of this engine yet.
//
// All numerics run on the GPU through libsfb.so; there is no CPU fallback: if no device is available the constructor of
// the solver throws std::runtime_error with the library's message.
#pragma once

#include <chrono>
#include <cstdint>
#include <functional>
#include <optional>
#include <span>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <vector>

#include "../sfb.h"

namespace smooth::feedback {

/// Solver exit codes (qp.hpp:82-92; the numeric values are the C ABI's sfb_qp_status)
enum class QPSolutionStatus {
  Optimal,
  PolishFailed,
  PrimalInfeasible,
  DualInfeasible,
  MaxIterations,
  MaxTime,
  Unknown
};

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

/// Solution (qp.hpp:95-108); PrimalT / DualT are the problem's own vector types
template<typename PrimalT, typename DualT, typename Scalar = double>
struct QPSolutionT
{
  QPSolutionStatus code = QPSolutionStatus::Unknown;
  uint32_t iter{0};
  PrimalT primal{};
  DualT dual{};
  Scalar objective{0};
};

namespace detail {

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

}  // namespace detail

template<typename Pbm>
class QPSolver
{
  using Scalar  = std::remove_cvref_t<decltype(std::declval<Pbm>().q(0))>;
  using PrimalT = std::remove_cvref_t<decltype(Pbm::q)>;
  using DualT   = std::remove_cvref_t<decltype(Pbm::l)>;
  static_assert(std::is_same_v<Scalar, double> || std::is_same_v<Scalar, float>, "double or float problems only");

public:
  using Solution = QPSolutionT<PrimalT, DualT, Scalar>;

  QPSolver(const QPSolverParams & prm = {}) : prm_(prm) {}
  QPSolver(const Pbm & pbm, const QPSolverParams & prm = {}) : prm_(prm) { analyze(pbm); }

  /// Access most recent QP solution (qp_solver.hpp:292)
  const Solution & sol() const { return sol_; }

  /// Prepare for solving problems (qp_solver.hpp:297-338): sizes the staging buffers, zeroes the solution
  void analyze(const Pbm & pbm)
  {
    n_ = static_cast<int>(pbm.A.cols());
    m_ = static_cast<int>(pbm.A.rows());
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

private:
  QPSolverParams prm_{};
  Solution sol_{};
  int n_{0}, m_{0};
  detail::Handle handle_{};
  std::vector<Scalar> P_, q_, A_, l_, u_, wx_, wy_, x_, y_, obj_;
  std::vector<int32_t> st_;
  std::vector<uint32_t> it_;
};

/// One-off solve (qp_solver.hpp:779-787): fresh solver per call
template<typename Pbm>
typename QPSolver<Pbm>::Solution solve_qp(
  const Pbm & pbm,
  const QPSolverParams & prm,
  std::optional<std::reference_wrapper<const typename QPSolver<Pbm>::Solution>> warmstart = {})
{
  QPSolver<Pbm> solver(pbm, prm);
  return solver.solve(pbm, warmstart);
}

}  // namespace smooth::feedback
