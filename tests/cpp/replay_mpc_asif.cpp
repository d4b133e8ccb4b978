// Replays, verbatim, the member declarations and the solver calls of the reference's two QP call sites against the overlay
// (include/smooth_feedback_b200/qp_solver.hpp), with the reference's OWN problem / solution type templates
// (<smooth/feedback/qp.hpp>; here the stand-in under tests/cpp/mock_include, with Eigen installed the reference's file):
//
//   mpc.hpp:629-635   QuadraticProgramSparse<double> qp_;  QPSolver<QuadraticProgramSparse<double>> qp_solver_;
//                     std::optional<QPSolution<-1, -1, double>> warmstart_{};
//   mpc.hpp:422-424   qp_solver_{prm_.qp};  ...  qp_solver_.analyze(qp_);
//   mpc.hpp:491       const auto & sol = qp_solver_.solve(qp_, warmstart_);
//   mpc.hpp:510-516   if (sol.code == Optimal || MaxTime || MaxIterations) { warmstart_ = sol; }
//   mpc.hpp:518       sol.primal.template segment<Nu>(uvar_B)        (element access here)
//   asif.hpp:107-109  QuadraticProgram<-1, -1, double> qp_;  std::optional<QPSolution<-1, -1, double>> warmstart_;
//   asif.hpp:97-99    auto sol = feedback::solve_qp(qp_, prm_.qp, warmstart_);  if (sol.code == Optimal) { warmstart_ = sol; }
//
// If the overlay redefined QPSolution / QPSolutionStatus, or returned its own solution type, this file would not compile --
// exactly the defect the round-1 overlay had.  Needs a GPU to run; compiles anywhere.
#include <cstdio>
#include <cmath>
#include <limits>
#include <smooth_feedback_b200/qp_solver.hpp>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

namespace smooth::feedback {

struct MPCParamsReplay { QPSolverParams qp{}; bool warmstart{true}; };

// the members and calls of MPC<...> that touch the QP solver (mpc.hpp:405-425, 458-519, 629-635)
class MpcReplay
{
public:
  explicit MpcReplay(MPCParamsReplay && prm = {}) : prm_{std::move(prm)}, qp_solver_{prm_.qp}
  {
    // ocp_to_qp_allocate / update stand-in: min 1/2 |x|^2 - 4 x0 + x1/4, -1 <= x <= 1 (tests/test_qp.cpp:100-122 BasicSparse)
    qp_.P.set(2, 2, {0, 1, 2}, {0, 1}, {1, 1});
    qp_.A.set(2, 2, {0, 1, 2}, {0, 1}, {1, 1});
    qp_.q.resize(2); qp_.q(0) = -4; qp_.q(1) = 0.25;
    qp_.l.resize(2); qp_.l(0) = -1; qp_.l(1) = -1;
    qp_.u.resize(2); qp_.u(0) = 1; qp_.u(1) = 1;
    qp_solver_.analyze(qp_);
  }

  std::pair<double, QPSolutionStatus> operator()()
  {
    qp_.A.makeCompressed();
    qp_.P.makeCompressed();

    // solve QP
    const auto & sol = qp_solver_.solve(qp_, warmstart_);

    // save solution to warmstart next iteration
    if (prm_.warmstart) {
      // clang-format off
      if (sol.code == QPSolutionStatus::Optimal || sol.code == QPSolutionStatus::MaxTime || sol.code == QPSolutionStatus::MaxIterations) {
        warmstart_ = sol;
      }
      // clang-format on
    }
    iters_ = sol.iter;
    return {sol.primal(0), sol.code};
  }

  void reset_warmstart() { warmstart_ = {}; }
  uint32_t iters_{0};

private:
  MPCParamsReplay prm_{};
  QuadraticProgramSparse<double> qp_;
  QPSolver<QuadraticProgramSparse<double>> qp_solver_;
  std::optional<QPSolution<-1, -1, double>> warmstart_{};
};

// the members and calls of ASIFilter<...> that touch the QP solver (asif.hpp:82-110)
class AsifReplay
{
public:
  AsifReplay()
  {
    qp_.A.resize(2, 2); qp_.P.resize(2, 2); qp_.q.resize(2); qp_.l.resize(2); qp_.u.resize(2);
    qp_.P(0, 0) = 1; qp_.P(1, 1) = 1; qp_.A(0, 0) = 1; qp_.A(1, 1) = 1;
    qp_.q(0) = -4; qp_.q(1) = 0.25; qp_.l(0) = -1; qp_.l(1) = -1; qp_.u(0) = 1; qp_.u(1) = 1;
    prm_qp_.polish = false;
  }

  std::pair<double, QPSolutionStatus> operator()()
  {
    auto sol = feedback::solve_qp(qp_, prm_qp_, warmstart_);

    if (sol.code == QPSolutionStatus::Optimal) { warmstart_ = sol; }

    iters_ = sol.iter;
    return {sol.primal(0), sol.code};
  }
  uint32_t iters_{0};

private:
  QuadraticProgram<-1, -1, double> qp_;
  QPSolverParams prm_qp_{};
  std::optional<QPSolution<-1, -1, double>> warmstart_;
};

}  // namespace smooth::feedback

int main()
{
  using namespace smooth::feedback;
  static_assert(std::is_same_v<detail::qp_solution_t<QuadraticProgramSparse<double>>, QPSolution<-1, -1, double>>);
  static_assert(std::is_same_v<detail::qp_solution_t<QuadraticProgram<2, 2, double>>, QPSolution<2, 2, double>>);
  static_assert(std::is_same_v<decltype(std::declval<QPSolver<QuadraticProgram<-1, -1, double>> &>().solve(std::declval<const QuadraticProgram<-1, -1, double> &>())),
                               const QPSolution<-1, -1, double> &>);
  static_assert(std::is_copy_constructible_v<QPSolver<QuadraticProgramSparse<double>>> && std::is_move_assignable_v<QPSolver<QuadraticProgramSparse<double>>>);

  MpcReplay mpc;
  auto [u0, c0] = mpc();
  CHECK(c0 == QPSolutionStatus::Optimal && std::fabs(u0 - 1) < 1e-4 && mpc.iters_ > 2);
  auto [u1, c1] = mpc();  // warm start from the stored solution: exits at the first stop check
  CHECK(c1 == QPSolutionStatus::Optimal && mpc.iters_ == 2 && std::fabs(u1 - u0) < 1e-9);
  MpcReplay copy = mpc;   // Mpc.Constructors (tests/test_mpc.cpp:120-164): a copy gives the same input
  auto [u2, c2] = copy();
  CHECK(c2 == QPSolutionStatus::Optimal && std::fabs(u2 - u0) < 1e-9);
  mpc.reset_warmstart();
  auto [u3, c3] = mpc();
  CHECK(c3 == QPSolutionStatus::Optimal && mpc.iters_ > 2 && u3 == u0);

  AsifReplay asif;
  auto [a0, d0] = asif();
  CHECK(d0 == QPSolutionStatus::Optimal && std::fabs(a0 - 1) < 5e-3 && asif.iters_ > 2);
  auto [a1, d1] = asif();
  CHECK(d1 == QPSolutionStatus::Optimal && asif.iters_ == 2);
  (void)a1;
  std::printf("replay ok\n");
  return 0;
}
