// The reference's own BasicDynamic / Portfolio / SolverAPI unit tests (tests/test_qp.cpp:75-98, :244-272, :338-372) written
// against the overlay header and the reference's own QuadraticProgram / QPSolution templates (Eigen stand-in: tests/cpp/mock_include).  Needs a GPU to run; compiles anywhere.
#include <cmath>
#include <cstdio>
#include <limits>
#include <smooth_feedback_b200/qp_solver.hpp>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
  using namespace smooth::feedback;
  using Pbm = QuadraticProgram<-1, -1, double>;
  constexpr double inf = std::numeric_limits<double>::infinity();
  const QPSolverParams test_prm{.verbose = false, .polish = true};

  Pbm basic;
  basic.P.resize(2, 2); basic.P(0, 0) = 1; basic.P(1, 1) = 1;
  basic.q.resize(2); basic.q(0) = -4; basic.q(1) = 0.25;
  basic.A.resize(2, 2); basic.A(0, 0) = 1; basic.A(1, 1) = 1;
  basic.l.resize(2); basic.l(0) = -1; basic.l(1) = -1;
  basic.u.resize(2); basic.u(0) = 1; basic.u(1) = 1;

  auto sol = solve_qp(basic, test_prm);
  CHECK(sol.code == QPSolutionStatus::Optimal);
  CHECK(std::fabs(sol.primal(0) - 1) < 1e-4 && std::fabs(sol.primal(1) + 0.25) < 1e-4);
  CHECK(std::fabs(sol.objective - (0.5 - 4 - 1. / 32)) < 1e-4);
  auto sol_hs = solve_qp(basic, test_prm, sol);  // warm start with own solution (tests/test_qp.cpp:92-97)
  static_assert(std::is_same_v<decltype(sol), QPSolution<-1, -1, double>>);
  CHECK(sol_hs.code == QPSolutionStatus::Optimal && sol_hs.iter == 2);

  // SolverAPI: copies and moved-from solvers give the same primal
  QPSolver<Pbm> s1(basic, test_prm);
  const auto x1 = s1.solve(basic).primal;
  QPSolver<Pbm> s2 = s1;
  const auto x2 = s2.solve(basic).primal;
  QPSolver<Pbm> s3(std::move(s2));
  const auto x3 = s3.solve(basic).primal;
  CHECK(x1(0) == x2(0) && x1(1) == x2(1) && x1(0) == x3(0) && x1(1) == x3(1));

  // trivially infeasible -> status only, never an exception
  Pbm bad = basic; bad.l(1) = 1; bad.u(1) = -1;
  CHECK(solve_qp(bad, test_prm).code == QPSolutionStatus::PrimalInfeasible);

  // extension: a batch
  std::vector<Pbm> batch(100, basic);
  for (int i = 0; i < 100; ++i) { batch[i].q(0) = -4 + 0.01 * i; }
  std::vector<QPSolver<Pbm>::Solution> sols(100);
  s1.solve_batch(batch, sols);
  for (int i = 0; i < 100; ++i) { CHECK(sols[i].code == QPSolutionStatus::Optimal && std::fabs(sols[i].primal(0) - 1) < 1e-4); }
  // BasicSparse (tests/test_qp.cpp:100-122): the same QP as QuadraticProgramSparse -- the problem type MPC instantiates
  {
    using SPbm = QuadraticProgramSparse<double>;
    SPbm sp;
    sp.P.set(2, 2, {0, 1, 2}, {0, 1}, {1, 1});
    sp.A.set(2, 2, {0, 1, 2}, {0, 1}, {1, 1});
    sp.q.resize(2); sp.q(0) = -4; sp.q(1) = 0.25;
    sp.l.resize(2); sp.l(0) = -1; sp.l(1) = -1;
    sp.u.resize(2); sp.u(0) = 1; sp.u(1) = 1;
    QPSolver<SPbm> ss(sp, test_prm);
    const auto ssol = ss.solve(sp);
    CHECK(ssol.code == QPSolutionStatus::Optimal);
    CHECK(std::fabs(ssol.primal(0) - 1) < 1e-4 && std::fabs(ssol.primal(1) + 0.25) < 1e-4);
    CHECK(std::fabs(ssol.objective - (0.5 - 4 - 1. / 32)) < 1e-4);
    const auto swarm = ss.solve(sp, ssol);
    CHECK(swarm.code == QPSolutionStatus::Optimal && swarm.iter == 2);
    QPSolver<SPbm> ss2 = ss;  // a copy drops the analysed pattern and re-analyses (LDLTWrapper semantics, SparseSolverAPI :374-415)
    const auto x2s = ss2.solve(sp).primal;
    CHECK(x2s(0) == ssol.primal(0) && x2s(1) == ssol.primal(1));
    std::vector<SPbm> sbatch(50, sp);
    for (int i = 0; i < 50; ++i) { sbatch[i].q(0) = -4 + 0.01 * i; }
    std::vector<QPSolver<SPbm>::Solution> ssols(50);
    ss.solve_batch(sbatch, ssols);
    for (int i = 0; i < 50; ++i) { CHECK(ssols[i].code == QPSolutionStatus::Optimal && std::fabs(ssols[i].primal(0) - 1) < 1e-4); }
    sbatch[7].A.inner = {1, 0};  // a different pattern inside one batch is an API error, not a status
    bool threw = false;
    try { ss.solve_batch(sbatch, ssols); } catch (const std::invalid_argument &) { threw = true; }
    CHECK(threw);
    sbatch[7].A.inner = {0, 1};
    sbatch[7].A.outer = {0, 2, 2};  // same inner indices, different row partition: also a different pattern
    threw = false;
    try { ss.solve_batch(sbatch, ssols); } catch (const std::invalid_argument &) { threw = true; }
    CHECK(threw);
    // BasicStatic (tests/test_qp.cpp:54-73): statically sized problem and solution types
    QuadraticProgram<2, 2, double> st;
    st.P(0, 0) = 1; st.P(1, 1) = 1; st.A(0, 0) = 1; st.A(1, 1) = 1;
    st.q(0) = -4; st.q(1) = 0.25; st.l(0) = -1; st.l(1) = -1; st.u(0) = 1; st.u(1) = 1;
    const QPSolution<2, 2, double> sts = solve_qp(st, test_prm);
    CHECK(sts.code == QPSolutionStatus::Optimal && std::fabs(sts.primal(0) - 1) < 1e-4);
  }
  (void)inf;
  std::printf("overlay ok\n");
  return 0;
}
