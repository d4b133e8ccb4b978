// TEST INFRASTRUCTURE: the type templates of the reference's include/smooth/feedback/qp.hpp (QuadraticProgram :31-45,
// QuadraticProgramSparse :60-79, QPSolutionStatus :82-92, QPSolution :95-108) with the same names, template parameters,
// member names and member types, declared on the Eigen stand-in of this directory.  It exists so that the overlay
// (include/smooth_feedback_b200/qp_solver.hpp, which includes <smooth/feedback/qp.hpp> and must NOT redefine these types)
// can be compile-checked in an image without Eigen; with the reference's include tree on the path, its own qp.hpp is used.
#pragma once
#include <cstdint>
#include <Eigen/Dense>
#include <Eigen/Sparse>
namespace smooth::feedback {
template<Eigen::Index M, Eigen::Index N, typename Scalar = double>
struct QuadraticProgram
{
  Eigen::Matrix<Scalar, N, N> P;
  Eigen::Matrix<Scalar, N, 1> q;
  Eigen::Matrix<Scalar, M, N> A;
  Eigen::Matrix<Scalar, M, 1> l;
  Eigen::Matrix<Scalar, M, 1> u;
};
template<typename Scalar = double>
struct QuadraticProgramSparse
{
  Eigen::SparseMatrix<Scalar> P;
  Eigen::Matrix<Scalar, -1, 1> q;
  Eigen::SparseMatrix<Scalar, Eigen::RowMajor> A;
  Eigen::Matrix<Scalar, -1, 1> l;
  Eigen::Matrix<Scalar, -1, 1> u;
};
enum class QPSolutionStatus { Optimal, PolishFailed, PrimalInfeasible, DualInfeasible, MaxIterations, MaxTime, Unknown };
template<Eigen::Index M, Eigen::Index N, typename Scalar = double>
struct QPSolution
{
  QPSolutionStatus code = QPSolutionStatus::Unknown;
  uint32_t iter;
  Eigen::Matrix<Scalar, N, 1> primal;
  Eigen::Matrix<Scalar, M, 1> dual;
  Scalar objective{0.};
};
}  // namespace smooth::feedback
