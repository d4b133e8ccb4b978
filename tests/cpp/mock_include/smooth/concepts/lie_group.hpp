// TEST INFRASTRUCTURE: the slice of pettni/smooth's LieGroup / Manifold vocabulary that smooth_feedback's ekf.hpp uses,
// for the one group family the reference's own EKF tests run on: R^n = Eigen::Matrix<S, n, 1> (tests/test_ekf.cpp:50-180).
// smooth itself is not installed in this image; with it on the include path this directory is not used.
#pragma once
#include <Eigen/Dense>
namespace smooth {
template<typename G> concept LieGroup = requires { typename G::Scalar; G::RowsAtCompileTime; };
template<typename G> concept Manifold = LieGroup<G>;
template<typename G> inline constexpr Eigen::Index Dof = G::RowsAtCompileTime;
template<typename G> using Scalar = typename G::Scalar;
template<typename G> using Tangent = Eigen::Matrix<Scalar<G>, Dof<G>, 1>;
template<typename G> using TangentMap = Eigen::Matrix<Scalar<G>, Dof<G>, Dof<G>>;
template<typename T, typename G> using CastT = Eigen::Matrix<T, Dof<G>, 1>;
template<typename G> G Default() { return G{}; }
// R^n is commutative: ad = 0; x (+) a = x + a; x (-) y = x - y (operator+ / operator- of the vector type)
template<typename G> TangentMap<G> ad(const Tangent<G> &) { return TangentMap<G>{}; }
}  // namespace smooth
