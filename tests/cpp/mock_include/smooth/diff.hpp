// TEST INFRASTRUCTURE: smooth::diff::dr<1, Type>(f, wrt(x)) for R^n by central differences (the reference's default is
// forward-mode autodiff; on the linear test models of tests/test_ekf.cpp both give the exact Jacobian up to rounding).
#pragma once
#include <tuple>
#include <utility>
#include "concepts/lie_group.hpp"
namespace smooth {
template<typename... Args> auto wrt(Args &&... a) { return std::forward_as_tuple(std::forward<Args>(a)...); }
namespace diff {
enum class Type { Numerical, Autodiff, Ceres, Analytic, Default };
template<int K, Type DT = Type::Default, typename F, typename Wrt>
auto dr(F && f, Wrt && x)
{
  static_assert(K == 1);
  using X = std::decay_t<std::tuple_element_t<0, std::decay_t<Wrt>>>;
  const X & x0 = std::get<0>(x);
  auto val = f(x0);
  using Y = decltype(val);
  Eigen::Matrix<typename X::Scalar, Y::RowsAtCompileTime, X::RowsAtCompileTime> J;
  const double h = 1e-6;
  for (Eigen::Index j = 0; j < x0.size(); ++j) {
    X xp = x0, xm = x0;
    xp(j) += h; xm(j) -= h;
    const auto fp = f(xp), fm = f(xm);
    for (Eigen::Index i = 0; i < val.size(); ++i) { J(i, j) = (fp(i) - fm(i)) / (2 * h); }
  }
  return std::make_pair(val, J);
}
}  // namespace diff
}  // namespace smooth
