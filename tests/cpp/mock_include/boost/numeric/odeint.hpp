// TEST INFRASTRUCTURE: Boost.odeint's euler / runge_kutta4 steppers with vector_space_algebra, restated from their
// published one-step formulas (Boost is not installed in this image).  Same template parameter list and do_step signature.
#pragma once
namespace boost::numeric::odeint {
struct vector_space_algebra {};
template<class State, class Value = double, class Deriv = State, class Time = Value, class Algebra = vector_space_algebra, class...>
struct euler
{
  template<class System> void do_step(System sys, State & x, Time t, Time dt)
  {
    Deriv d{};
    sys(x, d, t);
    x = x + static_cast<Value>(dt) * d;
  }
};
template<class State, class Value = double, class Deriv = State, class Time = Value, class Algebra = vector_space_algebra, class...>
struct runge_kutta4
{
  template<class System> void do_step(System sys, State & x, Time t, Time dt)
  {
    Deriv k1{}, k2{}, k3{}, k4{};
    sys(x, k1, t);
    sys(x + static_cast<Value>(dt / 2) * k1, k2, t + dt / 2);
    sys(x + static_cast<Value>(dt / 2) * k2, k3, t + dt / 2);
    sys(x + static_cast<Value>(dt) * k3, k4, t + dt);
    x = x + static_cast<Value>(dt / 6) * (k1 + static_cast<Value>(2) * k2 + static_cast<Value>(2) * k3 + k4);
  }
};
}  // namespace boost::numeric::odeint
