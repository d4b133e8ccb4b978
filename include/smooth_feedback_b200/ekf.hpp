// ekf.hpp -- drop-in replacement for the reference's include/smooth/feedback/ekf.hpp on top of the sfb C ABI.
//
// Same class template, members and semantics as the reference (pettni/smooth_feedback @ 9a08971):
//   EKF<G, DiffType, Stpr>::{reset, estimate, covariance}   ekf.hpp:45-59
//   EKF::predict(f, Q, tau, dt)                               ekf.hpp:79-103
//   EKF::update(h, y, R)                                      ekf.hpp:116-139
// What stays on the host is what cannot cross a C ABI: the user's callables and their differentiation (diff::dr), the
// group operations on the estimate (g (+) a, y (-) h) and the state stepper.  What moves to the engine is the covariance
// algebra: every sub-step of the covariance ODE  Pdot = symU(A P + P A^T + Q)  (ekf.hpp:84-89) is one
// sfb_ekf_predict_batch_f64 call, and the measurement update  S = triu(H symU(P) H^T + R), K = (S^-1 H P)^T,
// P = symU((I - K H) P)  (ekf.hpp:129-138) is one sfb_ekf_update_batch_f64 call (batch of one filter; many filters of one
// model are what the batch entry points are for).
//
// Step order as in the reference: per sub-step the covariance first (it depends on g_hat_ BEFORE its update, ekf.hpp:94-95),
// then the state; A = -ad(f(t, g_hat_)) + d^r f / dx is RE-EVALUATED at every sub-step (the engine call holds A constant
// over ONE sub-step only, sfb.h).  With runge_kutta4 the reference evaluates A at the four stage times of a sub-step
// (same g_hat_, times t, t + dt/2, t + dt/2, t + dt); here A is evaluated at the sub-step's start -- identical whenever f
// has no explicit time dependence (every EKF model in the reference's tests and examples).
// Stpr must be boost::numeric::odeint::euler (the default) or runge_kutta4: the covariance stepper runs on the device.
#pragma once

#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

#include <Eigen/Cholesky>
#include <boost/numeric/odeint.hpp>
#include <smooth/compat/odeint.hpp>
#include <smooth/concepts/lie_group.hpp>
#include <smooth/diff.hpp>

#include "../sfb.h"

namespace smooth::feedback {

namespace detail {

template<template<typename...> typename Stpr> struct ekf_stepper_id { static constexpr int value = -1; };
template<> struct ekf_stepper_id<boost::numeric::odeint::euler> { static constexpr int value = SFB_STEPPER_EULER; };
template<> struct ekf_stepper_id<boost::numeric::odeint::runge_kutta4> { static constexpr int value = SFB_STEPPER_RK4; };

/// RAII owner of an sfb handle; a copy of a filter gets its own handle
class EkfHandle
{
public:
  EkfHandle() { open(); }
  EkfHandle(const EkfHandle &) { open(); }
  EkfHandle(EkfHandle && o) noexcept : h_(o.h_) { o.h_ = nullptr; }
  EkfHandle & operator=(const EkfHandle &) { return *this; }
  EkfHandle & operator=(EkfHandle && o) noexcept { std::swap(h_, o.h_); return *this; }
  ~EkfHandle() { if (h_) { sfb_destroy(h_); } }
  sfb_handle_t get() const { return h_; }

private:
  void open()
  {
    if (sfb_create(0, nullptr, &h_) != SFB_OK) {
      throw std::runtime_error(std::string("smooth::feedback::EKF (B200 engine): ") + sfb_last_error_message(nullptr));
    }
  }
  sfb_handle_t h_{nullptr};
};

}  // namespace detail

template<
  LieGroup G,
  diff::Type DiffType                 = diff::Type::Default,
  template<typename...> typename Stpr = boost::numeric::odeint::euler>
  requires(Dof<G> > 0)
class EKF
{
  static_assert(detail::ekf_stepper_id<Stpr>::value >= 0, "the device covariance stepper supports odeint::euler and odeint::runge_kutta4");
  static_assert(std::is_same_v<Scalar<G>, double>, "the engine's EKF entry points are fp64 (the reference's arithmetic)");
  static constexpr int D = static_cast<int>(Dof<G>);
  static_assert(D <= 16, "1 <= d <= 16 (sfb.h)");

public:
  /// Covariance matrix type (ekf.hpp:37)
  using CovT = Eigen::Matrix<Scalar<G>, Dof<G>, Dof<G>>;

  /// ekf.hpp:45-49
  void reset(const G & g, const CovT & P)
  {
    g_hat_ = g;
    P_     = P;
  }

  /// ekf.hpp:54
  G estimate() const { return g_hat_; }

  /// ekf.hpp:59
  CovT covariance() const { return P_; }

  /// Propagate the filter through a dynamical model d^r x_t = f(t, x) with process covariance Q over [0, tau] (ekf.hpp:79-103)
  template<typename F, typename QDer>
  void predict(F && f, const Eigen::MatrixBase<QDer> & Qb, Scalar<G> tau, std::optional<Scalar<G>> dt = {})
  {
    const QDer & Q = static_cast<const QDer &>(Qb);
    const auto state_ode = [&f](const G & g, Tangent<G> & dg, Scalar<G> t) { dg = f(t, g); };

    // one covariance sub-step of length delta on the device, A evaluated at (t, g_hat_) -- ekf.hpp:84-89
    const auto cov_step = [this, &f, &Q](Scalar<G> t, Scalar<G> delta) {
      const auto f_x      = [&f, &t]<typename _T>(const CastT<_T, G> & x) -> Tangent<CastT<_T, G>> { return f(t, x); };
      const auto [fv, dr] = diff::dr<1, DiffType>(f_x, wrt(g_hat_));
      const auto adm      = ad<G>(fv);
      double A[D * D], Qc[D * D], Pc[D * D], Pn[D * D];
      for (int j = 0; j < D; ++j) {
        for (int i = 0; i < D; ++i) {
          A[i + D * j]  = -adm(i, j) + dr(i, j);
          Qc[i + D * j] = Q(i, j);
          Pc[i + D * j] = P_(i, j);
        }
      }
      // dt <= 0 -> one step of length tau' = delta (sfb.h); the stepper is the filter's Stpr
      if (sfb_ekf_predict_batch_f64(h_.get(), 1, D, detail::ekf_stepper_id<Stpr>::value, Pc, A, Qc, delta, -1.0, Pn) != SFB_OK) {
        throw std::runtime_error(std::string("sfb_ekf_predict_batch_f64: ") + sfb_last_error_message(h_.get()));
      }
      for (int j = 0; j < D; ++j) {
        for (int i = 0; i < D; ++i) { P_(i, j) = Pn[i + D * j]; }
      }
    };

    Scalar<G> t          = 0;
    const Scalar<G> dt_v = dt.value_or(2 * tau);
    while (t + dt_v < tau) {
      // step covariance first since it depends on g_hat_
      cov_step(t, dt_v);
      sst_.do_step(state_ode, g_hat_, t, dt_v);
      t += dt_v;
    }

    // last step up to time t
    cov_step(t, tau - t);
    sst_.do_step(state_ode, g_hat_, t, tau - t);
  }

  /// Update the filter with a measurement y = h(x) + w, w ~ N(0, R) (ekf.hpp:116-139)
  template<typename F, typename RDev, Manifold Y = std::invoke_result_t<F, G>>
  void update(F && h, const Y & y, const Eigen::MatrixBase<RDev> & Rb)
  {
    const RDev & R       = static_cast<const RDev &>(Rb);
    const auto [hval, H] = diff::dr<1, DiffType>(h, wrt(g_hat_));

    using Result = std::decay_t<decltype(hval)>;
    static_assert(Manifold<Result>, "h(x) is not a Manifold");
    static constexpr int Ny = static_cast<int>(Dof<Result>);
    static_assert(Ny > 0, "h(x) must be statically sized");
    static_assert(Ny <= 16, "1 <= ny <= 16 (sfb.h)");

    const auto innov = y - hval;  // y (-) h(g_hat)
    double Pc[D * D], Hc[Ny * D], Rc[Ny * Ny], iv[Ny], delta[D], Pn[D * D];
    for (int j = 0; j < D; ++j) {
      for (int i = 0; i < D; ++i) { Pc[i + D * j] = P_(i, j); }
      for (int i = 0; i < Ny; ++i) { Hc[i + Ny * j] = H(i, j); }
    }
    for (int j = 0; j < Ny; ++j) {
      for (int i = 0; i < Ny; ++i) { Rc[i + Ny * j] = R(i, j); }
      iv[j] = innov(j);
    }
    if (sfb_ekf_update_batch_f64(h_.get(), 1, D, Ny, Pc, Hc, Rc, iv, delta, Pn) != SFB_OK) {
      throw std::runtime_error(std::string("sfb_ekf_update_batch_f64: ") + sfb_last_error_message(h_.get()));
    }
    // update estimate and covariance (ekf.hpp:137-138): g_hat (+)= K (y (-) h), P = symU((I - K H) P)
    Tangent<G> dg;
    for (int i = 0; i < D; ++i) { dg(i) = delta[i]; }
    g_hat_ += dg;
    for (int j = 0; j < D; ++j) {
      for (int i = 0; i < D; ++i) { P_(i, j) = Pn[i + D * j]; }
    }
  }

private:
  // filter estimate and covariance
  G g_hat_ = Default<G>();
  CovT P_  = CovT::Identity();

  // stepper for the state ODE (the covariance stepper of the reference, cst_, runs on the device)
  Stpr<G, Scalar<G>, Tangent<G>, Scalar<G>, boost::numeric::odeint::vector_space_algebra> sst_{};

  detail::EkfHandle h_{};
};

}  // namespace smooth::feedback
