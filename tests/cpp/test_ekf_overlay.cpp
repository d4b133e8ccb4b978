// The reference's EKF unit tests (tests/test_ekf.cpp:50-180: UpdateLinear<3,3 | 10,3 | 3,10>, PredictLinear<3 | 6 | 9> with
// runge_kutta4 at dt = 1e-3, PredictTimeCut) transliterated against the overlay include/smooth_feedback_b200/ekf.hpp, on
// the stand-ins for Eigen / smooth / Boost.odeint under tests/cpp/mock_include (none of them is installed here).
// Same models, same tolerances.  Needs a GPU to run; compiles anywhere.
#include <cmath>
#include <cstdio>
#include <random>
#include <smooth_feedback_b200/ekf.hpp>

#define CHECK(c) do { if (!(c)) { std::printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

template<Eigen::Index R, Eigen::Index C> using Mat = Eigen::Matrix<double, R, C>;
static std::mt19937 rng(7);
static double rnd() { return std::uniform_real_distribution<double>(-1.0, 1.0)(rng); }
template<Eigen::Index R, Eigen::Index C> Mat<R, C> Random() { Mat<R, C> m; for (int k = 0; k < R * C; ++k) { m.data()[k] = rnd(); } return m; }
template<Eigen::Index N> Mat<N, N> RandomDiag() { Mat<N, N> m; for (int i = 0; i < N; ++i) { m(i, i) = rnd() + 1.1; } return m; }
template<typename S, Eigen::Index R, Eigen::Index K, Eigen::Index C> Eigen::Matrix<S, R, C> mul(const Mat<R, K> & a, const Eigen::Matrix<S, K, C> & b)
{
  Eigen::Matrix<S, R, C> o;
  for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) { S s = 0; for (int k = 0; k < K; ++k) { s += a(i, k) * b(k, j); } o(i, j) = s; } }
  return o;
}
template<Eigen::Index R, Eigen::Index C> Mat<C, R> tr(const Mat<R, C> & a) { Mat<C, R> o; for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) { o(j, i) = a(i, j); } } return o; }
template<Eigen::Index N> Mat<N, N> inverse(Mat<N, N> a)
{
  Mat<N, N> inv = Mat<N, N>::Identity();
  for (int k = 0; k < N; ++k) {
    int p = k;
    for (int i = k + 1; i < N; ++i) { if (std::fabs(a(i, k)) > std::fabs(a(p, k))) { p = i; } }
    for (int j = 0; j < N; ++j) { std::swap(a(k, j), a(p, j)); std::swap(inv(k, j), inv(p, j)); }
    const double d = a(k, k);
    for (int j = 0; j < N; ++j) { a(k, j) /= d; inv(k, j) /= d; }
    for (int i = 0; i < N; ++i) {
      if (i == k) { continue; }
      const double f = a(i, k);
      for (int j = 0; j < N; ++j) { a(i, j) -= f * a(k, j); inv(i, j) -= f * inv(k, j); }
    }
  }
  return inv;
}
template<Eigen::Index N> Mat<N, N> expm(const Mat<N, N> & a)  // scaling and squaring with a Taylor series
{
  Mat<N, N> s = a;
  for (int k = 0; k < N * N; ++k) { s.data()[k] /= 1024.0; }
  Mat<N, N> e = Mat<N, N>::Identity(), term = Mat<N, N>::Identity();
  for (int k = 1; k < 20; ++k) { term = mul(term, s); for (int q = 0; q < N * N; ++q) { term.data()[q] /= k; } e += term; }
  for (int k = 0; k < 10; ++k) { e = mul(e, e); }
  return e;
}
template<Eigen::Index R, Eigen::Index C> bool isApprox(const Mat<R, C> & a, const Mat<R, C> & b, double prec)  // Eigen: |a - b|_F <= prec min(|a|_F, |b|_F)
{
  double d = 0, na = 0, nb = 0;
  for (int k = 0; k < R * C; ++k) { d += (a.data()[k] - b.data()[k]) * (a.data()[k] - b.data()[k]); na += a.data()[k] * a.data()[k]; nb += b.data()[k] * b.data()[k]; }
  return std::sqrt(d) <= prec * std::sqrt(std::min(na, nb));
}

template<int Nx, int Ny> int test_update_linear()
{
  for (auto it = 0; it != 10; ++it) {
    Mat<Nx, 1> x = Random<Nx, 1>(), xhat = Random<Nx, 1>();
    smooth::feedback::EKF<Mat<Nx, 1>> ekf;
    Mat<Nx, Nx> P = RandomDiag<Nx>();
    ekf.reset(xhat, P);
    Mat<Ny, Nx> H = Random<Ny, Nx>();
    Mat<Ny, 1> h  = Random<Ny, 1>();
    Mat<Ny, Ny> R = RandomDiag<Ny>();
    ekf.update([&H, &h]<typename T>(const Eigen::Matrix<T, Nx, 1> & xvar) -> Eigen::Matrix<T, Ny, 1> { return mul(H, xvar) + h; },
               mul(H, x) + h, R);
    Mat<Ny, Ny> S = mul(mul(H, P), tr(H)) + R;
    Mat<Nx, Ny> K = mul(mul(P, tr(H)), inverse(S));
    Mat<Nx, 1> x_new  = xhat + mul(K, mul(H, x) + h - (mul(H, xhat) + h));
    Mat<Nx, Nx> P_new = mul(Mat<Nx, Nx>::Identity() - mul(K, H), P);
    CHECK(isApprox(x_new, ekf.estimate(), 1e-6));
    CHECK(isApprox(P_new, ekf.covariance(), 1e-6));
  }
  return 0;
}

template<int Nx> int test_predict_linear()
{
  for (auto it = 0; it != 3; ++it) {
    Mat<Nx, 1> xhat = Random<Nx, 1>();
    smooth::feedback::EKF<Mat<Nx, 1>, smooth::diff::Type::Numerical, boost::numeric::odeint::runge_kutta4> ekf;
    Mat<Nx, Nx> P = RandomDiag<Nx>();
    ekf.reset(xhat, P);
    Mat<Nx, Nx> A = Random<Nx, Nx>();
    Mat<Nx, Nx> Q;  // zero: the linear-system solution is complicated for Q != 0 (tests/test_ekf.cpp:127)
    double tau = 0.7;
    ekf.predict([&A]<typename T>(double, const Eigen::Matrix<T, Nx, 1> & xvar) -> Eigen::Matrix<T, Nx, 1> { return mul(A, xvar); }, Q, tau, 1e-3);
    Mat<Nx, Nx> At = A;
    for (int k = 0; k < Nx * Nx; ++k) { At.data()[k] *= tau; }
    Mat<Nx, Nx> F = expm(At);
    CHECK(isApprox(mul(F, xhat), ekf.estimate(), 1e-3));
    CHECK(isApprox(mul(mul(F, P), tr(F)), ekf.covariance(), 1e-3));
  }
  return 0;
}

int main()
{
  if (test_update_linear<3, 3>() || test_update_linear<10, 3>() || test_update_linear<3, 10>()) { return 1; }
  if (test_predict_linear<3>() || test_predict_linear<6>() || test_predict_linear<9>()) { return 1; }
  {  // PredictTimeCut (tests/test_ekf.cpp:155-180): Euler, tau = 0.7 with dt = 0.5 -> two sub-steps 0.5 + 0.2
    Mat<2, 1> xhat = Random<2, 1>();
    Mat<2, 2> P    = RandomDiag<2>();
    smooth::feedback::EKF<Mat<2, 1>> ekf;
    ekf.reset(xhat, P);
    Mat<2, 1> b = Random<2, 1>();
    Mat<2, 2> Q = RandomDiag<2>();
    double tau  = 0.7;
    ekf.predict([&b]<typename T>(T, const Eigen::Matrix<T, 2, 1> &) -> Eigen::Matrix<T, 2, 1> { return b; }, Q, tau, 0.5);
    CHECK(isApprox(ekf.estimate(), xhat + tau * b, 1e-12));
    // constant dynamics: A = 0, so the covariance grows by symU(Q) tau exactly (two Euler sub-steps)
    Mat<2, 2> Pexp = P;
    for (int k = 0; k < 4; ++k) { Pexp.data()[k] += tau * Q.data()[k]; }
    CHECK(isApprox(Pexp, ekf.covariance(), 1e-12));
    auto copy = ekf;  // a copy is an independent filter with the same estimate
    CHECK(isApprox(copy.estimate(), ekf.estimate(), 1e-15));
  }
  std::printf("ekf overlay ok\n");
  return 0;
}
