// sf_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code).  See sf_oracle.h.
//
// Every function cites the reference lines it restates (paths relative to /root/reference).
// Written from the algorithm, with plain loops over column-major arrays -- no Eigen.

#include "sf_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

static long long g_dx_zero_checks = 0;  // see check_stopping (instrumentation for tests)

namespace {

constexpr double kInf = std::numeric_limits<double>::infinity();

// QPSolutionStatus, qp.hpp:82-92 (order is part of the contract)
enum Status : int32_t {
  Optimal          = 0,
  PolishFailed     = 1,
  PrimalInfeasible = 2,
  DualInfeasible   = 3,
  MaxIterations    = 4,
  MaxTime          = 5,
  Unknown          = 6,
  Unset            = -1
};

inline double norm_inf(const double* v, int n)
{
  double r = 0;
  for (int i = 0; i < n; ++i) r = std::max(r, std::fabs(v[i]));
  return r;
}

// ---------------------------------------------------------------------------------------------
// Eigen 3.4.0 LDLT<Matrix, Upper> restated (third-party, absent from /root/reference; used at
// qp_solver.hpp:187-194,259,428,462 and ekf.hpp:134).
//
// Published algorithm (Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked run on the
// transpose view for Upper): unblocked left-looking LDL^T with symmetric pivoting on the largest
// |diagonal entry| of the *stored* (not yet updated) trailing diagonal, first maximum wins.
// W holds the lower triangle (W(i,j), i>=j) of the symmetric matrix, column-major, leading dim k.
// ---------------------------------------------------------------------------------------------
struct Ldlt
{
  int k = 0;
  std::vector<double> W;     // factor: unit-lower L strictly below diagonal, D on the diagonal
  std::vector<int> tr;       // transpositions
  std::vector<double> temp;  // workspace
  bool ok = false;

  double& at(int i, int j) { return W[size_t(i) + size_t(k) * j]; }
  double at(int i, int j) const { return W[size_t(i) + size_t(k) * j]; }

  void resize(int size)
  {
    k = size;
    W.assign(size_t(size) * size, 0.0);
    tr.assign(size, 0);
    temp.assign(size, 0.0);
  }

  // in-place factorisation of the lower triangle currently stored in W
  void compute()
  {
    const int size = k;
    bool found_zero_pivot = false;
    bool ret = true;
    if (size <= 1) {
      for (int i = 0; i < size; ++i) tr[i] = i;
      ok = true;
      return;
    }
    for (int kk = 0; kk < size; ++kk) {
      // largest |diagonal| in the trailing corner, first maximum
      int big = kk;
      double bigv = std::fabs(at(kk, kk));
      for (int j = kk + 1; j < size; ++j) {
        const double v = std::fabs(at(j, j));
        if (v > bigv) { bigv = v; big = j; }
      }
      tr[kk] = big;
      if (kk != big) {
        // symmetric transposition touching only the lower triangle
        const int s = size - big - 1;
        for (int j = 0; j < kk; ++j) std::swap(at(kk, j), at(big, j));
        for (int i = 0; i < s; ++i) std::swap(at(big + 1 + i, kk), at(big + 1 + i, big));
        std::swap(at(kk, kk), at(big, big));
        for (int i = kk + 1; i < big; ++i) std::swap(at(i, kk), at(big, i));
      }
      const int rs = size - kk - 1;
      if (kk > 0) {
        for (int j = 0; j < kk; ++j) temp[j] = at(j, j) * at(kk, j);
        double acc = 0;
        for (int j = 0; j < kk; ++j) acc += at(kk, j) * temp[j];
        at(kk, kk) -= acc;
        for (int i = kk + 1; i < size; ++i) {
          double a = 0;
          for (int j = 0; j < kk; ++j) a += at(i, j) * temp[j];
          at(i, kk) -= a;
        }
      }
      const double akk = at(kk, kk);
      const bool pivot_is_valid = std::fabs(akk) > 0.0;
      if (kk == 0 && !pivot_is_valid) {
        // entire diagonal is zero: success iff the matrix is zero
        for (int j = 0; j < size; ++j) {
          tr[j] = j;
          for (int i = j + 1; i < size; ++i) ret = ret && (at(i, j) == 0.0);
        }
        ok = ret;
        return;
      }
      if (rs > 0 && pivot_is_valid) {
        for (int i = kk + 1; i < size; ++i) at(i, kk) /= akk;
      } else if (rs > 0) {
        for (int i = kk + 1; i < size; ++i) ret = ret && (at(i, kk) == 0.0);
      }
      if (found_zero_pivot && pivot_is_valid) {
        ret = false;
      } else if (!pivot_is_valid) {
        found_zero_pivot = true;
      }
    }
    ok = ret;
  }

  // b <- A^{-1} b  (LDLT::solveInPlace: P, L^{-1}, pseudo-inverse of D, L^{-T}, P^T)
  void solve_in_place(double* b) const
  {
    const int size = k;
    for (int i = 0; i < size; ++i) std::swap(b[i], b[tr[i]]);
    for (int j = 0; j < size; ++j) {
      const double bj = b[j];
      if (bj != 0.0) {
        for (int i = j + 1; i < size; ++i) b[i] -= at(i, j) * bj;
      }
    }
    const double tol = std::numeric_limits<double>::min();
    for (int i = 0; i < size; ++i) {
      const double d = at(i, i);
      if (std::fabs(d) > tol) b[i] /= d; else b[i] = 0.0;
    }
    for (int j = size - 1; j >= 0; --j) {
      double a = b[j];
      for (int i = j + 1; i < size; ++i) a -= at(i, j) * b[i];
      b[j] = a;
    }
    for (int i = size - 1; i >= 0; --i) std::swap(b[i], b[tr[i]]);
  }
};

// ---------------------------------------------------------------------------------------------
// QPSolver<QuadraticProgram<-1,-1,double>> restated, dense branches only.
// ---------------------------------------------------------------------------------------------
struct QpOracle
{
  int n = 0, m = 0;
  sfo_qp_params prm{};

  // analyze(): qp_solver.hpp:297-338
  double c = 1;
  std::vector<double> sx, sy, sx_inc, sy_inc;
  std::vector<double> x, y, z, z_next, rho, p;
  std::vector<double> x_us, dx_us, y_us, dy_us, z_us, Px, Aty, Ax;
  Ldlt ldlt;

  void analyze(int n_, int m_)
  {
    n = n_; m = m_;
    const int k = n + m;
    x.assign(n, 0); y.assign(m, 0);
    c = 1;
    sx.assign(n, 1); sy.assign(m, 1);
    sx_inc.assign(n, 0); sy_inc.assign(m, 0);
    z.assign(m, 0); z_next.assign(m, 0); rho.assign(m, 0); p.assign(k, 0);
    x_us.assign(n, 0); dx_us.assign(n, 0); y_us.assign(m, 0); dy_us.assign(m, 0); z_us.assign(m, 0);
    Px.assign(n, 0); Aty.assign(n, 0); Ax.assign(m, 0);
    ldlt.resize(k);
  }

  // scale(): qp_solver.hpp:673-730.  P[i + n*j], A[i + m*j].
  void scale(const double* P, const double* q, const double* A)
  {
    std::fill(sx.begin(), sx.end(), 1.0);
    std::fill(sy.begin(), sy.end(), 1.0);
    std::fill(sx_inc.begin(), sx_inc.end(), 0.0);
    // :681-685 column-wise max |P_ij| over ALL entries
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) sx_inc[j] = std::max(sx_inc[j], std::fabs(P[i + size_t(n) * j]));
    for (int j = 0; j < n; ++j)
      if (sx_inc[j] == 0) sx_inc[j] = 1;
    // :693
    double mean = 0;
    for (int j = 0; j < n; ++j) mean += sx_inc[j];
    mean /= double(n);
    c = 1.0 / std::max({1e-6, mean, norm_inf(q, n)});

    int iter = 0;
    double dev;
    do {
      std::fill(sx_inc.begin(), sx_inc.end(), 0.0);
      std::fill(sy_inc.begin(), sy_inc.end(), 0.0);
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < n; ++i)
          sx_inc[j] = std::max(sx_inc[j], std::fabs(c * sx[i] * sx[j] * P[i + size_t(n) * j]));  // :704-707
      for (int j = 0; j < n; ++j)
        for (int i = 0; i < m; ++i) {
          const double Aij = std::fabs(sy[i] * sx[j] * A[i + size_t(m) * j]);  // :712
          sx_inc[j] = std::max(sx_inc[j], Aij);
          sy_inc[i] = std::max(sy_inc[i], Aij);
        }
      for (int j = 0; j < n; ++j)
        if (sx_inc[j] == 0) sx_inc[j] = 1;
      for (int i = 0; i < m; ++i)
        if (sy_inc[i] == 0) sy_inc[i] = 1;
      // :726-727  s <- s * sqrt(1 / max(inc, 1e-8))
      for (int j = 0; j < n; ++j) sx[j] = std::sqrt(1.0 / std::max(sx_inc[j], 1e-8)) * sx[j];
      for (int i = 0; i < m; ++i) sy[i] = std::sqrt(1.0 / std::max(sy_inc[i], 1e-8)) * sy[i];
      double dx = 0, dy = 0;
      for (int j = 0; j < n; ++j) dx = std::max(dx, std::fabs(sx_inc[j] - 1.0));
      for (int i = 0; i < m; ++i) dy = std::max(dy, std::fabs(sy_inc[i] - 1.0));
      dev = std::max(dx, dy);
    } while (iter++ < 10 && dev > 0.1);  // :728-729
  }

  // check_stopping(): qp_solver.hpp:574-644
  Status check_stopping(const double* P, const double* q, const double* A, const double* l, const double* u)
  {
    const double eps_abs = prm.eps_abs, eps_rel = prm.eps_rel;  // floats promoted
    const double eps_pinf = prm.eps_primal_inf, eps_dinf = prm.eps_dual_inf;

    // OPTIMALITY :584-594
    for (int i = 0; i < m; ++i) {
      double a = 0;
      for (int j = 0; j < n; ++j) a += A[i + size_t(m) * j] * x_us[j];
      Ax[i] = a;
    }
    const double Ax_norm = norm_inf(Ax.data(), m);
    for (int i = 0; i < m; ++i) Ax[i] -= z_us[i];
    if (norm_inf(Ax.data(), m) <= eps_abs + eps_rel * std::max(Ax_norm, norm_inf(z_us.data(), m))) {
      for (int i = 0; i < n; ++i) {
        double a = 0;
        for (int j = 0; j < n; ++j) a += P[i + size_t(n) * j] * x_us[j];
        Px[i] = a;
      }
      for (int j = 0; j < n; ++j) {
        double a = 0;
        for (int i = 0; i < m; ++i) a += A[i + size_t(m) * j] * y_us[i];
        Aty[j] = a;
      }
      const double dual_scale = std::max({norm_inf(Px.data(), n), norm_inf(q, n), norm_inf(Aty.data(), n)});
      for (int i = 0; i < n; ++i) Px[i] += q[i] + Aty[i];
      if (norm_inf(Px.data(), n) <= eps_abs + eps_rel * dual_scale) return Optimal;
    }

    // PRIMAL INFEASIBILITY :598-621
    for (int j = 0; j < n; ++j) {
      double a = 0;
      for (int i = 0; i < m; ++i) a += A[i + size_t(m) * j] * dy_us[i];
      Aty[j] = a;
    }
    const double Edy_norm = norm_inf(dy_us.data(), m);
    double s = 0;
    for (int i = 0; i < m; ++i) {
      if (u[i] != kInf) {
        s += u[i] * std::max(0.0, dy_us[i]);
      } else if (dy_us[i] > eps_pinf * Edy_norm) {
        s = kInf;
        break;
      }
      if (l[i] != -kInf) {
        s += l[i] * std::min(0.0, dy_us[i]);
      } else if (dy_us[i] < -eps_pinf * Edy_norm) {
        s = kInf;
        break;
      }
    }
    if (std::max(norm_inf(Aty.data(), n), s) < eps_pinf * Edy_norm) return PrimalInfeasible;

    // DUAL INFEASIBILITY :625-641
    for (int i = 0; i < m; ++i) {
      double a = 0;
      for (int j = 0; j < n; ++j) a += A[i + size_t(m) * j] * dx_us[j];
      Ax[i] = a;
    }
    const double dx_norm = norm_inf(dx_us.data(), n);
    // instrumentation only (never changes a result): how often a stop check sees an EXACTLY stationary primal iterate.
    // With dx == 0 every comparison of qp_solver.hpp:629-639 reads 0 <= 0; the CUDA engine guards its dual-infeasibility
    // certificate with dx != 0 (deviations) and tests/ use this counter to show that the reference algorithm
    // never meets that case on any parity workload, i.e. that the guard cannot change a status the reference returns.
    if (dx_norm == 0.0) __atomic_fetch_add(&g_dx_zero_checks, 1, __ATOMIC_RELAXED);
    for (int i = 0; i < n; ++i) {
      double a = 0;
      for (int j = 0; j < n; ++j) a += P[i + size_t(n) * j] * dx_us[j];
      Px[i] = a;
    }
    double qdx = 0;
    for (int j = 0; j < n; ++j) qdx += q[j] * dx_us[j];
    bool dual_infeasible = (norm_inf(Px.data(), n) <= eps_dinf * dx_norm) && (qdx <= eps_dinf * dx_norm);
    for (int i = 0; i < m && dual_infeasible; ++i) {
      if (u[i] == kInf) {
        dual_infeasible &= (Ax[i] >= -eps_dinf * dx_norm);
      } else if (l[i] == -kInf) {
        dual_infeasible &= (Ax[i] <= eps_dinf * dx_norm);
      } else {
        dual_infeasible &= std::fabs(Ax[i]) < eps_dinf * dx_norm;
      }
    }
    if (dual_infeasible) return DualInfeasible;
    return Unset;
  }

  // active set as polish_qp defines it, :113-123 (on the *scaled* dual)
  void active_set(const double* l, const double* u, int8_t* act) const
  {
    const double eps = std::numeric_limits<double>::epsilon();
    for (int i = 0; i < m; ++i) {
      int8_t a = 0;
      if (y[i] < -100 * eps && l[i] != -kInf) a = -1;
      if (y[i] > 100 * eps && u[i] != kInf) a = 1;
      act[i] = a;
    }
  }

  // detail::polish_qp, qp_solver.hpp:92-204 (dense branch).  Works on scaled x, y in place.
  bool polish(const double* P, const double* q, const double* A, const double* l, const double* u)
  {
    const double eps = std::numeric_limits<double>::epsilon();
    std::vector<int> LU_idx;
    int nl = 0, nu = 0;
    for (int i = 0; i < m; ++i) {
      if (y[i] < -100 * eps && l[i] != -kInf) nl++;
      if (y[i] > 100 * eps && u[i] != kInf) nu++;
    }
    LU_idx.assign(nl + nu, 0);
    for (int i = 0, lc = 0, uc = 0; i < m; ++i) {
      if (y[i] < -100 * eps && l[i] != -kInf) LU_idx[lc++] = i;
      if (y[i] > 100 * eps && u[i] != kInf) LU_idx[nl + uc++] = i;
    }
    const int na = nl + nu, K = n + na;
    const double delta = prm.delta;

    // H (upper triangle significant) :160-166 ; Hp = H + delta*I_n (+) -delta*I_na :175-176
    std::vector<double> H(size_t(K) * K, 0.0);
    auto Hat = [&](int r, int cc) -> double& { return H[size_t(r) + size_t(K) * cc]; };
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) Hat(i, j) = c * sx[i] * P[i + size_t(n) * j] * sx[j];
    for (int a = 0; a < na; ++a) {
      const int r = LU_idx[a];
      for (int j = 0; j < n; ++j) Hat(j, n + a) = sy[r] * A[r + size_t(m) * j] * sx[j];
    }
    Ldlt f;
    f.resize(K);
    // lower-triangle view of the transpose == upper triangle of Hp
    for (int j = 0; j < K; ++j)
      for (int i = j; i < K; ++i) f.at(i, j) = Hat(j, i);
    for (int i = 0; i < n; ++i) f.at(i, i) += delta;
    for (int i = 0; i < na; ++i) f.at(n + i, n + i) -= delta;

    std::vector<double> h(K), t(K, 0.0), r(K);
    for (int i = 0; i < n; ++i) h[i] = -c * (sx[i] * q[i]);  // :180
    for (int a = 0; a < nl; ++a) h[n + a] = sy[LU_idx[a]] * l[LU_idx[a]];
    for (int a = 0; a < nu; ++a) h[n + nl + a] = sy[LU_idx[nl + a]] * u[LU_idx[nl + a]];

    f.compute();
    if (!f.ok) return false;  // :190

    for (uint32_t it = 0; it != prm.polish_iter; ++it) {  // :193-195
      // r = h - selfadjointView<Upper>(H) * t
      for (int i = 0; i < K; ++i) {
        double a = 0;
        for (int j = 0; j < K; ++j) a += (i <= j ? Hat(i, j) : Hat(j, i)) * t[j];
        r[i] = h[i] - a;
      }
      f.solve_in_place(r.data());
      for (int i = 0; i < K; ++i) t[i] += r[i];
    }
    for (int i = 0; i < n; ++i) x[i] = t[i];  // :199-201
    for (int a = 0; a < na; ++a) y[LU_idx[a]] = t[n + a];
    return true;
  }

  // solve(): qp_solver.hpp:343-568
  void solve(const double* P, const double* q, const double* A, const double* l, const double* u,
             const double* warm_x, const double* warm_y, double* out_x, double* out_y, double* out_obj,
             int32_t* out_status, uint32_t* out_iter, int8_t* out_active)
  {
    if (prm.scaling) scale(P, q, A);  // :347

    const double rho_bar = double(prm.rho), alpha = double(prm.alpha), alpha_comp = 1.0 - alpha,
                 sigma = double(prm.sigma);  // :353-356
    Status ret = Unset;

    for (int i = 0; i < m; ++i) {  // :361-374
      if (l[i] == kInf || u[i] == -kInf || u[i] - l[i] < 0.0) ret = PrimalInfeasible;
      if (l[i] == -kInf && u[i] == kInf) {
        rho[i] = 1e-6;
      } else if (sy[i] * std::fabs(l[i] - u[i]) < 1e-5) {
        rho[i] = 1e3 * rho_bar;
      } else {
        rho[i] = rho_bar;
      }
    }

    // KKT matrix, upper triangle of H :399-404, handed to LDLT<.,Upper> as the lower triangle of H^T
    std::fill(ldlt.W.begin(), ldlt.W.end(), 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i) ldlt.at(i, j) = c * sx[j] * P[j + size_t(n) * i] * sx[i];  // H(j,i), j<=i
    for (int i = 0; i < n; ++i) ldlt.at(i, i) += sigma;
    for (int r = 0; r < m; ++r)
      for (int j = 0; j < n; ++j) ldlt.at(n + r, j) = sy[r] * A[r + size_t(m) * j] * sx[j];  // H(j, n+r)
    for (int r = 0; r < m; ++r) ldlt.at(n + r, n + r) = 1.0 / (-rho[r]);
    ldlt.compute();                 // :428
    if (!ldlt.ok) ret = Unknown;    // :433

    if (warm_x && warm_y) {  // :436-445
      for (int i = 0; i < n; ++i) x[i] = (1.0 / sx[i]) * warm_x[i];
      for (int i = 0; i < m; ++i) y[i] = c * ((1.0 / sy[i]) * warm_y[i]);
      for (int i = 0; i < m; ++i) {
        double a = 0;
        for (int j = 0; j < n; ++j) a += (sy[i] * A[i + size_t(m) * j]) * warm_x[j];
        z[i] = a;
      }
    } else {
      std::fill(x.begin(), x.end(), 0.0);
      std::fill(y.begin(), y.end(), 0.0);
      std::fill(z.begin(), z.end(), 0.0);
    }

    uint32_t iter = 0;
    for (; (!prm.has_max_iter || iter != prm.max_iter) && ret == Unset; ++iter) {  // :449
      for (int i = 0; i < n; ++i) p[i] = sigma * x[i] - (c * sx[i]) * q[i];       // :450
      for (int i = 0; i < m; ++i) p[n + i] = z[i] - (1.0 / rho[i]) * y[i];        // :451
      ldlt.solve_in_place(p.data());                                               // :462

      const bool chk = (iter % prm.stop_check_iter == 1);
      if (chk) {  // :465-468
        dx_us = x;
        dy_us = y;
      }
      for (int i = 0; i < n; ++i) x[i] = alpha * p[i] + alpha_comp * x[i];  // :470
      for (int i = 0; i < m; ++i) {                                          // :471-474
        const double rinv = 1.0 / rho[i];
        double v = alpha * (rinv * p[n + i]) + alpha_comp * (rinv * y[i]) + z[i];
        v = std::max(v, sy[i] * l[i]);
        v = std::min(v, sy[i] * u[i]);
        z_next[i] = v;
      }
      for (int i = 0; i < m; ++i)  // :475-476
        y[i] = alpha_comp * y[i] + alpha * p[n + i] + rho[i] * z[i] - rho[i] * z_next[i];
      std::swap(z, z_next);  // :477

      if (chk) {  // :479-509
        for (int i = 0; i < n; ++i) x_us[i] = sx[i] * x[i];
        for (int i = 0; i < m; ++i) y_us[i] = sy[i] * y[i] / c;
        for (int i = 0; i < m; ++i) z_us[i] = (1.0 / sy[i]) * z[i];
        for (int i = 0; i < n; ++i) dx_us[i] = sx[i] * (x[i] - dx_us[i]);
        for (int i = 0; i < m; ++i) dy_us[i] = sy[i] * (y[i] - dy_us[i]) / c;
        ret = check_stopping(P, q, A, l, u);
      }
    }

    if (out_active) active_set(l, u, out_active);

    if (ret == Optimal && prm.polish) {  // :515-539 ; a failed polish is silently ignored (:537 vs :544)
      (void)polish(P, q, A, l, u);
    }

    // :544-548
    *out_status = (ret == Unset) ? int32_t(MaxIterations) : int32_t(ret);
    for (int i = 0; i < n; ++i) out_x[i] = sx[i] * x[i];
    for (int i = 0; i < m; ++i) out_y[i] = sy[i] * y[i] / c;
    double obj = 0;
    for (int i = 0; i < n; ++i) {
      double a = 0;
      for (int j = 0; j < n; ++j) a += 0.5 * P[i + size_t(n) * j] * out_x[j];
      obj += out_x[i] * (a + q[i]);
    }
    *out_obj = obj;
    *out_iter = iter;
    // keep solver state like the reference does: sol_ holds the unscaled solution after solve()
  }
};

// ---------------------------------------------------------------------------------------------
// EKF dense algebra, ekf.hpp:79-139
// ---------------------------------------------------------------------------------------------

// dcov = selfadjointView<Upper>( A*cov + cov*A^T + Q )   ekf.hpp:88
void cov_ode(int d, const double* A, const double* Q, const double* cov, double* dcov)
{
  for (int j = 0; j < d; ++j) {
    for (int i = 0; i <= j; ++i) {
      double a = 0;
      for (int k = 0; k < d; ++k) a += A[i + d * k] * cov[k + d * j];
      for (int k = 0; k < d; ++k) a += cov[i + d * k] * A[j + d * k];
      a += Q[i + d * j];
      dcov[i + d * j] = a;
      dcov[j + d * i] = a;
    }
  }
}

// one stepper step on the covariance (Boost.odeint euler / runge_kutta4, vector_space_algebra)
void cov_step(int d, int stepper, const double* A, const double* Q, double* P, double dt, double* w)
{
  const int dd = d * d;
  double* k1 = w;
  double* k2 = w + dd;
  double* k3 = w + 2 * dd;
  double* k4 = w + 3 * dd;
  double* xt = w + 4 * dd;
  if (stepper == 0) {
    cov_ode(d, A, Q, P, k1);
    for (int i = 0; i < dd; ++i) P[i] = P[i] + dt * k1[i];
    return;
  }
  cov_ode(d, A, Q, P, k1);
  for (int i = 0; i < dd; ++i) xt[i] = P[i] + (dt * 0.5) * k1[i];
  cov_ode(d, A, Q, xt, k2);
  for (int i = 0; i < dd; ++i) xt[i] = P[i] + (dt * 0.5) * k2[i];
  cov_ode(d, A, Q, xt, k3);
  for (int i = 0; i < dd; ++i) xt[i] = P[i] + dt * k3[i];
  cov_ode(d, A, Q, xt, k4);
  for (int i = 0; i < dd; ++i)
    P[i] = P[i] + (dt / 6.0) * k1[i] + (dt / 3.0) * k2[i] + (dt / 3.0) * k3[i] + (dt / 6.0) * k4[i];
}

void ekf_predict_one(int d, int stepper, const double* P0, const double* A, const double* Q, double tau,
                     double dt, double* P, double* w)
{
  const int dd = d * d;
  std::memcpy(P, P0, sizeof(double) * dd);
  double t = 0;
  const double dt_v = (dt > 0) ? dt : 2 * tau;  // ekf.hpp:92
  while (t + dt_v < tau) {                      // :93-98
    cov_step(d, stepper, A, Q, P, dt_v, w);
    t += dt_v;
  }
  cov_step(d, stepper, A, Q, P, tau - t, w);    // :101
}

void ekf_update_one(int d, int ny, const double* P, const double* H, const double* R, const double* innov,
                    double* delta, double* outP, Ldlt& f, std::vector<double>& w)
{
  // S = triu(H * symU(P) * H^T + R)   ekf.hpp:129-130
  w.assign(size_t(ny) * d * 3 + size_t(d) * d, 0.0);
  double* HPs = w.data();                 // H * symU(P)   ny x d
  double* HP  = HPs + size_t(ny) * d;     // H * P (full)  ny x d
  double* Kt  = HP + size_t(ny) * d;      // S^{-1} H P    ny x d  (= K^T)
  double* IKH = Kt + size_t(ny) * d;      // I - K H       d x d
  for (int i = 0; i < ny; ++i)
    for (int j = 0; j < d; ++j) {
      double a = 0, b = 0;
      for (int k = 0; k < d; ++k) {
        const double ps = (k <= j) ? P[k + d * j] : P[j + d * k];
        a += H[i + ny * k] * ps;
        b += H[i + ny * k] * P[k + d * j];
      }
      HPs[i + ny * j] = a;
      HP[i + ny * j]  = b;
    }
  f.resize(ny);
  for (int j = 0; j < ny; ++j)
    for (int i = j; i < ny; ++i) {
      // upper entry S(j,i), j<=i
      double a = 0;
      for (int k = 0; k < d; ++k) a += HPs[j + ny * k] * H[i + ny * k];
      f.at(i, j) = a + R[j + ny * i];
    }
  f.compute();
  // K^T = S^{-1} (H P)   :133-134
  std::vector<double> col(ny);
  for (int j = 0; j < d; ++j) {
    for (int i = 0; i < ny; ++i) col[i] = HP[i + ny * j];
    f.solve_in_place(col.data());
    for (int i = 0; i < ny; ++i) Kt[i + ny * j] = col[i];
  }
  // delta = K * innov   :137
  for (int i = 0; i < d; ++i) {
    double a = 0;
    for (int k = 0; k < ny; ++k) a += Kt[k + ny * i] * innov[k];
    delta[i] = a;
  }
  // P = symU((I - K H) P)   :138
  for (int i = 0; i < d; ++i)
    for (int j = 0; j < d; ++j) {
      double a = 0;
      for (int k = 0; k < ny; ++k) a += Kt[k + ny * i] * H[k + ny * j];
      IKH[i + d * j] = (i == j ? 1.0 : 0.0) - a;
    }
  for (int j = 0; j < d; ++j)
    for (int i = 0; i <= j; ++i) {
      double a = 0;
      for (int k = 0; k < d; ++k) a += IKH[i + d * k] * P[k + d * j];
      outP[i + d * j] = a;
      outP[j + d * i] = a;
    }
}

}  // namespace

extern "C" {

long long sfo_debug_dx_zero_checks(int reset)
{
  const long long v = __atomic_load_n(&g_dx_zero_checks, __ATOMIC_RELAXED);
  if (reset) __atomic_store_n(&g_dx_zero_checks, 0, __ATOMIC_RELAXED);
  return v;
}

void sfo_qp_params_default(sfo_qp_params* p)
{
  // qp_solver.hpp:29-68
  p->alpha = 1.6f;
  p->rho = 0.1f;
  p->sigma = 1e-6f;
  p->scaling = 1;
  p->eps_abs = 1e-3f;
  p->eps_rel = 1e-3f;
  p->eps_primal_inf = 1e-4f;
  p->eps_dual_inf = 1e-4f;
  p->has_max_iter = 0;
  p->max_iter = 0;
  p->stop_check_iter = 25;
  p->polish = 1;
  p->polish_iter = 5;
  p->delta = 1e-6f;
}

int sfo_qp_solve_dense_batch_f64(const sfo_qp_params* prm, int64_t batch, int n, int m, const double* P,
                                 const double* q, const double* A, const double* l, const double* u,
                                 const double* warm_x, const double* warm_y, double* out_x, double* out_y,
                                 double* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active,
                                 int nthreads)
{
  if (!prm || n <= 0 || m < 0 || prm->stop_check_iter == 0) return 1;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    QpOracle s;  // one reusable workspace per thread == the QPSolver-object usage pattern
    s.prm = *prm;
    s.analyze(n, m);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (int64_t b = 0; b < batch; ++b) {
      // reference solve_qp builds a fresh solver per call: scaling state restarts at c=1, sx=sy=1
      s.c = 1;
      std::fill(s.sx.begin(), s.sx.end(), 1.0);
      std::fill(s.sy.begin(), s.sy.end(), 1.0);
      s.solve(P + b * size_t(n) * n, q + b * size_t(n), A + b * size_t(m) * n, l + b * size_t(m),
              u + b * size_t(m), warm_x ? warm_x + b * size_t(n) : nullptr,
              warm_y ? warm_y + b * size_t(m) : nullptr, out_x + b * size_t(n), out_y + b * size_t(m),
              out_obj + b, out_status + b, out_iter + b, out_active ? out_active + b * size_t(m) : nullptr);
    }
  }
  return 0;
}

int sfo_qp_scale_f64(int n, int m, const double* P, const double* q, const double* A, double* c, double* sx,
                     double* sy)
{
  QpOracle s;
  sfo_qp_params_default(&s.prm);
  s.analyze(n, m);
  s.scale(P, q, A);
  *c = s.c;
  std::copy(s.sx.begin(), s.sx.end(), sx);
  std::copy(s.sy.begin(), s.sy.end(), sy);
  return 0;
}

int sfo_ekf_predict_batch_f64(int64_t batch, int d, int stepper, const double* P, const double* A,
                              const double* Q, double tau, double dt, double* out_P, int nthreads)
{
  if (d <= 0 || (stepper != 0 && stepper != 1)) return 1;
  const size_t dd = size_t(d) * d;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    std::vector<double> w(5 * dd);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t b = 0; b < batch; ++b)
      ekf_predict_one(d, stepper, P + b * dd, A + b * dd, Q + b * dd, tau, dt, out_P + b * dd, w.data());
  }
  return 0;
}

int sfo_ekf_update_batch_f64(int64_t batch, int d, int ny, const double* P, const double* H, const double* R,
                             const double* innov, double* out_delta, double* out_P, int nthreads)
{
  if (d <= 0 || ny <= 0) return 1;
  const size_t dd = size_t(d) * d;
#ifdef _OPENMP
  if (nthreads < 1) nthreads = 1;
#pragma omp parallel num_threads(nthreads)
#endif
  {
    Ldlt f;
    std::vector<double> w;
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int64_t b = 0; b < batch; ++b)
      ekf_update_one(d, ny, P + b * dd, H + b * size_t(ny) * d, R + b * size_t(ny) * ny, innov + b * size_t(ny),
                     out_delta + b * size_t(d), out_P + b * dd, f, w);
  }
  return 0;
}

}  // extern "C"
