// qp_dense_skinny.cuh -- dense operator-splitting QP solver for TALL-SKINNY problems (n <= 4 variables, m <= 256 rows,
// polish off): one WARP per instance, the whole working set in REGISTERS.
//
// This is the shape ASIFilter::operator() produces (asif.hpp:97 -> solve_qp, dense, n = nu + 1, m = K nh + nu_ineq + 1:
// n = 3, m = 203 for examples/mpc_asif_vehicle.cpp:96-129, n = 4, m = 301 in tests/test_asif.cpp:103-129 -- the latter
// exceeds 256 rows and takes the generic kernel) with the reference's own ASIF setting polish = false
// (mpc_asif_vehicle.cpp:127, asif_doubleintegrator.cpp:54).  Same algorithm as qp_dense_group.cuh
// (qp_solver.hpp:343-568: scale :673-730, rho classes :361-374, reduced KKT system, ADMM loop :449-510, check_stopping
// :574-644, unscale + objective :544-548); what changes is where the data lives:
//   * lane L owns rows L, L + 32, ... (R = 2, 4 or 8 rows per lane): the row of Abar, its bounds, rho, z, y stay in
//     registers for the whole solve; every n-vector and the n x n inverse are replicated on all lanes;
//   * Abar^T w and the norms of check_stopping are warp shuffles (xor butterfly: every lane ends with the same bits, so
//     all control flow is warp-uniform); an ADMM iteration touches neither shared nor global memory.
// The generic kernel spends an iteration of this shape on shared-memory GEMV passes and block barriers for n = 3 columns.

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "qp_dense_group.cuh"

namespace sfb {

constexpr int kSkinnyMaxN = 4;
constexpr int kSkinnyMaxM = 256;
constexpr int kSkinnyWarps = 4;  // warps (= instances in flight) per CTA

template <typename T, int N, int R> struct QpSkinny
{
  int lane, m;
  bool valid[R];
  T A[R][N];                       // raw rows, then Abar
  T l[R], u[R], sy[R], rho[R], rinv[R], z[R], y[R], yold[R], lo[R], hi[R];
  T P[N][N], q[N], qb[N], x[N], xold[N], sx[N], Minv[N][N];
  T c;
  bool dinf_guard = true;  // see QpArgs::dinf_guard

  __device__ __forceinline__ T wmax(T v) const { return warp_max(v); }
  __device__ __forceinline__ T wsum(T v) const { return warp_sum(v); }

  __device__ void load(const QpArgs<T>& a, long long b)
  {
    lane = threadIdx.x & 31;
    m = a.m;
    const T inf = Num<T>::inf();
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int i = lane + 32 * k;
      valid[k] = i < m;
#pragma unroll
      for (int j = 0; j < N; ++j) A[k][j] = valid[k] ? __ldg(a.A + b * (long long)m * N + i + (long long)m * j) : T(0);
      l[k] = valid[k] ? __ldg(a.l + b * (long long)m + i) : -inf;  // padding rows: free, zero row
      u[k] = valid[k] ? __ldg(a.u + b * (long long)m + i) : inf;
    }
#pragma unroll
    for (int j = 0; j < N; ++j) {
      q[j] = __ldg(a.q + b * N + j);
#pragma unroll
      for (int i = 0; i < N; ++i) P[i][j] = __ldg(a.P + b * N * N + i + N * j);
    }
  }

  // QPSolver::scale, qp_solver.hpp:673-730 (same products, exact maxima: bit-identical to the CPU restatement)
  __device__ void scale()
  {
#pragma unroll
    for (int j = 0; j < N; ++j) sx[j] = T(1);
#pragma unroll
    for (int k = 0; k < R; ++k) sy[k] = T(1);
    T mean = T(0), qn = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      T g = T(0);
#pragma unroll
      for (int i = 0; i < N; ++i) g = fmax(g, fabs(P[i][j]));
      if (g == T(0)) g = T(1);
      mean += g;
      qn = fmax(qn, fabs(q[j]));
    }
    mean /= T(N);
    c = T(1) / fmax(fmax(T(1e-6), mean), qn);
    int it = 0;
    T dev;
    do {
      T gx[N], gy[R];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T g = T(0);
#pragma unroll
        for (int k = 0; k < R; ++k) g = fmax(g, fabs((sy[k] * sx[j]) * A[k][j]));
        g = wmax(g);
#pragma unroll
        for (int i = 0; i < N; ++i) g = fmax(g, fabs(((c * sx[i]) * sx[j]) * P[i][j]));
        gx[j] = g;
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        T g = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) g = fmax(g, fabs((sy[k] * sx[j]) * A[k][j]));
        gy[k] = g;
      }
      dev = T(0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T g = gx[j];
        if (g == T(0)) g = T(1);
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];
        dev = fmax(dev, fabs(g - T(1)));
      }
      T dvy = T(0);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        T g = gy[k];
        if (g == T(0)) g = T(1);
        sy[k] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[k];
        if (valid[k]) dvy = fmax(dvy, fabs(g - T(1)));
      }
      dev = fmax(dev, wmax(dvy));
    } while (it++ < 10 && dev > T(0.1));
  }

  // N x N SPD inverse, Gauss-Jordan without pivoting, replicated on every lane.  False on a bad pivot.
  __device__ bool invert(T (&M)[N][N])
  {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      const T p = M[k][k];
      if (!(p > T(0)) || !(p < Num<T>::inf())) ok = false;
      const T pinv = T(1) / p;
      T rowk[N], colk[N];
#pragma unroll
      for (int j = 0; j < N; ++j) { rowk[j] = M[k][j] * pinv; colk[j] = M[j][k]; }
#pragma unroll
      for (int i = 0; i < N; ++i)
#pragma unroll
        for (int j = 0; j < N; ++j) {
          T v;
          if (i == k) v = (j == k) ? pinv : rowk[j];
          else if (j == k) v = -colk[i] * pinv;
          else v = M[i][j] - colk[i] * rowk[j];
          M[i][j] = v;
        }
    }
    return ok;
  }

  // check_stopping, qp_solver.hpp:574-644 -- same evaluation as qp_dense_group.cuh::check_stopping
  __device__ int check_stopping(const sfb_qp_params& prm)
  {
    const T eps_abs = T(prm.eps_abs), eps_rel = T(prm.eps_rel);
    const T eps_pinf = T(prm.eps_primal_inf), eps_dinf = T(prm.eps_dual_inf);
    const T inf = Num<T>::inf();
    T xus[N], dxs[N], dxus[N];
    T qn = T(0), dxn = T(0), qdx = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T d = x[j] - xold[j];
      xus[j] = sx[j] * x[j];
      dxs[j] = d;
      dxus[j] = sx[j] * d;
      qn = fmax(qn, fabs(q[j]));
      dxn = fmax(dxn, fabs(dxus[j]));
      qdx += q[j] * dxus[j];
    }
    T dy[R], Edy = T(0);
#pragma unroll
    for (int k = 0; k < R; ++k) {
      dy[k] = y[k] - yold[k];
      if (valid[k]) Edy = fmax(Edy, fabs(sy[k] * dy[k] / c));
    }
    Edy = wmax(Edy);
    T n_Ax = T(0), n_r = T(0), n_z = T(0), s_pinf = T(0);
    bool pinf_blocked = false, dinf_rows_ok = true;
    T aty[N], atdy[N];
#pragma unroll
    for (int j = 0; j < N; ++j) { aty[j] = T(0); atdy[j] = T(0); }
#pragma unroll
    for (int k = 0; k < R; ++k) {
      T ax = T(0), adx = T(0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        ax += A[k][j] * x[j];
        adx += A[k][j] * dxs[j];
        aty[j] += A[k][j] * y[k];
        atdy[j] += A[k][j] * dy[k];
      }
      if (valid[k]) {
        const T syinv = T(1) / sy[k];
        ax *= syinv;
        adx *= syinv;
        const T zus = syinv * z[k];
        n_Ax = fmax(n_Ax, fabs(ax));
        n_r = fmax(n_r, fabs(ax - zus));
        n_z = fmax(n_z, fabs(zus));
        const T dyus = sy[k] * dy[k] / c;
        const T li = l[k], ui = u[k];
        if (ui != inf) s_pinf += ui * fmax(T(0), dyus);
        else if (dyus > eps_pinf * Edy) pinf_blocked = true;
        if (li != -inf) s_pinf += li * fmin(T(0), dyus);
        else if (dyus < -eps_pinf * Edy) pinf_blocked = true;
        if (ui == inf) dinf_rows_ok = dinf_rows_ok && (adx >= -eps_dinf * dxn);
        else if (li == -inf) dinf_rows_ok = dinf_rows_ok && (adx <= eps_dinf * dxn);
        else dinf_rows_ok = dinf_rows_ok && (fabs(adx) < eps_dinf * dxn);
      }
    }
    n_Ax = wmax(n_Ax); n_r = wmax(n_r); n_z = wmax(n_z); s_pinf = wsum(s_pinf);
    pinf_blocked = __any_sync(kFullMask, pinf_blocked);
    dinf_rows_ok = __all_sync(kFullMask, dinf_rows_ok);
    if (pinf_blocked) s_pinf = inf;
    T n_Px = T(0), n_Aty = T(0), n_res = T(0), n_Atdy = T(0), n_Pdx = T(0);
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const T sc = T(1) / (sx[j] * c);
      const T ay = wsum(aty[j]) * sc;
      const T ady = wsum(atdy[j]) * sc;
      T px = T(0), pdx = T(0);
#pragma unroll
      for (int k2 = 0; k2 < N; ++k2) {
        px += P[j][k2] * xus[k2];
        pdx += P[j][k2] * dxus[k2];
      }
      n_Px = fmax(n_Px, fabs(px));
      n_Aty = fmax(n_Aty, fabs(ay));
      n_res = fmax(n_res, fabs(px + q[j] + ay));
      n_Atdy = fmax(n_Atdy, fabs(ady));
      n_Pdx = fmax(n_Pdx, fabs(pdx));
    }
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;
    // dx == 0 guard: DESIGN.md, "deliberate deviations"
    if ((dxn > T(0) || !dinf_guard) && (n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  __device__ void solve(const QpArgs<T>& a, long long b)
  {
    const unsigned long long t0 = a.prm.has_max_time ? global_timer_ns() : 0ull;
    load(a, b);
    dinf_guard = a.dinf_guard != 0;
    int code;
    unsigned iter;
    run(a.prm, a.max_iter_eff, t0, a.warm_x ? a.warm_x + b * N : nullptr, a.warm_y ? a.warm_y + b * (long long)m : nullptr, code, iter);
    store(a, b, code, iter);
  }

  // Everything between the staged problem data (A, l, u, P, q, valid, m, lane set by the caller) and the final iterate:
  // scale, rho classes, reduced KKT inverse, warm start, ADMM loop with stop checks.  On return x / y hold the SCALED
  // iterate (unscaled: sx x, sy y / c), code the status (kStatusUnset = iteration budget exhausted), iter the counter.
  __device__ void run(const sfb_qp_params& prm, unsigned max_iter_eff, unsigned long long t0, const T* warm_x, const T* warm_y,
                      int& code_out, unsigned& iter_out)
  {
    const T inf = Num<T>::inf();
    if (prm.scaling) scale();
    else {
      c = T(1);
#pragma unroll
      for (int j = 0; j < N; ++j) sx[j] = T(1);
#pragma unroll
      for (int k = 0; k < R; ++k) sy[k] = T(1);
    }
    int code = kStatusUnset;
    const T rho_bar = T(prm.rho), sigma = T(prm.sigma), alpha = T(prm.alpha), alpha_comp = T(1) - alpha;
    bool triv = false;
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const T li = l[k], ui = u[k];
      if (valid[k] && (li == inf || ui == -inf || ui - li < T(0))) triv = true;
      T r;
      if (li == -inf && ui == inf) r = T(1e-6);
      else if (sy[k] * fabs(li - ui) < T(1e-5)) r = T(1e3) * rho_bar;
      else r = rho_bar;
      rho[k] = r;
      rinv[k] = T(1) / r;
      lo[k] = sy[k] * li;
      hi[k] = sy[k] * ui;
    }
    if (__any_sync(kFullMask, triv)) code = SFB_QP_PRIMAL_INFEASIBLE;
    // scaled data and the reduced KKT matrix  M = Pbar(upper, mirrored) + sigma I + Abar^T R Abar
#pragma unroll
    for (int j = 0; j < N; ++j) qb[j] = (c * sx[j]) * q[j];
#pragma unroll
    for (int k = 0; k < R; ++k)
#pragma unroll
      for (int j = 0; j < N; ++j) A[k][j] = (sy[k] * A[k][j]) * sx[j];
    T M[N][N];
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = i; j < N; ++j) {
        T acc = T(0);
#pragma unroll
        for (int k = 0; k < R; ++k) acc += (rho[k] * A[k][i]) * A[k][j];
        acc = wsum(acc);
        T h = ((c * sx[i]) * P[i][j]) * sx[j];
        if (i == j) h += sigma;
        M[i][j] = h + acc;
        M[j][i] = h + acc;
      }
    if (!invert(M)) code = SFB_QP_UNKNOWN;
#pragma unroll
    for (int i = 0; i < N; ++i)
#pragma unroll
      for (int j = 0; j < N; ++j) Minv[i][j] = M[i][j];
    // initial iterate  :436-445
    if (warm_x != nullptr) {
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = (T(1) / sx[j]) * warm_x[j];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int i = lane + 32 * k;
        y[k] = valid[k] ? c * ((T(1) / sy[k]) * warm_y[i]) : T(0);
        T zt = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) zt += A[k][j] * x[j];
        z[k] = zt;
      }
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) x[j] = T(0);
#pragma unroll
      for (int k = 0; k < R; ++k) { y[k] = T(0); z[k] = T(0); }
    }
#pragma unroll
    for (int j = 0; j < N; ++j) xold[j] = T(0);
#pragma unroll
    for (int k = 0; k < R; ++k) yold[k] = T(0);

    const unsigned sci = prm.stop_check_iter;
    unsigned iter = 0;
#pragma unroll 1
    for (; iter != max_iter_eff && code == kStatusUnset; ++iter) {
      // rhs = sigma x - qb + Abar^T (R z - y)
      T t[N];
#pragma unroll
      for (int j = 0; j < N; ++j) t[j] = T(0);
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const T wk = rho[k] * z[k] - y[k];
#pragma unroll
        for (int j = 0; j < N; ++j) t[j] += A[k][j] * wk;
      }
      T rhs[N], xt[N];
#pragma unroll
      for (int j = 0; j < N; ++j) rhs[j] = sigma * x[j] - qb[j] + wsum(t[j]);
      const bool chk = (iter % sci == 1u);
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) acc += Minv[i][j] * rhs[j];
        xt[i] = acc;
        if (chk) xold[i] = x[i];
        x[i] = alpha * acc + alpha_comp * x[i];
      }
#pragma unroll
      for (int k = 0; k < R; ++k) {
        T zt = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) zt += A[k][j] * xt[j];
        const T zi = z[k], yi = y[k], ri = rho[k], rinvi = rinv[k];
        if (chk) yold[k] = yi;
        const T nu = ri * (zt - zi) + yi;
        T zn = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;
        zn = fmax(zn, lo[k]);
        zn = fmin(zn, hi[k]);
        y[k] = alpha_comp * yi + alpha * nu + ri * zi - ri * zn;
        z[k] = zn;
      }
      if (chk) {
        code = check_stopping(prm);
        if (code == kStatusUnset && prm.has_max_time) {
          const bool late = (lane == 0) && ((long long)(global_timer_ns() - t0) > prm.max_time_ns);
          if (__any_sync(kFullMask, late)) code = SFB_QP_MAX_TIME;
        }
      }
    }

    code_out = code;
    iter_out = iter;
  }

  // active sets (qp_solver.hpp:113-123) on the scaled dual, outputs, objective  :544-548
  __device__ void store(const QpArgs<T>& a, long long b, int code, unsigned iter)
  {
    const T inf = Num<T>::inf();
    const T thr = T(100) * Num<T>::eps();
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const int i = lane + 32 * k;
      if (!valid[k]) continue;
      int act = 0;
      if (y[k] < -thr && l[k] != -inf) act = -1;
      if (y[k] > thr && u[k] != inf) act = 1;
      if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
      a.out_y[b * (long long)m + i] = sy[k] * y[k] / c;
    }
    if (lane == 0) {
      T xus[N], obj = T(0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        xus[j] = sx[j] * x[j];
        a.out_x[b * N + j] = xus[j];
      }
#pragma unroll
      for (int i = 0; i < N; ++i) {
        T acc = T(0);
#pragma unroll
        for (int j = 0; j < N; ++j) acc += T(0.5) * P[i][j] * xus[j];
        obj += xus[i] * (acc + q[i]);
      }
      a.out_obj[b] = obj;
      a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
      a.out_iter[b] = iter;
      if (a.out_flags) a.out_flags[b] = 0u;
    }
  }
};

template <typename T, int N, int R>
__global__ void __launch_bounds__(32 * kSkinnyWarps) qp_dense_skinny_kernel(const __grid_constant__ QpArgs<T> a)
{
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(a.work_counter, 1ull);
    b = __shfl_sync(kFullMask, b, 0);
    if ((long long)b >= a.batch) break;
    QpSkinny<T, N, R> s;
    s.solve(a, (long long)b);
  }
}

}  // namespace sfb
