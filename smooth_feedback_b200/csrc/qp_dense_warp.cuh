// qp_dense_warp.cuh -- batched dense operator-splitting QP solver for sm_100a, one warp per QP instance.
//
// Replaces (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   QPSolver::scale           include/smooth/feedback/qp_solver.hpp:673-730
//   QPSolver::solve           include/smooth/feedback/qp_solver.hpp:343-568
//   QPSolver::check_stopping  include/smooth/feedback/qp_solver.hpp:574-644
//   detail::polish_qp         include/smooth/feedback/qp_solver.hpp:92-204
//
// Design: the whole working set of one instance lives in one warp's slice of shared
// memory for the lifetime of the solve; HBM sees the problem data once and the solution once.
//   * The reference factorises the (n+m)x(n+m) quasi-definite KKT matrix with a pivoted LDL^T and does two
//     triangular sweeps per ADMM iteration.  Triangular sweeps are a serial dependency chain for a warp, so
//     this kernel eliminates the diagonal (2,2) block -1/rho analytically:
//         (Pbar + sigma I + Abar^T R Abar) xt = sigma x - qbar + Abar^T (R z - y),   nu = R (Abar xt - z) + y
//     and keeps the explicit n x n inverse Minv in shared memory, which turns every iteration into three
//     conflict-free GEMV passes (Abar^T w, Minv rhs, Abar xt).  Mathematically identical to the KKT solve.
//   * Abar (m x n) and Minv (n x n) are stored column-major with an ODD leading dimension so that both
//     lane-per-row (consecutive addresses) and lane-per-column (stride ld) accesses are bank-conflict free.
//   * polish solves the same reduced KKT system the reference builds, by block elimination with explicit
//     inverses (Kinv = (Pbar + delta I)^-1, Sinv = (delta I + Aa Kinv Aa^T)^-1) held in the same shared slice.

#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/sfb.h"

namespace sfb {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kStatusUnset = -1;

template <typename T> struct Num;
template <> struct Num<double>
{
  __device__ static double inf() { return CUDART_INF; }
  __device__ static double eps() { return 2.220446049250313e-16; }
};
template <> struct Num<float>
{
  __device__ static float inf() { return CUDART_INF_F; }
  __device__ static float eps() { return 1.1920928955078125e-07f; }
};

template <typename T> __device__ __forceinline__ T warp_max(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}
template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
  return v;
}
__device__ __forceinline__ bool warp_any(bool p) { return __any_sync(kFullMask, p); }
__device__ __forceinline__ bool warp_all(bool p) { return __all_sync(kFullMask, p); }

__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

template <typename T> struct QpArgs
{
  const T* P;
  const T* q;
  const T* A;
  const T* l;
  const T* u;
  const T* warm_x;
  const T* warm_y;
  T* out_x;
  T* out_y;
  T* out_obj;
  int32_t* out_status;
  uint32_t* out_iter;
  int8_t* out_active;
  uint32_t* out_flags;
  // scale-only mode outputs (mode == 1)
  T* out_c;
  T* out_sx;
  T* out_sy;
  long long batch;
  int n, m;
  int mode;  // 0 = solve, 1 = scale only
  sfb_qp_params prm;
  unsigned max_iter_eff;
  unsigned long long* work_counter;
};

// per-warp shared-memory layout (units of T)
struct QpLayout
{
  int ldA, ldN;
  int offAs, offMs, offN, offM, total;
  __host__ __device__ static int odd(int v) { return v | 1; }
  __host__ __device__ QpLayout(int n, int m)
  {
    ldA = odd(m > 0 ? m : 1);
    ldN = odd(n);
    offAs = 0;
    offMs = offAs + ldA * n;
    offN = offMs + ldN * n;
    offM = offN + 8 * n;
    total = offM + 10 * (m > 0 ? m : 1);
    total = (total + 1) & ~1;  // keep every warp slice 16-byte aligned
  }
};

template <typename T> struct QpWarp
{
  // geometry
  int n, m, ldA, ldN, lane;
  // matrices
  T* As;  // m x n   raw A, then Abar = Sy A Sx, (polish: active rows compacted on top, Sinv below)
  T* Ms;  // n x n   raw P, then M, then Minv, (polish: Kinv)
  // n-vectors
  T *sx, *q, *qb, *x, *xt, *xold, *nv1, *nv2;
  // m-vectors
  T *sy, *l, *u, *rho, *rinv, *z, *y, *w, *yold, *mv1;
  T c;

  __device__ QpWarp(T* base, int n_, int m_, int lane_) : n(n_), m(m_), lane(lane_)
  {
    QpLayout L(n_, m_);
    ldA = L.ldA;
    ldN = L.ldN;
    As = base + L.offAs;
    Ms = base + L.offMs;
    T* nv = base + L.offN;
    sx = nv; q = nv + n; qb = nv + 2 * n; x = nv + 3 * n; xt = nv + 4 * n; xold = nv + 5 * n; nv1 = nv + 6 * n;
    nv2 = nv + 7 * n;
    T* mv = base + L.offM;
    const int mm = m > 0 ? m : 1;
    sy = mv; l = mv + mm; u = mv + 2 * mm; rho = mv + 3 * mm; rinv = mv + 4 * mm; z = mv + 5 * mm; y = mv + 6 * mm;
    w = mv + 7 * mm; yold = mv + 8 * mm; mv1 = mv + 9 * mm;
    c = T(1);
  }

  // ------------------------------------------------------------------------------------------------
  // stage one instance HBM -> shared memory (coalesced: consecutive lanes read consecutive elements)
  // ------------------------------------------------------------------------------------------------
  __device__ void load(const QpArgs<T>& a, long long b)
  {
    const T* gA = a.A + b * (long long)m * n;
    const T* gP = a.P + b * (long long)n * n;
    {
      int i = lane, j = 0;
      while (m > 0 && i >= m) { i -= m; ++j; }
      for (int e = lane; e < m * n; e += 32) {
        As[i + ldA * j] = __ldg(gA + e);
        i += 32;
        while (i >= m) { i -= m; ++j; }
      }
    }
    {
      int i = lane, j = 0;
      while (i >= n) { i -= n; ++j; }
      for (int e = lane; e < n * n; e += 32) {
        Ms[i + ldN * j] = __ldg(gP + e);
        i += 32;
        while (i >= n) { i -= n; ++j; }
      }
    }
    for (int j = lane; j < n; j += 32) q[j] = __ldg(a.q + b * (long long)n + j);
    for (int i = lane; i < m; i += 32) {
      l[i] = __ldg(a.l + b * (long long)m + i);
      u[i] = __ldg(a.u + b * (long long)m + i);
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------------
  // QPSolver::scale, qp_solver.hpp:673-730.  Products are formed in the reference's order so that
  // c, sx, sy agree with the CPU restatement bit for bit (max/abs/mul/sqrt/div only -- nothing to contract).
  // ------------------------------------------------------------------------------------------------
  __device__ void scale()
  {
    for (int j = lane; j < n; j += 32) sx[j] = T(1);
    for (int i = lane; i < m; i += 32) sy[i] = T(1);
    T qn = T(0);
    for (int j = lane; j < n; j += 32) {
      T g = T(0);
      for (int i = 0; i < n; ++i) g = fmax(g, fabs(Ms[i + ldN * j]));  // :681-685
      if (g == T(0)) g = T(1);                                           // :688-690
      nv1[j] = g;
      qn = fmax(qn, fabs(q[j]));
    }
    qn = warp_max(qn);
    __syncwarp();
    T mean = T(0);
    if (lane == 0) {
      for (int j = 0; j < n; ++j) mean += nv1[j];  // sequential on purpose: same rounding as the oracle
      mean /= T(n);
    }
    mean = __shfl_sync(kFullMask, mean, 0);
    c = T(1) / fmax(fmax(T(1e-6), mean), qn);  // :693
    __syncwarp();

    int it = 0;
    bool again;
    do {
      // column norms of [Ps As'; As 0]  :701-716
      for (int j = lane; j < n; j += 32) {
        const T sxj = sx[j];
        T g = T(0);
        for (int i = 0; i < n; ++i) g = fmax(g, fabs(((c * sx[i]) * sxj) * Ms[i + ldN * j]));
        for (int i = 0; i < m; ++i) g = fmax(g, fabs((sy[i] * sxj) * As[i + ldA * j]));
        if (g == T(0)) g = T(1);
        nv1[j] = g;
      }
      for (int i = lane; i < m; i += 32) {
        const T syi = sy[i];
        T g = T(0);
        for (int j = 0; j < n; ++j) g = fmax(g, fabs((syi * sx[j]) * As[i + ldA * j]));
        if (g == T(0)) g = T(1);
        mv1[i] = g;
      }
      __syncwarp();
      T dev = T(0);
      for (int j = lane; j < n; j += 32) {
        const T g = nv1[j];
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];  // :726
        dev = fmax(dev, fabs(g - T(1)));
      }
      for (int i = lane; i < m; i += 32) {
        const T g = mv1[i];
        sy[i] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[i];  // :727
        dev = fmax(dev, fabs(g - T(1)));
      }
      dev = warp_max(dev);
      __syncwarp();
      again = (it++ < 10) && (dev > T(0.1));  // :728-729
    } while (again);
  }

  // ------------------------------------------------------------------------------------------------
  // In-place Gauss-Jordan inversion of an SPD matrix (no pivoting), lane-per-row.  `scr` holds >= sz scalars.
  // Returns false (warp-uniform) on a non-positive or non-finite pivot.
  // ------------------------------------------------------------------------------------------------
  __device__ bool gj_invert(T* Mx, int ld, int sz, T* scr)
  {
    for (int k = 0; k < sz; ++k) {
      const T p = Mx[k + ld * k];
      if (!(p > T(0)) || !(p < Num<T>::inf())) return false;
      const T pinv = T(1) / p;
      for (int j = lane; j < sz; j += 32) scr[j] = (j == k) ? T(0) : Mx[k + ld * j] * pinv;
      __syncwarp();
      for (int i = lane; i < sz; i += 32) {
        if (i != k) {
          const T f = Mx[i + ld * k];
          for (int j = 0; j < sz; ++j) Mx[i + ld * j] -= f * scr[j];
          Mx[i + ld * k] = -f * pinv;
        } else {
          for (int j = 0; j < sz; ++j) Mx[k + ld * j] = scr[j];
          Mx[k + ld * k] = pinv;
        }
      }
      __syncwarp();
    }
    return true;
  }

  // t_j = sum_i As[i, j] * v1[i]  (and the same with v2) for the columns j = lane + 32*s this lane owns.
  // Two strategies: lane-per-column (n large) or row-partials + shuffle reduction (tall-skinny, n <= 8).
  // Results are written to o1[j] (and o2[j] if v2 != nullptr); caller syncs.
  __device__ void At_vec(const T* v1, const T* v2, int rows, T* o1, T* o2)
  {
    if (n > 8) {
      for (int j = lane; j < n; j += 32) {
        const T* col = As + ldA * j;
        T a1 = T(0), a2 = T(0);
        if (v2) {
          for (int i = 0; i < rows; ++i) {
            const T aij = col[i];
            a1 += aij * v1[i];
            a2 += aij * v2[i];
          }
          o2[j] = a2;
        } else {
          T b1 = T(0);
          int i = 0;
          for (; i + 1 < rows; i += 2) {
            a1 += col[i] * v1[i];
            b1 += col[i + 1] * v1[i + 1];
          }
          if (i < rows) a1 += col[i] * v1[i];
          a1 += b1;
        }
        o1[j] = a1;
      }
    } else {
      for (int j = 0; j < n; ++j) {
        const T* col = As + ldA * j;
        T a1 = T(0), a2 = T(0);
        for (int i = lane; i < rows; i += 32) {
          const T aij = col[i];
          a1 += aij * v1[i];
          if (v2) a2 += aij * v2[i];
        }
        a1 = warp_sum(a1);
        if (v2) a2 = warp_sum(a2);
        if (lane == 0) {
          o1[j] = a1;
          if (v2) o2[j] = a2;
        }
      }
    }
  }

  // ------------------------------------------------------------------------------------------------
  // check_stopping, qp_solver.hpp:574-644.  Called right after the iterate update of a check iteration
  // with xold / yold holding the pre-update iterates.  Uses nv1, nv2, xold, w, mv1 as scratch.
  // A x_us is evaluated as Sy^-1 (Abar x) and A^T y_us as Sx^-1 Abar^T y / c (same quantities).
  // ------------------------------------------------------------------------------------------------
  __device__ int check_stopping(const QpArgs<T>& a, const T* gP)
  {
    const T eps_abs = T(a.prm.eps_abs), eps_rel = T(a.prm.eps_rel);
    const T eps_pinf = T(a.prm.eps_primal_inf), eps_dinf = T(a.prm.eps_dual_inf);
    const T inf = Num<T>::inf();

    // n-space: x_us -> nv1, dx (scaled) -> nv2, dx_us -> xold ;  norms of q, dx_us
    T qn = T(0), dxn = T(0), qdx = T(0);
    for (int j = lane; j < n; j += 32) {
      const T xj = x[j];
      const T d = xj - xold[j];
      nv1[j] = sx[j] * xj;      // :481
      nv2[j] = d;
      const T dus = sx[j] * d;  // :484
      xold[j] = dus;
      qn = fmax(qn, fabs(q[j]));
      dxn = fmax(dxn, fabs(dus));
      qdx += q[j] * dus;
    }
    qn = warp_max(qn);
    dxn = warp_max(dxn);
    qdx = warp_sum(qdx);
    // m-space: dy (scaled) -> w ; ||dy_us||
    T Edy = T(0);
    for (int i = lane; i < m; i += 32) {
      const T d = y[i] - yold[i];
      w[i] = d;
      Edy = fmax(Edy, fabs(sy[i] * d / c));  // :485
    }
    Edy = warp_max(Edy);
    __syncwarp();

    // row pass: A x_us, A dx_us
    T n_Ax = T(0), n_r = T(0), n_z = T(0), s_pinf = T(0);
    bool pinf_blocked = false, dinf_rows_ok = true;
    for (int i = lane; i < m; i += 32) {
      T ax = T(0), adx = T(0);
      for (int j = 0; j < n; ++j) {
        const T aij = As[i + ldA * j];
        ax += aij * x[j];
        adx += aij * nv2[j];
      }
      const T syinv = T(1) / sy[i];
      ax *= syinv;
      adx *= syinv;
      const T zus = syinv * z[i];  // :483
      n_Ax = fmax(n_Ax, fabs(ax));
      n_r = fmax(n_r, fabs(ax - zus));
      n_z = fmax(n_z, fabs(zus));
      const T dyus = sy[i] * w[i] / c;
      const T li = l[i], ui = u[i];
      // :602-617 (the reference's early break only matters through "any trigger -> +inf")
      if (ui != inf) s_pinf += ui * fmax(T(0), dyus);
      else if (dyus > eps_pinf * Edy) pinf_blocked = true;
      if (li != -inf) s_pinf += li * fmin(T(0), dyus);
      else if (dyus < -eps_pinf * Edy) pinf_blocked = true;
      // :631-639
      if (ui == inf) dinf_rows_ok = dinf_rows_ok && (adx >= -eps_dinf * dxn);
      else if (li == -inf) dinf_rows_ok = dinf_rows_ok && (adx <= eps_dinf * dxn);
      else dinf_rows_ok = dinf_rows_ok && (fabs(adx) < eps_dinf * dxn);
    }
    n_Ax = warp_max(n_Ax);
    n_r = warp_max(n_r);
    n_z = warp_max(n_z);
    s_pinf = warp_sum(s_pinf);
    pinf_blocked = warp_any(pinf_blocked);
    dinf_rows_ok = warp_all(dinf_rows_ok);
    if (pinf_blocked) s_pinf = inf;

    // column pass: Abar^T y -> mv1[0..n) is not usable (m may be < n); use xt / qb-free scratch: xt and nv2 reuse
    // (nv2 = scaled dx is no longer needed after the row pass)
    __syncwarp();
    At_vec(y, w, m, xt, nv2);
    __syncwarp();
    T n_Px = T(0), n_Aty = T(0), n_res = T(0), n_Atdy = T(0), n_Pdx = T(0);
    for (int j = lane; j < n; j += 32) {
      const T sc = T(1) / (sx[j] * c);
      const T aty = xt[j] * sc;
      const T atdy = nv2[j] * sc;
      T px = T(0), pdx = T(0);
      for (int k = 0; k < n; ++k) {
        const T pjk = __ldg(gP + j + (long long)n * k);  // row j of the unscaled P (coalesced across lanes)
        px += pjk * nv1[k];
        pdx += pjk * xold[k];
      }
      n_Px = fmax(n_Px, fabs(px));
      n_Aty = fmax(n_Aty, fabs(aty));
      n_res = fmax(n_res, fabs(px + q[j] + aty));
      n_Atdy = fmax(n_Atdy, fabs(atdy));
      n_Pdx = fmax(n_Pdx, fabs(pdx));
    }
    n_Px = warp_max(n_Px);
    n_Aty = warp_max(n_Aty);
    n_res = warp_max(n_res);
    n_Atdy = warp_max(n_Atdy);
    n_Pdx = warp_max(n_Pdx);
    __syncwarp();

    // OPTIMALITY :584-594
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    // PRIMAL INFEASIBILITY :619
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;
    // DUAL INFEASIBILITY :629-641
    if ((n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  // (Pbar v)_i with Pbar = upper triangle of c Sx P Sx mirrored (selfadjointView<Upper>), from global P
  __device__ T pbar_row_dot(const T* gP, int i, const T* v)
  {
    T acc = T(0);
    const T csxi = c * sx[i];
    for (int j = 0; j < n; ++j) {
      T h;
      if (i <= j) h = (csxi * __ldg(gP + i + (long long)n * j)) * sx[j];
      else h = ((c * sx[j]) * __ldg(gP + j + (long long)n * i)) * sx[i];
      acc += h * v[j];
    }
    return acc;
  }

  // ------------------------------------------------------------------------------------------------
  // detail::polish_qp, qp_solver.hpp:92-204.  `na` active rows (ascending) are listed in idx[], their scaled
  // bounds in bnd[].  Returns SFB_QP_FLAG_* bits.
  // ------------------------------------------------------------------------------------------------
  __device__ unsigned polish(const QpArgs<T>& a, const T* gP, int na, const int* idx, const T* bnd)
  {
    if (na > n || 2 * na > ldA) return SFB_QP_FLAG_POLISH_SKIPPED;
    const T delta = T(a.prm.delta);

    // compact the active rows of Abar to the top of every column (idx ascending => in-place safe)
    for (int j = lane; j < n; j += 32) {
      T* col = As + ldA * j;
      for (int r = 0; r < na; ++r) col[r] = col[idx[r]];
    }
    // Kinv = (Pbar + delta I)^-1 in Ms   :161,175
    for (int i = lane; i < n; i += 32) {
      const T csxi = c * sx[i];
      for (int j = 0; j < n; ++j) {
        T h;
        if (i <= j) h = (csxi * __ldg(gP + i + (long long)n * j)) * sx[j];
        else h = ((c * sx[j]) * __ldg(gP + j + (long long)n * i)) * sx[i];
        if (i == j) h += delta;
        Ms[i + ldN * j] = h;
      }
    }
    __syncwarp();
    if (!gj_invert(Ms, ldN, n, nv1)) return SFB_QP_FLAG_POLISH_FAILED;

    // S = delta I + Aa Kinv Aa^T, stored below the compacted rows: S(r, s) at As[na + r + ldA * s]
    T* S = As + na;
    for (int s = 0; s < na; ++s) {
      for (int i = lane; i < n; i += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += Ms[i + ldN * j] * As[s + ldA * j];
        nv1[i] = acc;
      }
      __syncwarp();
      for (int r = lane; r < na; r += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += As[r + ldA * j] * nv1[j];
        if (r == s) acc += delta;
        S[r + ldA * s] = acc;
      }
      __syncwarp();
    }
    if (na > 0 && !gj_invert(S, ldA, na, nv1)) return SFB_QP_FLAG_POLISH_FAILED;

    // iterative refinement  t += Hp^-1 (h - H t)   :192-195
    T* tx = xt;    // n
    T* ty = z;     // na
    T* rx = nv1;   // n
    T* ux = nv2;   // n
    T* ry = yold;  // na
    T* sv = rho;   // na
    T* dy = rinv;  // na
    for (int j = lane; j < n; j += 32) tx[j] = T(0);
    for (int r = lane; r < na; r += 32) ty[r] = T(0);
    __syncwarp();
    for (uint32_t it = 0; it != a.prm.polish_iter; ++it) {
      // residual r = h - sym(H) t,  H = [Pbar Aa^T; Aa 0]
      for (int i = lane; i < n; i += 32) {
        T acc = pbar_row_dot(gP, i, tx);
        const T* col = As + ldA * i;
        for (int r = 0; r < na; ++r) acc += col[r] * ty[r];
        rx[i] = -c * (sx[i] * q[i]) - acc;  // :180
      }
      for (int r = lane; r < na; r += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += As[r + ldA * j] * tx[j];
        ry[r] = bnd[r] - acc;  // :181-182
      }
      __syncwarp();
      // solve [K Aa^T; Aa -delta I] [dx; dy] = [rx; ry]:  dy = Sinv (Aa Kinv rx - ry),  dx = Kinv (rx - Aa^T dy)
      for (int i = lane; i < n; i += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += Ms[i + ldN * j] * rx[j];
        ux[i] = acc;
      }
      __syncwarp();
      for (int r = lane; r < na; r += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += As[r + ldA * j] * ux[j];
        sv[r] = acc - ry[r];
      }
      __syncwarp();
      for (int r = lane; r < na; r += 32) {
        T acc = T(0);
        for (int s = 0; s < na; ++s) acc += S[r + ldA * s] * sv[s];
        dy[r] = acc;
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        const T* col = As + ldA * i;
        T acc = rx[i];
        for (int r = 0; r < na; ++r) acc -= col[r] * dy[r];
        ux[i] = acc;
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += Ms[i + ldN * j] * ux[j];
        tx[i] += acc;
      }
      for (int r = lane; r < na; r += 32) ty[r] += dy[r];
      __syncwarp();
    }
    // :199-201
    for (int j = lane; j < n; j += 32) x[j] = tx[j];
    for (int r = lane; r < na; r += 32) y[idx[r]] = ty[r];
    __syncwarp();
    return SFB_QP_FLAG_POLISHED;
  }

  // ------------------------------------------------------------------------------------------------
  // QPSolver::solve, qp_solver.hpp:343-568
  // ------------------------------------------------------------------------------------------------
  __device__ void solve(const QpArgs<T>& a, long long b)
  {
    const T inf = Num<T>::inf();
    const T* gP = a.P + b * (long long)n * n;
    const unsigned long long t0 = a.prm.has_max_time ? global_timer_ns() : 0ull;

    load(a, b);
    if (a.prm.scaling) {
      scale();  // :347
    } else {
      for (int j = lane; j < n; j += 32) sx[j] = T(1);
      for (int i = lane; i < m; i += 32) sy[i] = T(1);
      c = T(1);
      __syncwarp();
    }
    if (a.mode == 1) {
      if (lane == 0) a.out_c[b] = c;
      for (int j = lane; j < n; j += 32) a.out_sx[b * (long long)n + j] = sx[j];
      for (int i = lane; i < m; i += 32) a.out_sy[b * (long long)m + i] = sy[i];
      __syncwarp();
      return;
    }

    const T rho_bar = T(a.prm.rho), alpha = T(a.prm.alpha), alpha_comp = T(1) - alpha, sigma = T(a.prm.sigma);
    int code = kStatusUnset;

    // rho per constraint class + trivially empty feasible set  :361-374
    bool triv = false;
    for (int i = lane; i < m; i += 32) {
      const T li = l[i], ui = u[i];
      if (li == inf || ui == -inf || ui - li < T(0)) triv = true;
      T r;
      if (li == -inf && ui == inf) r = T(1e-6);
      else if (sy[i] * fabs(li - ui) < T(1e-5)) r = T(1e3) * rho_bar;
      else r = rho_bar;
      rho[i] = r;
      rinv[i] = T(1) / r;
    }
    if (warp_any(triv)) code = SFB_QP_PRIMAL_INFEASIBLE;

    // scale the data in place:  qb = c Sx q,  Abar = Sy A Sx,  upper(Pbar) = c Sx P Sx   :401-403,:450
    for (int j = lane; j < n; j += 32) qb[j] = (c * sx[j]) * q[j];
    for (int i = lane; i < m; i += 32) {
      const T syi = sy[i];
      for (int j = 0; j < n; ++j) As[i + ldA * j] = (syi * As[i + ldA * j]) * sx[j];
    }
    for (int i = lane; i < n; i += 32) {
      const T csxi = c * sx[i];
      for (int j = i; j < n; ++j) Ms[i + ldN * j] = (csxi * Ms[i + ldN * j]) * sx[j];
    }
    __syncwarp();

    // reduced KKT matrix  M = Pbar + sigma I + Abar^T R Abar  (upper triangle computed, mirrored)
    for (int j = 0; j < n; ++j) {
      for (int r = lane; r < m; r += 32) w[r] = rho[r] * As[r + ldA * j];
      __syncwarp();
      for (int i = lane; i <= j; i += 32) {
        const T* col = As + ldA * i;
        T a0 = Ms[i + ldN * j], a1 = T(0);
        int r = 0;
        for (; r + 1 < m; r += 2) {
          a0 += col[r] * w[r];
          a1 += col[r + 1] * w[r + 1];
        }
        if (r < m) a0 += col[r] * w[r];
        a0 += a1;
        if (i == j) a0 += sigma;
        Ms[i + ldN * j] = a0;
        Ms[j + ldN * i] = a0;
      }
      __syncwarp();
    }
    if (!gj_invert(Ms, ldN, n, nv1)) code = SFB_QP_UNKNOWN;  // :433

    // initial iterate  :436-445
    if (a.warm_x != nullptr) {
      for (int j = lane; j < n; j += 32) x[j] = (T(1) / sx[j]) * __ldg(a.warm_x + b * (long long)n + j);
      for (int i = lane; i < m; i += 32) y[i] = c * ((T(1) / sy[i]) * __ldg(a.warm_y + b * (long long)m + i));
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        T acc = T(0);
        for (int j = 0; j < n; ++j) acc += As[i + ldA * j] * x[j];
        z[i] = acc;
      }
    } else {
      for (int j = lane; j < n; j += 32) x[j] = T(0);
      for (int i = lane; i < m; i += 32) {
        y[i] = T(0);
        z[i] = T(0);
      }
    }
    __syncwarp();

    // main ADMM loop  :449-510
    const unsigned sci = a.prm.stop_check_iter;
    unsigned iter = 0;
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      // rhs_x = sigma x - qb + Abar^T (R z - y)
      for (int i = lane; i < m; i += 32) w[i] = rho[i] * z[i] - y[i];
      __syncwarp();
      At_vec(w, nullptr, m, xt, nullptr);
      __syncwarp();
      for (int j = lane; j < n; j += 32) xt[j] = sigma * x[j] - qb[j] + xt[j];
      __syncwarp();
      // xtilde = Minv rhs_x
      for (int i = lane; i < n; i += 32) {
        T a0 = T(0), a1 = T(0);
        int j = 0;
        for (; j + 1 < n; j += 2) {
          a0 += Ms[i + ldN * j] * xt[j];
          a1 += Ms[i + ldN * (j + 1)] * xt[j + 1];
        }
        if (j < n) a0 += Ms[i + ldN * j] * xt[j];
        nv1[i] = a0 + a1;
      }
      const bool chk = (iter % sci == 1u);
      if (chk) {  // :465-468
        for (int j = lane; j < n; j += 32) xold[j] = x[j];
        for (int i = lane; i < m; i += 32) yold[i] = y[i];
      }
      __syncwarp();
      for (int j = lane; j < n; j += 32) x[j] = alpha * nv1[j] + alpha_comp * x[j];  // :470
      // nu = R (Abar xtilde - z) + y ; z, y updates  :471-477
      for (int i = lane; i < m; i += 32) {
        T a0 = T(0), a1 = T(0);
        int j = 0;
        for (; j + 1 < n; j += 2) {
          a0 += As[i + ldA * j] * nv1[j];
          a1 += As[i + ldA * (j + 1)] * nv1[j + 1];
        }
        if (j < n) a0 += As[i + ldA * j] * nv1[j];
        const T zt = a0 + a1;
        const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];
        const T nu = ri * (zt - zi) + yi;
        T v = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;
        v = fmax(v, sy[i] * l[i]);
        v = fmin(v, sy[i] * u[i]);
        y[i] = alpha_comp * yi + alpha * nu + ri * zi - ri * v;
        z[i] = v;
      }
      __syncwarp();
      if (chk) {
        code = check_stopping(a, gP);  // :488
        if (code == kStatusUnset && a.prm.has_max_time &&
            (long long)(global_timer_ns() - t0) > a.prm.max_time_ns)
          code = SFB_QP_MAX_TIME;  // :504-508
      }
    }

    // active sets as polish_qp builds them (:113-123), ascending order, on the scaled dual
    int na = 0;
    int* idx = reinterpret_cast<int*>(mv1);
    T* bnd = w;
    {
      const T thr = T(100) * Num<T>::eps();
      for (int base = 0; base < m; base += 32) {
        const int i = base + lane;
        int act = 0;
        T bv = T(0);
        if (i < m) {
          if (y[i] < -thr && l[i] != -inf) { act = -1; bv = sy[i] * l[i]; }
          if (y[i] > thr && u[i] != inf) { act = 1; bv = sy[i] * u[i]; }
          if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
        }
        const unsigned bal = __ballot_sync(kFullMask, act != 0);
        if (act != 0) {
          const int pos = na + __popc(bal & ((1u << lane) - 1u));
          idx[pos] = i;
          bnd[pos] = bv;
        }
        na += __popc(bal);
      }
      __syncwarp();
    }

    unsigned flags = 0;
    if (code == SFB_QP_OPTIMAL && a.prm.polish) flags = polish(a, gP, na, idx, bnd);  // :515-539

    // unscale + objective  :544-548
    for (int j = lane; j < n; j += 32) {
      const T v = sx[j] * x[j];
      nv1[j] = v;
      a.out_x[b * (long long)n + j] = v;
    }
    for (int i = lane; i < m; i += 32) a.out_y[b * (long long)m + i] = sy[i] * y[i] / c;
    __syncwarp();
    T obj = T(0);
    for (int i = lane; i < n; i += 32) {
      T acc = T(0);
      for (int j = 0; j < n; ++j) acc += T(0.5) * __ldg(gP + i + (long long)n * j) * nv1[j];
      obj += nv1[i] * (acc + q[i]);
    }
    obj = warp_sum(obj);
    if (lane == 0) {
      a.out_obj[b] = obj;
      a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
      a.out_iter[b] = iter;
      if (a.out_flags) a.out_flags[b] = flags;
    }
    __syncwarp();
  }
};

// One warp per instance; warps pull instances from a global work counter (iteration counts are
// heavy-tailed), so a slow instance never idles the rest of the grid.
template <typename T> __global__ void qp_dense_warp_kernel(const QpArgs<T> a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  QpLayout L(a.n, a.m);
  T* base = reinterpret_cast<T*>(smem_raw) + (size_t)warp * L.total;
  QpWarp<T> s(base, a.n, a.m, lane);
  for (;;) {
    unsigned long long b = 0;
    if (lane == 0) b = atomicAdd(a.work_counter, 1ull);
    b = __shfl_sync(kFullMask, b, 0);
    if ((long long)b >= a.batch) break;
    s.solve(a, (long long)b);
  }
}

}  // namespace sfb
