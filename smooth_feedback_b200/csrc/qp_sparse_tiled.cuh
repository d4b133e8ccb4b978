// qp_sparse_tiled.cuh -- batched sparse operator-splitting QP solver for sm_100a (shared sparsity pattern).
//
// Replaces the sparse branches of (pettni/smooth_feedback @ 9a08971)
//   QPSolver<QuadraticProgramSparse>::solve   include/smooth/feedback/qp_solver.hpp:343-568 (fill :380-397,
//                                             SimplicialLDLT factorize :424-426, permuted solve :456-460)
//   QPSolver::scale / check_stopping          :673-730 / :574-644 (InnerIterator walks only stored entries)
// i.e. the call MPC::operator() makes at mpc.hpp:491.
//
// Layout: thread-per-instance.  A warp owns a TILE of 32 instances; every per-instance array of the working set
// lives in global memory as [tile][element][32], so that the 32 lanes of a warp -- which all execute the same
// pattern-driven program (the index arrays are shared by the whole batch and broadcast) -- touch 32 consecutive
// scalars per access: every load and store of the solve is a fully coalesced 256-byte (fp64) line.  The per-iteration
// traffic is therefore exactly the north star's model: one pass over Abar twice and over the L D L^T factor twice,
// streamed from HBM/L2 (B_iter), plus the vectors.
//
// The KKT system is reduced as in the dense kernel:  (Pbar + sigma I + Abar^T R Abar) xt = sigma x - qbar + Abar^T (R z - y),
// nu = R (Abar xt - z) + y; the n x n matrix is factorised L D L^T without pivoting (it is SPD) in the fill-reducing
// order computed on the host (qp_sparse_host.hpp).  Variables are kept in the permuted order throughout.
//
// P is used exactly as the reference uses it: only entries with col >= row enter the factorised matrix
// (qp_solver.hpp:384), while scale(), the dual residual, the dual-infeasibility test and the objective multiply
// with the entries AS STORED (:589,627,547,681-707) -- MPC stores the upper triangle only (ocp_to_qp.hpp:97-105),
// and the reference's residuals inherit that.

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfb.h"
#include "qp_dense_group.cuh"  // Num<T>, kStatusUnset, global_timer_ns

namespace sfb {

constexpr int kSpNV = 10;  // n-vectors of the working set
constexpr int kSpMV = 9;   // m-vectors

struct SpPattern
{
  int n, m, nnzP, nnzA, nnzL;
  const int *perm, *iperm, *P_rowp, *P_colp, *P_tgt, *A_rowptr, *A_col, *A_pair_ptr, *A_pair_tgt, *L_colptr, *L_row,
    *F_ptr, *F_tgt;
};

template <typename T> struct SpArgs
{
  SpPattern pat;
  // inputs, array-of-instances:  P_vals [batch][nnzP], q [batch][n], A_vals [batch][nnzA], l,u [batch][m]
  const T *P, *q, *A, *l, *u, *warm_x, *warm_y;
  T *out_x, *out_y, *out_obj;
  int32_t* out_status;
  uint32_t* out_iter;
  int8_t* out_active;
  uint32_t* out_flags;
  // tiled workspace: [tile][len][32]
  T *wsA, *wsP, *wsW, *wsN, *wsM;
  long long batch;
  sfb_qp_params prm;
  unsigned max_iter_eff;
};

// element e of this lane's instance inside a [len][32] tile block
template <typename T> struct TV
{
  T* p;
  __device__ __forceinline__ T& operator[](int e) const { return p[(size_t)e * 32]; }
};

template <typename T> struct SpSolver
{
  const SpPattern& S;
  int n, m;
  TV<T> A, P, W;
  TV<T> q, qb, x, xold, v, sx, t1, t2, t3, t4;      // n-vectors (permuted order)
  TV<T> l, u, sy, rho, rinv, z, y, yold, w;        // m-vectors
  T c;

  __device__ SpSolver(const SpArgs<T>& a, long long tile, int lane) : S(a.pat), n(a.pat.n), m(a.pat.m)
  {
    A.p = a.wsA + (size_t)tile * S.nnzA * 32 + lane;
    P.p = a.wsP + (size_t)tile * S.nnzP * 32 + lane;
    W.p = a.wsW + (size_t)tile * (S.nnzL + n) * 32 + lane;
    T* nv = a.wsN + (size_t)tile * kSpNV * n * 32 + lane;
    T* mv = a.wsM + (size_t)tile * kSpMV * m * 32 + lane;
    auto N = [&](int k) { return TV<T>{nv + (size_t)k * n * 32}; };
    auto M = [&](int k) { return TV<T>{mv + (size_t)k * m * 32}; };
    q = N(0); qb = N(1); x = N(2); xold = N(3); v = N(4); sx = N(5); t1 = N(6); t2 = N(7); t3 = N(8); t4 = N(9);
    l = M(0); u = M(1); sy = M(2); rho = M(3); rinv = M(4); z = M(5); y = M(6); yold = M(7); w = M(8);
    c = T(1);
  }

  // ---------------------------------------------------------------- QPSolver::scale, qp_solver.hpp:673-730
  __device__ void scale()
  {
    for (int j = 0; j < n; ++j) { sx[j] = T(1); t1[j] = T(0); }
    for (int i = 0; i < m; ++i) sy[i] = T(1);
    for (int e = 0; e < S.nnzP; ++e) {
      const int cj = S.P_colp[e];
      t1[cj] = fmax(t1[cj], fabs(P[e]));
    }
    T mean = T(0), qn = T(0);
    for (int j = 0; j < n; ++j) {  // original column order: the mean is summed as the reference sums it
      const int pj = S.iperm[j];
      T g = t1[pj];
      if (g == T(0)) g = T(1);
      mean += g;
      qn = fmax(qn, fabs(q[pj]));
    }
    mean /= T(n);
    c = T(1) / fmax(fmax(T(1e-6), mean), qn);
    int it = 0;
    T dev;
    do {
      for (int j = 0; j < n; ++j) t1[j] = T(0);
      for (int e = 0; e < S.nnzP; ++e) {
        const int rj = S.P_rowp[e], cj = S.P_colp[e];
        t1[cj] = fmax(t1[cj], fabs(((c * sx[rj]) * sx[cj]) * P[e]));
      }
      for (int i = 0; i < m; ++i) {
        const T syi = sy[i];
        T g = T(0);
        for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) {
          const int cj = S.A_col[e];
          const T aij = fabs((syi * sx[cj]) * A[e]);
          t1[cj] = fmax(t1[cj], aij);
          g = fmax(g, aij);
        }
        w[i] = g;
      }
      dev = T(0);
      for (int j = 0; j < n; ++j) {
        T g = t1[j];
        if (g == T(0)) g = T(1);
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];
        dev = fmax(dev, fabs(g - T(1)));
      }
      for (int i = 0; i < m; ++i) {
        T g = w[i];
        if (g == T(0)) g = T(1);
        sy[i] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[i];
        dev = fmax(dev, fabs(g - T(1)));
      }
    } while (it++ < 10 && dev > T(0.1));
  }

  // ---------------------------------------------------------------- W <- shift I + c Sx triu(P) Sx + Abar^T diag(wt) Abar
  // (lower triangle in the factor's slots, diagonal in W[nnzL + i]); wt = rho for the ADMM system
  __device__ void assemble(T shift, const TV<T>& wt)
  {
    const int nW = S.nnzL + n;
    for (int e = 0; e < S.nnzL; ++e) W[e] = T(0);
    for (int e = S.nnzL; e < nW; ++e) W[e] = shift;
    for (int e = 0; e < S.nnzP; ++e) {
      const int t = S.P_tgt[e];
      if (t >= 0) W[t] += ((c * sx[S.P_rowp[e]]) * sx[S.P_colp[e]]) * P[e];  // qp_solver.hpp:386
    }
    for (int i = 0; i < m; ++i) {
      const T ri = wt[i];
      const int e0 = S.A_rowptr[i], e1 = S.A_rowptr[i + 1];
      int p = S.A_pair_ptr[i];
      if (ri == T(0)) continue;
      for (int ea = e0; ea < e1; ++ea) {
        const T ra = ri * A[ea];
        for (int eb = ea; eb < e1; ++eb) {
          const int t = S.A_pair_tgt[p++];
          W[t] += ra * A[eb];
        }
      }
    }
  }

  // right-looking L D L^T in place; afterwards W[e] = L entries, W[nnzL + k] = 1 / D_k.  False on a non-positive pivot.
  __device__ bool factor()
  {
    bool ok = true;
    const int nL = S.nnzL;
    for (int k = 0; k < n; ++k) {
      const T dk = W[nL + k];
      if (!(dk > T(0)) || !(dk < Num<T>::inf())) ok = false;
      const T dinv = T(1) / dk;
      const int c0 = S.L_colptr[k], c1 = S.L_colptr[k + 1];
      int p = S.F_ptr[k];
      for (int a = c0; a < c1; ++a) {
        const T la = W[a] * dinv;
#pragma unroll 4
        for (int b = a; b < c1; ++b) {
          const int t = S.F_tgt[p++];
          W[t] -= W[b] * la;
        }
      }
      for (int a = c0; a < c1; ++a) W[a] *= dinv;
      W[nL + k] = dinv;
    }
    return ok;
  }

  // v <- (L D L^T)^-1 v
  __device__ void solve()
  {
    const int nL = S.nnzL;
    for (int k = 0; k < n; ++k) {
      const T vk = v[k];
      const int c0 = S.L_colptr[k], c1 = S.L_colptr[k + 1];
#pragma unroll 4
      for (int e = c0; e < c1; ++e) {
        const int r = S.L_row[e];
        v[r] -= W[e] * vk;
      }
    }
    for (int k = n - 1; k >= 0; --k) {
      T acc = v[k] * W[nL + k];
      const int c0 = S.L_colptr[k], c1 = S.L_colptr[k + 1];
#pragma unroll 4
      for (int e = c0; e < c1; ++e) acc -= W[e] * v[S.L_row[e]];
      v[k] = acc;
    }
  }

  // out[col] (+)= sum_i Abar_ij in_i   (row-wise scatter; out must be zeroed by the caller)
  __device__ void At_acc(const TV<T>& in, const TV<T>& out)
  {
    for (int i = 0; i < m; ++i) {
      const T wi = in[i];
#pragma unroll 4
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) {
        const int cj = S.A_col[e];
        out[cj] += A[e] * wi;
      }
    }
  }

  // ---------------------------------------------------------------- check_stopping, qp_solver.hpp:574-644
  // Same evaluation as the dense kernel (qp_dense_group.cuh::check_stopping): A x_us = Sy^-1 (Abar x), A^T y_us = Sx^-1 Abar^T y / c.
  __device__ int check_stopping(const sfb_qp_params& prm)
  {
    const T eps_abs = T(prm.eps_abs), eps_rel = T(prm.eps_rel);
    const T eps_pinf = T(prm.eps_primal_inf), eps_dinf = T(prm.eps_dual_inf);
    const T inf = Num<T>::inf();
    T qn = T(0), dxn = T(0), qdx = T(0), Edy = T(0);
    for (int j = 0; j < n; ++j) {
      const T xj = x[j];
      const T d = xj - xold[j];
      t1[j] = sx[j] * xj;       // x_us   :481
      t2[j] = d;                // scaled dx
      const T dus = sx[j] * d;  // dx_us  :484
      xold[j] = dus;
      qn = fmax(qn, fabs(q[j]));
      dxn = fmax(dxn, fabs(dus));
      t3[j] = T(0);
      t4[j] = T(0);
    }
    for (int jo = 0; jo < n; ++jo) {  // q . dx_us summed in the original variable order
      const int j = S.iperm[jo];
      qdx += q[j] * xold[j];
    }
    for (int i = 0; i < m; ++i) {
      const T d = y[i] - yold[i];
      w[i] = d;
      Edy = fmax(Edy, fabs(sy[i] * d / c));  // :485
    }
    T n_Ax = T(0), n_r = T(0), n_z = T(0), s_pinf = T(0);
    bool pinf_blocked = false, dinf_rows_ok = true;
    for (int i = 0; i < m; ++i) {
      T ax = T(0), adx = T(0);
#pragma unroll 4
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) {
        const int cj = S.A_col[e];
        const T aij = A[e];
        ax += aij * x[cj];
        adx += aij * t2[cj];
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
      if (ui != inf) s_pinf += ui * fmax(T(0), dyus);  // :602-617
      else if (dyus > eps_pinf * Edy) pinf_blocked = true;
      if (li != -inf) s_pinf += li * fmin(T(0), dyus);
      else if (dyus < -eps_pinf * Edy) pinf_blocked = true;
      if (ui == inf) dinf_rows_ok = dinf_rows_ok && (adx >= -eps_dinf * dxn);  // :631-639
      else if (li == -inf) dinf_rows_ok = dinf_rows_ok && (adx <= eps_dinf * dxn);
      else dinf_rows_ok = dinf_rows_ok && (fabs(adx) < eps_dinf * dxn);
    }
    if (pinf_blocked) s_pinf = inf;
    // Abar^T y -> v, Abar^T dy -> t2 (scaled dx is dead)
    for (int j = 0; j < n; ++j) { v[j] = T(0); t2[j] = T(0); }
    At_acc(y, v);
    At_acc(w, t2);
    // P x_us -> t3, P dx_us -> t4 with the entries as stored
    for (int e = 0; e < S.nnzP; ++e) {
      const int rj = S.P_rowp[e], cj = S.P_colp[e];
      const T pv = P[e];
      t3[rj] += pv * t1[cj];
      t4[rj] += pv * xold[cj];
    }
    T n_Px = T(0), n_Aty = T(0), n_res = T(0), n_Atdy = T(0), n_Pdx = T(0);
    for (int j = 0; j < n; ++j) {
      const T sc = T(1) / (sx[j] * c);
      const T aty = v[j] * sc;
      const T atdy = t2[j] * sc;
      const T px = t3[j], pdx = t4[j];
      n_Px = fmax(n_Px, fabs(px));
      n_Aty = fmax(n_Aty, fabs(aty));
      n_res = fmax(n_res, fabs(px + q[j] + aty));
      n_Atdy = fmax(n_Atdy, fabs(atdy));
      n_Pdx = fmax(n_Pdx, fabs(pdx));
    }
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {  // :584-594
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;  // :619
    // dx == 0 guard: see qp_dense_group.cuh / DESIGN.md ("deliberate deviations")
    if ((dxn > T(0)) && (n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  // ---------------------------------------------------------------- detail::polish_qp, qp_solver.hpp:92-204
  // The regularised system  Hp s = r,  Hp = [Pbar + delta I, Aa^T; Aa, -delta I],  is solved through the SAME symbolic
  // factor: eliminating the (2,2) block gives (Pbar + delta I + Aa^T Aa / delta) s1 = r1 + Aa^T r2 / delta,
  // s2 = (Aa s1 - r2) / delta, whose pattern is a subset of M's.  That reduced matrix is ill conditioned (1/delta = 1e6
  // against delta), so every application of Hp^-1 is followed by two steps of iterative refinement on Hp itself; the
  // outer iteration t += Hp^-1 (h - H t) then follows the reference's sequence to ~1e-9.
  // On entry: w[i] = 1 for active rows else 0, yold = scaled active bound.  Uses rho, rinv, z, l, u and t1..t3, v as scratch.

  // in: r1 in v, r2 in R2 (active rows) -> out: s1 in v, s2 in R2 (in place)
  __device__ void polish_reduced_solve(const TV<T>& R2, T dinv)
  {
    for (int i = 0; i < m; ++i) {
      if (w[i] == T(0)) continue;
      const T sc = R2[i] * dinv;
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) v[S.A_col[e]] += A[e] * sc;
    }
    solve();
    for (int i = 0; i < m; ++i) {
      if (w[i] == T(0)) continue;
      T as = T(0);
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) as += A[e] * v[S.A_col[e]];
      R2[i] = (as - R2[i]) * dinv;
    }
  }

  // out1 (n) = r1 - (sym(Pbar) X + shift X + Aa^T Y),  out2 (m, active rows) = r2 - (Aa X - shift Y)
  __device__ void polish_residual(const TV<T>& r1, const TV<T>& r2, const TV<T>& X, const TV<T>& Y, T shift,
                                  const TV<T>& out1, const TV<T>& out2)
  {
    for (int j = 0; j < n; ++j) out1[j] = r1[j] - shift * X[j];
    for (int e = 0; e < S.nnzP; ++e) {
      if (S.P_tgt[e] < 0) continue;
      const int rj = S.P_rowp[e], cj = S.P_colp[e];
      const T pb = ((c * sx[rj]) * sx[cj]) * P[e];
      out1[rj] -= pb * X[cj];
      if (rj != cj) out1[cj] -= pb * X[rj];
    }
    for (int i = 0; i < m; ++i) {
      if (w[i] == T(0)) continue;
      T at = T(0);
      const T yi = Y[i];
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) {
        const int cj = S.A_col[e];
        at += A[e] * X[cj];
        out1[cj] -= A[e] * yi;
      }
      out2[i] = r2[i] - (at - shift * yi);
    }
  }

  __device__ unsigned polish(const sfb_qp_params& prm)
  {
    const T delta = T(prm.delta), dinv = T(1) / delta;
    for (int i = 0; i < m; ++i) rho[i] = (w[i] != T(0)) ? dinv : T(0);
    assemble(delta, rho);
    if (!factor()) return SFB_QP_FLAG_POLISH_FAILED;
    // t = (t1 [n], rinv [m]);  h = (-qb, bnd)
    for (int j = 0; j < n; ++j) { t1[j] = T(0); xold[j] = -qb[j]; }
    for (int i = 0; i < m; ++i) rinv[i] = T(0);
    for (unsigned it = 0; it < prm.polish_iter; ++it) {
      // r = h - H t  -> (t2, z)        (H = Hp without the delta blocks: shift 0)
      polish_residual(xold, yold, t1, rinv, T(0), t2, z);
      // s = Hp^-1 r -> (t3, l), refined twice against Hp
      for (int j = 0; j < n; ++j) v[j] = t2[j];
      for (int i = 0; i < m; ++i) l[i] = z[i];
      polish_reduced_solve(l, dinv);
      for (int j = 0; j < n; ++j) t3[j] = v[j];
      for (int rf = 0; rf < 2; ++rf) {
        // Hp s = [ (Pbar + delta I) s1 + Aa^T s2 ; Aa s1 - delta s2 ]
        for (int i = 0; i < m; ++i) u[i] = T(0);
        polish_residual_hp(delta);
        polish_reduced_solve(u, dinv);
        for (int j = 0; j < n; ++j) t3[j] += v[j];
        for (int i = 0; i < m; ++i)
          if (w[i] != T(0)) l[i] += u[i];
      }
      for (int j = 0; j < n; ++j) t1[j] += t3[j];
      for (int i = 0; i < m; ++i)
        if (w[i] != T(0)) rinv[i] += l[i];
    }
    bool finite = true;
    for (int j = 0; j < n; ++j) finite = finite && (fabs(t1[j]) < Num<T>::inf());
    if (!finite) return SFB_QP_FLAG_POLISH_FAILED;
    for (int j = 0; j < n; ++j) x[j] = t1[j];  // :199
    for (int i = 0; i < m; ++i)
      if (w[i] != T(0)) y[i] = rinv[i];  // :200-201 (other duals unchanged)
    return SFB_QP_FLAG_POLISHED;
  }

  // (v, u) <- (t2, z) - Hp (t3, l):  first block has +delta on the diagonal, second block -delta
  __device__ void polish_residual_hp(T delta)
  {
    for (int j = 0; j < n; ++j) v[j] = t2[j] - delta * t3[j];
    for (int e = 0; e < S.nnzP; ++e) {
      if (S.P_tgt[e] < 0) continue;
      const int rj = S.P_rowp[e], cj = S.P_colp[e];
      const T pb = ((c * sx[rj]) * sx[cj]) * P[e];
      v[rj] -= pb * t3[cj];
      if (rj != cj) v[cj] -= pb * t3[rj];
    }
    for (int i = 0; i < m; ++i) {
      if (w[i] == T(0)) continue;
      T at = T(0);
      const T yi = l[i];
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) {
        const int cj = S.A_col[e];
        at += A[e] * t3[cj];
        v[cj] -= A[e] * yi;
      }
      u[i] = z[i] - (at - delta * yi);
    }
  }

  // ---------------------------------------------------------------- QPSolver::solve, qp_solver.hpp:343-568
  __device__ void run(const SpArgs<T>& a, long long b)
  {
    const T inf = Num<T>::inf();
    const sfb_qp_params& prm = a.prm;
    const unsigned long long t0 = prm.has_max_time ? global_timer_ns() : 0ull;
    // ---- ingest (array-of-instances -> tile; each lane streams its own rows, lines are reused through L1)
    {
      const T* gA = a.A + b * (long long)S.nnzA;
      const T* gP = a.P + b * (long long)S.nnzP;
      for (int e = 0; e < S.nnzA; ++e) A[e] = __ldg(gA + e);
      for (int e = 0; e < S.nnzP; ++e) P[e] = __ldg(gP + e);
      for (int j = 0; j < n; ++j) q[S.iperm[j]] = __ldg(a.q + b * (long long)n + j);
      for (int i = 0; i < m; ++i) {
        l[i] = __ldg(a.l + b * (long long)m + i);
        u[i] = __ldg(a.u + b * (long long)m + i);
      }
    }
    if (prm.scaling) scale();  // :347
    else {
      c = T(1);
      for (int j = 0; j < n; ++j) sx[j] = T(1);
      for (int i = 0; i < m; ++i) sy[i] = T(1);
    }
    // ---- rho classes + trivially empty feasible set  :361-374
    int code = kStatusUnset;
    const T rho_bar = T(prm.rho), sigma = T(prm.sigma), alpha = T(prm.alpha), alpha_comp = T(1) - alpha;
    for (int i = 0; i < m; ++i) {
      const T li = l[i], ui = u[i];
      if (li == inf || ui == -inf || ui - li < T(0)) code = SFB_QP_PRIMAL_INFEASIBLE;
      T r;
      if (li == -inf && ui == inf) r = T(1e-6);
      else if (sy[i] * fabs(li - ui) < T(1e-5)) r = T(1e3) * rho_bar;
      else r = rho_bar;
      rho[i] = r;
      rinv[i] = T(1) / r;
    }
    // ---- scaled data: qb = c Sx q, Abar = Sy A Sx  (:401-403, :450)
    for (int j = 0; j < n; ++j) qb[j] = (c * sx[j]) * q[j];
    for (int i = 0; i < m; ++i) {
      const T syi = sy[i];
      for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) A[e] = (syi * sx[S.A_col[e]]) * A[e];
    }
    assemble(sigma, rho);
    if (!factor()) code = SFB_QP_UNKNOWN;  // :433
    // ---- initial iterate  :436-445
    if (a.warm_x != nullptr) {
      for (int j = 0; j < n; ++j) {
        const int pj = S.iperm[j];
        x[pj] = (T(1) / sx[pj]) * __ldg(a.warm_x + b * (long long)n + j);
      }
      for (int i = 0; i < m; ++i) {
        y[i] = c * ((T(1) / sy[i]) * __ldg(a.warm_y + b * (long long)m + i));
        T zt = T(0);
        for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) zt += A[e] * x[S.A_col[e]];
        z[i] = zt;
      }
    } else {
      for (int j = 0; j < n; ++j) x[j] = T(0);
      for (int i = 0; i < m; ++i) { y[i] = T(0); z[i] = T(0); }
    }
    for (int i = 0; i < m; ++i) w[i] = rho[i] * z[i] - y[i];

    // ---- main loop  :449-510
    const unsigned sci = prm.stop_check_iter;
    unsigned iter = 0;
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      for (int j = 0; j < n; ++j) v[j] = T(0);
      At_acc(w, v);
      for (int j = 0; j < n; ++j) v[j] = sigma * x[j] - qb[j] + v[j];
      solve();
      const bool chk = (iter % sci == 1u);
      for (int j = 0; j < n; ++j) {
        const T xi = x[j];
        if (chk) xold[j] = xi;  // :465-468
        x[j] = alpha * v[j] + alpha_comp * xi;  // :470
      }
      for (int i = 0; i < m; ++i) {
        T zt = T(0);
#pragma unroll 4
        for (int e = S.A_rowptr[i]; e < S.A_rowptr[i + 1]; ++e) zt += A[e] * v[S.A_col[e]];
        const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];
        if (chk) yold[i] = yi;
        const T nu = ri * (zt - zi) + yi;
        T vv = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;  // :471-474
        vv = fmax(vv, sy[i] * l[i]);
        vv = fmin(vv, sy[i] * u[i]);
        const T yn = alpha_comp * yi + alpha * nu + ri * zi - ri * vv;  // :475-477
        y[i] = yn;
        z[i] = vv;
        w[i] = ri * vv - yn;
      }
      if (chk) {
        code = check_stopping(prm);  // :488 (clobbers w, v, t1..t4, xold)
        if (code == kStatusUnset && prm.has_max_time && (long long)(global_timer_ns() - t0) > prm.max_time_ns)
          code = SFB_QP_MAX_TIME;  // :504-508
        for (int i = 0; i < m; ++i) w[i] = rho[i] * z[i] - y[i];
      }
    }

    // ---- active sets as polish_qp builds them (:113-123) on the scaled dual
    const T thr = T(100) * Num<T>::eps();
    int na = 0;
    for (int i = 0; i < m; ++i) {
      int act = 0;
      T bv = T(0);
      if (y[i] < -thr && l[i] != -inf) { act = -1; bv = sy[i] * l[i]; }
      if (y[i] > thr && u[i] != inf) { act = 1; bv = sy[i] * u[i]; }
      if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
      w[i] = (act != 0) ? T(1) : T(0);
      yold[i] = bv;
      na += (act != 0);
    }
    unsigned flags = 0;
    if (code == SFB_QP_OPTIMAL && prm.polish) {
      if (sizeof(T) == 4) flags = SFB_QP_FLAG_POLISH_SKIPPED;  // delta = 1e-6 is not resolvable in fp32 (as in the dense kernel)
      else flags = polish(prm);
    }
    // ---- unscale + objective  :544-548
    for (int j = 0; j < n; ++j) {
      t1[j] = sx[j] * x[j];
      t3[j] = T(0);
    }
    for (int e = 0; e < S.nnzP; ++e) t3[S.P_rowp[e]] += (T(0.5) * P[e]) * t1[S.P_colp[e]];
    T obj = T(0);
    for (int jo = 0; jo < n; ++jo) {
      const int j = S.iperm[jo];
      const T xv = t1[j];
      a.out_x[b * (long long)n + jo] = xv;
      obj += xv * (t3[j] + q[j]);
    }
    for (int i = 0; i < m; ++i) a.out_y[b * (long long)m + i] = sy[i] * y[i] / c;
    a.out_obj[b] = obj;
    a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
    a.out_iter[b] = iter;
    if (a.out_flags) a.out_flags[b] = flags;
    (void)na;
  }
};

// One warp per tile of 32 instances, one warp per CTA (small batches still spread over all SMs).
template <typename T> __global__ void __launch_bounds__(32) qp_sparse_tiled_kernel(const SpArgs<T> a)
{
  const int lane = threadIdx.x;
  const long long ntiles = (a.batch + 31) / 32;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b = tile * 32 + lane;
    if (b < a.batch) {
      SpSolver<T> s(a, tile, lane);
      s.run(a, b);
    }
  }
}

}  // namespace sfb
