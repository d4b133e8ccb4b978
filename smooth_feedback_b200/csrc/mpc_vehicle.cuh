// mpc_vehicle.cuh -- device side of MPC::operator() for the built-in SE(2) x R^3 vehicle family (SURVEY 8(f) rows f2 + f3).
//
// Replaces, per control step and agent (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   MPC::operator()          include/smooth/feedback/mpc.hpp:458-519   (update problem :473-478, transcribe :481-488,
//                                                                        solve :491, warm-start retention :510-516, u :518)
//   ocp_to_qp_update_ce      include/smooth/feedback/ocp_to_qp.hpp:326-373  with MPCCE (mpc.hpp:268-300):
//                            ce = x0 (-) x0_fix evaluated at x0 = xdes(t), Jacobian dr_expinv(ce), l = u = -ce
// The collocation / running-constraint rows and the cost are constants of the fleet (mpc_vehicle_host.hpp); this kernel
// emits each agent's value arrays straight into the layout sfb_qp_solve_sparse_batch_* ingests (the pattern order of
// QuadraticProgramSparse after makeCompressed), so nothing but (t, x) per agent ever crosses PCIe.
// Group operations of pettni/smooth (SE(2) exp / log / dr_expinv) are restated from their published closed forms, in the
// operation order of oracle/transcribe.py.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfb.h"

namespace sfb {

struct MpcVehicleDev
{
  int n, m, nnzP, nnzA, xvar_L, ce_row0;
  int ce_slot[6][6];
  double g0[3], vdes[3], udes[2];
  const double *P_vals, *A_base, *l_base, *u_base;  // fleet constants (device)
};

template <typename T> struct MpcArgs
{
  MpcVehicleDev mdl;
  long long batch;
  const T* t;  // [batch] absolute time of each agent's control step
  const T* x;  // [batch][7] (x, y, sin, cos, v1, v2, v3)
  T *P_vals, *q, *A_vals, *l, *u;  // per-agent QP values [batch][...]
};

__device__ inline void se2_exp_d(double vx, double vy, double w, double (&g)[4])
{
  double s, c;
  sincos(w, &s, &c);
  double A, B;
  if (fabs(w) < 1e-9) { A = 1.0 - w * w / 6.0; B = w / 2.0 - w * w * w / 24.0; }
  else { A = s / w; B = (1.0 - c) / w; }
  g[0] = A * vx - B * vy; g[1] = B * vx + A * vy; g[2] = s; g[3] = c;
}

// One CTA per agent: thread 0 evaluates the end constraint, all threads then stream the constant blocks.
template <typename T> __global__ void __launch_bounds__(128) mpc_vehicle_transcribe_kernel(const __grid_constant__ MpcArgs<T> a)
{
  const MpcVehicleDev& M = a.mdl;
  __shared__ double e_s[6];
  __shared__ double J_s[3][3];
  for (long long b = blockIdx.x; b < a.batch; b += gridDim.x) {
    if (threadIdx.x == 0) {
      const T* xs = a.x + b * 7;
      const double t = (double)a.t[b];
      // xl0 = xdes(t) = g0 * exp(t vdes)   (mpc.hpp:473-478 set t0; XDes::operator(), :29-33)
      double ex[4];
      se2_exp_d(t * M.vdes[0], t * M.vdes[1], t * M.vdes[2], ex);
      double s0, c0;
      sincos(M.g0[2], &s0, &c0);
      const double gx = M.g0[0] + c0 * ex[0] - s0 * ex[1], gy = M.g0[1] + s0 * ex[0] + c0 * ex[1];
      const double gs = s0 * ex[3] + c0 * ex[2], gc = c0 * ex[3] - s0 * ex[2];
      // ce = rminus(xl0, x0_fix) = log(x^-1 * xl0)   (mpc.hpp:282-285)
      const double px = (double)xs[0], py = (double)xs[1], sn = (double)xs[2], cs = (double)xs[3];
      const double ix = -(cs * px + sn * py), iy = -(-sn * px + cs * py), is = -sn, ic = cs;  // inverse of x
      const double rx = ix + ic * gx - is * gy, ry = iy + is * gx + ic * gy;
      const double rs = is * gc + ic * gs, rc = ic * gc - is * gs;
      const double w = atan2(rs, rc);
      double A;
      if (fabs(w) < 1e-9) A = 1.0 - w * w / 12.0;
      else A = 0.5 * w / tan(0.5 * w);
      const double B = w / 2.0;
      const double e0 = A * rx + B * ry, e1 = -B * rx + A * ry;
      e_s[0] = e0; e_s[1] = e1; e_s[2] = w;
      for (int k = 0; k < 3; ++k) e_s[3 + k] = M.vdes[k] - (double)xs[4 + k];
      // dr_expinv(ce) = I + ad / 2 + c2 ad^2   (mpc.hpp:287-296)
      double c2;
      if (fabs(w) < 1e-5) c2 = 1.0 / 12.0 + w * w / 720.0;
      else c2 = 1.0 / (w * w) - (1.0 + cos(w)) / (2.0 * w * sin(w));
      const double ad[3][3] = {{0.0, -w, e1}, {w, 0.0, -e0}, {0.0, 0.0, 0.0}};
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
          double sq = 0.0;
          for (int k = 0; k < 3; ++k) sq += ad[i][k] * ad[k][j];
          J_s[i][j] = (i == j ? 1.0 : 0.0) + 0.5 * ad[i][j] + c2 * sq;
        }
    }
    __syncthreads();
    T* Av = a.A_vals + b * (long long)M.nnzA;
    for (int e = threadIdx.x; e < M.nnzA; e += blockDim.x) Av[e] = (T)M.A_base[e];
    T* Pv = a.P_vals + b * (long long)M.nnzP;
    for (int e = threadIdx.x; e < M.nnzP; e += blockDim.x) Pv[e] = (T)M.P_vals[e];
    for (int j = threadIdx.x; j < M.n; j += blockDim.x) a.q[b * (long long)M.n + j] = T(0);
    for (int i = threadIdx.x; i < M.m; i += blockDim.x) {
      double lo = M.l_base[i], hi = M.u_base[i];
      if (i >= M.ce_row0) { lo = 0.0 - e_s[i - M.ce_row0]; hi = lo; }  // cel - ceval, ceu - ceval with cel = ceu = 0 (:369-370)
      a.l[b * (long long)M.m + i] = (T)lo;
      a.u[b * (long long)M.m + i] = (T)hi;
    }
    __syncthreads();  // the constant rows are in place before the agent's own entries overwrite the placeholders
    if (threadIdx.x < 36) {
      const int r = threadIdx.x / 6, c = threadIdx.x % 6;
      const int slot = M.ce_slot[r][c];
      if (slot >= 0) Av[slot] = (T)((r < 3 && c < 3) ? J_s[r][c] : (r == c ? 1.0 : 0.0));
    }
    __syncthreads();
  }
}

template <typename T> struct MpcEpilogueArgs
{
  long long batch;
  int n, m, uvar_B;
  double udes[2];
  int keep_warm;  // MPCParams::warmstart
  const T *sol_x, *sol_y;
  const int32_t* status;
  T *warm_x, *warm_y;
  uint8_t* warm_valid;
  T* out_u;
};

// u = udes(0) (+) primal.segment<Nu>(uvar_B) (mpc.hpp:518); the solution becomes the next warm start if the code is
// Optimal, MaxTime or MaxIterations (mpc.hpp:510-516)
template <typename T> __global__ void __launch_bounds__(128) mpc_vehicle_epilogue_kernel(const __grid_constant__ MpcEpilogueArgs<T> a)
{
  for (long long b = blockIdx.x; b < a.batch; b += gridDim.x) {
    const int st = a.status[b];
    if (threadIdx.x < 2) a.out_u[b * 2 + threadIdx.x] = (T)a.udes[threadIdx.x] + a.sol_x[b * (long long)a.n + a.uvar_B + threadIdx.x];
    if (a.keep_warm && (st == SFB_QP_OPTIMAL || st == SFB_QP_MAX_TIME || st == SFB_QP_MAX_ITERATIONS)) {
      for (int j = threadIdx.x; j < a.n; j += blockDim.x) a.warm_x[b * (long long)a.n + j] = a.sol_x[b * (long long)a.n + j];
      for (int i = threadIdx.x; i < a.m; i += blockDim.x) a.warm_y[b * (long long)a.m + i] = a.sol_y[b * (long long)a.m + i];
      if (threadIdx.x == 0) a.warm_valid[b] = 1;
    }
  }
}

// The optional outputs of MPC::operator() (mpc.hpp:493-507): u_traj[i] = udes(t + tf tau_i) + primal.segment<Nu>(uvar_B + i Nu),
// i < N, and x_traj[i] = xdes(t + tf tau_i) (+) primal.segment<Nx>(i Nx), i <= N, with (+) of Bundle<SE2, R^3>:
// (g * exp(a_0..2), v + a_3..5).  One thread per (agent, node); evaluated in fp64 whatever the fleet's scalar type.
template <typename T> struct MpcTrajArgs
{
  long long batch;
  int N, n, xvar_L;
  double tf, g0[3], vdes[3], udes[2];
  const double* tau;  // [N + 1]
  const T* t;         // [batch] the absolute time given to the step whose solution is read
  const T* sol_x;     // [batch][n]
  T* u_traj;          // [batch][N][2] or nullptr
  T* x_traj;          // [batch][N + 1][7] (x, y, sin, cos, v1, v2, v3) or nullptr
};

template <typename T> __global__ void __launch_bounds__(128) mpc_vehicle_traj_kernel(const __grid_constant__ MpcTrajArgs<T> a)
{
  const long long total = a.batch * (long long)(a.N + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long b = idx / (a.N + 1);
    const int i = (int)(idx - b * (a.N + 1));
    const T* sol = a.sol_x + b * (long long)a.n;
    if (a.u_traj != nullptr && i < a.N) {
      T* uo = a.u_traj + (b * a.N + i) * 2;
      uo[0] = (T)(a.udes[0] + (double)sol[a.xvar_L + 2 * i]);
      uo[1] = (T)(a.udes[1] + (double)sol[a.xvar_L + 2 * i + 1]);
    }
    if (a.x_traj != nullptr) {
      const double ta = (double)a.t[b] + a.tf * a.tau[i];
      double ex[4], es[4];
      se2_exp_d(ta * a.vdes[0], ta * a.vdes[1], ta * a.vdes[2], ex);
      double s0, c0;
      sincos(a.g0[2], &s0, &c0);
      const double gx = a.g0[0] + c0 * ex[0] - s0 * ex[1], gy = a.g0[1] + s0 * ex[0] + c0 * ex[1];
      const double gs = s0 * ex[3] + c0 * ex[2], gc = c0 * ex[3] - s0 * ex[2];
      se2_exp_d((double)sol[6 * i], (double)sol[6 * i + 1], (double)sol[6 * i + 2], es);
      T* xo = a.x_traj + (b * (a.N + 1) + i) * 7;
      xo[0] = (T)(gx + gc * es[0] - gs * es[1]);
      xo[1] = (T)(gy + gs * es[0] + gc * es[1]);
      xo[2] = (T)(gs * es[3] + gc * es[2]);
      xo[3] = (T)(gc * es[3] - gs * es[2]);
      for (int k = 0; k < 3; ++k) xo[4 + k] = (T)(a.vdes[k] + (double)sol[6 * i + 3 + k]);
    }
  }
}

}  // namespace sfb
