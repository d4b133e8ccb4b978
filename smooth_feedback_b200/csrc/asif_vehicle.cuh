// asif_vehicle.cuh -- ASIFilter::operator() on the device for the built-in SE(2) x R^3 vehicle family: ONE launch maps
// (state g, desired input u_des) -> filtered input u for every agent of a fleet.
//
// Replaces (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   asif_to_qp_update      include/smooth/feedback/asif_func.hpp:104-199   (backup-trajectory + sensitivity Euler
//                                                                            integration, barrier rows, input bounds)
//   ASIFilter::operator()  include/smooth/feedback/asif.hpp:82-102         (transcribe, solve_qp, keep the warm start
//                                                                            only if Optimal :99, u = u_des (+) primal.head :101)
// for the model family of examples/mpc_asif_vehicle.cpp:42-52 (dynamics), :96-100 (safe set), :103 (backup controller).
// The reference differentiates user lambdas on the host with autodiff; here the family's derivatives are closed forms.
//
// Structure: one launch, two work queues.
//   phase 1 (a warp takes a tile of 32 agents, lane = agent): the backup trajectory x(t) and its 6 x 6 sensitivity S(t) are integrated in registers (fp64
//     whatever the QP precision) along the step schedule the host computed with the reference's exact time arithmetic
//     (asif_func.hpp:139-143,170-176 -- it does not depend on the state).  Per constraint k the only agent-dependent QP
//     entries are (A_k0, A_k1, l_k); they go to a [3][K] block per agent.
//   phase 2 (a warp takes ONE agent at a time from a queue over the whole batch; tiles are published through a ready flag):
//     rows are re-read lane-strided (coalesced), the remaining rows / P / q are
//     constants of the parameter set, and the register-resident tall-skinny ADMM solver of qp_dense_skinny.cuh runs
//     (n = 3, m = K + 3; polish off like mpc_asif_vehicle.cpp:127).  Warm starts stay resident in device memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "qp_dense_skinny.cuh"

namespace sfb {

struct AsifVehicleDev
{
  int K;
  double alpha, relax_cost;
  double w_u[2], ulim_l[2], ulim_u[2];
  double drag1, drag3, cx, cy, radius, bu_gain, bu_const;
  const int* nsteps;     // [K] number of Euler steps taken after constraint k was emitted
  const double* dt_act;  // [K] their common length
};

template <typename T> struct AsifArgs
{
  AsifVehicleDev mdl;
  sfb_qp_params prm;
  unsigned max_iter_eff;
  int dinf_guard;  // see QpArgs::dinf_guard
  long long batch;
  const T* x;      // [batch][7]  (x, y, sin, cos, v1, v2, v3): smooth::Bundle<SE2, R^3> coefficient order
  const T* u_des;  // [batch][2]
  T* rows;         // [batch][3][K] scratch: A_k0, A_k1, l_k
  T* warm_x;       // [batch][3]      resident warm start (asif.hpp:109), nullptr = never warm start
  T* warm_y;       // [batch][K + 3]
  uint8_t* warm_valid;  // [batch]
  T* out_u;        // [batch][2]
  int32_t* out_status;
  uint32_t* out_iter;
  // transcription-only mode (asif_to_qp, asif_func.hpp:246-261): dense QP in the reference's column-major storage
  T *qp_P, *qp_q, *qp_A, *qp_l, *qp_u;
  unsigned long long* work_counter;   // tiles of 32 agents to transcribe
  unsigned long long* solve_counter;  // agents to solve
  unsigned* tile_ready;               // [ceil(batch / 32)] 0 until the tile's rows are in memory (zeroed before every launch)
};

// phase 1 for one agent.  All arithmetic in double, in the operation order of oracle/transcribe.py::asif_to_qp.
template <typename T>
__device__ __noinline__ void asif_vehicle_rows(const AsifVehicleDev& M, const T* __restrict__ xin, const T* __restrict__ ud, T* __restrict__ rows)
{
  double px = (double)xin[0], py = (double)xin[1], sn = (double)xin[2], cs = (double)xin[3];
  double v1 = (double)xin[4], v2 = (double)xin[5], v3 = (double)xin[6];
  const double u0 = (double)ud[0], u1 = (double)ud[1];
  const double f0[6] = {v1, v2, v3, -M.drag1 * v1 + u0, 0.0, -M.drag3 * v3 + u1};  // f(x0, u_des)   :146-147
  const double j33 = -M.drag1 + M.bu_gain;  // d/dv1 [-drag1 v1 + bu_gain v1], summed like the autodiff chain rule
  double S[6][6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 6; ++j) S[i][j] = (i == j) ? 1.0 : 0.0;
  const int K = M.K;
#pragma unroll 1
  for (int k = 0; k < K; ++k) {
    // barrier function and its right derivative  :152-156
    const double dx = px - M.cx, dy = py - M.cy;
    const double nrm = sqrt(dx * dx + dy * dy);
    const double e0 = dx / nrm, e1 = dy / nrm;
    const double hval = (dx * e0 + dy * e1) - M.radius;
    const double g0 = e0 * cs + e1 * sn, g1 = -e0 * sn + e1 * cs;  // e^T R
    double d[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) d[j] = g0 * S[0][j] + g1 * S[1][j];  // dh_dx0 = dh_dx * dx_dx0   :159
    double dot = 0.0;
#pragma unroll
    for (int j = 0; j < 6; ++j) dot += d[j] * f0[j];
    rows[k] = (T)d[3];                                   // dh_dx0 * d_f0_du: unit columns 3 and 5   :160
    rows[K + k] = (T)d[5];
    rows[2 * K + k] = (T)((-0.0 - M.alpha * hval) - dot);  // -dh_dt - alpha h - dh_dx0 f0   :161
    const int ns = M.nsteps[k];
    const double dt = M.dt_act[k];
#pragma unroll 1
    for (int s = 0; s < ns; ++s) {
      // state stepper: x <- x (+) dt f(x, bu(x))   :173
      {
        const double fc3 = -M.drag1 * v1 + M.bu_gain * v1, fc5 = -M.drag3 * v3 + M.bu_const;
        const double a = dt * v1, b = dt * v2, w = dt * v3;
        double sw, cw;
        sincos(w, &sw, &cw);
        double A_, B_;
        if (fabs(w) < 1e-9) { A_ = 1.0 - w * w / 6.0; B_ = w / 2.0 - w * w * w / 24.0; }
        else { A_ = sw / w; B_ = (1.0 - cw) / w; }
        const double ex = A_ * a - B_ * b, ey = B_ * a + A_ * b;
        const double npx = px + cs * ex - sn * ey, npy = py + sn * ex + cs * ey;
        const double nsn = sn * cw + cs * sw, ncs = cs * cw - sn * sw;
        px = npx; py = npy; sn = nsn; cs = ncs;
        v1 = v1 + dt * fc3; v2 = v2 + dt * 0.0; v3 = v3 + dt * fc5;
      }
      // sensitivity stepper, linearised at the ALREADY-STEPPED state (the reference's lambda captures x by reference)   :130-134,174
      double dS[6][6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        dS[0][j] = (v3 * S[1][j] + (-v2) * S[2][j]) + S[3][j];
        dS[1][j] = ((-v3) * S[0][j] + v1 * S[2][j]) + S[4][j];
        dS[2][j] = S[5][j];
        dS[3][j] = j33 * S[3][j];
        dS[4][j] = 0.0;
        dS[5][j] = -M.drag3 * S[5][j];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) S[i][j] = S[i][j] + dt * dS[i][j];
    }
  }
}

template <typename T, int R>
__global__ void __launch_bounds__(32 * kSkinnyWarps, sizeof(T) == 4 ? 3 : 2) asif_vehicle_filter_kernel(const __grid_constant__ AsifArgs<T> a)
{
  const int lane = threadIdx.x & 31;
  const AsifVehicleDev& M = a.mdl;
  const int K = M.K, m = K + 3;
  const long long ntiles = (a.batch + 31) / 32;
  // ---- phase 1: warps pull tiles of 32 agents (lane = agent) until none is left, publishing each finished tile
#pragma unroll 1
  for (;;) {
    unsigned long long tile = 0;
    if (lane == 0) tile = atomicAdd(a.work_counter, 1ull);
    tile = __shfl_sync(kFullMask, tile, 0);
    if ((long long)tile >= ntiles) break;
    const long long b = (long long)tile * 32 + lane;
    if (b < a.batch) asif_vehicle_rows<T>(M, a.x + b * 7, a.u_des + b * 2, a.rows + b * 3 * (long long)K);
    __threadfence();  // every lane's rows are visible device-wide before the tile is published
    __syncwarp();
    if (lane == 0) atomicExch(a.tile_ready + tile, 1u);
  }
  // ---- phase 2: warps pull single agents from ONE queue over the whole batch (iteration counts are heavy-tailed: median 2,
  // mean ~300, max > 2500 on the vehicle workload), whichever tile they came from.  A warp gets here only after every tile has
  // been claimed by a running warp, and a warp transcribing a tile never waits, so the spin below always terminates.
#pragma unroll 1
  for (;;) {
    unsigned long long bb = 0;
    if (lane == 0) {
      bb = atomicAdd(a.solve_counter, 1ull);
      if ((long long)bb < a.batch) {
        const volatile unsigned* flag = a.tile_ready + (bb >> 5);
        while (*flag == 0u) {}
        __threadfence();
      }
    }
    bb = __shfl_sync(kFullMask, bb, 0);
    if ((long long)bb >= a.batch) break;
    {
      const long long b = (long long)bb;
      QpSkinny<T, 3, R> s;
      s.lane = lane;
      s.m = m;
      s.dinf_guard = a.dinf_guard != 0;
      const T inf = Num<T>::inf();
      const T* rw = a.rows + b * 3 * (long long)K;
      const double ud0 = (double)a.u_des[b * 2], ud1 = (double)a.u_des[b * 2 + 1];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        const int i = lane + 32 * k;
        s.valid[k] = i < m;
        T a0 = T(0), a1 = T(0), a2 = T(0), lo = -inf, hi = inf;
        if (i < K) { a0 = __ldcg(rw + i); a1 = __ldcg(rw + K + i); a2 = T(1); lo = __ldcg(rw + 2 * K + i); }                       // :160-162,180
        else if (i == K) { a0 = T(1); lo = (T)(M.ulim_l[0] - ud0); hi = (T)(M.ulim_u[0] - ud0); }          // :183-185 (ulim.A = I, c = 0)
        else if (i == K + 1) { a1 = T(1); lo = (T)(M.ulim_l[1] - ud1); hi = (T)(M.ulim_u[1] - ud1); }
        else if (i == K + 2) { a2 = T(1); lo = T(0); }                                                     // :188-190
        s.A[k][0] = a0; s.A[k][1] = a1; s.A[k][2] = a2;
        s.l[k] = lo; s.u[k] = hi;
      }
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        s.q[i] = T(0);
#pragma unroll
        for (int j = 0; j < 3; ++j) s.P[i][j] = T(0);
      }
      s.P[0][0] = (T)M.w_u[0]; s.P[1][1] = (T)M.w_u[1]; s.P[2][2] = (T)M.relax_cost;                       // :192-195

      if (a.qp_A != nullptr) {  // asif_to_qp: emit the QP instead of solving it
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int i = lane + 32 * k;
          if (i >= m) continue;
#pragma unroll
          for (int j = 0; j < 3; ++j) a.qp_A[b * 3 * (long long)m + i + (long long)m * j] = s.A[k][j];
          a.qp_l[b * (long long)m + i] = s.l[k];
          a.qp_u[b * (long long)m + i] = s.u[k];
        }
        if (lane < 9) {
          const int pi = lane % 3, pj = lane / 3;
          a.qp_P[b * 9 + lane] = (pi != pj) ? T(0) : (pi == 0 ? (T)M.w_u[0] : (pi == 1 ? (T)M.w_u[1] : (T)M.relax_cost));
        }
        if (lane < 3) a.qp_q[b * 3 + lane] = T(0);
        continue;
      }

      const unsigned long long t0 = a.prm.has_max_time ? global_timer_ns() : 0ull;
      const bool warm = a.warm_x != nullptr && a.warm_valid[b] != 0;
      int code;
      unsigned iter;
      s.run(a.prm, a.max_iter_eff, t0, warm ? a.warm_x + b * 3 : nullptr, warm ? a.warm_y + b * (long long)m : nullptr, code, iter);
      const bool optimal = code == SFB_QP_OPTIMAL;
      if (optimal && a.warm_x != nullptr) {  // asif.hpp:99: the solution replaces the warm start only if Optimal
#pragma unroll
        for (int k = 0; k < R; ++k) {
          const int i = lane + 32 * k;
          if (i < m) a.warm_y[b * (long long)m + i] = s.sy[k] * s.y[k] / s.c;
        }
      }
      if (lane == 0) {
        T xus[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) xus[j] = s.sx[j] * s.x[j];
        a.out_u[b * 2] = (T)ud0 + xus[0];       // rplus(u_des, primal.head<2>)   asif.hpp:101
        a.out_u[b * 2 + 1] = (T)ud1 + xus[1];
        a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
        a.out_iter[b] = iter;
        if (optimal && a.warm_x != nullptr) {
#pragma unroll
          for (int j = 0; j < 3; ++j) a.warm_x[b * 3 + j] = xus[j];
          a.warm_valid[b] = 1;
        }
      }
    }
    __syncwarp();
  }
}

}  // namespace sfb
