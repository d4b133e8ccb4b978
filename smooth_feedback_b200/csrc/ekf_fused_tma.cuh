// ekf_fused_tma.cuh -- fused, size-specialised EKF covariance step for sm_100a: predict and/or update in ONE pass
// over HBM, operands staged by the TMA engine (1-D bulk copies, cp.async.bulk + mbarrier), algebra in registers.
//
// Replaces (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   EKF::predict, covariance half   include/smooth/feedback/ekf.hpp:79-103   (boost euler stepper, :147)
//   EKF::update                     include/smooth/feedback/ekf.hpp:116-139
//
// Why this shape.  A d = 6, ny = 3 filter moves 1104 B in and 336 B out for ~1.5 kflop: the step is HBM-bound by
// two orders of magnitude, so the only thing that matters is keeping bulk transfers in flight:
//   * the batch is in the reference's array-of-matrices layout, so a tile of TILE consecutive instances is ONE
//     contiguous block per field -> six 1-D bulk copies per tile, issued by one thread, completion counted on an
//     mbarrier (complete_tx).  No per-thread address arithmetic, no LDG instruction stream.
//   * a thread owns one instance: it pulls its operands from the staged tile into registers (LDS.128, rows are
//     16-byte aligned), the CTA barriers, and the tile buffer is immediately re-armed with the CTA's next tile, so
//     the next transfer overlaps the register-resident algebra of this one (2 CTAs per SM -> ~140 KB in flight per
//     SM, far above the ~45 KB Little's-law requirement at 6.5 TB/s).
//   * results go to a staging tile and leave with bulk stores (cp.async.bulk.global.shared::cta, bulk_group).
//   * the pivoted ny x ny LDL^T (Eigen LDLT semantics, dynamic pivot indices) runs out of a per-thread strided
//     shared-memory scratch: dynamic register indexing would spill to local memory.
// Arithmetic order follows oracle/sf_oracle.cpp::{cov_step, ekf_update_one} term by term.

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "ekf_kernels.cuh"

namespace sfb {

struct EkfStepArgs
{
  const double* P;      // [batch][d*d]
  const double* A;      // [batch][d*d]   (predict)
  const double* Q;      // [batch][d*d]   (predict)
  const double* H;      // [batch][ny*d]  (update)
  const double* R;      // [batch][ny*ny] (update)
  const double* innov;  // [batch][ny]    (update)
  double* out_delta;    // [batch][d]     (update)
  double* out_P;        // [batch][d*d]
  long long batch;
  double tau, dt;
};

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
// global -> shared, completion signalled on the mbarrier (bytes % 16 == 0, both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace tma

template <int D, int NY, bool PRED, bool UPD, int TILE> struct EkfFusedLayout
{
  static constexpr int DD = D * D, ND = NY * D, NN = NY * NY;
  static constexpr int ev(int x) { return (x + 1) & ~1; }  // every block starts 16-byte aligned
  static constexpr int oP = 0;
  static constexpr int oA = oP + TILE * DD;
  static constexpr int oQ = oA + (PRED ? TILE * DD : 0);
  static constexpr int oH = oQ + (PRED ? TILE * DD : 0);
  static constexpr int oR = oH + (UPD ? ev(TILE * ND) : 0);
  static constexpr int oI = oR + (UPD ? ev(TILE * NN) : 0);
  static constexpr int oOutP = oI + (UPD ? ev(TILE * NY) : 0);
  static constexpr int oOutD = oOutP + TILE * DD;
  static constexpr int oScr = oOutD + (UPD ? ev(TILE * D) : 0);
  static constexpr int scr_elems = UPD ? (NN + NY + ND) : 0;  // S, tmp, Kt   per thread, stride TILE+1
  static constexpr int oBar = oScr + ev(scr_elems * (TILE + 1));
  static constexpr int total = oBar + 2;
  static constexpr size_t bytes = sizeof(double) * (size_t)total;
  static constexpr uint32_t in_bytes_per_inst = 8u * (DD + (PRED ? 2 * DD : 0) + (UPD ? ND + NN + NY : 0));
};

// copy `len` doubles of this thread's row out of the staged tile (16-byte vector loads; rows of even length
// start 16-byte aligned for every thread, odd-length rows are read scalar)
template <int LEN> __device__ __forceinline__ void row_to_regs(const double* tile, int t, double (&r)[LEN])
{
  const double* p = tile + (size_t)t * LEN;
  if constexpr (LEN % 2 == 0) {
#pragma unroll
    for (int e = 0; e < LEN; e += 2) {
      const double2 v = *reinterpret_cast<const double2*>(p + e);
      r[e] = v.x;
      r[e + 1] = v.y;
    }
  } else {
#pragma unroll
    for (int e = 0; e < LEN; ++e) r[e] = p[e];
  }
}
template <int LEN> __device__ __forceinline__ void regs_to_row(double* tile, int t, const double (&r)[LEN])
{
  double* p = tile + (size_t)t * LEN;
  if constexpr (LEN % 2 == 0) {
#pragma unroll
    for (int e = 0; e < LEN; e += 2) *reinterpret_cast<double2*>(p + e) = make_double2(r[e], r[e + 1]);
  } else {
#pragma unroll
    for (int e = 0; e < LEN; ++e) p[e] = r[e];
  }
}

// cooperative plain copy for the ragged last tile (cnt < TILE: bulk-copy sizes would not be multiples of 16 bytes)
__device__ __forceinline__ void plain_copy(double* dst, const double* src, int count)
{
  for (int k = threadIdx.x; k < count; k += blockDim.x) dst[k] = src[k];
}

// ldlt_small (ekf_kernels.cuh) with a compile-time size: the pivot loop is unrolled so tr[] stays in registers.
template <int SZ, typename T>
__device__ __forceinline__ void ldlt_small_static(const SmMat<T>& W, int (&tr)[SZ], const SmMat<T>& tmp)
{
  if constexpr (SZ <= 1) {
    tr[0] = 0;
  } else {
#pragma unroll
  for (int kk = 0; kk < SZ; ++kk) {
    int big = kk;
    T bigv = fabs(W(kk, kk));
#pragma unroll
    for (int j = kk + 1; j < SZ; ++j) {
      const T v = fabs(W(j, j));
      if (v > bigv) { bigv = v; big = j; }
    }
    tr[kk] = big;
    if (kk != big) {
      const int s = SZ - big - 1;
      for (int j = 0; j < kk; ++j) { const T t = W(kk, j); W(kk, j) = W(big, j); W(big, j) = t; }
      for (int i = 0; i < s; ++i) { const T t = W(big + 1 + i, kk); W(big + 1 + i, kk) = W(big + 1 + i, big); W(big + 1 + i, big) = t; }
      { const T t = W(kk, kk); W(kk, kk) = W(big, big); W(big, big) = t; }
      for (int i = kk + 1; i < big; ++i) { const T t = W(i, kk); W(i, kk) = W(big, i); W(big, i) = t; }
    }
    if (kk > 0) {
#pragma unroll
      for (int j = 0; j < kk; ++j) tmp[j] = W(j, j) * W(kk, j);
      T acc = T(0);
#pragma unroll
      for (int j = 0; j < kk; ++j) acc += W(kk, j) * tmp[j];
      W(kk, kk) -= acc;
#pragma unroll
      for (int i = kk + 1; i < SZ; ++i) {
        T a2 = T(0);
#pragma unroll
        for (int j = 0; j < kk; ++j) a2 += W(i, j) * tmp[j];
        W(i, kk) -= a2;
      }
    }
    const T akk = W(kk, kk);
    if (fabs(akk) > T(0)) {
#pragma unroll
      for (int i = kk + 1; i < SZ; ++i) W(i, kk) /= akk;
    }
  }
  }
}

template <int D, int NY, bool PRED, bool UPD, int TILE>
__global__ void __launch_bounds__(TILE, 2) ekf_fused_tma_kernel(const EkfStepArgs a)
{
  using L = EkfFusedLayout<D, NY, PRED, UPD, TILE>;
  constexpr int DD = L::DD, ND = L::ND, NN = L::NN;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* sm = reinterpret_cast<double*>(smem_raw);
  double* sP = sm + L::oP;
  double* sA = sm + L::oA;
  double* sQ = sm + L::oQ;
  double* sH = sm + L::oH;
  double* sR = sm + L::oR;
  double* sI = sm + L::oI;
  double* sOutP = sm + L::oOutP;
  double* sOutD = sm + L::oOutD;
  double* sScr = sm + L::oScr;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + L::oBar);
  const int t = threadIdx.x;
  const long long full_tiles = a.batch / TILE;
  const long long ntiles = (a.batch + TILE - 1) / TILE;

  if (t == 0) {
    tma::mbar_init(bar, 1);
    tma::fence_mbar_init();
  }
  __syncthreads();

  auto issue_loads = [&](long long tile) {  // one thread
    const long long b0 = tile * TILE;
    tma::mbar_expect_tx(bar, L::in_bytes_per_inst * (uint32_t)TILE);
    tma::bulk_g2s(sP, a.P + b0 * DD, 8u * DD * TILE, bar);
    if constexpr (PRED) {
      tma::bulk_g2s(sA, a.A + b0 * DD, 8u * DD * TILE, bar);
      tma::bulk_g2s(sQ, a.Q + b0 * DD, 8u * DD * TILE, bar);
    }
    if constexpr (UPD) {
      tma::bulk_g2s(sH, a.H + b0 * ND, 8u * ND * TILE, bar);
      tma::bulk_g2s(sR, a.R + b0 * NN, 8u * NN * TILE, bar);
      tma::bulk_g2s(sI, a.innov + b0 * NY, 8u * NY * TILE, bar);
    }
  };

  uint32_t phase = 0;
  if (t == 0 && (long long)blockIdx.x < full_tiles) issue_loads(blockIdx.x);

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * TILE;
    const bool full = tile < full_tiles;
    const int cnt = full ? TILE : (int)(a.batch - b0);
    if (full) {
      tma::mbar_wait(bar, phase);
      phase ^= 1u;
    } else {
      plain_copy(sP, a.P + b0 * DD, cnt * DD);
      if constexpr (PRED) {
        plain_copy(sA, a.A + b0 * DD, cnt * DD);
        plain_copy(sQ, a.Q + b0 * DD, cnt * DD);
      }
      if constexpr (UPD) {
        plain_copy(sH, a.H + b0 * ND, cnt * ND);
        plain_copy(sR, a.R + b0 * NN, cnt * NN);
        plain_copy(sI, a.innov + b0 * NY, cnt * NY);
      }
      __syncthreads();
    }

    // ---- operands to registers -------------------------------------------------------------------------
    double P[DD];
    row_to_regs<DD>(sP, t, P);
    if constexpr (PRED) {
      double A[DD];
      row_to_regs<DD>(sA, t, A);
      // step schedule of ekf.hpp:91-102 (dt <= 0 -> default 2 tau -> exactly one step of length tau)
      const double dt_v = (a.dt > 0.0) ? a.dt : 2.0 * a.tau;
      double tt = 0.0;
      bool last = false;
      do {
        double h;
        if (tt + dt_v < a.tau) { h = dt_v; } else { h = a.tau - tt; last = true; }
        tt += dt_v;
        // dcov = symU(A P + P A^T + Q), P += h dcov  (euler, ekf.hpp:84-89,147); upper entry mirrored
        double K[DD];
#pragma unroll
        for (int j = 0; j < D; ++j) {
#pragma unroll
          for (int i = 0; i <= j; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < D; ++k) acc += A[i + D * k] * P[k + D * j];
#pragma unroll
            for (int k = 0; k < D; ++k) acc += P[i + D * k] * A[j + D * k];
            acc += sQ[(size_t)t * DD + i + D * j];
            K[i + D * j] = acc;
          }
        }
#pragma unroll
        for (int j = 0; j < D; ++j) {
#pragma unroll
          for (int i = 0; i <= j; ++i) {
            const double k_ij = K[i + D * j];
            P[i + D * j] = P[i + D * j] + h * k_ij;
            if (i != j) P[j + D * i] = P[j + D * i] + h * k_ij;
          }
        }
      } while (!last);
    }

    double H[UPD ? ND : 1];
    double inn[UPD ? NY : 1];
    if constexpr (UPD) {
      row_to_regs<ND>(sH, t, H);
      row_to_regs<NY>(sI, t, inn);
    }
    // R (triu only) goes straight into the per-thread LDL^T scratch below; read it before the tile is re-armed
    const int stride = TILE + 1;
    const SmMat<double> S{sScr + t, stride, NY};
    const SmMat<double> Tmp{sScr + (size_t)NN * stride + t, stride, NY};
    const SmMat<double> Kt{sScr + (size_t)(NN + NY) * stride + t, stride, NY};
    if constexpr (UPD) {
#pragma unroll
      for (int j = 0; j < NY; ++j)
#pragma unroll
        for (int i = j; i < NY; ++i) S(i, j) = sR[(size_t)t * NN + j + NY * i];  // upper entry R(j,i) -> lower slot (i,j)
    }

    // every thread has consumed the staged tile: drain the previous bulk store (it reads sOut*), re-arm the loads
    if (t == 0) tma::bulk_wait_read0();
    __syncthreads();
    if (t == 0) {
      const long long next = tile + gridDim.x;
      if (next < full_tiles) issue_loads(next);
    }

    if constexpr (UPD) {
      // HPs = H symU(P), HP = H P     ekf.hpp:129-130,134
      double HPs[ND], HP[ND];
#pragma unroll
      for (int i = 0; i < NY; ++i)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s1 = 0.0, s2 = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) {
            const double hik = H[i + NY * k];
            s1 += hik * ((k <= j) ? P[k + D * j] : P[j + D * k]);
            s2 += hik * P[k + D * j];
          }
          HPs[i + NY * j] = s1;
          HP[i + NY * j] = s2;
        }
      // S = triu(HPs H^T + R), kept as the lower triangle of S^T
#pragma unroll
      for (int j = 0; j < NY; ++j)
#pragma unroll
        for (int i = j; i < NY; ++i) {
          double s1 = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s1 += HPs[j + NY * k] * H[i + NY * k];
          S(i, j) = s1 + S(i, j);
        }
      int tr[NY];
      ldlt_small_static<NY>(S, tr, Tmp);
      // K^T = S^-1 (H P)   :133-134.  The factor moves to registers; only the row permutation (dynamic pivot
      // indices) goes through the scratch.  1/D is formed once per pivot (Eigen divides each right-hand side;
      // the difference is one rounding, far below the parity tolerance) with the same pseudo-inverse rule.
      double Lr[NN], invD[NY];
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        const double dgl = S(i, i);
        invD[i] = (fabs(dgl) > 2.2250738585072014e-308) ? 1.0 / dgl : 0.0;
#pragma unroll
        for (int j = 0; j < i; ++j) Lr[i + NY * j] = S(i, j);
      }
      // composite of the transpositions, 4 bits per row in one word (dynamic swaps without an indexable array):
      // permuted row i is original row (pw >> 4 i) & 15
      static_assert(NY <= 8, "row permutation is packed 4 bits per row into 32 bits");
      uint32_t pw = 0x76543210u;
#pragma unroll
      for (int i = 0; i < NY; ++i) {
        const int k = tr[i];
        const uint32_t x = ((pw >> (4 * i)) ^ (pw >> (4 * k))) & 15u;
        pw ^= (x << (4 * i)) ^ (x << (4 * k));  // k == i: both terms cancel
      }
      int perm[NY];
#pragma unroll
      for (int i = 0; i < NY; ++i) perm[i] = (int)((pw >> (4 * i)) & 15u);
#pragma unroll
      for (int e = 0; e < ND; ++e) Kt[e] = HP[e];
      double K[ND];  // K^T
#pragma unroll
      for (int j = 0; j < D; ++j) {
        double b[NY];
#pragma unroll
        for (int i = 0; i < NY; ++i) b[i] = Kt(perm[i], j);
#pragma unroll
        for (int c = 0; c < NY; ++c)
#pragma unroll
          for (int i = c + 1; i < NY; ++i) b[i] -= Lr[i + NY * c] * b[c];
#pragma unroll
        for (int i = 0; i < NY; ++i) b[i] *= invD[i];
#pragma unroll
        for (int c = NY - 1; c >= 0; --c) {
          double acc = b[c];
#pragma unroll
          for (int i = c + 1; i < NY; ++i) acc -= Lr[i + NY * c] * b[i];
          b[c] = acc;
        }
#pragma unroll
        for (int i = 0; i < NY; ++i) Kt(perm[i], j) = b[i];
      }
#pragma unroll
      for (int e = 0; e < ND; ++e) K[e] = Kt[e];
      // delta = K innov   :137
      double del[D];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double s1 = 0.0;
#pragma unroll
        for (int k = 0; k < NY; ++k) s1 += K[k + NY * i] * inn[k];
        del[i] = s1;
      }
      regs_to_row<D>(sOutD, t, del);
      // P = symU((I - K H) P)   :138
      double out[DD];
#pragma unroll
      for (int i = 0; i < D; ++i) {
        double ikh[D];  // row i of I - K H
#pragma unroll
        for (int j = 0; j < D; ++j) {
          double s1 = 0.0;
#pragma unroll
          for (int k = 0; k < NY; ++k) s1 += K[k + NY * i] * H[k + NY * j];
          ikh[j] = ((i == j) ? 1.0 : 0.0) - s1;
        }
#pragma unroll
        for (int j = i; j < D; ++j) {
          double s1 = 0.0;
#pragma unroll
          for (int k = 0; k < D; ++k) s1 += ikh[k] * P[k + D * j];
          out[i + D * j] = s1;
          out[j + D * i] = s1;
        }
      }
      regs_to_row<DD>(sOutP, t, out);
    } else {
      regs_to_row<DD>(sOutP, t, P);
    }

    // ---- results leave with bulk stores ----------------------------------------------------------------
    if (full) {
      tma::fence_async_smem();
      __syncthreads();
      if (t == 0) {
        tma::bulk_s2g(a.out_P + b0 * DD, sOutP, 8u * DD * TILE);
        if constexpr (UPD) tma::bulk_s2g(a.out_delta + b0 * D, sOutD, 8u * D * TILE);
        tma::bulk_commit();
      }
    } else {
      __syncthreads();
      plain_copy(a.out_P + b0 * DD, sOutP, cnt * DD);
      if constexpr (UPD) plain_copy(a.out_delta + b0 * D, sOutD, cnt * D);
    }
  }
  if (t == 0) tma::bulk_wait0();
}

}  // namespace sfb
