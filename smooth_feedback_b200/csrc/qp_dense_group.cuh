// qp_dense_group.cuh -- batched dense operator-splitting QP solver for sm_100a.
// One group of G warps (G = 1, 2 or 4; one CTA per group) owns one QP instance at a time and keeps its whole
// working set in shared memory for the lifetime of the solve.
//
// Replaces (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   QPSolver::scale           include/smooth/feedback/qp_solver.hpp:673-730
//   QPSolver::solve           include/smooth/feedback/qp_solver.hpp:343-568
//   QPSolver::check_stopping  include/smooth/feedback/qp_solver.hpp:574-644
//   detail::polish_qp         include/smooth/feedback/qp_solver.hpp:92-204
//
// Design (details and measurements in DESIGN.md):
//   * HBM sees the problem data once (coalesced load) and the solution once; everything else is on chip.
//   * The reference factorises the (n+m)x(n+m) quasi-definite KKT matrix with a pivoted LDL^T and runs two
//     triangular sweeps per ADMM iteration -- a serial dependency chain on a GPU.  This kernel eliminates the
//     diagonal (2,2) block -1/rho analytically,
//         (Pbar + sigma I + Abar^T R Abar) xt = sigma x - qbar + Abar^T (R z - y),   nu = R (Abar xt - z) + y,
//     and keeps the explicit n x n inverse Minv on chip, so that an iteration is three conflict-free GEMV
//     passes (Abar^T w, Minv rhs, Abar xt).  Mathematically identical to the KKT solve (Eigen's pivoting
//     eliminates the large |-1/rho| diagonal first as well).
//   * Abar (m x n) and Minv (n x n) are column-major with an ODD leading dimension: thread-per-row accesses are
//     consecutive, thread-per-column accesses have an odd stride -- both shared-memory bank-conflict free.
//   * G is chosen on the host so that an SM holds >= ~12 warps: the first version (one warp per QP, profiles/)
//     was shared-memory-capacity limited to 3 warps per SM and stalled 65% of the time on its own latencies.
//   * polish solves the reference's regularised KKT system by block elimination in the order Eigen's diagonal
//     pivoting takes (primal block first): Kinv = (Pbar + delta I)^-1, Sinv = (delta I + Aa Kinv Aa^T)^-1; when
//     there are more active rows than variables the equivalent Woodbury form (n x n) is used instead.

#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "../../include/sfb.h"

namespace sfb {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kStatusUnset = -1;
constexpr int kRedSlots = 12;  // scalars per batched group reduction
constexpr int kGjPad = 64;    // padded length of the pivot row / column buffers of the register-blocked inverse

template <typename T> struct Vec2T;
template <> struct Vec2T<double> { using type = double2; };
template <> struct Vec2T<float> { using type = float2; };
// two consecutive scalars in one shared-memory transaction (address must be 2-scalar aligned)
template <typename T> __device__ __forceinline__ typename Vec2T<T>::type ld2(const T* p)
{
  return *reinterpret_cast<const typename Vec2T<T>::type*>(p);
}

__device__ __forceinline__ double rcp(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float rcp(float x) { return __frcp_rn(x); }
// Branch-free reciprocal for the pivots of the register-blocked inverse: MUFU.RCP64H seed (20 bits) + Newton steps, <= 1 ulp
// for normal inputs.  __drcp_rn carries a slow-path call (BSSY / CALL / BSYNC) that splits the basic block, so ptxas
// could not overlap it with the rank-1 update; this one is plain DFMAs the scheduler interleaves with the update.
__device__ __forceinline__ double rcp_inline(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
__device__ __forceinline__ float rcp_inline(float x) { return __frcp_rn(x); }
// the value is materialised at this point of the program (an empty asm the optimiser cannot move the definition across)
__device__ __forceinline__ void keep_here(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void keep_here(float& x) { asm volatile("" : "+f"(x)); }

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

__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// T: the scalar the kernel computes in; TIO: the scalar of the caller's arrays.  TIO != T only in the mixed-precision
// polish pass (mode 2: T = double over float data, see solve()).
template <typename T, typename TIO = T> struct QpArgs
{
  const TIO* P;
  const TIO* q;
  const TIO* A;
  const TIO* l;
  const TIO* u;
  const TIO* warm_x;
  const TIO* warm_y;
  TIO* out_x;
  TIO* out_y;
  TIO* out_obj;
  int32_t* out_status;
  uint32_t* out_iter;
  int8_t* out_active;
  uint32_t* out_flags;
  // scale-only mode outputs (mode == 1)
  T* out_c;
  T* out_sx;
  T* out_sy;
  T* scratch;            // global polish workspace, scratch_per_cta scalars per CTA (may be null if 0)
  long long scratch_per_cta;
  long long batch;
  int n, m;
  int mode;  // 0 = solve, 1 = scale only, 2 = polish only (instances already solved: out_* hold the unpolished result)
  sfb_qp_params prm;
  unsigned max_iter_eff;
  int force_polish_scratch;  // debug (SFB_OPT_FORCE_POLISH_SCRATCH): keep the polish Schur block in the global workspace even when it fits on chip
  int polish_form;  // SFB_OPT_POLISH_FORM: 0 reduced form with accuracy check and schur fallback (default), 1 schur always, 2 reduced always
  int dinf_guard;  // 1: the dual-infeasibility certificate needs dx != 0 (SFB_OPT_DUAL_INF_DX_GUARD, default); 0: literal reference rule
  unsigned long long* work_counter;
};

// per-CTA shared-memory layout (units of T)
struct QpLayout
{
  int ldA, ldN, npad, mpad, npart;
  int offAs, offMs, offN, offM, offPart, offRed, offGj, offSlot, total;
  __host__ __device__ static int odd(int v) { return v | 1; }
  __host__ __device__ QpLayout(int n, int m, int nthreads)
  {
    const int mm = m > 0 ? m : 1;
    // Abar: leading dimension == 2 (mod 4).  Thread-per-row accesses are consecutive; thread-per-column accesses
    // fetch two rows per 2-scalar load and the column stride ldA/2 is odd: both patterns are bank-conflict free.
    ldA = mm + ((2 - mm % 4) + 4) % 4;
    // (The polish Schur block lives below the compacted active rows when 2 na <= ldA, else in a global workspace; both
    // placements give bit-identical results, SFB_OPT_FORCE_POLISH_SCRATCH A/B in profiles/r02_polish_scratch_vs_onchip.txt.)
    ldN = odd(n);  // Minv / P: only row-wise and scalar column accesses -> odd stride
    npad = (n + 1) & ~1;
    mpad = (mm + 1) & ~1;
    npart = nthreads > n ? nthreads : n;
    offAs = 0;
    offMs = offAs + ldA * n;
    offN = (offMs + ldN * n + 1) & ~1;  // every vector starts on a 2-scalar boundary (vector loads)
    offM = offN + 8 * npad;
    offPart = offM + 10 * mpad;
    offRed = (offPart + npart + 1) & ~1;
    offGj = offRed + 2 * 4 * kRedSlots;
    offSlot = (offGj + 4 * kGjPad + 1) & ~1;  // 8-byte aligned work-queue slot
    total = offSlot + 2;
  }
};

// out-of-line stages (defined below the struct).  They re-derive every pointer from the dynamic shared-memory
// symbol, so the compiler still knows the address space (LDS/STS, not generic LD/ST) while the kernel keeps ONE
// copy of each cold stage: the fully inlined version was 430 KB of SASS and spent half its time on instruction
// fetch (profiles/).
template <typename T, int G, int NS, int MS> __device__ __noinline__ bool qp_stage_gj(int n, int m, int where, int sz);
template <typename T, int G, int NS, int MS> __device__ __noinline__ bool qp_stage_gj_generic(int n, int m, T* Mx, int ld, int sz);
template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ T qp_stage_load_scale(const QpArgs<T, TIO>* a, long long b);
template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ int qp_stage_setup(const QpArgs<T, TIO>* a, T c);
template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ int qp_stage_check(const QpArgs<T, TIO>* a, long long b, T c);
template <typename T, int G, int NS, int MS, typename TIO>
__device__ __noinline__ unsigned qp_stage_polish(const QpArgs<T, TIO>* a, long long b, T c, int na, T* gscratch);
template <typename T, int G, int NS, int MS, typename TIO>
__device__ __noinline__ int qp_stage_loop(const QpArgs<T, TIO>* a, long long b, T c, unsigned long long t0, int code, unsigned* iter_out);

// NS, MS: compile-time problem shape (0 = runtime).  A shape-specialised instantiation lets the compiler fold every
// leading dimension / trip count into immediates, which removes most of the integer overhead of the GEMV passes.
template <typename T, int G, int NS, int MS> struct QpGroup
{
  static constexpr int NT = 32 * G;
  int n, m, ldA, ldN, npad, mpad, tid, lane, warp;
  // column-pass geometry: cw threads side by side own consecutive columns, csegs such bands split the rows
  int cw, csegs, cseg, c0;
  T* As;  // m x n   raw A, then Abar = Sy A Sx  (polish: active rows compacted on top, Sinv below when it fits)
  T* Ms;  // n x n   raw P, then M, then Minv    (polish: Kinv or the Woodbury inverse)
  T *sx, *q, *qb, *x, *xt, *xold, *nv1, *nv2;           // n-vectors
  T *sy, *l, *u, *rho, *rinv, *z, *y, *w, *yold, *mv1;  // m-vectors
  T* part;  // max(NT, n) scalars: partial results of column passes
  T* red;   // reduction scratch, double buffered
  T* gjbuf; // 4 * kGjPad scalars for the register-blocked inverse
  int rsel;
  T c;

  __device__ QpGroup(T* base, int n_, int m_) : n(n_), m(m_)
  {
    tid = threadIdx.x;
    lane = tid & 31;
    warp = tid >> 5;
    QpLayout L(n_, m_, NT);
    ldA = L.ldA;
    ldN = L.ldN;
    As = base + L.offAs;
    Ms = base + L.offMs;
    npad = L.npad;
    mpad = L.mpad;
    T* nv = base + L.offN;
    sx = nv; q = nv + npad; qb = nv + 2 * npad; x = nv + 3 * npad; xt = nv + 4 * npad; xold = nv + 5 * npad;
    nv1 = nv + 6 * npad; nv2 = nv + 7 * npad;
    T* mv = base + L.offM;
    sy = mv; l = mv + mpad; u = mv + 2 * mpad; rho = mv + 3 * mpad; rinv = mv + 4 * mpad; z = mv + 5 * mpad;
    y = mv + 6 * mpad; w = mv + 7 * mpad; yold = mv + 8 * mpad; mv1 = mv + 9 * mpad;
    part = base + L.offPart;
    red = base + L.offRed;
    gjbuf = base + L.offGj;
    rsel = 0;
    c = T(1);
    const int ru = ((n + 31) / 32) * 32;
    cw = ru < NT ? ru : NT;
    csegs = NT / cw;
    cseg = tid / cw;
    c0 = tid - cseg * cw;
    if (cseg >= csegs) { cseg = csegs; c0 = n; }  // NT not a multiple of cw: leftover threads own no column
  }

  // ---------------------------------------------------------------- group primitives
  __device__ __forceinline__ void gsync()
  {
    if (G == 1) __syncwarp(); else __syncthreads();
  }
  __device__ __forceinline__ bool gany(bool p)
  {
    if (G == 1) return __any_sync(kFullMask, p);
    return __syncthreads_or(p) != 0;
  }
  __device__ __forceinline__ bool gall(bool p)
  {
    if (G == 1) return __all_sync(kFullMask, p);
    return __syncthreads_and(p) != 0;
  }
  // batched reductions: K maxima and K2 sums in one barrier
  template <int KM, int KS> __device__ __forceinline__ void greduce(T (&mx)[KM], T (&sm)[KS])
  {
#pragma unroll
    for (int k = 0; k < KM; ++k) mx[k] = warp_max(mx[k]);
#pragma unroll
    for (int k = 0; k < KS; ++k) sm[k] = warp_sum(sm[k]);
    if (G > 1) {
      T* r = red + rsel * (4 * kRedSlots);
      rsel ^= 1;
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < KM; ++k) r[warp * kRedSlots + k] = mx[k];
#pragma unroll
        for (int k = 0; k < KS; ++k) r[warp * kRedSlots + KM + k] = sm[k];
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < KM; ++k) {
        T v = r[k];
#pragma unroll 1
        for (int wv = 1; wv < G; ++wv) v = fmax(v, r[wv * kRedSlots + k]);
        mx[k] = v;
      }
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        T v = r[KM + k];
#pragma unroll 1
        for (int wv = 1; wv < G; ++wv) v += r[wv * kRedSlots + KM + k];
        sm[k] = v;
      }
    }
  }
  __device__ __forceinline__ T gmax(T v)
  {
    T a[1] = {v}, b[1] = {T(0)};
    greduce<1, 0 + 1>(a, b);
    return a[0];
  }
  __device__ __forceinline__ T gsum(T v)
  {
    T a[1] = {T(0)}, b[1] = {v};
    greduce<1, 1>(a, b);
    return b[0];
  }
  __device__ __forceinline__ T part_sum(int idx, int len) const
  {
    T v = part[idx];
#pragma unroll 1
    for (int s = 1; s < csegs; ++s) v += part[s * len + idx];
    return v;
  }
  __device__ __forceinline__ T part_max(int idx, int len) const
  {
    T v = part[idx];
#pragma unroll 1
    for (int s = 1; s < csegs; ++s) v = fmax(v, part[s * len + idx]);
    return v;
  }

  // ---------------------------------------------------------------- stage one instance HBM -> shared memory
  template <typename TIO> __device__ void load(const QpArgs<T, TIO>& a, long long b)
  {
    const TIO* gA = a.A + b * (long long)m * n;
    const TIO* gP = a.P + b * (long long)n * n;
    // Coalesced streaming loads, 8 independent requests in flight per thread (the copy is pure DRAM latency:
    // with one request in flight it cost ~40k cycles per instance, profiles/r01_ncu_v5_summary.txt).
    auto copy = [&](const TIO* g, T* sm, int rows, int ld, int total) {
      constexpr int U = 8;
      int e = tid;
#pragma unroll 1
      for (; e + (U - 1) * NT < total; e += U * NT) {
        T v[U];
#pragma unroll
        for (int k = 0; k < U; ++k) v[k] = (T)__ldg(g + e + k * NT);
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const int ee = e + k * NT;
          const int j = ee / rows, i = ee - j * rows;
          sm[i + ld * j] = v[k];
        }
      }
#pragma unroll 1
      for (; e < total; e += NT) {
        const int j = e / rows, i = e - j * rows;
        sm[i + ld * j] = (T)__ldg(g + e);
      }
    };
    if (m > 0) copy(gA, As, m, ldA, m * n);
    copy(gP, Ms, n, ldN, n * n);
#pragma unroll 1
    for (int j = tid; j < n; j += NT) q[j] = (T)__ldg(a.q + b * (long long)n + j);
#pragma unroll 1
    for (int i = tid; i < m; i += NT) {
      l[i] = (T)__ldg(a.l + b * (long long)m + i);
      u[i] = (T)__ldg(a.u + b * (long long)m + i);
    }
    gsync();
  }

  // ---------------------------------------------------------------- QPSolver::scale, qp_solver.hpp:673-730
  // Products are formed in the reference's order and max() is exact, so c, sx, sy agree with the CPU
  // restatement bit for bit whatever the thread decomposition.
  __device__ void scale()
  {
#pragma unroll 1
    for (int j = tid; j < n; j += NT) sx[j] = T(1);
#pragma unroll 1
    for (int i = tid; i < m; i += NT) sy[i] = T(1);
    const int rpsP = (n + csegs - 1) / csegs, rpsA = (m + csegs - 1) / csegs;
#pragma unroll 1
    for (int j = c0; j < n; j += cw) {
      const int i0 = cseg * rpsP, i1 = min(n, i0 + rpsP);
      T g = T(0);
#pragma unroll 1
      for (int i = i0; i < i1; ++i) g = fmax(g, fabs(Ms[i + ldN * j]));  // :681-685
      part[cseg * n + j] = g;
    }
    gsync();
    T qn = T(0);
#pragma unroll 1
    for (int j = tid; j < n; j += NT) {
      T g = part_max(j, n);
      if (g == T(0)) g = T(1);  // :688-690
      nv1[j] = g;
      qn = fmax(qn, fabs(q[j]));
    }
    qn = gmax(qn);
    gsync();
    if (tid == 0) {
      T mean = T(0);
#pragma unroll 1
      for (int j = 0; j < n; ++j) mean += nv1[j];  // sequential on purpose: same rounding as the CPU restatement
      mean /= T(n);
      nv2[0] = mean;
    }
    gsync();
    c = T(1) / fmax(fmax(T(1e-6), nv2[0]), qn);  // :693
    gsync();

    int it = 0;
    bool again;
#pragma unroll 1
    do {
      // column norms of [Ps As'; As 0]  :701-716
#pragma unroll 1
      for (int j = c0; j < n; j += cw) {
        const T sxj = sx[j];
        T g = T(0);
        {
          const int i0 = cseg * rpsP, i1 = min(n, i0 + rpsP);
          T g1 = T(0);
          int i = i0;
#pragma unroll 2
          for (; i + 1 < i1; i += 2) {  // two independent maxima: max is exact, the split does not change the result
            g = fmax(g, fabs(((c * sx[i]) * sxj) * Ms[i + ldN * j]));
            g1 = fmax(g1, fabs(((c * sx[i + 1]) * sxj) * Ms[i + 1 + ldN * j]));
          }
          if (i < i1) g = fmax(g, fabs(((c * sx[i]) * sxj) * Ms[i + ldN * j]));
          g = fmax(g, g1);
        }
        {
          const int i0 = cseg * rpsA, i1 = min(m, i0 + rpsA);
          T g1 = T(0);
          int i = i0;
#pragma unroll 2
          for (; i + 1 < i1; i += 2) {
            g = fmax(g, fabs((sy[i] * sxj) * As[i + ldA * j]));
            g1 = fmax(g1, fabs((sy[i + 1] * sxj) * As[i + 1 + ldA * j]));
          }
          if (i < i1) g = fmax(g, fabs((sy[i] * sxj) * As[i + ldA * j]));
          g = fmax(g, g1);
        }
        part[cseg * n + j] = g;
      }
#pragma unroll 1
      for (int i = tid; i < m; i += NT) {
        const T syi = sy[i];
        T g = T(0);
        T g1 = T(0);
        int j = 0;
#pragma unroll 2
        for (; j + 1 < n; j += 2) {
          g = fmax(g, fabs((syi * sx[j]) * As[i + ldA * j]));
          g1 = fmax(g1, fabs((syi * sx[j + 1]) * As[i + ldA * (j + 1)]));
        }
        if (j < n) g = fmax(g, fabs((syi * sx[j]) * As[i + ldA * j]));
        g = fmax(g, g1);
        if (g == T(0)) g = T(1);
        mv1[i] = g;
      }
      gsync();
      T dev = T(0);
#pragma unroll 1
      for (int j = tid; j < n; j += NT) {
        T g = part_max(j, n);
        if (g == T(0)) g = T(1);
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];  // :726
        dev = fmax(dev, fabs(g - T(1)));
      }
#pragma unroll 1
      for (int i = tid; i < m; i += NT) {
        const T g = mv1[i];
        sy[i] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[i];  // :727
        dev = fmax(dev, fabs(g - T(1)));
      }
      dev = gmax(dev);
      gsync();
      again = (it++ < 10) && (dev > T(0.1));  // :728-729
    } while (again);
  }

  // ---------------------------------------------------------------- SPD inverse, register-blocked Gauss-Jordan
  // Thread (ti, tj) of a TR x TC grid owns the cyclic block {(ti + TR a, tj + TC b)} (a < RB, b < CB) in registers for
  // the whole elimination; per step only the pivot row and column travel through shared memory (double buffered:
  // one barrier per step).  Needs sz <= TR*RB and sz <= TC*CB; buf: 4*sz scalars of shared scratch.
  static constexpr int TR = (G == 4) ? 16 : 8;
  static constexpr int TC = NT / TR;
  static constexpr int RB = 4, CB = 8;
  __device__ __forceinline__ bool gj_fits_regs(int sz) const { return sz <= TR * RB && sz <= TC * CB; }
  __device__ bool gj_invert_reg(T* Mx, int ld, int sz, T* buf)
  {
    // buf: 4 * kGjPad scalars (two {row, column} buffer pairs, padded so that no bounds checks are needed)
    const int ti = tid % TR, tj = tid / TR;
    T v[RB][CB];
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
      for (int b2 = 0; b2 < CB; ++b2) {
        const int i = ti + TR * a, j = tj + TC * b2;
        v[a][b2] = (i < sz && j < sz) ? Mx[i + ld * j] : T(0);
      }
    bool ok = true;
    // The pivot index runs as k = ak*TR + h*TC + t2 so that the register slots (ak, bk) holding row k and
    // column k are compile-time constants: no dynamic register indexing, no branches in the step body.
#pragma unroll
    for (int ak = 0; ak < RB; ++ak) {
#pragma unroll
      for (int h = 0; h < TR / TC; ++h) {
        constexpr int Q = TR / TC;
        const int bk = ak * Q + h;
        // 1 / pivot is computed by every thread on ITS slot (ak, bk) one step ahead, overlapped with the rank-1 update; the
        // thread that owns the pivot publishes it in colk[k] (that entry of the pivot column is never read: the pivot row
        // zeroes its own multiplier) -- the reciprocal is off the barrier -> update critical path of every step
        T pn = rcp_inline(v[ak][bk]);
#pragma unroll 1
        for (int t2 = 0; t2 < TC; ++t2) {
          const int t = h * TC + t2;
          const int k = ak * TR + t;
          if (k >= sz || !ok) break;
          T* rowk = buf + (k & 1) * (2 * kGjPad);
          T* colk = rowk + kGjPad;
          if (ti == t) {
#pragma unroll
            for (int b2 = 0; b2 < CB; ++b2) rowk[tj + TC * b2] = v[ak][b2];
          }
          if (tj == t2) {
#pragma unroll
            for (int a = 0; a < RB; ++a) colk[ti + TR * a] = (ti == t && a == ak) ? pn : v[a][bk];
          }
          gsync();
          const T p = rowk[k];
          const T pinv = colk[k];
          if (!(p > T(0)) || !(p < Num<T>::inf())) { ok = false; break; }
          T ck[RB], rk[CB];
#pragma unroll
          for (int a = 0; a < RB; ++a) ck[a] = colk[ti + TR * a];
#pragma unroll
          for (int b2 = 0; b2 < CB; ++b2) rk[b2] = rowk[tj + TC * b2] * pinv;
          // pivot column (j == k): result -c_i * pinv  ==  0 - c_i * rk'  with rk' = pinv
          if (tj == t2) {
            rk[bk] = pinv;
#pragma unroll
            for (int a = 0; a < RB; ++a) v[a][bk] = T(0);
          }
          // pivot row (i == k): result rk_j (pinv at the pivot)  ==  rk_j - 0 * rk_j
          if (ti == t) {
            ck[ak] = T(0);
#pragma unroll
            for (int b2 = 0; b2 < CB; ++b2) v[ak][b2] = rk[b2];
          }
          v[ak][bk] -= ck[ak] * rk[bk];  // the next pivot's slot first: its reciprocal overlaps the rest of the update
          pn = rcp_inline(v[ak][bk]);
#pragma unroll
          for (int a = 0; a < RB; ++a)
#pragma unroll
            for (int b2 = 0; b2 < CB; ++b2)
              if (a != ak || b2 != bk) v[a][b2] -= ck[a] * rk[b2];
          keep_here(pn);  // the reciprocal stays in THIS iteration (not sunk to the top of the next one)
        }
      }
    }
    gsync();
    if (!ok) return false;
#pragma unroll
    for (int a = 0; a < RB; ++a)
#pragma unroll
      for (int b2 = 0; b2 < CB; ++b2) {
        const int i = ti + TR * a, j = tj + TC * b2;
        if (i < sz && j < sz) Mx[i + ld * j] = v[a][b2];
      }
    gsync();
    return true;
  }

  // ---------------------------------------------------------------- SPD inverse in place (Gauss-Jordan, no pivoting)
  // Mx may live in shared or global memory.  rowk / colk: >= sz scalars of shared scratch.
  // Returns false (group-uniform) on a non-positive or non-finite pivot.
  // where = 0: Ms (n x n, ldN);  where = 1: the Schur block kept below the compacted rows, As + sz (sz x sz, ldA)
  __device__ __forceinline__ bool gj_invert_at(int where, int sz)
  {
    if (gj_fits_regs(sz)) return qp_stage_gj<T, G, NS, MS>(n, m, where, sz);
    return qp_stage_gj_generic<T, G, NS, MS>(n, m, where == 0 ? Ms : As + sz, where == 0 ? ldN : ldA, sz);
  }
  __device__ bool gj_invert_smem(T* Mx, int ld, int sz, T* rowk, T* colk)
  {
#pragma unroll 1
    for (int k = 0; k < sz; ++k) {
      const T p = Mx[k + ld * k];
      if (!(p > T(0)) || !(p < Num<T>::inf())) return false;
      const T pinv = T(1) / p;
#pragma unroll 1
      for (int j = tid; j < sz; j += NT) {
        rowk[j] = Mx[k + ld * j] * pinv;
        colk[j] = Mx[j + ld * k];
      }
      gsync();
      int i = tid % sz, j = tid / sz;
      const int di = NT % sz, dj = NT / sz;
#pragma unroll 1
      for (int e = tid; e < sz * sz; e += NT) {
        T v;
        if (i == k) v = (j == k) ? pinv : rowk[j];
        else if (j == k) v = -colk[i] * pinv;
        else v = Mx[i + ld * j] - colk[i] * rowk[j];
        Mx[i + ld * j] = v;
        i += di; j += dj;
        if (i >= sz) { i -= sz; ++j; }
      }
      gsync();
    }
    return true;
  }

  // ---------------------------------------------------------------- transposed product, partials into part[]
  // part[seg * n + j] = sum_{i in seg} As[i, j] * v[i]   (and a second vector into part2 if given)
  __device__ __forceinline__ void colpass(const T* v, int rows)
  {
    int rps = (rows + csegs - 1) / csegs;
    rps = (rps + 1) & ~1;  // segments start on even rows: 2-scalar loads stay aligned
    const int i0 = min(rows, cseg * rps), i1 = min(rows, i0 + rps);
#pragma unroll 1
    for (int j = c0; j < n; j += cw) {
      const T* col = As + ldA * j;
      T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
      int i = i0;
#pragma unroll(NS > 0 ? 4 : 2)
      for (; i + 3 < i1; i += 4) {
        const auto c01 = ld2(col + i), c23 = ld2(col + i + 2);
        const auto v01 = ld2(v + i), v23 = ld2(v + i + 2);
        a0 += c01.x * v01.x;
        a1 += c01.y * v01.y;
        a2 += c23.x * v23.x;
        a3 += c23.y * v23.y;
      }
#pragma unroll 1
      for (; i < i1; ++i) a0 += col[i] * v[i];
      part[cseg * n + j] = (a0 + a1) + (a2 + a3);
    }
  }
  // tall-skinny variant (n <= 8): every thread strides over rows, one group reduction per column; result in o[]
  __device__ void colpass_skinny(const T* v, int rows, T* o)
  {
    T acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = T(0);
#pragma unroll 1
    for (int i = tid; i < rows; i += NT) {
      const T vi = v[i];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < n) acc[j] += As[i + ldA * j] * vi;
    }
    T dummy[1] = {T(0)};
    greduce<1, 8>(dummy, acc);
    if (tid < n) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j == tid) o[j] = acc[j];
    }
  }
  // o[j] = (Abar^T v)_j  for all j, visible to the whole group on return
  __device__ void At_vec(const T* v, int rows, T* o)
  {
    if (n <= 8) {
      colpass_skinny(v, rows, o);
      gsync();
    } else {
      colpass(v, rows);
      gsync();
#pragma unroll 1
      for (int j = tid; j < n; j += NT) o[j] = part_sum(j, n);
      gsync();
    }
  }
  // dot of two contiguous, 2-scalar aligned shared-memory vectors (4 accumulators)
  __device__ __forceinline__ T vecdot(const T* a, const T* v, int len) const
  {
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    int i = 0;
#pragma unroll 1
    for (; i + 3 < len; i += 4) {
      const auto c01 = ld2(a + i), c23 = ld2(a + i + 2);
      const auto v01 = ld2(v + i), v23 = ld2(v + i + 2);
      a0 += c01.x * v01.x;
      a1 += c01.y * v01.y;
      a2 += c23.x * v23.x;
      a3 += c23.y * v23.y;
    }
#pragma unroll 1
    for (; i < len; ++i) a0 += a[i] * v[i];
    return (a0 + a1) + (a2 + a3);
  }
  // dot of row i of a column-major matrix with v (4 accumulators for ILP)
  __device__ __forceinline__ T rowdot(const T* Mx, int ld, int i, int cols, const T* v) const
  {
    const T* p = Mx + i;
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    int j = 0;
#pragma unroll(NS > 0 ? 4 : 2)
    for (; j + 3 < cols; j += 4) {
      const auto v01 = ld2(v + j), v23 = ld2(v + j + 2);  // v is 2-scalar aligned (vector strides are padded)
      a0 += p[ld * j] * v01.x;
      a1 += p[ld * (j + 1)] * v01.y;
      a2 += p[ld * (j + 2)] * v23.x;
      a3 += p[ld * (j + 3)] * v23.y;
    }
#pragma unroll 1
    for (; j < cols; ++j) a0 += p[ld * j] * v[j];
    return (a0 + a1) + (a2 + a3);
  }

  // ---------------------------------------------------------------- check_stopping, qp_solver.hpp:574-644
  // Called right after the iterate update of a check iteration; xold / yold hold the pre-update iterates.
  // A x_us is evaluated as Sy^-1 (Abar x) and A^T y_us as Sx^-1 Abar^T y / c (the same quantities).
  template <typename TIO> __device__ int check_stopping(const QpArgs<T, TIO>& a, const TIO* gP)
  {
    const T eps_abs = T(a.prm.eps_abs), eps_rel = T(a.prm.eps_rel);
    const T eps_pinf = T(a.prm.eps_primal_inf), eps_dinf = T(a.prm.eps_dual_inf);
    const T inf = Num<T>::inf();

    // n-space: x_us -> nv1, dx (scaled) -> nv2, dx_us -> xold ; m-space: dy (scaled) -> w
    T qn = T(0), dxn = T(0), qdx = T(0), Edy = T(0);
#pragma unroll 1
    for (int j = tid; j < n; j += NT) {
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
#pragma unroll 1
    for (int i = tid; i < m; i += NT) {
      const T d = y[i] - yold[i];
      w[i] = d;
      Edy = fmax(Edy, fabs(sy[i] * d / c));  // :485
    }
    {
      T mx[3] = {qn, dxn, Edy}, sm[1] = {qdx};
      greduce<3, 1>(mx, sm);
      qn = mx[0]; dxn = mx[1]; Edy = mx[2]; qdx = sm[0];
    }
    gsync();

    // row pass: A x_us, A dx_us
    T n_Ax = T(0), n_r = T(0), n_z = T(0), s_pinf = T(0);
    bool pinf_blocked = false, dinf_rows_ok = true;
#pragma unroll 1
    for (int i = tid; i < m; i += NT) {
      T ax = T(0), adx = T(0);
      const T* p = As + i;
#pragma unroll 1
      for (int j = 0; j < n; ++j) {
        const T aij = p[ldA * j];
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
    {
      T mx[3] = {n_Ax, n_r, n_z}, sm[1] = {s_pinf};
      greduce<3, 1>(mx, sm);
      n_Ax = mx[0]; n_r = mx[1]; n_z = mx[2]; s_pinf = sm[0];
    }
    pinf_blocked = gany(pinf_blocked);
    dinf_rows_ok = gall(dinf_rows_ok);
    if (pinf_blocked) s_pinf = inf;

    // column passes: Abar^T y -> xt, Abar^T dy -> nv2 (scaled dx is dead after the row pass)
    gsync();
    At_vec(y, m, xt);
    At_vec(w, m, nv2);
    T n_Px = T(0), n_Aty = T(0), n_res = T(0), n_Atdy = T(0), n_Pdx = T(0);
#pragma unroll 1
    for (int j = tid; j < n; j += NT) {
      const T sc = T(1) / (sx[j] * c);
      const T aty = xt[j] * sc;
      const T atdy = nv2[j] * sc;
      T px = T(0), pdx = T(0);
#pragma unroll 10
      for (int k = 0; k < n; ++k) {  // unrolled: the L2 loads are independent, keep ~10 in flight
        const T pjk = (T)__ldg(gP + j + (long long)n * k);  // row j of the unscaled P (coalesced across threads)
        px += pjk * nv1[k];
        pdx += pjk * xold[k];
      }
      n_Px = fmax(n_Px, fabs(px));
      n_Aty = fmax(n_Aty, fabs(aty));
      n_res = fmax(n_res, fabs(px + q[j] + aty));
      n_Atdy = fmax(n_Atdy, fabs(atdy));
      n_Pdx = fmax(n_Pdx, fabs(pdx));
    }
    {
      T mx[5] = {n_Px, n_Aty, n_res, n_Atdy, n_Pdx}, sm[1] = {T(0)};
      greduce<5, 1>(mx, sm);
      n_Px = mx[0]; n_Aty = mx[1]; n_res = mx[2]; n_Atdy = mx[3]; n_Pdx = mx[4];
    }
    gsync();

    // OPTIMALITY :584-594
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    // PRIMAL INFEASIBILITY :619
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;
    // DUAL INFEASIBILITY :629-641
    // Guard (deviations; sfb_set_option(SFB_OPT_DUAL_INF_DX_GUARD)): with dx == 0 exactly every test of :629-639
    // reads 0 <= 0 and the literal rule reports DualInfeasible for an iterate that merely stopped moving.  The reference's
    // x carries rounding noise from its (n+m)-long LDL^T sweeps and is exactly stationary only on scalar (n = 1) problems
    // (tests/test_oracle_qp_known_answers.py); the reduced system here can be during active-set plateaus of tall problems.
    if ((dxn > T(0) || !a.dinf_guard) && (n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  // Pbar(i, j): upper triangle of c Sx P Sx mirrored (selfadjointView<Upper>), from the unscaled global P
  template <typename TIO> __device__ __forceinline__ T pbar(const TIO* gP, int i, int j) const
  {
    const int r = i <= j ? i : j, cc = i <= j ? j : i;
    return ((c * sx[r]) * (T)__ldg(gP + r + (long long)n * cc)) * sx[cc];
  }

  // ---------------------------------------------------------------- detail::polish_qp, qp_solver.hpp:92-204
  // `na` active rows (ascending) are listed in idx[], their scaled bounds in bnd[].  Returns SFB_QP_FLAG_* bits.
  template <typename TIO> __device__ unsigned polish(const QpArgs<T, TIO>& a, const TIO* gP, int na, const int* idx, const T* bnd, T* gscratch)
  {
    // fp32: the delta = 1e-6 regularised polish systems are not resolvable in single precision (8 ulp): an fp32 solve flags
    // the polish as skipped here and the host launches this kernel again in fp64 on the fp32 iterate and active set (mode 2,
    // sfb_api.cu), which is where the _f32 entry points get SFB_QP_FLAG_POLISHED from.
    if (sizeof(T) == 4) return SFB_QP_FLAG_POLISH_SKIPPED;
    const T delta = T(a.prm.delta);
    const T dinv = T(1) / delta;
    // Two block eliminations of the same regularised system Hp = [K Aa^T; Aa -delta I], K = Pbar + delta I:
    //   schur    primal block first (the order Eigen's diagonal pivoting takes): Kinv, S = delta I + Aa Kinv Aa^T (na x na), Sinv;
    //   reduced  duals first: N = K + Aa^T Aa / delta (n x n), ONE inverse, no S -- 40 % less work at the headline shape, but
    //            1 / delta sits inside N, so one application of the computed inverse is only good to cond(N) eps.
    // a.polish_form 0 (default): reduced, refined with literal residual sweeps until the correction is below 1e-8 of the
    // solution; an instance that does not get there within polish_iter sweeps is redone in the schur form.  1: schur always.  2: reduced always.
    // na > n has no schur form (S would be singular-ish).  A CTA owns one instance: the choice is CTA-uniform.
    T* S = nullptr;
    int ldS = 0;
    bool s_shared = false;  // NOT derivable from ldS == ldA: with every row of an m == ldA problem active, na == ldA too (the
                            // root cause of the round-1 "2 x 2" polish failures: S in the workspace was addressed as if on chip)
    if (na > 0 && na <= n) {
      if (2 * na <= ldA && !a.force_polish_scratch) { S = As + na; ldS = ldA; s_shared = true; }  // below the compacted rows
      else if (gscratch != nullptr && (long long)na * na <= a.scratch_per_cta) { S = gscratch; ldS = na; }
    }
    const bool schur_possible = (na == 0) || (S != nullptr);
    bool woodbury = (na > n) || (na > 0 && a.polish_form != 1);
    if (!woodbury && !schur_possible) return SFB_QP_FLAG_POLISH_SKIPPED;
    const bool may_fall_back = woodbury && na <= n && a.polish_form == 0 && schur_possible;

    // compact the active rows of Abar to the top of every column (idx ascending => in-place safe)
#pragma unroll 1
    for (int j = tid; j < n; j += NT) {
      T* col = As + ldA * j;
#pragma unroll 1
      for (int r = 0; r < na; ++r) col[r] = col[idx[r]];
    }
    gsync();

    T* tx = xt;    // n
    T* ty = z;     // na
    T* rx = nv1;   // n
    T* ux = nv2;   // n
    T* ry = yold;  // na
    T* sv = rho;   // na
    T* dy = rinv;  // na
    bool used_scratch = false;
#pragma unroll 1
    for (int attempt = 0; attempt < 2; ++attempt) {
      // K = Pbar + delta I  (+ Aa^T Aa / delta in the reduced form)   :161,175
      // one coalesced sweep over the unscaled P: every upper-triangle entry is scaled and written to both (i,j), (j,i)
      {
        int i = tid % n, j = tid / n;
        const int di = NT % n, dj = NT / n;
#pragma unroll 1
        for (int e = tid; e < n * n; e += NT) {
          if (i <= j) {
            T h = ((c * sx[i]) * (T)__ldg(gP + e)) * sx[j];
            if (i == j) h += delta;
            Ms[i + ldN * j] = h;
            Ms[j + ldN * i] = h;
          }
          i += di; j += dj;
          if (i >= n) { i -= n; ++j; }
        }
      }
      gsync();
      if (woodbury) form_weighted_gram(T(0), na, nullptr, dinv);
      bool ok = gj_invert_at(0, n);

      if (ok && !woodbury && S != nullptr) {
        // S = delta I + Aa Kinv Aa^T, four columns at a time: T4 = Kinv Aa[s0..s0+3,:]^T (n x 4 in xt,xold,nv1,nv2; stride npad),
        // then S[:, s0..s0+3] = Aa T4
#pragma unroll 1
        for (int s0 = 0; s0 < na; s0 += 4) {
          const int ns = min(4, na - s0);
#pragma unroll 1
          for (int e = tid; e < n * ns; e += NT) {
            const int i = e % n, sc = e / n;
            const T* p = Ms + i;
            const T* arow = As + (s0 + sc);
            T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
            int j = 0;
#pragma unroll 2
            for (; j + 3 < n; j += 4) {
              a0 += p[ldN * j] * arow[ldA * j];
              a1 += p[ldN * (j + 1)] * arow[ldA * (j + 1)];
              a2 += p[ldN * (j + 2)] * arow[ldA * (j + 2)];
              a3 += p[ldN * (j + 3)] * arow[ldA * (j + 3)];
            }
#pragma unroll 1
            for (; j < n; ++j) a0 += p[ldN * j] * arow[ldA * j];
            xt[sc * npad + i] = (a0 + a1) + (a2 + a3);
          }
          gsync();
#pragma unroll 1
          for (int e = tid; e < na * ns; e += NT) {
            const int r = e % na, sc = e / na;
            T acc = rowdot(As, ldA, r, n, xt + sc * npad);
            if (r == s0 + sc) acc += delta;
            if (s_shared) As[na + r + ldA * (s0 + sc)] = acc; else S[r + ldS * (s0 + sc)] = acc;
          }
          gsync();
        }
        ok = s_shared ? gj_invert_at(1, na) : qp_stage_gj_generic<T, G, NS, MS>(n, m, S, ldS, na);
        used_scratch = !s_shared;
      }

      bool accurate = true;
      if (ok) {
        // iterative refinement  t += Hp^-1 (h - H t)   :192-195
#pragma unroll 1
        for (int j = tid; j < n; j += NT) tx[j] = T(0);
#pragma unroll 1
        for (int r = tid; r < na; r += NT) ty[r] = T(0);
        gsync();
        // The reference iterates t <- t + Hp^-1 (h - H t) with H = Hp - D, D = diag(delta I_n, -delta I_na).  Since
        // h - H t = (h + D t) - Hp t, the same sequence is t <- Hp^-1 (h + D t): no product with H (whose Pbar block lives
        // in HBM) is needed.  The closing sweep(s) use the literal residual form, so that rounding errors of the explicit
        // inverses are corrected, exactly like iterative refinement does.
        bool converged = false;
#pragma unroll 1
        for (uint32_t it = 0; it != a.prm.polish_iter; ++it) {
          // schur: once t stops changing (to a few ulp) further sweeps of the fixed-point form are no-ops up to rounding: jump
          // to ONE closing literal sweep.  Well-conditioned systems contract by delta / lambda_min ~ 1e-6 per sweep.
          // reduced: the fixed point of the COMPUTED inverse is only good to cond(N) eps (y: 1e-2), so every sweep after the
          // first is a literal one, repeated until the correction is negligible (typically two: 1e-2 -> 1e-9 -> accepted).
          const bool literal = (it > 0) && (woodbury || (it + 1 == a.prm.polish_iter) || converged);
          if (literal) {
            // residual r = h - sym(H) t,  H = [Pbar Aa^T; Aa 0]
#pragma unroll 1
            for (int i = tid; i < n; i += NT) {
              T acc = T(0);
#pragma unroll 10
              for (int j = 0; j < n; ++j) acc += pbar(gP, i, j) * tx[j];
              rx[i] = -c * (sx[i] * q[i]) - (acc + vecdot(As + ldA * i, ty, na));  // :180
            }
#pragma unroll 1
            for (int r = tid; r < na; r += NT) ry[r] = bnd[r] - rowdot(As, ldA, r, n, tx);  // :181-182
          } else {
            // rhs = h + D t ; the solve below then yields the new t directly
#pragma unroll 1
            for (int i = tid; i < n; i += NT) rx[i] = -c * (sx[i] * q[i]) + delta * tx[i];
#pragma unroll 1
            for (int r = tid; r < na; r += NT) ry[r] = bnd[r] - delta * ty[r];
          }
          gsync();
          T diff = T(0), mag = T(0), diffy = T(0), magy = T(0);  // literal sweeps: x and y blocks apart (their scales differ)
          if (!woodbury) {
            // [K Aa^T; Aa -delta I] [dx; dy] = [rx; ry]:  dy = Sinv (Aa Kinv rx - ry),  dx = Kinv (rx - Aa^T dy)
#pragma unroll 1
            for (int i = tid; i < n; i += NT) ux[i] = rowdot(Ms, ldN, i, n, rx);
            gsync();
#pragma unroll 1
            for (int r = tid; r < na; r += NT) sv[r] = rowdot(As, ldA, r, n, ux) - ry[r];
            gsync();
#pragma unroll 1
            for (int r = tid; r < na; r += NT)
              dy[r] = s_shared ? rowdot(As + na, ldA, r, na, sv) : rowdot(S, ldS, r, na, sv);  // keep LDS on the common path
            gsync();
#pragma unroll 1
            for (int i = tid; i < n; i += NT) ux[i] = rx[i] - vecdot(As + ldA * i, dy, na);
            gsync();
#pragma unroll 1
            for (int i = tid; i < n; i += NT) {
              const T d = rowdot(Ms, ldN, i, n, ux);
              if (literal) {
                tx[i] += d;
              } else {
                diff = fmax(diff, fabs(d - tx[i]));
                mag = fmax(mag, fabs(d));
                tx[i] = d;
              }
            }
#pragma unroll 1
            for (int r = tid; r < na; r += NT) {
              const T d = dy[r];
              if (literal) {
                ty[r] += d;
              } else {
                diff = fmax(diff, fabs(d - ty[r]));
                mag = fmax(mag, fabs(d));
                ty[r] = d;
              }
            }
          } else {
            // same system, duals eliminated first:  (K + Aa^T Aa / delta) dx = rx + Aa^T ry / delta,  dy = (Aa dx - ry) / delta
#pragma unroll 1
            for (int i = tid; i < n; i += NT) ux[i] = rx[i] + dinv * vecdot(As + ldA * i, ry, na);
            gsync();
#pragma unroll 1
            for (int i = tid; i < n; i += NT) rx[i] = rowdot(Ms, ldN, i, n, ux);  // rx now holds dx
            gsync();
#pragma unroll 1
            for (int r = tid; r < na; r += NT) {
              const T d = (rowdot(As, ldA, r, n, rx) - ry[r]) * dinv;
              if (literal) {
                const T tn = ty[r] + d;
                diffy = fmax(diffy, fabs(d));
                magy = fmax(magy, fabs(tn));
                ty[r] = tn;
              } else {
                diff = fmax(diff, fabs(d - ty[r]));
                mag = fmax(mag, fabs(d));
                ty[r] = d;
              }
            }
#pragma unroll 1
            for (int i = tid; i < n; i += NT) {
              const T d = rx[i];
              if (literal) {
                const T tn = tx[i] + d;
                diff = fmax(diff, fabs(d));
                mag = fmax(mag, fabs(tn));
                tx[i] = tn;
              } else {
                diff = fmax(diff, fabs(d - tx[i]));
                mag = fmax(mag, fabs(d));
                tx[i] = d;
              }
            }
          }
          if (literal && !woodbury) {
            gsync();
            break;
          }
          {
            T mx[4] = {diff, mag, diffy, magy}, sm[1] = {T(0)};
            greduce<4, 1>(mx, sm);
            // a literal sweep's correction is the error BEFORE it (relative, per block); what is left after it is smaller, but
            // by a factor that is anything between 1e-7 (first sweep) and 0.5 (measured) -- so only the correction itself is
            // trusted: 1e-8 of the solution, two orders inside the 1e-6 parity bar
            if (literal) accurate = (mx[0] <= T(1e-8) * mx[1]) && (mx[2] <= T(1e-8) * mx[3]);
            else converged = mx[0] <= T(8) * Num<T>::eps() * mx[1];
          }
          gsync();
          if (literal && accurate) break;
        }
      }
      if (ok && accurate) break;
      if (!may_fall_back || !woodbury) {
        if (!ok) return SFB_QP_FLAG_POLISH_FAILED;
        break;  // refined as far as polish_iter sweeps go; there is no other form to try
      }
      woodbury = false;  // redo this instance in the schur form
    }
    // :199-201
#pragma unroll 1
    for (int j = tid; j < n; j += NT) x[j] = tx[j];
    gsync();  // ty aliases z, y is a distinct vector: scatter after everyone is done reading
#pragma unroll 1
    for (int r = tid; r < na; r += NT) y[idx[r]] = ty[r];
    gsync();
    return SFB_QP_FLAG_POLISHED | (used_scratch ? SFB_QP_FLAG_POLISH_SCRATCH : 0u) | (woodbury ? SFB_QP_FLAG_POLISH_REDUCED : 0u);
  }

  // ---------------------------------------------------------------- M = Pbar + sigma I + Abar^T R Abar
  // 4x4 register tiles over the upper triangle; each thread walks the m rows in a rotated order so that the
  // threads of a warp hit different shared-memory banks.
  __device__ void form_reduced_kkt(T sigma) { form_weighted_gram(sigma, m, rho, T(0)); }
  // Ms += diag + As[0..rows)^T W As[0..rows), W = diag(wts) or wconst I (wts == nullptr)
  __device__ void form_weighted_gram(T sigma, int rows, const T* wts, T wconst)
  {
    const int nb = (n + 3) / 4;
    const int ntiles = nb * (nb + 1) / 2;
#pragma unroll 1
    for (int t = tid; t < ntiles; t += NT) {
      int J = 0, rem = t;
#pragma unroll 1
      while (rem > J) { rem -= (J + 1); ++J; }
      const int I = rem;  // I <= J
      int ci[4], cj[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ci[k] = min(4 * I + k, n - 1) * ldA;
        cj[k] = min(4 * J + k, n - 1) * ldA;
      }
      T acc[4][4];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int s = 0; s < 4; ++s) acc[p][s] = T(0);
      int r = (rows > 0) ? (tid % rows) : 0;
#pragma unroll 1
      for (int k = 0; k < rows; ++k) {
        const T rr = wts != nullptr ? wts[r] : wconst;
        T av[4], bv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          av[p] = As[r + ci[p]];
          bv[p] = rr * As[r + cj[p]];
        }
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int s = 0; s < 4; ++s) acc[p][s] += av[p] * bv[s];
        if (++r == rows) r = 0;
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const int gi = 4 * I + p, gj = 4 * J + s;
          if (gi < n && gj < n && gi <= gj) {
            T v = Ms[gi + ldN * gj] + acc[p][s];
            if (gi == gj) v += sigma;
            Ms[gi + ldN * gj] = v;
            Ms[gj + ldN * gi] = v;
          }
        }
    }
    gsync();
  }

  // rho classes, trivially empty feasible set, in-place data scaling, reduced KKT matrix and its inverse
  template <typename TIO> __device__ int setup(const QpArgs<T, TIO>& a)
  {
    const T inf = Num<T>::inf();
    const T rho_bar = T(a.prm.rho), sigma = T(a.prm.sigma);
    int code = kStatusUnset;

    // rho per constraint class + trivially empty feasible set  :361-374
    bool triv = false;
#pragma unroll 1
    for (int i = tid; i < m; i += NT) {
      const T li = l[i], ui = u[i];
      if (li == inf || ui == -inf || ui - li < T(0)) triv = true;
      T r;
      if (li == -inf && ui == inf) r = T(1e-6);
      else if (sy[i] * fabs(li - ui) < T(1e-5)) r = T(1e3) * rho_bar;
      else r = rho_bar;
      rho[i] = r;
      rinv[i] = T(1) / r;
    }
    if (gany(triv)) code = SFB_QP_PRIMAL_INFEASIBLE;

    // scale the data in place:  qb = c Sx q,  Abar = Sy A Sx,  upper(Pbar) = c Sx P Sx   :401-403,:450
#pragma unroll 1
    for (int j = tid; j < n; j += NT) qb[j] = (c * sx[j]) * q[j];
#pragma unroll 1
    for (int i = tid; i < m; i += NT) {
      const T syi = sy[i];
      T* p = As + i;
#pragma unroll 1
      for (int j = 0; j < n; ++j) p[ldA * j] = (syi * p[ldA * j]) * sx[j];
    }
#pragma unroll 1
    for (int i = tid; i < n; i += NT) {
      const T csxi = c * sx[i];
#pragma unroll 1
      for (int j = i; j < n; ++j) Ms[i + ldN * j] = (csxi * Ms[i + ldN * j]) * sx[j];
    }
    gsync();

    if (a.mode == 2) return code;  // polish-only pass: the ADMM matrix is not needed
    form_reduced_kkt(sigma);
    if (!gj_invert_at(0, n)) code = SFB_QP_UNKNOWN;  // :433
    return code;
  }

  // ---------------------------------------------------------------- main ADMM loop  qp_solver.hpp:449-510
  // Outlined as its own stage (qp_stage_loop).  Returns the status code (kStatusUnset when the iteration budget ran out).
  template <typename TIO> __device__ int admm_loop(const QpArgs<T, TIO>& a, long long b, unsigned long long t0, int code, unsigned* iter_out)
  {
    const T alpha = T(a.prm.alpha), alpha_comp = T(1) - alpha, sigma = T(a.prm.sigma);
    const unsigned sci = a.prm.stop_check_iter;
    unsigned iter = 0;
#pragma unroll 1
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      // rhs_x = sigma x - qb + Abar^T w,  w = R z - y
      if (n <= 8) {
        colpass_skinny(w, m, nv2);
        gsync();
#pragma unroll 1
        for (int j = tid; j < n; j += NT) xt[j] = sigma * x[j] - qb[j] + nv2[j];
      } else {
        colpass(w, m);
        gsync();
#pragma unroll 1
        for (int j = tid; j < n; j += NT) xt[j] = sigma * x[j] - qb[j] + part_sum(j, n);
      }
      gsync();
      // xtilde = Minv rhs_x ; x <- alpha xtilde + (1 - alpha) x   :470
      const bool chk = (iter % sci == 1u);
#pragma unroll 1
      for (int i = tid; i < n; i += NT) {
        const T xti = rowdot(Ms, ldN, i, n, xt);
        nv1[i] = xti;
        const T xi = x[i];
        if (chk) xold[i] = xi;  // :465-468
        x[i] = alpha * xti + alpha_comp * xi;
      }
      gsync();
      // nu = R (Abar xtilde - z) + y ; z, y updates  :471-477 ; next w
#pragma unroll 1
      for (int i = tid; i < m; i += NT) {
        const T zt = rowdot(As, ldA, i, n, nv1);
        const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];
        if (chk) yold[i] = yi;
        const T nu = ri * (zt - zi) + yi;
        T v = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;
        v = fmax(v, sy[i] * l[i]);
        v = fmin(v, sy[i] * u[i]);
        const T yn = alpha_comp * yi + alpha * nu + ri * zi - ri * v;
        y[i] = yn;
        z[i] = v;
        w[i] = ri * v - yn;
      }
      gsync();
      if (chk) {
        code = qp_stage_check<T, G, NS, MS>(&a, b, c);  // :488  (clobbers w)
        if (code == kStatusUnset && a.prm.has_max_time) {
          // :504-508 ; one thread reads the clock so that the decision is group-uniform
          const bool late = (tid == 0) && ((long long)(global_timer_ns() - t0) > a.prm.max_time_ns);
          if (gany(late)) code = SFB_QP_MAX_TIME;
        }
#pragma unroll 1
        for (int i = tid; i < m; i += NT) w[i] = rho[i] * z[i] - y[i];
        gsync();
      }
    }

    *iter_out = iter;
    return code;
  }


  // ---------------------------------------------------------------- main ADMM loop, REGISTER-RESIDENT Abar
  // For compile-time shapes with n <= 56, m <= 112 and G = 4 (the headline n = 50, m = 100) every thread keeps a 7 x 7
  // block of Abar in registers for the whole loop: threads form a 16 x 8 grid (row group rg = 4 warp + lane / 8 owns rows
  // rg + 16 a, column group cg = lane % 8 owns columns cg + 8 b).  Per iteration
  //   Abar^T w : 49 FMAs on registers, reduce-scatter over the 4 row groups of the warp (6 shuffles), 4 warp partials
  //              through shared memory;
  //   Minv rhs : two threads per row of Minv (shared memory), one shuffle;
  //   Abar xt  : 49 FMAs on registers, reduce-scatter over the 8 column groups (7 shuffles) -- lane (rl, cg) ends up with
  //              row rg + 16 cg, whose z, y, rho, bounds live in ITS registers, so the z / y update touches no memory.
  // The shared-memory version above streams Abar twice per iteration (10 000 of its 12 500 operand loads); this one loads
  // it once per solve.  Same arithmetic, different summation order (parity with the oracle is unaffected: tests).
  static constexpr bool kRegLoop = (G == 4) && (NS > 0) && (NS <= 56) && (MS > 0) && (MS <= 112);
  template <typename TIO> __device__ int admm_loop_reg(const QpArgs<T, TIO>& a, long long b, unsigned long long t0, int code, unsigned* iter_out)
  {
    constexpr int RA = 7, CBK = 7;
    const T alpha = T(a.prm.alpha), alpha_comp = T(1) - alpha, sigma = T(a.prm.sigma);
    const unsigned sci = a.prm.stop_check_iter;
    const int cg = lane & 7, rl = lane >> 3, rg = warp * 4 + rl;
    T Ar[RA][CBK];
#pragma unroll
    for (int ia = 0; ia < RA; ++ia)
#pragma unroll
      for (int jb = 0; jb < CBK; ++jb) {
        const int i = rg + 16 * ia, j = cg + 8 * jb;
        Ar[ia][jb] = (i < m && j < n) ? As[i + ldA * j] : T(0);
      }
    // the row whose iterate this thread owns: a == cg after the reduce-scatter of Abar xt
    const int myrow = rg + 16 * cg;
    const bool has_row = (cg < RA) && (myrow < m);
    T zr = T(0), yr = T(0), rr = T(0), rinvr = T(0), lor = T(0), hir = T(0), yoldr = T(0);
    if (has_row) {
      zr = z[myrow]; yr = y[myrow]; rr = rho[myrow]; rinvr = rinv[myrow];
      lor = sy[myrow] * l[myrow]; hir = sy[myrow] * u[myrow];
    }
    T* part4 = gjbuf;  // 4 warps x kGjPad column partials (the inverse's scratch is idle during the loop)
    // Minv rows: lanes 0-15 of a warp take columns 0, 2, 4, ... of rows 16 warp + lane, lanes 16-31 the odd columns of the
    // same rows (each half-warp then reads 16 consecutive scalars: conflict-free; tid / 2, tid % 2 put rows r and r + 51 into
    // one half-warp and measured 35 % excess wavefronts)
    const int r2 = warp * 16 + (lane & 15), h2 = lane >> 4;
    const bool has2 = r2 < n;
    // iterations until the next stop check (iter % sci == 1) without a division per iteration; sci == 1 never checks (x % 1 != 1)
    unsigned until_check = (sci == 1u) ? 0xffffffffu : 1u;
    const int c0 = cg + 16 * rl, c1 = c0 + 8;  // the two columns this lane holds after the reduce-scatter of Abar^T w
    unsigned iter = 0;
#pragma unroll 1
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      // ---- rhs_x = sigma x - qb + Abar^T w,  w = R z - y
      {
        T wv[RA];
#pragma unroll
        for (int ia = 0; ia < RA; ++ia) {
          const int i = rg + 16 * ia;
          wv[ia] = (i < m) ? w[i] : T(0);
        }
        T pc[8];
#pragma unroll
        for (int jb = 0; jb < CBK; ++jb) {
          T acc = T(0);
#pragma unroll
          for (int ia = 0; ia < RA; ++ia) acc += Ar[ia][jb] * wv[ia];
          pc[jb] = acc;
        }
        pc[7] = T(0);
        const bool up16 = (lane & 16) != 0, up8 = (lane & 8) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const T keep = up16 ? pc[4 + k] : pc[k], send = up16 ? pc[k] : pc[4 + k];
          pc[k] = keep + __shfl_xor_sync(kFullMask, send, 16);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const T keep = up8 ? pc[2 + k] : pc[k], send = up8 ? pc[k] : pc[2 + k];
          pc[k] = keep + __shfl_xor_sync(kFullMask, send, 8);
        }
        if (c0 < n) part4[warp * kGjPad + c0] = pc[0];
        if (c1 < n) part4[warp * kGjPad + c1] = pc[1];
      }
      gsync();
      if (tid < n) {
        const T t = (part4[tid] + part4[kGjPad + tid]) + (part4[2 * kGjPad + tid] + part4[3 * kGjPad + tid]);
        xt[tid] = sigma * x[tid] - qb[tid] + t;
      }
      gsync();
      // ---- xtilde = Minv rhs_x ; x <- alpha xtilde + (1 - alpha) x   :470   (two threads per row: columns h2, h2 + 2, ...)
      const bool chk = (until_check == 0u);
      until_check = chk ? sci - 1u : until_check - 1u;
      {
        T a0 = T(0), a1 = T(0);
        if (has2) {
          const T* pr = Ms + r2 + ldN * h2;
          int k = 0;
#pragma unroll 6
          for (; k + 1 < (n - h2 + 1) / 2; k += 2) {
            a0 += pr[ldN * (2 * k)] * xt[h2 + 2 * k];
            a1 += pr[ldN * (2 * k + 2)] * xt[h2 + 2 * k + 2];
          }
          if (k < (n - h2 + 1) / 2) a0 += pr[ldN * (2 * k)] * xt[h2 + 2 * k];
        }
        T xti = a0 + a1;
        xti += __shfl_xor_sync(kFullMask, xti, 16);
        if (has2 && h2 == 0) {
          nv1[r2] = xti;
          const T xi = x[r2];
          if (chk) xold[r2] = xi;  // :465-468
          x[r2] = alpha * xti + alpha_comp * xi;
        }
      }
      gsync();
      // ---- nu = R (Abar xtilde - z) + y ; z, y updates  :471-477 ; next w
      {
        T xv[CBK];
#pragma unroll
        for (int jb = 0; jb < CBK; ++jb) {
          const int j = cg + 8 * jb;
          xv[jb] = (j < n) ? nv1[j] : T(0);
        }
        T zp[8];
#pragma unroll
        for (int ia = 0; ia < RA; ++ia) {
          T acc = T(0);
#pragma unroll
          for (int jb = 0; jb < CBK; ++jb) acc += Ar[ia][jb] * xv[jb];
          zp[ia] = acc;
        }
        zp[7] = T(0);
        const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const T keep = up4 ? zp[4 + k] : zp[k], send = up4 ? zp[k] : zp[4 + k];
          zp[k] = keep + __shfl_xor_sync(kFullMask, send, 4);
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const T keep = up2 ? zp[2 + k] : zp[k], send = up2 ? zp[k] : zp[2 + k];
          zp[k] = keep + __shfl_xor_sync(kFullMask, send, 2);
        }
        {
          const T keep = up1 ? zp[1] : zp[0], send = up1 ? zp[0] : zp[1];
          zp[0] = keep + __shfl_xor_sync(kFullMask, send, 1);
        }
        if (has_row) {
          const T zt = zp[0];
          if (chk) yoldr = yr;
          const T nu = rr * (zt - zr) + yr;
          T v = alpha * (rinvr * nu) + alpha_comp * (rinvr * yr) + zr;
          v = fmax(v, lor);
          v = fmin(v, hir);
          const T yn = alpha_comp * yr + alpha * nu + rr * zr - rr * v;
          yr = yn;
          zr = v;
          w[myrow] = rr * v - yn;
        }
      }
      if (chk) {
        if (has_row) { z[myrow] = zr; y[myrow] = yr; yold[myrow] = yoldr; }
        gsync();
        code = qp_stage_check<T, G, NS, MS>(&a, b, c);  // :488  (clobbers w)
        if (code == kStatusUnset && a.prm.has_max_time) {
          const bool late = (tid == 0) && ((long long)(global_timer_ns() - t0) > a.prm.max_time_ns);
          if (gany(late)) code = SFB_QP_MAX_TIME;
        }
        if (has_row) w[myrow] = rr * zr - yr;
      }
      // w[rg + 16 ia] is produced and consumed by the SAME warp (rg = 4 warp + rl on both sides), and every other buffer of
      // the iteration is already ordered by the three barriers above (part4: read before barrier 2, rewritten after barrier 3
      // of the same iteration; xt, nv1, x: rewritten only after the next barrier 1): a warp-level sync is enough here
      __syncwarp();
    }
    if (has_row) { z[myrow] = zr; y[myrow] = yr; }
    gsync();
    *iter_out = iter;
    return code;
  }

  // ---------------------------------------------------------------- QPSolver::solve, qp_solver.hpp:343-568
  template <typename TIO> __device__ void solve(const QpArgs<T, TIO>& a, long long b, T* gscratch)
  {
    const T inf = Num<T>::inf();
    const TIO* gP = a.P + b * (long long)n * n;
    const unsigned long long t0 = a.prm.has_max_time ? global_timer_ns() : 0ull;

    c = qp_stage_load_scale<T, G, NS, MS>(&a, b);  // :347
    if (a.mode == 1) {
      if (tid == 0) a.out_c[b] = c;
#pragma unroll 1
      for (int j = tid; j < n; j += NT) a.out_sx[b * (long long)n + j] = sx[j];
#pragma unroll 1
      for (int i = tid; i < m; i += NT) a.out_sy[b * (long long)m + i] = sy[i];
      gsync();
      return;
    }

    // Mixed-precision polish (mode 2, T = double over TIO = float data): the instance was solved by the single-precision
    // kernel, whose delta = 1e-6 regularised polish systems are not resolvable in fp32; this pass re-stages the problem in
    // double, takes the unpolished iterate and the active set from the outputs and runs only polish_qp on them.
    const bool polish_only = a.mode == 2;
    if (polish_only && a.out_status[b] != (int32_t)SFB_QP_OPTIMAL) return;  // uniform: every thread reads the same word

    int code = qp_stage_setup<T, G, NS, MS>(&a, c);  // rho classes, trivial infeasibility, in-place scaling, M, Minv

    // initial iterate  :436-445
    if (polish_only) {
#pragma unroll 1
      for (int j = tid; j < n; j += NT) x[j] = (T(1) / sx[j]) * (T)a.out_x[b * (long long)n + j];
#pragma unroll 1
      for (int i = tid; i < m; i += NT) {
        y[i] = c * ((T(1) / sy[i]) * (T)a.out_y[b * (long long)m + i]);
        z[i] = T(0);
      }
    } else if (a.warm_x != nullptr) {
#pragma unroll 1
      for (int j = tid; j < n; j += NT) x[j] = (T(1) / sx[j]) * (T)__ldg(a.warm_x + b * (long long)n + j);
#pragma unroll 1
      for (int i = tid; i < m; i += NT) y[i] = c * ((T(1) / sy[i]) * (T)__ldg(a.warm_y + b * (long long)m + i));
      gsync();
#pragma unroll 1
      for (int i = tid; i < m; i += NT) z[i] = rowdot(As, ldA, i, n, x);
    } else {
#pragma unroll 1
      for (int j = tid; j < n; j += NT) x[j] = T(0);
#pragma unroll 1
      for (int i = tid; i < m; i += NT) {
        y[i] = T(0);
        z[i] = T(0);
      }
    }
    gsync();
#pragma unroll 1
    for (int i = tid; i < m; i += NT) w[i] = rho[i] * z[i] - y[i];
    gsync();

    unsigned iter = 0;
    if (polish_only) { code = SFB_QP_OPTIMAL; iter = a.out_iter[b]; }
    else code = qp_stage_loop<T, G, NS, MS>(&a, b, c, t0, code, &iter);

    // active sets as polish_qp builds them (:113-123), ascending order, on the scaled dual
    int na = 0;
    int* idx = reinterpret_cast<int*>(mv1);
    T* bnd = w;
    {
      const T thr = T(100) * Num<T>::eps();
      int* wcnt = reinterpret_cast<int*>(part);
#pragma unroll 1
      for (int base = 0; base < m; base += NT) {
        const int i = base + tid;
        int act = 0;
        T bv = T(0);
        if (i < m) {
          if (polish_only) {  // the active set the lower-precision solve determined
            act = a.out_active[b * (long long)m + i];
            if (act < 0) bv = sy[i] * l[i];
            if (act > 0) bv = sy[i] * u[i];
          } else {
            if (y[i] < -thr && l[i] != -inf) { act = -1; bv = sy[i] * l[i]; }
            if (y[i] > thr && u[i] != inf) { act = 1; bv = sy[i] * u[i]; }
            if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
          }
        }
        const unsigned bal = __ballot_sync(kFullMask, act != 0);
        int before = 0, total = __popc(bal);
        if (G > 1) {
          if (lane == 0) wcnt[warp] = total;
          __syncthreads();
          total = 0;
#pragma unroll 1
          for (int wv = 0; wv < G; ++wv) {
            if (wv < warp) before += wcnt[wv];
            total += wcnt[wv];
          }
        }
        if (act != 0) {
          const int pos = na + before + __popc(bal & ((1u << lane) - 1u));
          idx[pos] = i;
          bnd[pos] = bv;
        }
        na += total;
        gsync();
      }
    }

    unsigned flags = 0;
    if (code == SFB_QP_OPTIMAL && a.prm.polish) flags = qp_stage_polish<T, G, NS, MS>(&a, b, c, na, gscratch);  // :515-539

    // unscale + objective  :544-548
#pragma unroll 1
    for (int j = tid; j < n; j += NT) {
      const T v = sx[j] * x[j];
      nv1[j] = v;
      a.out_x[b * (long long)n + j] = (TIO)v;
    }
#pragma unroll 1
    for (int i = tid; i < m; i += NT) a.out_y[b * (long long)m + i] = (TIO)(sy[i] * y[i] / c);
    gsync();
    T obj = T(0);
#pragma unroll 1
    for (int i = tid; i < n; i += NT) {
      T acc = T(0);
#pragma unroll 10
      for (int j = 0; j < n; ++j) acc += T(0.5) * (T)__ldg(gP + i + (long long)n * j) * nv1[j];
      obj += nv1[i] * (acc + q[i]);
    }
    obj = gsum(obj);
    if (tid == 0) {
      a.out_obj[b] = (TIO)obj;
      a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
      a.out_iter[b] = iter;
      if (a.out_flags) a.out_flags[b] = flags;
    }
    gsync();
  }
};

// ------------------------------------------------------------------------------------------------------
// out-of-line stages
// ------------------------------------------------------------------------------------------------------
template <typename T, int G, int NS, int MS> __device__ __forceinline__ QpGroup<T, G, NS, MS> qp_view(int n, int m, T c)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  QpGroup<T, G, NS, MS> s(reinterpret_cast<T*>(smem_raw), NS > 0 ? NS : n, MS > 0 ? MS : m);
  s.c = c;
  s.gsync();  // every thread is out of the caller's last reduction: the scratch parity may restart at 0
  return s;
}

template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ T qp_stage_load_scale(const QpArgs<T, TIO>* a, long long b)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(a->n, a->m, T(1));
  s.load(*a, b);
  if (a->prm.scaling) {
    s.scale();
  } else {
#pragma unroll 1
    for (int j = s.tid; j < s.n; j += 32 * G) s.sx[j] = T(1);
#pragma unroll 1
    for (int i = s.tid; i < s.m; i += 32 * G) s.sy[i] = T(1);
    s.c = T(1);
  }
  s.gsync();
  return s.c;
}

template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ int qp_stage_setup(const QpArgs<T, TIO>* a, T c)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(a->n, a->m, c);
  const int code = s.setup(*a);
  s.gsync();
  return code;
}

template <typename T, int G, int NS, int MS> __device__ __noinline__ bool qp_stage_gj(int n, int m, int where, int sz)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(n, m, T(1));
  const bool ok = s.gj_invert_reg(where == 0 ? s.Ms : s.As + sz, where == 0 ? s.ldN : s.ldA, sz, s.gjbuf);
  s.gsync();
  return ok;
}

// generic-pointer variant (matrix may live in global memory; shared scratch): sizes beyond the register blocking
template <typename T, int G, int NS, int MS> __device__ __noinline__ bool qp_stage_gj_generic(int n, int m, T* Mx, int ld, int sz)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(n, m, T(1));
  // row / column buffers: xt..nv2 hold 4n scalars; a Schur block has sz <= n, Ms itself has sz == n
  const bool ok = s.gj_invert_smem(Mx, ld, sz, s.xt, s.xt + sz);
  s.gsync();
  return ok;
}

template <typename T, int G, int NS, int MS, typename TIO> __device__ __noinline__ int qp_stage_check(const QpArgs<T, TIO>* a, long long b, T c)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(a->n, a->m, c);
  const int code = s.check_stopping(*a, a->P + b * (long long)a->n * a->n);
  s.gsync();
  return code;
}

template <typename T, int G, int NS, int MS, typename TIO>
__device__ __noinline__ unsigned qp_stage_polish(const QpArgs<T, TIO>* a, long long b, T c, int na, T* gscratch)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(a->n, a->m, c);
  const unsigned fl = s.polish(*a, a->P + b * (long long)a->n * a->n, na, reinterpret_cast<const int*>(s.mv1), s.w, gscratch);
  s.gsync();
  return fl;
}

template <typename T, int G, int NS, int MS, typename TIO>
__device__ __noinline__ int qp_stage_loop(const QpArgs<T, TIO>* a, long long b, T c, unsigned long long t0, int code, unsigned* iter_out)
{
  QpGroup<T, G, NS, MS> s = qp_view<T, G, NS, MS>(a->n, a->m, c);
  int r;
  if constexpr (QpGroup<T, G, NS, MS>::kRegLoop) r = s.admm_loop_reg(*a, b, t0, code, iter_out);
  else r = s.admm_loop(*a, b, t0, code, iter_out);
  s.gsync();
  return r;
}

// One CTA of G warps per instance; CTAs pull instances from a global work counter (iteration counts are
// heavy-tailed), so a slow instance never idles the rest of the grid.
template <typename T, int G, int MINB, int NS, int MS, typename TIO = T>
__global__ void __launch_bounds__(32 * G, MINB) qp_dense_group_kernel(const __grid_constant__ QpArgs<T, TIO> a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* base = reinterpret_cast<T*>(smem_raw);
  QpGroup<T, G, NS, MS> s(base, NS > 0 ? NS : a.n, MS > 0 ? MS : a.m);
  T* gscratch = a.scratch ? a.scratch + (long long)blockIdx.x * a.scratch_per_cta : nullptr;
  unsigned long long* slot = reinterpret_cast<unsigned long long*>(base + QpLayout(NS > 0 ? NS : a.n, MS > 0 ? MS : a.m, 32 * G).offSlot);
#pragma unroll 1
  for (;;) {
    unsigned long long b = 0;
    if (G == 1) {
      if (s.lane == 0) b = atomicAdd(a.work_counter, 1ull);
      b = __shfl_sync(kFullMask, b, 0);
    } else {
      if (threadIdx.x == 0) *slot = atomicAdd(a.work_counter, 1ull);
      __syncthreads();
      b = *slot;
      __syncthreads();
    }
    if ((long long)b >= a.batch) break;
    s.solve(a, (long long)b, gscratch);
  }
}

}  // namespace sfb
