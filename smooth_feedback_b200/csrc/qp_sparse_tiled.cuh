// qp_sparse_tiled.cuh -- batched sparse operator-splitting QP solver for sm_100a (shared sparsity pattern).
//
// Replaces the sparse branches of (pettni/smooth_feedback @ 9a08971)
//   QPSolver<QuadraticProgramSparse>::solve   include/smooth/feedback/qp_solver.hpp:343-568 (fill :380-397,
//                                             SimplicialLDLT factorize :424-426, permuted solve :456-460)
//   QPSolver::scale / check_stopping          :673-730 / :574-644 (InnerIterator walks only stored entries)
//   detail::polish_qp                         :92-204
// i.e. the call MPC::operator() makes at mpc.hpp:491.
//
// Layout.  A warp owns a TILE of TW instances (TW = 32, 8 or 4); every per-instance array of the working set lives in
// global memory as [tile][element][TW].  All lanes of a warp execute the same pattern-driven program (the index arrays
// are shared by the whole batch and broadcast), lane -> (instance = lane % TW, row-lane r = lane / TW): an access to
// element e touches TW consecutive scalars, so every load and store of the solve is a full, coalesced line and the
// per-iteration traffic is exactly the north star's model (Abar twice, the L D L^T factor twice, the iterate vectors;
// B_iter).  With TW < 32 the RL = 32 / TW lanes of an instance split the entries of a row between them
// (small batches: RL times more warps to hide the memory latency, which is what bounds this kernel), the solve vector
// lives in shared memory and the triangular sweeps run as a register-pipelined list of fixed-width steps; with TW = 32
// a lane owns its instance alone and the code is free of shuffles and barriers.
//
// The KKT system is reduced as in the dense kernel:  (Pbar + sigma I + Abar^T R Abar) xt = sigma x - qbar + Abar^T (R z - y),
// nu = R (Abar xt - z) + y; the n x n matrix is factorised L D L^T without pivoting (it is SPD) in the fill-reducing
// order computed on the host (qp_sparse_host.hpp).  Variables are kept in the permuted order throughout.  Every hot
// loop is in GATHER form (index mirrors built on the host) or updates provably distinct targets, so that a batch of
// kSpU independent loads is in flight per lane before the first one is consumed.
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
constexpr int kSpMV = 11;  // m-vectors
constexpr int kSpU = 8;    // independent loads kept in flight per lane in the gather / update loops

struct SpPattern
{
  int n, m, nnzP, nnzA, nnzL;
  const int *perm, *iperm, *P_rowp, *P_colp, *P_tgt, *A_rowptr, *A_col, *A_pair_ptr, *A_pair_tgt, *L_colptr, *L_row,
    *F_ptr, *F_tgt;
  // gather-form mirrors and packed pair lists (qp_sparse_host.hpp)
  const int *LR_ptr, *LR_col, *LR_slot, *AT_ptr, *AT_row, *AT_slot, *PR_ptr, *PR_col, *PR_slot, *PS_ptr, *PS_col, *PS_slot,
    *PC_ptr, *PC_slot, *LB_ptr, *LB_row, *LB_slot, *A_pair_ab, *F_ab;
  // padded sweep schedules (TW == 8): nFS / nBS steps of kSpStep entries each
  const int *FS_meta, *FS_col, *FS_slot, *BS_meta, *BS_col, *BS_slot;
  int nFS, nBS;
  // padded row / column streams of A (TW < 32; WR == 0: not available for this pattern)
  const int *RP_col, *RP_slot, *ATP_row, *ATP_slot;
  int WR, WA, m_pad, n_pad;
};

constexpr int kSpStep = 32;  // == SparseSymbolic::kStepWidth
constexpr int kSpDepth = 4;  // sweep steps in flight per lane
constexpr bool kSpPaddedF64 = false;  // fp64 at 128 registers: measured slower than the generic passes even with 4-wide chunks (91.8 vs 82.4 ms at cfg3)
constexpr int kSpPolishRefine = 0;  // extra refinement steps on Hp per application of the reduced polish solve (0, 1 and 2 give
                                    // the same agreement with the oracle on every parity workload; 0 is 10 % faster)

// scalars per instance of the W block: factor (nnzL + n) followed by its two stream-ordered copies
__host__ __device__ inline size_t sp_w_len(const SpPattern& p, int tw)
{
  const size_t copies = (tw < 32) ? (size_t)(p.nFS + p.nBS) * kSpStep + (size_t)p.nBS : 2 * (size_t)p.nnzL;
  return (size_t)p.nnzL + p.n + copies;
}
// scalars per instance of the A block: Abar (CSR order) followed by its padded row-stream and column-stream copies
__host__ __device__ inline size_t sp_a_len(const SpPattern& p, int tw)
{
  return (size_t)p.nnzA + ((tw < 32 && p.WR > 0) ? (size_t)p.m_pad * p.WR + (size_t)p.n_pad * p.WA : 0);
}
__host__ __device__ inline size_t sp_fwd_len(const SpPattern& p, int tw) { return (tw < 32) ? (size_t)p.nFS * kSpStep : (size_t)p.nnzL; }

// T: the scalar the kernel computes in; TIO: the scalar of the caller's arrays (TIO != T only in the mixed-precision
// polish pass, mode 2: T = double over float data).
template <typename T, typename TIO = T> struct SpArgs
{
  SpPattern pat;
  // inputs, array-of-instances:  P_vals [batch][nnzP], q [batch][n], A_vals [batch][nnzA], l,u [batch][m]
  const TIO *P, *q, *A, *l, *u, *warm_x, *warm_y;
  TIO *out_x, *out_y, *out_obj;
  int32_t* out_status;
  uint32_t* out_iter;
  int8_t* out_active;
  uint32_t* out_flags;
  // tiled workspace: [tile][len][TW]
  T *wsA, *wsP, *wsW, *wsN, *wsM;
  long long batch;
  sfb_qp_params prm;
  unsigned max_iter_eff;
  int dinf_guard;  // see QpArgs::dinf_guard
  int mode;  // 0 = solve, 2 = polish only (instances already solved in lower precision: out_* hold the unpolished result)
};

// element e of this lane's instance inside a [len][TW] tile block
template <typename T, int TW> struct TV
{
  T* p;
  __device__ __forceinline__ T& operator[](int e) const { return p[(size_t)e * TW]; }
};

template <typename T, int TW> struct SpSolver
{
  static constexpr int RL = 32 / TW;  // lanes cooperating on one instance
  using V = TV<T, TW>;
  const SpPattern& S;
  int n, m, r;
  unsigned gmask;
  V A, P, W, LRW, LBW;  // LRW / LBW: the factor's values again, in the forward / backward sweep's stream order
  V LBD;                // 1 / D of the row of every backward step (TW < 32)
  V APW, ATW;           // Abar again: rows padded to WR entries, columns padded to WA entries (TW < 32)
  V lo, hi;             // scaled bounds sy l, sy u
  V q, qb, x, xold, v, sx, t1, t2, t3, t4;  // n-vectors (permuted order)
  V l, u, sy, rho, rinv, z, y, yold, w;    // m-vectors
  T c;
  int inst_ = 0;    // instance slot inside the tile

  // smem_v: [n][TW] scalars of shared memory for the solve vector (TW == 8 only, nullptr otherwise)
  template <typename TIO> __device__ SpSolver(const SpArgs<T, TIO>& a, long long tile, int lane, T* smem_v) : S(a.pat), n(a.pat.n), m(a.pat.m)
  {
    const int inst = lane & (TW - 1);
    inst_ = inst;
    r = lane / TW;
    gmask = 0u;
#pragma unroll
    for (int k = 0; k < RL; ++k) gmask |= 1u << (inst + k * TW);
    A.p = a.wsA + (size_t)tile * sp_a_len(S, TW) * TW + inst;
    APW.p = A.p + (size_t)S.nnzA * TW;
    ATW.p = APW.p + (size_t)S.m_pad * S.WR * TW;
    P.p = a.wsP + (size_t)tile * S.nnzP * TW + inst;
    W.p = a.wsW + (size_t)tile * sp_w_len(S, TW) * TW + inst;
    LRW.p = W.p + (size_t)(S.nnzL + n) * TW;
    LBW.p = LRW.p + sp_fwd_len(S, TW) * TW;
    LBD.p = LBW.p + (size_t)S.nBS * kSpStep * TW;
    T* nv = a.wsN + (size_t)tile * kSpNV * n * TW + inst;
    T* mv = a.wsM + (size_t)tile * kSpMV * m * TW + inst;
    auto N = [&](int k) { return V{nv + (size_t)k * n * TW}; };
    auto M = [&](int k) { return V{mv + (size_t)k * m * TW}; };
    q = N(0); qb = N(1); x = N(2); xold = N(3); v = N(4); sx = N(5); t1 = N(6); t2 = N(7); t3 = N(8); t4 = N(9);
    l = M(0); u = M(1); sy = M(2); rho = M(3); rinv = M(4); z = M(5); y = M(6); yold = M(7); w = M(8); lo = M(9); hi = M(10);
    if (smem_v != nullptr) v.p = smem_v + inst;  // the solve vector is gathered from ~40 times per row sweep: keep it on chip
    c = T(1);
  }

  // ---------------------------------------------------------------- group primitives (the RL lanes of one instance)
  __device__ __forceinline__ void gsync() const
  {
    if (RL > 1) __syncwarp(gmask);
  }
  __device__ __forceinline__ T gsum(T val) const
  {
#pragma unroll
    for (int o = TW; o < 32; o <<= 1) val += __shfl_xor_sync(gmask, val, o);
    return val;
  }
  __device__ __forceinline__ T gmax(T val) const
  {
#pragma unroll
    for (int o = TW; o < 32; o <<= 1) val = fmax(val, __shfl_xor_sync(gmask, val, o));
    return val;
  }
  __device__ __forceinline__ bool gany(bool pr) const
  {
    int val = pr ? 1 : 0;
#pragma unroll
    for (int o = TW; o < 32; o <<= 1) val |= __shfl_xor_sync(gmask, val, o);
    return val != 0;
  }

  // sum over e = e0 + first, e0 + first + step, ... < e1 of val(e) * vec(idx[e]); a batch of kSpU entries is loaded
  // before any of them is used
  template <class VAL, class VEC>
  __device__ __forceinline__ T gather_sum(int e0, int e1, int first, int step, const int* __restrict__ idx, VAL val, VEC vec) const
  {
    T acc = T(0);
    for (int e = e0 + first; e < e1; e += kSpU * step) {
      int j[kSpU];
      T a[kSpU], b[kSpU];
#pragma unroll
      for (int k = 0; k < kSpU; ++k) j[k] = (e + k * step < e1) ? idx[e + k * step] : 0;
#pragma unroll
      for (int k = 0; k < kSpU; ++k) a[k] = (e + k * step < e1) ? val(e + k * step) : T(0);
#pragma unroll
      for (int k = 0; k < kSpU; ++k) b[k] = (e + k * step < e1) ? vec(j[k]) : T(0);
#pragma unroll
      for (int k = 0; k < kSpU; ++k)
        if (e + k * step < e1) acc += a[k] * b[k];
    }
    return acc;
  }
  // max over e in [e0, e1) of f(e), batched the same way
  template <class F> __device__ __forceinline__ T gather_max(T g, int e0, int e1, F f) const
  {
    for (int e = e0; e < e1; e += kSpU) {
      T a[kSpU];
#pragma unroll
      for (int k = 0; k < kSpU; ++k) a[k] = (e + k < e1) ? f(e + k) : T(0);
#pragma unroll
      for (int k = 0; k < kSpU; ++k) g = fmax(g, a[k]);
    }
    return g;
  }
  // W[tgt[p]] += delta(p) for p = p0 + r, p0 + r + RL, ... < p1.  The targets of one call are DISTINCT (pairs of one
  // row of A / one column of L), so a batch is loaded, updated and stored without intermediate dependencies.
  template <class F> __device__ __forceinline__ void rmw_flat(int p0, int p1, const int* __restrict__ tgt, F delta) const
  {
    const V Wl = W;
    for (int p = p0 + r; p < p1; p += kSpU * RL) {
      int t[kSpU];
      T d[kSpU], o[kSpU];
#pragma unroll
      for (int k = 0; k < kSpU; ++k) t[k] = (p + k * RL < p1) ? tgt[p + k * RL] : 0;
#pragma unroll
      for (int k = 0; k < kSpU; ++k) d[k] = (p + k * RL < p1) ? delta(p + k * RL) : T(0);
#pragma unroll
      for (int k = 0; k < kSpU; ++k) o[k] = (p + k * RL < p1) ? Wl[t[k]] : T(0);
#pragma unroll
      for (int k = 0; k < kSpU; ++k)
        if (p + k * RL < p1) Wl[t[k]] = o[k] + d[k];
    }
  }

  // ---------------------------------------------------------------- QPSolver::scale, qp_solver.hpp:673-730
  // Column / row maxima in gather form (PC_*, AT_* mirrors, CSR rows): max() is exact and order independent and the
  // products are formed in the reference's order, so sx, sy agree with the CPU restatement bit for bit (c too: every
  // lane sums the column maxima in the original column order).
  __device__ void scale()
  {
    const V Pl = P, Al = A, sxl = sx, syl = sy;
    for (int j = r; j < n; j += RL) sx[j] = T(1);
    for (int i = r; i < m; i += RL) sy[i] = T(1);
    for (int pj = r; pj < n; pj += RL)
      t1[pj] = gather_max(T(0), S.PC_ptr[pj], S.PC_ptr[pj + 1], [&](int e) { return fabs(Pl[S.PC_slot[e]]); });
    gsync();
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
    const T cc = c;
    int it = 0;
    T dev;
    do {
      gsync();
      for (int pj = r; pj < n; pj += RL) {
        const T sxj = sxl[pj];
        T g = gather_max(T(0), S.PC_ptr[pj], S.PC_ptr[pj + 1], [&](int e) {
          const int sl = S.PC_slot[e];
          return fabs(((cc * sxl[S.P_rowp[sl]]) * sxj) * Pl[sl]);
        });
        g = gather_max(g, S.AT_ptr[pj], S.AT_ptr[pj + 1], [&](int e) { return fabs((syl[S.AT_row[e]] * sxj) * Al[S.AT_slot[e]]); });
        t1[pj] = g;
      }
      for (int i = r; i < m; i += RL) {
        const T syi = syl[i];
        w[i] = gather_max(T(0), S.A_rowptr[i], S.A_rowptr[i + 1], [&](int e) { return fabs((syi * sxl[S.A_col[e]]) * Al[e]); });
      }
      gsync();  // every maximum is formed before any scale factor changes
      dev = T(0);
      for (int j = r; j < n; j += RL) {
        T g = t1[j];
        if (g == T(0)) g = T(1);
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];
        dev = fmax(dev, fabs(g - T(1)));
      }
      for (int i = r; i < m; i += RL) {
        T g = w[i];
        if (g == T(0)) g = T(1);
        sy[i] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[i];
        dev = fmax(dev, fabs(g - T(1)));
      }
      dev = gmax(dev);
    } while (it++ < 10 && dev > T(0.1));
    gsync();
  }

  // ---------------------------------------------------------------- W <- shift I + c Sx triu(P) Sx + Abar^T diag(wt) Abar
  // (lower triangle in the factor's slots, diagonal in W[nnzL + i]); wt = rho for the ADMM system
  __device__ void assemble(T shift, const V& wt)
  {
    const int nW = S.nnzL + n;
    const V Al = A;
    for (int e = r; e < S.nnzL; e += RL) W[e] = T(0);
    for (int e = S.nnzL + r; e < nW; e += RL) W[e] = shift;
    gsync();
    if (r == 0) {  // compressed P: distinct targets; kept on one lane, nnzP is small
      for (int e = 0; e < S.nnzP; e += kSpU) {
        int t[kSpU];
        T d[kSpU], o[kSpU];
#pragma unroll
        for (int k = 0; k < kSpU; ++k) t[k] = (e + k < S.nnzP) ? S.P_tgt[e + k] : -1;
#pragma unroll
        for (int k = 0; k < kSpU; ++k)
          d[k] = (t[k] >= 0) ? ((c * sx[S.P_rowp[e + k]]) * sx[S.P_colp[e + k]]) * P[e + k] : T(0);  // qp_solver.hpp:386
#pragma unroll
        for (int k = 0; k < kSpU; ++k) o[k] = (t[k] >= 0) ? W[t[k]] : T(0);
#pragma unroll
        for (int k = 0; k < kSpU; ++k)
          if (t[k] >= 0) W[t[k]] = o[k] + d[k];
      }
    }
    gsync();
    for (int i = 0; i < m; ++i) {
      const T ri = wt[i];
      if (ri != T(0)) {  // group-uniform
        const int e0 = S.A_rowptr[i];
        rmw_flat(S.A_pair_ptr[i], S.A_pair_ptr[i + 1], S.A_pair_tgt, [&](int p) {
          const int ab = S.A_pair_ab[p];
          return (ri * Al[e0 + (ab >> 16)]) * Al[e0 + (ab & 0xffff)];
        });
        gsync();
      }
    }
  }

  // right-looking L D L^T in place; afterwards W[e] = L entries, W[nnzL + k] = 1 / D_k.  False on a non-positive pivot.
  __device__ bool factor()
  {
    bool ok = true;
    const int nL = S.nnzL;
    const V Wl = W;
    for (int k = 0; k < n; ++k) {
      const T dk = W[nL + k];
      if (!(dk > T(0)) || !(dk < Num<T>::inf())) ok = false;
      const T dinv = T(1) / dk;
      const int c0 = S.L_colptr[k], c1 = S.L_colptr[k + 1];
      rmw_flat(S.F_ptr[k], S.F_ptr[k + 1], S.F_tgt, [&](int p) {
        const int ab = S.F_ab[p];
        return -(Wl[c0 + (ab & 0xffff)] * (Wl[c0 + (ab >> 16)] * dinv));
      });
      gsync();
      for (int a = c0 + r; a < c1; a += RL) W[a] *= dinv;
      if (r == 0) W[nL + k] = dinv;
      gsync();
    }
    // stream-order copies for the sweeps (contiguous, no slot indirection, prefetchable); TW == 8: padded step layout
    auto copy_stream = [&](const V& dst, const int* __restrict__ slot, int len) {
      for (int e = r; e < len; e += kSpU * RL) {
        int sl[kSpU];
        T a0[kSpU];
#pragma unroll
        for (int k = 0; k < kSpU; ++k) sl[k] = (e + k * RL < len) ? slot[e + k * RL] : -1;
#pragma unroll
        for (int k = 0; k < kSpU; ++k) a0[k] = (sl[k] >= 0) ? Wl[sl[k]] : T(0);
#pragma unroll
        for (int k = 0; k < kSpU; ++k)
          if (e + k * RL < len) dst[e + k * RL] = a0[k];
      }
    };
    if (RL > 1) {
      copy_stream(LRW, S.FS_slot, S.nFS * kSpStep);
      copy_stream(LBW, S.BS_slot, S.nBS * kSpStep);
      for (int st = r; st < S.nBS; st += RL) LBD[st] = W[nL + (S.BS_meta[st] >> 1)];
    } else {
      copy_stream(LRW, S.LR_slot, nL);
      copy_stream(LBW, S.LB_slot, nL);
    }
    gsync();
    return ok;
  }

  // The solve vector through the shared-memory SYMBOL (TW < 32): via the member pointer the compiler only sees a generic
  // address and emits generic loads, which cost more latency and are tracked like global loads; derived from the extern
  // __shared__ array they are LDS / STS.
  __device__ __forceinline__ T* v_shared() const
  {
    extern __shared__ __align__(16) unsigned char sp_smem_raw[];
    return reinterpret_cast<T*>(sp_smem_raw) + inst_;
  }

  // One sweep of the triangular solve in gather form over the stream-ordered factor copy LW:
  //   forward:  v[k] <- v[k] - sum_e LW[e] v[col[e]]          backward:  v[k] <- v[k] / D_k - sum_e LW[e] v[col[e]]
  // TW == 8: the sweep is a list of fixed-width STEPS (kSpStep padded entries, RL lanes x kSpU each); a row of L spans one
  // or more consecutive steps.  Factor entries and column indices do not depend on the solve's dependency chain, so they
  // are fetched TWO steps ahead into registers: what is left on the chain per step is the gather from v (shared memory),
  // the RL-lane shuffles and one store.
  static constexpr int SU = (RL > 1) ? kSpStep / RL : kSpU;  // entries of a step per lane
  struct StepBuf
  {
    int meta;
    int j[SU];
    T a[SU];
    T dk;
  };
  // The sweep is instruction-bound per warp (one warp's ~100 dependent instructions per step, not memory latency, set
  // the pace: measured 1350 cycles per step), so the step is kept lean: every address is a running pointer plus an
  // immediate, nothing depends on another load, the step lists are padded to a multiple of kSpDepth (no bounds checks).
  template <bool FWD> __device__ __forceinline__ void step_load(StepBuf& B, const int* __restrict__ pm, const int* __restrict__ pc,
                                                                const T* __restrict__ pa, const T* __restrict__ pd) const
  {
    B.meta = *pm;
#pragma unroll
    for (int t = 0; t < SU; ++t) {
      B.j[t] = pc[t * RL];
      B.a[t] = pa[(size_t)t * RL * TW];
    }
    B.dk = FWD ? T(1) : *pd;
  }
  // shared-memory accesses with 32-bit addresses (one shift-add per gather instead of 64-bit pointer arithmetic)
  static __device__ __forceinline__ T lds(unsigned addr)
  {
    T val;
    if constexpr (sizeof(T) == 8) asm volatile("ld.shared.f64 %0, [%1];" : "=d"(val) : "r"(addr));
    else asm volatile("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(addr));
    return val;
  }
  static __device__ __forceinline__ void sts(unsigned addr, T val)
  {
    if constexpr (sizeof(T) == 8) asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(val) : "memory");
    else asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(val) : "memory");
  }
  template <bool FWD> __device__ __forceinline__ void step_apply(const StepBuf& B, T& acc, unsigned sv)
  {
    constexpr unsigned kStride = TW * sizeof(T);
    const unsigned ak = sv + (unsigned)(B.meta >> 1) * kStride;
    const T vk = (r == 0) ? lds(ak) : T(0);  // read and written by the row's first lane only; rhs row k is untouched until here
#pragma unroll
    for (int t = 0; t < SU; ++t) acc += B.a[t] * lds(sv + (unsigned)B.j[t] * kStride);  // padding: a = 0, j = n (dummy zero slot)
    if (B.meta & 1) {  // last step of the row (warp-uniform)
      const T s = gsum(acc);
      if (r == 0) sts(ak, FWD ? vk - s : vk * B.dk - s);
      acc = T(0);
      gsync();
    }
  }
  // kSpDepth step buffers rotate through a loop unrolled kSpDepth times (static register indexing, no moves)
  template <bool FWD> __device__ void sweep_steps(int nsteps, const int* __restrict__ meta, const int* __restrict__ col, const V& LW)
  {
    if (nsteps == 0) return;  // nsteps is a multiple of kSpDepth (host padding)
    const int* pm = meta;
    const int* pc = col + r;
    const T* pa = LW.p + (size_t)r * TW;
    const T* pd = LBD.p;
    StepBuf b[kSpDepth];
#pragma unroll
    for (int d = 0; d < kSpDepth; ++d) {
      step_load<FWD>(b[d], pm, pc, pa, pd);
      pm += 1; pc += kSpStep; pa += (size_t)kSpStep * TW; pd += TW;
    }
    T acc = T(0);
    const unsigned sv = (unsigned)__cvta_generic_to_shared(v_shared());
    int st = 0;
    for (; st + kSpDepth < nsteps; st += kSpDepth) {
#pragma unroll
      for (int d = 0; d < kSpDepth; ++d) {
        step_apply<FWD>(b[d], acc, sv);
        step_load<FWD>(b[d], pm, pc, pa, pd);
        pm += 1; pc += kSpStep; pa += (size_t)kSpStep * TW; pd += TW;
      }
    }
#pragma unroll
    for (int d = 0; d < kSpDepth; ++d) step_apply<FWD>(b[d], acc, sv);
  }
  // TW == 32: one lane per instance, rows walked with batched gathers over the contiguous copies
  template <bool FWD> __device__ void sweep_rows(const int* __restrict__ ptr, const int* __restrict__ col, const V& LW)
  {
    const V vv = v;
    for (int kk = 0; kk < n; ++kk) {
      const int k = FWD ? kk : n - 1 - kk;
      const int e0 = ptr[kk], e1 = ptr[kk + 1];
      if (FWD && e0 == e1) continue;
      const T vk = v[k], dk = FWD ? T(1) : W[S.nnzL + k];
      const T s = gather_sum(e0, e1, 0, 1, col, [&](int e) { return LW[e]; }, [&](int j) { return vv[j]; });
      v[k] = FWD ? vk - s : vk * dk - s;
    }
  }

  // v <- (L D L^T)^-1 v
  __device__ void solve()
  {
    if (RL > 1) {
      sweep_steps<true>(S.nFS, S.FS_meta, S.FS_col, LRW);
      sweep_steps<false>(S.nBS, S.BS_meta, S.BS_col, LBW);
    } else {
      sweep_rows<true>(S.LR_ptr, S.LR_col, LRW);
      sweep_rows<false>(S.LB_ptr, S.LB_row, LBW);
    }
  }

  // out[j] = fin(j, sum_i Abar_ij in_i)  (column gather through the AT_* mirror; columns are dealt to the row-lanes)
  template <class FIN> __device__ void At_gather(const V& in, const V& out, FIN fin)
  {
    const V Al = A;
    for (int j = r; j < n; j += RL)
      out[j] = fin(j, gather_sum(S.AT_ptr[j], S.AT_ptr[j + 1], 0, 1, S.AT_row, [&](int e) { return Al[S.AT_slot[e]]; },
                                 [&](int i) { return in[i]; }));
    gsync();
  }
  // out[j] = fin(j, sum_k sym(Pbar)_jk in_k)  (upper triangle of c Sx P Sx mirrored, PS_* mirror)
  template <class FIN> __device__ void Psym_gather(const V& in, const V& out, FIN fin)
  {
    const V Pl = P, sxl = sx;
    const T cc = c;
    for (int j = r; j < n; j += RL)
      out[j] = fin(j, gather_sum(S.PS_ptr[j], S.PS_ptr[j + 1], 0, 1, S.PS_col,
                                 [&](int e) { const int sl = S.PS_slot[e]; return ((cc * sxl[S.P_rowp[sl]]) * sxl[S.P_colp[sl]]) * Pl[sl]; },
                                 [&](int k) { return in[k]; }));
    gsync();
  }
  // out[j] = sum_k (sc P_jk) in_k with the entries as stored (PR_* mirror)
  __device__ void P_gather(const V& in, const V& out, T sc)
  {
    const V Pl = P;
    for (int j = r; j < n; j += RL)
      out[j] = gather_sum(S.PR_ptr[j], S.PR_ptr[j + 1], 0, 1, S.PR_col, [&](int e) { return sc * Pl[S.PR_slot[e]]; },
                          [&](int k) { return in[k]; });
    gsync();
  }
  // sum_j Abar_ij vec_j for row i (whole row on the calling lane)
  __device__ __forceinline__ T A_row_dot(int i, const V& vec) const
  {
    const V Al = A;
    return gather_sum(S.A_rowptr[i], S.A_rowptr[i + 1], 0, 1, S.A_col, [&](int e) { return Al[e]; }, [&](int j) { return vec[j]; });
  }

  // ---------------------------------------------------------------- pipelined SpMV passes of the ADMM iteration (TW < 32)
  // Rows (columns) of Abar are dealt round-robin to the RL lanes of an instance and padded to a fixed width, so every
  // address is a running pointer plus an immediate; the operands of the NEXT row are requested before the current one is
  // reduced (two register buffers, loop unrolled twice).
  // fp64 at 128 registers spills in these loops and ends up slower than the generic passes (89.6 vs 81.7 ms at cfg3),
  // fp32 gains 45 % (128k -> 186k solves/s): enabled for single precision only
  __device__ __forceinline__ bool padded_passes() const { return RL > 1 && S.WR > 0 && (sizeof(T) == 4 || kSpPaddedF64); }

  // A padded row (column) is a run of CHUNKS of kSpU entries; the flattened chunk stream of a lane is walked with two
  // register buffers (loop unrolled twice): the operands of the next chunk are requested before the current one is used.
  static constexpr int CW = (sizeof(T) == 8) ? 4 : 8;  // chunk width: fp64 at 128 registers cannot afford two 8-wide buffers
  struct Chunk
  {
    int j[CW];
    T a[CW];
  };
  static __device__ __forceinline__ void chunk_load(Chunk& B, const int*& pc, const T*& pa)
  {
#pragma unroll
    for (int t = 0; t < CW; ++t) {
      B.j[t] = pc[t];
      B.a[t] = pa[(size_t)t * TW];
    }
    pc += CW;
    pa += (size_t)CW * TW;
  }
  // walk `nslots` padded rows of `width` entries each (lane r owns rows r, r + RL, ...): dot(B, acc) accumulates one
  // chunk, fin(slot, acc) closes a row
  template <class DOT, class FIN>
  __device__ __forceinline__ void walk_chunks(int nslots, int width, const int* idx, const T* val, DOT dot, FIN fin)
  {
    const int cpr = width / CW;            // chunks per row (the host pads widths to a multiple of 8)
    const int total = nslots * cpr;        // chunks of this lane
    const int* pc = idx + (size_t)r * width;
    const T* pa = val + (size_t)r * width * TW;
    const int skip = (RL - 1) * width;     // from the end of this lane's row to the start of its next one
    Chunk b0, b1;
    int c_in_row = 0, slot = 0;
    auto advance_row = [&](int& cload) {   // called after loading a chunk: jump to the lane's next row at a row end
      if (++cload == cpr) { cload = 0; pc += skip; pa += (size_t)skip * TW; }
    };
    int cload = 0;
    chunk_load(b0, pc, pa); advance_row(cload);
    T acc = T(0);
    int q = 0;
    auto consume = [&](const Chunk& B) {
      dot(B, acc);
      if (++c_in_row == cpr) { fin(slot, acc); acc = T(0); c_in_row = 0; ++slot; }
    };
    for (; q + 2 < total; q += 2) {
      chunk_load(b1, pc, pa); advance_row(cload);
      consume(b0);
      chunk_load(b0, pc, pa); advance_row(cload);
      consume(b1);
    }
    for (; q < total; ++q) {  // at most two chunks left
      if (q + 1 < total) { chunk_load(b1, pc, pa); advance_row(cload); }
      consume(b0);
      b0 = b1;
    }
  }

  // v[j] = sigma x[j] - qb[j] + sum_i Abar_ij w_i      (qp_solver.hpp:450-451, reduced system)
  __device__ void rhs_pass(T sigma)
  {
    const V wl = w;
    walk_chunks(S.n_pad / RL, S.WA, S.ATP_row, ATW.p,
                [&](const Chunk& B, T& acc) {
                  T g[CW];
#pragma unroll
                  for (int t = 0; t < CW; ++t) g[t] = wl[B.j[t]];  // padding: a = 0, row 0
#pragma unroll
                  for (int t = 0; t < CW; ++t) acc += B.a[t] * g[t];
                },
                [&](int slot, T acc) {
                  const int j = r + slot * RL;
                  if (j < n) v[j] = sigma * x[j] - qb[j] + acc;
                });
    gsync();
  }

  // zt = Abar_i . xt, then the z / y / w updates of row i      (qp_solver.hpp:470-477)
  __device__ void row_pass(T alpha, T alpha_comp, bool chk)
  {
    const unsigned sv = (unsigned)__cvta_generic_to_shared(v_shared());
    constexpr unsigned kStride = TW * sizeof(T);
    walk_chunks(S.m_pad / RL, S.WR, S.RP_col, APW.p,
                [&](const Chunk& B, T& acc) {
#pragma unroll
                  for (int t = 0; t < CW; ++t) acc += B.a[t] * lds(sv + (unsigned)B.j[t] * kStride);  // padding: a = 0, column n (dummy zero slot)
                },
                [&](int slot, T zt) {
                  const int i = r + slot * RL;
                  if (i < m) {
                    const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];
                    if (chk) yold[i] = yi;
                    const T nu = ri * (zt - zi) + yi;
                    T zn = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;  // :471-474
                    zn = fmax(zn, lo[i]);
                    zn = fmin(zn, hi[i]);
                    const T yn = alpha_comp * yi + alpha * nu + ri * zi - ri * zn;  // :475-477
                    y[i] = yn;
                    z[i] = zn;
                    w[i] = ri * zn - yn;
                  }
                });
    gsync();
  }

  // ---------------------------------------------------------------- check_stopping, qp_solver.hpp:574-644
  // Same evaluation as the dense kernel (qp_dense_group.cuh::check_stopping): A x_us = Sy^-1 (Abar x), A^T y_us = Sx^-1 Abar^T y / c.
  __device__ int check_stopping(const sfb_qp_params& prm, bool dinf_guard)
  {
    const T eps_abs = T(prm.eps_abs), eps_rel = T(prm.eps_rel);
    const T eps_pinf = T(prm.eps_primal_inf), eps_dinf = T(prm.eps_dual_inf);
    const T inf = Num<T>::inf();
    T qn = T(0), dxn = T(0), qdx = T(0), Edy = T(0);
    for (int j = r; j < n; j += RL) {
      const T xj = x[j];
      const T d = xj - xold[j];
      t1[j] = sx[j] * xj;       // x_us   :481
      t2[j] = d;                // scaled dx
      const T dus = sx[j] * d;  // dx_us  :484
      xold[j] = dus;
      qn = fmax(qn, fabs(q[j]));
      dxn = fmax(dxn, fabs(dus));
    }
    gsync();
    for (int jo = r; jo < n; jo += RL) {  // q . dx_us in the original variable order (exactly so when RL == 1)
      const int j = S.iperm[jo];
      qdx += q[j] * xold[j];
    }
    for (int i = r; i < m; i += RL) {
      const T d = y[i] - yold[i];
      w[i] = d;
      Edy = fmax(Edy, fabs(sy[i] * d / c));  // :485
    }
    qn = gmax(qn); dxn = gmax(dxn); Edy = gmax(Edy); qdx = gsum(qdx);
    gsync();
    T n_Ax = T(0), n_r = T(0), n_z = T(0), s_pinf = T(0);
    bool pinf_blocked = false, dinf_rows_ok = true;
    for (int i = r; i < m; i += RL) {
      T ax = A_row_dot(i, x), adx = A_row_dot(i, t2);
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
    n_Ax = gmax(n_Ax); n_r = gmax(n_r); n_z = gmax(n_z); s_pinf = gsum(s_pinf);
    pinf_blocked = gany(pinf_blocked);
    dinf_rows_ok = !gany(!dinf_rows_ok);
    if (pinf_blocked) s_pinf = inf;
    gsync();
    // Abar^T y -> v, Abar^T dy -> t2 (scaled dx is dead); P x_us -> t3, P dx_us -> t4 with the entries as stored
    At_gather(y, v, [](int, T sum) { return sum; });
    At_gather(w, t2, [](int, T sum) { return sum; });
    P_gather(t1, t3, T(1));
    P_gather(xold, t4, T(1));
    T n_Px = T(0), n_Aty = T(0), n_res = T(0), n_Atdy = T(0), n_Pdx = T(0);
    for (int j = r; j < n; j += RL) {
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
    n_Px = gmax(n_Px); n_Aty = gmax(n_Aty); n_res = gmax(n_res); n_Atdy = gmax(n_Atdy); n_Pdx = gmax(n_Pdx);
    gsync();
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {  // :584-594
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;  // :619
    // dx == 0 guard: see qp_dense_group.cuh / DESIGN.md ("deliberate deviations")
    if ((dxn > T(0) || !dinf_guard) && (n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  // ---------------------------------------------------------------- detail::polish_qp, qp_solver.hpp:92-204
  // The regularised system  Hp s = r,  Hp = [Pbar + delta I, Aa^T; Aa, -delta I],  is solved through the SAME symbolic
  // factor: eliminating the (2,2) block gives (Pbar + delta I + Aa^T Aa / delta) s1 = r1 + Aa^T r2 / delta,
  // s2 = (Aa s1 - r2) / delta, whose pattern is a subset of M's.  That reduced matrix is ill conditioned (1/delta = 1e6
  // against delta), so an application of Hp^-1 can be followed by kSpPolishRefine steps of iterative refinement on Hp
  // itself.  Measured: 0, 1 and 2 steps are indistinguishable on every parity workload -- the reference's outer iteration
  // t += Hp^-1 (h - H t) is itself a refinement (contraction ~1e-4 per sweep on MPC problems) and absorbs the error of the
  // reduced solve -- so none is taken; agreement with the oracle stays at ~1e-9 or the oracle's own conditioning.
  // On entry: w[i] = 1 for active rows else 0, yold = scaled active bound.  Uses rho, rinv, z, l, u, xold, t1..t3, v as
  // scratch.  Every row-space vector of the polish is kept EXACTLY zero on inactive rows, so the gathers need no mask.

  // in: r1 in v, r2 in R2 -> out: s1 in v, s2 in R2 (in place)
  __device__ void polish_reduced_solve(const V& R2, T dinv)
  {
    const V vv = v;
    At_gather(R2, v, [&](int j, T sum) { return vv[j] + dinv * sum; });
    solve();
    for (int i = r; i < m; i += RL) {
      const T as = A_row_dot(i, v);
      R2[i] = (w[i] != T(0)) ? (as - R2[i]) * dinv : T(0);
    }
    gsync();
  }

  // out1 (n) = r1 - (sym(Pbar) X + shift1 X + Aa^T Y),  out2 (m) = r2 - (Aa X + shift2 Y) on active rows, 0 elsewhere
  __device__ void polish_residual(const V& r1, const V& r2, const V& X, const V& Y, T shift1, T shift2, const V& out1,
                                  const V& out2)
  {
    Psym_gather(X, out1, [&](int j, T sum) { return r1[j] - shift1 * X[j] - sum; });
    const V o1 = out1;
    At_gather(Y, out1, [&](int j, T sum) { return o1[j] - sum; });
    for (int i = r; i < m; i += RL) {
      const T at = A_row_dot(i, X);
      out2[i] = (w[i] != T(0)) ? r2[i] - (at + shift2 * Y[i]) : T(0);
    }
    gsync();
  }

  __device__ unsigned polish(const sfb_qp_params& prm)
  {
    const T delta = T(prm.delta), dinv = T(1) / delta;
    for (int i = r; i < m; i += RL) rho[i] = (w[i] != T(0)) ? dinv : T(0);
    gsync();
    assemble(delta, rho);
    if (!factor()) return SFB_QP_FLAG_POLISH_FAILED;
    // t = (t1 [n], rinv [m]);  h = (-qb, bnd) in (xold, yold)
    for (int j = r; j < n; j += RL) { t1[j] = T(0); xold[j] = -qb[j]; }
    for (int i = r; i < m; i += RL) rinv[i] = T(0);
    gsync();
    for (unsigned it = 0; it < prm.polish_iter; ++it) {
      // r = h - H t -> (t2, z)     (H = Hp without the delta blocks)
      polish_residual(xold, yold, t1, rinv, T(0), T(0), t2, z);
      // s = Hp^-1 r -> (t3, l), refined twice against Hp = H + diag(delta I, -delta I)
      for (int j = r; j < n; j += RL) v[j] = t2[j];
      for (int i = r; i < m; i += RL) l[i] = z[i];
      gsync();
      polish_reduced_solve(l, dinv);
      for (int j = r; j < n; j += RL) t3[j] = v[j];
      gsync();
      for (int rf = 0; rf < kSpPolishRefine; ++rf) {
        polish_residual(t2, z, t3, l, delta, -delta, v, u);
        polish_reduced_solve(u, dinv);
        for (int j = r; j < n; j += RL) t3[j] += v[j];
        for (int i = r; i < m; i += RL) l[i] += u[i];
        gsync();
      }
      T dmax = T(0), tmax = T(0);
      for (int j = r; j < n; j += RL) {
        const T tn = t1[j] + t3[j];
        dmax = fmax(dmax, fabs(t3[j]));
        tmax = fmax(tmax, fabs(tn));
        t1[j] = tn;
      }
      for (int i = r; i < m; i += RL) {
        const T tn = rinv[i] + l[i];
        dmax = fmax(dmax, fabs(l[i]));
        tmax = fmax(tmax, fabs(tn));
        rinv[i] = tn;
      }
      dmax = gmax(dmax);
      tmax = gmax(tmax);
      gsync();
      // once the correction is negligible (1e-12 relative: the ill-conditioned reduced solves have a noise floor of
      // ~1e-13, parity is demanded at 1e-6 and observed at ~1e-9) the remaining sweeps of the reference's fixed iteration
      // count are no-ops up to rounding (well-conditioned systems contract by ~delta / lambda_min per sweep): stop early,
      // as the dense kernel does
      if (dmax <= T(1e-12) * tmax) break;
    }
    bool bad = false;
    for (int j = r; j < n; j += RL) bad = bad || !(fabs(t1[j]) < Num<T>::inf());
    if (gany(bad)) return SFB_QP_FLAG_POLISH_FAILED;
    for (int j = r; j < n; j += RL) x[j] = t1[j];  // :199
    for (int i = r; i < m; i += RL)
      if (w[i] != T(0)) y[i] = rinv[i];  // :200-201 (other duals unchanged)
    gsync();
    return SFB_QP_FLAG_POLISHED;
  }

  // ---------------------------------------------------------------- QPSolver::solve, qp_solver.hpp:343-568
  template <typename TIO> __device__ void run(const SpArgs<T, TIO>& a, long long b)
  {
    const T inf = Num<T>::inf();
    const sfb_qp_params& prm = a.prm;
    const unsigned long long t0 = prm.has_max_time ? global_timer_ns() : 0ull;
    // ---- ingest (array-of-instances -> tile; each lane streams its own rows, lines are reused through L1)
    {
      const TIO* gA = a.A + b * (long long)S.nnzA;
      const TIO* gP = a.P + b * (long long)S.nnzP;
      for (int e = r; e < S.nnzA; e += RL) A[e] = (T)__ldg(gA + e);
      for (int e = r; e < S.nnzP; e += RL) P[e] = (T)__ldg(gP + e);
      for (int j = r; j < n; j += RL) q[S.iperm[j]] = (T)__ldg(a.q + b * (long long)n + j);
      for (int i = r; i < m; i += RL) {
        l[i] = (T)__ldg(a.l + b * (long long)m + i);
        u[i] = (T)__ldg(a.u + b * (long long)m + i);
      }
      gsync();
    }
    if (prm.scaling) scale();  // :347
    else {
      c = T(1);
      for (int j = r; j < n; j += RL) sx[j] = T(1);
      for (int i = r; i < m; i += RL) sy[i] = T(1);
      gsync();
    }
    // ---- rho classes + trivially empty feasible set  :361-374
    int code = kStatusUnset;
    const T rho_bar = T(prm.rho), sigma = T(prm.sigma), alpha = T(prm.alpha), alpha_comp = T(1) - alpha;
    bool triv = false;
    for (int i = r; i < m; i += RL) {
      const T li = l[i], ui = u[i];
      if (li == inf || ui == -inf || ui - li < T(0)) triv = true;
      T rr;
      if (li == -inf && ui == inf) rr = T(1e-6);
      else if (sy[i] * fabs(li - ui) < T(1e-5)) rr = T(1e3) * rho_bar;
      else rr = rho_bar;
      rho[i] = rr;
      rinv[i] = T(1) / rr;
    }
    if (gany(triv)) code = SFB_QP_PRIMAL_INFEASIBLE;
    // ---- scaled data: qb = c Sx q, Abar = Sy A Sx  (:401-403, :450)
    for (int j = r; j < n; j += RL) qb[j] = (c * sx[j]) * q[j];
    {
      const V Al = A, sxl = sx;
      for (int i = r; i < m; i += RL) {
        const T syi = sy[i];
        const int e1 = S.A_rowptr[i + 1];
        for (int e = S.A_rowptr[i]; e < e1; e += kSpU) {
          T a0[kSpU], s0[kSpU];
#pragma unroll
          for (int k = 0; k < kSpU; ++k) a0[k] = (e + k < e1) ? Al[e + k] : T(0);
#pragma unroll
          for (int k = 0; k < kSpU; ++k) s0[k] = (e + k < e1) ? sxl[S.A_col[e + k]] : T(0);
#pragma unroll
          for (int k = 0; k < kSpU; ++k)
            if (e + k < e1) Al[e + k] = (syi * s0[k]) * a0[k];
        }
      }
    }
    for (int i = r; i < m; i += RL) { lo[i] = sy[i] * l[i]; hi[i] = sy[i] * u[i]; }
    gsync();
    if (padded_passes()) {  // padded row / column stream copies of Abar
      const V Al = A;
      auto copy_a = [&](const V& dst, const int* __restrict__ slot, int len) {
        for (int e = r; e < len; e += kSpU * RL) {
          int sl[kSpU];
          T a0[kSpU];
#pragma unroll
          for (int k = 0; k < kSpU; ++k) sl[k] = (e + k * RL < len) ? slot[e + k * RL] : -1;
#pragma unroll
          for (int k = 0; k < kSpU; ++k) a0[k] = (sl[k] >= 0) ? Al[sl[k]] : T(0);
#pragma unroll
          for (int k = 0; k < kSpU; ++k)
            if (e + k * RL < len) dst[e + k * RL] = a0[k];
        }
      };
      copy_a(APW, S.RP_slot, S.m_pad * S.WR);
      copy_a(ATW, S.ATP_slot, S.n_pad * S.WA);
      gsync();
    }
    // Mixed-precision polish (mode 2, T = double over TIO = float data): the instance was solved by the single-precision
    // kernel; this pass re-stages it in double, takes the unpolished iterate and the active set from the outputs and runs
    // only polish_qp (which assembles and factorises its own system).  Every lane of an instance reads the same status.
    const bool polish_only = a.mode == 2;
    const bool skip = polish_only && a.out_status[b] != (int32_t)SFB_QP_OPTIMAL;
    if (!polish_only) {
      assemble(sigma, rho);
      if (!factor()) code = SFB_QP_UNKNOWN;  // :433
    }
    // ---- initial iterate  :436-445
    if (polish_only) {
      for (int j = r; j < n; j += RL) {
        const int pj = S.iperm[j];
        x[pj] = (T(1) / sx[pj]) * (T)a.out_x[b * (long long)n + j];
      }
      for (int i = r; i < m; i += RL) {
        y[i] = c * ((T(1) / sy[i]) * (T)a.out_y[b * (long long)m + i]);
        z[i] = T(0);
      }
    } else if (a.warm_x != nullptr) {
      for (int j = r; j < n; j += RL) {
        const int pj = S.iperm[j];
        x[pj] = (T(1) / sx[pj]) * (T)__ldg(a.warm_x + b * (long long)n + j);
      }
      gsync();
      for (int i = r; i < m; i += RL) {
        y[i] = c * ((T(1) / sy[i]) * (T)__ldg(a.warm_y + b * (long long)m + i));
        z[i] = A_row_dot(i, x);
      }
    } else {
      for (int j = r; j < n; j += RL) x[j] = T(0);
      for (int i = r; i < m; i += RL) { y[i] = T(0); z[i] = T(0); }
    }
    for (int i = r; i < m; i += RL) w[i] = rho[i] * z[i] - y[i];
    gsync();

    // ---- main loop  :449-510
    const unsigned sci = prm.stop_check_iter;
    unsigned iter = 0;
    if (polish_only) { code = skip ? kStatusUnset - 1 : (int)SFB_QP_OPTIMAL; iter = a.out_iter[b]; }
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      if (padded_passes()) rhs_pass(sigma);
      else At_gather(w, v, [&](int j, T sum) { return sigma * x[j] - qb[j] + sum; });  // rhs = sigma x - qb + Abar^T w
      solve();
      const bool chk = (iter % sci == 1u);
      for (int j = r; j < n; j += RL) {
        const T xi = x[j];
        if (chk) xold[j] = xi;  // :465-468
        x[j] = alpha * v[j] + alpha_comp * xi;  // :470
      }
      if (padded_passes()) {
        row_pass(alpha, alpha_comp, chk);
      } else {
        for (int i = r; i < m; i += RL) {
          const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];  // issued before the dot: their latency overlaps it
          const T lob = lo[i], hib = hi[i];
          const T zt = A_row_dot(i, v);
          if (chk) yold[i] = yi;
          const T nu = ri * (zt - zi) + yi;
          T zn = alpha * (rinvi * nu) + alpha_comp * (rinvi * yi) + zi;  // :471-474
          zn = fmax(zn, lob);
          zn = fmin(zn, hib);
          const T yn = alpha_comp * yi + alpha * nu + ri * zi - ri * zn;  // :475-477
          y[i] = yn;
          z[i] = zn;
          w[i] = ri * zn - yn;
        }
        gsync();
      }
      if (chk) {
        code = check_stopping(prm, a.dinf_guard != 0);  // :488 (clobbers w, v, t1..t4, xold)
        if (code == kStatusUnset && prm.has_max_time) {
          const bool late = (long long)(global_timer_ns() - t0) > prm.max_time_ns;  // :504-508
          if (gany(late)) code = SFB_QP_MAX_TIME;
        }
        for (int i = r; i < m; i += RL) w[i] = rho[i] * z[i] - y[i];
        gsync();
      }
    }

    // ---- active sets as polish_qp builds them (:113-123) on the scaled dual
    const T thr = T(100) * Num<T>::eps();
    for (int i = r; i < m; i += RL) {
      int act = 0;
      T bv = T(0);
      if (polish_only) {  // the active set the lower-precision solve determined
        act = a.out_active[b * (long long)m + i];
        if (act < 0) bv = sy[i] * l[i];
        if (act > 0) bv = sy[i] * u[i];
      } else {
        if (y[i] < -thr && l[i] != -inf) { act = -1; bv = sy[i] * l[i]; }
        if (y[i] > thr && u[i] != inf) { act = 1; bv = sy[i] * u[i]; }
        if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
      }
      w[i] = (act != 0) ? T(1) : T(0);
      yold[i] = bv;
    }
    gsync();
    unsigned flags = 0;
    if (code == SFB_QP_OPTIMAL && prm.polish) {
      if (sizeof(T) == 4) flags = SFB_QP_FLAG_POLISH_SKIPPED;  // delta = 1e-6 is not resolvable in fp32 (as in the dense kernel)
      else flags = polish(prm);
    }
    if (skip) return;  // polish-only pass over an instance that is not Optimal: outputs stay as they are (instance-uniform)
    // ---- unscale + objective  :544-548
    for (int j = r; j < n; j += RL) t1[j] = sx[j] * x[j];
    gsync();
    P_gather(t1, t3, T(0.5));
    T obj = T(0);
    for (int jo = r; jo < n; jo += RL) {
      const int j = S.iperm[jo];
      const T xv = t1[j];
      a.out_x[b * (long long)n + jo] = (TIO)xv;
      obj += xv * (t3[j] + q[j]);
    }
    obj = gsum(obj);
    for (int i = r; i < m; i += RL) a.out_y[b * (long long)m + i] = (TIO)(sy[i] * y[i] / c);
    if (r == 0) {
      a.out_obj[b] = (TIO)obj;
      a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
      a.out_iter[b] = iter;
      if (a.out_flags) a.out_flags[b] = flags;
    }
  }
};

// One warp per tile of TW instances, one warp per CTA (small batches still spread over all SMs).
// register budget: all instances of a batch should be resident at once with slack (TW = 4 at batch 8192 needs 14 warps
// per SM; 144 registers fit exactly 14 and measured 1.6x slower because the last blocks wait for a second wave)
template <typename T, int TW, typename TIO = T> __global__ void __launch_bounds__(32, TW == 8 ? 8 : 16) qp_sparse_tiled_kernel(const SpArgs<T, TIO> a)
{
  extern __shared__ __align__(16) unsigned char sp_smem_raw[];
  const int lane = threadIdx.x;
  T* smem_v = (TW < 32) ? reinterpret_cast<T*>(sp_smem_raw) : nullptr;
  if (TW < 32) {  // dummy slot v[n] == 0: the target of every padding gather
    if (lane < TW) smem_v[(size_t)a.pat.n * TW + lane] = T(0);
    __syncwarp();
  }
  const long long ntiles = (a.batch + TW - 1) / TW;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b = tile * TW + (lane & (TW - 1));
    if (b < a.batch) {
      SpSolver<T, TW> s(a, tile, lane, smem_v);
      s.run(a, b);
    }
    __syncwarp();
  }
}

}  // namespace sfb
