// qp_sparse_cta.cuh -- ON-CHIP batched sparse operator-splitting QP solver for sm_100a: one CTA per instance, the
// supernodal L D L^T factor, Abar, P, the vectors of the iteration and every index table of the hot loops resident in
// shared memory.
//
// Replaces the sparse branches of (pettni/smooth_feedback @ 9a08971)
//   QPSolver<QuadraticProgramSparse>::solve   include/smooth/feedback/qp_solver.hpp:343-568 (fill :380-397,
//                                             SimplicialLDLT factorize :424-426, permuted solve :456-460)
//   QPSolver::scale / check_stopping          :673-730 / :574-644
//   detail::polish_qp                         :92-204
// i.e. the call MPC::operator() makes at mpc.hpp:491 -- same algorithm and arithmetic as qp_sparse_tiled.cuh (which
// stays as the fallback for patterns whose working set does not fit in shared memory), different machine mapping:
//   * HBM traffic is the compulsory bytes only (inputs once, outputs once); the tiled kernel re-streamed Abar and the
//     factor from HBM in every iteration (27 MB per solve at n = m = 422);
//   * the triangular sweeps are 4 x (supernode levels) barrier-separated stages of independent dot products over dense
//     supernodal blocks with INVERTED diagonal blocks (qp_sparse_cta_host.hpp) instead of ~850 dependent steps;
//   * the numeric factorisation advances all supernodes of one level column by column together (each with its own
//     warps), then pushes their rank-s updates into the levels above;
//   * assembly of Abar^T R Abar runs colour by colour (rows of one colour touch disjoint columns): deterministic, no atomics;
//   * nothing on a dependent chain is fetched from global memory: the first version kept the schedules in global memory
//     and spent 2.7 k cycles per sweep stage on chains of L2 round trips (profiles/r02_cta_phases.txt).
// The reduced system (Pbar + sigma I + Abar^T R Abar) xt = sigma x - qbar + Abar^T (R z - y) is the tiled / dense kernels'.
// CTAs are persistent and pull instances from a work counter (iteration counts differ between instances).

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sfb.h"
#include "qp_dense_group.cuh"  // Num<T>, kStatusUnset, global_timer_ns
#include "qp_sparse_cta_host.hpp"

namespace sfb {

struct CtaPattern
{
  int n, m, np, nnzP, nnzA, nW, ns, smax, nstages, nrounds, nfacrounds, nints;
  unsigned tscalars;  // scalars of type T in front of the integer tables (cta_smem_scalars)
  int ioff[kI_count];
  const int *ints;    // the tables copied into shared memory (CtaSymbolic::smem_ints)
  const int *perm, *iperm, *P_rowp, *P_colp, *P_tgt, *PR_ptr, *PR_col, *PR_slot, *PS_ptr, *PS_col, *PS_slot, *PC_ptr, *PC_slot, *asm_sync;
  const int2* asm_desc;
  const int *ATword_g, *extword_g;  // the two big tables in global memory (kernels instantiated with GT = true read them there)
};

template <typename T, typename TIO = T> struct CtaArgs
{
  CtaPattern pat;
  const TIO *P, *q, *A, *l, *u, *warm_x, *warm_y;
  TIO *out_x, *out_y, *out_obj;
  int32_t* out_status;
  uint32_t* out_iter;
  int8_t* out_active;
  uint32_t* out_flags;
  T* ws;  // per CTA: np + m scalars (x and y of the previous stop check)
  TIO* scale_ws;  // [batch][n + m + 1] or nullptr: the scaling (Sx in the caller's column order, Sy, c) an fp32 solve hands to its fp64 polish pass
  unsigned long long* work_counter;
  long long batch;
  sfb_qp_params prm;
  unsigned max_iter_eff;
  int dinf_guard;
  int mode;  // 0 = solve, 2 = polish only (see SpArgs::mode)
  unsigned long long* prof;  // dev instrumentation (SFB_CTA_PROF=1): cycles per phase summed over CTAs, nullptr otherwise
};

enum CtaPhase { kPhLoad = 0, kPhScale, kPhPrep, kPhAssemble, kPhFactorCols, kPhFactorExt, kPhFactorInv, kPhRhs, kPhSolve, kPhUpdate,
                kPhCheck, kPhPolish, kPhOut, kPhCount, kPhStage0 = kPhCount, kPhTotal = kPhCount + 32 };

// a vector in shared memory, addressed through the extern symbol (LDS / STS with 32-bit addresses)
template <typename T> struct SV
{
  unsigned off;
  __device__ __forceinline__ T& operator[](int i) const
  {
    extern __shared__ __align__(16) unsigned char cta_smem_raw[];
    return reinterpret_cast<T*>(cta_smem_raw)[off + (unsigned)i];
  }
};

__device__ __forceinline__ int lo16(int x) { return (int)((unsigned)x & 0xffffu); }
__device__ __forceinline__ int hi16(int x) { return (int)((unsigned)x >> 16); }

// 16-byte vectors of the compute type: block rows of the factor are padded to whole vectors (qp_sparse_cta_host.hpp)
template <typename T> struct VecOf;
template <> struct VecOf<float>
{
  using type = float4;
  static constexpr int PAD = 4, LPAD = 2;
  static __device__ __forceinline__ void unpack(const float4& v, float (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }
  static __device__ __forceinline__ float4 pack(const float (&o)[4]) { return make_float4(o[0], o[1], o[2], o[3]); }
};
template <> struct VecOf<double>
{
  using type = double2;
  static constexpr int PAD = 2, LPAD = 1;
  static __device__ __forceinline__ void unpack(const double2& v, double (&o)[2]) { o[0] = v.x; o[1] = v.y; }
  static __device__ __forceinline__ double2 pack(const double (&o)[2]) { return make_double2(o[0], o[1]); }
};

template <typename T, typename TIO, int NT, bool GT> struct CtaSolver
{
  static constexpr int PAD = VecOf<T>::PAD, LPAD = VecOf<T>::LPAD;
  static constexpr int BW = 4, LBW = 2;  // pivot columns per step of the blocked factorisation (one fp32 vector, two fp64 vectors)
  using VT = typename VecOf<T>::type;
  static constexpr int NW = NT / 32;
  using V = SV<T>;
  const CtaArgs<T, TIO>& a;
  const CtaPattern& S;
  const int n, m, np, nW, tid, lane, warp;
  V W, A;                             // factor slots (+ 1/D), Abar (CSR order)
  V qb, x, v, sv, sx, t1;             // n-vectors over the padded ids (holes stay zero); v / sv: the two buffers of the sweeps, sv = result
  V sy, rho, rinv, z, y, w, lo, hi;   // m-vectors
  V red, piv;                         // reduction scratch; parked pivot-block factorisations (2 x NW slots)
  T *xold, *yold;                     // global: iterate at the previous stop check
  unsigned ibase;                     // byte offset of the integer tables
  T c;
  T qn_us;                            // max |q| (unscaled)
  long long b = 0;                    // current instance
  unsigned long long* prof = nullptr;
  long long tlast = 0;

  __device__ CtaSolver(const CtaArgs<T, TIO>& args)
      : a(args), S(args.pat), n(args.pat.n), m(args.pat.m), np(args.pat.np), nW(args.pat.nW), tid(threadIdx.x), lane(threadIdx.x & 31),
        warp(threadIdx.x >> 5)
  {
    unsigned o = 0;
    auto r4 = [](unsigned k) { return (k + 3u) & ~3u; };
    auto take = [&](unsigned len) { V r{o}; o += r4(len); return r; };  // same carving as cta_smem_scalars
    W = take(nW + np); A = take(S.nnzA);
    qb = take(np); x = take(np); v = take(np); sv = take(np); sx = take(np); t1 = take(np);
    sy = take(m); rho = take(m); rinv = take(m); z = take(m); y = take(m); w = take(m); lo = take(m); hi = take(m);
    red = take(kCtaRed);
    piv = take(2 * NW * kCtaPivot);
    ibase = S.tscalars * (unsigned)sizeof(T);
    xold = a.ws + (size_t)blockIdx.x * (size_t)(np + m);
    yold = xold + np;
    c = T(1);
    qn_us = T(0);
  }

  __device__ __forceinline__ const int* itab(int which) const
  {
    extern __shared__ __align__(16) unsigned char cta_smem_raw[];
    return reinterpret_cast<const int*>(cta_smem_raw + ibase) + S.ioff[which];
  }
  __device__ __forceinline__ const unsigned short* htab(int which) const { return reinterpret_cast<const unsigned short*>(itab(which)); }
  // once per CTA: schedules into shared memory, every scalar zero (holes of the padded vectors and 1 / D of the holes stay zero)
  __device__ void load_tables()
  {
    extern __shared__ __align__(16) unsigned char cta_smem_raw[];
    int* dst = reinterpret_cast<int*>(cta_smem_raw + ibase);
    for (int k = tid; k < S.nints; k += NT) dst[k] = __ldg(S.ints + k);
    T* sc = reinterpret_cast<T*>(cta_smem_raw);
    for (unsigned k = tid; k < S.tscalars; k += NT) sc[k] = T(0);
    __syncthreads();
  }
  // 16-byte accesses (slot a multiple of PAD)
  __device__ __forceinline__ void ldv(const V& vec, int slot, T (&o)[PAD]) const { VecOf<T>::unpack(*reinterpret_cast<const VT*>(&vec[slot]), o); }
  __device__ __forceinline__ void stv(const V& vec, int slot, const T (&o)[PAD]) const { *reinterpret_cast<VT*>(&vec[slot]) = VecOf<T>::pack(o); }
  // inputs of the current instance, read in place (global, read-only); pj: padded id
  __device__ __forceinline__ T qg(int pj) const
  {
    const int jo = __ldg(S.perm + pj);
    return jo >= 0 ? (T)__ldg(a.q + b * (long long)n + jo) : T(0);
  }
  __device__ __forceinline__ T Pg(int e) const { return (T)__ldg(a.P + b * (long long)S.nnzP + e); }  // P as given (only cold paths read it)
  __device__ __forceinline__ T lg(int i) const { return (T)__ldg(a.l + b * (long long)m + i); }
  __device__ __forceinline__ T ug(int i) const { return (T)__ldg(a.u + b * (long long)m + i); }

  __device__ __forceinline__ void mark(int phase)
  {
    if (prof != nullptr && tid == 0) {
      const long long t = clock64();
      atomicAdd(prof + phase, (unsigned long long)(t - tlast));
      tlast = t;
    }
  }

  // ---------------------------------------------------------------- CTA reductions (results on every thread, fixed order)
  // K values at once; bit k of MAXMASK: maximum instead of sum
  template <int K, unsigned MAXMASK> __device__ __forceinline__ void reduce(T (&val)[K])
  {
    static_assert(K * NW <= kCtaRed, "reduction scratch too small");
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const T other = __shfl_xor_sync(0xffffffffu, val[k], o);
        val[k] = ((MAXMASK >> k) & 1u) ? fmax(val[k], other) : val[k] + other;
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < K; ++k) red[warp * K + k] = val[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
      T acc = red[k];
#pragma unroll
      for (int wv = 1; wv < NW; ++wv) acc = ((MAXMASK >> k) & 1u) ? fmax(acc, red[wv * K + k]) : acc + red[wv * K + k];
      val[k] = acc;
    }
    __syncthreads();
  }
  __device__ __forceinline__ T bmax(T val) { T r[1] = {val}; reduce<1, 1u>(r); return r[0]; }
  __device__ __forceinline__ T bsum(T val) { T r[1] = {val}; reduce<1, 0u>(r); return r[0]; }
  __device__ __forceinline__ bool bany(bool pr) { return __syncthreads_or(pr ? 1 : 0) != 0; }

  // ---------------------------------------------------------------- gathers over A (index tables in shared memory)
  // sum_i Abar_ij in_i for column j (padded id); one word per entry: slot | row << 16
  __device__ __forceinline__ T At_col_dot(int j, const V& in) const
  {
    const unsigned short* atp = htab(kI_ATptr);
    const int* atw = GT ? S.ATword_g : itab(kI_ATword);
    const int e0 = atp[j], e1 = atp[j + 1];
    T acc = T(0);
    for (int e = e0; e < e1; e += 4) {
      int wd[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) wd[k] = (e + k < e1) ? atw[e + k] : -1;
      T av[4], iv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        av[k] = (wd[k] != -1) ? A[lo16(wd[k])] : T(0);
        iv[k] = (wd[k] != -1) ? in[hi16(wd[k])] : T(0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += av[k] * iv[k];
    }
    return acc;
  }
  // two column dots at once (check_stopping)
  __device__ __forceinline__ void At_col_dot2(int j, const V& in1, const V& in2, T& o1, T& o2) const
  {
    const unsigned short* atp = htab(kI_ATptr);
    const int* atw = GT ? S.ATword_g : itab(kI_ATword);
    const int e1 = atp[j + 1];
    T a1 = T(0), a2 = T(0);
#pragma unroll 2
    for (int e = atp[j]; e < e1; ++e) {
      const int wd = atw[e];
      const T av = A[lo16(wd)];
      a1 += av * in1[hi16(wd)];
      a2 += av * in2[hi16(wd)];
    }
    o1 = a1; o2 = a2;
  }
  __device__ __forceinline__ T A_row_dot(int i, const V& vec) const
  {
    const unsigned short *arp = htab(kI_Arowptr), *acol = htab(kI_Acol);
    const int e0 = arp[i], e1 = arp[i + 1];
    T acc = T(0);
    for (int e = e0; e < e1; e += 4) {
      int cj[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) cj[k] = (e + k < e1) ? (int)acol[e + k] : -1;
      T av[4], xv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        av[k] = (cj[k] >= 0) ? A[e + k] : T(0);
        xv[k] = (cj[k] >= 0) ? vec[cj[k]] : T(0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) acc += av[k] * xv[k];
    }
    return acc;
  }
  // sum_k sym(Pbar)_jk in_k  (upper triangle of c Sx P Sx mirrored)
  __device__ __forceinline__ T Psym_row_dot(int j, const V& in) const
  {
    T acc = T(0);
    for (int e = S.PS_ptr[j]; e < S.PS_ptr[j + 1]; ++e) {
      const int sl = __ldg(S.PS_slot + e);
      acc += (((c * sx[__ldg(S.P_rowp + sl)]) * sx[__ldg(S.P_colp + sl)]) * Pg(sl)) * in[__ldg(S.PS_col + e)];
    }
    return acc;
  }

  // ---------------------------------------------------------------- QPSolver::scale, qp_solver.hpp:673-730
  // Same evaluation as qp_sparse_tiled.cuh::scale (bit-exact against the CPU restatement: max / abs / mul / div / sqrt only,
  // products in the reference's order, the column mean summed sequentially in the original column order).
  __device__ void scale()
  {
    const unsigned short *arp = htab(kI_Arowptr), *acol = htab(kI_Acol), *atp = htab(kI_ATptr);
    const int* atw = GT ? S.ATword_g : itab(kI_ATword);
    for (int j = tid; j < np; j += NT) sx[j] = T(1);
    for (int i = tid; i < m; i += NT) sy[i] = T(1);
    for (int pj = tid; pj < np; pj += NT) {
      T g = T(0);
      for (int e = S.PC_ptr[pj]; e < S.PC_ptr[pj + 1]; ++e) g = fmax(g, fabs(Pg(__ldg(S.PC_slot + e))));
      t1[pj] = g;
    }
    __syncthreads();
    if (tid == 0) {
      T mean = T(0);
      for (int j = 0; j < n; ++j) {
        T g = t1[__ldg(S.iperm + j)];
        if (g == T(0)) g = T(1);
        mean += g;
      }
      mean /= T(n);
      red[0] = T(1) / fmax(fmax(T(1e-6), mean), qn_us);
    }
    __syncthreads();
    c = red[0];
    __syncthreads();
    const T cc = c;
    int it = 0;
    T dev;
    do {
      for (int pj = tid; pj < np; pj += NT) {
        const T sxj = sx[pj];
        T g = T(0);
        for (int e = S.PC_ptr[pj]; e < S.PC_ptr[pj + 1]; ++e) {
          const int sl = __ldg(S.PC_slot + e);
          g = fmax(g, fabs(((cc * sx[__ldg(S.P_rowp + sl)]) * sxj) * Pg(sl)));
        }
        const int e1 = atp[pj + 1];
#pragma unroll 4
        for (int e = atp[pj]; e < e1; ++e) {
          const int wd = atw[e];
          g = fmax(g, fabs((sy[hi16(wd)] * sxj) * A[lo16(wd)]));
        }
        t1[pj] = g;
      }
      for (int i = tid; i < m; i += NT) {
        const T syi = sy[i];
        T g = T(0);
        const int e1 = arp[i + 1];
#pragma unroll 4
        for (int e = arp[i]; e < e1; ++e) g = fmax(g, fabs((syi * sx[acol[e]]) * A[e]));
        w[i] = g;
      }
      __syncthreads();  // every maximum is formed before any scale factor changes
      dev = T(0);
      for (int j = tid; j < np; j += NT) {
        T g = t1[j];
        if (g == T(0)) g = T(1);
        sx[j] = sqrt(T(1) / fmax(g, T(1e-8))) * sx[j];
        dev = fmax(dev, fabs(g - T(1)));
      }
      for (int i = tid; i < m; i += NT) {
        T g = w[i];
        if (g == T(0)) g = T(1);
        sy[i] = sqrt(T(1) / fmax(g, T(1e-8))) * sy[i];
        dev = fmax(dev, fabs(g - T(1)));
      }
      dev = bmax(dev);
    } while (it++ < 10 && dev > T(0.1));
  }

  // ---------------------------------------------------------------- W <- shift I + c Sx triu(P) Sx + Abar^T diag(wt) Abar
  __device__ void assemble(T shift, const V& wt)
  {
    for (int e = tid; e < nW; e += NT) W[e] = T(0);
    __syncthreads();
    {  // the diagonal of every real column lives in its row of the diagonal block
      const int *fd = itab(kI_fdiag), *sn = itab(kI_sntab);
      const unsigned short* prow = htab(kI_prow);
      for (int e = tid; e < n; e += NT) {
        const int wd = fd[e], q = hi16(wd), i = lo16(wd) - sn[8 * q];
        W[prow[sn[8 * q + 7] + i] + i] = shift;
      }
    }
    __syncthreads();
    for (int e = tid; e < S.nnzP; e += NT) {  // compressed P: distinct targets
      const int t = __ldg(S.P_tgt + e);
      if (t >= 0) W[t] += ((c * sx[__ldg(S.P_rowp + e)]) * sx[__ldg(S.P_colp + e)]) * Pg(e);  // qp_solver.hpp:386
    }
    __syncthreads();
    // colour by colour; the descriptors of four rounds are requested together (one L2 round trip per four rounds)
    constexpr int U = 4;
    for (int r0 = 0; r0 < S.nrounds; r0 += U) {
      int2 d[U];
      int fl[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        const bool in = r0 + k < S.nrounds;
        d[k] = in ? __ldg(S.asm_desc + (size_t)(r0 + k) * NT + tid) : make_int2(0, -1);
        fl[k] = in ? __ldg(S.asm_sync + r0 + k) : 0;
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        if (d[k].y != -1) {
          const T wr = wt[hi16(d[k].y)];
          if (wr != T(0)) W[lo16(d[k].y)] += (wr * A[lo16(d[k].x)]) * A[hi16(d[k].x)];
        }
        if (fl[k]) __syncthreads();
      }
    }
  }

  // ---------------------------------------------------------------- pieces of the blocked factorisation
  // The first nv (1 .. BW) of BW consecutive scalars from slot (a multiple of PAD), zeros behind them: only the vectors that hold
  // one of the nv columns are read -- a below row ends with its supernode's last (padded) column
  __device__ __forceinline__ void ldb(int slot, int nv, T (&o)[BW]) const
  {
#pragma unroll
    for (int h = 0; h < BW / PAD; ++h) {
      T part[PAD];
      const bool in = h * PAD < nv;
      if (in) ldv(W, slot + h * PAD, part);
#pragma unroll
      for (int cc = 0; cc < PAD; ++cc) o[h * PAD + cc] = (in && h * PAD + cc < nv) ? part[cc] : T(0);
    }
  }
  // y = u L11^-T for the unit-lower pivot block L11 (multipliers lm[c][c'], c' < c): y[c] = u[c] - sum_{c' < c} y[c'] lm[c][c']
  __device__ __forceinline__ void fwd_sub(const T (&u)[BW], const T (&lm)[BW][BW], T (&y)[BW]) const
  {
#pragma unroll
    for (int cc = 0; cc < BW; ++cc) {
      T acc = u[cc];
#pragma unroll
      for (int c2 = 0; c2 < cc; ++c2) acc -= y[c2] * lm[cc][c2];
      y[cc] = acc;
    }
  }
  // L D L^T of the BW x BW pivot block at (kb, kb) of a supernode's diagonal block, in registers: yd[i][c] (c <= i) the
  // forward-substituted entries (yd[i][i] = D_i), lm the unit-lower multipliers, dinv = 1 / D (0 for columns >= s).  False on a
  // non-positive pivot.
  __device__ __forceinline__ bool pivot_block(const unsigned short* pr, int kb, int s, T (&lm)[BW][BW], T (&dinv)[BW], T (&yd)[BW][BW]) const
  {
    const T inf = Num<T>::inf();
    T av[BW][BW];
#pragma unroll
    for (int i = 0; i < BW; ++i) {
      const bool in = kb + i < s;
      if (in) ldb(pr[kb + i] + kb, i + 1, av[i]);  // row i of the block holds its columns 0 .. i only (a longer read would touch the next row)
#pragma unroll
      for (int cc = 0; cc < BW; ++cc) av[i][cc] = in ? av[i][cc] : T(0);
    }
    bool good = true;
#pragma unroll
    for (int cc = 0; cc < BW; ++cc) {
#pragma unroll
      for (int i = cc; i < BW; ++i) {
        T acc = av[i][cc];
#pragma unroll
        for (int c2 = 0; c2 < cc; ++c2) acc -= yd[i][c2] * lm[cc][c2];
        yd[i][cc] = acc;
      }
      const bool valid = kb + cc < s;
      const T d = yd[cc][cc];
      if (valid && (!(d > T(0)) || !(d < inf))) good = false;
      dinv[cc] = valid ? T(1) / d : T(0);
#pragma unroll
      for (int i = cc + 1; i < BW; ++i) lm[i][cc] = yd[i][cc] * dinv[cc];
    }
    return good;
  }

  // ---------------------------------------------------------------- supernodal right-looking L D L^T in place
  // Afterwards: below blocks hold L, diagonal blocks hold X' = -(strict lower part of L_SS^-1) (diagonal and padding zero),
  // W[nW + k] = 1 / D_k.  False on a non-positive pivot.
  // Rounds (qp_sparse_cta_host.hpp): the supernodes of a round advance together, each with its own warps, PAD columns per step.
  // The diagonal D_i lives IN row i of its block during the factorisation, so a step is one uniform update of the trailing
  // panel: a thread owns a group of PAD adjacent columns and walks panel rows with 16-byte read-modify-writes; the padding of
  // a row (columns > i) collects garbage that the scaling pass clears.
  __device__ bool factor()
  {
    const T inf = Num<T>::inf();
    bool ok = true;
    const int *rounds = itab(kI_facrounds), *fext = itab(kI_facext), *fwarp = itab(kI_facwarp), *sn = itab(kI_sntab), *extptr = itab(kI_extptr),
              *extw = GT ? S.extword_g : itab(kI_extword);
    const unsigned short* prow = htab(kI_prow);
    const int Di = nW;  // 1 / D
    for (int r = 0; r < S.nfacrounds; ++r) {
      const int maxs = rounds[4 * r], cnt = lo16(rounds[4 * r + 1]), totw = hi16(rounds[4 * r + 1]), e0 = rounds[4 * r + 2], w0 = rounds[4 * r + 3];
      int q = -1, rank = 0, nw = 1;
      if (warp < totw) {
        const unsigned wd = (unsigned)fwarp[w0 + warp];
        q = (int)(wd & 0xffffu); rank = (int)((wd >> 16) & 0xffu); nw = (int)(wd >> 24);
      }
      int c0 = 0, s = 0, t = 0, sp = 0;
      const unsigned short* pr = prow;
      if (q >= 0) {
        const int4 h = *reinterpret_cast<const int4*>(sn + 8 * q);
        c0 = h.x; s = h.y; t = h.z; sp = sn[8 * q + 6]; pr = prow + sn[8 * q + 7];
      }
      // thread -> (column group cg0 of NCG = sp / PAD, row chunk rc of RC); up to 32 column groups side by side (NCG rounded up to a
      // power of two), a wider supernode is walked in strides of 32 groups
      const int gsize = nw * 32, gtid = rank * 32 + lane;
      const int NCG = sp >> LPAD;
      int lcg = 0;
      while ((1 << lcg) < NCG && lcg < 5) ++lcg;
      const int CGS = 1 << lcg, cg0 = gtid & (CGS - 1), rc = gtid >> lcg, RC = gsize >> lcg;
      const bool active = q >= 0 && rc < RC;
      const int nrows = s + t;
      // Blocked right-looking steps, BW = 4 pivot columns at a time.  In step kb a thread of a column group right of the block
      // (A) factorises the BW x BW pivot block in registers (redundantly: a dozen loads and a few dozen flops instead of a
      // barrier), forward-substitutes the pivot-column entries of its PAD group rows and of every panel row it owns
      // (y = u L11^-T) and applies the rank-BW update to its vectors.  The pivot columns themselves are read by everybody during
      // the step, so their final values y (and D on the diagonal, 1 / D beside it) are written back one step LATER, by the
      // threads of the block's first column group (B), from the pivot factorisation that one otherwise idle thread of that group
      // (P) parked in shared memory during the step -- nobody re-reads rows that are being overwritten.
      const int nsteps = ((maxs + BW - 1) >> LBW) + 1;
      const int pslot = (warp - rank) * kCtaPivot;  // this supernode's slot in the pivot scratch (double-buffered by step parity)
      for (int b = 0; b < nsteps; ++b) {
        const int kb = b << LBW;
        __syncthreads();
        if (!active) continue;
        const int kp = kb - BW;  // the previous block
        for (int cg = cg0; cg < NCG; cg += CGS) {
        const int j0 = cg << LPAD;
        if (j0 == kb && rc == 0 && kb < s) {  // (P) park this block's factorisation: yd (lower triangle, row-major), then 1 / D
          T lm[BW][BW], dinv[BW], yd[BW][BW];
          if (!pivot_block(pr, kb, s, lm, dinv, yd)) ok = false;
          const int sl = (b & 1) * NW * kCtaPivot + pslot;
#pragma unroll
          for (int i = 0; i < BW; ++i) {
#pragma unroll
            for (int cc = 0; cc <= i; ++cc) piv[sl + i * (i + 1) / 2 + cc] = yd[i][cc];
            piv[sl + BW * (BW + 1) / 2 + i] = dinv[i];
          }
        }
        if (kp >= 0 && j0 == kp && kp < s) {  // (B) write the previous block back: all its BW columns of the rows this thread owns
          T lm[BW][BW], dinv[BW], yd[BW][BW];
          const int sl = ((b - 1) & 1) * NW * kCtaPivot + pslot;
#pragma unroll
          for (int i = 0; i < BW; ++i) {
#pragma unroll
            for (int cc = 0; cc <= i; ++cc) yd[i][cc] = piv[sl + i * (i + 1) / 2 + cc];
            dinv[i] = piv[sl + BW * (BW + 1) / 2 + i];
          }
#pragma unroll
          for (int i = 1; i < BW; ++i) {
#pragma unroll
            for (int cc = 0; cc < i; ++cc) lm[i][cc] = yd[i][cc] * dinv[cc];
          }
          const int kend = min(kb, s), nvp = min(BW, s - kp);
          for (int p = kp + rc; p < nrows; p += RC) {
            T y[BW];
            const int base = pr[p];
            if (p < kend) {  // a row of the pivot block: multipliers left of the diagonal, D on it, zeros right of it
              const int i = p - kp;
#pragma unroll
              for (int cc = 0; cc < BW; ++cc) {
                T val = T(0);
#pragma unroll
                for (int ii = 0; ii < BW; ++ii) val = (ii == i && cc <= ii) ? yd[ii][cc] : val;
                y[cc] = val;
              }
              T di = dinv[0];
#pragma unroll
              for (int ii = 1; ii < BW; ++ii) di = (ii == i) ? dinv[ii] : di;
              W[Di + c0 + p] = di;
            } else {
              T u[BW];
              ldb(base + kp, nvp, u);
              fwd_sub(u, lm, y);
            }
#pragma unroll
            for (int h = 0; h < BW / PAD; ++h) {
              if (h * PAD < nvp && (p >= s || kp + h * PAD <= p)) {  // the vector exists: inside the supernode's columns, and a row of the diagonal block reaches that far
                T out[PAD];
#pragma unroll
                for (int cc = 0; cc < PAD; ++cc) out[cc] = y[h * PAD + cc];
                stv(W, base + kp + h * PAD, out);
              }
            }
          }
        }
        if (kb < s && j0 >= kb + BW && j0 < s) {  // (A) rank-BW update of this thread's column group
          T lm[BW][BW], dinv[BW], yd[BW][BW];
          pivot_block(pr, kb, s, lm, dinv, yd);
          T ys[PAD][BW];  // ys[jj][c] = y(j0 + jj, c) / D_c for the rows of the diagonal block that carry this group's columns
#pragma unroll
          for (int jj = 0; jj < PAD; ++jj) {
            T u[BW], y[BW];
            const bool in = j0 + jj < s;
            if (in) ldb(pr[j0 + jj] + kb, BW, u);
#pragma unroll
            for (int cc = 0; cc < BW; ++cc) u[cc] = in ? u[cc] : T(0);
            fwd_sub(u, lm, y);
#pragma unroll
            for (int cc = 0; cc < BW; ++cc) ys[jj][cc] = y[cc] * dinv[cc];
          }
          for (int p = max(min(kb + BW, s), j0) + rc; p < nrows; p += RC) {  // rows below the pivot block that hold this group
            T u[BW], y[BW], old[PAD];
            const int base = pr[p];
            ldb(base + kb, BW, u); ldv(W, base + j0, old);
            fwd_sub(u, lm, y);
#pragma unroll
            for (int jj = 0; jj < PAD; ++jj) {
              T acc = old[jj];
#pragma unroll
              for (int cc = 0; cc < BW; ++cc) acc -= y[cc] * ys[jj][cc];
              old[jj] = acc;
            }
            stv(W, base + j0, old);
          }
        }
        }  // column groups of this thread
      }
      __syncthreads();
      mark(kPhFactorCols);
      // external updates, one supernode after the other (supernodes of one round may share targets): U D^-1 U^T of the below block
      for (int k = 0; k < cnt; ++k) {
        const int qe = fext[e0 + k];
        const int ec0 = sn[8 * qe], ebb = sn[8 * qe + 4], esp = sn[8 * qe + 6];
        const int p1 = extptr[qe + 1];
        for (int p = extptr[qe] + tid; p < p1; p += NT) {
          const unsigned wd = (unsigned)extw[p];
          const int ra = ebb + (int)(wd & 0xffu) * esp, rb = ebb + (int)((wd >> 8) & 0xffu) * esp, tg = (int)(wd >> 16);
          T acc[PAD];
#pragma unroll
          for (int cc = 0; cc < PAD; ++cc) acc[cc] = T(0);
          for (int ch = 0; ch < esp; ch += PAD) {  // whole padded rows: the padding columns hold zeros (and 1 / D of the holes is zero)
            T ua[PAD], ub[PAD], di[PAD];
            ldv(W, ra + ch, ua); ldv(W, rb + ch, ub); ldv(W, Di + ec0 + ch, di);
#pragma unroll
            for (int cc = 0; cc < PAD; ++cc) acc[cc] += (ua[cc] * di[cc]) * ub[cc];
          }
          T tot = acc[0];
#pragma unroll
          for (int cc = 1; cc < PAD; ++cc) tot += acc[cc];
          W[tg] -= tot;
        }
        __syncthreads();
      }
      mark(kPhFactorExt);
    }
    // scale the panels, L = (unscaled columns) D^-1, and clear the diagonal and the padding of the rows of the diagonal blocks:
    // a thread owns (panel row, column group); 1 / D as a vector (zero in the holes)
    for (int sq = 0; sq < S.ns; ++sq) {
      const int4 h = *reinterpret_cast<const int4*>(sn + 8 * sq);  // c0, s, t, dbase
      const int sp = sn[8 * sq + 6];
      const unsigned short* pr = prow + sn[8 * sq + 7];
      const int NCG = sp >> LPAD, tasks = (h.y + h.z) * NCG;
      for (int e = tid; e < tasks; e += NT) {
        const int p = e / NCG, j0 = (e - p * NCG) << LPAD;
        if (p < h.y && j0 > p) continue;  // row p of the diagonal block holds the columns <= p
        T lw[PAD], di[PAD];
        const int slot = pr[p] + j0;
        ldv(W, slot, lw); ldv(W, Di + h.x + j0, di);
#pragma unroll
        for (int cc = 0; cc < PAD; ++cc) lw[cc] = (p < h.y && j0 + cc >= p) ? T(0) : lw[cc] * di[cc];
        stv(W, slot, lw);
      }
    }
    __syncthreads();
    // invert the unit-lower diagonal blocks in place: L^-1 = (I - l_{s-1} e^T) ... (I - l_0 e^T), applied right-looking.  With
    // X' = -(strict lower part of L^-1) stored, step k is X'(i, c) -= L(i, k) X'(k, c) for i > k, c < k (column k itself is already
    // X').  A thread owns a row i of a diagonal block and updates it with 16-byte vectors; row k is final after step k - 1 and its
    // diagonal and padding are zero, so whole vectors are exact.  One barrier per step, all supernodes at once.
    {
      constexpr int KS = 2;  // n <= KS * NT (checked on the host)
      int rbase[KS], ri[KS];
      const unsigned short* rq[KS];
      const int* fd = itab(kI_fdiag);
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int e = tid + ks * NT;
        ri[ks] = 0; rbase[ks] = 0; rq[ks] = prow;
        if (e < n) {
          const int wd = fd[e];  // column | supernode << 16, every real column once
          const int q = hi16(wd);
          ri[ks] = lo16(wd) - sn[8 * q];
          rq[ks] = prow + sn[8 * q + 7];
          rbase[ks] = rq[ks][ri[ks]];
        }
      }
      for (int k = 1; k + 1 < S.smax; ++k) {
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          if (k < ri[ks]) {
            const T lik = W[rbase[ks] + k];
            const int rowk = rq[ks][k];
            for (int ch = 0; ch < k; ch += 2 * PAD) {  // two vectors per trip: the loads of the second do not wait for the store of the first
              T x0[PAD], r0[PAD], x1[PAD], r1[PAD];
              const bool two = ch + PAD < k;
              ldv(W, rbase[ks] + ch, x0); ldv(W, rowk + ch, r0);
              if (two) { ldv(W, rbase[ks] + ch + PAD, x1); ldv(W, rowk + ch + PAD, r1); }
#pragma unroll
              for (int cc = 0; cc < PAD; ++cc) x0[cc] -= lik * r0[cc];
              stv(W, rbase[ks] + ch, x0);
              if (two) {
#pragma unroll
                for (int cc = 0; cc < PAD; ++cc) x1[cc] -= lik * r1[cc];
                stv(W, rbase[ks] + ch + PAD, x1);
              }
            }
          }
        }
        __syncthreads();
      }
    }
    ok = !bany(!ok);
    mark(kPhFactorInv);
    return ok;
  }

  // ---------------------------------------------------------------- v <- (L D L^T)^-1 v, result in sv
  // Stage list of qp_sparse_cta_host.hpp.  G = 1 << lg lanes cooperate on one output; every inner loop moves whole 16-byte
  // vectors of the factor (block rows are padded with zeros).  Each stage ends with a barrier.  The stage bodies are kept SMALL
  // (run-time lane-group size, no unrolling): the ADMM loop runs 15 of them per iteration and must stay inside the instruction
  // cache -- templated on the group size the sweeps alone were 72 KB of SASS and every stage started with instruction-fetch misses.
  __device__ void stage_fwd_diag(int first, int count, int lg)
  {
    const int G = 1 << lg;
    const int *fd = itab(kI_fdiag), *sn = itab(kI_sntab);
    const unsigned short* doff = htab(kI_diagoff);
    const int group = tid >> lg, g = tid & (G - 1);  // lanes of a group are adjacent: contiguous reads of a row
    for (int o0 = 0; o0 < count; o0 += NT >> lg) {
      const int o = o0 + group;
      const bool valid = o < count;
      T acc = T(0);
      int k = 0;
      if (valid) {
        const int wd = fd[first + o];
        k = lo16(wd);
        const int q = hi16(wd);
        const int c0 = sn[8 * q], i = k - c0, row = sn[8 * q + 3] + doff[i];
#pragma unroll 2
        for (int ch = g << LPAD; ch < i; ch += G << LPAD) {
          T lw[PAD], vv[PAD];
          ldv(W, row + ch, lw); ldv(v, c0 + ch, vv);
#pragma unroll
          for (int cc = 0; cc < PAD; ++cc) acc += lw[cc] * vv[cc];
        }
      }
      for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (valid && g == 0) sv[k] = v[k] - acc;
    }
  }
  __device__ void stage_fwd_push(int first, int count, int lg)
  {
    const int G = 1 << lg;
    const int *pout = itab(kI_pushout), *ptask = itab(kI_pushtask), *sn = itab(kI_sntab);
    const int group = tid >> lg, g = tid & (G - 1);
    for (int o0 = 0; o0 < count; o0 += NT >> lg) {
      const int o = o0 + group;
      const bool valid = o < count;
      T acc = T(0);
      int dst = 0;
      if (valid) {
        const int w0 = pout[first + o], w1 = pout[first + o + 1];
        dst = lo16(w0);
        for (int k = hi16(w0); k < hi16(w1); ++k) {
          const int tw = ptask[k];
          const int q = lo16(tw);
          const int c0 = sn[8 * q], sp = sn[8 * q + 6], row = sn[8 * q + 4] + hi16(tw) * sp;
#pragma unroll 2
          for (int ch = g << LPAD; ch < sp; ch += G << LPAD) {
            T lw[PAD], yv[PAD];
            ldv(W, row + ch, lw); ldv(sv, c0 + ch, yv);
#pragma unroll
            for (int cc = 0; cc < PAD; ++cc) acc += lw[cc] * yv[cc];
          }
        }
      }
      for (int off = G >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
      if (valid && g == 0) v[dst] -= acc;
    }
  }
  // backward stages: an output is a group of PAD adjacent columns; lanes of adjacent outputs are adjacent
  template <bool PULL> __device__ void stage_bwd(int first, int count, int lg)
  {
    const int G = 1 << lg, lopw = 5 - lg, OPW = 1 << lopw;
    const int *bo = itab(kI_bwd), *sn = itab(kI_sntab);
    const unsigned short *doff = htab(kI_diagoff), *rl = htab(kI_rlist);
    const int g = lane >> lopw, ow = lane & (OPW - 1);
    for (int o0 = 0; o0 < count; o0 += NT >> lg) {
      const int o = o0 + (warp << lopw) + ow;
      const bool valid = o < count;
      T acc[PAD];
#pragma unroll
      for (int cc = 0; cc < PAD; ++cc) acc[cc] = T(0);
      int k0 = 0;
      if (valid) {
        const int wd = bo[first + o];
        k0 = lo16(wd);
        const int q = hi16(wd);
        const int4 h = *reinterpret_cast<const int4*>(sn + 8 * q);  // c0, s, t, dbase
        const int kc = k0 - h.x;
        if (PULL) {
          const int sp = sn[8 * q + 6], base = sn[8 * q + 4] + kc, rb = sn[8 * q + 5];
#pragma unroll 2
          for (int r = g; r < h.z; r += G) {
            T lw[PAD];
            ldv(W, base + r * sp, lw);
            const T xr = sv[rl[rb + r]];
#pragma unroll
            for (int cc = 0; cc < PAD; ++cc) acc[cc] += lw[cc] * xr;
          }
        } else {
          const int base = h.w + kc;
#pragma unroll 2
          for (int i = kc + 1 + g; i < h.y; i += G) {  // columns >= i of row i are zeros
            T lw[PAD];
            ldv(W, base + doff[i], lw);
            const T vi = v[h.x + i];
#pragma unroll
            for (int cc = 0; cc < PAD; ++cc) acc[cc] += lw[cc] * vi;
          }
        }
      }
      for (int off = 16; off >= OPW; off >>= 1) {
#pragma unroll
        for (int cc = 0; cc < PAD; ++cc) acc[cc] += __shfl_xor_sync(0xffffffffu, acc[cc], off);
      }
      if (valid && g == 0) {  // whole vectors: the holes come out as zero (1 / D of a hole and every padding entry are zero)
        T o4[PAD];
        if (PULL) {
          T di[PAD], yv[PAD];
          ldv(W, nW + k0, di); ldv(sv, k0, yv);
#pragma unroll
          for (int cc = 0; cc < PAD; ++cc) o4[cc] = di[cc] * yv[cc] - acc[cc];
          stv(v, k0, o4);
        } else {
          T vv[PAD];
          ldv(v, k0, vv);
#pragma unroll
          for (int cc = 0; cc < PAD; ++cc) o4[cc] = vv[cc] - acc[cc];
          stv(sv, k0, o4);
        }
      }
    }
  }
  __device__ void solve()
  {
    const int4* stages = reinterpret_cast<const int4*>(itab(kI_stages));
    long long tstage = (prof != nullptr && tid == 0) ? clock64() : 0;
    int4 nxt = stages[0];  // kind, first, count, log2(lanes per output); the next descriptor is fetched a stage ahead
    for (int st = 0; st < S.nstages; ++st) {
      const int4 sd = nxt;
      if (st + 1 < S.nstages) nxt = stages[st + 1];
      if (sd.x == kStageFwdDiag) stage_fwd_diag(sd.y, sd.z, sd.w);
      else if (sd.x == kStageFwdPush) stage_fwd_push(sd.y, sd.z, sd.w);
      else if (sd.x == kStageBwdPull) stage_bwd<true>(sd.y, sd.z, sd.w);
      else stage_bwd<false>(sd.y, sd.z, sd.w);
      __syncthreads();
      if (prof != nullptr && tid == 0 && st < 32) {  // per-stage split of the solve phase (the phase total is accumulated by mark() as well)
        const long long t = clock64();
        atomicAdd(prof + kPhStage0 + st, (unsigned long long)(t - tstage));
        tstage = t;
      }
    }
  }

  // ---------------------------------------------------------------- check_stopping, qp_solver.hpp:574-644
  // Same evaluation as qp_sparse_tiled.cuh::check_stopping.  Clobbers v, sv, w, t1.
  __device__ int check_stopping(const sfb_qp_params& prm, bool dinf_guard)
  {
    const T eps_abs = T(prm.eps_abs), eps_rel = T(prm.eps_rel);
    const T eps_pinf = T(prm.eps_primal_inf), eps_dinf = T(prm.eps_dual_inf);
    const T inf = Num<T>::inf();
    T r1[3] = {T(0), T(0), T(0)};  // dxn, Edy (max), qdx (sum)
    for (int j = tid; j < np; j += NT) {
      const T xj = x[j];
      const T d = xj - xold[j];
      t1[j] = sx[j] * xj;       // x_us   :481
      sv[j] = d;                // scaled dx
      const T dus = sx[j] * d;  // dx_us  :484
      v[j] = dus;
      r1[0] = fmax(r1[0], fabs(dus));
      r1[2] += qg(j) * dus;
    }
    for (int i = tid; i < m; i += NT) {
      const T d = y[i] - yold[i];
      w[i] = d;
      r1[1] = fmax(r1[1], fabs(sy[i] * d / c));  // :485
    }
    reduce<3, 0x3u>(r1);
    const T qn = qn_us, dxn = r1[0], Edy = r1[1], qdx = r1[2];
    T r2[4] = {T(0), T(0), T(0), T(0)};  // n_Ax, n_r, n_z (max), s_pinf (sum)
    bool pinf_blocked = false, dinf_rows_bad = false;
    for (int i = tid; i < m; i += NT) {
      T ax = A_row_dot(i, x), adx = A_row_dot(i, sv);
      const T syinv = T(1) / sy[i];
      ax *= syinv;
      adx *= syinv;
      const T zus = syinv * z[i];  // :483
      r2[0] = fmax(r2[0], fabs(ax));
      r2[1] = fmax(r2[1], fabs(ax - zus));
      r2[2] = fmax(r2[2], fabs(zus));
      const T dyus = sy[i] * w[i] / c;
      const T li = lg(i), ui = ug(i);
      if (ui != inf) r2[3] += ui * fmax(T(0), dyus);  // :602-617
      else if (dyus > eps_pinf * Edy) pinf_blocked = true;
      if (li != -inf) r2[3] += li * fmin(T(0), dyus);
      else if (dyus < -eps_pinf * Edy) pinf_blocked = true;
      bool okrow;
      if (ui == inf) okrow = (adx >= -eps_dinf * dxn);  // :631-639
      else if (li == -inf) okrow = (adx <= eps_dinf * dxn);
      else okrow = (fabs(adx) < eps_dinf * dxn);
      dinf_rows_bad = dinf_rows_bad || !okrow;
    }
    reduce<4, 0x7u>(r2);
    const T n_Ax = r2[0], n_r = r2[1], n_z = r2[2];
    T s_pinf = r2[3];
    pinf_blocked = bany(pinf_blocked);
    const bool dinf_rows_ok = !bany(dinf_rows_bad);
    if (pinf_blocked) s_pinf = inf;
    // Abar^T y, Abar^T dy; P x_us, P dx_us with the entries as stored -- all consumed by the thread that forms them
    T r3[5] = {T(0), T(0), T(0), T(0), T(0)};  // n_Px, n_Aty, n_res, n_Atdy, n_Pdx
    for (int j = tid; j < np; j += NT) {
      T a1, a2;
      At_col_dot2(j, y, w, a1, a2);
      T px = T(0), pdx = T(0);
      for (int e = S.PR_ptr[j]; e < S.PR_ptr[j + 1]; ++e) {
        const T pv = Pg(__ldg(S.PR_slot + e));
        const int k = __ldg(S.PR_col + e);
        px += pv * t1[k];
        pdx += pv * v[k];
      }
      const T sc = T(1) / (sx[j] * c);
      const T aty = a1 * sc, atdy = a2 * sc;
      r3[0] = fmax(r3[0], fabs(px));
      r3[1] = fmax(r3[1], fabs(aty));
      r3[2] = fmax(r3[2], fabs(px + qg(j) + aty));
      r3[3] = fmax(r3[3], fabs(atdy));
      r3[4] = fmax(r3[4], fabs(pdx));
    }
    reduce<5, 0x1fu>(r3);
    const T n_Px = r3[0], n_Aty = r3[1], n_res = r3[2], n_Atdy = r3[3], n_Pdx = r3[4];
    if (n_r <= eps_abs + eps_rel * fmax(n_Ax, n_z)) {  // :584-594
      const T dual_scale = fmax(fmax(n_Px, qn), n_Aty);
      if (n_res <= eps_abs + eps_rel * dual_scale) return SFB_QP_OPTIMAL;
    }
    if (fmax(n_Atdy, s_pinf) < eps_pinf * Edy) return SFB_QP_PRIMAL_INFEASIBLE;  // :619
    if ((dxn > T(0) || !dinf_guard) && (n_Pdx <= eps_dinf * dxn) && (qdx <= eps_dinf * dxn) && dinf_rows_ok) return SFB_QP_DUAL_INFEASIBLE;
    return kStatusUnset;
  }

  // ---------------------------------------------------------------- detail::polish_qp, qp_solver.hpp:92-204
  // As in qp_sparse_tiled.cuh::polish: the regularised system is solved through the SAME symbolic factor,
  // (Pbar + delta I + Aa^T Aa / delta) s1 = r1 + Aa^T r2 / delta, s2 = (Aa s1 - r2) / delta; the reference's outer iteration
  // t += Hp^-1 (h - H t) absorbs the error of the ill-conditioned reduced solve.  On entry: w[i] = 1 on active rows else 0,
  // yold = scaled active bound.  Row-space vectors are kept exactly zero on inactive rows.  t = (t1, rinv), h = (-qb, yold).
  __device__ unsigned polish(const sfb_qp_params& prm)
  {
    const T delta = T(prm.delta), dinv = T(1) / delta;
    for (int i = tid; i < m; i += NT) rho[i] = (w[i] != T(0)) ? dinv : T(0);
    __syncthreads();
    assemble(delta, rho);
    if (!factor()) return SFB_QP_FLAG_POLISH_FAILED;
    for (int j = tid; j < np; j += NT) t1[j] = T(0);
    for (int i = tid; i < m; i += NT) rinv[i] = T(0);
    __syncthreads();
    for (unsigned it = 0; it < prm.polish_iter; ++it) {
      // r = h - H t:  r1 = -qb - sym(Pbar) t1 - Aa^T rinv -> v,  r2 = bnd - Aa t1 on active rows -> z
      for (int j = tid; j < np; j += NT) v[j] = -qb[j] - Psym_row_dot(j, t1) - At_col_dot(j, rinv);
      for (int i = tid; i < m; i += NT) z[i] = (w[i] != T(0)) ? yold[i] - A_row_dot(i, t1) : T(0);
      __syncthreads();
      // s = Hp^-1 r -> (sv, s2):  (..) s1 = r1 + Aa^T r2 / delta,  s2 = (Aa s1 - r2) / delta
      for (int j = tid; j < np; j += NT) v[j] += dinv * At_col_dot(j, z);
      __syncthreads();
      solve();
      T dm[2] = {T(0), T(0)};
      for (int i = tid; i < m; i += NT) {
        const T s2 = (w[i] != T(0)) ? (A_row_dot(i, sv) - z[i]) * dinv : T(0);
        const T tn = rinv[i] + s2;
        dm[0] = fmax(dm[0], fabs(s2));
        dm[1] = fmax(dm[1], fabs(tn));
        rinv[i] = tn;  // read by its own thread only in this loop
      }
      for (int j = tid; j < np; j += NT) {
        const T s1 = sv[j], tn = t1[j] + s1;
        dm[0] = fmax(dm[0], fabs(s1));
        dm[1] = fmax(dm[1], fabs(tn));
        t1[j] = tn;
      }
      reduce<2, 0x3u>(dm);
      // see qp_sparse_tiled.cuh::polish.  A sweep's correction is the error BEFORE it; the fp64 polish pass of an fp32 solve
      // (TIO = float: parity bar 1e-3, the inputs themselves carry 6e-8) stops at 1e-8 instead of 1e-12
      if (dm[0] <= T(sizeof(TIO) == 4 ? 1e-8 : 1e-12) * dm[1]) break;
    }
    bool bad = false;
    for (int j = tid; j < np; j += NT) bad = bad || !(fabs(t1[j]) < Num<T>::inf());
    if (bany(bad)) return SFB_QP_FLAG_POLISH_FAILED;
    for (int j = tid; j < np; j += NT) x[j] = t1[j];  // :199
    for (int i = tid; i < m; i += NT)
      if (w[i] != T(0)) y[i] = rinv[i];  // :200-201
    __syncthreads();
    return SFB_QP_FLAG_POLISHED;
  }

  // ---------------------------------------------------------------- QPSolver::solve, qp_solver.hpp:343-568
  __device__ void run(long long inst)
  {
    b = inst;
    const T inf = Num<T>::inf();
    const sfb_qp_params& prm = a.prm;
    const unsigned long long t0 = prm.has_max_time ? global_timer_ns() : 0ull;
    const unsigned short *arp = htab(kI_Arowptr), *acol = htab(kI_Acol);
    prof = a.prof;
    if (prof != nullptr && tid == 0) tlast = clock64();
    {
      const TIO* gA = a.A + b * (long long)S.nnzA;
      for (int e = tid; e < S.nnzA; e += NT) A[e] = (T)__ldg(gA + e);
      T qn = T(0);
      for (int j = tid; j < n; j += NT) qn = fmax(qn, fabs((T)__ldg(a.q + b * (long long)n + j)));
      qn_us = bmax(qn);
    }
    mark(kPhLoad);
    TIO* const sws = a.scale_ws != nullptr ? a.scale_ws + b * (long long)(n + m + 1) : nullptr;
    if (a.mode == 2 && sws != nullptr) {
      // polish pass of a lower-precision solve: reuse ITS equilibration (any positive scaling is a valid one for polish_qp, and the
      // iterate to polish was computed in that scaling) instead of ten more passes over P and A
      for (int pj = tid; pj < np; pj += NT) {
        const int jo = __ldg(S.perm + pj);
        sx[pj] = jo >= 0 ? (T)sws[jo] : T(1);
      }
      for (int i = tid; i < m; i += NT) sy[i] = (T)sws[n + i];
      c = (T)sws[n + m];
      __syncthreads();
    } else if (prm.scaling) scale();  // :347
    else {
      c = T(1);
      for (int j = tid; j < np; j += NT) sx[j] = T(1);
      for (int i = tid; i < m; i += NT) sy[i] = T(1);
      __syncthreads();
    }
    if (a.mode != 2 && sws != nullptr) {
      for (int pj = tid; pj < np; pj += NT) {
        const int jo = __ldg(S.perm + pj);
        if (jo >= 0) sws[jo] = (TIO)sx[pj];
      }
      for (int i = tid; i < m; i += NT) sws[n + i] = (TIO)sy[i];
      if (tid == 0) sws[n + m] = (TIO)c;
    }
    mark(kPhScale);
    // ---- rho classes + trivially empty feasible set  :361-374
    int code = kStatusUnset;
    const T rho_bar = T(prm.rho), sigma = T(prm.sigma), alpha = T(prm.alpha), alpha_comp = T(1) - alpha;
    bool triv = false;
    for (int i = tid; i < m; i += NT) {
      const T li = lg(i), ui = ug(i);
      if (li == inf || ui == -inf || ui - li < T(0)) triv = true;
      T rr;
      if (li == -inf && ui == inf) rr = T(1e-6);
      else if (sy[i] * fabs(li - ui) < T(1e-5)) rr = T(1e3) * rho_bar;
      else rr = rho_bar;
      rho[i] = rr;
      rinv[i] = T(1) / rr;
      lo[i] = sy[i] * li;
      hi[i] = sy[i] * ui;
    }
    if (bany(triv)) code = SFB_QP_PRIMAL_INFEASIBLE;
    // ---- scaled data: qb = c Sx q, Abar = Sy A Sx  (:401-403, :450)
    for (int j = tid; j < np; j += NT) qb[j] = (c * sx[j]) * qg(j);
    for (int i = tid; i < m; i += NT) {
      const T syi = sy[i];
      const int e1 = arp[i + 1];
      for (int e = arp[i]; e < e1; ++e) A[e] = (syi * sx[acol[e]]) * A[e];
    }
    __syncthreads();
    const bool polish_only = a.mode == 2;
    const bool skip = polish_only && a.out_status[b] != (int32_t)SFB_QP_OPTIMAL;
    mark(kPhPrep);
    if (!polish_only) {
      assemble(sigma, rho);
      mark(kPhAssemble);
      if (!factor()) code = SFB_QP_UNKNOWN;  // :433
    }
    // ---- initial iterate  :436-445
    if (polish_only) {
      for (int pj = tid; pj < np; pj += NT) {
        const int jo = __ldg(S.perm + pj);
        x[pj] = jo >= 0 ? (T(1) / sx[pj]) * (T)a.out_x[b * (long long)n + jo] : T(0);
      }
      for (int i = tid; i < m; i += NT) {
        y[i] = c * ((T(1) / sy[i]) * (T)a.out_y[b * (long long)m + i]);
        z[i] = T(0);
      }
    } else if (a.warm_x != nullptr) {
      for (int pj = tid; pj < np; pj += NT) {
        const int jo = __ldg(S.perm + pj);
        x[pj] = jo >= 0 ? (T(1) / sx[pj]) * (T)__ldg(a.warm_x + b * (long long)n + jo) : T(0);
      }
      __syncthreads();
      for (int i = tid; i < m; i += NT) {
        y[i] = c * ((T(1) / sy[i]) * (T)__ldg(a.warm_y + b * (long long)m + i));
        z[i] = A_row_dot(i, x);
      }
    } else {
      for (int j = tid; j < np; j += NT) x[j] = T(0);
      for (int i = tid; i < m; i += NT) { y[i] = T(0); z[i] = T(0); }
    }
    __syncthreads();
    for (int i = tid; i < m; i += NT) w[i] = rho[i] * z[i] - y[i];
    __syncthreads();
    mark(kPhPrep);

    // ---- main loop  :449-510
    const unsigned sci = prm.stop_check_iter;
    unsigned iter = 0;
    if (polish_only) { code = skip ? kStatusUnset - 1 : (int)SFB_QP_OPTIMAL; iter = a.out_iter[b]; }
    for (; iter != a.max_iter_eff && code == kStatusUnset; ++iter) {
      // rhs = sigma x - qb + Abar^T w   (reduced system)
      for (int j = tid; j < np; j += NT) v[j] = sigma * x[j] - qb[j] + At_col_dot(j, w);
      __syncthreads();
      mark(kPhRhs);
      solve();
      mark(kPhSolve);
      const bool chk = (iter % sci == 1u);
      for (int j = tid; j < np; j += NT) {
        const T xi = x[j];
        if (chk) xold[j] = xi;  // :465-468
        x[j] = alpha * sv[j] + alpha_comp * xi;  // :470
      }
      for (int i = tid; i < m; i += NT) {
        const T zi = z[i], yi = y[i], ri = rho[i], rinvi = rinv[i];
        const T lob = lo[i], hib = hi[i];
        const T zt = A_row_dot(i, sv);
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
      __syncthreads();
      mark(kPhUpdate);
      if (chk) {
        code = check_stopping(prm, a.dinf_guard != 0);  // :488 (clobbers w, v, sv, t1)
        if (code == kStatusUnset && prm.has_max_time) {
          const bool late = (long long)(global_timer_ns() - t0) > prm.max_time_ns;  // :504-508
          if (bany(late)) code = SFB_QP_MAX_TIME;
        }
        __syncthreads();
        for (int i = tid; i < m; i += NT) w[i] = rho[i] * z[i] - y[i];
        __syncthreads();
        mark(kPhCheck);
      }
    }

    // ---- active sets as polish_qp builds them (:113-123) on the scaled dual; yold <- scaled active bound
    const T thr = T(100) * Num<T>::eps();
    for (int i = tid; i < m; i += NT) {
      int act = 0;
      T bv = T(0);
      const T li = lg(i), ui = ug(i);
      if (polish_only) {
        act = a.out_active[b * (long long)m + i];
        if (act < 0) bv = sy[i] * li;
        if (act > 0) bv = sy[i] * ui;
      } else {
        if (y[i] < -thr && li != -inf) { act = -1; bv = sy[i] * li; }
        if (y[i] > thr && ui != inf) { act = 1; bv = sy[i] * ui; }
        if (a.out_active) a.out_active[b * (long long)m + i] = (int8_t)act;
      }
      w[i] = (act != 0) ? T(1) : T(0);
      yold[i] = bv;
    }
    __syncthreads();
    unsigned flags = 0;
    if (code == SFB_QP_OPTIMAL && prm.polish) {
      if (sizeof(T) == 4) flags = SFB_QP_FLAG_POLISH_SKIPPED;  // delta = 1e-6 is not resolvable in fp32: second pass in fp64 (mode 2)
      else flags = polish(prm);
    }
    mark(kPhPolish);
    if (skip) return;
    // ---- unscale + objective  :544-548
    for (int j = tid; j < np; j += NT) t1[j] = sx[j] * x[j];
    __syncthreads();
    T obj = T(0);
    for (int j = tid; j < np; j += NT) {
      const int jo = __ldg(S.perm + j);
      if (jo < 0) continue;
      T px = T(0);
      for (int e = S.PR_ptr[j]; e < S.PR_ptr[j + 1]; ++e) px += (T(0.5) * Pg(__ldg(S.PR_slot + e))) * t1[__ldg(S.PR_col + e)];
      const T xv = t1[j];
      a.out_x[b * (long long)n + jo] = (TIO)xv;
      obj += xv * (px + (T)__ldg(a.q + b * (long long)n + jo));
    }
    obj = bsum(obj);
    for (int i = tid; i < m; i += NT) a.out_y[b * (long long)m + i] = (TIO)(sy[i] * y[i] / c);
    if (tid == 0) {
      a.out_obj[b] = (TIO)obj;
      a.out_status[b] = (code == kStatusUnset) ? (int32_t)SFB_QP_MAX_ITERATIONS : (int32_t)code;
      a.out_iter[b] = iter;
      if (a.out_flags) a.out_flags[b] = flags;
    }
    mark(kPhOut);
  }
};

// GT: the two big index tables stay in global memory (fp32: the working set then fits twice per SM)
template <typename T, typename TIO, bool GT> __global__ void __launch_bounds__(kCtaNT, GT ? 2 : 1) qp_sparse_cta_kernel(const CtaArgs<T, TIO> a)
{
  __shared__ long long next_inst;
  CtaSolver<T, TIO, kCtaNT, GT> s(a);
  s.load_tables();
  for (;;) {
    if (threadIdx.x == 0) next_inst = (long long)atomicAdd(a.work_counter, 1ull);
    __syncthreads();
    const long long b = next_inst;
    __syncthreads();
    if (b >= a.batch) break;
    s.run(b);
  }
}

}  // namespace sfb
