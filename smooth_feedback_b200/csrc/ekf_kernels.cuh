// ekf_kernels.cuh -- batched EKF covariance algebra for sm_100a, one thread per filter instance.
//
// Replaces (reference paths relative to pettni/smooth_feedback @ 9a08971):
//   EKF::predict, covariance half   include/smooth/feedback/ekf.hpp:79-103
//       Pdot = selfadjointView<Upper>(A P + P A^T + Q) stepped with Boost.odeint euler / runge_kutta4
//   EKF::update                     include/smooth/feedback/ekf.hpp:116-139
//       S = triu(H symU(P) H^T + R), K = (S.ldlt().solve(H P))^T, P <- symU((I - K H) P)
//
// These are d x d problems with d ~ 6: ~1.3 flop per byte, i.e. HBM-bound.  The batch arrives in the
// reference's array-of-matrices layout; a CTA stages a tile of instances through shared memory with fully
// coalesced loads, transposing to [element][thread] (stride blockDim+1: conflict-free for both the
// cooperative copy and the per-thread algebra), runs the algebra thread-per-instance out of shared
// memory, and stores the results coalesced again.

#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sfb {

template <typename T> struct EkfPredictArgs
{
  const T* P;
  const T* A;
  const T* Q;
  T* out_P;
  long long batch;
  int d;
  int stepper;  // 0 euler, 1 rk4
  T tau, dt;
};

template <typename T> struct EkfUpdateArgs
{
  const T* P;
  const T* H;
  const T* R;
  const T* innov;
  T* out_delta;
  T* out_P;
  long long batch;
  int d, ny;
};

// cooperative, coalesced tile copy: global [cnt][len] contiguous  <->  shared [len][stride] (+ thread index)
template <typename T>
__device__ __forceinline__ void tile_load(T* sm, int stride, const T* g, int cnt, int len)
{
  const int total = cnt * len;
  int t = threadIdx.x / len, e = threadIdx.x - t * len;
  const int dt_ = blockDim.x / len, de = blockDim.x - dt_ * len;
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    sm[e * stride + t] = __ldg(g + k);
    t += dt_;
    e += de;
    if (e >= len) { e -= len; ++t; }
  }
}
template <typename T>
__device__ __forceinline__ void tile_store(const T* sm, int stride, T* g, int cnt, int len)
{
  const int total = cnt * len;
  int t = threadIdx.x / len, e = threadIdx.x - t * len;
  const int dt_ = blockDim.x / len, de = blockDim.x - dt_ * len;
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    g[k] = sm[e * stride + t];
    t += dt_;
    e += de;
    if (e >= len) { e -= len; ++t; }
  }
}

// per-thread view of a column-major matrix living in the [element][thread] shared layout
template <typename T> struct SmMat
{
  T* p;
  int stride, ld;
  __device__ __forceinline__ T& operator()(int i, int j) const { return p[(i + ld * j) * stride]; }
  __device__ __forceinline__ T& operator[](int e) const { return p[e * stride]; }
};

// dcov = selfadjointView<Upper>(A cov + cov A^T + Q)   ekf.hpp:88
template <typename T>
__device__ __forceinline__ void cov_ode(int d, const SmMat<T>& A, const SmMat<T>& Q, const SmMat<T>& cov,
                                        const SmMat<T>& dcov)
{
  for (int j = 0; j < d; ++j) {
    for (int i = 0; i <= j; ++i) {
      T a = T(0);
      for (int k = 0; k < d; ++k) a += A(i, k) * cov(k, j);
      for (int k = 0; k < d; ++k) a += cov(i, k) * A(j, k);
      a += Q(i, j);
      dcov(i, j) = a;
      dcov(j, i) = a;
    }
  }
}

// shared memory: 6 matrices of d*d per thread
template <typename T> __global__ void ekf_predict_kernel(const EkfPredictArgs<T> a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int d = a.d, dd = d * d, BD = blockDim.x, stride = BD + 1;
  const size_t msz = (size_t)dd * stride;
  T* sP = sm;
  T* sA = sm + msz;
  T* sQ = sm + 2 * msz;
  T* sX = sm + 3 * msz;
  T* sK = sm + 4 * msz;
  T* sAcc = sm + 5 * msz;
  const long long ntiles = (a.batch + BD - 1) / BD;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * BD;
    const int cnt = (int)((a.batch - b0 < BD) ? (a.batch - b0) : BD);
    tile_load(sP, stride, a.P + b0 * dd, cnt, dd);
    tile_load(sA, stride, a.A + b0 * dd, cnt, dd);
    tile_load(sQ, stride, a.Q + b0 * dd, cnt, dd);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const int t = threadIdx.x;
      const SmMat<T> P{sP + t, stride, d}, A{sA + t, stride, d}, Q{sQ + t, stride, d}, X{sX + t, stride, d},
        K{sK + t, stride, d}, Acc{sAcc + t, stride, d};
      auto step = [&](T h) {
        if (a.stepper == 0) {
          cov_ode(d, A, Q, P, K);
          for (int e = 0; e < dd; ++e) P[e] = P[e] + h * K[e];
        } else {
          cov_ode(d, A, Q, P, K);  // k1
          for (int e = 0; e < dd; ++e) {
            X[e] = P[e] + (h * T(0.5)) * K[e];
            Acc[e] = P[e] + (h / T(6)) * K[e];
          }
          cov_ode(d, A, Q, X, K);  // k2
          for (int e = 0; e < dd; ++e) {
            X[e] = P[e] + (h * T(0.5)) * K[e];
            Acc[e] = Acc[e] + (h / T(3)) * K[e];
          }
          cov_ode(d, A, Q, X, K);  // k3
          for (int e = 0; e < dd; ++e) {
            X[e] = P[e] + h * K[e];
            Acc[e] = Acc[e] + (h / T(3)) * K[e];
          }
          cov_ode(d, A, Q, X, K);  // k4
          for (int e = 0; e < dd; ++e) P[e] = Acc[e] + (h / T(6)) * K[e];
        }
      };
      // step schedule of ekf.hpp:91-102
      T tt = T(0);
      const T dt_v = (a.dt > T(0)) ? a.dt : T(2) * a.tau;
      while (tt + dt_v < a.tau) {
        step(dt_v);
        tt += dt_v;
      }
      step(a.tau - tt);
    }
    __syncthreads();
    tile_store(sP, stride, a.out_P + b0 * dd, cnt, dd);
    __syncthreads();
  }
}

// Eigen-style LDL^T (largest stored |diagonal| first, left-looking) of the ny x ny matrix whose lower triangle is
// in W; then in-place solves.  Mirrors oracle/sf_oracle.cpp::Ldlt (Eigen 3.4 LDLT<.,Upper> semantics).
template <typename T>
__device__ void ldlt_small(int sz, const SmMat<T>& W, int* tr, const SmMat<T>& tmp)
{
  if (sz <= 1) {
    for (int i = 0; i < sz; ++i) tr[i] = i;
    return;
  }
  for (int kk = 0; kk < sz; ++kk) {
    int big = kk;
    T bigv = fabs(W(kk, kk));
    for (int j = kk + 1; j < sz; ++j) {
      const T v = fabs(W(j, j));
      if (v > bigv) { bigv = v; big = j; }
    }
    tr[kk] = big;
    if (kk != big) {
      const int s = sz - big - 1;
      for (int j = 0; j < kk; ++j) { const T t = W(kk, j); W(kk, j) = W(big, j); W(big, j) = t; }
      for (int i = 0; i < s; ++i) { const T t = W(big + 1 + i, kk); W(big + 1 + i, kk) = W(big + 1 + i, big); W(big + 1 + i, big) = t; }
      { const T t = W(kk, kk); W(kk, kk) = W(big, big); W(big, big) = t; }
      for (int i = kk + 1; i < big; ++i) { const T t = W(i, kk); W(i, kk) = W(big, i); W(big, i) = t; }
    }
    if (kk > 0) {
      for (int j = 0; j < kk; ++j) tmp[j] = W(j, j) * W(kk, j);
      T acc = T(0);
      for (int j = 0; j < kk; ++j) acc += W(kk, j) * tmp[j];
      W(kk, kk) -= acc;
      for (int i = kk + 1; i < sz; ++i) {
        T a2 = T(0);
        for (int j = 0; j < kk; ++j) a2 += W(i, j) * tmp[j];
        W(i, kk) -= a2;
      }
    }
    const T akk = W(kk, kk);
    if (fabs(akk) > T(0)) {
      for (int i = kk + 1; i < sz; ++i) W(i, kk) /= akk;
    }
  }
}

// b (strided column in shared memory, element e at b[e]) <- S^-1 b
template <typename T>
__device__ void ldlt_small_solve(int sz, const SmMat<T>& W, const int* tr, T* b, int bstride, T tiny)
{
  auto B = [&](int i) -> T& { return b[i * bstride]; };
  for (int i = 0; i < sz; ++i) { const T t = B(i); B(i) = B(tr[i]); B(tr[i]) = t; }
  for (int j = 0; j < sz; ++j) {
    const T bj = B(j);
    for (int i = j + 1; i < sz; ++i) B(i) -= W(i, j) * bj;
  }
  for (int i = 0; i < sz; ++i) {
    const T dgl = W(i, i);
    if (fabs(dgl) > tiny) B(i) /= dgl; else B(i) = T(0);
  }
  for (int j = sz - 1; j >= 0; --j) {
    T acc = B(j);
    for (int i = j + 1; i < sz; ++i) acc -= W(i, j) * B(i);
    B(j) = acc;
  }
  for (int i = sz - 1; i >= 0; --i) { const T t = B(i); B(i) = B(tr[i]); B(tr[i]) = t; }
}

constexpr int kEkfMaxNy = 16;

// shared memory per thread: P dd, H ny*d, S ny*ny, HP ny*d, Kt ny*d, IKH dd, out dd, innov ny, delta d, tmp ny
template <typename T> __host__ __device__ inline size_t ekf_update_elems(int d, int ny)
{
  return (size_t)3 * d * d + (size_t)3 * ny * d + (size_t)ny * ny + 2 * (size_t)ny + d;
}

template <typename T> __global__ void ekf_update_kernel(const EkfUpdateArgs<T> a)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const int d = a.d, ny = a.ny, dd = d * d, nd = ny * d, nn = ny * ny, BD = blockDim.x, stride = BD + 1;
  T* sP = sm;
  T* sH = sP + (size_t)dd * stride;
  T* sS = sH + (size_t)nd * stride;
  T* sHP = sS + (size_t)nn * stride;
  T* sKt = sHP + (size_t)nd * stride;
  T* sIKH = sKt + (size_t)nd * stride;
  T* sOut = sIKH + (size_t)dd * stride;
  T* sInn = sOut + (size_t)dd * stride;
  T* sDel = sInn + (size_t)ny * stride;
  T* sTmp = sDel + (size_t)d * stride;
  const long long ntiles = (a.batch + BD - 1) / BD;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * BD;
    const int cnt = (int)((a.batch - b0 < BD) ? (a.batch - b0) : BD);
    tile_load(sP, stride, a.P + b0 * dd, cnt, dd);
    tile_load(sH, stride, a.H + b0 * nd, cnt, nd);
    tile_load(sS, stride, a.R + b0 * nn, cnt, nn);
    tile_load(sInn, stride, a.innov + b0 * ny, cnt, ny);
    __syncthreads();
    if ((int)threadIdx.x < cnt) {
      const int t = threadIdx.x;
      const SmMat<T> P{sP + t, stride, d}, H{sH + t, stride, ny}, S{sS + t, stride, ny}, HP{sHP + t, stride, ny},
        Kt{sKt + t, stride, ny}, IKH{sIKH + t, stride, d}, Out{sOut + t, stride, d}, Inn{sInn + t, stride, ny},
        Del{sDel + t, stride, d}, Tmp{sTmp + t, stride, ny};
      // HPs = H * symU(P) (into Kt for now), HP = H * P     ekf.hpp:129-130,134
      for (int i = 0; i < ny; ++i)
        for (int j = 0; j < d; ++j) {
          T s1 = T(0), s2 = T(0);
          for (int k = 0; k < d; ++k) {
            const T hik = H(i, k);
            s1 += hik * ((k <= j) ? P(k, j) : P(j, k));
            s2 += hik * P(k, j);
          }
          Kt(i, j) = s1;
          HP(i, j) = s2;
        }
      // lower triangle of S^T == upper triangle of S = HPs H^T + R (R currently in S; only triu(R) is used)
      for (int j = 0; j < ny; ++j)
        for (int i = j; i < ny; ++i) {
          T s1 = T(0);
          for (int k = 0; k < d; ++k) s1 += Kt(j, k) * H(i, k);
          Tmp[0] = s1 + S(j, i);  // S(j,i), j<=i : upper entry of R
          S(i, j) = Tmp[0];
        }
      // the upper part of S above still holds R's upper entries that later columns need: the loop above reads
      // S(j,i) with j<=i and writes S(i,j) with i>=j; for i==j it is the same cell (read before write), and a
      // strictly-lower write never clobbers an upper entry.
      int tr[kEkfMaxNy];
      ldlt_small(ny, S, tr, Tmp);
      for (int j = 0; j < d; ++j) {
        for (int i = 0; i < ny; ++i) Kt(i, j) = HP(i, j);
        ldlt_small_solve(ny, S, tr, &Kt(0, j), stride, T(sizeof(T) == 8 ? 2.2250738585072014e-308 : 1.17549435e-38));
      }
      // delta = K innov   :137
      for (int i = 0; i < d; ++i) {
        T s1 = T(0);
        for (int k = 0; k < ny; ++k) s1 += Kt(k, i) * Inn[k];
        Del[i] = s1;
      }
      // P = symU((I - K H) P)   :138
      for (int i = 0; i < d; ++i)
        for (int j = 0; j < d; ++j) {
          T s1 = T(0);
          for (int k = 0; k < ny; ++k) s1 += Kt(k, i) * H(k, j);
          IKH(i, j) = ((i == j) ? T(1) : T(0)) - s1;
        }
      for (int j = 0; j < d; ++j)
        for (int i = 0; i <= j; ++i) {
          T s1 = T(0);
          for (int k = 0; k < d; ++k) s1 += IKH(i, k) * P(k, j);
          Out(i, j) = s1;
          Out(j, i) = s1;
        }
    }
    __syncthreads();
    tile_store(sOut, stride, a.out_P + b0 * dd, cnt, dd);
    tile_store(sDel, stride, a.out_delta + b0 * d, cnt, d);
    __syncthreads();
  }
}

}  // namespace sfb
