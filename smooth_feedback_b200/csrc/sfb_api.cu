// sfb_api.cu -- host side of the C ABI declared in include/sfb.h.
//
// Thin by design: argument validation, launch geometry, and the host-buffer staging pipeline.  All numerics
// live in the kernels (qp_dense_warp.cuh, ekf_kernels.cuh).  There is no CPU implementation of anything here:
// without a CUDA device every entry point fails.

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/sfb.h"
#include "ekf_fused_tma.cuh"
#include "ekf_kernels.cuh"
#include "qp_dense_group.cuh"
#include "qp_dense_skinny.cuh"
#include "qp_sparse_host.hpp"
#include "qp_sparse_tiled.cuh"

namespace {

constexpr int kNumSlots = 3;  // staging slots of the host-buffer pipeline (H2D / compute / D2H overlap)

std::string g_create_error;

struct Slot
{
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  void* dev = nullptr;
  size_t bytes = 0;
};

// global polish workspace (only touched when an instance's Schur block does not fit in shared memory)
struct Scratch
{
  void* dev = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct sfb_context
{
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaDeviceProp prop{};
  unsigned long long* counters = nullptr;  // work-queue heads, one per launch in flight
  int num_counters = 64;
  int next_counter = 0;
  uint64_t launches = 0;
  std::string last_error;
  Slot slots[kNumSlots];
  Scratch scratch[kNumSlots + 1];  // [kNumSlots] belongs to the handle's own stream
  cudaEvent_t ev_start = nullptr;
  bool ekf_force_generic = false;
  bool dense_force_generic = false;  // SFB_DENSE_FORCE_GENERIC=1: bypass the tall-skinny register kernel (A/B measurements)
  int sparse_tw = 0;     // SFB_SPARSE_TW=4|8|32 overrides the tile-width heuristic of the sparse QP path (A/B measurements)
  Scratch sparse_ws;     // tiled working set of the sparse QP path
  Scratch sparse_stage;  // device copies of host buffers (sparse path)
};


// device-resident result of sfb_qp_sparse_analyze (index arrays shared by every instance of a batch)
struct sfb_qp_sparse_pattern
{
  int device = 0;
  sfb::SparseSymbolic sym;
  int* dev = nullptr;  // one allocation holding all index arrays
  sfb::SpPattern pat{};
};

namespace {

int fail(sfb_context* h, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->last_error = buf; else g_create_error = buf;
  return code;
}

#define SFB_CUDA(h, call)                                                                              \
  do {                                                                                                 \
    cudaError_t e__ = (call);                                                                          \
    if (e__ != cudaSuccess)                                                                            \
      return fail(h, SFB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__,  \
                  __LINE__);                                                                           \
  } while (0)

// 0 = host, 1 = device, -1 = unknown/error
int mem_space(const void* p)
{
  cudaPointerAttributes at{};
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return 0;  // plain malloc'ed memory on older runtimes
  }
  if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return 1;
  return 0;
}

// classify a set of pointers (nullptr entries ignored): 0 all host, 1 all device, -1 mixed
int classify(std::initializer_list<const void*> ps)
{
  int seen = -2;
  for (const void* p : ps) {
    if (!p) continue;
    const int s = mem_space(p);
    if (seen == -2) seen = s;
    else if (seen != s) return -1;
  }
  return seen == -2 ? 1 : seen;
}

unsigned long long* next_counter(sfb_context* h)
{
  unsigned long long* c = h->counters + h->next_counter;
  h->next_counter = (h->next_counter + 1) % h->num_counters;
  return c;
}

int ensure_slot(sfb_context* h, Slot& s, size_t bytes)
{
  if (s.bytes >= bytes) return SFB_OK;
  if (s.dev) {
    SFB_CUDA(h, cudaStreamSynchronize(s.stream));
    SFB_CUDA(h, cudaFree(s.dev));
    s.dev = nullptr;
    s.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&s.dev, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, SFB_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) for staging failed: %s", bytes, cudaGetErrorString(e));
  }
  s.bytes = bytes;
  return SFB_OK;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// -------------------------------------------------------------------------------------------------------
// dense QP launch
// -------------------------------------------------------------------------------------------------------
struct QpGeom
{
  int G = 0;            // warps cooperating on one instance (= warps per CTA)
  int ctas_per_sm = 0;
  size_t smem_per_cta = 0;
};

template <typename T, int G> constexpr int qp_minb() { return G == 4 ? 3 : (G == 2 ? 8 : 16); }

template <typename T, int G> int qp_occupancy(sfb_context* h, int n, int m, QpGeom* g)
{
  sfb::QpLayout L(n, m, 32 * G);
  const size_t bytes = (size_t)L.total * sizeof(T);
  if (bytes > h->prop.sharedMemPerBlockOptin) return 0;
  auto kern = sfb::qp_dense_group_kernel<T, G, qp_minb<T, G>(), 0, 0>;
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int nb = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * G, bytes) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  g->G = G;
  g->ctas_per_sm = nb;
  g->smem_per_cta = bytes;
  return nb * G;  // resident warps per SM
}

// Smallest group size that still puts >= 16 warps on an SM; otherwise the one with the most resident warps.
template <typename T> int qp_geometry(sfb_context* h, int n, int m, QpGeom* g)
{
  QpGeom c1, c2, c4;
  const int w1 = qp_occupancy<T, 1>(h, n, m, &c1);
  const int w2 = qp_occupancy<T, 2>(h, n, m, &c2);
  const int w4 = qp_occupancy<T, 4>(h, n, m, &c4);
  if (w1 == 0 && w2 == 0 && w4 == 0) return SFB_ERR_UNSUPPORTED_SIZE;
  if (w1 >= 16) *g = c1;
  else if (w2 >= 16) *g = c2;
  else if (w4 >= w2 && w4 >= w1) *g = c4;
  else if (w2 >= w1) *g = c2;
  else *g = c1;
  return SFB_OK;
}

int ensure_scratch(sfb_context* h, Scratch& s, size_t bytes, cudaStream_t st)
{
  if (s.bytes >= bytes) return SFB_OK;
  if (s.dev) {
    SFB_CUDA(h, cudaStreamSynchronize(st));
    SFB_CUDA(h, cudaFree(s.dev));
    s.dev = nullptr;
    s.bytes = 0;
  }
  cudaError_t e = cudaMalloc(&s.dev, bytes);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(h, SFB_ERR_OUT_OF_MEMORY, "cudaMalloc(%zu) for the polish workspace failed: %s", bytes, cudaGetErrorString(e));
  }
  s.bytes = bytes;
  return SFB_OK;
}

template <typename T, int G, int NS, int MS>
int qp_launch_g(sfb_context* h, cudaStream_t st, sfb::QpArgs<T>& args, const QpGeom& g, int grid)
{
  auto kern = sfb::qp_dense_group_kernel<T, G, qp_minb<T, G>(), NS, MS>;
  SFB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_per_cta));
  kern<<<grid, 32 * G, g.smem_per_cta, st>>>(args);
  SFB_CUDA(h, cudaGetLastError());
  return SFB_OK;
}

// tall-skinny problems without polish (the ASIF shape): warp per instance, working set in registers
template <typename T, int N, int R> int qp_launch_skinny_nr(sfb_context* h, cudaStream_t st, sfb::QpArgs<T>& args)
{
  auto kern = sfb::qp_dense_skinny_kernel<T, N, R>;
  int nb = 0;
  SFB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * sfb::kSkinnyWarps, 0));
  if (nb < 1) return fail(h, SFB_ERR_CUDA, "skinny QP kernel cannot be resident");
  const long long ctas = (args.batch + sfb::kSkinnyWarps - 1) / sfb::kSkinnyWarps;
  const int grid = (int)std::max<long long>(1, std::min<long long>(ctas, (long long)h->prop.multiProcessorCount * nb));
  kern<<<grid, 32 * sfb::kSkinnyWarps, 0, st>>>(args);
  SFB_CUDA(h, cudaGetLastError());
  return SFB_OK;
}
template <typename T, int N> int qp_launch_skinny_n(sfb_context* h, cudaStream_t st, sfb::QpArgs<T>& args)
{
  return args.m <= 128 ? qp_launch_skinny_nr<T, N, 4>(h, st, args) : qp_launch_skinny_nr<T, N, 8>(h, st, args);
}
bool qp_is_skinny(const sfb_context* h, int n, int m, int mode, const sfb_qp_params& prm)
{
  return !h->dense_force_generic && mode == 0 && !prm.polish && n <= sfb::kSkinnyMaxN && m >= 1 && m <= sfb::kSkinnyMaxM;
}

template <typename T>
int qp_launch(sfb_context* h, cudaStream_t st, int scratch_slot, const sfb::QpArgs<T>& args_in)
{
  sfb::QpArgs<T> args = args_in;
  if (qp_is_skinny(h, args.n, args.m, args.mode, args.prm)) {
    args.scratch = nullptr;
    args.scratch_per_cta = 0;
    args.work_counter = next_counter(h);
    SFB_CUDA(h, cudaMemsetAsync(args.work_counter, 0, sizeof(unsigned long long), st));
    int rc;
    switch (args.n) {
      case 1: rc = qp_launch_skinny_n<T, 1>(h, st, args); break;
      case 2: rc = qp_launch_skinny_n<T, 2>(h, st, args); break;
      case 3: rc = qp_launch_skinny_n<T, 3>(h, st, args); break;
      default: rc = qp_launch_skinny_n<T, 4>(h, st, args); break;
    }
    if (rc != SFB_OK) return rc;
    h->launches += 1;
    return SFB_OK;
  }
  QpGeom g;
  if (qp_geometry<T>(h, args.n, args.m, &g) != SFB_OK)
    return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "dense QP n=%d m=%d (%zu-byte scalars) does not fit in shared memory",
                args.n, args.m, sizeof(T));
  const long long persistent = (long long)h->prop.multiProcessorCount * g.ctas_per_sm;
  const int grid = (int)std::max<long long>(1, std::min<long long>(args.batch, persistent));
  // polish workspace: the Schur block S (na x na, na <= min(n, m)) lives in shared memory when 2 na <= ldA
  args.scratch = nullptr;
  args.scratch_per_cta = 0;
  if (args.mode == 0 && args.prm.polish) {
    const long long k = std::min(args.n, args.m);
    const sfb::QpLayout L(args.n, args.m, 32 * g.G);
    if (2 * k > L.ldA) {
      const size_t bytes = sizeof(T) * (size_t)(k * k) * (size_t)grid;
      int rc = ensure_scratch(h, h->scratch[scratch_slot], bytes, st);
      if (rc != SFB_OK) return rc;
      args.scratch = static_cast<T*>(h->scratch[scratch_slot].dev);
      args.scratch_per_cta = k * k;
    }
  }
  args.work_counter = next_counter(h);
  SFB_CUDA(h, cudaMemsetAsync(args.work_counter, 0, sizeof(unsigned long long), st));
  int rc;
  // shape-specialised instantiations (compile-time n, m) for the headline shapes, generic kernels otherwise
  if (g.G == 4) {
    if (std::is_same<T, double>::value && args.n == 50 && args.m == 100) rc = qp_launch_g<T, 4, 50, 100>(h, st, args, g, grid);
    else rc = qp_launch_g<T, 4, 0, 0>(h, st, args, g, grid);
  } else if (g.G == 2) {
    rc = qp_launch_g<T, 2, 0, 0>(h, st, args, g, grid);
  } else {
    if (std::is_same<T, double>::value && args.n == 10 && args.m == 20) rc = qp_launch_g<T, 1, 10, 20>(h, st, args, g, grid);
    else rc = qp_launch_g<T, 1, 0, 0>(h, st, args, g, grid);
  }
  if (rc != SFB_OK) return rc;
  h->launches += 1;
  return SFB_OK;
}

int check_params(sfb_context* h, const sfb_qp_params* prm, int64_t batch, int n, int m)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!prm) return fail(h, SFB_ERR_INVALID_ARGUMENT, "prm is NULL");
  if (batch < 0 || n <= 0 || m < 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad sizes batch=%lld n=%d m=%d", (long long)batch, n, m);
  if (prm->stop_check_iter == 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "stop_check_iter must be > 0");
  return SFB_OK;
}

template <typename T>
int qp_solve_impl(sfb_context* h, const sfb_qp_params* prm, int64_t batch, int n, int m, const T* P, const T* q,
                  const T* A, const T* l, const T* u, const T* warm_x, const T* warm_y, T* out_x, T* out_y,
                  T* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active, uint32_t* out_flags)
{
  int rc = check_params(h, prm, batch, n, m);
  if (rc != SFB_OK) return rc;
  if (!P || !q || !out_x || !out_y || !out_obj || !out_status || !out_iter || (m > 0 && (!A || !l || !u)))
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  if ((warm_x == nullptr) != (warm_y == nullptr))
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "warm_x and warm_y must both be given or both be NULL");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  {
    QpGeom g;
    if (qp_geometry<T>(h, n, m, &g) != SFB_OK)
      return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "dense QP n=%d m=%d (%zu-byte scalars) does not fit in shared memory", n,
                  m, sizeof(T));
  }
  const int space = classify({P, q, A, l, u, warm_x, warm_y, out_x, out_y, out_obj, out_status, out_iter, out_active, out_flags});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");

  sfb::QpArgs<T> a{};
  a.batch = batch;
  a.n = n;
  a.m = m;
  a.mode = 0;
  a.prm = *prm;
  a.max_iter_eff = prm->has_max_iter ? prm->max_iter : SFB_QP_DEVICE_ITER_CAP;

  if (space == 1) {
    a.P = P; a.q = q; a.A = A; a.l = l; a.u = u; a.warm_x = warm_x; a.warm_y = warm_y;
    a.out_x = out_x; a.out_y = out_y; a.out_obj = out_obj; a.out_status = out_status; a.out_iter = out_iter;
    a.out_active = out_active; a.out_flags = out_flags;
    return qp_launch<T>(h, h->stream, kNumSlots, a);
  }

  // ---- host buffers: pipelined staging (chunk k uses slot k % kNumSlots on its own stream) ----
  const size_t sP = align_up(sizeof(T) * n * n, 16), sq = align_up(sizeof(T) * n, 16),
               sA = align_up(sizeof(T) * m * n, 16), sm_ = align_up(sizeof(T) * m, 16);
  (void)sP; (void)sq; (void)sA; (void)sm_;
  const size_t in_per = sizeof(T) * ((size_t)n * n + n + (size_t)m * n + 2 * (size_t)m) +
                        (warm_x ? sizeof(T) * ((size_t)n + m) : 0);
  const size_t out_per = sizeof(T) * ((size_t)n + m + 1) + 8 + (size_t)m + 4;
  const size_t per = in_per + out_per;
  // chunk: ~64 MB of staging per slot, at least one instance, and at least ~4 chunks when the batch allows
  long long chunk = std::max<long long>(1, (long long)((64ull << 20) / per));
  chunk = std::min<long long>(chunk, std::max<long long>(1, (batch + 3) / 4));
  chunk = std::min<long long>(chunk, batch);
  // every per-field sub-buffer 256-byte aligned
  auto fld = [&](size_t elems_bytes) { return align_up(elems_bytes * (size_t)chunk, 256); };
  const size_t oP = 0, oq = oP + fld(sizeof(T) * n * n), oA = oq + fld(sizeof(T) * n), ol = oA + fld(sizeof(T) * m * n),
               ou = ol + fld(sizeof(T) * m), owx = ou + fld(sizeof(T) * m), owy = owx + fld(sizeof(T) * n),
               oox = owy + fld(sizeof(T) * m), ooy = oox + fld(sizeof(T) * n), oobj = ooy + fld(sizeof(T) * m),
               ost = oobj + fld(sizeof(T)), oit = ost + fld(4), oact = oit + fld(4), ofl = oact + fld(m),
               slot_bytes = ofl + fld(4);

  SFB_CUDA(h, cudaEventRecord(h->ev_start, h->stream));
  for (int s = 0; s < kNumSlots; ++s) SFB_CUDA(h, cudaStreamWaitEvent(h->slots[s].stream, h->ev_start, 0));
  int k = 0;
  for (long long b0 = 0; b0 < batch; b0 += chunk, ++k) {
    Slot& s = h->slots[k % kNumSlots];
    rc = ensure_slot(h, s, slot_bytes);
    if (rc != SFB_OK) return rc;
    const long long cnt = std::min<long long>(chunk, batch - b0);
    char* d = static_cast<char*>(s.dev);
    auto h2d = [&](size_t off, const T* src, size_t per_inst) -> cudaError_t {
      return cudaMemcpyAsync(d + off, src + (size_t)b0 * per_inst, sizeof(T) * per_inst * (size_t)cnt,
                             cudaMemcpyHostToDevice, s.stream);
    };
    SFB_CUDA(h, h2d(oP, P, (size_t)n * n));
    SFB_CUDA(h, h2d(oq, q, n));
    if (m > 0) {
      SFB_CUDA(h, h2d(oA, A, (size_t)m * n));
      SFB_CUDA(h, h2d(ol, l, m));
      SFB_CUDA(h, h2d(ou, u, m));
    }
    if (warm_x) {
      SFB_CUDA(h, h2d(owx, warm_x, n));
      if (m > 0) SFB_CUDA(h, h2d(owy, warm_y, m));
    }
    sfb::QpArgs<T> c = a;
    c.batch = cnt;
    c.P = reinterpret_cast<const T*>(d + oP); c.q = reinterpret_cast<const T*>(d + oq);
    c.A = reinterpret_cast<const T*>(d + oA); c.l = reinterpret_cast<const T*>(d + ol);
    c.u = reinterpret_cast<const T*>(d + ou);
    c.warm_x = warm_x ? reinterpret_cast<const T*>(d + owx) : nullptr;
    c.warm_y = warm_x ? reinterpret_cast<const T*>(d + owy) : nullptr;
    c.out_x = reinterpret_cast<T*>(d + oox); c.out_y = reinterpret_cast<T*>(d + ooy);
    c.out_obj = reinterpret_cast<T*>(d + oobj); c.out_status = reinterpret_cast<int32_t*>(d + ost);
    c.out_iter = reinterpret_cast<uint32_t*>(d + oit);
    c.out_active = out_active ? reinterpret_cast<int8_t*>(d + oact) : nullptr;
    c.out_flags = out_flags ? reinterpret_cast<uint32_t*>(d + ofl) : nullptr;
    rc = qp_launch<T>(h, s.stream, k % kNumSlots, c);
    if (rc != SFB_OK) return rc;
    auto d2h = [&](void* dst, size_t off, size_t bytes_per_inst) -> cudaError_t {
      return cudaMemcpyAsync(static_cast<char*>(dst) + (size_t)b0 * bytes_per_inst, d + off,
                             bytes_per_inst * (size_t)cnt, cudaMemcpyDeviceToHost, s.stream);
    };
    SFB_CUDA(h, d2h(out_x, oox, sizeof(T) * n));
    if (m > 0) SFB_CUDA(h, d2h(out_y, ooy, sizeof(T) * m));
    SFB_CUDA(h, d2h(out_obj, oobj, sizeof(T)));
    SFB_CUDA(h, d2h(out_status, ost, 4));
    SFB_CUDA(h, d2h(out_iter, oit, 4));
    if (out_active && m > 0) SFB_CUDA(h, d2h(out_active, oact, m));
    if (out_flags) SFB_CUDA(h, d2h(out_flags, ofl, 4));
  }
  for (int s = 0; s < kNumSlots; ++s) {
    SFB_CUDA(h, cudaEventRecord(h->slots[s].done, h->slots[s].stream));
    SFB_CUDA(h, cudaStreamWaitEvent(h->stream, h->slots[s].done, 0));
  }
  // host results must be in place when the call returns
  for (int s = 0; s < kNumSlots; ++s) SFB_CUDA(h, cudaStreamSynchronize(h->slots[s].stream));
  return SFB_OK;
}

// -------------------------------------------------------------------------------------------------------
// EKF launches
// -------------------------------------------------------------------------------------------------------
int ekf_block_threads(sfb_context* h, size_t elems_per_thread, size_t scalar, size_t* smem)
{
  const size_t cap = h->prop.sharedMemPerBlockOptin;
  for (int bd = 128; bd >= 32; bd -= 32) {
    // target >= 2 CTAs per SM when possible
    const size_t need = elems_per_thread * (size_t)(bd + 1) * scalar;
    const size_t budget = (bd > 32) ? cap / 2 : cap;
    if (need <= budget) { *smem = need; return bd; }
  }
  return 0;
}

// ---- fused, size-specialised TMA path (ekf_fused_tma.cuh) -----------------------------------------------
constexpr int kEkfTile = 64;

bool aligned16(std::initializer_list<const void*> ps)
{
  for (const void* p : ps)
    if (p && (reinterpret_cast<uintptr_t>(p) & 15u)) return false;
  return true;
}

template <int D, int NY, bool PRED, bool UPD>
int ekf_fused_launch(sfb_context* h, const sfb::EkfStepArgs& a)
{
  using L = sfb::EkfFusedLayout<D, NY, PRED, UPD, kEkfTile>;
  auto kern = sfb::ekf_fused_tma_kernel<D, NY, PRED, UPD, kEkfTile>;
  SFB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  int nb = 0;
  SFB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kEkfTile, L::bytes));
  if (nb < 1) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "fused EKF kernel does not fit on this device");
  const long long tiles = (a.batch + kEkfTile - 1) / kEkfTile;
  const int grid = (int)std::min<long long>(tiles, (long long)h->prop.multiProcessorCount * nb);
  kern<<<grid, kEkfTile, L::bytes, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return SFB_OK;
}

// returns -1 when no specialisation exists for (d, ny, mode): the caller takes the generic kernels
template <bool PRED, bool UPD> int ekf_fused_dispatch(sfb_context* h, int d, int ny, const sfb::EkfStepArgs& a)
{
  if (h->ekf_force_generic) return -1;  // SFB_EKF_FORCE_GENERIC=1: A/B measurements against the generic kernels
  if (!UPD) {
    if (d == 2) return ekf_fused_launch<2, 1, PRED, false>(h, a);
    if (d == 3) return ekf_fused_launch<3, 1, PRED, false>(h, a);
    if (d == 4) return ekf_fused_launch<4, 1, PRED, false>(h, a);
    if (d == 6) return ekf_fused_launch<6, 1, PRED, false>(h, a);
    return -1;
  }
  // (state dof, measurement dim) pairs with a register-resident specialisation; everything else takes the generic kernels
  if (d == 6 && ny == 3) return ekf_fused_launch<6, 3, PRED, true>(h, a);
  if (d == 6 && ny == 6) return ekf_fused_launch<6, 6, PRED, true>(h, a);
  if (d == 6 && ny == 2) return ekf_fused_launch<6, 2, PRED, true>(h, a);
  if (d == 6 && ny == 1) return ekf_fused_launch<6, 1, PRED, true>(h, a);
  if (d == 4 && ny == 2) return ekf_fused_launch<4, 2, PRED, true>(h, a);
  if (d == 3 && ny == 3) return ekf_fused_launch<3, 3, PRED, true>(h, a);
  if (d == 3 && ny == 1) return ekf_fused_launch<3, 1, PRED, true>(h, a);
  if (d == 2 && ny == 2) return ekf_fused_launch<2, 2, PRED, true>(h, a);
  return -1;
}

int ekf_predict_generic(sfb_context* h, int64_t batch, int d, int stepper, const double* P, const double* A,
                        const double* Q, double tau, double dt, double* out_P)
{
  size_t smem = 0;
  const int bd = ekf_block_threads(h, (size_t)6 * d * d, sizeof(double), &smem);
  if (bd == 0) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "EKF predict d=%d does not fit in shared memory", d);
  auto kern = sfb::ekf_predict_kernel<double>;
  SFB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sfb::EkfPredictArgs<double> a{P, A, Q, out_P, batch, d, stepper, tau, dt};
  const long long tiles = (batch + bd - 1) / bd;
  const int per_sm = (int)std::max<size_t>(1, h->prop.sharedMemPerMultiprocessor / (smem + 1024));
  const int grid = (int)std::min<long long>(tiles, (long long)h->prop.multiProcessorCount * per_sm);
  kern<<<grid, bd, smem, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return SFB_OK;
}

int ekf_update_generic(sfb_context* h, int64_t batch, int d, int ny, const double* P, const double* H,
                       const double* R, const double* innov, double* out_delta, double* out_P)
{
  size_t smem = 0;
  const int bd = ekf_block_threads(h, sfb::ekf_update_elems<double>(d, ny), sizeof(double), &smem);
  if (bd == 0) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "EKF update d=%d ny=%d does not fit in shared memory", d, ny);
  auto kern = sfb::ekf_update_kernel<double>;
  SFB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sfb::EkfUpdateArgs<double> a{P, H, R, innov, out_delta, out_P, batch, d, ny};
  const long long tiles = (batch + bd - 1) / bd;
  const int per_sm = (int)std::max<size_t>(1, h->prop.sharedMemPerMultiprocessor / (smem + 1024));
  const int grid = (int)std::min<long long>(tiles, (long long)h->prop.multiProcessorCount * per_sm);
  kern<<<grid, bd, smem, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return SFB_OK;
}

// ---- sparse QP (shared pattern) --------------------------------------------------------------------------
template <typename T>
int qp_sparse_solve_impl(sfb_context* h, const sfb_qp_sparse_pattern* pt, const sfb_qp_params* prm, int64_t batch,
                         const T* P, const T* q, const T* A, const T* l, const T* u, const T* warm_x, const T* warm_y,
                         T* out_x, T* out_y, T* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active,
                         uint32_t* out_flags)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!pt) return fail(h, SFB_ERR_INVALID_ARGUMENT, "pattern is NULL");
  const int n = pt->sym.n, m = pt->sym.m;
  int rc = check_params(h, prm, batch, n, m);
  if (rc != SFB_OK) return rc;
  if (pt->device != h->device) return fail(h, SFB_ERR_INVALID_ARGUMENT, "pattern was analysed for device %d, handle is on %d", pt->device, h->device);
  if (!q || !out_x || !out_y || !out_obj || !out_status || !out_iter || (pt->sym.nnzP > 0 && !P) ||
      (m > 0 && (!l || !u)) || (pt->sym.nnzA > 0 && !A))
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  if ((warm_x == nullptr) != (warm_y == nullptr))
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "warm_x and warm_y must both be given or both be NULL");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  const int space = classify({P, q, A, l, u, warm_x, warm_y, out_x, out_y, out_obj, out_status, out_iter, out_active, out_flags});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");

  const sfb::SparseSymbolic& S = pt->sym;
  // tile width.  4 instances per warp with 8 lanes cooperating on each wins at every batch size measured (n = m = 422:
  // batch 8192 -> 99k solves/s fp64 / 186k fp32 against 14k / - with one lane per instance; batch 65536 -> 105k / 235k
  // against 100k / 150k; profiles/README.md); one lane per instance (32 per warp) needs ~40 % less workspace and is kept
  // for batches whose 4-wide working set would not fit in half of the free device memory.
  int tw = 4;
  {
    const size_t per_inst4 = (sfb::sp_a_len(pt->pat, 4) + pt->sym.nnzP + sfb::sp_w_len(pt->pat, 4) + (size_t)sfb::kSpNV * n +
                              (size_t)sfb::kSpMV * m) * sizeof(T);
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
    if (h->sparse_ws.bytes < per_inst4 * (size_t)batch && per_inst4 * (size_t)batch > free_b / 2) tw = 32;
  }
  if (h->sparse_tw == 4 || h->sparse_tw == 8 || h->sparse_tw == 32) tw = h->sparse_tw;
  const long long tiles = (batch + tw - 1) / tw;
  const size_t wlen = sfb::sp_w_len(pt->pat, tw);  // factor + its stream-ordered copies
  const size_t alen = sfb::sp_a_len(pt->pat, tw);  // Abar + its padded row / column stream copies
  const size_t per_tile = (alen + S.nnzP + wlen + (size_t)sfb::kSpNV * n + (size_t)sfb::kSpMV * m) * tw * sizeof(T);
  rc = ensure_scratch(h, h->sparse_ws, per_tile * (size_t)tiles, h->stream);
  if (rc != SFB_OK) return rc;

  sfb::SpArgs<T> a{};
  a.pat = pt->pat;
  a.batch = batch;
  a.prm = *prm;
  a.max_iter_eff = prm->has_max_iter ? prm->max_iter : SFB_QP_DEVICE_ITER_CAP;
  {
    T* w = static_cast<T*>(h->sparse_ws.dev);
    a.wsA = w; w += (size_t)tiles * alen * tw;
    a.wsP = w; w += (size_t)tiles * S.nnzP * tw;
    a.wsW = w; w += (size_t)tiles * wlen * tw;
    a.wsN = w; w += (size_t)tiles * sfb::kSpNV * n * tw;
    a.wsM = w;
  }
  auto launch = [&]() -> int {
    const unsigned grid = (unsigned)std::min<long long>(tiles, 1 << 30);
    if (tw < 32) {
      const size_t smem = (size_t)(n + 1) * tw * sizeof(T);  // the solve vector of the tile + the dummy zero slot
      if (smem > h->prop.sharedMemPerBlockOptin) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "sparse QP n=%d: solve vector does not fit in shared memory", n);
      if (tw == 8) {
        SFB_CUDA(h, cudaFuncSetAttribute(sfb::qp_sparse_tiled_kernel<T, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sfb::qp_sparse_tiled_kernel<T, 8><<<grid, 32, smem, h->stream>>>(a);
      } else {
        SFB_CUDA(h, cudaFuncSetAttribute(sfb::qp_sparse_tiled_kernel<T, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        sfb::qp_sparse_tiled_kernel<T, 4><<<grid, 32, smem, h->stream>>>(a);
      }
    } else {
      sfb::qp_sparse_tiled_kernel<T, 32><<<grid, 32, 0, h->stream>>>(a);
    }
    SFB_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return SFB_OK;
  };
  if (space == 1) {
    a.P = P; a.q = q; a.A = A; a.l = l; a.u = u; a.warm_x = warm_x; a.warm_y = warm_y;
    a.out_x = out_x; a.out_y = out_y; a.out_obj = out_obj; a.out_status = out_status; a.out_iter = out_iter;
    a.out_active = out_active; a.out_flags = out_flags;
    return launch();
  }
  // host buffers: one staged round trip on the handle's stream (inputs are ~1 % of the per-solve traffic of this path)
  auto al = [](size_t v) { return (v + 255) / 256 * 256; };
  const size_t B = (size_t)batch;
  const size_t sP = al(sizeof(T) * S.nnzP * B), sq = al(sizeof(T) * n * B), sA = al(sizeof(T) * S.nnzA * B),
               sm_ = al(sizeof(T) * m * B), s4 = al(4 * B), sact = al((size_t)m * B), s1 = al(sizeof(T) * B);
  const size_t total = sP + sq + sA + 2 * sm_ + (warm_x ? sq + sm_ : 0) + sq + sm_ + s1 + 3 * s4 + sact;
  rc = ensure_scratch(h, h->sparse_stage, total, h->stream);
  if (rc != SFB_OK) return rc;
  char* d = static_cast<char*>(h->sparse_stage.dev);
  auto take = [&](size_t bytes) { char* r = d; d += bytes; return r; };
  T* dP = (T*)take(sP); T* dq = (T*)take(sq); T* dA = (T*)take(sA); T* dl = (T*)take(sm_); T* du = (T*)take(sm_);
  T* dwx = warm_x ? (T*)take(sq) : nullptr; T* dwy = warm_x ? (T*)take(sm_) : nullptr;
  T* dox = (T*)take(sq); T* doy = (T*)take(sm_); T* dobj = (T*)take(s1);
  int32_t* dst = (int32_t*)take(s4); uint32_t* dit = (uint32_t*)take(s4); uint32_t* dfl = (uint32_t*)take(s4);
  int8_t* dact = (int8_t*)take(sact);
  auto up = [&](void* dst_, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst_, src, bytes, cudaMemcpyHostToDevice, h->stream) : cudaSuccess; };
  SFB_CUDA(h, up(dP, P, sizeof(T) * S.nnzP * B));
  SFB_CUDA(h, up(dq, q, sizeof(T) * n * B));
  SFB_CUDA(h, up(dA, A, sizeof(T) * S.nnzA * B));
  SFB_CUDA(h, up(dl, l, sizeof(T) * m * B));
  SFB_CUDA(h, up(du, u, sizeof(T) * m * B));
  if (warm_x) {
    SFB_CUDA(h, up(dwx, warm_x, sizeof(T) * n * B));
    SFB_CUDA(h, up(dwy, warm_y, sizeof(T) * m * B));
  }
  a.P = dP; a.q = dq; a.A = dA; a.l = dl; a.u = du; a.warm_x = dwx; a.warm_y = dwy;
  a.out_x = dox; a.out_y = doy; a.out_obj = dobj; a.out_status = dst; a.out_iter = dit;
  a.out_active = out_active ? dact : nullptr; a.out_flags = out_flags ? dfl : nullptr;
  rc = launch();
  if (rc != SFB_OK) return rc;
  auto down = [&](void* dst_, const void* src, size_t bytes) { return bytes ? cudaMemcpyAsync(dst_, src, bytes, cudaMemcpyDeviceToHost, h->stream) : cudaSuccess; };
  SFB_CUDA(h, down(out_x, dox, sizeof(T) * n * B));
  SFB_CUDA(h, down(out_y, doy, sizeof(T) * m * B));
  SFB_CUDA(h, down(out_obj, dobj, sizeof(T) * B));
  SFB_CUDA(h, down(out_status, dst, 4 * B));
  SFB_CUDA(h, down(out_iter, dit, 4 * B));
  if (out_active) SFB_CUDA(h, down(out_active, dact, (size_t)m * B));
  if (out_flags) SFB_CUDA(h, down(out_flags, dfl, 4 * B));
  SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SFB_OK;
}

// device staging of host buffers for the EKF entry points (doubles; every block 256-byte aligned)
struct EkfStage
{
  sfb_context* h;
  char* base = nullptr;
  size_t off = 0;
  explicit EkfStage(sfb_context* h_) : h(h_) {}
  static size_t al(size_t elems) { return (elems * sizeof(double) + 255) / 256 * 256; }
  int reserve(size_t bytes)
  {
    const int rc = ensure_scratch(h, h->sparse_stage, bytes, h->stream);
    if (rc == SFB_OK) base = static_cast<char*>(h->sparse_stage.dev);
    return rc;
  }
  double* out(size_t elems)
  {
    double* p = reinterpret_cast<double*>(base + off);
    off += al(elems);
    return p;
  }
  cudaError_t up(const double** dev, const double* host, size_t elems)
  {
    double* p = out(elems);
    *dev = p;
    return cudaMemcpyAsync(p, host, elems * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  }
  cudaError_t down(double* host, const double* dev, size_t elems)
  {
    return cudaMemcpyAsync(host, dev, elems * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  }
};

}  // namespace

// =======================================================================================================
// C ABI
// =======================================================================================================
extern "C" {

int sfb_version(void) { return SFB_VERSION; }

const char* sfb_error_string(int err)
{
  switch (err) {
    case SFB_OK: return "ok";
    case SFB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SFB_ERR_NO_DEVICE: return "no usable CUDA device (this engine has no CPU path)";
    case SFB_ERR_CUDA: return "CUDA runtime error";
    case SFB_ERR_UNSUPPORTED_SIZE: return "problem size not supported by the shared-memory resident kernels";
    case SFB_ERR_MIXED_MEMORY: return "host and device pointers mixed in one call";
    case SFB_ERR_OUT_OF_MEMORY: return "out of device memory";
    default: return "unknown error";
  }
}

const char* sfb_last_error_message(sfb_handle_t h) { return h ? h->last_error.c_str() : g_create_error.c_str(); }

int sfb_create(int device, void* stream, sfb_handle_t* out)
{
  if (!out) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    cudaGetLastError();
    return fail(nullptr, SFB_ERR_NO_DEVICE, "no CUDA device visible (%s); libsfb has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  }
  if (device < 0 || device >= count) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, count);
  sfb_context* h = new sfb_context();
  h->device = device;
  { const char* e = getenv("SFB_EKF_FORCE_GENERIC"); h->ekf_force_generic = e && e[0] == '1'; }
  { const char* e = getenv("SFB_SPARSE_TW"); h->sparse_tw = e ? atoi(e) : 0; }
  { const char* e = getenv("SFB_DENSE_FORCE_GENERIC"); h->dense_force_generic = e && e[0] == '1'; }
  h->stream = static_cast<cudaStream_t>(stream);
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&h->prop, device) != cudaSuccess) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    delete h;
    return fail(nullptr, SFB_ERR_CUDA, "cannot open device %d: %s", device, msg);
  }
  if (h->prop.major != 10) {
    const int mj = h->prop.major, mn = h->prop.minor;
    delete h;
    return fail(nullptr, SFB_ERR_NO_DEVICE, "device %d is sm_%d%d; libsfb is built for sm_100a only", device, mj, mn);
  }
  bool ok = cudaMalloc(&h->counters, sizeof(unsigned long long) * h->num_counters) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming) == cudaSuccess;
  for (int s = 0; ok && s < kNumSlots; ++s) {
    ok = ok && cudaStreamCreateWithFlags(&h->slots[s].stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&h->slots[s].done, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    sfb_destroy(h);
    return fail(nullptr, SFB_ERR_CUDA, "handle setup failed: %s", msg);
  }
  *out = h;
  return SFB_OK;
}

int sfb_destroy(sfb_handle_t h)
{
  if (!h) return SFB_OK;
  cudaSetDevice(h->device);
  for (int s = 0; s < kNumSlots; ++s) {
    if (h->slots[s].stream) { cudaStreamSynchronize(h->slots[s].stream); cudaStreamDestroy(h->slots[s].stream); }
    if (h->slots[s].done) cudaEventDestroy(h->slots[s].done);
    if (h->slots[s].dev) cudaFree(h->slots[s].dev);
  }
  if (h->ev_start) cudaEventDestroy(h->ev_start);
  for (auto& sc : h->scratch)
    if (sc.dev) cudaFree(sc.dev);
  if (h->sparse_ws.dev) cudaFree(h->sparse_ws.dev);
  if (h->sparse_stage.dev) cudaFree(h->sparse_stage.dev);
  if (h->counters) cudaFree(h->counters);
  delete h;
  return SFB_OK;
}

int sfb_set_stream(sfb_handle_t h, void* stream)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  h->stream = static_cast<cudaStream_t>(stream);
  return SFB_OK;
}

int sfb_synchronize(sfb_handle_t h)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  SFB_CUDA(h, cudaSetDevice(h->device));
  SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SFB_OK;
}

int sfb_kernel_launch_count(sfb_handle_t h, uint64_t* out)
{
  if (!h || !out) return SFB_ERR_INVALID_ARGUMENT;
  *out = h->launches;
  return SFB_OK;
}

void sfb_qp_params_default(sfb_qp_params* p)
{
  if (!p) return;
  // qp_solver.hpp:29-68
  p->verbose = 0;
  p->alpha = 1.6f;
  p->rho = 0.1f;
  p->sigma = 1e-6f;
  p->scaling = 1;
  p->eps_abs = 1e-3f;
  p->eps_rel = 1e-3f;
  p->eps_primal_inf = 1e-4f;
  p->eps_dual_inf = 1e-4f;
  p->has_max_iter = 0;
  p->max_iter = 0;
  p->has_max_time = 0;
  p->max_time_ns = 0;
  p->stop_check_iter = 25;
  p->polish = 1;
  p->polish_iter = 5;
  p->delta = 1e-6f;
}

int sfb_qp_solve_dense_batch_f64(sfb_handle_t h, const sfb_qp_params* prm, int64_t batch, int n, int m,
                                 const double* P, const double* q, const double* A, const double* l,
                                 const double* u, const double* warm_x, const double* warm_y, double* out_x,
                                 double* out_y, double* out_obj, int32_t* out_status, uint32_t* out_iter,
                                 int8_t* out_active, uint32_t* out_flags)
{
  return qp_solve_impl<double>(h, prm, batch, n, m, P, q, A, l, u, warm_x, warm_y, out_x, out_y, out_obj,
                               out_status, out_iter, out_active, out_flags);
}

int sfb_qp_solve_dense_batch_f32(sfb_handle_t h, const sfb_qp_params* prm, int64_t batch, int n, int m,
                                 const float* P, const float* q, const float* A, const float* l, const float* u,
                                 const float* warm_x, const float* warm_y, float* out_x, float* out_y,
                                 float* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active,
                                 uint32_t* out_flags)
{
  return qp_solve_impl<float>(h, prm, batch, n, m, P, q, A, l, u, warm_x, warm_y, out_x, out_y, out_obj, out_status,
                              out_iter, out_active, out_flags);
}

int sfb_qp_dense_max_m(sfb_handle_t h, int n, int scalar_bytes)
{
  if (!h || n <= 0 || (scalar_bytes != 4 && scalar_bytes != 8)) return 0;
  const size_t cap = h->prop.sharedMemPerBlockOptin;
  int lo = 0, hi = 1 << 16;
  if ((size_t)sfb::QpLayout(n, 1, 32).total * scalar_bytes > cap) return 0;
  while (lo + 1 < hi) {  // largest m that fits
    const int mid = (lo + hi) / 2;
    if ((size_t)sfb::QpLayout(n, mid, 32).total * scalar_bytes <= cap) lo = mid; else hi = mid;
  }
  return lo;
}

int sfb_qp_scale_dense_batch_f64(sfb_handle_t h, int64_t batch, int n, int m, const double* P, const double* q,
                                 const double* A, double* out_c, double* out_sx, double* out_sy)
{
  sfb_qp_params prm;
  sfb_qp_params_default(&prm);
  int rc = check_params(h, &prm, batch, n, m);
  if (rc != SFB_OK) return rc;
  if (!P || !q || (m > 0 && !A) || !out_c || !out_sx || !out_sy) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  if (classify({P, q, A, out_c, out_sx, out_sy}) != 1) return fail(h, SFB_ERR_MIXED_MEMORY, "sfb_qp_scale_dense_batch_f64 takes device pointers only");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  sfb::QpArgs<double> a{};
  a.batch = batch; a.n = n; a.m = m; a.mode = 1; a.prm = prm; a.max_iter_eff = 0;
  a.P = P; a.q = q; a.A = A; a.l = A; a.u = A;  // l/u are staged but unused in scale-only mode: point at valid memory
  a.out_c = out_c; a.out_sx = out_sx; a.out_sy = out_sy;
  return qp_launch<double>(h, h->stream, kNumSlots, a);
}

int sfb_ekf_predict_batch_f64(sfb_handle_t h, int64_t batch, int d, int stepper, const double* P,
                              const double* A, const double* Q, double tau, double dt, double* out_P)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (batch < 0 || d <= 0 || (stepper != SFB_STEPPER_EULER && stepper != SFB_STEPPER_RK4) || !P || !A || !Q || !out_P)
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_ekf_predict_batch_f64");
  const int space = classify({P, A, Q, out_P});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 0) {  // host buffers: one staged round trip on the handle's stream
    const size_t dd = (size_t)d * d * batch;
    EkfStage stg(h);
    const double *dP, *dA, *dQ;
    double* dO;
    int rc = stg.reserve(3 * stg.al(dd) + stg.al(dd));
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.up(&dP, P, dd)); SFB_CUDA(h, stg.up(&dA, A, dd)); SFB_CUDA(h, stg.up(&dQ, Q, dd));
    dO = stg.out(dd);
    rc = sfb_ekf_predict_batch_f64(h, batch, d, stepper, dP, dA, dQ, tau, dt, dO);
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.down(out_P, dO, dd));
    SFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return SFB_OK;
  }
  if (stepper == SFB_STEPPER_EULER && aligned16({P, A, Q, out_P})) {
    sfb::EkfStepArgs a{P, A, Q, nullptr, nullptr, nullptr, nullptr, out_P, batch, tau, dt};
    const int rc = ekf_fused_dispatch<true, false>(h, d, 1, a);
    if (rc >= 0) return rc;
  }
  return ekf_predict_generic(h, batch, d, stepper, P, A, Q, tau, dt, out_P);
}

int sfb_ekf_update_batch_f64(sfb_handle_t h, int64_t batch, int d, int ny, const double* P, const double* H,
                             const double* R, const double* innov, double* out_delta, double* out_P)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (batch < 0 || d <= 0 || ny <= 0 || ny > sfb::kEkfMaxNy || !P || !H || !R || !innov || !out_delta || !out_P)
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_ekf_update_batch_f64");
  const int space = classify({P, H, R, innov, out_delta, out_P});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 0) {
    const size_t B = (size_t)batch, dd = (size_t)d * d * B, nd = (size_t)ny * d * B, nn = (size_t)ny * ny * B;
    EkfStage stg(h);
    const double *dP, *dH, *dR, *dI;
    int rc = stg.reserve(stg.al(dd) + stg.al(nd) + stg.al(nn) + stg.al(ny * B) + stg.al(d * B) + stg.al(dd));
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.up(&dP, P, dd)); SFB_CUDA(h, stg.up(&dH, H, nd)); SFB_CUDA(h, stg.up(&dR, R, nn)); SFB_CUDA(h, stg.up(&dI, innov, ny * B));
    double* dD = stg.out(d * B);
    double* dO = stg.out(dd);
    rc = sfb_ekf_update_batch_f64(h, batch, d, ny, dP, dH, dR, dI, dD, dO);
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.down(out_delta, dD, d * B)); SFB_CUDA(h, stg.down(out_P, dO, dd));
    SFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return SFB_OK;
  }
  if (aligned16({P, H, R, innov, out_delta, out_P})) {
    sfb::EkfStepArgs a{P, nullptr, nullptr, H, R, innov, out_delta, out_P, batch, 0.0, 0.0};
    const int rc = ekf_fused_dispatch<false, true>(h, d, ny, a);
    if (rc >= 0) return rc;
  }
  return ekf_update_generic(h, batch, d, ny, P, H, R, innov, out_delta, out_P);
}

int sfb_ekf_step_batch_f64(sfb_handle_t h, int64_t batch, int d, int ny, int stepper, const double* P,
                           const double* A, const double* Q, double tau, double dt, const double* H,
                           const double* R, const double* innov, double* out_delta, double* out_P)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (batch < 0 || d <= 0 || ny <= 0 || ny > sfb::kEkfMaxNy || (stepper != SFB_STEPPER_EULER && stepper != SFB_STEPPER_RK4) ||
      !P || !A || !Q || !H || !R || !innov || !out_delta || !out_P)
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_ekf_step_batch_f64");
  const int space = classify({P, A, Q, H, R, innov, out_delta, out_P});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  if (batch == 0) return SFB_OK;
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 0) {
    const size_t B = (size_t)batch, dd = (size_t)d * d * B, nd = (size_t)ny * d * B, nn = (size_t)ny * ny * B;
    EkfStage stg(h);
    const double *dP, *dA, *dQ, *dH, *dR, *dI;
    int rc = stg.reserve(3 * stg.al(dd) + stg.al(nd) + stg.al(nn) + stg.al(ny * B) + stg.al(d * B) + stg.al(dd));
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.up(&dP, P, dd)); SFB_CUDA(h, stg.up(&dA, A, dd)); SFB_CUDA(h, stg.up(&dQ, Q, dd));
    SFB_CUDA(h, stg.up(&dH, H, nd)); SFB_CUDA(h, stg.up(&dR, R, nn)); SFB_CUDA(h, stg.up(&dI, innov, ny * B));
    double* dD = stg.out(d * B);
    double* dO = stg.out(dd);
    rc = sfb_ekf_step_batch_f64(h, batch, d, ny, stepper, dP, dA, dQ, tau, dt, dH, dR, dI, dD, dO);
    if (rc != SFB_OK) return rc;
    SFB_CUDA(h, stg.down(out_delta, dD, d * B)); SFB_CUDA(h, stg.down(out_P, dO, dd));
    SFB_CUDA(h, cudaStreamSynchronize(h->stream));
    return SFB_OK;
  }
  if (stepper == SFB_STEPPER_EULER && aligned16({P, A, Q, H, R, innov, out_delta, out_P})) {
    sfb::EkfStepArgs a{P, A, Q, H, R, innov, out_delta, out_P, batch, tau, dt};
    const int rc = ekf_fused_dispatch<true, true>(h, d, ny, a);
    if (rc >= 0) return rc;
  }
  // generic sizes / RK4: the two generic kernels back to back; the update works in place on out_P (each CTA stages its
  // tile of P completely before it stores)
  int rc = ekf_predict_generic(h, batch, d, stepper, P, A, Q, tau, dt, out_P);
  if (rc != SFB_OK) return rc;
  return ekf_update_generic(h, batch, d, ny, out_P, H, R, innov, out_delta, out_P);
}

int sfb_qp_sparse_analyze(sfb_handle_t h, int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx,
                          const int32_t* A_rowptr, const int32_t* A_colidx, sfb_qp_sparse_pattern_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out) return fail(h, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n <= 0 || m < 0 || !P_colptr || (m > 0 && !A_rowptr)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad pattern arguments");
  if ((P_colptr[n] > 0 && !P_rowidx) || (m > 0 && A_rowptr[m] > 0 && !A_colidx)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "index array is NULL");
  auto* p = new sfb_qp_sparse_pattern();
  p->device = h->device;
  if (!sfb::sparse_analyze(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx, p->sym)) {
    const std::string msg = p->sym.error;
    delete p;
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "sparse pattern rejected: %s", msg.c_str());
  }
  const sfb::SparseSymbolic& S = p->sym;
  const std::vector<int>* arrs[] = {&S.perm, &S.iperm, &S.P_rowp, &S.P_colp, &S.P_tgt, &S.A_rowptr, &S.A_col, &S.A_pair_ptr,
                                    &S.A_pair_tgt, &S.L_colptr, &S.L_row, &S.F_ptr, &S.F_tgt, &S.LR_ptr, &S.LR_col, &S.LR_slot,
                                    &S.AT_ptr, &S.AT_row, &S.AT_slot, &S.PR_ptr, &S.PR_col, &S.PR_slot, &S.PS_ptr, &S.PS_col,
                                    &S.PS_slot, &S.PC_ptr, &S.PC_slot, &S.LB_ptr, &S.LB_row, &S.LB_slot, &S.A_pair_ab, &S.F_ab, &S.FS_meta, &S.FS_col, &S.FS_slot,
                                    &S.BS_meta, &S.BS_col, &S.BS_slot, &S.RP_col, &S.RP_slot, &S.ATP_row, &S.ATP_slot};
  size_t total = 0;
  std::vector<size_t> off;
  for (auto* v : arrs) { off.push_back(total); total += (v->size() + 31) / 32 * 32; }
  std::vector<int> flat(total, 0);
  for (size_t k = 0; k < off.size(); ++k) std::copy(arrs[k]->begin(), arrs[k]->end(), flat.begin() + off[k]);
  if (cudaSetDevice(h->device) != cudaSuccess || cudaMalloc(&p->dev, total * sizeof(int)) != cudaSuccess ||
      cudaMemcpy(p->dev, flat.data(), total * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    if (p->dev) cudaFree(p->dev);
    delete p;
    return fail(h, SFB_ERR_CUDA, "uploading the sparse pattern failed: %s", msg);
  }
  sfb::SpPattern& d = p->pat;
  d.n = n; d.m = m; d.nnzP = S.nnzP; d.nnzA = S.nnzA; d.nnzL = S.nnzL;
  const int* base = p->dev;
  d.perm = base + off[0]; d.iperm = base + off[1]; d.P_rowp = base + off[2]; d.P_colp = base + off[3]; d.P_tgt = base + off[4];
  d.A_rowptr = base + off[5]; d.A_col = base + off[6]; d.A_pair_ptr = base + off[7]; d.A_pair_tgt = base + off[8];
  d.L_colptr = base + off[9]; d.L_row = base + off[10]; d.F_ptr = base + off[11]; d.F_tgt = base + off[12];
  d.LR_ptr = base + off[13]; d.LR_col = base + off[14]; d.LR_slot = base + off[15];
  d.AT_ptr = base + off[16]; d.AT_row = base + off[17]; d.AT_slot = base + off[18];
  d.PR_ptr = base + off[19]; d.PR_col = base + off[20]; d.PR_slot = base + off[21];
  d.PS_ptr = base + off[22]; d.PS_col = base + off[23]; d.PS_slot = base + off[24];
  d.PC_ptr = base + off[25]; d.PC_slot = base + off[26];
  d.LB_ptr = base + off[27]; d.LB_row = base + off[28]; d.LB_slot = base + off[29];
  d.A_pair_ab = base + off[30]; d.F_ab = base + off[31];
  d.FS_meta = base + off[32]; d.FS_col = base + off[33]; d.FS_slot = base + off[34];
  d.BS_meta = base + off[35]; d.BS_col = base + off[36]; d.BS_slot = base + off[37];
  d.nFS = (int)S.FS_meta.size(); d.nBS = (int)S.BS_meta.size();
  d.RP_col = base + off[38]; d.RP_slot = base + off[39]; d.ATP_row = base + off[40]; d.ATP_slot = base + off[41];
  d.WR = S.WR; d.WA = S.WA; d.m_pad = S.m_pad; d.n_pad = S.n_pad;
  *out = p;
  return SFB_OK;
}

int sfb_qp_sparse_symbolic(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                           const int32_t* A_colidx, int64_t* nnz_L, int64_t* factor_flops, int32_t* perm_out,
                           int32_t* L_colptr_out)
{
  if (n <= 0 || m < 0 || !P_colptr || (m > 0 && !A_rowptr)) return SFB_ERR_INVALID_ARGUMENT;
  sfb::SparseSymbolic S;
  if (!sfb::sparse_analyze(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx, S)) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "sparse pattern rejected: %s", S.error.c_str());
  {
    std::string why;  // every schedule the device kernel relies on is self-checked on this (test-facing) entry point
    if (!sfb::sparse_validate(S, why)) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "internal: inconsistent sparse schedules: %s", why.c_str());
  }
  if (nnz_L) *nnz_L = S.nnzL;
  if (factor_flops) *factor_flops = S.flops;
  if (perm_out) std::copy(S.perm.begin(), S.perm.end(), perm_out);
  if (L_colptr_out) std::copy(S.L_colptr.begin(), S.L_colptr.end(), L_colptr_out);
  return SFB_OK;
}

int sfb_qp_sparse_pattern_destroy(sfb_qp_sparse_pattern_t p)
{
  if (!p) return SFB_OK;
  cudaSetDevice(p->device);
  if (p->dev) cudaFree(p->dev);
  delete p;
  return SFB_OK;
}

int sfb_qp_sparse_pattern_info(sfb_qp_sparse_pattern_t p, int64_t* nnz_L, int64_t* factor_flops, int32_t* perm_out)
{
  if (!p) return SFB_ERR_INVALID_ARGUMENT;
  if (nnz_L) *nnz_L = p->sym.nnzL;
  if (factor_flops) *factor_flops = p->sym.flops;
  if (perm_out) std::copy(p->sym.perm.begin(), p->sym.perm.end(), perm_out);
  return SFB_OK;
}

int sfb_qp_solve_sparse_batch_f64(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                  int64_t batch, const double* P_vals, const double* q, const double* A_vals,
                                  const double* l, const double* u, const double* warm_x, const double* warm_y,
                                  double* out_x, double* out_y, double* out_obj, int32_t* out_status,
                                  uint32_t* out_iter, int8_t* out_active, uint32_t* out_flags)
{
  return qp_sparse_solve_impl<double>(h, pattern, prm, batch, P_vals, q, A_vals, l, u, warm_x, warm_y, out_x, out_y,
                                      out_obj, out_status, out_iter, out_active, out_flags);
}

int sfb_qp_solve_sparse_batch_f32(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                  int64_t batch, const float* P_vals, const float* q, const float* A_vals,
                                  const float* l, const float* u, const float* warm_x, const float* warm_y, float* out_x,
                                  float* out_y, float* out_obj, int32_t* out_status, uint32_t* out_iter,
                                  int8_t* out_active, uint32_t* out_flags)
{
  return qp_sparse_solve_impl<float>(h, pattern, prm, batch, P_vals, q, A_vals, l, u, warm_x, warm_y, out_x, out_y,
                                     out_obj, out_status, out_iter, out_active, out_flags);
}

}  // extern "C"
