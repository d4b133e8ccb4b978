// sfb_api.cu -- host side of the C ABI declared in include/sfb.h.
//
// Thin by design: argument validation, launch geometry, and the host-buffer staging pipeline.  All numerics
// live in the kernels (qp_dense_warp.cuh, ekf_kernels.cuh).  There is no CPU implementation of anything here:
// without a CUDA device every entry point fails.


#include "sfb_internal.hpp"

#include "qp_dense_group.cuh"
#include "qp_dense_skinny.cuh"

using namespace sfbi;

namespace sfbi {

std::string& create_error()
{
  static std::string e;
  return e;
}

int fail(sfb_context* h, int code, const char* fmt, ...)
{
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->last_error = buf; else create_error() = buf;
  return code;
}

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

// Work-queue heads.  One ring per stream the handle launches on (slot < kNumSlots: the staging streams, kNumSlots: the
// handle's own stream): a counter is reused only after kCountersPerStream later launches on the SAME stream, i.e. in
// stream order, so a long-running launch on another stream can never share a counter with a later one.
unsigned long long* next_counter(sfb_context* h, int slot)
{
  int& nx = h->next_counter[slot];
  unsigned long long* c = h->counters + slot * kCountersPerStream + nx;
  nx = (nx + 1) % kCountersPerStream;
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

int check_params(sfb_context* h, const sfb_qp_params* prm, int64_t batch, int n, int m)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!prm) return fail(h, SFB_ERR_INVALID_ARGUMENT, "prm is NULL");
  if (batch < 0 || n <= 0 || m < 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad sizes batch=%lld n=%d m=%d", (long long)batch, n, m);
  if (prm->stop_check_iter == 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "stop_check_iter must be > 0");
  return SFB_OK;
}

}  // namespace sfbi

namespace {

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


template <typename T, int G, int NS, int MS, typename TIO = T>
int qp_launch_g(sfb_context* h, cudaStream_t st, sfb::QpArgs<T, TIO>& args, const QpGeom& g, int grid)
{
  auto kern = sfb::qp_dense_group_kernel<T, G, qp_minb<T, G>(), NS, MS, TIO>;
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

template <typename T, typename TIO>
int qp_launch_group(sfb_context* h, cudaStream_t st, int scratch_slot, sfb::QpArgs<T, TIO>& args);

template <typename T>
int qp_launch(sfb_context* h, cudaStream_t st, int scratch_slot, const sfb::QpArgs<T>& args_in)
{
  sfb::QpArgs<T> args = args_in;
  if (qp_is_skinny(h, args.n, args.m, args.mode, args.prm)) {
    args.scratch = nullptr;
    args.scratch_per_cta = 0;
    args.work_counter = next_counter(h, scratch_slot);
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
  int rc = qp_launch_group<T, T>(h, st, scratch_slot, args);
  if (rc != SFB_OK) return rc;
  // fp32 + polish: the delta = 1e-6 regularised polish systems are not resolvable in single precision.  Mixed precision:
  // the ADMM iterations ran in fp32 above (the instance is flagged POLISH_SKIPPED); a second pass re-stages the Optimal
  // instances in fp64 and runs polish_qp (qp_solver.hpp:92-204) on the fp32 iterate and its active set.  Shapes whose
  // fp64 working set does not fit in shared memory keep the unpolished solution and the POLISH_SKIPPED flag.
  if (std::is_same<T, float>::value && args.mode == 0 && args.prm.polish && (args.out_active != nullptr || args.m == 0)) {
    QpGeom gd;
    if (qp_geometry<double>(h, args.n, args.m, &gd) == SFB_OK) {
      sfb::QpArgs<double, T> pa{};
      pa.P = args.P; pa.q = args.q; pa.A = args.A; pa.l = args.l; pa.u = args.u;
      pa.out_x = args.out_x; pa.out_y = args.out_y; pa.out_obj = args.out_obj; pa.out_status = args.out_status;
      pa.out_iter = args.out_iter; pa.out_active = args.out_active; pa.out_flags = args.out_flags;
      pa.batch = args.batch; pa.n = args.n; pa.m = args.m; pa.mode = 2; pa.prm = args.prm; pa.max_iter_eff = args.max_iter_eff; pa.dinf_guard = args.dinf_guard; pa.force_polish_scratch = args.force_polish_scratch; pa.polish_form = args.polish_form;
      rc = qp_launch_group<double, T>(h, st, scratch_slot, pa);
      if (rc != SFB_OK) return rc;
    }
  }
  return SFB_OK;
}

template <typename T, typename TIO>
int qp_launch_group(sfb_context* h, cudaStream_t st, int scratch_slot, sfb::QpArgs<T, TIO>& args)
{
  QpGeom g;
  if (qp_geometry<T>(h, args.n, args.m, &g) != SFB_OK)
    return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "dense QP n=%d m=%d (%zu-byte scalars) does not fit in shared memory",
                args.n, args.m, sizeof(T));
  const long long persistent = (long long)h->prop.multiProcessorCount * g.ctas_per_sm;
  const int grid = (int)std::max<long long>(1, std::min<long long>(args.batch, persistent));
  // polish workspace: the Schur block S (na x na, na <= min(n, m)) lives in shared memory when 2 na <= ldA
  args.scratch = nullptr;
  args.scratch_per_cta = 0;
  if (args.mode != 1 && args.prm.polish) {
    const long long k = std::min(args.n, args.m);
    const sfb::QpLayout L(args.n, args.m, 32 * g.G);
    if (2 * k > L.ldA || args.force_polish_scratch) {
      const size_t bytes = sizeof(T) * (size_t)(k * k) * (size_t)grid;
      int rc = ensure_scratch(h, h->scratch[scratch_slot], bytes, st);
      if (rc != SFB_OK) return rc;
      args.scratch = static_cast<T*>(h->scratch[scratch_slot].dev);
      args.scratch_per_cta = k * k;
    }
  }
  args.work_counter = next_counter(h, scratch_slot);
  SFB_CUDA(h, cudaMemsetAsync(args.work_counter, 0, sizeof(unsigned long long), st));
  int rc;
  // shape-specialised instantiations (compile-time n, m) for the headline shapes, generic kernels otherwise
  constexpr bool kPlain = std::is_same<T, TIO>::value;  // shape-specialised kernels exist for the plain fp64 path only
  if (g.G == 4) {
    if (kPlain && std::is_same<T, double>::value && args.n == 50 && args.m == 100) rc = qp_launch_g<T, 4, kPlain ? 50 : 0, kPlain ? 100 : 0, TIO>(h, st, args, g, grid);
    else rc = qp_launch_g<T, 4, 0, 0, TIO>(h, st, args, g, grid);
  } else if (g.G == 2) {
    rc = qp_launch_g<T, 2, 0, 0, TIO>(h, st, args, g, grid);
  } else {
    if (kPlain && std::is_same<T, double>::value && args.n == 10 && args.m == 20) rc = qp_launch_g<T, 1, kPlain ? 10 : 0, kPlain ? 20 : 0, TIO>(h, st, args, g, grid);
    else rc = qp_launch_g<T, 1, 0, 0, TIO>(h, st, args, g, grid);
  }
  if (rc != SFB_OK) return rc;
  h->launches += 1;
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
  a.dinf_guard = h->dinf_guard;
  a.force_polish_scratch = h->force_polish_scratch;
  a.polish_form = h->polish_form;

  if (space == 1) {
    a.P = P; a.q = q; a.A = A; a.l = l; a.u = u; a.warm_x = warm_x; a.warm_y = warm_y;
    a.out_x = out_x; a.out_y = out_y; a.out_obj = out_obj; a.out_status = out_status; a.out_iter = out_iter;
    a.out_active = out_active; a.out_flags = out_flags;
    if (std::is_same<T, float>::value && prm->polish && !out_active && m > 0) {
      // the mixed-precision polish pass reads the active set the fp32 solve determined
      rc = ensure_scratch(h, h->act_tmp, (size_t)batch * m, h->stream);
      if (rc != SFB_OK) return rc;
      a.out_active = static_cast<int8_t*>(h->act_tmp.dev);
    }
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
    c.out_active = (out_active || (std::is_same<T, float>::value && prm->polish)) ? reinterpret_cast<int8_t*>(d + oact) : nullptr;
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

const char* sfb_last_error_message(sfb_handle_t h) { return h ? h->last_error.c_str() : create_error().c_str(); }

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
  { const char* e = getenv("SFB_SPARSE_KERNEL"); h->sparse_kernel = (e && !strcmp(e, "tiled")) ? 1 : ((e && !strcmp(e, "cta")) ? 2 : 0); }
  if (h->sparse_tw == 4 || h->sparse_tw == 8 || h->sparse_tw == 32) h->sparse_kernel = 1;  // a forced tile width means the tiled kernel
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
  bool ok = cudaMalloc(&h->counters, sizeof(unsigned long long) * kCountersPerStream * (kNumSlots + 1)) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_start, cudaEventDisableTiming) == cudaSuccess;
  ok = ok && cudaEventCreateWithFlags(&h->ev_order, cudaEventDisableTiming) == cudaSuccess;
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
  if (h->ev_order) cudaEventDestroy(h->ev_order);
  for (auto& sc : h->scratch)
    if (sc.dev) cudaFree(sc.dev);
  if (h->sparse_ws.dev) cudaFree(h->sparse_ws.dev);
  if (h->sparse_cta_ws.dev) cudaFree(h->sparse_cta_ws.dev);
  if (h->sparse_scale_ws.dev) cudaFree(h->sparse_scale_ws.dev);
  if (h->sparse_stage.dev) cudaFree(h->sparse_stage.dev);
  if (h->act_tmp.dev) cudaFree(h->act_tmp.dev);
  if (h->csc_tmp.dev) cudaFree(h->csc_tmp.dev);
  if (h->counters) cudaFree(h->counters);
  delete h;
  return SFB_OK;
}

int sfb_set_stream(sfb_handle_t h, void* stream)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  cudaStream_t ns = static_cast<cudaStream_t>(stream);
  if (ns != h->stream) {
    // The handle owns device workspaces (sparse working set, polish scratch, staging buffers, work counters) that every call
    // reuses.  Work already enqueued on the old stream must finish with them before work on the new stream touches them.
    SFB_CUDA(h, cudaSetDevice(h->device));
    SFB_CUDA(h, cudaEventRecord(h->ev_order, h->stream));
    SFB_CUDA(h, cudaStreamWaitEvent(ns, h->ev_order, 0));
    h->stream = ns;
  }
  return SFB_OK;
}

int sfb_set_option(sfb_handle_t h, int option, int value)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  switch (option) {
    case SFB_OPT_DUAL_INF_DX_GUARD: h->dinf_guard = value ? 1 : 0; return SFB_OK;
    case SFB_OPT_FORCE_POLISH_SCRATCH: h->force_polish_scratch = value ? 1 : 0; return SFB_OK;
    case SFB_OPT_POLISH_FORM:
      if (value < 0 || value > 2) return fail(h, SFB_ERR_INVALID_ARGUMENT, "SFB_OPT_POLISH_FORM takes 0, 1 or 2");
      h->polish_form = value;
      return SFB_OK;
    default: return fail(h, SFB_ERR_INVALID_ARGUMENT, "unknown option %d", option);
  }
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


}  // extern "C"
