// sfb_sparse.cu -- host side of the sparse (shared-pattern) QP entry points of include/sfb.h.
#include "sfb_internal.hpp"

#include <cmath>

#include "qp_dense_group.cuh"
#include "qp_sparse_host.hpp"
#include "qp_sparse_cta_host.hpp"
#include "qp_sparse_tiled.cuh"
#include "qp_sparse_cta.cuh"

using namespace sfbi;

// device-resident result of sfb_qp_sparse_analyze (index arrays shared by every instance of a batch)
struct sfb_qp_sparse_pattern
{
  int device = 0;
  sfb::SparseSymbolic sym;
  int* dev = nullptr;  // one allocation holding all index arrays
  sfb::SpPattern pat{};
  // sfb_qp_sparse_analyze_csc: CSR slot e of A takes the caller's CSC value csc2csr[e]
  std::vector<int> csc2csr;
  int* d_csc2csr = nullptr;
  // on-chip kernel (qp_sparse_cta.cuh): its own ordering / supernodal schedules; cta_ok == false: tiled kernel only
  struct Cta
  {
    bool ok = false;
    sfb::CtaSymbolic sym;
    int* dev = nullptr;
    sfb::CtaPattern pat{};
  } cta[2];  // [0]: fp32 layout (rows padded to 4 scalars), [1]: fp64 layout (2 scalars)
  template <typename T> const Cta& cta_for() const { return cta[sizeof(T) == 4 ? 0 : 1]; }
};

namespace {

// ---- sparse QP (shared pattern) --------------------------------------------------------------------------
// workspace bytes of one tile of tw instances computed in scalar T
template <typename T> size_t sp_tile_bytes(const sfb_qp_sparse_pattern* pt, int tw)
{
  const sfb::SparseSymbolic& S = pt->sym;
  return (sfb::sp_a_len(pt->pat, tw) + S.nnzP + sfb::sp_w_len(pt->pat, tw) + (size_t)sfb::kSpNV * S.n + (size_t)sfb::kSpMV * S.m) * tw * sizeof(T);
}

// carve the handle's tiled workspace for scalar T and launch the kernel (a's I/O pointers, mode, prm are set by the caller)
template <typename T, typename TIO>
int sp_launch(sfb_context* h, const sfb_qp_sparse_pattern* pt, sfb::SpArgs<T, TIO>& a, int tw)
{
  const sfb::SparseSymbolic& S = pt->sym;
  const int n = S.n, m = S.m;
  const long long tiles = (a.batch + tw - 1) / tw;
  const size_t wlen = sfb::sp_w_len(pt->pat, tw);  // factor + its stream-ordered copies
  const size_t alen = sfb::sp_a_len(pt->pat, tw);  // Abar + its padded row / column stream copies
  {
    T* w = static_cast<T*>(h->sparse_ws.dev);
    a.wsA = w; w += (size_t)tiles * alen * tw;
    a.wsP = w; w += (size_t)tiles * S.nnzP * tw;
    a.wsW = w; w += (size_t)tiles * wlen * tw;
    a.wsN = w; w += (size_t)tiles * sfb::kSpNV * n * tw;
    a.wsM = w;
  }
  const unsigned grid = (unsigned)std::min<long long>(tiles, 1 << 30);
  if (tw < 32) {
    const size_t smem = (size_t)(n + 1) * tw * sizeof(T);  // the solve vector of the tile + the dummy zero slot
    if (smem > h->prop.sharedMemPerBlockOptin) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "sparse QP n=%d: solve vector does not fit in shared memory", n);
    if (tw == 8) {
      SFB_CUDA(h, cudaFuncSetAttribute(sfb::qp_sparse_tiled_kernel<T, 8, TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sfb::qp_sparse_tiled_kernel<T, 8, TIO><<<grid, 32, smem, h->stream>>>(a);
    } else {
      SFB_CUDA(h, cudaFuncSetAttribute(sfb::qp_sparse_tiled_kernel<T, 4, TIO>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      sfb::qp_sparse_tiled_kernel<T, 4, TIO><<<grid, 32, smem, h->stream>>>(a);
    }
  } else {
    sfb::qp_sparse_tiled_kernel<T, 32, TIO><<<grid, 32, 0, h->stream>>>(a);
  }
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return SFB_OK;
}

// ---- on-chip kernel: one CTA per instance, working set in shared memory ----
template <typename T> size_t cta_smem_bytes(const sfb_qp_sparse_pattern* pt) { return sfb::cta_smem_bytes(pt->cta_for<T>().sym, sizeof(T)); }

template <typename T> bool cta_fits(const sfb_context* h, const sfb_qp_sparse_pattern* pt)
{
  return pt->cta_for<T>().ok && h->sparse_kernel != 1 && cta_smem_bytes<T>(pt) + 64 <= (size_t)h->prop.sharedMemPerBlockOptin;
}

// ws: device workspace for the launch (grid x (n + m) scalars of T), carved by the caller
template <typename T, typename TIO> int cta_launch(sfb_context* h, const sfb_qp_sparse_pattern* pt, sfb::CtaArgs<T, TIO>& a)
{
  const size_t smem = cta_smem_bytes<T>(pt);
  constexpr bool GT = sizeof(T) == 4;  // matches the analysis of the fp32 layout (big tables in global memory)
  auto* kern = sfb::qp_sparse_cta_kernel<T, TIO, GT>;
  SFB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 1;
  SFB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, sfb::kCtaNT, smem));
  if (occ < 1) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "on-chip sparse kernel does not fit on an SM");
  const unsigned grid = (unsigned)std::min<long long>(a.batch, (long long)h->prop.multiProcessorCount * occ);
  int rc = ensure_scratch(h, h->sparse_cta_ws, (size_t)grid * (size_t)(pt->cta_for<T>().sym.np + pt->cta_for<T>().sym.m) * sizeof(T), h->stream);
  if (rc != SFB_OK) return rc;
  a.ws = static_cast<T*>(h->sparse_cta_ws.dev);
  a.work_counter = next_counter(h, kNumSlots);
  SFB_CUDA(h, cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), h->stream));
  static const bool prof_on = getenv("SFB_CTA_PROF") != nullptr;  // dev instrumentation: cycles per phase, printed after a blocking sync
  unsigned long long* prof = nullptr;
  if (prof_on && cudaMalloc(&prof, sizeof(unsigned long long) * sfb::kPhTotal) == cudaSuccess) cudaMemset(prof, 0, sizeof(unsigned long long) * sfb::kPhTotal);
  a.prof = prof;
  kern<<<grid, sfb::kCtaNT, smem, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  if (prof) {
    unsigned long long hp[sfb::kPhTotal];
    cudaStreamSynchronize(h->stream);
    cudaMemcpy(hp, prof, sizeof(hp), cudaMemcpyDeviceToHost);
    cudaFree(prof);
    static const char* names[] = {"load", "scale", "prep", "assemble", "factor_cols", "factor_ext", "factor_inv", "rhs", "solve", "update", "check", "polish", "out"};
    double tot = 0;
    for (int k = 0; k < sfb::kPhCount; ++k) tot += (double)hp[k];
    fprintf(stderr, "[cta prof] T=%d mode=%d batch=%lld grid=%u smem=%zu cycles/instance=%.0f:", (int)sizeof(T), a.mode, a.batch, grid, smem, tot / (double)a.batch);
    for (int k = 0; k < sfb::kPhCount; ++k) fprintf(stderr, " %s=%.0f(%.1f%%)", names[k], (double)hp[k] / (double)a.batch, 100.0 * hp[k] / tot);
    fprintf(stderr, "\n[cta prof] sweep stages (kind:lgG:count cycles per solve call), %d stages:", pt->cta_for<T>().pat.nstages);
    {
      const sfb::CtaSymbolic& Cs = pt->cta_for<T>().sym;
      double calls = 0;  // solve() calls per instance = stage-0 total / ... : normalise by the solve phase instead
      double stot = 0;
      for (int k = 0; k < 32; ++k) stot += (double)hp[sfb::kPhStage0 + k];
      (void)calls;
      for (size_t k = 0; k < Cs.stages.size() / 4 && k < 32; ++k)
        fprintf(stderr, " [%d:%d:%d %.1f%%]", Cs.stages[4 * k], Cs.stages[4 * k + 3], Cs.stages[4 * k + 2], 100.0 * hp[sfb::kPhStage0 + k] / std::max(stot, 1.0));
    }
    fprintf(stderr, "\n");
  }
  return SFB_OK;
}

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
  const bool use_cta = cta_fits<T>(h, pt);
  const bool polish_cta = cta_fits<double>(h, pt);  // the fp64 second pass of an fp32 solve
  int tw = 4;
  const bool mixed = std::is_same<T, float>::value && prm->polish;
  const bool needs_tiled = !use_cta || (mixed && !polish_cta);  // only the tiled kernel has a per-instance HBM workspace
  if (needs_tiled) {
    const size_t per_inst4 = (sfb::sp_a_len(pt->pat, 4) + pt->sym.nnzP + sfb::sp_w_len(pt->pat, 4) + (size_t)sfb::kSpNV * n +
                              (size_t)sfb::kSpMV * m) * sizeof(T);
    // cudaMemGetInfo goes through the resource manager (an ioctl that serialises with NVML / nvidia-smi polling: measured
    // 10 - 70 ms stalls per call while nvidia-smi -lms ran, profiles/r02_host_stall_diag.txt), so it is asked only when the
    // workspace really has to grow
    if (h->sparse_ws.bytes < per_inst4 * (size_t)batch) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); free_b = 0; }
      if (per_inst4 * (size_t)batch > free_b / 2) tw = 32;
    }
  }
  if (h->sparse_tw == 4 || h->sparse_tw == 8 || h->sparse_tw == 32) tw = h->sparse_tw;
  // fp32 + polish runs a second, fp64 pass over the same instances (mixed precision, see sp_polish_pass): size the workspace once
  const long long tiles = (batch + tw - 1) / tw;
  const size_t ws_bytes = std::max(use_cta ? (size_t)0 : sp_tile_bytes<T>(pt, tw), (mixed && !polish_cta) ? sp_tile_bytes<double>(pt, tw) : (size_t)0) * (size_t)tiles;
  rc = ensure_scratch(h, h->sparse_ws, ws_bytes, h->stream);
  if (rc != SFB_OK) return rc;

  sfb::SpArgs<T> a{};
  a.pat = pt->pat;
  a.batch = batch;
  a.prm = *prm;
  a.max_iter_eff = prm->has_max_iter ? prm->max_iter : SFB_QP_DEVICE_ITER_CAP;
  a.mode = 0;
  a.dinf_guard = h->dinf_guard;
  // Mixed-precision polish: the ADMM iterations ran in fp32 (instances flagged POLISH_SKIPPED); the Optimal ones are
  // re-staged in fp64 and polish_qp (qp_solver.hpp:92-204) runs on the fp32 iterate and its active set.
  auto launch = [&]() -> int {
    int rc2;
    if (use_cta) {
      sfb::CtaArgs<T, T> ca{};
      ca.pat = pt->cta_for<T>().pat; ca.batch = a.batch; ca.prm = a.prm; ca.max_iter_eff = a.max_iter_eff; ca.mode = 0; ca.dinf_guard = a.dinf_guard;
      ca.P = a.P; ca.q = a.q; ca.A = a.A; ca.l = a.l; ca.u = a.u; ca.warm_x = a.warm_x; ca.warm_y = a.warm_y;
      ca.out_x = a.out_x; ca.out_y = a.out_y; ca.out_obj = a.out_obj; ca.out_status = a.out_status; ca.out_iter = a.out_iter;
      ca.out_active = a.out_active; ca.out_flags = a.out_flags;
      ca.scale_ws = nullptr;
      if (mixed && polish_cta) {  // the fp64 polish pass reuses this solve's equilibration
        rc2 = ensure_scratch(h, h->sparse_scale_ws, (size_t)a.batch * (size_t)(n + m + 1) * sizeof(T), h->stream);
        if (rc2 != SFB_OK) return rc2;
        ca.scale_ws = static_cast<T*>(h->sparse_scale_ws.dev);
      }
      rc2 = cta_launch<T, T>(h, pt, ca);
    } else {
      rc2 = sp_launch<T, T>(h, pt, a, tw);
    }
    if (rc2 != SFB_OK || !mixed) return rc2;
    if (polish_cta) {
      sfb::CtaArgs<double, T> ca{};
      ca.pat = pt->cta_for<double>().pat; ca.batch = a.batch; ca.prm = a.prm; ca.max_iter_eff = a.max_iter_eff; ca.mode = 2; ca.dinf_guard = a.dinf_guard;
      ca.P = a.P; ca.q = a.q; ca.A = a.A; ca.l = a.l; ca.u = a.u;
      ca.out_x = a.out_x; ca.out_y = a.out_y; ca.out_obj = a.out_obj; ca.out_status = a.out_status; ca.out_iter = a.out_iter;
      ca.out_active = a.out_active; ca.out_flags = a.out_flags;
      ca.scale_ws = use_cta ? static_cast<T*>(h->sparse_scale_ws.dev) : nullptr;
      return cta_launch<double, T>(h, pt, ca);
    }
    sfb::SpArgs<double, T> pa{};
    pa.pat = a.pat; pa.batch = a.batch; pa.prm = a.prm; pa.max_iter_eff = a.max_iter_eff; pa.mode = 2; pa.dinf_guard = a.dinf_guard;
    pa.P = a.P; pa.q = a.q; pa.A = a.A; pa.l = a.l; pa.u = a.u;
    pa.out_x = a.out_x; pa.out_y = a.out_y; pa.out_obj = a.out_obj; pa.out_status = a.out_status; pa.out_iter = a.out_iter;
    pa.out_active = a.out_active; pa.out_flags = a.out_flags;
    return sp_launch<double, T>(h, pt, pa, tw);
  };
  if (space == 1) {
    a.P = P; a.q = q; a.A = A; a.l = l; a.u = u; a.warm_x = warm_x; a.warm_y = warm_y;
    a.out_x = out_x; a.out_y = out_y; a.out_obj = out_obj; a.out_status = out_status; a.out_iter = out_iter;
    a.out_active = out_active; a.out_flags = out_flags;
    if (mixed && !out_active && m > 0) {  // the polish pass reads the active set the fp32 solve determined
      rc = ensure_scratch(h, h->act_tmp, (size_t)batch * m, h->stream);
      if (rc != SFB_OK) return rc;
      a.out_active = static_cast<int8_t*>(h->act_tmp.dev);
    }
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
  a.out_active = (out_active || mixed) ? dact : nullptr; a.out_flags = out_flags ? dfl : nullptr;
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

}  // namespace


namespace {
// A_vals in the caller's CSC order -> the CSR order the solver ingests (per instance gather through the pattern's map)
template <typename T> __global__ void csc_to_csr_values_kernel(const T* __restrict__ in, T* __restrict__ out, const int* __restrict__ map,
                                                             int nnz, long long batch)
{
  const long long total = batch * (long long)nnz;
  for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < total; k += (long long)gridDim.x * blockDim.x) {
    const long long b = k / nnz;
    const int e = (int)(k - b * nnz);
    out[k] = in[b * (long long)nnz + map[e]];
  }
}
}  // namespace

extern "C" {

int sfb_qp_sparse_analyze(sfb_handle_t h, int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx,
                          const int32_t* A_rowptr, const int32_t* A_colidx, sfb_qp_sparse_pattern_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out) return fail(h, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n <= 0 || m < 0 || !P_colptr || (m > 0 && !A_rowptr)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad pattern arguments");
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
  // ---- on-chip kernel: its own analyses (ordering for a short supernodal elimination tree; one layout per precision);
  // a failure only disables that path
  for (int which = 0; which < 2 && n <= 2 * sfb::kCtaNT; ++which) {
    sfb_qp_sparse_pattern::Cta& C = p->cta[which];
    sfb::CtaSymbolic& Cs = C.sym;
    if (!sfb::cta_analyze(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx, S.perm, Cs, which == 0 ? 2 : 1, -1, which == 0)) continue;
    const std::vector<int>* carr[] = {&Cs.smem_ints, &Cs.perm, &Cs.iperm, &Cs.P_rowp, &Cs.P_colp, &Cs.P_tgt, &Cs.PR_ptr, &Cs.PR_col, &Cs.PR_slot,
                                      &Cs.PS_ptr, &Cs.PS_col, &Cs.PS_slot, &Cs.PC_ptr, &Cs.PC_slot, &Cs.asm_round_sync, &Cs.asm_round_desc, &Cs.AT_word, &Cs.ext_word};
    size_t ctotal = 0;
    std::vector<size_t> coff;
    for (auto* v : carr) { coff.push_back(ctotal); ctotal += (v->size() + 31) / 32 * 32; }
    std::vector<int> cflat(ctotal, 0);
    for (size_t k = 0; k < coff.size(); ++k) std::copy(carr[k]->begin(), carr[k]->end(), cflat.begin() + coff[k]);
    if (cudaMalloc(&C.dev, ctotal * sizeof(int)) == cudaSuccess &&
        cudaMemcpy(C.dev, cflat.data(), ctotal * sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess) {
      sfb::CtaPattern& c = C.pat;
      const int* cb = C.dev;
      c.n = n; c.m = m; c.np = Cs.np; c.nnzP = Cs.nnzP; c.nnzA = Cs.nnzA; c.nW = Cs.nW; c.ns = Cs.ns; c.smax = Cs.smax;
      c.nstages = (int)Cs.stages.size() / 4; c.nrounds = (int)Cs.asm_round_sync.size(); c.nfacrounds = (int)Cs.fac_rounds.size() / 4;
      c.nints = (int)Cs.smem_ints.size();
      c.tscalars = (unsigned)sfb::cta_smem_scalars(Cs);
      for (int k = 0; k < sfb::kI_count; ++k) c.ioff[k] = Cs.smem_off[k];
      c.ints = cb + coff[0]; c.perm = cb + coff[1]; c.iperm = cb + coff[2]; c.P_rowp = cb + coff[3]; c.P_colp = cb + coff[4];
      c.P_tgt = cb + coff[5]; c.PR_ptr = cb + coff[6]; c.PR_col = cb + coff[7]; c.PR_slot = cb + coff[8];
      c.PS_ptr = cb + coff[9]; c.PS_col = cb + coff[10]; c.PS_slot = cb + coff[11]; c.PC_ptr = cb + coff[12]; c.PC_slot = cb + coff[13];
      c.asm_sync = cb + coff[14];
      c.asm_desc = reinterpret_cast<const int2*>(cb + coff[15]);
      c.ATword_g = cb + coff[16]; c.extword_g = cb + coff[17];
      C.ok = true;
    } else {
      cudaGetLastError();
      if (C.dev) { cudaFree(C.dev); C.dev = nullptr; }
    }
  }
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

int sfb_qp_sparse_cta_selfcheck(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                                const int32_t* A_colidx, int ordering, int64_t* info_out, double* max_rel_err_out)
{
  if (n <= 0 || m < 0 || !P_colptr || (m > 0 && !A_rowptr)) return SFB_ERR_INVALID_ARGUMENT;
  sfb::SparseSymbolic S0;
  if (!sfb::sparse_analyze(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx, S0)) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "sparse pattern rejected: %s", S0.error.c_str());
  double worst = 0;
  for (int lpad = 1; lpad <= 2; ++lpad) {  // the fp64 (pad 2) and fp32 (pad 4) layouts
    sfb::CtaSymbolic S;
    if (!sfb::cta_analyze(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx, S0.perm, S, lpad, ordering))
      return fail(nullptr, SFB_ERR_UNSUPPORTED_SIZE, "on-chip sparse analysis failed: %s", S.error.c_str());
    if (info_out && lpad == 2) {
      info_out[0] = S.ns; info_out[1] = S.nlev; info_out[2] = S.nW; info_out[3] = S.nnzL_true; info_out[4] = S.flops;
      info_out[5] = S.smax; info_out[6] = (int64_t)S.stages.size() / 4; info_out[7] = S.ordering;
    }
    const int np = S.np;
    // seeded values (xorshift), M = shift I + mirrored triu(P) + A^T diag(w) A in the padded order (holes: shift I), dense reference
    uint64_t st = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0 - 0.5; };
    std::vector<double> Pv(S.nnzP), Av(S.nnzA), w(m), b(np, 0.0);
    for (auto& v : Av) v = rnd();
    for (auto& v : w) v = (rnd() > -0.2) ? 0.1 + rnd() * 0.1 : 0.0;
    for (int j = 0; j < n; ++j) b[S.iperm[j]] = rnd();
    std::vector<double> M((size_t)np * np, 0.0);
    for (int j = 0; j < n; ++j)
      for (int e = P_colptr[j]; e < P_colptr[j + 1]; ++e) {
        const int r = P_rowidx[e];
        Pv[e] = (r == j) ? 1.0 + rnd() : 0.05 * rnd();
        if (j >= r) {
          const int pr = S.iperm[r], pc = S.iperm[j];
          M[(size_t)pr * np + pc] += Pv[e];
          if (pr != pc) M[(size_t)pc * np + pr] += Pv[e];
        }
      }
    const double shift = 1.0;  // keeps the random matrix positive definite (off-diagonal P entries are small)
    for (int k = 0; k < np; ++k) M[(size_t)k * np + k] += shift;
    for (int i = 0; i < m; ++i)
      for (int e1 = A_rowptr[i]; e1 < A_rowptr[i + 1]; ++e1)
        for (int e2 = A_rowptr[i]; e2 < A_rowptr[i + 1]; ++e2)
          M[(size_t)S.iperm[A_colidx[e1]] * np + S.iperm[A_colidx[e2]]] += w[i] * Av[e1] * Av[e2];
    sfb::CtaHostExec ex(S);
    ex.assemble(shift, Pv.data(), Av.data(), w.data());
    const bool ok = ex.factor();
    std::vector<double> x = b;
    ex.solve(x);
    std::vector<double> ref = b;  // dense Cholesky of M (in place), solve
    for (int k = 0; k < np; ++k) {
      double d = M[(size_t)k * np + k];
      for (int j = 0; j < k; ++j) d -= M[(size_t)k * np + j] * M[(size_t)k * np + j];
      if (!(d > 0)) return fail(nullptr, SFB_ERR_INVALID_ARGUMENT, "selfcheck: random matrix not positive definite");
      d = std::sqrt(d);
      M[(size_t)k * np + k] = d;
      for (int i = k + 1; i < np; ++i) {
        double v = M[(size_t)i * np + k];
        for (int j = 0; j < k; ++j) v -= M[(size_t)i * np + j] * M[(size_t)k * np + j];
        M[(size_t)i * np + k] = v / d;
      }
    }
    for (int i = 0; i < np; ++i) {
      double v = ref[i];
      for (int j = 0; j < i; ++j) v -= M[(size_t)i * np + j] * ref[j];
      ref[i] = v / M[(size_t)i * np + i];
    }
    for (int i = np - 1; i >= 0; --i) {
      double v = ref[i];
      for (int j = i + 1; j < np; ++j) v -= M[(size_t)j * np + i] * ref[j];
      ref[i] = v / M[(size_t)i * np + i];
    }
    double err = 0, scale = 0;
    for (int i = 0; i < np; ++i) { err = std::max(err, std::fabs(x[i] - ref[i])); scale = std::max(scale, std::fabs(ref[i])); }
    worst = std::max(worst, ok ? err / std::max(scale, 1e-300) : 1e300);
  }
  if (max_rel_err_out) *max_rel_err_out = worst;
  return SFB_OK;
}

int sfb_qp_sparse_pattern_destroy(sfb_qp_sparse_pattern_t p)
{
  if (!p) return SFB_OK;
  cudaSetDevice(p->device);
  if (p->dev) cudaFree(p->dev);
  if (p->d_csc2csr) cudaFree(p->d_csc2csr);
  for (auto& C : p->cta)
    if (C.dev) cudaFree(C.dev);
  delete p;
  return SFB_OK;
}

int sfb_qp_sparse_uses_onchip(sfb_handle_t h, sfb_qp_sparse_pattern_t p, int scalar_bytes, int64_t* info_out)
{
  if (!h || !p || (scalar_bytes != 4 && scalar_bytes != 8)) return 0;
  const sfb_qp_sparse_pattern::Cta& C = p->cta[scalar_bytes == 4 ? 0 : 1];
  if (info_out) {
    const sfb::CtaSymbolic& S = C.sym;
    info_out[0] = S.ns; info_out[1] = S.nlev; info_out[2] = S.nW; info_out[3] = S.nnzL_true; info_out[4] = S.flops;
    info_out[5] = S.smax; info_out[6] = (int64_t)S.stages.size() / 4;
    info_out[7] = C.ok ? (int64_t)sfb::cta_smem_bytes(S, (size_t)scalar_bytes) : 0;
  }
  return (scalar_bytes == 4 ? cta_fits<float>(h, p) : cta_fits<double>(h, p)) ? 1 : 0;
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

int sfb_qp_sparse_analyze_csc(sfb_handle_t h, int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx,
                              const int32_t* A_colptr, const int32_t* A_rowidx, sfb_qp_sparse_pattern_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out) return fail(h, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (n <= 0 || m < 0 || !P_colptr || !A_colptr) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad pattern arguments");
  if (A_colptr[0] != 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "pattern pointers must start at 0");
  for (int j = 0; j < n; ++j)
    if (A_colptr[j + 1] < A_colptr[j]) return fail(h, SFB_ERR_INVALID_ARGUMENT, "A_colptr not monotone");
  const int nnz = A_colptr[n];
  if (nnz > 0 && !A_rowidx) return fail(h, SFB_ERR_INVALID_ARGUMENT, "index array is NULL");
  // CSC -> CSR: counting sort by row; within a row the columns come out ascending because columns are visited in order
  std::vector<int32_t> rowptr(m + 1, 0), colidx(nnz);
  std::vector<int> map(nnz);
  for (int k = 0; k < nnz; ++k) {
    if (A_rowidx[k] < 0 || A_rowidx[k] >= m) return fail(h, SFB_ERR_INVALID_ARGUMENT, "A row index out of range");
    rowptr[A_rowidx[k] + 1] += 1;
  }
  for (int i = 0; i < m; ++i) rowptr[i + 1] += rowptr[i];
  std::vector<int32_t> fill(rowptr.begin(), rowptr.end() - 1);
  for (int j = 0; j < n; ++j)
    for (int k = A_colptr[j]; k < A_colptr[j + 1]; ++k) {
      const int e = fill[A_rowidx[k]]++;
      colidx[e] = j;
      map[e] = k;
    }
  int rc = sfb_qp_sparse_analyze(h, n, m, P_colptr, P_rowidx, rowptr.data(), colidx.data(), out);
  if (rc != SFB_OK) return rc;
  sfb_qp_sparse_pattern* p = *out;
  p->csc2csr = map;
  if (nnz > 0 && (cudaMalloc(&p->d_csc2csr, sizeof(int) * nnz) != cudaSuccess ||
                  cudaMemcpy(p->d_csc2csr, map.data(), sizeof(int) * nnz, cudaMemcpyHostToDevice) != cudaSuccess)) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    sfb_qp_sparse_pattern_destroy(p);
    *out = nullptr;
    return fail(h, SFB_ERR_CUDA, "uploading the CSC map failed: %s", msg);
  }
  return SFB_OK;
}

int sfb_qp_solve_sparse_batch_csc_f64(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                      int64_t batch, const double* P_vals, const double* q, const double* A_vals_csc,
                                      const double* l, const double* u, const double* warm_x, const double* warm_y,
                                      double* out_x, double* out_y, double* out_obj, int32_t* out_status,
                                      uint32_t* out_iter, int8_t* out_active, uint32_t* out_flags)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!pattern) return fail(h, SFB_ERR_INVALID_ARGUMENT, "pattern is NULL");
  const int nnz = pattern->sym.nnzA;
  if (nnz == 0 || batch <= 0)
    return sfb_qp_solve_sparse_batch_f64(h, pattern, prm, batch, P_vals, q, A_vals_csc, l, u, warm_x, warm_y, out_x, out_y, out_obj,
                                         out_status, out_iter, out_active, out_flags);
  if ((int)pattern->csc2csr.size() != nnz) return fail(h, SFB_ERR_INVALID_ARGUMENT, "pattern was not analysed with sfb_qp_sparse_analyze_csc");
  if (!A_vals_csc) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (mem_space(A_vals_csc) == 1) {  // device: gather kernel into a workspace, then the CSR entry point
    int rc = ensure_scratch(h, h->csc_tmp, sizeof(double) * (size_t)nnz * (size_t)batch, h->stream);
    if (rc != SFB_OK) return rc;
    double* tmp = static_cast<double*>(h->csc_tmp.dev);
    const long long total = (long long)batch * nnz;
    const int grid = (int)std::min<long long>((total + 255) / 256, (long long)h->prop.multiProcessorCount * 16);
    csc_to_csr_values_kernel<double><<<grid, 256, 0, h->stream>>>(A_vals_csc, tmp, pattern->d_csc2csr, nnz, batch);
    SFB_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return sfb_qp_solve_sparse_batch_f64(h, pattern, prm, batch, P_vals, q, tmp, l, u, warm_x, warm_y, out_x, out_y, out_obj, out_status,
                                         out_iter, out_active, out_flags);
  }
  std::vector<double> tmp((size_t)nnz * (size_t)batch);  // host: permute in place of the staging copy's source
  for (int64_t b = 0; b < batch; ++b)
    for (int e = 0; e < nnz; ++e) tmp[(size_t)b * nnz + e] = A_vals_csc[(size_t)b * nnz + pattern->csc2csr[e]];
  return sfb_qp_solve_sparse_batch_f64(h, pattern, prm, batch, P_vals, q, tmp.data(), l, u, warm_x, warm_y, out_x, out_y, out_obj,
                                       out_status, out_iter, out_active, out_flags);
}

}  // extern "C"
