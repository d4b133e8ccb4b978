// sfb_ekf.cu -- host side of the EKF entry points of include/sfb.h (launch geometry, staging of host buffers).
#include "sfb_internal.hpp"

#include "ekf_fused_tma.cuh"
#include "ekf_kernels.cuh"

using namespace sfbi;

namespace {

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

extern "C" {

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

}  // extern "C"
