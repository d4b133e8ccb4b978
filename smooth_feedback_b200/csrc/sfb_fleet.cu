// sfb_fleet.cu -- host side of the fleet entry points of include/sfb.h: controllers / filters of the reference that wrap
// the QP hot path (ASIFilter, asif.hpp:82-102) run for a whole fleet of agents with device-resident state.
#include "sfb_internal.hpp"

#include "asif_vehicle.cuh"

using namespace sfbi;

struct sfb_asif_fleet
{
  sfb_context* h = nullptr;
  sfb_asif_vehicle_params prm{};
  int64_t batch = 0;
  int scalar_bytes = 8;
  int m = 0;
  int use_warm = 1;
  // device state
  int* d_nsteps = nullptr;
  double* d_dt = nullptr;
  void* d_rows = nullptr;        // [batch][3][K]
  void* d_warm_x = nullptr;      // [batch][3]
  void* d_warm_y = nullptr;      // [batch][m]
  uint8_t* d_warm_valid = nullptr;
  // staging of host buffers
  void* d_x = nullptr;
  void* d_ud = nullptr;
  void* d_u = nullptr;
  int32_t* d_status = nullptr;
  uint32_t* d_iter = nullptr;
  int total_steps = 0;
};

namespace {

// The step schedule of asif_to_qp_update (asif_func.hpp:139-143,170-176), with the reference's exact time arithmetic:
// dt_act is fixed BEFORE the inner while loop, so an interval may overshoot its end by up to one step.
void asif_schedule(double T, int K, double dt_max, std::vector<int>& nsteps, std::vector<double>& dts)
{
  const double tau = T / static_cast<double>(K);
  const double dt = std::min<double>(dt_max, tau);
  double t = 0;
  nsteps.assign(K, 0);
  dts.assign(K, 0.0);
  for (int k = 0; k != K; ++k) {
    const double dt_act = std::min(dt, tau * (k + 1) - t);
    dts[k] = dt_act;
    while (t < tau * (k + 1)) {
      nsteps[k] += 1;
      t += dt_act;
    }
  }
}

template <typename T>
int asif_launch(sfb_asif_fleet* f, const T* x, const T* ud, T* out_u, int32_t* out_status, uint32_t* out_iter, T* qP, T* qq,
                T* qA, T* ql, T* qu)
{
  sfb_context* h = f->h;
  sfb::AsifArgs<T> a{};
  const sfb_asif_vehicle_params& p = f->prm;
  a.mdl.K = p.K;
  a.mdl.alpha = p.alpha;
  a.mdl.relax_cost = p.relax_cost;
  for (int i = 0; i < 2; ++i) {
    a.mdl.w_u[i] = p.u_weight[i];
    a.mdl.ulim_l[i] = p.ulim_l[i];
    a.mdl.ulim_u[i] = p.ulim_u[i];
  }
  a.mdl.drag1 = p.drag1; a.mdl.drag3 = p.drag3; a.mdl.cx = p.centre[0]; a.mdl.cy = p.centre[1];
  a.mdl.radius = p.radius; a.mdl.bu_gain = p.bu_gain; a.mdl.bu_const = p.bu_const;
  a.mdl.nsteps = f->d_nsteps;
  a.mdl.dt_act = f->d_dt;
  a.prm = p.qp;
  a.max_iter_eff = p.qp.has_max_iter ? p.qp.max_iter : SFB_QP_DEVICE_ITER_CAP;
  a.dinf_guard = h->dinf_guard;
  a.batch = f->batch;
  a.x = x; a.u_des = ud;
  a.rows = static_cast<T*>(f->d_rows);
  const bool warm = f->use_warm && qA == nullptr;
  a.warm_x = warm ? static_cast<T*>(f->d_warm_x) : nullptr;
  a.warm_y = warm ? static_cast<T*>(f->d_warm_y) : nullptr;
  a.warm_valid = f->d_warm_valid;
  a.out_u = out_u; a.out_status = out_status; a.out_iter = out_iter;
  a.qp_P = qP; a.qp_q = qq; a.qp_A = qA; a.qp_l = ql; a.qp_u = qu;
  a.work_counter = next_counter(h, kNumSlots);
  SFB_CUDA(h, cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), h->stream));
  const long long tiles = (f->batch + 31) / 32;
  const long long ctas = (tiles + sfb::kSkinnyWarps - 1) / sfb::kSkinnyWarps;
  auto go = [&](auto kern) -> int {
    int nb = 0;
    SFB_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, 32 * sfb::kSkinnyWarps, 0));
    if (nb < 1) return fail(h, SFB_ERR_CUDA, "ASIF kernel cannot be resident");
    const int grid = (int)std::max<long long>(1, std::min<long long>(ctas, (long long)h->prop.multiProcessorCount * nb));
    kern<<<grid, 32 * sfb::kSkinnyWarps, 0, h->stream>>>(a);
    SFB_CUDA(h, cudaGetLastError());
    h->launches += 1;
    return SFB_OK;
  };
  return f->m <= 128 ? go(sfb::asif_vehicle_filter_kernel<T, 4>) : go(sfb::asif_vehicle_filter_kernel<T, 8>);
}

template <typename T>
int asif_filter_impl(sfb_asif_fleet* f, const T* x, const T* ud, T* out_u, int32_t* out_status, uint32_t* out_iter)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  if (f->scalar_bytes != (int)sizeof(T)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "fleet was created for %d-byte scalars", f->scalar_bytes);
  if (!x || !ud || !out_u || !out_status || !out_iter) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  const int space = classify({x, ud, out_u, out_status, out_iter});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 1) return asif_launch<T>(f, x, ud, out_u, out_status, out_iter, nullptr, nullptr, nullptr, nullptr, nullptr);
  const size_t B = (size_t)f->batch;
  SFB_CUDA(h, cudaMemcpyAsync(f->d_x, x, sizeof(T) * 7 * B, cudaMemcpyHostToDevice, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(f->d_ud, ud, sizeof(T) * 2 * B, cudaMemcpyHostToDevice, h->stream));
  int rc = asif_launch<T>(f, static_cast<const T*>(f->d_x), static_cast<const T*>(f->d_ud), static_cast<T*>(f->d_u), f->d_status,
                          f->d_iter, nullptr, nullptr, nullptr, nullptr, nullptr);
  if (rc != SFB_OK) return rc;
  SFB_CUDA(h, cudaMemcpyAsync(out_u, f->d_u, sizeof(T) * 2 * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(out_status, f->d_status, 4 * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(out_iter, f->d_iter, 4 * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SFB_OK;
}

}  // namespace

extern "C" {

void sfb_asif_vehicle_params_default(sfb_asif_vehicle_params* p)
{
  if (!p) return;
  // examples/mpc_asif_vehicle.cpp:96-129
  p->T = 2.5;
  p->K = 200;
  p->alpha = 5;
  p->dt = 0.01;
  p->relax_cost = 100;
  p->u_weight[0] = 20; p->u_weight[1] = 1;
  p->ulim_l[0] = -0.2; p->ulim_l[1] = -0.5;
  p->ulim_u[0] = 0.5; p->ulim_u[1] = 0.5;
  p->drag1 = 0.2; p->drag3 = 0.4;
  p->centre[0] = 0; p->centre[1] = -2.3;
  p->radius = 0.7;
  p->bu_gain = 0.2;
  p->bu_const = -0.5;
  sfb_qp_params_default(&p->qp);
  p->qp.polish = 0;
}

int sfb_asif_fleet_create(sfb_handle_t h, const sfb_asif_vehicle_params* p, int64_t batch, int scalar_bytes,
                          sfb_asif_fleet_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out) return fail(h, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!p || batch <= 0 || (scalar_bytes != 4 && scalar_bytes != 8)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_asif_fleet_create");
  if (p->K < 1 || !(p->T > 0) || !(p->dt > 0) || p->qp.stop_check_iter == 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad ASIF parameters (K, T, dt, stop_check_iter)");
  if (p->K + 3 > sfb::kSkinnyMaxM) return fail(h, SFB_ERR_UNSUPPORTED_SIZE, "K + 3 = %d rows exceed the %d of the register-resident solver", p->K + 3, sfb::kSkinnyMaxM);
  if (p->qp.polish) return fail(h, SFB_ERR_INVALID_ARGUMENT, "the fleet filter runs with qp.polish = 0 (as examples/mpc_asif_vehicle.cpp:127 does)");
  SFB_CUDA(h, cudaSetDevice(h->device));
  auto* f = new sfb_asif_fleet();
  f->h = h; f->prm = *p; f->batch = batch; f->scalar_bytes = scalar_bytes; f->m = p->K + 3;
  std::vector<int> ns;
  std::vector<double> dts;
  asif_schedule(p->T, p->K, p->dt, ns, dts);
  for (int v : ns) f->total_steps += v;
  const size_t B = (size_t)batch, sb = (size_t)scalar_bytes;
  bool ok = cudaMalloc(&f->d_nsteps, sizeof(int) * p->K) == cudaSuccess && cudaMalloc(&f->d_dt, sizeof(double) * p->K) == cudaSuccess &&
            cudaMalloc(&f->d_rows, sb * 3 * p->K * B) == cudaSuccess && cudaMalloc(&f->d_warm_x, sb * 3 * B) == cudaSuccess &&
            cudaMalloc(&f->d_warm_y, sb * f->m * B) == cudaSuccess && cudaMalloc(&f->d_warm_valid, B) == cudaSuccess &&
            cudaMalloc(&f->d_x, sb * 7 * B) == cudaSuccess && cudaMalloc(&f->d_ud, sb * 2 * B) == cudaSuccess &&
            cudaMalloc(&f->d_u, sb * 2 * B) == cudaSuccess && cudaMalloc(&f->d_status, 4 * B) == cudaSuccess &&
            cudaMalloc(&f->d_iter, 4 * B) == cudaSuccess;
  ok = ok && cudaMemcpy(f->d_nsteps, ns.data(), sizeof(int) * p->K, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(f->d_dt, dts.data(), sizeof(double) * p->K, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemset(f->d_warm_valid, 0, B) == cudaSuccess;
  if (!ok) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    sfb_asif_fleet_destroy(f);
    return fail(h, SFB_ERR_OUT_OF_MEMORY, "allocating the ASIF fleet state failed: %s", msg);
  }
  *out = f;
  return SFB_OK;
}

int sfb_asif_fleet_destroy(sfb_asif_fleet_t f)
{
  if (!f) return SFB_OK;
  cudaSetDevice(f->h->device);
  cudaStreamSynchronize(f->h->stream);
  void* ptrs[] = {f->d_nsteps, f->d_dt, f->d_rows, f->d_warm_x, f->d_warm_y, f->d_warm_valid, f->d_x, f->d_ud, f->d_u, f->d_status, f->d_iter};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  delete f;
  return SFB_OK;
}

int sfb_asif_fleet_reset_warmstart(sfb_asif_fleet_t f)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  SFB_CUDA(h, cudaSetDevice(h->device));
  SFB_CUDA(h, cudaMemsetAsync(f->d_warm_valid, 0, (size_t)f->batch, h->stream));
  return SFB_OK;
}

int sfb_asif_fleet_set_warmstart(sfb_asif_fleet_t f, int warm)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  f->use_warm = warm ? 1 : 0;
  return SFB_OK;
}

int sfb_asif_fleet_filter_f64(sfb_asif_fleet_t f, const double* x, const double* u_des, double* out_u,
                              int32_t* out_status, uint32_t* out_iter)
{
  return asif_filter_impl<double>(f, x, u_des, out_u, out_status, out_iter);
}

int sfb_asif_fleet_filter_f32(sfb_asif_fleet_t f, const float* x, const float* u_des, float* out_u, int32_t* out_status,
                              uint32_t* out_iter)
{
  return asif_filter_impl<float>(f, x, u_des, out_u, out_status, out_iter);
}

int sfb_asif_fleet_to_qp_f64(sfb_asif_fleet_t f, const double* x, const double* u_des, double* P, double* q, double* A,
                             double* l, double* u)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  if (f->scalar_bytes != 8) return fail(h, SFB_ERR_INVALID_ARGUMENT, "sfb_asif_fleet_to_qp_f64 needs an fp64 fleet");
  if (!x || !u_des || !P || !q || !A || !l || !u) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  const int space = classify({x, u_des, P, q, A, l, u});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 1) return asif_launch<double>(f, x, u_des, nullptr, nullptr, nullptr, P, q, A, l, u);
  const size_t B = (size_t)f->batch, m = (size_t)f->m;
  double *dP, *dq, *dA, *dl, *du;
  const size_t total = sizeof(double) * B * (9 + 3 + 3 * m + 2 * m);
  int rc = ensure_scratch(h, h->sparse_stage, total, h->stream);
  if (rc != SFB_OK) return rc;
  dP = static_cast<double*>(h->sparse_stage.dev); dq = dP + 9 * B; dA = dq + 3 * B; dl = dA + 3 * m * B; du = dl + m * B;
  SFB_CUDA(h, cudaMemcpyAsync(f->d_x, x, sizeof(double) * 7 * B, cudaMemcpyHostToDevice, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(f->d_ud, u_des, sizeof(double) * 2 * B, cudaMemcpyHostToDevice, h->stream));
  rc = asif_launch<double>(f, static_cast<const double*>(f->d_x), static_cast<const double*>(f->d_ud), nullptr, nullptr, nullptr, dP, dq, dA, dl, du);
  if (rc != SFB_OK) return rc;
  SFB_CUDA(h, cudaMemcpyAsync(P, dP, sizeof(double) * 9 * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(q, dq, sizeof(double) * 3 * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(A, dA, sizeof(double) * 3 * m * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(l, dl, sizeof(double) * m * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(u, du, sizeof(double) * m * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SFB_OK;
}

}  // extern "C"
