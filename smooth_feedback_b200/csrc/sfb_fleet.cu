// sfb_fleet.cu -- host side of the fleet entry points of include/sfb.h: controllers / filters of the reference that wrap
// the QP hot path (ASIFilter, asif.hpp:82-102) run for a whole fleet of agents with device-resident state.
#include "sfb_internal.hpp"

#include "asif_vehicle.cuh"
#include "mpc_vehicle.cuh"
#include "mpc_vehicle_host.hpp"

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
  unsigned* d_tile_ready = nullptr;  // [ceil(batch / 32)] per-launch publication flags of the transcription tiles
  // staging of host buffers
  void* d_x = nullptr;
  void* d_ud = nullptr;
  void* d_u = nullptr;
  int32_t* d_status = nullptr;
  uint32_t* d_iter = nullptr;
  int total_steps = 0;
};


struct sfb_mpc_fleet
{
  sfb_context* h = nullptr;
  sfb_mpc_vehicle_params prm{};
  sfb::MpcVehicleHost host;
  sfb_qp_sparse_pattern_t pattern = nullptr;
  int64_t batch = 0;
  int scalar_bytes = 8;
  // fleet constants on the device
  double *d_Pc = nullptr, *d_Ac = nullptr, *d_lc = nullptr, *d_uc = nullptr, *d_tau = nullptr;
  // per-agent QP values, solutions and resident warm starts (scalar type of the fleet)
  void *d_P = nullptr, *d_q = nullptr, *d_A = nullptr, *d_l = nullptr, *d_u = nullptr;
  void *d_sx = nullptr, *d_sy = nullptr, *d_obj = nullptr, *d_wx = nullptr, *d_wy = nullptr;
  uint8_t* d_warm_valid = nullptr;
  int any_warm = 0;  // host-side: has any step stored warm starts yet?
  // staging of host buffers
  void *d_t = nullptr, *d_x = nullptr, *d_uout = nullptr;
  int32_t* d_status = nullptr;
  uint32_t* d_iter = nullptr;
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
  const long long tiles = (f->batch + 31) / 32;
  a.work_counter = next_counter(h, kNumSlots);
  a.solve_counter = next_counter(h, kNumSlots);
  a.tile_ready = f->d_tile_ready;
  SFB_CUDA(h, cudaMemsetAsync(a.work_counter, 0, sizeof(unsigned long long), h->stream));
  SFB_CUDA(h, cudaMemsetAsync(a.solve_counter, 0, sizeof(unsigned long long), h->stream));
  SFB_CUDA(h, cudaMemsetAsync(a.tile_ready, 0, sizeof(unsigned) * (size_t)tiles, h->stream));
  const long long ctas = (f->batch + sfb::kSkinnyWarps - 1) / sfb::kSkinnyWarps;  // phase 2 hands out single agents
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
            cudaMalloc(&f->d_tile_ready, sizeof(unsigned) * ((B + 31) / 32)) == cudaSuccess &&
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
  void* ptrs[] = {f->d_nsteps, f->d_dt, f->d_rows, f->d_warm_x, f->d_warm_y, f->d_warm_valid, f->d_tile_ready, f->d_x, f->d_ud, f->d_u, f->d_status, f->d_iter};
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

// ---------------------------------------------------------------------------------------------------------------------
// MPC fleet
// ---------------------------------------------------------------------------------------------------------------------
namespace {

template <typename T> sfb::MpcArgs<T> mpc_args(sfb_mpc_fleet* f, const T* t, const T* x)
{
  sfb::MpcArgs<T> a{};
  const sfb::MpcVehicleHost& H = f->host;
  a.mdl.n = H.n; a.mdl.m = H.m; a.mdl.nnzP = (int)H.P_vals.size(); a.mdl.nnzA = (int)H.A_base.size();
  a.mdl.xvar_L = H.xvar_L; a.mdl.ce_row0 = H.ce_row0;
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 6; ++c) a.mdl.ce_slot[r][c] = H.ce_slot[r][c];
  for (int k = 0; k < 3; ++k) { a.mdl.g0[k] = f->prm.g0[k]; a.mdl.vdes[k] = f->prm.vdes[k]; }
  a.mdl.udes[0] = f->prm.udes[0]; a.mdl.udes[1] = f->prm.udes[1];
  a.mdl.P_vals = f->d_Pc; a.mdl.A_base = f->d_Ac; a.mdl.l_base = f->d_lc; a.mdl.u_base = f->d_uc;
  a.batch = f->batch;
  a.t = t; a.x = x;
  return a;
}

template <typename T> int mpc_transcribe(sfb_mpc_fleet* f, const T* t, const T* x, T* P, T* q, T* A, T* l, T* u)
{
  sfb_context* h = f->h;
  sfb::MpcArgs<T> a = mpc_args<T>(f, t, x);
  a.P_vals = P; a.q = q; a.A_vals = A; a.l = l; a.u = u;
  const int grid = (int)std::min<long long>(f->batch, (long long)h->prop.multiProcessorCount * 16);
  sfb::mpc_vehicle_transcribe_kernel<T><<<grid, 128, 0, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  return SFB_OK;
}

template <typename T> int sparse_solve(sfb_mpc_fleet* f, const T* P, const T* q, const T* A, const T* l, const T* u, const T* wx,
                                       const T* wy, T* ox, T* oy, T* oobj, int32_t* st, uint32_t* it);
template <> int sparse_solve<double>(sfb_mpc_fleet* f, const double* P, const double* q, const double* A, const double* l,
                                     const double* u, const double* wx, const double* wy, double* ox, double* oy, double* oobj,
                                     int32_t* st, uint32_t* it)
{
  return sfb_qp_solve_sparse_batch_f64(f->h, f->pattern, &f->prm.qp, f->batch, P, q, A, l, u, wx, wy, ox, oy, oobj, st, it, nullptr, nullptr);
}
template <> int sparse_solve<float>(sfb_mpc_fleet* f, const float* P, const float* q, const float* A, const float* l, const float* u,
                                    const float* wx, const float* wy, float* ox, float* oy, float* oobj, int32_t* st, uint32_t* it)
{
  return sfb_qp_solve_sparse_batch_f32(f->h, f->pattern, &f->prm.qp, f->batch, P, q, A, l, u, wx, wy, ox, oy, oobj, st, it, nullptr, nullptr);
}

template <typename T>
int mpc_step_impl(sfb_mpc_fleet* f, const T* t, const T* x, T* out_u, int32_t* out_status, uint32_t* out_iter, T* out_primal, T* out_dual)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  if (f->scalar_bytes != (int)sizeof(T)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "fleet was created for %d-byte scalars", f->scalar_bytes);
  if (!t || !x || !out_u || !out_status || !out_iter) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  const int space = classify({t, x, out_u, out_status, out_iter, out_primal, out_dual});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  SFB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = (size_t)f->batch;
  const int n = f->host.n, m = f->host.m;
  const T* dt = t; const T* dx = x;
  T* du = out_u; int32_t* dst = out_status; uint32_t* dit = out_iter;
  if (space == 0) {
    SFB_CUDA(h, cudaMemcpyAsync(f->d_t, t, sizeof(T) * B, cudaMemcpyHostToDevice, h->stream));
    SFB_CUDA(h, cudaMemcpyAsync(f->d_x, x, sizeof(T) * 7 * B, cudaMemcpyHostToDevice, h->stream));
    dt = static_cast<const T*>(f->d_t); dx = static_cast<const T*>(f->d_x);
    du = static_cast<T*>(f->d_uout); dst = f->d_status; dit = f->d_iter;
  }
  T *P = static_cast<T*>(f->d_P), *q = static_cast<T*>(f->d_q), *A = static_cast<T*>(f->d_A), *l = static_cast<T*>(f->d_l),
    *u = static_cast<T*>(f->d_u), *sx = static_cast<T*>(f->d_sx), *sy = static_cast<T*>(f->d_sy), *obj = static_cast<T*>(f->d_obj),
    *wx = static_cast<T*>(f->d_wx), *wy = static_cast<T*>(f->d_wy);
  int rc = mpc_transcribe<T>(f, dt, dx, P, q, A, l, u);  // mpc.hpp:473-488
  if (rc != SFB_OK) return rc;
  // mpc.hpp:491.  The reference passes its optional warm start; an agent without one starts from zeros, which is what a
  // zero warm start reproduces exactly (x = 0 / sx, y = c 0 / sy, z = Abar 0), so one launch serves both kinds of agent.
  const bool warm = f->prm.warmstart && f->any_warm;
  rc = sparse_solve<T>(f, P, q, A, l, u, warm ? wx : nullptr, warm ? wy : nullptr, sx, sy, obj, dst, dit);
  if (rc != SFB_OK) return rc;
  sfb::MpcEpilogueArgs<T> e{};
  e.batch = f->batch; e.n = n; e.m = m; e.uvar_B = f->host.xvar_L;
  e.udes[0] = f->prm.udes[0]; e.udes[1] = f->prm.udes[1];
  e.keep_warm = f->prm.warmstart;
  e.sol_x = sx; e.sol_y = sy; e.status = dst; e.warm_x = wx; e.warm_y = wy; e.warm_valid = f->d_warm_valid; e.out_u = du;
  const int grid = (int)std::min<long long>(f->batch, (long long)h->prop.multiProcessorCount * 16);
  sfb::mpc_vehicle_epilogue_kernel<T><<<grid, 128, 0, h->stream>>>(e);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  if (f->prm.warmstart) f->any_warm = 1;
  if (space == 0) {
    SFB_CUDA(h, cudaMemcpyAsync(out_u, du, sizeof(T) * 2 * B, cudaMemcpyDeviceToHost, h->stream));
    SFB_CUDA(h, cudaMemcpyAsync(out_status, dst, 4 * B, cudaMemcpyDeviceToHost, h->stream));
    SFB_CUDA(h, cudaMemcpyAsync(out_iter, dit, 4 * B, cudaMemcpyDeviceToHost, h->stream));
    if (out_primal) SFB_CUDA(h, cudaMemcpyAsync(out_primal, sx, sizeof(T) * n * B, cudaMemcpyDeviceToHost, h->stream));
    if (out_dual) SFB_CUDA(h, cudaMemcpyAsync(out_dual, sy, sizeof(T) * m * B, cudaMemcpyDeviceToHost, h->stream));
    SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  } else {
    if (out_primal) SFB_CUDA(h, cudaMemcpyAsync(out_primal, sx, sizeof(T) * n * B, cudaMemcpyDeviceToDevice, h->stream));
    if (out_dual) SFB_CUDA(h, cudaMemcpyAsync(out_dual, sy, sizeof(T) * m * B, cudaMemcpyDeviceToDevice, h->stream));
  }
  return SFB_OK;
}

// mpc.hpp:493-507 on the solution the last step left on the device
template <typename T> int mpc_traj_impl(sfb_mpc_fleet* f, const T* t, T* out_u_traj, T* out_x_traj)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  if (f->scalar_bytes != (int)sizeof(T)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "fleet was created for %d-byte scalars", f->scalar_bytes);
  if (!t) return fail(h, SFB_ERR_INVALID_ARGUMENT, "t is NULL");
  if (!out_u_traj && !out_x_traj) return SFB_OK;
  const int space = classify({t, out_u_traj, out_x_traj});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  SFB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = (size_t)f->batch;
  const int N = f->host.N;
  const size_t ub = sizeof(T) * 2 * (size_t)N * B, xb = sizeof(T) * 7 * (size_t)(N + 1) * B;
  sfb::MpcTrajArgs<T> a{};
  a.batch = f->batch; a.N = N; a.n = f->host.n; a.xvar_L = f->host.xvar_L; a.tf = f->prm.tf;
  for (int k = 0; k < 3; ++k) { a.g0[k] = f->prm.g0[k]; a.vdes[k] = f->prm.vdes[k]; }
  a.udes[0] = f->prm.udes[0]; a.udes[1] = f->prm.udes[1];
  a.tau = f->d_tau; a.sol_x = static_cast<const T*>(f->d_sx);
  a.t = t; a.u_traj = out_u_traj; a.x_traj = out_x_traj;
  if (space == 0) {  // host buffers: staged through the handle's scratch
    int rc = sfbi::ensure_scratch(h, h->sparse_stage, sizeof(T) * B + ub + xb + 512, h->stream);
    if (rc != SFB_OK) return rc;
    char* d = static_cast<char*>(h->sparse_stage.dev);
    auto al = [](size_t v) { return (v + 255) / 256 * 256; };
    T* dt = reinterpret_cast<T*>(d);
    T* du = reinterpret_cast<T*>(d + al(sizeof(T) * B));
    T* dx = reinterpret_cast<T*>(d + al(sizeof(T) * B) + al(ub));
    SFB_CUDA(h, cudaMemcpyAsync(dt, t, sizeof(T) * B, cudaMemcpyHostToDevice, h->stream));
    a.t = dt; a.u_traj = out_u_traj ? du : nullptr; a.x_traj = out_x_traj ? dx : nullptr;
  }
  const long long total = (long long)B * (N + 1);
  const int grid = (int)std::min<long long>((total + 127) / 128, (long long)h->prop.multiProcessorCount * 16);
  sfb::mpc_vehicle_traj_kernel<T><<<grid, 128, 0, h->stream>>>(a);
  SFB_CUDA(h, cudaGetLastError());
  h->launches += 1;
  if (space == 0) {
    if (out_u_traj) SFB_CUDA(h, cudaMemcpyAsync(out_u_traj, a.u_traj, ub, cudaMemcpyDeviceToHost, h->stream));
    if (out_x_traj) SFB_CUDA(h, cudaMemcpyAsync(out_x_traj, a.x_traj, xb, cudaMemcpyDeviceToHost, h->stream));
    SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  }
  return SFB_OK;
}

}  // namespace

extern "C" {

void sfb_mpc_vehicle_params_default(sfb_mpc_vehicle_params* p)
{
  if (!p) return;
  p->K = 50; p->tf = 5; p->Kmesh = 4; p->warmstart = 1;
  for (int i = 0; i < 6; ++i) { p->Q[i] = 1; p->Qtf[i] = 1; }
  p->R[0] = p->R[1] = 1;
  p->crl[0] = p->crl[1] = -0.5;
  p->cru[0] = p->cru[1] = 0.5;
  p->drag1 = 0.2; p->drag3 = 0.4;
  p->g0[0] = 2.5; p->g0[1] = 0; p->g0[2] = 1.57079632679489661923;  // M_PI_2
  p->vdes[0] = 1; p->vdes[1] = 0; p->vdes[2] = 0.4;
  p->udes[0] = p->udes[1] = 0;
  sfb_qp_params_default(&p->qp);
}

int sfb_mpc_fleet_create(sfb_handle_t h, const sfb_mpc_vehicle_params* p, int64_t batch, int scalar_bytes, sfb_mpc_fleet_t* out)
{
  if (!h) return SFB_ERR_INVALID_ARGUMENT;
  if (!out) return fail(h, SFB_ERR_INVALID_ARGUMENT, "out is NULL");
  *out = nullptr;
  if (!p || batch <= 0 || (scalar_bytes != 4 && scalar_bytes != 8)) return fail(h, SFB_ERR_INVALID_ARGUMENT, "bad argument to sfb_mpc_fleet_create");
  if (p->qp.stop_check_iter == 0) return fail(h, SFB_ERR_INVALID_ARGUMENT, "stop_check_iter must be > 0");
  auto* f = new sfb_mpc_fleet();
  f->h = h; f->prm = *p; f->batch = batch; f->scalar_bytes = scalar_bytes;
  if (!sfb::mpc_vehicle_build(*p, f->host)) {
    const std::string msg = f->host.error;
    delete f;
    return fail(h, SFB_ERR_INVALID_ARGUMENT, "MPC parameters rejected: %s", msg.c_str());
  }
  const sfb::MpcVehicleHost& H = f->host;
  int rc = sfb_qp_sparse_analyze(h, H.n, H.m, H.P_colptr.data(), H.P_rowidx.data(), H.A_rowptr.data(), H.A_colidx.data(), &f->pattern);  // mpc.hpp:424
  if (rc != SFB_OK) { delete f; return rc; }
  const size_t B = (size_t)batch, sb = (size_t)scalar_bytes, nP = H.P_vals.size(), nA = H.A_base.size(), n = H.n, m = H.m;
  auto dm = [&](void** ptr, size_t bytes) { return cudaMalloc(ptr, std::max<size_t>(bytes, 16)) == cudaSuccess; };
  bool ok = cudaSetDevice(h->device) == cudaSuccess;
  ok = ok && dm((void**)&f->d_Pc, 8 * nP) && dm((void**)&f->d_Ac, 8 * nA) && dm((void**)&f->d_lc, 8 * m) && dm((void**)&f->d_uc, 8 * m);
  ok = ok && dm((void**)&f->d_tau, 8 * H.tau.size()) &&
       cudaMemcpy(f->d_tau, H.tau.data(), 8 * H.tau.size(), cudaMemcpyHostToDevice) == cudaSuccess;
  ok = ok && dm(&f->d_P, sb * nP * B) && dm(&f->d_q, sb * n * B) && dm(&f->d_A, sb * nA * B) && dm(&f->d_l, sb * m * B) && dm(&f->d_u, sb * m * B);
  ok = ok && dm(&f->d_sx, sb * n * B) && dm(&f->d_sy, sb * m * B) && dm(&f->d_obj, sb * B) && dm(&f->d_wx, sb * n * B) && dm(&f->d_wy, sb * m * B);
  ok = ok && dm((void**)&f->d_warm_valid, B) && dm(&f->d_t, sb * B) && dm(&f->d_x, sb * 7 * B) && dm(&f->d_uout, sb * 2 * B);
  ok = ok && dm((void**)&f->d_status, 4 * B) && dm((void**)&f->d_iter, 4 * B);
  ok = ok && cudaMemcpy(f->d_Pc, H.P_vals.data(), 8 * nP, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(f->d_Ac, H.A_base.data(), 8 * nA, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(f->d_lc, H.l_base.data(), 8 * m, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemcpy(f->d_uc, H.u_base.data(), 8 * m, cudaMemcpyHostToDevice) == cudaSuccess &&
       cudaMemset(f->d_warm_valid, 0, B) == cudaSuccess && cudaMemset(f->d_wx, 0, sb * n * B) == cudaSuccess &&
       cudaMemset(f->d_wy, 0, sb * m * B) == cudaSuccess;
  if (!ok) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    sfb_mpc_fleet_destroy(f);
    return fail(h, SFB_ERR_OUT_OF_MEMORY, "allocating the MPC fleet state failed: %s", msg);
  }
  *out = f;
  return SFB_OK;
}

int sfb_mpc_fleet_destroy(sfb_mpc_fleet_t f)
{
  if (!f) return SFB_OK;
  cudaSetDevice(f->h->device);
  cudaStreamSynchronize(f->h->stream);
  void* ptrs[] = {f->d_tau, f->d_Pc, f->d_Ac, f->d_lc, f->d_uc, f->d_P, f->d_q, f->d_A, f->d_l, f->d_u, f->d_sx, f->d_sy, f->d_obj, f->d_wx,
                  f->d_wy, f->d_warm_valid, f->d_t, f->d_x, f->d_uout, f->d_status, f->d_iter};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (f->pattern) sfb_qp_sparse_pattern_destroy(f->pattern);
  delete f;
  return SFB_OK;
}

int sfb_mpc_fleet_reset_warmstart(sfb_mpc_fleet_t f)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  SFB_CUDA(h, cudaSetDevice(h->device));
  const size_t B = (size_t)f->batch, sb = (size_t)f->scalar_bytes;
  SFB_CUDA(h, cudaMemsetAsync(f->d_warm_valid, 0, B, h->stream));
  SFB_CUDA(h, cudaMemsetAsync(f->d_wx, 0, sb * f->host.n * B, h->stream));
  SFB_CUDA(h, cudaMemsetAsync(f->d_wy, 0, sb * f->host.m * B, h->stream));
  f->any_warm = 0;
  return SFB_OK;
}

int sfb_mpc_fleet_dims(sfb_mpc_fleet_t f, int* n, int* m, int* nnzP, int* nnzA, int64_t* nnzL)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  if (n) *n = f->host.n;
  if (m) *m = f->host.m;
  if (nnzP) *nnzP = (int)f->host.P_vals.size();
  if (nnzA) *nnzA = (int)f->host.A_base.size();
  if (nnzL) return sfb_qp_sparse_pattern_info(f->pattern, nnzL, nullptr, nullptr);
  return SFB_OK;
}

int sfb_mpc_fleet_pattern(sfb_mpc_fleet_t f, int32_t* P_colptr, int32_t* P_rowidx, int32_t* A_rowptr, int32_t* A_colidx)
{
  if (!f || !P_colptr || !P_rowidx || !A_rowptr || !A_colidx) return SFB_ERR_INVALID_ARGUMENT;
  const sfb::MpcVehicleHost& H = f->host;
  std::copy(H.P_colptr.begin(), H.P_colptr.end(), P_colptr);
  std::copy(H.P_rowidx.begin(), H.P_rowidx.end(), P_rowidx);
  std::copy(H.A_rowptr.begin(), H.A_rowptr.end(), A_rowptr);
  std::copy(H.A_colidx.begin(), H.A_colidx.end(), A_colidx);
  return SFB_OK;
}

int sfb_mpc_fleet_to_qp_f64(sfb_mpc_fleet_t f, const double* t, const double* x, double* P_vals, double* q, double* A_vals,
                            double* l, double* u)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  sfb_context* h = f->h;
  if (f->scalar_bytes != 8) return fail(h, SFB_ERR_INVALID_ARGUMENT, "sfb_mpc_fleet_to_qp_f64 needs an fp64 fleet");
  if (!t || !x || !P_vals || !q || !A_vals || !l || !u) return fail(h, SFB_ERR_INVALID_ARGUMENT, "required pointer is NULL");
  const int space = classify({t, x, P_vals, q, A_vals, l, u});
  if (space < 0) return fail(h, SFB_ERR_MIXED_MEMORY, "host and device pointers mixed in one call");
  SFB_CUDA(h, cudaSetDevice(h->device));
  if (space == 1) return mpc_transcribe<double>(f, t, x, P_vals, q, A_vals, l, u);
  const size_t B = (size_t)f->batch, nP = f->host.P_vals.size(), nA = f->host.A_base.size(), n = f->host.n, m = f->host.m;
  SFB_CUDA(h, cudaMemcpyAsync(f->d_t, t, 8 * B, cudaMemcpyHostToDevice, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(f->d_x, x, 8 * 7 * B, cudaMemcpyHostToDevice, h->stream));
  double *P = static_cast<double*>(f->d_P), *qq = static_cast<double*>(f->d_q), *A = static_cast<double*>(f->d_A),
         *ll = static_cast<double*>(f->d_l), *uu = static_cast<double*>(f->d_u);
  int rc = mpc_transcribe<double>(f, static_cast<const double*>(f->d_t), static_cast<const double*>(f->d_x), P, qq, A, ll, uu);
  if (rc != SFB_OK) return rc;
  SFB_CUDA(h, cudaMemcpyAsync(P_vals, P, 8 * nP * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(q, qq, 8 * n * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(A_vals, A, 8 * nA * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(l, ll, 8 * m * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaMemcpyAsync(u, uu, 8 * m * B, cudaMemcpyDeviceToHost, h->stream));
  SFB_CUDA(h, cudaStreamSynchronize(h->stream));
  return SFB_OK;
}

int sfb_mpc_fleet_nodes(sfb_mpc_fleet_t f, int* N, double* tau)
{
  if (!f) return SFB_ERR_INVALID_ARGUMENT;
  if (N) *N = f->host.N;
  if (tau) std::copy(f->host.tau.begin(), f->host.tau.end(), tau);
  return SFB_OK;
}

int sfb_mpc_fleet_trajectories_f64(sfb_mpc_fleet_t f, const double* t, double* out_u_traj, double* out_x_traj)
{
  return mpc_traj_impl<double>(f, t, out_u_traj, out_x_traj);
}

int sfb_mpc_fleet_trajectories_f32(sfb_mpc_fleet_t f, const float* t, float* out_u_traj, float* out_x_traj)
{
  return mpc_traj_impl<float>(f, t, out_u_traj, out_x_traj);
}

int sfb_mpc_fleet_step_f64(sfb_mpc_fleet_t f, const double* t, const double* x, double* out_u, int32_t* out_status,
                           uint32_t* out_iter, double* out_primal, double* out_dual)
{
  return mpc_step_impl<double>(f, t, x, out_u, out_status, out_iter, out_primal, out_dual);
}

int sfb_mpc_fleet_step_f32(sfb_mpc_fleet_t f, const float* t, const float* x, float* out_u, int32_t* out_status,
                           uint32_t* out_iter, float* out_primal, float* out_dual)
{
  return mpc_step_impl<float>(f, t, x, out_u, out_status, out_iter, out_primal, out_dual);
}

}  // extern "C"
