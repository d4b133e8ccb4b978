/*
 * sfb.h -- C ABI of the B200-native batched QP / EKF engine (libsfb.so).
 *
 * This is the drop-in boundary for smooth_feedback's numerical hot path.  The reference has no FFI of
 * its own (header-only C++20 templates), so each entry point cites the reference *call* it replaces;
 * file:line are into the reference tree (pettni/smooth_feedback @ 9a08971).
 *
 *   sfb_qp_solve_dense_batch_f64/_f32   <- QPSolver<Pbm>::solve      include/smooth/feedback/qp_solver.hpp:343-568
 *                                          solve_qp                   include/smooth/feedback/qp_solver.hpp:779-787
 *                                          (call sites: mpc.hpp:491, asif.hpp:97)
 *   sfb_qp_params / _default            <- QPSolverParams             qp_solver.hpp:29-68
 *   sfb_qp_status                       <- QPSolutionStatus           qp.hpp:82-92
 *   sfb_qp_sparse_analyze               <- SimplicialLDLT::analyzePattern, call site qp_solver.hpp:424 (mpc.hpp:424)
 *   sfb_qp_solve_sparse_batch_f64/_f32  <- QPSolver<QuadraticProgramSparse>::solve   qp_solver.hpp:343-568 (mpc.hpp:491)
 *   sfb_ekf_predict_batch_f64           <- EKF::predict (cov. ODE)    ekf.hpp:79-103
 *   sfb_ekf_update_batch_f64            <- EKF::update                ekf.hpp:116-139
 *   sfb_ekf_step_batch_f64              <- EKF::predict + EKF::update ekf.hpp:79-139 (fused, one HBM pass)
 *
 * Conventions
 *   - plain pointers and sizes only; every array is a contiguous batch of per-instance blocks laid out
 *     exactly like the reference's Eigen members (column-major matrices):
 *        P [batch][n*n]   P_ij at  i + n*j        q [batch][n]
 *        A [batch][m*n]   A_ij at  i + m*j        l,u [batch][m]   (+-INFINITY allowed)
 *   - every data pointer of one call must live in the same memory space: all device (resident on the
 *     handle's device) or all host.  Host buffers are staged through the engine's own device workspace
 *     in pipelined chunks (pinned host memory makes the copies asynchronous).
 *   - functions return 0 (SFB_OK) or an sfb_error; per-instance outcomes are reported only through
 *     out_status (never an error code), like the reference which never throws on the numeric path.
 *   - there is NO CPU fallback: without a CUDA device sfb_create fails with SFB_ERR_NO_DEVICE.
 *   - a handle is not re-entrant (like a QPSolver object, qp_solver.hpp:734-756); use one per host thread.  A handle owns
 *     device workspaces that every call reuses: calls are ordered on the handle's stream, and sfb_set_stream makes the new
 *     stream wait for the work already enqueued on the old one, so switching streams never lets two calls share a workspace
 *     concurrently.  For genuinely concurrent streams use one handle per stream.
 *     Work is enqueued on the handle's stream; calls with device pointers are asynchronous with respect
 *     to the host unless stated otherwise, calls with host pointers return after the results are in place.
 */
#ifndef SFB_H
#define SFB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFB_VERSION 200 /* 0.2.0 */

typedef struct sfb_context* sfb_handle_t;

typedef enum {
  SFB_OK = 0,
  SFB_ERR_INVALID_ARGUMENT = 1, /* null pointer, non-positive size, stop_check_iter == 0, ... */
  SFB_ERR_NO_DEVICE = 2,        /* no CUDA device / wrong architecture: the engine has no CPU path */
  SFB_ERR_CUDA = 3,             /* a CUDA runtime call failed; see sfb_last_error_message */
  SFB_ERR_UNSUPPORTED_SIZE = 4, /* (n, m) does not fit the shared-memory resident kernels */
  SFB_ERR_MIXED_MEMORY = 5,     /* host and device pointers mixed in one call */
  SFB_ERR_OUT_OF_MEMORY = 6
} sfb_error;

/* QPSolutionStatus, qp.hpp:82-92 -- values and order are part of the contract */
typedef enum {
  SFB_QP_OPTIMAL = 0,
  SFB_QP_POLISH_FAILED = 1, /* never produced, exactly like the reference (qp_solver.hpp:537 is overwritten at :544) */
  SFB_QP_PRIMAL_INFEASIBLE = 2,
  SFB_QP_DUAL_INFEASIBLE = 3,
  SFB_QP_MAX_ITERATIONS = 4,
  SFB_QP_MAX_TIME = 5,
  SFB_QP_UNKNOWN = 6
} sfb_qp_status;

/*
 * QPSolverParams, qp_solver.hpp:29-68, field for field.  The float members stay float: the reference
 * casts them to Scalar (rho = (double)0.1f, not 0.1) and that is part of the algorithm.
 */
typedef struct {
  int32_t verbose;          /* accepted, ignored on device */
  float alpha;              /* 1.6f   relaxation */
  float rho;                /* 0.1f   first dual step size */
  float sigma;              /* 1e-6f  second dual step size */
  int32_t scaling;          /* 1      Ruiz-style equilibration (qp_solver.hpp:673-730) */
  float eps_abs;            /* 1e-3f */
  float eps_rel;            /* 1e-3f */
  float eps_primal_inf;     /* 1e-4f */
  float eps_dual_inf;       /* 1e-4f */
  int32_t has_max_iter;     /* 0      std::optional engaged? */
  uint32_t max_iter;        /*        if !has_max_iter the device loop is still bounded by SFB_QP_DEVICE_ITER_CAP */
  int32_t has_max_time;     /* 0 */
  int64_t max_time_ns;      /*        per-instance budget measured with the device clock, checked at stop checks */
  uint32_t stop_check_iter; /* 25     checks fire when iter % stop_check_iter == 1 (qp_solver.hpp:465,479) */
  int32_t polish;           /* 1 */
  uint32_t polish_iter;     /* 5 */
  float delta;              /* 1e-6f  polish regularisation */
} sfb_qp_params;

/* an unbounded device loop would let one diverging instance hold a warp forever */
#define SFB_QP_DEVICE_ITER_CAP 1000000u

/* out_flags bits (optional diagnostics, not part of the reference's observable state) */
#define SFB_QP_FLAG_POLISHED 1u       /* polish ran and its result was accepted */
#define SFB_QP_FLAG_POLISH_SKIPPED 2u /* active set too large for the on-chip polish workspace: solution left unpolished,
                                         which is also what the reference returns when its polish fails */
#define SFB_QP_FLAG_POLISH_FAILED 4u  /* non-positive / non-finite pivot in the polish systems */
#define SFB_QP_FLAG_POLISH_SCRATCH 8u /* (with POLISHED) the polish Schur block lived in the global workspace, not on chip */
#define SFB_QP_FLAG_POLISH_REDUCED 16u /* (with POLISHED, dense kernels) the polish system was solved in its reduced n x n form */

int sfb_version(void);
const char* sfb_error_string(int err);
/* message of the last failing call on this handle (or of sfb_create when h == NULL); never NULL */
const char* sfb_last_error_message(sfb_handle_t h);

/* stream: a cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream) or NULL for the default stream */
int sfb_create(int device, void* stream, sfb_handle_t* out);
int sfb_destroy(sfb_handle_t h);
int sfb_set_stream(sfb_handle_t h, void* stream);
int sfb_synchronize(sfb_handle_t h);

/*
 * Engine options (not part of QPSolverParams, which sfb_qp_params mirrors field for field).
 *   SFB_OPT_DUAL_INF_DX_GUARD (default 1): the dual-infeasibility certificate of check_stopping (qp_solver.hpp:625-641)
 *     additionally requires ||dx|| != 0.  With an exactly stationary primal iterate every comparison of :629-639 reads
 *     0 <= 0 and the literal rule reports DualInfeasible.  In the reference that only happens on scalar (n = 1) problems
 *     (oracle instrumentation, tests/test_oracle_qp_known_answers.py); the engine's reduced KKT system reaches exactly
 *     stationary iterates on tall problems (n = 3, m = 203) where the reference does not, so the guard is what keeps the
 *     engine's statuses equal to the reference's there.  0 = the literal rule.
 */
#define SFB_OPT_DUAL_INF_DX_GUARD 1
/*   SFB_OPT_FORCE_POLISH_SCRATCH (default 0, debug / tests): dense polish keeps its Schur block in the global workspace even
 *     when it fits in shared memory, so that both placements can be compared on the same problems. */
#define SFB_OPT_FORCE_POLISH_SCRATCH 2
/*   SFB_OPT_POLISH_FORM (default 0): how the dense kernel eliminates the regularised polish system of polish_qp
 *     (qp_solver.hpp:160-195).  0: duals first (one n x n inverse of Pbar + delta I + Aa^T Aa / delta), refined with literal
 *     residual sweeps until the correction is below 1e-8 of the solution, and redone in the Schur form if polish_iter sweeps
 *     do not get there; 1: always the Schur form (primal block first, Eigen's pivot order; the round-1 / early round-2
 *     behaviour); 2: always the reduced form (A/B measurements). */
#define SFB_OPT_POLISH_FORM 3
int sfb_set_option(sfb_handle_t h, int option, int value);
/* number of kernels of this library launched through the handle since creation */
int sfb_kernel_launch_count(sfb_handle_t h, uint64_t* out);

void sfb_qp_params_default(sfb_qp_params* p);

/*
 * Replaces QPSolver<QuadraticProgram<M,N,double>>::solve / solve_qp (qp_solver.hpp:343-568, :779-787) for a
 * batch of independent dense problems of one shape.
 *
 *   warm_x [batch][n], warm_y [batch][m]  both NULL (cold) or both given      (qp_solver.hpp:436-445)
 *   out_x [batch][n], out_y [batch][m], out_obj [batch]                        QPSolution::{primal,dual,objective}
 *   out_status [batch] int32 (sfb_qp_status), out_iter [batch] uint32          QPSolution::{code,iter}
 *   out_active [batch][m] int8, may be NULL: -1 lower-active, +1 upper-active, 0 inactive -- the sets
 *       polish_qp builds (qp_solver.hpp:113-123) evaluated on the scaled dual when the ADMM loop ends
 *   out_flags [batch] uint32, may be NULL: SFB_QP_FLAG_* diagnostics
 *
 * Kernel selection is internal and does not change results beyond rounding: problems that fit in shared memory run the
 * group kernel (1, 2 or 4 warps per instance); tall-skinny problems without polish (n <= 4, m <= 256 -- the shape and the
 * setting of ASIFilter's call, asif.hpp:97) run a warp-per-instance kernel that keeps the working set in registers.
 */
int sfb_qp_solve_dense_batch_f64(sfb_handle_t h, const sfb_qp_params* prm, int64_t batch, int n, int m,
                                 const double* P, const double* q, const double* A, const double* l,
                                 const double* u, const double* warm_x, const double* warm_y, double* out_x,
                                 double* out_y, double* out_obj, int32_t* out_status, uint32_t* out_iter,
                                 int8_t* out_active, uint32_t* out_flags);

/* Same in single precision (new functionality: the reference has no float instantiation): ADMM iterations in
 * fp32; polish_qp (delta = 1e-6 is 8 ulp in fp32) runs as a mixed-precision second pass in fp64 on the fp32 iterate and active
 * set (SFB_QP_FLAG_POLISHED); shapes whose fp64 working set does not fit in shared memory keep SFB_QP_FLAG_POLISH_SKIPPED. */
int sfb_qp_solve_dense_batch_f32(sfb_handle_t h, const sfb_qp_params* prm, int64_t batch, int n, int m,
                                 const float* P, const float* q, const float* A, const float* l, const float* u,
                                 const float* warm_x, const float* warm_y, float* out_x, float* out_y,
                                 float* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active,
                                 uint32_t* out_flags);

/* Largest m (for a given n) the shared-memory resident dense kernel accepts; 0 if n itself is too large. */
int sfb_qp_dense_max_m(sfb_handle_t h, int n, int scalar_bytes);

/* QPSolver::scale alone (qp_solver.hpp:673-730): out_c [batch], out_sx [batch][n], out_sy [batch][m]. Device pointers only. */
int sfb_qp_scale_dense_batch_f64(sfb_handle_t h, int64_t batch, int n, int m, const double* P, const double* q,
                                 const double* A, double* out_c, double* out_sx, double* out_sy);

typedef enum { SFB_STEPPER_EULER = 0, SFB_STEPPER_RK4 = 1 } sfb_stepper;

/*
 * Replaces the covariance half of EKF::predict (ekf.hpp:79-103): integrates Pdot = symU(A P + P A^T + Q)
 * over [0, tau] with the reference's step schedule (dt <= 0 -> one step of length tau, ekf.hpp:92-102).
 * A = -ad(f) + d^r f/dx is evaluated by the caller at the pre-step estimate (user lambdas + autodiff stay
 * on the host) and held constant over the call.  P, A, Q, out_P: [batch][d*d] column-major.  1 <= d <= 16.
 * All-device pointers (asynchronous on the handle's stream) or all-host pointers (staged, returns with the result in place),
 * like every entry point of this header.
 */
int sfb_ekf_predict_batch_f64(sfb_handle_t h, int64_t batch, int d, int stepper, const double* P,
                              const double* A, const double* Q, double tau, double dt, double* out_P);

/*
 * Replaces the algebra of EKF::update (ekf.hpp:116-139): S = triu(H symU(P) H^T + R), K = (S^-1 H P)^T,
 * out_delta = K * innov (the caller applies g_hat (+) delta), out_P = symU((I - K H) P).
 * H [batch][ny*d] (H_ij at i + ny*j), R [batch][ny*ny], innov = y (-) h(g_hat) [batch][ny].  1 <= d, ny <= 16.
 */
int sfb_ekf_update_batch_f64(sfb_handle_t h, int64_t batch, int d, int ny, const double* P, const double* H,
                             const double* R, const double* innov, double* out_delta, double* out_P);

/*
 * One EKF cycle of the covariance in a single pass over HBM: predict (euler stepper, the reference default,
 * ekf.hpp:147) followed by update, i.e. sfb_ekf_predict_batch_f64 + sfb_ekf_update_batch_f64 without the round
 * trip of the predicted covariance through memory.  H and innov are evaluated by the caller at the PREDICTED
 * estimate g_hat (+) tau f (ekf.hpp:96,119 -- they do not depend on P).  Arguments as in the two calls above;
 * stepper must be SFB_STEPPER_EULER or SFB_STEPPER_RK4 (RK4 runs the two generic kernels back to back).
 * out_P may alias P.
 */
int sfb_ekf_step_batch_f64(sfb_handle_t h, int64_t batch, int d, int ny, int stepper, const double* P,
                           const double* A, const double* Q, double tau, double dt, const double* H,
                           const double* R, const double* innov, double* out_delta, double* out_P);

/*
 * ---- sparse problems with a SHARED sparsity pattern (QuadraticProgramSparse, qp.hpp:60-79) ---------------------
 *
 * sfb_qp_sparse_analyze replaces SimplicialLDLT::analyzePattern (call site qp_solver.hpp:424, done once per solver
 * object -- MPC does it in its constructor, mpc.hpp:424): fill-reducing ordering, symbolic factor and assembly
 * schedules of the reduced KKT matrix, uploaded to the handle's device.  The pattern arrays are HOST pointers in
 * Eigen's compressed storage:  P column-major (SparseMatrix<double>: outerIndexPtr = P_colptr [n+1], innerIndexPtr
 * = P_rowidx), A row-major (SparseMatrix<double, RowMajor>: A_rowptr [m+1], A_colidx).  Explicit zeros are part of the
 * pattern.  Like the reference, only entries of P with col >= row enter the factorisation (qp_solver.hpp:384) while
 * scaling, residuals and the objective use the entries as stored.
 */
typedef struct sfb_qp_sparse_pattern* sfb_qp_sparse_pattern_t;

int sfb_qp_sparse_analyze(sfb_handle_t h, int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx,
                          const int32_t* A_rowptr, const int32_t* A_colidx, sfb_qp_sparse_pattern_t* out);
int sfb_qp_sparse_pattern_destroy(sfb_qp_sparse_pattern_t p);
/* The symbolic analysis alone (host only, no device needed): nnz(L), multiply-adds per factorisation, the ordering
 * perm_out [n] (perm[new] = old) and the column pointers of L, L_colptr_out [n+1]; outputs may be NULL. */
int sfb_qp_sparse_symbolic(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                           const int32_t* A_colidx, int64_t* nnz_L, int64_t* factor_flops, int32_t* perm_out,
                           int32_t* L_colptr_out);
/* Host-only self-check of the schedules of the on-chip (one CTA per instance) sparse kernel: analyses the pattern with
 * ordering = -1 (cost model picks), 0 (minimum degree) or 1 (nested dissection), executes assembly, supernodal
 * factorisation, diagonal-block inversion and the staged triangular sweeps on the host on seeded random values and compares
 * the solution of (L D L^T) x = b with a dense Cholesky solve.  info_out [8] = {supernodes, levels, factor slots nW,
 * structural nnz(L), multiply-adds, largest supernode, sweep stages, ordering used}; max_rel_err_out = the largest relative error. */
int sfb_qp_sparse_cta_selfcheck(int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx, const int32_t* A_rowptr,
                                const int32_t* A_colidx, int ordering, int64_t* info_out, double* max_rel_err_out);
/* 1 if sfb_qp_solve_sparse_batch_* on this handle runs the on-chip kernel (one CTA per instance, working set in shared memory)
 * for problems of scalar_bytes = 4 (fp32) or 8 (fp64) with this pattern, 0 if it runs the HBM-tiled kernel (working set too large for
 * shared memory, or SFB_SPARSE_KERNEL=tiled in the environment when the handle was created).  info_out [8] (may be NULL) = the
 * on-chip analysis: {supernodes, levels, factor slots, structural nnz(L), multiply-adds, largest supernode, sweep stages,
 * shared-memory bytes per CTA}. */
int sfb_qp_sparse_uses_onchip(sfb_handle_t h, sfb_qp_sparse_pattern_t p, int scalar_bytes, int64_t* info_out);
/* nnz of the strictly lower factor L, multiply-adds of one numeric factorisation, ordering (perm_out [n], may be NULL) */
int sfb_qp_sparse_pattern_info(sfb_qp_sparse_pattern_t p, int64_t* nnz_L, int64_t* factor_flops, int32_t* perm_out);

/*
 * Replaces QPSolver<QuadraticProgramSparse<double>>::solve (qp_solver.hpp:343-568, sparse branches; call site
 * mpc.hpp:491) for a batch of problems sharing `pattern`:
 *   P_vals [batch][nnzP], A_vals [batch][nnzA]  the valuePtr() arrays of each instance, in pattern order
 *   q [batch][n], l,u [batch][m]; warm starts and outputs exactly as in sfb_qp_solve_dense_batch_f64.
 * All-host or all-device pointers.
 */
int sfb_qp_solve_sparse_batch_f64(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                  int64_t batch, const double* P_vals, const double* q, const double* A_vals,
                                  const double* l, const double* u, const double* warm_x, const double* warm_y,
                                  double* out_x, double* out_y, double* out_obj, int32_t* out_status,
                                  uint32_t* out_iter, int8_t* out_active, uint32_t* out_flags);
/*
 * OSQP-style ingestion (compat/osqp.hpp:36-49 hands OSQP both matrices in CSC): the same problems with A
 * given column-compressed -- A_colptr [n+1], A_rowidx [nnzA], values per instance in that order.  The analysis converts the
 * pattern to the solver's CSR order once and keeps the value permutation; results are identical to the CSR entry point.
 */
int sfb_qp_sparse_analyze_csc(sfb_handle_t h, int n, int m, const int32_t* P_colptr, const int32_t* P_rowidx,
                              const int32_t* A_colptr, const int32_t* A_rowidx, sfb_qp_sparse_pattern_t* out);
int sfb_qp_solve_sparse_batch_csc_f64(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                      int64_t batch, const double* P_vals, const double* q, const double* A_vals_csc,
                                      const double* l, const double* u, const double* warm_x, const double* warm_y,
                                      double* out_x, double* out_y, double* out_obj, int32_t* out_status,
                                      uint32_t* out_iter, int8_t* out_active, uint32_t* out_flags);
/* single precision (new functionality): ADMM iterations in fp32, polish_qp as a mixed-precision second pass in fp64 */
int sfb_qp_solve_sparse_batch_f32(sfb_handle_t h, sfb_qp_sparse_pattern_t pattern, const sfb_qp_params* prm,
                                  int64_t batch, const float* P_vals, const float* q, const float* A_vals,
                                  const float* l, const float* u, const float* warm_x, const float* warm_y, float* out_x,
                                  float* out_y, float* out_obj, int32_t* out_status, uint32_t* out_iter,
                                  int8_t* out_active, uint32_t* out_flags);

/*
 * ---- ASIFilter on the device for the built-in SE(2) x R^3 vehicle family (SURVEY 8(f) rows f1 + f3) -----------------
 *
 * Replaces ASIFilter<G, U, Dyn>::operator()(g, u_des, h, bu) (asif.hpp:82-102) -- asif_to_qp_update (asif_func.hpp:104-199)
 * followed by solve_qp (asif.hpp:97) and the warm-start retention rule (asif.hpp:99: kept only if Optimal) -- for a FLEET of
 * agents that share one parameter set, for the model family of examples/mpc_asif_vehicle.cpp:
 *     G = Bundle<SE2, R^3>, coefficients (x, y, sin, cos, v1, v2, v3);  U = R^2
 *     d^r g = (v1, v2, v3, -drag1 v1 + u1, 0, -drag3 v3 + u2)                  (:42-52)
 *     h(g)  = |p - centre| - radius                (nh = 1)                     (:96-100)
 *     bu(g) = (bu_gain v1, bu_const)                                            (:103)
 * The reference differentiates user lambdas with autodiff on the host (they cannot cross a C ABI); for this family the
 * derivatives are closed forms evaluated on the device.  One kernel launch maps (g, u_des) -> u; the QP (n = 3, m = K + 3)
 * never exists in host memory and the warm starts stay resident on the device between control steps.
 * K + 3 <= 256 (the register-resident tall-skinny solver); qp.polish must be 0 (as in mpc_asif_vehicle.cpp:127).
 */
typedef struct {
  double T;             /* 2.5    ASIFilterParams::T        asif.hpp:20   look-ahead horizon */
  int32_t K;            /* 200    ASIFtoQPParams::K         asif_func.hpp:61 */
  double alpha;         /* 5      ASIFtoQPParams::alpha     :63 */
  double dt;            /* 0.01   ASIFtoQPParams::dt        :65 */
  double relax_cost;    /* 100    ASIFtoQPParams::relax_cost :67 */
  double u_weight[2];   /* 20, 1  ASIFilterParams::u_weight asif.hpp:24 */
  double ulim_l[2];     /* -0.2, -0.5   ManifoldBounds with A = I, c = 0 (mpc_asif_vehicle.cpp:105-110) */
  double ulim_u[2];     /* 0.5, 0.5 */
  double drag1, drag3;  /* 0.2, 0.4 */
  double centre[2];     /* 0, -2.3 */
  double radius;        /* 0.7 */
  double bu_gain;       /* 0.2 */
  double bu_const;      /* -0.5 */
  sfb_qp_params qp;     /* ASIFilterParams::qp; defaults + polish = 0 */
} sfb_asif_vehicle_params;

/* the parameter set of examples/mpc_asif_vehicle.cpp:96-129 */
void sfb_asif_vehicle_params_default(sfb_asif_vehicle_params* p);

typedef struct sfb_asif_fleet* sfb_asif_fleet_t;

/* scalar_bytes: 8 (the reference's arithmetic) or 4.  The trajectory / sensitivity integration always runs in fp64. */
int sfb_asif_fleet_create(sfb_handle_t h, const sfb_asif_vehicle_params* p, int64_t batch, int scalar_bytes,
                          sfb_asif_fleet_t* out);
int sfb_asif_fleet_destroy(sfb_asif_fleet_t f);
/* forget every agent's warm start (a freshly constructed ASIFilter, asif.hpp:109) */
int sfb_asif_fleet_reset_warmstart(sfb_asif_fleet_t f);
/* warm = 0: ignore and do not update the resident warm starts (every call is a cold solve_qp) */
int sfb_asif_fleet_set_warmstart(sfb_asif_fleet_t f, int warm);
/*
 * One control step of every agent: x [batch][7], u_des [batch][2] -> out_u [batch][2] = u_des + primal.head<2>(),
 * out_status [batch] (sfb_qp_status), out_iter [batch].  All-host or all-device pointers (host: 72 B in, 24 B out per
 * agent cross PCIe per step).
 */
int sfb_asif_fleet_filter_f64(sfb_asif_fleet_t f, const double* x, const double* u_des, double* out_u,
                              int32_t* out_status, uint32_t* out_iter);
int sfb_asif_fleet_filter_f32(sfb_asif_fleet_t f, const float* x, const float* u_des, float* out_u, int32_t* out_status,
                              uint32_t* out_iter);
/*
 * asif_to_qp (asif_func.hpp:246-261) alone: the dense QP of every agent in the reference's storage,
 * P [batch][9], q [batch][3], A [batch][(K+3)*3] column-major, l, u [batch][K+3].  fp64 fleets only.
 */
int sfb_asif_fleet_to_qp_f64(sfb_asif_fleet_t f, const double* x, const double* u_des, double* P, double* q, double* A,
                             double* l, double* u);

/*
 * ---- MPC on the device for the built-in SE(2) x R^3 vehicle family (SURVEY 8(f) rows f2 + f3) -----------------------
 *
 * Replaces MPC<T, X, U, F, CR, Kmesh>::operator()(t, x) (mpc.hpp:458-519) -- ocp_to_qp_update_dyn / _ce (ocp_to_qp.hpp:
 * 198-276, 326-373), makeCompressed, qp_solver_.solve(qp_, warmstart_) (:491), the warm-start retention rule (:510-516:
 * kept if Optimal, MaxTime or MaxIterations) and u = udes(0) (+) primal.segment<Nu>(uvar_B) (:518) -- for a FLEET of agents
 * that share one parameter set, for the model family of examples/mpc_asif_vehicle.cpp:42-89:
 *     X = Bundle<SE2, R^3>, coefficients (x, y, sin, cos, v1, v2, v3);  U = R^2
 *     d^r x = (v1, v2, v3, -drag1 v1 + u1, 0, -drag3 v3 + u2),  running constraint crl <= u <= cru
 *     xdes(t) = X{SE2(g0) * exp(t vdes), vdes}  (mpc_asif_vehicle.cpp:72-78),  udes(t) = udes
 * The sparse QP (K = 50: n = m = 422) is transcribed on the device straight into the value layout of
 * sfb_qp_solve_sparse_batch_* (pattern = QuadraticProgramSparse after makeCompressed); symbolic analysis is done once at
 * creation (MPC's constructor, mpc.hpp:424); warm starts stay resident on the device.  Per control step only (t, x) of
 * every agent go in and (u, code, iter) come out.
 * The cost is transcribed at creation from Q, R, Qtf exactly as the reference's constructor does (mpc.hpp:423); the
 * reference never re-transcribes it, so MPC::set_weights after construction has no effect there either.
 */
typedef struct {
  int32_t K;         /* 50   MPCParams::K       mpc.hpp:317   -> ceil(K / Kmesh) mesh intervals */
  double tf;         /* 5    MPCParams::tf      :322 */
  int32_t Kmesh;     /* 4    template parameter Kmesh of MPC (collocation points per interval) */
  int32_t warmstart; /* 1    MPCParams::warmstart :327 */
  double Q[6];       /* 1..  MPCWeights (diagonal), in effect at construction */
  double R[2];
  double Qtf[6];
  double crl[2];     /* -0.5, -0.5 */
  double cru[2];     /*  0.5,  0.5 */
  double drag1, drag3; /* 0.2, 0.4 */
  double g0[3];      /* 2.5, 0, pi/2   SE(2) element (x, y, angle) the desired trajectory starts from */
  double vdes[3];    /* 1, 0, 0.4      its constant body velocity = the desired R^3 part */
  double udes[2];    /* 0, 0 */
  sfb_qp_params qp;  /* MPCParams::qp (defaults: polish on) */
} sfb_mpc_vehicle_params;

/* the parameter set of examples/mpc_asif_vehicle.cpp:42-89 with K = 50 (BASELINE configs[2]); weights as the reference
 * actually uses them (identity: its set_weights call comes after the cost was transcribed) */
void sfb_mpc_vehicle_params_default(sfb_mpc_vehicle_params* p);

typedef struct sfb_mpc_fleet* sfb_mpc_fleet_t;

int sfb_mpc_fleet_create(sfb_handle_t h, const sfb_mpc_vehicle_params* p, int64_t batch, int scalar_bytes,
                         sfb_mpc_fleet_t* out);
int sfb_mpc_fleet_destroy(sfb_mpc_fleet_t f);
int sfb_mpc_fleet_reset_warmstart(sfb_mpc_fleet_t f); /* MPC::reset_warmstart, mpc.hpp:611 */
/* QP sizes: n, m, nnz(P), nnz(A), nnz of the factor; any output may be NULL */
int sfb_mpc_fleet_dims(sfb_mpc_fleet_t f, int* n, int* m, int* nnzP, int* nnzA, int64_t* nnzL);
/* the shared sparsity pattern (host arrays): P_colptr [n+1], P_rowidx [nnzP], A_rowptr [m+1], A_colidx [nnzA] */
int sfb_mpc_fleet_pattern(sfb_mpc_fleet_t f, int32_t* P_colptr, int32_t* P_rowidx, int32_t* A_rowptr, int32_t* A_colidx);
/* the transcription alone: t [batch], x [batch][7] -> P_vals [batch][nnzP], q [batch][n], A_vals [batch][nnzA], l, u [batch][m] */
int sfb_mpc_fleet_to_qp_f64(sfb_mpc_fleet_t f, const double* t, const double* x, double* P_vals, double* q,
                            double* A_vals, double* l, double* u);
/*
 * One control step of every agent: t [batch] (absolute time), x [batch][7] -> out_u [batch][2], out_status [batch],
 * out_iter [batch]; out_primal [batch][n] / out_dual [batch][m] may be NULL.  All-host or all-device pointers.
 */
int sfb_mpc_fleet_step_f64(sfb_mpc_fleet_t f, const double* t, const double* x, double* out_u, int32_t* out_status,
                           uint32_t* out_iter, double* out_primal, double* out_dual);
int sfb_mpc_fleet_step_f32(sfb_mpc_fleet_t f, const float* t, const float* x, float* out_u, int32_t* out_status,
                           uint32_t* out_iter, float* out_primal, float* out_dual);
/*
 * The optional outputs of MPC::operator() (mpc.hpp:453-454, :493-507), computed from the solution the LAST step left on
 * the device: out_u_traj [batch][N][2], u_traj[i] = udes(t + tf tau_i) + primal.segment<Nu>(uvar_B + i Nu);
 * out_x_traj [batch][N + 1][7], x_traj[i] = xdes(t + tf tau_i) (+) primal.segment<Nx>(i Nx).  t [batch] must be the times
 * given to that step; either output may be NULL.  sfb_mpc_fleet_nodes: N = Mesh::N_colloc() and tau [N + 1] =
 * Mesh::all_nodes() (collocation/mesh.hpp:213-224), either may be NULL.
 */
int sfb_mpc_fleet_nodes(sfb_mpc_fleet_t f, int* N, double* tau);
int sfb_mpc_fleet_trajectories_f64(sfb_mpc_fleet_t f, const double* t, double* out_u_traj, double* out_x_traj);
int sfb_mpc_fleet_trajectories_f32(sfb_mpc_fleet_t f, const float* t, float* out_u_traj, float* out_x_traj);

/*
 * ---- multi-GPU: the one collective of the path -----------------------------------------------------
 *
 * Instances are independent: every rank (one process per GPU) solves a contiguous shard, inputs are never exchanged, and
 * the results of all shards are made visible everywhere by ONE all-gather over NCCL (NVLink 5 / NVSwitch).  The reference
 * is a single-instance library and has no counterpart; this is what a C / C++ host of the engine binds for 8 x B200.
 *
 *   sfb_comm_unique_id    rank 0 creates the 128-byte NCCL id and distributes it (MPI, a file, torch.distributed ...)
 *   sfb_comm_create       ncclCommInitRank on the handle's device (collective: every rank calls it)
 *   sfb_allgather_results the `count` result arrays of this rank's shard (device pointers; send[k] holds bytes_per_rank[k]
 *                         bytes, recv[k] world * bytes_per_rank[k]: rank r's block at offset r * bytes_per_rank[k]) are
 *                         gathered as ONE grouped NCCL operation -- no packing kernel.  It is enqueued on the
 *                         communicator's own stream, ordered after the work already enqueued on the handle's stream, and
 *                         returns immediately: the next solve (into other output buffers) overlaps the exchange.
 *   sfb_comm_wait         host_sync = 0: the handle's stream waits for the exchange; 1: the host does.
 * NCCL is loaded at run time (libnccl.so.2); without it these calls fail with SFB_ERR_CUDA, nothing else is affected.
 */
#define SFB_COMM_ID_BYTES 128
typedef struct sfb_comm* sfb_comm_t;
int sfb_comm_unique_id(void* out128);
int sfb_comm_create(sfb_handle_t h, int world, int rank, const void* id128, sfb_comm_t* out);
int sfb_comm_destroy(sfb_comm_t c);
int sfb_allgather_results(sfb_comm_t c, int count, const void* const* send, void* const* recv, const size_t* bytes_per_rank);
int sfb_comm_wait(sfb_comm_t c, int host_sync);

#ifdef __cplusplus
}
#endif
#endif /* SFB_H */
