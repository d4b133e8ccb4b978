/*
 * sf_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * Plain restatement of the smooth_feedback numerical hot path, used ONLY as the
 * checker in tests/, in __graft_entry__.smoke() and as bench.py's cpu_baseline /
 * --impl reference leg.  Nothing under smooth_feedback_b200/ may link, import or
 * call this library.
 *
 * Restates (file:line are into /root/reference):
 *   include/smooth/feedback/qp_solver.hpp:297-730   QPSolver::{analyze,solve,check_stopping,scale}
 *   include/smooth/feedback/qp_solver.hpp:92-204    detail::polish_qp
 *   include/smooth/feedback/qp.hpp:31-108           QuadraticProgram / QPSolution / QPSolutionStatus
 *   include/smooth/feedback/ekf.hpp:79-139          EKF::predict / EKF::update (dense algebra part)
 * plus the third-party pieces that are absent from /root/reference:
 *   Eigen 3.4.0  LDLT<Matrix,Upper>::compute / solveInPlace  (diagonal-pivoted LDL^T, see
 *                oracle/sf_oracle.cpp::ldlt_compute for the published algorithm restated)
 *   Boost.odeint euler / runge_kutta4 with vector_space_algebra (classical one-step formulas)
 *
 * PARITY PIN: the reference cannot be compiled in this image (Eigen, Boost, smooth absent).
 * The oracle is pinned by the reference's own known-answer tests, transliterated in
 * tests/test_oracle_qp_known_answers.py and tests/test_oracle_ekf.py (tests/test_qp.cpp:54-336,
 * tests/test_ekf.cpp:50-180).  Pivot order / rounding of Eigen's LDLT itself is "parity unpinned".
 */
#ifndef SF_ORACLE_H
#define SF_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* mirrors QPSolverParams (qp_solver.hpp:29-68); float members stay float on purpose */
typedef struct {
  float alpha;            /* 1.6f  */
  float rho;              /* 0.1f  */
  float sigma;            /* 1e-6f */
  int32_t scaling;        /* 1     */
  float eps_abs;          /* 1e-3f */
  float eps_rel;          /* 1e-3f */
  float eps_primal_inf;   /* 1e-4f */
  float eps_dual_inf;     /* 1e-4f */
  int32_t has_max_iter;   /* 0     */
  uint32_t max_iter;
  uint32_t stop_check_iter; /* 25  */
  int32_t polish;         /* 1     */
  uint32_t polish_iter;   /* 5     */
  float delta;            /* 1e-6f */
} sfo_qp_params;

void sfo_qp_params_default(sfo_qp_params* p);

/* test instrumentation: number of stop checks so far that saw an exactly stationary primal iterate (||dx_us|| == 0);
 * reset != 0 zeroes the counter after reading it */
long long sfo_debug_dx_zero_checks(int reset);

/*
 * Batched dense QP solve, fp64.  Layout per instance b (all column-major like Eigen defaults):
 *   P + b*n*n, q + b*n, A + b*m*n (A[i + m*j]), l + b*m, u + b*m
 * warm_x / warm_y may be NULL (cold start).  out_active[b*m + i] in {-1,0,+1} is the active set as
 * polish_qp defines it (qp_solver.hpp:113-123) evaluated on the scaled dual before polishing.
 * nthreads <= 1 -> serial.  Returns 0.
 */
int sfo_qp_solve_dense_batch_f64(const sfo_qp_params* prm, int64_t batch, int n, int m, const double* P,
                                 const double* q, const double* A, const double* l, const double* u,
                                 const double* warm_x, const double* warm_y, double* out_x, double* out_y,
                                 double* out_obj, int32_t* out_status, uint32_t* out_iter, int8_t* out_active,
                                 int nthreads);

/* debug: expose the scaling (c, sx[n], sy[m]) the solver computed for one instance */
int sfo_qp_scale_f64(int n, int m, const double* P, const double* q, const double* A, double* c, double* sx,
                     double* sy);

/*
 * EKF covariance propagation (ekf.hpp:79-103) with A held constant over the call:
 *   stepper 0 = euler, 1 = runge_kutta4; dt <= 0 -> reference default dt = 2*tau (one step of tau).
 * P, A, Q, out_P: [batch][d*d] column-major.
 */
int sfo_ekf_predict_batch_f64(int64_t batch, int d, int stepper, const double* P, const double* A,
                              const double* Q, double tau, double dt, double* out_P, int nthreads);

/*
 * EKF measurement update (ekf.hpp:116-139): H [batch][ny*d] col-major (H[i + ny*j]), R [batch][ny*ny],
 * innov = y (-) h(g_hat) [batch][ny].  Outputs delta = K*innov [batch][d] and out_P [batch][d*d].
 */
int sfo_ekf_update_batch_f64(int64_t batch, int d, int ny, const double* P, const double* H, const double* R,
                             const double* innov, double* out_delta, double* out_P, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
