"""Pins the CPU oracle's EKF algebra with the reference's closed-form tests (tests/test_ekf.cpp:50-180).

CPU only.  State propagation (g_hat (+) tau*f) is host-side group arithmetic and is covered in tests/cpp/test_ekf_overlay.cpp (the same tests through the C++ EKF overlay).
"""
import numpy as np
import pytest
from scipy.linalg import expm

from qp_cases import is_approx


@pytest.mark.parametrize("nx,ny", [(3, 3), (10, 3), (3, 10)])
def test_update_linear(oracle, nx, ny):
    # test_ekf.cpp:50-103: random linear h, diagonal P,R in [0.1, 2.1]; equals the textbook KF at 1e-6
    rng = np.random.default_rng(nx * 100 + ny)
    B = 20
    x = rng.uniform(-1, 1, (B, nx)); xhat = rng.uniform(-1, 1, (B, nx))
    P = np.zeros((B, nx, nx)); R = np.zeros((B, ny, ny))
    P[:, np.arange(nx), np.arange(nx)] = rng.uniform(-1, 1, (B, nx)) + 1.1
    R[:, np.arange(ny), np.arange(ny)] = rng.uniform(-1, 1, (B, ny)) + 1.1
    H = rng.uniform(-1, 1, (B, ny, nx)); h = rng.uniform(-1, 1, (B, ny))
    ymeas = np.einsum("bij,bj->bi", H, x) + h
    innov = ymeas - (np.einsum("bij,bj->bi", H, xhat) + h)
    delta, Pn = oracle.ekf_update_batch(P, H, R, innov)
    for b in range(B):
        S = H[b] @ P[b] @ H[b].T + R[b]
        K = P[b] @ H[b].T @ np.linalg.inv(S)
        assert is_approx(xhat[b] + K @ innov[b], xhat[b] + delta[b], 1e-6)
        assert is_approx((np.eye(nx) - K @ H[b]) @ P[b], Pn[b], 1e-6)


@pytest.mark.parametrize("nx", [3, 6, 9])
def test_predict_linear_rk4(oracle, nx):
    # test_ekf.cpp:105-153: xdot = A x, Q = 0, tau = 0.7, RK4 dt = 1e-3; P_new = F P F^T with F = expm(A tau) at 1e-3
    rng = np.random.default_rng(nx)
    B = 10
    A = rng.uniform(-1, 1, (B, nx, nx))
    P = np.zeros((B, nx, nx))
    P[:, np.arange(nx), np.arange(nx)] = rng.uniform(-1, 1, (B, nx)) + 1.1
    Q = np.zeros((B, nx, nx))
    tau = 0.7
    Pn = oracle.ekf_predict_batch(P, A, Q, tau, dt=1e-3, stepper="rk4")
    for b in range(B):
        F = expm(A[b] * tau)
        assert is_approx(F @ P[b] @ F.T, Pn[b], 1e-3)


def test_predict_time_cut_euler(oracle):
    # test_ekf.cpp:155-180: tau = 0.7, dt = 0.5 -> two Euler steps 0.5 + 0.2 (ekf.hpp:93-102).
    # For xdot = b the Jacobian A is 0, so Pdot = Q and Euler is exact: P + tau * symU(Q).
    rng = np.random.default_rng(7)
    P = np.diag(rng.uniform(-1, 1, 2) + 1.1)[None]
    Q = np.diag(rng.uniform(-1, 1, 2) + 1.1)[None]
    A = np.zeros((1, 2, 2))
    Pn = oracle.ekf_predict_batch(P, A, Q, 0.7, dt=0.5, stepper="euler")
    assert np.allclose(Pn[0], P[0] + 0.7 * Q[0], rtol=1e-14)


def test_predict_default_is_one_euler_step(oracle):
    # default dt = 2*tau -> the while loop is skipped, exactly one step of length tau (ekf.hpp:92-102);
    # Euler gives P + tau*(AP + PA^T + Q), NOT F P F^T + tau Q
    rng = np.random.default_rng(3)
    d = 6
    A = rng.normal(size=(4, d, d)); M = rng.uniform(-1, 1, (4, d, d))
    P = M @ np.transpose(M, (0, 2, 1)) + 0.1 * np.eye(d)
    Q = 0.01 * np.eye(d)[None].repeat(4, 0)
    tau = 0.1
    Pn = oracle.ekf_predict_batch(P, A, Q, tau)
    ref = P + tau * (A @ P + P @ np.transpose(A, (0, 2, 1)) + Q)
    assert np.allclose(Pn, ref, rtol=1e-13, atol=1e-15)


def test_only_upper_triangle_of_q_is_used(oracle):
    rng = np.random.default_rng(4)
    d = 3
    A = rng.normal(size=(1, d, d)); P = np.eye(d)[None] * 2.0
    Q = rng.normal(size=(1, d, d))
    Qsym = np.triu(Q[0]) + np.triu(Q[0], 1).T
    a = oracle.ekf_predict_batch(P, A, Q, 0.3)
    b = oracle.ekf_predict_batch(P, A, Qsym[None], 0.3)
    assert np.array_equal(a, b)
