"""Pins the CPU oracle with the reference's own known-answer tests (tests/test_qp.cpp:54-336).

CPU only.  Also records the iteration counts the oracle takes so that drift is visible.
"""
import numpy as np
import pytest

from qp_cases import CASES, OPTIMAL, as_batch, is_approx


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_known_answer(oracle, case):
    P, q, A, l, u = as_batch(case)
    r = oracle.qp_solve_batch(P, q, A, l, u)  # test_prm == library defaults with polish=true (test_qp.cpp:32-35)
    assert r.status[0] == case["status"]
    if case["x"] is not None:
        assert is_approx(r.x[0], case["x"], case["x_rtol"])
    if case["obj"] is not None:
        assert abs(r.obj[0] - case["obj"]) <= case["obj_atol"]
    if case["status"] == OPTIMAL:
        # every numeric reference test re-solves with its own solution as warm start
        r2 = oracle.qp_solve_batch(P, q, A, l, u, warm_x=r.x, warm_y=r.y)
        assert r2.status[0] == OPTIMAL
        assert is_approx(r2.x[0], case["x"], case["x_rtol"])
        if case["obj"] is not None:
            assert abs(r2.obj[0] - case["obj"]) <= case["obj_atol"]
        assert r2.iter[0] == 2  # exits at the first stop check (iter % 25 == 1)


def test_iteration_cadence(oracle):
    # sol.iter is the loop counter after increment: 0 (trivial infeasibility) or 25k+2  (qp_solver.hpp:449,465,548)
    for case in CASES:
        P, q, A, l, u = as_batch(case)
        r = oracle.qp_solve_batch(P, q, A, l, u)
        it = int(r.iter[0])
        assert it == 0 or it % 25 == 2, (case["name"], it)
    P, q, A, l, u = as_batch(CASES[3])  # PrimalInfeasibleEasy -> trivial check, no iterations
    assert oracle.qp_solve_batch(P, q, A, l, u).iter[0] == 0


def test_max_iter(oracle):
    P, q, A, l, u = as_batch(CASES[7])  # Portfolio needs > 100 iterations
    prm = oracle.default_params(max_iter=10)
    r = oracle.qp_solve_batch(P, q, A, l, u, params=prm)
    assert r.status[0] == 4 and r.iter[0] == 10  # MaxIterations


def test_float_params_are_floats(oracle):
    # rho = (double)0.1f, not 0.1 (qp_solver.hpp:37,353)
    prm = oracle.default_params()
    assert float(prm.rho) == float(np.float32(0.1)) != 0.1


def test_scipy_crosscheck_random(oracle):
    """Independent check of the restated algorithm: polished solutions satisfy the KKT conditions."""
    rng = np.random.default_rng(0)
    B, n, m = 16, 10, 20
    A = rng.uniform(-1, 1, (B, m, n))
    L = np.tril(rng.uniform(-1, 1, (B, n, n)))
    idx = np.arange(n)
    L[:, idx, idx] = np.maximum(np.abs(L[:, idx, idx]), 0.05)
    P = L @ np.transpose(L, (0, 2, 1))
    q = rng.uniform(-1, 1, (B, n)); v = rng.uniform(-1, 1, (B, n))
    l = np.full((B, m), -np.inf)
    u = np.einsum("bij,bj->bi", A, v) + rng.uniform(0, 1, (B, m))
    prm = oracle.default_params(max_iter=4000)
    r = oracle.qp_solve_batch(P, q, A, l, u, params=prm)
    assert (r.status == 0).all()
    Ax = np.einsum("bij,bj->bi", A, r.x)
    assert (Ax <= u + 1e-7).all()
    stat = np.einsum("bij,bj->bi", P, r.x) + q + np.einsum("bij,bi->bj", A, r.y)
    assert np.abs(stat).max() < 1e-6
    assert (r.y >= -1e-9).all()                       # l = -inf -> multipliers of upper bounds are >= 0
    assert np.abs(r.y * (Ax - u)).max() < 1e-6        # complementarity
