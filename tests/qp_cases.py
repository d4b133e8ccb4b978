"""Known-answer QP cases transliterated from the reference's own tests.

Source: /root/reference/tests/test_qp.cpp (line ranges per case).  Same numbers, same tolerances.
Each case: dict(name, P, q, A, l, u, status, x (or None), x_rtol, obj (or None), obj_atol).
Matrices are in math layout (row-major numpy).
"""
import numpy as np

inf = np.inf

OPTIMAL, POLISH_FAILED, PRIMAL_INFEASIBLE, DUAL_INFEASIBLE, MAX_ITERATIONS, MAX_TIME, UNKNOWN = range(7)

_P3 = np.array([[4.0, 2, 2], [2, 4, 2], [2, 2, 4]])

CASES = [
    dict(  # test_qp.cpp:54-73 BasicStatic (== BasicDynamic :75-98, BasicSparse :100-122, BasicPartialDynamic :124-147)
        name="Basic",
        P=np.eye(2), q=[-4, 0.25], A=np.eye(2), l=[-1, -1], u=[1, 1],
        status=OPTIMAL, x=[1, -0.25], x_rtol=1e-4, obj=0.5 - 4 - 1.0 / 32, obj_atol=1e-4,
    ),
    dict(  # :149-166
        name="Unconstrained",
        P=_P3, q=[-8, -6, -10], A=np.zeros((1, 3)), l=[-inf], u=[inf],
        status=OPTIMAL, x=[1, 0, 2], x_rtol=1e-4, obj=None, obj_atol=None,
    ),
    dict(  # :168-185
        name="HalfConstrained",
        P=_P3, q=[-8, -6, -10], A=np.eye(3), l=[-inf, -inf, -10], u=[inf, 10, inf],
        status=OPTIMAL, x=[1, 0, 2], x_rtol=1e-4, obj=None, obj_atol=None,
    ),
    dict(  # :187-199
        name="PrimalInfeasibleEasy",
        P=np.eye(2), q=[0.1, 0.1], A=np.eye(2), l=[-1, 1], u=[1, -1],
        status=PRIMAL_INFEASIBLE, x=None, x_rtol=None, obj=None, obj_atol=None,
    ),
    dict(  # :201-213
        name="PrimalInfeasibleHard",
        P=np.eye(2), q=[0.1, 0.1], A=[[1, 1], [-1, -1]], l=[0.5, 0.5], u=[1, 1],
        status=PRIMAL_INFEASIBLE, x=None, x_rtol=None, obj=None, obj_atol=None,
    ),
    dict(  # :215-227
        name="PrimalInfeasibleInfinity",
        P=np.eye(2), q=[0.1, 0.1], A=[[1, 1], [-1, -1], [1, 0], [0, 1]], l=[0.5, 0.5, -inf, -inf],
        u=[1, 1, inf, inf],
        status=PRIMAL_INFEASIBLE, x=None, x_rtol=None, obj=None, obj_atol=None,
    ),
    dict(  # :229-242
        name="DualInfeasible",
        P=np.diag([1.0, 0.0]), q=[1, -1], A=np.eye(2), l=[-1, -inf], u=[1, inf],
        status=DUAL_INFEASIBLE, x=None, x_rtol=None, obj=None, obj_atol=None,
    ),
    dict(  # :244-272 PortfolioOptimization (== Sparse variant :274-312)
        name="Portfolio",
        P=[[0.018641, 0.00359853, 0.00130976], [0.00359853, 0.00643694, 0.00488727],
           [0.00130976, 0.00488727, 0.0686828]],
        q=[0, 0, 0],
        A=[[1, 1, 1], [0.0260022, 0.00810132, 0.0737159], [1, 0, 0], [0, 1, 0], [0, 0, 1]],
        l=[-inf, 50, 0, 0, 0], u=[1000, inf, inf, inf, inf],
        status=OPTIMAL, x=[497.04552984986384, 0.0, 502.9544801594811], x_rtol=1e-4,
        obj=22634.417849884154 / 2, obj_atol=5e-2,
    ),
    dict(  # :314-336
        name="TwoDimensional",
        P=[[0.0100131, 0], [0, 0.01]], q=[-0.329554, 0.536459], A=[[-0.0639209, -0.168], [-0.467, 0]],
        l=[-inf, -inf], u=[-0.034974, 0.46571],
        status=OPTIMAL, x=[46.6338, -17.5351], x_rtol=1e-4, obj=None, obj_atol=None,
    ),
]


def as_batch(case):
    """-> P[1,n,n], q[1,n], A[1,m,n], l[1,m], u[1,m] float64"""
    f = lambda a: np.asarray(a, dtype=np.float64)[None]
    return f(case["P"]), f(case["q"]), f(case["A"]), f(case["l"]), f(case["u"])


def is_approx(a, b, prec):
    """Eigen's isApprox: ||a-b|| <= prec * min(||a||, ||b||)   (2-norms)."""
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= prec * min(np.linalg.norm(a), np.linalg.norm(b))
