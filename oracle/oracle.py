"""ctypes binding of the CPU ORACLE (oracle/sf_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs.  The product package ``smooth_feedback_b200`` never imports this module.

Restates /root/reference/include/smooth/feedback/{qp_solver.hpp:92-730, ekf.hpp:79-139}; see sf_oracle.h.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "build")


class SfoQpParams(C.Structure):
    # field-for-field mirror of sfo_qp_params (QPSolverParams, qp_solver.hpp:29-68)
    _fields_ = [
        ("alpha", C.c_float),
        ("rho", C.c_float),
        ("sigma", C.c_float),
        ("scaling", C.c_int32),
        ("eps_abs", C.c_float),
        ("eps_rel", C.c_float),
        ("eps_primal_inf", C.c_float),
        ("eps_dual_inf", C.c_float),
        ("has_max_iter", C.c_int32),
        ("max_iter", C.c_uint32),
        ("stop_check_iter", C.c_uint32),
        ("polish", C.c_int32),
        ("polish_iter", C.c_uint32),
        ("delta", C.c_float),
    ]


def build(force: bool = False) -> None:
    """Compile oracle/build/liboracle{,_fast}.so with the committed Makefile."""
    if force:
        subprocess.check_call(["make", "-C", _HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_libs: dict[str, C.CDLL] = {}


def _lib(fast: bool = False) -> C.CDLL:
    name = "liboracle_fast.so" if fast else "liboracle.so"
    if name in _libs:
        return _libs[name]
    path = os.path.join(_BUILD, name)
    if not os.path.exists(path):
        build()
    lib = C.CDLL(path)
    dp = C.POINTER(C.c_double)
    lib.sfo_qp_params_default.argtypes = [C.POINTER(SfoQpParams)]
    lib.sfo_qp_params_default.restype = None
    lib.sfo_qp_solve_dense_batch_f64.argtypes = [
        C.POINTER(SfoQpParams), C.c_int64, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, dp, dp, dp, dp,
        C.POINTER(C.c_int32), C.POINTER(C.c_uint32), C.POINTER(C.c_int8), C.c_int,
    ]
    lib.sfo_qp_solve_dense_batch_f64.restype = C.c_int
    lib.sfo_debug_dx_zero_checks.argtypes = [C.c_int]
    lib.sfo_debug_dx_zero_checks.restype = C.c_longlong
    lib.sfo_qp_scale_f64.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
    lib.sfo_qp_scale_f64.restype = C.c_int
    lib.sfo_ekf_predict_batch_f64.argtypes = [C.c_int64, C.c_int, C.c_int, dp, dp, dp, C.c_double, C.c_double, dp, C.c_int]
    lib.sfo_ekf_predict_batch_f64.restype = C.c_int
    lib.sfo_ekf_update_batch_f64.argtypes = [C.c_int64, C.c_int, C.c_int, dp, dp, dp, dp, dp, dp, C.c_int]
    lib.sfo_ekf_update_batch_f64.restype = C.c_int
    _libs[name] = lib
    return lib


def default_params(**kw) -> SfoQpParams:
    p = SfoQpParams()
    _lib().sfo_qp_params_default(C.byref(p))
    if "max_iter" in kw and kw["max_iter"] is not None:
        p.has_max_iter = 1
        p.max_iter = int(kw.pop("max_iter"))
    else:
        kw.pop("max_iter", None)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape):
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.shape == tuple(shape), (a.shape, shape)
    return a


@dataclass
class OracleQPResult:
    x: np.ndarray       # [B, n]
    y: np.ndarray       # [B, m]
    obj: np.ndarray     # [B]
    status: np.ndarray  # [B] int32, QPSolutionStatus order (qp.hpp:82-92)
    iter: np.ndarray    # [B] uint32
    active: np.ndarray  # [B, m] int8


def qp_solve_batch(P, q, A, l, u, params: SfoQpParams | None = None, warm_x=None, warm_y=None,
                   nthreads: int = 1, fast: bool = False) -> OracleQPResult:
    """P [B,n,n], A [B,m,n] given as *row-major numpy views of the math matrices* (P[b,i,j] = P_ij).

    Internally converted to the reference's column-major storage.
    """
    P = np.asarray(P, dtype=np.float64)
    A = np.asarray(A, dtype=np.float64)
    B, n, _ = P.shape
    m = A.shape[1]
    Pc = np.ascontiguousarray(np.transpose(P, (0, 2, 1)))  # column-major per instance
    Ac = np.ascontiguousarray(np.transpose(A, (0, 2, 1)))
    q = _f64(q, (B, n)); l = _f64(l, (B, m)); u = _f64(u, (B, m))
    if warm_x is not None:
        warm_x = _f64(warm_x, (B, n)); warm_y = _f64(warm_y, (B, m))
    prm = params if params is not None else default_params()
    x = np.empty((B, n)); y = np.empty((B, m)); obj = np.empty(B)
    st = np.empty(B, dtype=np.int32); it = np.empty(B, dtype=np.uint32); act = np.empty((B, m), dtype=np.int8)
    rc = _lib(fast).sfo_qp_solve_dense_batch_f64(
        C.byref(prm), B, n, m, _dp(Pc), _dp(q), _dp(Ac), _dp(l), _dp(u), _dp(warm_x), _dp(warm_y),
        _dp(x), _dp(y), _dp(obj), st.ctypes.data_as(C.POINTER(C.c_int32)),
        it.ctypes.data_as(C.POINTER(C.c_uint32)), act.ctypes.data_as(C.POINTER(C.c_int8)), int(nthreads))
    if rc != 0:
        raise ValueError(f"oracle rejected the call (rc={rc})")
    return OracleQPResult(x, y, obj, st, it, act)


def qp_scale(P, q, A):
    """Return (c, sx, sy) of QPSolver::scale for one instance (P [n,n], A [m,n] math layout)."""
    P = np.asarray(P, dtype=np.float64); A = np.asarray(A, dtype=np.float64)
    n = P.shape[0]; m = A.shape[0]
    Pc = np.ascontiguousarray(P.T); Ac = np.ascontiguousarray(A.T)
    q = _f64(q, (n,))
    c = C.c_double(); sx = np.empty(n); sy = np.empty(m)
    _lib().sfo_qp_scale_f64(n, m, _dp(Pc), _dp(q), _dp(Ac), C.byref(c), _dp(sx), _dp(sy))
    return c.value, sx, sy


def ekf_predict_batch(P, A, Q, tau: float, dt: float | None = None, stepper: str = "euler", nthreads: int = 1,
                      fast: bool = False):
    """P, A, Q: [B,d,d] math layout.  Returns propagated covariance [B,d,d]."""
    P = np.asarray(P, dtype=np.float64)
    B, d, _ = P.shape
    tc = lambda M: np.ascontiguousarray(np.transpose(np.asarray(M, dtype=np.float64), (0, 2, 1)))
    Pc, Ac, Qc = tc(P), tc(A), tc(Q)
    out = np.empty_like(Pc)
    rc = _lib(fast).sfo_ekf_predict_batch_f64(B, d, {"euler": 0, "rk4": 1}[stepper], _dp(Pc), _dp(Ac), _dp(Qc),
                                              float(tau), -1.0 if dt is None else float(dt), _dp(out), int(nthreads))
    if rc != 0:
        raise ValueError(f"oracle rejected the call (rc={rc})")
    return np.transpose(out, (0, 2, 1)).copy()


def ekf_update_batch(P, H, R, innov, nthreads: int = 1, fast: bool = False):
    """P [B,d,d], H [B,ny,d], R [B,ny,ny], innov [B,ny] -> (delta [B,d], P_new [B,d,d])."""
    P = np.asarray(P, dtype=np.float64)
    B, d, _ = P.shape
    ny = np.asarray(H).shape[1]
    tc = lambda M: np.ascontiguousarray(np.transpose(np.asarray(M, dtype=np.float64), (0, 2, 1)))
    Pc, Hc, Rc = tc(P), tc(H), tc(R)
    innov = _f64(innov, (B, ny))
    delta = np.empty((B, d)); out = np.empty_like(Pc)
    rc = _lib(fast).sfo_ekf_update_batch_f64(B, d, ny, _dp(Pc), _dp(Hc), _dp(Rc), _dp(innov), _dp(delta), _dp(out),
                                             int(nthreads))
    if rc != 0:
        raise ValueError(f"oracle rejected the call (rc={rc})")
    return delta, np.transpose(out, (0, 2, 1)).copy()


def dx_zero_checks(reset: bool = False, fast: bool = False) -> int:
    """Stop checks so far (in this process) at which the oracle saw an exactly stationary primal iterate; see sf_oracle.cpp."""
    return int(_lib(fast).sfo_debug_dx_zero_checks(int(reset)))
