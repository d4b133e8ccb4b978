"""Host-side mirror of smooth_feedback's QP interface on top of the C ABI (include/sfb.h).

Mirrors, name for name (reference paths relative to pettni/smooth_feedback @ 9a08971):
  QPSolverParams      include/smooth/feedback/qp_solver.hpp:29-68
  QuadraticProgram    include/smooth/feedback/qp.hpp:31-45
  QPSolutionStatus    include/smooth/feedback/qp.hpp:82-92
  QPSolution          include/smooth/feedback/qp.hpp:95-108
  QPSolver            include/smooth/feedback/qp_solver.hpp:242-757   (analyze / solve / sol)
  solve_qp            include/smooth/feedback/qp_solver.hpp:779-787
plus the one thing the reference does not have: ``solve_batch`` / ``solve_dense_batch`` over many independent
problems of one shape, which is what the GPU engine is for.

All numerics run in libsfb.so on the GPU.  There is no CPU path: without the library or a device these
functions raise ``SfbError``.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import Handle, SfbQpParams


class QPSolutionStatus(enum.IntEnum):
    """qp.hpp:82-92 -- values and order are part of the contract."""

    Optimal = 0
    PolishFailed = 1
    PrimalInfeasible = 2
    DualInfeasible = 3
    MaxIterations = 4
    MaxTime = 5
    Unknown = 6


FLAG_POLISHED = 1
FLAG_POLISH_SKIPPED = 2
FLAG_POLISH_FAILED = 4
FLAG_POLISH_SCRATCH = 8
FLAG_POLISH_REDUCED = 16


@dataclass
class QPSolverParams:
    """qp_solver.hpp:29-68.  The float-typed members are rounded to float32 exactly as the reference stores them."""

    verbose: bool = False
    alpha: float = 1.6
    rho: float = 0.1
    sigma: float = 1e-6
    scaling: bool = True
    eps_abs: float = 1e-3
    eps_rel: float = 1e-3
    eps_primal_inf: float = 1e-4
    eps_dual_inf: float = 1e-4
    max_iter: Optional[int] = None
    max_time: Optional[float] = None  # seconds (std::chrono::nanoseconds in the reference)
    stop_check_iter: int = 25
    polish: bool = True
    polish_iter: int = 5
    delta: float = 1e-6

    def to_c(self) -> SfbQpParams:
        p = SfbQpParams()
        p.verbose = int(self.verbose)
        p.alpha, p.rho, p.sigma = self.alpha, self.rho, self.sigma
        p.scaling = int(self.scaling)
        p.eps_abs, p.eps_rel = self.eps_abs, self.eps_rel
        p.eps_primal_inf, p.eps_dual_inf = self.eps_primal_inf, self.eps_dual_inf
        p.has_max_iter = int(self.max_iter is not None)
        p.max_iter = int(self.max_iter or 0)
        p.has_max_time = int(self.max_time is not None)
        p.max_time_ns = int((self.max_time or 0.0) * 1e9)
        p.stop_check_iter = int(self.stop_check_iter)
        p.polish = int(self.polish)
        p.polish_iter = int(self.polish_iter)
        p.delta = self.delta
        return p


@dataclass
class QuadraticProgram:
    """min 1/2 x'Px + q'x  s.t.  l <= Ax <= u   (qp.hpp:31-45).  Arrays in math layout: P[i, j] = P_ij."""

    P: np.ndarray
    q: np.ndarray
    A: np.ndarray
    l: np.ndarray
    u: np.ndarray


@dataclass
class QPSolution:
    """qp.hpp:95-108"""

    code: QPSolutionStatus = QPSolutionStatus.Unknown
    iter: int = 0
    primal: np.ndarray = field(default_factory=lambda: np.zeros(0))
    dual: np.ndarray = field(default_factory=lambda: np.zeros(0))
    objective: float = 0.0


@dataclass
class QPBatchResult:
    """Outputs of one batched call; numpy arrays for host inputs, torch CUDA tensors for device inputs."""

    x: object       # [B, n]
    y: object       # [B, m]
    obj: object     # [B]
    status: object  # [B] int32 (QPSolutionStatus values)
    iter: object    # [B] uint32 (int32 view for torch)
    active: object  # [B, m] int8: -1 lower-active, +1 upper-active (polish_qp's sets, qp_solver.hpp:113-123)
    flags: object   # [B] FLAG_* diagnostics


_default_handles: dict[int, Handle] = {}


def default_handle(device: int = 0) -> Handle:
    h = _default_handles.get(device)
    if h is None:
        h = _default_handles[device] = Handle(device)
    return h


def _is_torch(t) -> bool:
    return type(t).__module__.startswith("torch")


def _ptr(t):
    if t is None:
        return None
    if _is_torch(t):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


def solve_dense_batch(P_cm, q, A_cm, l, u, prm: QPSolverParams | None = None, warm_x=None, warm_y=None,
                      handle: Handle | None = None, out: QPBatchResult | None = None) -> QPBatchResult:
    """Solve B independent dense QPs of one shape on the GPU (sfb_qp_solve_dense_batch_f64/_f32).

    Storage layout is the reference's: per instance column-major, i.e.
        P_cm [B, n, n] with P_cm[b, j, i] = P_ij        q [B, n]
        A_cm [B, n, m] with A_cm[b, j, i] = A_ij        l, u [B, m]
    (``to_colmajor`` converts from math layout).  Inputs are either all contiguous numpy arrays (host path:
    staged through the engine in pipelined chunks) or all contiguous torch CUDA tensors (device path:
    asynchronous on the handle's stream).  dtype float64 or float32.
    """
    torch_mode = _is_torch(P_cm)
    B, n, n2 = P_cm.shape
    m = A_cm.shape[2] if A_cm is not None else 0
    assert n == n2 and tuple(q.shape) == (B, n)
    if m > 0:
        assert tuple(A_cm.shape) == (B, n, m) and tuple(l.shape) == (B, m) and tuple(u.shape) == (B, m)
    prm = prm or QPSolverParams()
    cprm = prm.to_c()
    if torch_mode:
        import torch

        f64 = P_cm.dtype == torch.float64
        assert P_cm.dtype in (torch.float64, torch.float32)
        ins = [P_cm, q, A_cm, l, u] + ([warm_x, warm_y] if warm_x is not None else [])
        for t in ins:
            assert t.is_cuda and t.is_contiguous() and t.dtype == P_cm.dtype, "need contiguous CUDA tensors of one dtype"
        dev = P_cm.device
        h = handle or default_handle(dev.index or 0)
        h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        if out is None:
            out = QPBatchResult(
                x=torch.empty((B, n), dtype=P_cm.dtype, device=dev), y=torch.empty((B, m), dtype=P_cm.dtype, device=dev),
                obj=torch.empty((B,), dtype=P_cm.dtype, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
                iter=torch.empty((B,), dtype=torch.int32, device=dev), active=torch.empty((B, m), dtype=torch.int8, device=dev),
                flags=torch.empty((B,), dtype=torch.int32, device=dev))
    else:
        P_cm = np.ascontiguousarray(P_cm)
        f64 = P_cm.dtype == np.float64
        assert P_cm.dtype in (np.float64, np.float32)
        dt = P_cm.dtype
        q, A_cm, l, u = (np.ascontiguousarray(t, dtype=dt) for t in (q, A_cm, l, u))
        if warm_x is not None:
            warm_x = np.ascontiguousarray(warm_x, dtype=dt)
            warm_y = np.ascontiguousarray(warm_y, dtype=dt)
        h = handle or default_handle(0)
        if out is None:
            out = QPBatchResult(x=np.empty((B, n), dt), y=np.empty((B, m), dt), obj=np.empty((B,), dt),
                                status=np.empty((B,), np.int32), iter=np.empty((B,), np.uint32),
                                active=np.empty((B, m), np.int8), flags=np.empty((B,), np.uint32))
    fn = _lib.lib().sfb_qp_solve_dense_batch_f64 if f64 else _lib.lib().sfb_qp_solve_dense_batch_f32
    rc = fn(h.raw, C.byref(cprm), B, n, m, _ptr(P_cm), _ptr(q), _ptr(A_cm), _ptr(l), _ptr(u), _ptr(warm_x), _ptr(warm_y),
            _ptr(out.x), _ptr(out.y), _ptr(out.obj), _ptr(out.status), _ptr(out.iter), _ptr(out.active), _ptr(out.flags))
    h.check(rc)
    return out


def to_colmajor(M):
    """math layout [..., r, c] -> the reference's column-major storage [..., c, r], contiguous."""
    if _is_torch(M):
        return M.transpose(-1, -2).contiguous()
    return np.ascontiguousarray(np.swapaxes(np.asarray(M), -1, -2))


def qp_scale_batch(P_cm, q, A_cm, handle: Handle | None = None):
    """QPSolver::scale alone (qp_solver.hpp:673-730) on CUDA tensors -> (c [B], sx [B,n], sy [B,m])."""
    import torch

    B, n, _ = P_cm.shape
    m = A_cm.shape[2]
    dev = P_cm.device
    h = handle or default_handle(dev.index or 0)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    c = torch.empty((B,), dtype=torch.float64, device=dev)
    sx = torch.empty((B, n), dtype=torch.float64, device=dev)
    sy = torch.empty((B, m), dtype=torch.float64, device=dev)
    h.check(_lib.lib().sfb_qp_scale_dense_batch_f64(h.raw, B, n, m, _ptr(P_cm), _ptr(q), _ptr(A_cm), _ptr(c), _ptr(sx), _ptr(sy)))
    return c, sx, sy


class QPSolver:
    """QPSolver<QuadraticProgram<-1,-1,double>> (qp_solver.hpp:242-757) with the same call surface.

    ``solve`` marshals a batch of one through the C ABI; ``solve_batch`` is the extension the engine exists for.
    Copies of a solver are independent objects (the reference's LDLTWrapper drops the factor on copy,
    qp_solver.hpp:209-231; the engine refactorises on every solve, so copies behave identically).
    """

    def __init__(self, pbm: QuadraticProgram | None = None, prm: QPSolverParams | None = None, *,
                 handle: Handle | None = None):
        if isinstance(pbm, QPSolverParams) and prm is None:  # QPSolver(prm) overload, qp_solver.hpp:267
            pbm, prm = None, pbm
        self.prm_ = prm or QPSolverParams()
        self._handle = handle
        self.sol_ = QPSolution()
        if pbm is not None:
            self.analyze(pbm)

    def analyze(self, pbm: QuadraticProgram) -> None:
        """qp_solver.hpp:297-338: size the working memory, zero the solution."""
        n = np.asarray(pbm.A).shape[1]
        m = np.asarray(pbm.A).shape[0]
        self.sol_ = QPSolution(primal=np.zeros(n), dual=np.zeros(m))

    def sol(self) -> QPSolution:
        return self.sol_

    def solve(self, pbm: QuadraticProgram, warmstart: QPSolution | None = None) -> QPSolution:
        """qp_solver.hpp:343-568"""
        sols = self.solve_batch([pbm], None if warmstart is None else [warmstart])
        self.sol_ = sols[0]
        return self.sol_

    def solve_batch(self, pbms: Sequence[QuadraticProgram], warmstarts: Sequence[QPSolution] | None = None) -> list[QPSolution]:
        f = lambda name: np.stack([np.asarray(getattr(p, name), dtype=np.float64) for p in pbms])
        P, q, A, l, u = f("P"), f("q"), f("A"), f("l"), f("u")
        if A.ndim == 2:  # M == 1 row vectors
            A = A[:, None, :]
        wx = wy = None
        if warmstarts is not None:
            wx = np.stack([np.asarray(w.primal, dtype=np.float64) for w in warmstarts])
            wy = np.stack([np.asarray(w.dual, dtype=np.float64) for w in warmstarts])
        r = solve_dense_batch(to_colmajor(P), q, to_colmajor(A), l, u, self.prm_, wx, wy, handle=self._handle)
        return [QPSolution(QPSolutionStatus(int(r.status[b])), int(r.iter[b]), r.x[b].copy(), r.y[b].copy(), float(r.obj[b]))
                for b in range(len(pbms))]


def solve_qp(pbm: QuadraticProgram, prm: QPSolverParams | None = None, warmstart: QPSolution | None = None) -> QPSolution:
    """solve_qp, qp_solver.hpp:779-787: fresh solver per call."""
    return QPSolver(pbm, prm or QPSolverParams()).solve(pbm, warmstart)
