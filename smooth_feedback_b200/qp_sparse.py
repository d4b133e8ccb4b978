"""Host-side mirror of the reference's sparse QP interface on top of the C ABI.

Reference (pettni/smooth_feedback @ 9a08971): ``QuadraticProgramSparse`` (qp.hpp:60-79: P column-major sparse, A ROW-major
sparse, dense q/l/u) solved by ``QPSolver<QuadraticProgramSparse<double>>`` (qp_solver.hpp:343-568, sparse branches) -- the
call ``MPC::operator()`` makes at mpc.hpp:491.  The engine solves a BATCH of such problems that share one sparsity
pattern (one OCP structure, many agents / time steps); the pattern is analysed once (``SparsePattern`` =
``SimplicialLDLT::analyzePattern``, qp_solver.hpp:424).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import Handle
from .qp import QPBatchResult, QPSolverParams, _is_torch, _ptr, default_handle


@dataclass
class QuadraticProgramSparse:
    """qp.hpp:60-79 in compressed storage: P CSC (only col >= row enters the KKT matrix), A CSR."""
    n: int
    m: int
    P_colptr: np.ndarray
    P_rowidx: np.ndarray
    P_vals: np.ndarray
    q: np.ndarray
    A_rowptr: np.ndarray
    A_colidx: np.ndarray
    A_vals: np.ndarray
    l: np.ndarray
    u: np.ndarray

    def dense(self):
        """-> (P [n,n], A [m,n]) with the stored entries (what the dense reference path would see)."""
        P = np.zeros((self.n, self.n)); A = np.zeros((self.m, self.n))
        for j in range(self.n):
            for e in range(self.P_colptr[j], self.P_colptr[j + 1]):
                P[self.P_rowidx[e], j] = self.P_vals[e]
        for i in range(self.m):
            for e in range(self.A_rowptr[i], self.A_rowptr[i + 1]):
                A[i, self.A_colidx[e]] = self.A_vals[e]
        return P, A


class SparsePattern:
    """Owns one sfb_qp_sparse_pattern_t: ordering, symbolic L D L^T factor and assembly schedules on the device."""

    def __init__(self, n: int, m: int, P_colptr, P_rowidx, A_rowptr, A_colidx, handle: Handle | None = None, a_csc: bool = False):
        """a_csc = True: A is given column-compressed like OSQP's csc_matrix (compat/osqp.hpp:36-49): A_rowptr / A_colidx then
        hold A's COLUMN pointers [n+1] / ROW indices, and solve_sparse_batch expects A_vals in that order."""
        self.n, self.m = int(n), int(m)
        self.a_csc = bool(a_csc)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.P_colptr, self.P_rowidx, self.A_rowptr, self.A_colidx = i32(P_colptr), i32(P_rowidx), i32(A_rowptr), i32(A_colidx)
        assert self.P_colptr.shape == (n + 1,) and self.A_rowptr.shape == ((n if a_csc else m) + 1,)
        self.nnzP, self.nnzA = int(self.P_colptr[-1]), int(self.A_rowptr[-1])
        self.handle = handle or default_handle(0)
        self._p = C.c_void_p()
        ip = lambda a: a.ctypes.data_as(C.c_void_p)
        analyze = _lib.lib().sfb_qp_sparse_analyze_csc if a_csc else _lib.lib().sfb_qp_sparse_analyze
        self.handle.check(analyze(self.handle.raw, self.n, self.m, ip(self.P_colptr), ip(self.P_rowidx),
                                  ip(self.A_rowptr), ip(self.A_colidx), C.byref(self._p)))
        nnzL, flops = C.c_int64(), C.c_int64()
        self.perm = np.empty(n, np.int32)
        _lib.lib().sfb_qp_sparse_pattern_info(self._p, C.byref(nnzL), C.byref(flops), ip(self.perm))
        self.nnzL, self.factor_flops = int(nnzL.value), int(flops.value)

    @property
    def raw(self):
        return self._p

    def uses_onchip(self, scalar_bytes: int = 8):
        """-> (bool, info dict): whether solves with this pattern run the on-chip kernel (one CTA per instance, working set in
        shared memory; qp_sparse_cta.cuh) and its analysis (supernodes, levels, factor slots, sweep stages, smem bytes)."""
        info = np.zeros(8, np.int64)
        r = _lib.lib().sfb_qp_sparse_uses_onchip(self.handle.raw, self._p, int(scalar_bytes), info.ctypes.data_as(C.c_void_p))
        keys = ("supernodes", "levels", "factor_slots", "nnzL", "factor_flops", "largest_supernode", "sweep_stages", "smem_bytes")
        return bool(r), dict(zip(keys, info.tolist()))

    def bytes_per_iteration(self, scalar: int = 8) -> int:
        """Algorithmic bytes one ADMM iteration streams per instance: Abar twice, the L D L^T factor twice, and the
        vector traffic of the iterate updates."""
        return scalar * (2 * self.nnzA + 2 * self.nnzL + self.n + 5 * self.n + 10 * self.m)

    def bytes_compulsory(self, scalar: int = 8) -> int:
        return scalar * (self.nnzP + self.nnzA + self.n + 2 * self.m) + scalar * (self.n + self.m + 1) + 8

    def __del__(self):
        try:
            if self._p:
                _lib.lib().sfb_qp_sparse_pattern_destroy(self._p)
                self._p = C.c_void_p()
        except Exception:
            pass


def sparse_onchip_selfcheck(n: int, m: int, P_colptr, P_rowidx, A_rowptr, A_colidx, ordering: int = -1):
    """Host-only check of the on-chip kernel's schedules (sfb_qp_sparse_cta_selfcheck): analyses the pattern (ordering -1: cost
    model, 0: minimum degree, 1: nested dissection), runs assembly / factorisation / inversion / staged sweeps on the host from the
    very tables the kernel uses and compares with a dense solve.  -> (info dict, max relative error over both layouts)."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    pc, pr, ar, ac = i32(P_colptr), i32(P_rowidx), i32(A_rowptr), i32(A_colidx)
    ip = lambda a: a.ctypes.data_as(C.c_void_p)
    info = np.zeros(8, np.int64); err = C.c_double()
    rc = _lib.lib().sfb_qp_sparse_cta_selfcheck(int(n), int(m), ip(pc), ip(pr), ip(ar), ip(ac), int(ordering), ip(info), C.byref(err))
    if rc != 0:
        raise _lib.SfbError(rc, _lib.lib().sfb_last_error_message(None).decode())
    keys = ("supernodes", "levels", "factor_slots", "nnzL", "factor_flops", "largest_supernode", "sweep_stages", "ordering")
    return dict(zip(keys, info.tolist())), float(err.value)


def sparse_symbolic(n: int, m: int, P_colptr, P_rowidx, A_rowptr, A_colidx):
    """Host-only symbolic analysis (sfb_qp_sparse_symbolic): -> dict(nnzL, factor_flops, perm [n], L_colptr [n+1])."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    pc, pr, ar, ac = i32(P_colptr), i32(P_rowidx), i32(A_rowptr), i32(A_colidx)
    ip = lambda a: a.ctypes.data_as(C.c_void_p)
    nnzL, flops = C.c_int64(), C.c_int64()
    perm = np.empty(n, np.int32); lcp = np.empty(n + 1, np.int32)
    rc = _lib.lib().sfb_qp_sparse_symbolic(int(n), int(m), ip(pc), ip(pr), ip(ar), ip(ac), C.byref(nnzL), C.byref(flops), ip(perm), ip(lcp))
    if rc != 0:
        raise _lib.SfbError(rc, _lib.lib().sfb_last_error_message(None).decode())
    return dict(nnzL=int(nnzL.value), factor_flops=int(flops.value), perm=perm, L_colptr=lcp)


def solve_sparse_batch(pattern: SparsePattern, P_vals, q, A_vals, l, u, prm: QPSolverParams | None = None, warm_x=None,
                       warm_y=None, out: QPBatchResult | None = None) -> QPBatchResult:
    """Solve B sparse QPs sharing ``pattern`` (sfb_qp_solve_sparse_batch_f64/_f32).

    P_vals [B, nnzP], A_vals [B, nnzA] are the value arrays in pattern order; q [B, n]; l, u [B, m].  All contiguous numpy
    arrays (host path) or all contiguous torch CUDA tensors (device path, asynchronous on the current stream).
    """
    torch_mode = _is_torch(q)
    B, n = q.shape
    m = pattern.m
    assert n == pattern.n and tuple(P_vals.shape) == (B, pattern.nnzP) and tuple(A_vals.shape) == (B, pattern.nnzA)
    assert tuple(l.shape) == (B, m) and tuple(u.shape) == (B, m)
    prm = prm or QPSolverParams()
    cprm = prm.to_c()
    h = pattern.handle
    if torch_mode:
        import torch

        f64 = q.dtype == torch.float64
        assert q.dtype in (torch.float64, torch.float32)
        ins = [P_vals, q, A_vals, l, u] + ([warm_x, warm_y] if warm_x is not None else [])
        for t in ins:
            assert t.is_cuda and t.is_contiguous() and t.dtype == q.dtype, "need contiguous CUDA tensors of one dtype"
        dev = q.device
        assert (dev.index or 0) == h.device
        h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
        if out is None:
            out = QPBatchResult(
                x=torch.empty((B, n), dtype=q.dtype, device=dev), y=torch.empty((B, m), dtype=q.dtype, device=dev),
                obj=torch.empty((B,), dtype=q.dtype, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
                iter=torch.empty((B,), dtype=torch.int32, device=dev), active=torch.empty((B, m), dtype=torch.int8, device=dev),
                flags=torch.empty((B,), dtype=torch.int32, device=dev))
    else:
        q = np.ascontiguousarray(q)
        f64 = q.dtype == np.float64
        assert q.dtype in (np.float64, np.float32)
        dt = q.dtype
        P_vals, A_vals, l, u = (np.ascontiguousarray(t, dtype=dt) for t in (P_vals, A_vals, l, u))
        if warm_x is not None:
            warm_x = np.ascontiguousarray(warm_x, dtype=dt)
            warm_y = np.ascontiguousarray(warm_y, dtype=dt)
        if out is None:
            out = QPBatchResult(x=np.empty((B, n), dt), y=np.empty((B, m), dt), obj=np.empty((B,), dt),
                                status=np.empty((B,), np.int32), iter=np.empty((B,), np.uint32),
                                active=np.empty((B, m), np.int8), flags=np.empty((B,), np.uint32))
    fn = _lib.lib().sfb_qp_solve_sparse_batch_f64 if f64 else _lib.lib().sfb_qp_solve_sparse_batch_f32
    if pattern.a_csc:
        assert f64, "the CSC ingestion entry point is fp64 (OSQP's c_float)"
        fn = _lib.lib().sfb_qp_solve_sparse_batch_csc_f64
    rc = fn(h.raw, pattern.raw, C.byref(cprm), B, _ptr(P_vals), _ptr(q), _ptr(A_vals), _ptr(l), _ptr(u), _ptr(warm_x),
            _ptr(warm_y), _ptr(out.x), _ptr(out.y), _ptr(out.obj), _ptr(out.status), _ptr(out.iter), _ptr(out.active),
            _ptr(out.flags))
    h.check(rc)
    return out
