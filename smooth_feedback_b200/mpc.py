"""Host-side mirror of smooth_feedback's MPC for a FLEET of agents on top of the C ABI (include/sfb.h).

Mirrors (reference paths relative to pettni/smooth_feedback @ 9a08971):
  MPCParams / MPCWeights     include/smooth/feedback/mpc.hpp:309-356
  MPC::MPC                   include/smooth/feedback/mpc.hpp:405-425   -> MPCVehicleFleet.__init__ (mesh, cost, symbolic analysis)
  MPC::operator()(t, x)      include/smooth/feedback/mpc.hpp:458-519   -> MPCVehicleFleet.__call__
  MPC::reset_warmstart       include/smooth/feedback/mpc.hpp:611       -> MPCVehicleFleet.reset_warmstart
for the built-in model family of examples/mpc_asif_vehicle.cpp:42-89 (user lambdas + autodiff cannot cross the C ABI).
Transcription (ocp_to_qp_update_dyn / _ce), the sparse QP solve, the warm-start retention rule (mpc.hpp:510-516) and the
extraction of u (mpc.hpp:518) run on the device; per step only (t, x) of every agent go in and (u, code, iter) come out.
No CPU path: without libsfb.so / a B200 every call raises SfbError.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Handle, SfbQpParams
from .qp import QPSolverParams, _is_torch, _ptr, default_handle


class SfbMpcVehicleParams(C.Structure):
    """sfb_mpc_vehicle_params, field for field."""

    _fields_ = [
        ("K", C.c_int32), ("tf", C.c_double), ("Kmesh", C.c_int32), ("warmstart", C.c_int32),
        ("Q", C.c_double * 6), ("R", C.c_double * 2), ("Qtf", C.c_double * 6), ("crl", C.c_double * 2), ("cru", C.c_double * 2),
        ("drag1", C.c_double), ("drag3", C.c_double), ("g0", C.c_double * 3), ("vdes", C.c_double * 3), ("udes", C.c_double * 2),
        ("qp", SfbQpParams),
    ]


@dataclass
class MPCVehicleParams:
    """MPCParams (mpc.hpp:309-333), MPCWeights (diagonal, in effect at construction -- the reference transcribes the cost in
    its constructor only, mpc.hpp:423) and the model constants of examples/mpc_asif_vehicle.cpp:42-89; K = 50 is BASELINE
    configs[2] (the shipped example uses K = 30)."""

    K: int = 50
    tf: float = 5.0
    Kmesh: int = 4
    warmstart: bool = True
    Q: tuple = (1.0,) * 6
    R: tuple = (1.0, 1.0)
    Qtf: tuple = (1.0,) * 6
    crl: tuple = (-0.5, -0.5)
    cru: tuple = (0.5, 0.5)
    drag1: float = 0.2
    drag3: float = 0.4
    g0: tuple = (2.5, 0.0, math.pi / 2)
    vdes: tuple = (1.0, 0.0, 0.4)
    udes: tuple = (0.0, 0.0)
    qp: QPSolverParams = field(default_factory=QPSolverParams)

    def to_c(self) -> SfbMpcVehicleParams:
        p = SfbMpcVehicleParams()
        p.K, p.tf, p.Kmesh, p.warmstart = int(self.K), float(self.tf), int(self.Kmesh), int(self.warmstart)
        for name in ("Q", "R", "Qtf", "crl", "cru", "g0", "vdes", "udes"):
            arr = getattr(p, name)
            for i, v in enumerate(getattr(self, name)):
                arr[i] = float(v)
        p.drag1, p.drag3 = self.drag1, self.drag3
        p.qp = self.qp.to_c()
        return p


class MPCVehicleFleet:
    """``batch`` MPC<Time, Bundle<SE2, R^3>, R^2, F, CR> objects sharing one parameter set (mpc.hpp:372-638).

    ``fleet(t, x) -> (u, code, iter)`` is MPC::operator()(t, x) for every agent: t [B] absolute time, x [B, 7] in smooth's
    coefficient order (x, y, sin, cos, v1, v2, v3).  numpy arrays (host path) or torch CUDA tensors (device path).
    """

    def __init__(self, batch: int, prm: MPCVehicleParams | None = None, dtype=np.float64, handle: Handle | None = None,
                 device: int = 0):
        self.prm = prm or MPCVehicleParams()
        self.batch = int(batch)
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float64), np.dtype(np.float32))
        self._h = handle or default_handle(device)
        self._f = C.c_void_p()
        cp = self.prm.to_c()
        L = _lib.lib()
        self._h.check(L.sfb_mpc_fleet_create(self._h.raw, C.byref(cp), self.batch, self.dtype.itemsize, C.byref(self._f)))
        n, m, nP, nA, nL = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int64()
        self._h.check(L.sfb_mpc_fleet_dims(self._f, C.byref(n), C.byref(m), C.byref(nP), C.byref(nA), C.byref(nL)))
        self.n, self.m, self.nnzP, self.nnzA, self.nnzL = n.value, m.value, nP.value, nA.value, nL.value
        N = ((self.prm.K + self.prm.Kmesh - 1) // self.prm.Kmesh) * self.prm.Kmesh
        self.uvar_B = 6 * (N + 1)

    def pattern(self) -> dict:
        """The shared sparsity pattern in the reference's storage (CSC P, CSR A), as generators / tests consume it."""
        pc = np.empty(self.n + 1, np.int32); pr = np.empty(self.nnzP, np.int32)
        ar = np.empty(self.m + 1, np.int32); ac = np.empty(self.nnzA, np.int32)
        self._h.check(_lib.lib().sfb_mpc_fleet_pattern(self._f, _ptr(pc), _ptr(pr), _ptr(ar), _ptr(ac)))
        return dict(n=self.n, m=self.m, P_colptr=pc, P_rowidx=pr, A_rowptr=ar, A_colidx=ac)

    def reset_warmstart(self) -> None:
        self._h.check(_lib.lib().sfb_mpc_fleet_reset_warmstart(self._f))

    def _prep(self, t, x):
        if _is_torch(x):
            import torch

            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            for v in (t, x):
                assert v.is_cuda and v.is_contiguous() and v.dtype == tdt
            self._h.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
            return t, x, True
        return np.ascontiguousarray(t, dtype=self.dtype), np.ascontiguousarray(x, dtype=self.dtype), False

    def __call__(self, t, x, return_solution: bool = False):
        t, x, tm = self._prep(t, x)
        B = self.batch
        assert tuple(t.shape) == (B,) and tuple(x.shape) == (B, 7)
        if tm:
            import torch

            mk = lambda shape, dt=x.dtype: torch.empty(shape, dtype=dt, device=x.device)
            u, st, it = mk((B, 2)), mk((B,), torch.int32), mk((B,), torch.int32)
        else:
            mk = lambda shape, dt=self.dtype: np.empty(shape, dt)
            u, st, it = mk((B, 2)), mk((B,), np.int32), mk((B,), np.uint32)
        px = mk((B, self.n)) if return_solution else None
        py = mk((B, self.m)) if return_solution else None
        L = _lib.lib()
        fn = L.sfb_mpc_fleet_step_f64 if self.dtype == np.float64 else L.sfb_mpc_fleet_step_f32
        self._h.check(fn(self._f, _ptr(t), _ptr(x), _ptr(u), _ptr(st), _ptr(it), _ptr(px), _ptr(py)))
        return (u, st, it, px, py) if return_solution else (u, st, it)

    def nodes(self):
        """(N, tau [N + 1]): Mesh::N_colloc() and Mesh::all_nodes(); x_i lives at t + tf tau[i], u_i at t + tf tau[i], i < N."""
        N = C.c_int(0)
        self._h.check(_lib.lib().sfb_mpc_fleet_nodes(self._f, C.byref(N), None))
        tau = np.empty(N.value + 1, np.float64)
        self._h.check(_lib.lib().sfb_mpc_fleet_nodes(self._f, None, tau.ctypes.data_as(C.c_void_p)))
        return N.value, tau

    def trajectories(self, t):
        """The optional outputs of MPC::operator() (mpc.hpp:493-507) for the solution of the LAST step (t = the times given
        to it): u_traj [B, N, 2] and x_traj [B, N + 1, 7]."""
        N, _ = self.nodes()
        B = self.batch
        if _is_torch(t):
            import torch

            assert t.is_cuda and t.is_contiguous() and t.dtype == (torch.float64 if self.dtype == np.float64 else torch.float32)
            self._h.set_stream(torch.cuda.current_stream(t.device).cuda_stream)
            ut = torch.empty((B, N, 2), dtype=t.dtype, device=t.device)
            xt = torch.empty((B, N + 1, 7), dtype=t.dtype, device=t.device)
        else:
            t = np.ascontiguousarray(t, dtype=self.dtype)
            ut, xt = np.empty((B, N, 2), self.dtype), np.empty((B, N + 1, 7), self.dtype)
        assert tuple(t.shape) == (B,)
        L = _lib.lib()
        fn = L.sfb_mpc_fleet_trajectories_f64 if self.dtype == np.float64 else L.sfb_mpc_fleet_trajectories_f32
        self._h.check(fn(self._f, _ptr(t), _ptr(ut), _ptr(xt)))
        return ut, xt

    def to_qp(self, t, x):
        """The transcription alone -> P_vals [B,nnzP], q [B,n], A_vals [B,nnzA], l, u [B,m] (pattern(): shared pattern)."""
        assert self.dtype == np.float64
        t, x, tm = self._prep(t, x)
        B = self.batch
        if tm:
            import torch

            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=x.device)
        else:
            mk = lambda *s: np.empty(s, np.float64)
        P, q, A, l, u = mk(B, self.nnzP), mk(B, self.n), mk(B, self.nnzA), mk(B, self.m), mk(B, self.m)
        self._h.check(_lib.lib().sfb_mpc_fleet_to_qp_f64(self._f, _ptr(t), _ptr(x), _ptr(P), _ptr(q), _ptr(A), _ptr(l), _ptr(u)))
        return P, q, A, l, u

    def close(self) -> None:
        if self._f:
            _lib.lib().sfb_mpc_fleet_destroy(self._f)
            self._f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
