"""Host-side mirror of smooth_feedback's ASIFilter for a FLEET of agents on top of the C ABI (include/sfb.h).

Mirrors (reference paths relative to pettni/smooth_feedback @ 9a08971):
  ASIFilterParams / ASIFtoQPParams   include/smooth/feedback/asif.hpp:17-32, asif_func.hpp:58-68
  ASIFilter::operator()              include/smooth/feedback/asif.hpp:82-102   -> ASIFVehicleFleet.__call__
  asif_to_qp                         include/smooth/feedback/asif_func.hpp:246-261 -> ASIFVehicleFleet.to_qp
for the built-in model family of examples/mpc_asif_vehicle.cpp (user lambdas + autodiff cannot cross the C ABI).
Transcription, QP solve and the warm-start retention rule (asif.hpp:99) all run on the device in one launch; the
per-step host traffic is the state and desired input of every agent in, the filtered input and solver code out.
No CPU path: without libsfb.so / a B200 every call raises SfbError.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import Handle, SfbQpParams
from .qp import QPSolverParams, _is_torch, _ptr, default_handle


class SfbAsifVehicleParams(C.Structure):
    """sfb_asif_vehicle_params, field for field."""

    _fields_ = [
        ("T", C.c_double), ("K", C.c_int32), ("alpha", C.c_double), ("dt", C.c_double), ("relax_cost", C.c_double),
        ("u_weight", C.c_double * 2), ("ulim_l", C.c_double * 2), ("ulim_u", C.c_double * 2),
        ("drag1", C.c_double), ("drag3", C.c_double), ("centre", C.c_double * 2), ("radius", C.c_double),
        ("bu_gain", C.c_double), ("bu_const", C.c_double), ("qp", SfbQpParams),
    ]


@dataclass
class ASIFVehicleParams:
    """ASIFilterParams<U> (asif.hpp:17-32) + ASIFtoQPParams (asif_func.hpp:58-68) + the model constants of
    examples/mpc_asif_vehicle.cpp:42-52,96-129; defaults are that example's values."""

    T: float = 2.5
    K: int = 200
    alpha: float = 5.0
    dt: float = 0.01
    relax_cost: float = 100.0
    u_weight: tuple = (20.0, 1.0)
    ulim_l: tuple = (-0.2, -0.5)
    ulim_u: tuple = (0.5, 0.5)
    drag1: float = 0.2
    drag3: float = 0.4
    centre: tuple = (0.0, -2.3)
    radius: float = 0.7
    bu_gain: float = 0.2
    bu_const: float = -0.5
    qp: QPSolverParams = field(default_factory=lambda: QPSolverParams(polish=False))

    def to_c(self) -> SfbAsifVehicleParams:
        p = SfbAsifVehicleParams()
        p.T, p.K, p.alpha, p.dt, p.relax_cost = self.T, int(self.K), self.alpha, self.dt, self.relax_cost
        for name in ("u_weight", "ulim_l", "ulim_u", "centre"):
            v = getattr(self, name)
            getattr(p, name)[0], getattr(p, name)[1] = float(v[0]), float(v[1])
        p.drag1, p.drag3, p.radius, p.bu_gain, p.bu_const = self.drag1, self.drag3, self.radius, self.bu_gain, self.bu_const
        p.qp = self.qp.to_c()
        return p


class ASIFVehicleFleet:
    """``batch`` ASIFilter<Bundle<SE2, R^3>, R^2, Dyn> objects sharing one parameter set (asif.hpp:40-110).

    ``fleet(x, u_des) -> (u, code, iter)`` is ASIFilter::operator() for every agent: x [B, 7] in smooth's coefficient order
    (x, y, sin, cos, v1, v2, v3), u_des [B, 2].  numpy arrays (host path) or torch CUDA tensors (device path, asynchronous
    on the current stream); dtype must match the fleet's.
    """

    def __init__(self, batch: int, prm: ASIFVehicleParams | None = None, dtype=np.float64, handle: Handle | None = None,
                 device: int = 0):
        self.prm = prm or ASIFVehicleParams()
        self.batch = int(batch)
        self.dtype = np.dtype(dtype)
        assert self.dtype in (np.dtype(np.float64), np.dtype(np.float32))
        self.m = int(self.prm.K) + 3
        self._h = handle or default_handle(device)
        self._f = C.c_void_p()
        cp = self.prm.to_c()
        self._h.check(_lib.lib().sfb_asif_fleet_create(self._h.raw, C.byref(cp), self.batch, self.dtype.itemsize, C.byref(self._f)))

    def reset_warmstart(self) -> None:
        self._h.check(_lib.lib().sfb_asif_fleet_reset_warmstart(self._f))

    def set_warmstart(self, enabled: bool) -> None:
        self._h.check(_lib.lib().sfb_asif_fleet_set_warmstart(self._f, int(enabled)))

    def _prep(self, x, u_des):
        if _is_torch(x):
            import torch

            tdt = torch.float64 if self.dtype == np.float64 else torch.float32
            for t in (x, u_des):
                assert t.is_cuda and t.is_contiguous() and t.dtype == tdt
            self._h.set_stream(torch.cuda.current_stream(x.device).cuda_stream)
            return x, u_des, True
        return np.ascontiguousarray(x, dtype=self.dtype), np.ascontiguousarray(u_des, dtype=self.dtype), False

    def __call__(self, x, u_des):
        x, u_des, tm = self._prep(x, u_des)
        B = self.batch
        assert tuple(x.shape) == (B, 7) and tuple(u_des.shape) == (B, 2)
        if tm:
            import torch

            u = torch.empty((B, 2), dtype=x.dtype, device=x.device)
            st = torch.empty((B,), dtype=torch.int32, device=x.device)
            it = torch.empty((B,), dtype=torch.int32, device=x.device)
        else:
            u = np.empty((B, 2), self.dtype); st = np.empty(B, np.int32); it = np.empty(B, np.uint32)
        L = _lib.lib()
        fn = L.sfb_asif_fleet_filter_f64 if self.dtype == np.float64 else L.sfb_asif_fleet_filter_f32
        self._h.check(fn(self._f, _ptr(x), _ptr(u_des), _ptr(u), _ptr(st), _ptr(it)))
        return u, st, it

    def to_qp(self, x, u_des):
        """asif_to_qp for every agent -> P [B,3,3], q [B,3], A_cm [B,3,m], l, u [B,m] (column-major A, like solve_dense_batch)."""
        assert self.dtype == np.float64
        x, u_des, tm = self._prep(x, u_des)
        B, m = self.batch, self.m
        if tm:
            import torch

            mk = lambda *s: torch.empty(s, dtype=torch.float64, device=x.device)
        else:
            mk = lambda *s: np.empty(s, np.float64)
        P, q, A, l, u = mk(B, 3, 3), mk(B, 3), mk(B, 3, m), mk(B, m), mk(B, m)
        self._h.check(_lib.lib().sfb_asif_fleet_to_qp_f64(self._f, _ptr(x), _ptr(u_des), _ptr(P), _ptr(q), _ptr(A), _ptr(l), _ptr(u)))
        return P, q, A, l, u

    def close(self) -> None:
        if self._f:
            _lib.lib().sfb_asif_fleet_destroy(self._f)
            self._f = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
