"""Host-side mirror of smooth_feedback's EKF (include/smooth/feedback/ekf.hpp:45-139) on top of the C ABI.

The reference's predict/update take user lambdas that are differentiated on the host; the engine takes over
after linearisation: given A = -ad(f) + d^r f/dx (predict) or H, R and the innovation (update) it does the
covariance algebra for a whole batch of filters on the GPU.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from ._lib import Handle
from .qp import _ptr, default_handle

STEPPERS = {"euler": 0, "rk4": 1}


def ekf_predict_batch(P_cm, A_cm, Q_cm, tau: float, dt: float | None = None, stepper: str = "euler",
                      handle: Handle | None = None, out=None):
    """Covariance half of EKF::predict (ekf.hpp:79-103) for B filters.

    P_cm, A_cm, Q_cm: contiguous float64 CUDA tensors [B, d, d] in column-major storage (X_cm[b, j, i] = X_ij).
    dt=None reproduces the reference default (one step of length tau).
    """
    import torch

    B, d, _ = P_cm.shape
    dev = P_cm.device
    for t in (P_cm, A_cm, Q_cm):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float64 and tuple(t.shape) == (B, d, d)
    h = handle or default_handle(dev.index or 0)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    if out is None:
        out = torch.empty_like(P_cm)
    h.check(_lib.lib().sfb_ekf_predict_batch_f64(h.raw, B, d, STEPPERS[stepper], _ptr(P_cm), _ptr(A_cm), _ptr(Q_cm),
                                                 float(tau), -1.0 if dt is None else float(dt), _ptr(out)))
    return out


def ekf_update_batch(P_cm, H_cm, R_cm, innov, handle: Handle | None = None, out_delta=None, out_P=None):
    """Algebra of EKF::update (ekf.hpp:116-139) for B filters.

    P_cm [B,d,d], H_cm [B,d,ny] (H_cm[b, j, i] = H_ij), R_cm [B,ny,ny], innov [B,ny] = y (-) h(g_hat).
    Returns (delta [B,d] = K innov, P_new_cm [B,d,d]); the caller applies g_hat (+) delta.
    """
    import torch

    B, d, _ = P_cm.shape
    ny = innov.shape[1]
    dev = P_cm.device
    assert tuple(H_cm.shape) == (B, d, ny) and tuple(R_cm.shape) == (B, ny, ny)
    for t in (P_cm, H_cm, R_cm, innov):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float64
    h = handle or default_handle(dev.index or 0)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    if out_delta is None:
        out_delta = torch.empty((B, d), dtype=torch.float64, device=dev)
    if out_P is None:
        out_P = torch.empty_like(P_cm)
    h.check(_lib.lib().sfb_ekf_update_batch_f64(h.raw, B, d, ny, _ptr(P_cm), _ptr(H_cm), _ptr(R_cm), _ptr(innov),
                                                _ptr(out_delta), _ptr(out_P)))
    return out_delta, out_P


def ekf_step_batch(P_cm, A_cm, Q_cm, tau: float, H_cm, R_cm, innov, dt: float | None = None, stepper: str = "euler",
                   handle: Handle | None = None, out_delta=None, out_P=None):
    """One predict + update cycle of the covariance (ekf.hpp:79-139) in a single pass over HBM.

    Same arguments as ekf_predict_batch followed by ekf_update_batch; H_cm and innov are evaluated by the caller at
    the predicted estimate.  Returns (delta [B,d], P_new_cm [B,d,d]).
    """
    import torch

    B, d, _ = P_cm.shape
    ny = innov.shape[1]
    dev = P_cm.device
    assert tuple(A_cm.shape) == (B, d, d) and tuple(Q_cm.shape) == (B, d, d)
    assert tuple(H_cm.shape) == (B, d, ny) and tuple(R_cm.shape) == (B, ny, ny)
    for t in (P_cm, A_cm, Q_cm, H_cm, R_cm, innov):
        assert t.is_cuda and t.is_contiguous() and t.dtype == torch.float64
    h = handle or default_handle(dev.index or 0)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    if out_delta is None:
        out_delta = torch.empty((B, d), dtype=torch.float64, device=dev)
    if out_P is None:
        out_P = torch.empty_like(P_cm)
    h.check(_lib.lib().sfb_ekf_step_batch_f64(h.raw, B, d, ny, STEPPERS[stepper], _ptr(P_cm), _ptr(A_cm), _ptr(Q_cm),
                                              float(tau), -1.0 if dt is None else float(dt), _ptr(H_cm), _ptr(R_cm),
                                              _ptr(innov), _ptr(out_delta), _ptr(out_P)))
    return out_delta, out_P


def ekf_step_batch_host(P_cm, A_cm, Q_cm, tau: float, H_cm, R_cm, innov, dt: float | None = None, stepper: str = "euler",
                        handle: Handle | None = None):
    """`ekf_step_batch` on contiguous float64 NUMPY arrays (host path of the C ABI: staged through the engine, returns with
    the results in place).  Same layouts as `ekf_step_batch`."""
    import numpy as np

    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (P_cm, A_cm, Q_cm, H_cm, R_cm, innov)]
    P_cm, A_cm, Q_cm, H_cm, R_cm, innov = arrs
    B, d, _ = P_cm.shape
    ny = innov.shape[1]
    h = handle or default_handle(0)
    out_delta = np.empty((B, d)); out_P = np.empty((B, d, d))
    h.check(_lib.lib().sfb_ekf_step_batch_f64(h.raw, B, d, ny, STEPPERS[stepper], _ptr(P_cm), _ptr(A_cm), _ptr(Q_cm),
                                              float(tau), -1.0 if dt is None else float(dt), _ptr(H_cm), _ptr(R_cm),
                                              _ptr(innov), _ptr(out_delta), _ptr(out_P)))
    return out_delta, out_P
