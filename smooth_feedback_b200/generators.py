"""Synthetic workloads for the hot path (host numpy for parity tests, torch on device for the bench).

random_qp restates the reference's only random-QP recipe, benchmarks/bench_types.hpp:19-41, at density 1.0:
    A_ij ~ U(-1,1);  L = tril(U(-1,1)), L_ii = max(|L_ii|, 0.05);  P = L L^T;  q, v ~ U(-1,1)^n;
    l = -inf;  u = A v + delta
with delta ~ U(0,1) ("G+", feasible by construction: x = v satisfies every row) or the literal
delta ~ U(-1,1) ("G+-", about half of the instances are infeasible at m = 2n).
The reference's std::default_random_engine stream is not reproduced (it carries no parity value);
numpy's counter-based Philox seeded with `seed` is used instead so CPU and GPU sides see identical bits.
"""
from __future__ import annotations

import numpy as np


def random_qp_numpy(B: int, n: int, m: int, seed: int = 5, feasible: bool = True, dtype=np.float64):
    """-> P [B,n,n], q [B,n], A [B,m,n], l [B,m], u [B,m] in math layout."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    A = rng.uniform(-1.0, 1.0, (B, m, n))
    L = np.tril(rng.uniform(-1.0, 1.0, (B, n, n)))
    i = np.arange(n)
    L[:, i, i] = np.maximum(np.abs(L[:, i, i]), 0.05)
    P = L @ np.transpose(L, (0, 2, 1))
    P = 0.5 * (P + np.transpose(P, (0, 2, 1)))  # exactly symmetric, as Eigen's L*L^T product is
    v = rng.uniform(-1.0, 1.0, (B, n))
    delta = rng.uniform(0.0, 1.0, (B, m)) if feasible else rng.uniform(-1.0, 1.0, (B, m))
    q = rng.uniform(-1.0, 1.0, (B, n))
    l = np.full((B, m), -np.inf)
    u = np.einsum("bij,bj->bi", A, v) + delta
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    return c(P), c(q), c(A), c(l), c(u)


def random_qp_torch(B: int, n: int, m: int, seed: int = 5, feasible: bool = True, device="cuda", dtype=None,
                    chunk: int = 8192):
    """Same recipe generated on the device, directly in the engine's column-major storage.

    -> P_cm [B,n,n], q [B,n], A_cm [B,n,m], l [B,m], u [B,m]   (see qp.solve_dense_batch)
    """
    import torch

    dtype = dtype or torch.float64
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    P_cm = torch.empty((B, n, n), dtype=dtype, device=device)
    A_cm = torch.empty((B, n, m), dtype=dtype, device=device)
    q = torch.empty((B, n), dtype=dtype, device=device)
    l = torch.full((B, m), -float("inf"), dtype=dtype, device=device)
    u = torch.empty((B, m), dtype=dtype, device=device)
    idx = torch.arange(n, device=device)
    for b0 in range(0, B, chunk):
        b1 = min(B, b0 + chunk)
        k = b1 - b0
        U = lambda *s: torch.rand(*s, generator=g, device=device, dtype=torch.float64) * 2.0 - 1.0
        A = U(k, m, n)
        L = torch.tril(U(k, n, n))
        L[:, idx, idx] = torch.clamp(L[:, idx, idx].abs(), min=0.05)
        P = L @ L.transpose(1, 2)
        P = 0.5 * (P + P.transpose(1, 2))
        v = U(k, n)
        delta = torch.rand(k, m, generator=g, device=device, dtype=torch.float64)
        if not feasible:
            delta = delta * 2.0 - 1.0
        P_cm[b0:b1] = P.transpose(1, 2).to(dtype)
        A_cm[b0:b1] = A.transpose(1, 2).to(dtype)
        q[b0:b1] = U(k, n).to(dtype)
        u[b0:b1] = (torch.einsum("bij,bj->bi", A, v) + delta).to(dtype)
    return P_cm, q, A_cm, l, u


def random_ekf_numpy(B: int, d: int, ny: int, seed: int = 5):
    """EKF workload of SURVEY section 8(d) cfg4 shape: A = random Jacobian, Q = 0.01 I, R = 0.01 I, P0 = M M^T + 0.1 I."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    A = rng.normal(size=(B, d, d))
    M = rng.uniform(-1.0, 1.0, (B, d, d))
    P = M @ np.transpose(M, (0, 2, 1)) + 0.1 * np.eye(d)
    P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
    Q = np.broadcast_to(0.01 * np.eye(d), (B, d, d)).copy()
    H = rng.normal(size=(B, ny, d))
    R = np.broadcast_to(0.01 * np.eye(ny), (B, ny, ny)).copy()
    innov = rng.normal(size=(B, ny)) * 0.1
    return P, A, Q, H, R, innov
