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


# ---------------------------------------------------------------------------------------------------------------------
# sparse workloads with a SHARED pattern (sfb_qp_solve_sparse_batch_*)
# ---------------------------------------------------------------------------------------------------------------------
def random_sparse_qp_numpy(B: int, n: int, m: int, density: float = 0.15, seed: int = 5, feasible: bool = True):
    """bench_types.hpp:19-41 at density < 1 with ONE Bernoulli mask for the whole batch (the reference draws a new mask
    per instance; a batch that shares its symbolic factorisation needs a shared one).  P = L L^T is stored with both
    triangles, exactly what qp_dense_to_sparse (bench_types.hpp:47-63, sparseView) produces.

    -> dict(n, m, P_colptr, P_rowidx, A_rowptr, A_colidx), P_vals [B,nnzP], q [B,n], A_vals [B,nnzA], l, u [B,m]
    """
    rng = np.random.Generator(np.random.Philox(key=seed))
    maskA = rng.random((m, n)) < density
    maskL = np.tril(rng.random((n, n)) < density)
    maskL[np.arange(n), np.arange(n)] = True
    A = rng.uniform(-1.0, 1.0, (B, m, n)) * maskA
    L = rng.uniform(-1.0, 1.0, (B, n, n)) * maskL
    i = np.arange(n)
    L[:, i, i] = np.maximum(np.abs(L[:, i, i]), 0.05)
    P = L @ np.transpose(L, (0, 2, 1))
    P = 0.5 * (P + np.transpose(P, (0, 2, 1)))
    maskP = (maskL.astype(np.int64) @ maskL.T.astype(np.int64)) > 0
    v = rng.uniform(-1.0, 1.0, (B, n))
    delta = rng.uniform(0.0, 1.0, (B, m)) if feasible else rng.uniform(-1.0, 1.0, (B, m))
    q = rng.uniform(-1.0, 1.0, (B, n))
    l = np.full((B, m), -np.inf)
    u = np.einsum("bij,bj->bi", A, v) + delta
    # CSC of P, CSR of A
    pc, pr = np.nonzero(maskP.T)  # sorted by column, then row
    P_colptr = np.concatenate([[0], np.cumsum(np.bincount(pc, minlength=n))]).astype(np.int32)
    ar, ac = np.nonzero(maskA)
    A_rowptr = np.concatenate([[0], np.cumsum(np.bincount(ar, minlength=m))]).astype(np.int32)
    pat = dict(n=n, m=m, P_colptr=P_colptr, P_rowidx=pr.astype(np.int32), A_rowptr=A_rowptr, A_colidx=ac.astype(np.int32))
    return pat, np.ascontiguousarray(P[:, pr, pc]), q, np.ascontiguousarray(A[:, ar, ac]), l, u


def _lgr_diff_matrix(K: int):
    """Differentiation matrix of the Lagrange basis on the K Legendre-Gauss-Radau points of [-1, 1) plus the end point 1
    (the mesh the reference collocates on, collocation/mesh.hpp): D [K, K+1], row k = derivative at node k."""
    from numpy.polynomial import legendre as Lg

    cK = np.zeros(K + 1); cK[K] = 1.0
    cK1 = np.zeros(K); cK1[K - 1] = 1.0
    nodes = np.sort(np.real(Lg.legroots(Lg.legadd(cK, cK1))))  # roots of P_K + P_{K-1}: LGR points, -1 included
    tau = np.concatenate([nodes, [1.0]])
    w = np.array([1.0 / np.prod([tau[j] - tau[k] for k in range(K + 1) if k != j]) for j in range(K + 1)])
    D = np.zeros((K + 1, K + 1))
    for i in range(K + 1):
        for j in range(K + 1):
            if i != j:
                D[i, j] = (w[j] / w[i]) / (tau[i] - tau[j])
        D[i, i] = -np.sum(D[i, :])
    # Radau quadrature weights for the cost
    PK1 = Lg.legval(nodes, cK1)
    qw = (1.0 - nodes) / (K * K * PK1 * PK1)
    return D[:K, :], qw, tau


def mpc_structured_pattern(Nx: int = 6, Nu: int = 2, nivals: int = 13, Ki: int = 4):
    """Sparsity pattern of the QP that MPC builds (ocp_to_qp.hpp:40-108) for Nx states, Nu inputs, Ncr = Nu input-bound
    rows per node and Nce = Nx initial-state rows, on a mesh of `nivals` intervals with Ki collocation nodes each.

    Variable layout [x_0 .. x_N, u_0 .. u_{N-1}] (ocp_to_qp.hpp:56); rows: dynamics Nx*N (each with Nx + Ki + Nu entries:
    df/dx block, one differentiation-matrix row, df/du block; :77-80), running constraints Ncr*N (Nx + Nu entries, :81),
    end constraints Nce (2 Nx entries, :82).  P holds the UPPER triangle only (:86-96): per node an Nx x Nx block, for the
    last node additionally the x_0 cross block, for every input node the x_i cross block and an Nu x Nu block.
    BASELINE configs[2] (SE(2) x R^3 bus, K = 50, Kmesh = 4 -> 13 intervals, N = 52): n = m = 422.
    """
    N = nivals * Ki
    Ncr, Nce = Nu, Nx
    nx = Nx * (N + 1)
    n = nx + Nu * N
    rows = []  # per row: list of (col, kind, a, b, c)
    for iv in range(nivals):
        I0 = iv * Ki
        for k in range(Ki):
            i = I0 + k
            for s in range(Nx):
                ent = {}
                for kk in range(Ki + 1):
                    ent.setdefault(Nx * (I0 + kk) + s, []).append(("D", k, kk, 0))
                for t in range(Nx):
                    ent.setdefault(Nx * i + t, []).append(("fx", i, s, t))
                for t in range(Nu):
                    ent.setdefault(nx + Nu * i + t, []).append(("fu", i, s, t))
                rows.append(("dyn", i, s, sorted(ent.items())))
    for i in range(N):
        for r in range(Ncr):
            ent = {Nx * i + t: [("zero", 0, 0, 0)] for t in range(Nx)}
            for t in range(Nu):
                ent[nx + Nu * i + t] = [("one", 0, 0, 0)] if t == r else [("zero", 0, 0, 0)]
            rows.append(("cr", i, r, sorted(ent.items())))
    for r in range(Nce):
        ent = {t: ([("one", 0, 0, 0)] if t == r else [("zero", 0, 0, 0)]) for t in range(Nx)}
        for t in range(Nx):
            ent[Nx * N + t] = [("zero", 0, 0, 0)]
        rows.append(("ce", 0, r, sorted(ent.items())))
    m = len(rows)
    A_rowptr = np.zeros(m + 1, np.int32)
    A_colidx, A_terms = [], []
    for ri, (_, _, _, ents) in enumerate(rows):
        for col, terms in ents:
            A_colidx.append(col)
            A_terms.append(terms)
        A_rowptr[ri + 1] = len(A_colidx)
    # P upper triangle, CSC
    pcols = [[] for _ in range(n)]
    for i in range(N + 1):
        for b in range(Nx):
            col = Nx * i + b
            if i == N:
                pcols[col] += [(t, ("zero", 0, 0, 0)) for t in range(Nx)]  # x_0 / x_N cross block (explicit zeros here)
            pcols[col] += [(Nx * i + a, ("Q", i, a, b)) for a in range(b + 1)]
    for i in range(N):
        for b in range(Nu):
            col = nx + Nu * i + b
            pcols[col] += [(Nx * i + t, ("S", i, t, b)) for t in range(Nx)]
            pcols[col] += [(nx + Nu * i + a, ("R", i, a, b)) for a in range(b + 1)]
    P_colptr = np.zeros(n + 1, np.int32)
    P_rowidx, P_terms = [], []
    for j in range(n):
        for r, term in sorted(pcols[j]):
            P_rowidx.append(r)
            P_terms.append(term)
        P_colptr[j + 1] = len(P_rowidx)
    return dict(n=n, m=m, Nx=Nx, Nu=Nu, N=N, Ki=Ki, nivals=nivals, rows=rows, A_terms=A_terms, P_terms=P_terms,
                P_colptr=P_colptr, P_rowidx=np.asarray(P_rowidx, np.int32), A_rowptr=A_rowptr,
                A_colidx=np.asarray(A_colidx, np.int32))


def mpc_structured_batch(pat, B: int, seed: int = 5, tf: float = 5.0, dtype=np.float64):
    """Values for `mpc_structured_pattern`: a linear time-varying surrogate of the MPC QP (mpc.hpp:458-519) per agent.

    Dynamics rows collocate  x' = Jx_i x + Ju_i u + e_i  on the LGR mesh (real differentiation matrix, interval length
    tf / nivals); Jx_i, Ju_i ~ small random Jacobians per agent and node; cost = Radau-weighted sum of x'Qx + u'Ru with
    diagonal Q, R (1 + 0.2 U(0,1)) -- off-diagonal pattern entries are explicit zeros; |u| <= 0.5 (mpc_asif_vehicle.cpp:60-66);
    x_0 fixed.  The right-hand sides are chosen so that a random trajectory with |u| <= 0.4 is feasible.

    -> P_vals [B,nnzP], q [B,n], A_vals [B,nnzA], l, u [B,m]
    """
    rng = np.random.Generator(np.random.Philox(key=seed))
    Nx, Nu, N, Ki, nivals, n, m = pat["Nx"], pat["Nu"], pat["N"], pat["Ki"], pat["nivals"], pat["n"], pat["m"]
    D, qw, _ = _lgr_diff_matrix(Ki)
    h = 0.5 * tf / nivals
    Jx = 0.5 * rng.uniform(-1, 1, (B, N, Nx, Nx))
    Ju = rng.uniform(-1, 1, (B, N, Nx, Nu))
    nnzA = len(pat["A_terms"])
    A_vals = np.zeros((B, nnzA))
    for e, terms in enumerate(pat["A_terms"]):
        for kind, a, b, c in terms:
            if kind == "D":
                A_vals[:, e] += D[a, b]
            elif kind == "fx":
                A_vals[:, e] -= h * Jx[:, a, b, c]
            elif kind == "fu":
                A_vals[:, e] -= h * Ju[:, a, b, c]
            elif kind == "one":
                A_vals[:, e] += 1.0
    wnode = np.concatenate([np.tile(qw, nivals) * h, [0.1]])  # running-cost quadrature weights, terminal weight 0.1
    # Weights as in the reference's MPC examples (mpc_asif_vehicle.cpp:79-83: Q = I, Qtf = 0.1 I, R = I): diagonal, so
    # the off-diagonal entries of the pattern are explicit zeros.  NB this matters for parity with the reference: its
    # check_stopping multiplies with P AS STORED (upper triangle only, qp_solver.hpp:589), so with a non-diagonal
    # weight its dual residual never vanishes and every solve ends in MaxIterations (verified with the oracle).
    dq = 1.0 + 0.2 * rng.uniform(0, 1, (B, N + 1, Nx))
    Qm = np.einsum("bni,ij->bnij", dq, np.eye(Nx))
    dr = 1.0 + 0.2 * rng.uniform(0, 1, (B, N, Nu))
    Rm = np.einsum("bni,ij->bnij", dr, np.eye(Nu))
    Sm = np.zeros((B, N, Nx, Nu))
    nnzP = len(pat["P_terms"])
    P_vals = np.zeros((B, nnzP))
    for e, (kind, i, a, b) in enumerate(pat["P_terms"]):
        if kind == "Q":
            P_vals[:, e] = wnode[i] * Qm[:, i, a, b]
        elif kind == "R":
            P_vals[:, e] = wnode[i] * Rm[:, i, a, b]
        elif kind == "S":
            P_vals[:, e] = wnode[i] * Sm[:, i, a, b]
    # feasible reference trajectory -> right-hand sides
    var = np.concatenate([rng.uniform(-1, 1, (B, Nx * (N + 1))), rng.uniform(-0.4, 0.4, (B, Nu * N))], axis=1)
    rowsum = np.zeros((B, m))
    rp, ci = pat["A_rowptr"], pat["A_colidx"]
    rid = np.repeat(np.arange(m), np.diff(rp))
    np.add.at(rowsum, (slice(None), rid), A_vals * var[:, ci])
    l = rowsum.copy(); u = rowsum.copy()
    for ri, (kind, _, _, _) in enumerate(pat["rows"]):
        if kind == "cr":
            l[:, ri] = -0.5; u[:, ri] = 0.5
    q = rng.uniform(-1, 1, (B, n)) * np.concatenate([np.repeat(wnode, Nx), np.repeat(wnode[:N], Nu)])
    c = lambda a: np.ascontiguousarray(a, dtype=dtype)
    return c(P_vals), c(q), c(A_vals), c(l), c(u)


def sparse_to_dense(pat, P_vals, A_vals):
    """Densify a shared-pattern batch: -> P [B,n,n] (entries as stored), A [B,m,n] in math layout."""
    B = P_vals.shape[0]
    n, m = pat["n"], pat["m"]
    P = np.zeros((B, n, n)); A = np.zeros((B, m, n))
    pc = np.repeat(np.arange(n), np.diff(pat["P_colptr"]))
    P[:, pat["P_rowidx"], pc] = P_vals
    ar = np.repeat(np.arange(m), np.diff(pat["A_rowptr"]))
    A[:, ar, pat["A_colidx"]] = A_vals
    return P, A


# ---------------------------------------------------------------------------------------------------------------------
# fleet workloads for the built-in vehicle family (sfb_asif_fleet_*, sfb_mpc_fleet_*): SURVEY 8(d) cfg3 / cfg5 sampling
# ---------------------------------------------------------------------------------------------------------------------
def _se2_exp_np(a):
    vx, vy, w = a[..., 0], a[..., 1], a[..., 2]
    small = np.abs(w) < 1e-9
    ws = np.where(small, 1.0, w)
    A = np.where(small, 1.0 - w * w / 6.0, np.sin(ws) / ws)
    Bc = np.where(small, w / 2.0, (1.0 - np.cos(ws)) / ws)
    return np.stack([A * vx - Bc * vy, Bc * vx + A * vy, np.sin(w), np.cos(w)], axis=-1)


def _se2_compose_np(g1, g2):
    x1, y1, s1, c1 = (g1[..., k] for k in range(4))
    x2, y2, s2, c2 = (g2[..., k] for k in range(4))
    return np.stack([x1 + c1 * x2 - s1 * y2, y1 + s1 * x2 + c1 * y2, s1 * c2 + c1 * s2, c1 * c2 - s1 * s2], axis=-1)


def vehicle_xdes_numpy(t, g0=(2.5, 0.0, np.pi / 2), vdes=(1.0, 0.0, 0.4)):
    """Desired trajectory of examples/mpc_asif_vehicle.cpp:72-78: X{SE2(g0) * exp(t vdes), vdes} -> [B, 7] coefficients
    (x, y, sin, cos, v1, v2, v3)."""
    t = np.asarray(t, dtype=np.float64)
    vd = np.asarray(vdes, dtype=np.float64)
    g0c = np.array([g0[0], g0[1], np.sin(g0[2]), np.cos(g0[2])])
    g = _se2_compose_np(np.broadcast_to(g0c, t.shape + (4,)), _se2_exp_np(t[..., None] * vd))
    return np.concatenate([g, np.broadcast_to(vd, t.shape + (3,))], axis=-1)


def vehicle_rplus_numpy(x, a):
    """x (+) a on Bundle<SE2, R^3>: x [B, 7], a [B, 6]."""
    return np.concatenate([_se2_compose_np(x[..., :4], _se2_exp_np(a[..., :3])), x[..., 4:] + a[..., 3:]], axis=-1)


def vehicle_dynamics_numpy(x, u, drag1=0.2, drag3=0.4):
    """d^r x = (v1, v2, v3, -drag1 v1 + u1, 0, -drag3 v3 + u2)   (mpc_asif_vehicle.cpp:42-52)."""
    v = x[..., 4:7]
    return np.stack([v[..., 0], v[..., 1], v[..., 2], -drag1 * v[..., 0] + u[..., 0], np.zeros_like(v[..., 0]),
                     -drag3 * v[..., 2] + u[..., 1]], axis=-1)


def vehicle_fleet_numpy(B: int, seed: int = 5, sigma: float = 0.1):
    """SURVEY 8(d) cfg3 / cfg5 sampling: t0 ~ U(0, 30), x0 = xdes(t0) (+) xi with xi ~ N(0, sigma^2 I_6), and a desired input
    u_des ~ U(-0.5, 0.5)^2 for the safety filter.  -> t0 [B], x0 [B, 7], u_des [B, 2]"""
    rng = np.random.Generator(np.random.Philox(key=seed))
    t0 = rng.uniform(0.0, 30.0, B)
    xi = sigma * rng.normal(size=(B, 6))
    u_des = rng.uniform(-0.5, 0.5, (B, 2))
    return t0, vehicle_rplus_numpy(vehicle_xdes_numpy(t0), xi), u_des
