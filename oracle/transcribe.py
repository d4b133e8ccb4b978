"""numpy restatement of the reference's QP *transcriptions* (the callers either side of the QP hot path).

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the checker for the engine's on-device transcription
kernels (sfb_asif_*, sfb_mpc_*) and the generator of the golden fixtures.  Nothing under smooth_feedback_b200/
imports this module.

Restates (file:line into /root/reference):
  include/smooth/feedback/asif_func.hpp:104-199   asif_to_qp_update   (ASIFilter::operator(), asif.hpp:82-102)
  include/smooth/feedback/ocp_to_qp.hpp:40-400    ocp_to_qp_allocate / update_{cost,dyn,cr,ce}
  include/smooth/feedback/mpc.hpp:22-302,405-519  MPC functors (MPCObj, MPCIntegrand, MPCCE ...), ctor, operator()
  include/smooth/feedback/collocation/mesh.hpp    Mesh<Kmin,Kmax>: LGR nodes / weights / differentiation matrices
  include/smooth/feedback/collocation/mesh_function.hpp:285-399   mesh_integrate (2nd-order part used by the cost)
for model families given through callbacks that return ANALYTIC derivatives (what the reference's forward-mode
autodiff computes up to rounding), plus the model family of examples/mpc_asif_vehicle.cpp:42-129.

Third-party pieces that are absent from /root/reference and therefore restated from their published definitions
("parity unpinned"): pettni/smooth SE(2) / Bundle group operations (exp, log, rplus, rminus, ad,
dr_expinv; coefficient order x, y, qz = sin, qw = cos; tangent order x, y, theta), smooth::lgr_nodes (Legendre-Gauss-
Radau nodes and weights), Boost.odeint euler with vector_space_algebra (x <- x (+) dt * f).

Reference quirks reproduced on purpose (each verified in the cited lines):
  * asif_func.hpp:170-176  the inner step size dt_act is computed ONCE per constraint interval, before the while loop,
    so a trajectory overshoots tau*(k+1) by up to one step; the state stepper runs first and the sensitivity ODE then
    linearises around the ALREADY-STEPPED state (dx_dx0_ode captures x by reference).
  * mpc.hpp:423 vs :481-488  the cost (P, q) is transcribed once in the constructor; MPC::set_weights called afterwards
    (as examples/mpc_asif_vehicle.cpp:79-83 does) never reaches the QP.  The example therefore runs with the default
    weights Q = I, R = I, Qtf = I.
  * mpc.hpp:103-107 + ocp_to_qp.hpp:190  MPCObj::hessian writes Qtf into the x0 block (rows/cols 1..Nx of (tf, x0, xf,
    q)), so the "terminal" weight lands on x_0 (scaled 0.5) and x_N carries no cost at all.
  * ocp_to_qp.hpp:86-96 / block_add(..., upper_only=true): only the upper triangle of P is stored.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

INF = np.inf


# ---------------------------------------------------------------------------------------------------------------------
# SE(2) and Bundle<SE(2), R^k> (pettni/smooth conventions, restated)
# ---------------------------------------------------------------------------------------------------------------------
def se2_exp(a):
    """exp of a tangent (vx, vy, w) -> coefficients (x, y, sin, cos)."""
    vx, vy, w = a
    if abs(w) < 1e-9:  # series (smooth switches to a Taylor expansion near zero as well)
        A = 1.0 - w * w / 6.0
        B = w / 2.0 - w ** 3 / 24.0
    else:
        A = np.sin(w) / w
        B = (1.0 - np.cos(w)) / w
    return np.array([A * vx - B * vy, B * vx + A * vy, np.sin(w), np.cos(w)])


def se2_compose(g1, g2):
    x1, y1, s1, c1 = g1
    x2, y2, s2, c2 = g2
    return np.array([x1 + c1 * x2 - s1 * y2, y1 + s1 * x2 + c1 * y2, s1 * c2 + c1 * s2, c1 * c2 - s1 * s2])


def se2_inverse(g):
    x, y, s, c = g
    return np.array([-(c * x + s * y), -(-s * x + c * y), -s, c])


def se2_log(g):
    x, y, s, c = g
    w = np.arctan2(s, c)
    if abs(w) < 1e-9:
        A = 1.0 - w * w / 12.0
    else:
        A = 0.5 * w / np.tan(0.5 * w)  # (w/2) cot(w/2), half-angle form: no cancellation for small w
    B = w / 2.0
    return np.array([A * x + B * y, -B * x + A * y, w])


def se2_ad(a):
    vx, vy, w = a
    return np.array([[0.0, -w, vy], [w, 0.0, -vx], [0.0, 0.0, 0.0]])


def se2_dr_expinv(a):
    """inverse of the right Jacobian of exp at a:  I + ad/2 + c2 ad^2,  c2 = 1/w^2 - (1 + cos w) / (2 w sin w)."""
    w = a[2]
    ad = se2_ad(a)
    if abs(w) < 1e-5:
        c2 = 1.0 / 12.0 + w * w / 720.0
    else:
        c2 = 1.0 / (w * w) - (1.0 + np.cos(w)) / (2.0 * w * np.sin(w))
    return np.eye(3) + 0.5 * ad + c2 * (ad @ ad)


class BundleSE2Rk:
    """smooth::Bundle<SE2d, Eigen::Vector<double, k>>: coefficients (x, y, sin, cos, r_1..r_k), tangent (vx, vy, w, r')."""

    def __init__(self, k: int):
        self.k = k
        self.dof = 3 + k
        self.rep = 4 + k

    def identity(self):
        g = np.zeros(self.rep)
        g[3] = 1.0
        return g

    def rplus(self, g, a):
        return np.concatenate([se2_compose(g[:4], se2_exp(a[:3])), g[4:] + a[3:]])

    def rminus(self, g1, g2):
        return np.concatenate([se2_log(se2_compose(se2_inverse(g2[:4]), g1[:4])), g1[4:] - g2[4:]])

    def ad(self, a):
        M = np.zeros((self.dof, self.dof))
        M[:3, :3] = se2_ad(a[:3])
        return M

    def dr_expinv(self, a):
        M = np.eye(self.dof)
        M[:3, :3] = se2_dr_expinv(a[:3])
        return M

    commutative = False


class Rn:
    """Eigen::Vector<double, n> as a (commutative) Lie group."""

    def __init__(self, n: int):
        self.dof = n
        self.rep = n

    def identity(self):
        return np.zeros(self.dof)

    def rplus(self, g, a):
        return g + a

    def rminus(self, g1, g2):
        return g1 - g2

    def ad(self, a):
        return np.zeros((self.dof, self.dof))

    def dr_expinv(self, a):
        return np.eye(self.dof)

    commutative = True


class SE2(BundleSE2Rk):
    def __init__(self):
        super().__init__(0)


# ---------------------------------------------------------------------------------------------------------------------
# asif_to_qp_update, asif_func.hpp:104-199
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class ASIFtoQPParams:  # asif_func.hpp:58-68
    K: int = 10
    alpha: float = 1.0
    dt: float = 0.1
    relax_cost: float = 100.0


@dataclass
class ASIFProblem:  # asif_func.hpp:40-53 (+ ManifoldBounds, common.hpp:20-32)
    T: float
    x0: np.ndarray
    u_des: np.ndarray
    W_u: np.ndarray
    ulim_A: np.ndarray
    ulim_c: np.ndarray
    ulim_l: np.ndarray
    ulim_u: np.ndarray


def asif_step_schedule(T: float, K: int, dt_max: float):
    """The (state-independent) sequence of Euler steps asif_to_qp_update takes, asif_func.hpp:139-143,170-176.

    -> (t_k [K] time at which constraint k is evaluated, steps: list of K lists of step lengths)"""
    tau = T / float(K)
    dt = min(dt_max, tau)
    t = 0.0
    tk, steps = [], []
    for k in range(K):
        tk.append(t)
        dt_act = min(dt, tau * (k + 1) - t)
        s = []
        while t < tau * (k + 1):
            s.append(dt_act)
            t += dt_act
        steps.append(s)
    return np.array(tk), steps


def asif_to_qp(group, pbm: ASIFProblem, prm: ASIFtoQPParams, f_u, f_cl, h, nh: int):
    """Dense QP of asif_to_qp_update.

    f_u(x, u)  -> (f(x, u), d f / d u  [nx, nu])                                   (:146-147)
    f_cl(t, x) -> (f(x, bu(t, x)), d^r/dx [f(x, bu(t, x))]  [nx, nx])              (:130-134)
    h(t, x)    -> (h [nh], dh/dt [nh], d^r h / dx [nh, nx])                        (:152-156)
    -> P [N,N], q [N], A [M,N], l [M], u [M]   with N = nu + 1, M = K nh + nu_ineq + 1
    """
    nx = group.dof
    nu = len(pbm.u_des)
    nu_ineq = pbm.ulim_A.shape[0]
    K = prm.K
    M, N = K * nh + nu_ineq + 1, nu + 1
    A = np.zeros((M, N)); l = np.zeros(M); u = np.zeros(M)
    P = np.zeros((N, N)); q = np.zeros(N)

    tau = pbm.T / float(K)
    dt = min(prm.dt, tau)
    t = 0.0
    x = np.array(pbm.x0, dtype=np.float64)
    S = np.eye(nx)
    f0, d_f0_du = f_u(x, pbm.u_des)
    for k in range(K):
        hval, dh_dt, dh_dx = h(t, x)
        dh_dx0 = dh_dx @ S
        A[k * nh:(k + 1) * nh, :nu] = dh_dx0 @ d_f0_du
        l[k * nh:(k + 1) * nh] = -dh_dt - prm.alpha * hval - dh_dx0 @ f0
        u[k * nh:(k + 1) * nh] = INF
        dt_act = min(dt, tau * (k + 1) - t)
        while t < tau * (k + 1):
            fx, _ = f_cl(t, x)
            x = group.rplus(x, dt_act * fx)                  # state stepper first (:173)
            fcl, dfcl = f_cl(t, x)                           # sensitivity ODE sees the stepped state (:130-134,174)
            S = S + dt_act * ((-group.ad(fcl) + dfcl) @ S)
            t += dt_act
    A[:K * nh, nu] = 1.0                                                  # :180
    A[K * nh:K * nh + nu_ineq, :nu] = pbm.ulim_A                          # :183
    e = pbm.u_des - pbm.ulim_c                                            # rminus on R^nu
    l[K * nh:K * nh + nu_ineq] = pbm.ulim_l - pbm.ulim_A @ e              # :184
    u[K * nh:K * nh + nu_ineq] = pbm.ulim_u - pbm.ulim_A @ e              # :185
    A[K * nh + nu_ineq, nu] = 1.0                                         # :188-190
    l[K * nh + nu_ineq] = 0.0
    u[K * nh + nu_ineq] = INF
    P[:nu, :nu] = np.diag(pbm.W_u)                                        # :192
    P[nu, nu] = prm.relax_cost                                            # :194
    return P, q, A, l, u


# ---- the model family of examples/mpc_asif_vehicle.cpp:42-129 ------------------------------------------------------
@dataclass
class VehicleModel:
    """SE(2) x R^3 "bus": d^r x = (v1, v2, v3, -drag1 v1 + u1, 0, -drag3 v3 + u2)   (mpc_asif_vehicle.cpp:42-52);
    safe set h = |p - centre| - radius (:96-100); backup controller bu = (bu_gain v1, bu_const) (:103)."""
    drag1: float = 0.2
    drag3: float = 0.4
    centre: tuple = (0.0, -2.3)
    radius: float = 0.7
    bu_gain: float = 0.2
    bu_const: float = -0.5
    group: BundleSE2Rk = field(default_factory=lambda: BundleSE2Rk(3))

    def f(self, x, u):
        v = x[4:7]
        return np.array([v[0], v[1], v[2], -self.drag1 * v[0] + u[0], 0.0, -self.drag3 * v[2] + u[1]])

    def f_u(self, x, u):
        B = np.zeros((6, 2)); B[3, 0] = 1.0; B[5, 1] = 1.0
        return self.f(x, u), B

    def df_dx(self):
        J = np.zeros((6, 6))
        J[0, 3] = J[1, 4] = J[2, 5] = 1.0
        J[3, 3] = -self.drag1
        J[5, 5] = -self.drag3
        return J

    def bu(self, t, x):
        return np.array([self.bu_gain * x[4], self.bu_const])

    def f_cl(self, t, x):
        J = self.df_dx()
        J[3, 3] = -self.drag1 + self.bu_gain  # d/dv1 [-drag1 v1 + bu_gain v1], summed like the autodiff chain rule does
        return self.f(x, self.bu(t, x)), J

    def h(self, t, x):
        d = x[:2] - np.asarray(self.centre)
        e = d / np.sqrt(d[0] * d[0] + d[1] * d[1])          # .normalized() of the double cast: a constant for autodiff
        s, c = x[2], x[3]
        dh = np.zeros((1, 6))
        dh[0, 0] = e[0] * c + e[1] * s                       # e^T R: d p / d a_xy = R (right derivative)
        dh[0, 1] = -e[0] * s + e[1] * c
        return np.array([d @ e - self.radius]), np.zeros(1), dh


def vehicle_asif_problem(x0, u_des):
    """ASIFilterParams of mpc_asif_vehicle.cpp:105-129 applied to state x0 and desired input u_des."""
    pbm = ASIFProblem(T=2.5, x0=np.asarray(x0, float), u_des=np.asarray(u_des, float), W_u=np.array([20.0, 1.0]),
                      ulim_A=np.eye(2), ulim_c=np.zeros(2), ulim_l=np.array([-0.2, -0.5]), ulim_u=np.array([0.5, 0.5]))
    prm = ASIFtoQPParams(K=200, alpha=5.0, dt=0.01, relax_cost=100.0)
    return pbm, prm


def vehicle_asif_qp_batch(x0, u_des, model: VehicleModel | None = None):
    """x0 [B,7], u_des [B,2] -> P [B,3,3], q [B,3], A [B,203,3], l, u [B,203] (math layout)."""
    model = model or VehicleModel()
    B = x0.shape[0]
    out = [[], [], [], [], []]
    for b in range(B):
        pbm, prm = vehicle_asif_problem(x0[b], u_des[b])
        r = asif_to_qp(model.group, pbm, prm, model.f_u, model.f_cl, model.h, 1)
        for o, v in zip(out, r):
            o.append(v)
    return tuple(np.stack(o) for o in out)


# ---------------------------------------------------------------------------------------------------------------------
# Mesh<Kmin, Kmax>, collocation/mesh.hpp
# ---------------------------------------------------------------------------------------------------------------------
def lgr_nodes(K: int):
    """K Legendre-Gauss-Radau nodes on [-1, 1) (roots of P_{K-1} + P_K, -1 included) and their quadrature weights
    w_i = (1 - x_i) / (K^2 P_{K-1}(x_i)^2)   (smooth::lgr_nodes, restated from the published definition)."""
    from numpy.polynomial import legendre as Lg

    cK = np.zeros(K + 1); cK[K] = 1.0
    cK1 = np.zeros(K); cK1[K - 1] = 1.0
    x = np.sort(np.real(Lg.legroots(Lg.legadd(cK, cK1))))
    for _ in range(3):  # Newton polish of the companion-matrix roots
        fx = Lg.legval(x, cK) + Lg.legval(x, cK1)
        dfx = Lg.legval(x, Lg.legder(cK)) + Lg.legval(x, Lg.legder(cK1))
        x = x - fx / dfx
    x[0] = -1.0
    w = (1.0 - x) / (K * K * Lg.legval(x, cK1) ** 2)
    return x, w


def lagrange_diffmat(nodes):
    """D[j, i] = l_j'(nodes[i]): derivative of the j-th Lagrange basis polynomial at node i (barycentric form)."""
    n = len(nodes)
    wb = np.array([1.0 / np.prod([nodes[j] - nodes[k] for k in range(n) if k != j]) for j in range(n)])
    D = np.zeros((n, n))
    for i in range(n):
        for j in range(n):
            if i != j:
                D[j, i] = (wb[j] / wb[i]) / (nodes[i] - nodes[j])
        D[i, i] = -sum(D[j, i] for j in range(n) if j != i)
    return D


class Mesh:
    """Mesh<Kmin, Kmax> (mesh.hpp:69-488): a list of intervals (K, tau0) on [0, 1]."""

    def __init__(self, n: int = 1, k: int = 5, Kmin: int | None = None, Kmax: int | None = None):
        self.Kmin = Kmin if Kmin is not None else k
        self.Kmax = Kmax if Kmax is not None else k
        if n < 2:
            self.intervals = [[k, 0.0]]
        else:
            dx = 1.0 / float(n)
            self.intervals = [[k, float(i) * dx] for i in range(n)]  # mesh.hpp:97-99

    def refine_ph(self, i: int, D: int):  # mesh.hpp:125-148
        K, tau0 = self.intervals[i]
        if D > self.Kmax or K > self.Kmax:
            n = max(2, (D + self.Kmin - 1) // self.Kmin)
            tauf = self.intervals[i + 1][1] if i + 1 < len(self.intervals) else 1.0
            taum = (tauf - tau0) / float(n)
            while n > 1:
                n -= 1
                self.intervals.insert(i + 1, [self.Kmin, tau0 + float(n) * taum])
            # NB: the reference leaves the split interval's own degree unchanged
        elif D < K:
            return
        elif D <= self.Kmax:
            self.intervals[i][0] = D

    def N_ivals(self):
        return len(self.intervals)

    def N_colloc_ival(self, i):
        return self.intervals[i][0]

    def N_colloc(self):
        return sum(iv[0] for iv in self.intervals)

    def _span(self, i):
        tau0 = self.intervals[i][1]
        tauf = self.intervals[i + 1][1] if i + 1 < len(self.intervals) else 1.0
        return tau0, tauf

    def interval_nodes(self, i):  # mesh.hpp:182-206: K + 1 values (the LGR nodes plus the interval end)
        K = self.intervals[i][0]
        x, _ = lgr_nodes(K)
        ext = np.concatenate([x, [1.0]])
        tau0, tauf = self._span(i)
        al = (tauf - tau0) / 2
        return tau0 + al * (ext + 1)

    def interval_weights(self, i):  # mesh.hpp:230-254
        K = self.intervals[i][0]
        _, w = lgr_nodes(K)
        ext = np.concatenate([w, [0.0]])
        tau0, tauf = self._span(i)
        return (tauf - tau0) / 2 * ext

    def all_nodes(self):  # mesh.hpp:213-224: N + 1 values
        n = self.N_ivals()
        return np.concatenate([self.interval_nodes(i)[: self.intervals[i][0] + (1 if i + 1 == n else 0)] for i in range(n)])

    def all_weights(self):
        n = self.N_ivals()
        return np.concatenate([self.interval_weights(i)[: self.intervals[i][0] + (1 if i + 1 == n else 0)] for i in range(n)])

    def interval_diffmat_unscaled(self, i):  # mesh.hpp:341-366 -> (alpha, Dus [K+1, K]),  Dus(j, i) = l_j'(tau_i)
        K = self.intervals[i][0]
        x, _ = lgr_nodes(K)
        ext = np.concatenate([x, [1.0]])
        tau0, tauf = self._span(i)
        return 2.0 / (tauf - tau0), lagrange_diffmat(ext)[:, :K]


# ---------------------------------------------------------------------------------------------------------------------
# ocp_to_qp, ocp_to_qp.hpp:40-400  (variable layout [x_0 .. x_N, u_0 .. u_{N-1}], rows [dyn | cr | ce])
# ---------------------------------------------------------------------------------------------------------------------
class TripletQP:
    """A sparse QP under construction: dict-of-keys for P (upper triangle) and A, dense q, l, u.  Explicit zeros are
    kept (Eigen's coeffRef inserts them), insertion by coeffRef semantics (+= / =)."""

    def __init__(self, nvar, ncon):
        self.n, self.m = nvar, ncon
        self.P = {}
        self.A = {}
        self.q = np.zeros(nvar); self.l = np.zeros(ncon); self.u = np.zeros(ncon)

    def add(self, M, r, c, v):
        M[(r, c)] = M.get((r, c), 0.0) + v

    def block_add(self, M, r0, c0, B, scale=1.0, upper_only=False, mask=None):
        # utils/sparse.hpp:21-37: every stored entry of the source (a dense source stores all of them)
        B = np.atleast_2d(B)
        for j in range(B.shape[1]):
            for i in range(B.shape[0]):
                if mask is not None and not mask[i, j]:
                    continue
                if (not upper_only) or (r0 + i <= c0 + j):
                    self.add(M, r0 + i, c0 + j, scale * B[i, j])

    def block_write(self, M, r0, c0, B, scale=1.0, mask=None):
        B = np.atleast_2d(B)
        for j in range(B.shape[1]):
            for i in range(B.shape[0]):
                if mask is not None and not mask[i, j]:
                    continue
                M[(r0 + i, c0 + j)] = scale * B[i, j]

    def zero_rows(self, r0, nr):  # set_zero(qp.A.middleRows(...)): values to zero, pattern kept
        for key in self.A:
            if r0 <= key[0] < r0 + nr:
                self.A[key] = 0.0

    def dense(self):
        P = np.zeros((self.n, self.n)); A = np.zeros((self.m, self.n))
        for (r, c), v in self.P.items():
            P[r, c] = v
        for (r, c), v in self.A.items():
            A[r, c] = v
        return P, A

    def csc_P(self):
        keys = sorted(self.P.keys(), key=lambda rc: (rc[1], rc[0]))
        colptr = np.zeros(self.n + 1, np.int32)
        for _, c in keys:
            colptr[c + 1] += 1
        return np.cumsum(colptr).astype(np.int32), np.array([r for r, _ in keys], np.int32), np.array([self.P[k] for k in keys])

    def csr_A(self):
        keys = sorted(self.A.keys())
        rowptr = np.zeros(self.m + 1, np.int32)
        for r, _ in keys:
            rowptr[r + 1] += 1
        return np.cumsum(rowptr).astype(np.int32), np.array([c for _, c in keys], np.int32), np.array([self.A[k] for k in keys])


def ocp_dims(mesh: Mesh, Nx, Nu, Ncr, Nce):
    N = mesh.N_colloc()
    return dict(N=N, xvar_L=Nx * (N + 1), uvar_L=Nu * N, dcon_L=Nx * N, crcon_L=Ncr * N, cecon_L=Nce,
                Nvar=Nx * (N + 1) + Nu * N, Ncon=Nx * N + Ncr * N + Nce)


def ocp_to_qp_update_dyn(qp: TripletQP, group, mesh: Mesh, tf, xl_fun, ul_fun, f_fun, Nx, Nu):
    """ocp_to_qp.hpp:198-276.   xl_fun(t) -> (x_l, d^r x_l / dt);  f_fun(t, x, u) -> (f, df/dx [Nx,Nx], df/du [Nx,Nu])."""
    d = ocp_dims(mesh, Nx, Nu, 0, 0)
    N, xvar_L = d["N"], d["xvar_L"]
    uvar_B = xvar_L
    t0 = 0.0
    qp.zero_rows(0, Nx * N)
    M = 0
    for ival in range(mesh.N_ivals()):
        Ki = mesh.N_colloc_ival(ival)
        alpha, Dus = mesh.interval_diffmat_unscaled(ival)
        nodes = mesh.interval_nodes(ival)
        for i in range(Ki):
            t_i = t0 + (tf - t0) * nodes[i]
            xl_i, dxl_i = xl_fun(t_i)
            ul_i = ul_fun(t_i)
            f_i, dfx, dfu = f_fun(t_i, xl_i, ul_i)
            r0 = (M + i) * Nx
            qp.block_add(qp.A, r0, (M + i) * Nx, dfx, tf)                          # :251
            qp.block_add(qp.A, r0, uvar_B + (M + i) * Nu, dfu, tf)                # :252
            if not group.commutative:
                qp.block_add(qp.A, r0, (M + i) * Nx, group.ad(f_i + dxl_i), -tf / 2)  # :255-257
            for j in range(Ki + 1):
                for dg in range(Nx):
                    qp.add(qp.A, r0 + dg, (M + j) * Nx + dg, -(alpha * Dus[j, i]))   # :259-263
            qp.l[r0:r0 + Nx] = -tf * (f_i - dxl_i)                                 # :265
            qp.u[r0:r0 + Nx] = qp.l[r0:r0 + Nx]
        M += Ki


def ocp_to_qp_update_cr(qp: TripletQP, mesh: Mesh, tf, xl_fun, ul_fun, cr_fun, crl, cru, Nx, Nu):
    """ocp_to_qp.hpp:279-323 (mesh_eval<1>: F, dF per node, mesh_function.hpp:62-170 with scale = false).
    cr_fun(t, x, u) -> (c, dc/dx [Ncr,Nx], dc/du [Ncr,Nu])."""
    Ncr = len(crl)
    d = ocp_dims(mesh, Nx, Nu, Ncr, 0)
    N, xvar_L, crcon_B = d["N"], d["xvar_L"], d["dcon_L"]
    nodes = mesh.all_nodes()
    for i in range(N):
        t_i = tf * nodes[i]
        xl_i, _ = xl_fun(t_i)
        c, dcx, dcu = cr_fun(t_i, xl_i, ul_fun(t_i))
        r0 = crcon_B + i * Ncr
        qp.block_write(qp.A, r0, i * Nx, dcx)
        qp.block_write(qp.A, r0, xvar_L + i * Nu, dcu)
        qp.l[r0:r0 + Ncr] = crl - c
        qp.u[r0:r0 + Ncr] = cru - c


def ocp_to_qp_update_ce(qp: TripletQP, mesh: Mesh, tf, xl_fun, ce_fun, cel, ceu, Nx, Nu, Ncr):
    """ocp_to_qp.hpp:326-373.  ce_fun(tf, x0, xf) -> (ce, dce/dx0 [Nce,Nx] or None, dce/dxf or None, masks)"""
    Nce = len(cel)
    d = ocp_dims(mesh, Nx, Nu, Ncr, Nce)
    cecon_B = d["dcon_L"] + d["crcon_L"]
    xl0, _ = xl_fun(0.0)
    xlf, _ = xl_fun(tf)
    ce, d0, df_, m0, mf = ce_fun(tf, xl0, xlf)
    if d0 is not None:
        qp.block_write(qp.A, cecon_B, 0, d0, mask=m0)                              # :366
    if df_ is not None:
        qp.block_write(qp.A, cecon_B, d["xvar_L"] - Nx, df_, mask=mf)              # :367
    qp.l[cecon_B:cecon_B + Nce] = cel - ce
    qp.u[cecon_B:cecon_B + Nce] = ceu - ce


def mpc_update_cost(qp: TripletQP, mesh: Mesh, tf, Q, R, Qtf, Nx, Nu):
    """ocp_to_qp_update_cost (ocp_to_qp.hpp:111-195) specialised to the MPC functors (mpc.hpp:60-232): integrand with
    zero value / gradient and Hessian blockdiag(Q, R) restricted to non-zero Q(i,j) (R entries are gated by Q(i,j) != 0,
    mpc.hpp:219-223); end cost with zero gradient w.r.t. x and its Hessian Qtf stored in the x0 block (mpc.hpp:103-107)."""
    d = ocp_dims(mesh, Nx, Nu, 0, 0)
    N, xvar_L = d["N"], d["xvar_L"]
    for key in qp.P:
        qp.P[key] = 0.0
    qp.q[:] = 0.0
    nodes, weights = mesh.all_nodes(), mesh.all_weights()
    for i in range(N):  # mesh_integrate<2>: xx and uu blocks, wl (tf - t0), upper only (mesh_function.hpp:385-392)
        w = weights[i]
        for a in range(Nx):
            for b in range(Nx):
                if Q[a, b] != 0 and a <= b:
                    qp.add(qp.P, i * Nx + a, i * Nx + b, 1.0 * ((w * 1.0) * (tf - 0.0)) * Q[a, b])
        for a in range(Nu):
            for b in range(Nu):
                if Q[a, b] != 0 and a <= b:
                    qp.add(qp.P, xvar_L + i * Nu + a, xvar_L + i * Nu + b, 1.0 * ((w * 1.0) * (tf - 0.0)) * R[a, b])
    for a in range(Nx):  # ocp_to_qp.hpp:190: d2th.block(1, 1, Nx, Nx) * 0.5 into the x0 block
        for b in range(Nx):
            if Qtf[a, b] != 0 and a <= b:
                qp.add(qp.P, a, b, 0.5 * Qtf[a, b])
    # :191-192 add the (x0,xf) and (xf,xf) blocks of the end-cost Hessian: structurally empty for MPCObj


class MPCRestated:
    """MPC<T, X, U, F, CR, Kmesh = 4> (mpc.hpp:372-638) for a model given by analytic callbacks.

    f_fun(t_abs, x, u) -> (f, df/dx, df/du);  cr = u (the vehicle example's running constraint; Jacobian [0 | I] with
    all Ncr x (Nx + Nu) entries stored, as block_write of a dense autodiff Jacobian does).
    xdes(t_abs) -> (x_des, d^r x_des / dt), udes(t_abs) -> u_des.
    """

    def __init__(self, group, Nu, f_fun, crl, cru, xdes, udes, K=10, tf=1.0, Kmesh=4, Q=None, R=None, Qtf=None):
        self.group, self.Nx, self.Nu = group, group.dof, Nu
        self.f_fun, self.crl, self.cru = f_fun, np.asarray(crl, float), np.asarray(cru, float)
        self.xdes, self.udes, self.tf = xdes, udes, float(tf)
        self.mesh = Mesh((K + Kmesh - 1) // Kmesh, Kmesh)                       # mpc.hpp:408
        self.Ncr = len(self.crl)
        d = ocp_dims(self.mesh, self.Nx, Nu, self.Ncr, self.Nx)
        self.dims = d
        self.qp = TripletQP(d["Nvar"], d["Ncon"])
        Q = np.eye(self.Nx) if Q is None else Q
        R = np.eye(Nu) if R is None else R
        Qtf = np.eye(self.Nx) if Qtf is None else Qtf
        mpc_update_cost(self.qp, self.mesh, self.tf, Q, R, Qtf, self.Nx, Nu)    # ctor only (mpc.hpp:423)
        self.warm = None

    def _ce_fun(self, x0_fix):
        g = self.group

        def ce(tf, x0, xf):
            e = g.rminus(x0, x0_fix)                                            # mpc.hpp:284
            J = g.dr_expinv(e)                                                  # mpc.hpp:290-296
            mask = np.zeros((g.dof, g.dof), bool)
            if isinstance(g, BundleSE2Rk):                                      # d_exp_sparse_pattern<X>
                mask[:2, :3] = True; mask[2, 2] = True
                for k in range(3, g.dof):
                    mask[k, k] = True
            else:
                mask[:] = np.eye(g.dof, dtype=bool) if g.commutative else True
            return e, J, None, mask, None
        return ce

    def transcribe(self, t, x):
        """Updates the QP like MPC::operator() does (mpc.hpp:473-488) and returns the TripletQP."""
        g, Nx, Nu, tf = self.group, self.Nx, self.Nu, self.tf
        xl = lambda tr: self.xdes(t + tr)
        ul = lambda tr: self.udes(t + tr)
        ocp_to_qp_update_dyn(self.qp, g, self.mesh, tf, xl, ul, lambda tr, xx, uu: self.f_fun(t + tr, xx, uu), Nx, Nu)
        if not hasattr(self, "_cr_done"):  # time-invariant cr: transcribed by the ctor's ocp_to_qp_update only (:482-485)
            cr = lambda tr, xx, uu: (uu, np.zeros((self.Ncr, Nx)), np.eye(self.Ncr, Nu))
            ocp_to_qp_update_cr(self.qp, self.mesh, tf, xl, ul, cr, self.crl, self.cru, Nx, Nu)
            self._cr_done = True
        ocp_to_qp_update_ce(self.qp, self.mesh, tf, xl, self._ce_fun(np.asarray(x, float)), np.zeros(Nx), np.zeros(Nx), Nx, Nu, self.Ncr)
        return self.qp

    def input_from_primal(self, t, primal):
        return self.udes(t) + primal[self.dims["xvar_L"]: self.dims["xvar_L"] + self.Nu]   # mpc.hpp:518

    def trajectories(self, t, primal):
        """The optional outputs of MPC::operator() (mpc.hpp:493-507): u_traj[i] = udes(t + tf tau_i) + primal.segment<Nu>(uvar_B
        + i Nu), i < N; x_traj[i] = xdes(t + tf tau_i) (+) primal.segment<Nx>(i Nx), i <= N; tau = Mesh::all_nodes()."""
        tau = self.mesh.all_nodes()
        N, Nx, Nu, uB = len(tau) - 1, self.Nx, self.Nu, self.dims["xvar_L"]
        u_traj = np.stack([self.udes(t + self.tf * tau[i]) + primal[uB + i * Nu: uB + (i + 1) * Nu] for i in range(N)])
        x_traj = np.stack([self.group.rplus(self.xdes(t + self.tf * tau[i])[0], primal[i * Nx: (i + 1) * Nx]) for i in range(N + 1)])
        return u_traj, x_traj

    @staticmethod
    def keeps_warmstart(code: int) -> bool:
        return code in (0, 5, 4)  # Optimal, MaxTime, MaxIterations  (mpc.hpp:510-516)


# vehicle MPC of examples/mpc_asif_vehicle.cpp:42-89 -----------------------------------------------------------------
VEHICLE_VDES = np.array([1.0, 0.0, 0.4])


def vehicle_xdes(t):
    """xdes(t) = X{SE2(SO2(pi/2), (2.5, 0)) + t vdes, vdes} and its body velocity (mpc_asif_vehicle.cpp:72-78)."""
    g0 = np.array([2.5, 0.0, np.sin(np.pi / 2), np.cos(np.pi / 2)])
    g = se2_compose(g0, se2_exp(t * VEHICLE_VDES))
    return np.concatenate([g, VEHICLE_VDES]), np.concatenate([VEHICLE_VDES, np.zeros(3)])


def vehicle_mpc(K=50, tf=5.0, model: VehicleModel | None = None):
    model = model or VehicleModel()
    f = lambda t, x, u: (model.f(x, u), model.df_dx(), model.f_u(x, u)[1])
    return MPCRestated(model.group, 2, f, [-0.5, -0.5], [0.5, 0.5], vehicle_xdes, lambda t: np.zeros(2), K=K, tf=tf)


def sample_vehicle_states(B: int, seed: int = 5, sigma: float = 0.1):
    """SURVEY 8(d) cfg3/cfg5 sampling: t0 ~ U(0, 30), x0 = xdes(t0) (+) xi, xi ~ N(0, sigma^2 I_6)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    g = BundleSE2Rk(3)
    t0 = rng.uniform(0.0, 30.0, B)
    xi = sigma * rng.normal(size=(B, 6))
    x0 = np.stack([g.rplus(vehicle_xdes(t0[b])[0], xi[b]) for b in range(B)])
    return t0, x0
