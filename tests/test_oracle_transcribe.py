"""CPU tests of oracle/transcribe.py: the numpy restatement of asif_to_qp (asif_func.hpp:104-199) and of the MPC
transcription (ocp_to_qp.hpp:40-400, mpc.hpp:405-519), pinned by the reference's own tests for these functions:

  tests/test_asif.cpp:37-95     Asif.Basic       (QP shape and block structure; SE(2), K = 3, nh = 2)
  tests/test_ocp_to_qp.cpp:41-107  OcpToQp.Basic (a known trajectory satisfies l <= A var <= u to 1e-8)
  tests/test_mpc.cpp:73-118     Mpc.Api          (Optimal; u(no warm start) == u(warm start))

plus derivative checks of the restated smooth group operations (their source is not in /root/reference).
"""
import numpy as np
import pytest

from oracle import transcribe as tr


def test_se2_group_ops():
    rng = np.random.default_rng(0)
    g = tr.BundleSE2Rk(3)
    for _ in range(20):
        a = rng.normal(size=6)
        x = g.rplus(g.identity(), rng.normal(size=6))
        y = g.rplus(x, a)
        assert np.allclose(g.rminus(y, x), a, atol=1e-12)
        assert abs(y[2] ** 2 + y[3] ** 2 - 1) < 1e-12
    # small angles go through the series branches
    a = np.array([0.3, -0.2, 1e-11, 0.1, 0.2, 0.3])
    assert np.allclose(g.rminus(g.rplus(g.identity(), a), g.identity()), a, atol=1e-14)


def test_dr_expinv_is_the_right_jacobian_of_rminus():
    # MPCCE::jacobian (mpc.hpp:287-296): d^r/dx0 [x0 (-) x0_fix] = dr_expinv(x0 (-) x0_fix)
    rng = np.random.default_rng(1)
    g = tr.BundleSE2Rk(3)
    for _ in range(5):
        x0 = g.rplus(g.identity(), rng.normal(size=6))
        xf = g.rplus(x0, 0.5 * rng.normal(size=6))
        e = g.rminus(x0, xf)
        J = g.dr_expinv(e)
        Jn = np.zeros((6, 6))
        h = 1e-4
        for k in range(6):
            d = np.zeros(6); d[k] = h
            Jn[:, k] = (g.rminus(g.rplus(x0, d), xf) - g.rminus(g.rplus(x0, -d), xf)) / (2 * h)
        assert np.allclose(J, Jn, atol=1e-8)
    assert np.allclose(tr.se2_dr_expinv(np.array([0.3, 0.1, 1e-7])), tr.se2_dr_expinv(np.array([0.3, 0.1, 2e-5])), atol=1e-5)


def test_lgr_mesh():
    for K in (3, 4, 5, 8):
        x, w = tr.lgr_nodes(K)
        assert x[0] == -1.0 and np.all(np.diff(x) > 0) and x[-1] < 1.0
        assert abs(w.sum() - 2.0) < 1e-13
        for p in range(2 * K - 1):  # Radau quadrature is exact to degree 2K - 2
            assert abs(w @ x ** p - (1 - (-1) ** (p + 1)) / (p + 1)) < 1e-12
    mesh = tr.Mesh(13, 4)
    assert mesh.N_colloc() == 52 and mesh.N_ivals() == 13
    nodes, wts = mesh.all_nodes(), mesh.all_weights()
    assert len(nodes) == 53 and nodes[0] == 0.0 and nodes[-1] == 1.0 and abs(wts.sum() - 1.0) < 1e-13 and wts[-1] == 0.0
    al, D = mesh.interval_diffmat_unscaled(0)
    assert abs(al - 26.0) < 1e-12 and D.shape == (5, 4)
    ext = np.concatenate([tr.lgr_nodes(4)[0], [1.0]])
    for p in range(5):  # differentiates polynomials of degree <= K exactly at the collocation nodes
        assert np.allclose((ext ** p) @ D, p * ext[:4] ** max(p - 1, 0) if p else 0.0, atol=1e-12)


def test_asif_basic_structure():
    # tests/test_asif.cpp:37-95
    K, Nu, Nh = 3, 2, 2
    g = tr.SE2()
    rng = np.random.default_rng(3)
    x0 = g.rplus(g.identity(), rng.normal(size=3))
    f_u = lambda x, u: (np.array([u[0], 0.0, u[1]]), np.array([[1.0, 0], [0, 0], [0, 1.0]]))
    f_cl = lambda t, x: (np.array([-0.1, 0.0, 1.0]), np.zeros((3, 3)))

    def h(t, x):
        s, c = x[2], x[3]
        return x[:2].copy(), np.zeros(2), np.array([[c, -s, 0.0], [s, c, 0.0]])  # d r2 / d a = R

    pbm = tr.ASIFProblem(T=1.0, x0=x0, u_des=np.array([0.5, 0.5]), W_u=np.ones(2), ulim_A=np.eye(2), ulim_c=np.zeros(2),
                         ulim_l=-np.ones(2), ulim_u=np.ones(2))
    P, q, A, l, u = tr.asif_to_qp(g, pbm, tr.ASIFtoQPParams(K=K), f_u, f_cl, h, Nh)
    niq = 2
    assert P.shape == (Nu + 1, Nu + 1) and q.shape == (Nu + 1,)
    assert A.shape == (Nh * K + niq + 1, Nu + 1) and l.shape == u.shape == (A.shape[0],)
    assert np.allclose(A[:Nh * K, Nu], 1.0)
    assert np.allclose(A[Nh * K:Nh * K + niq, :Nu], pbm.ulim_A)
    assert np.allclose(A[Nh * K + niq], [0, 0, 1])
    assert u[:Nh * K].min() == np.inf
    assert np.allclose(l[Nh * K:Nh * K + niq], pbm.ulim_l - pbm.ulim_A @ pbm.u_des)
    assert np.allclose(u[Nh * K:Nh * K + niq], pbm.ulim_u - pbm.ulim_A @ pbm.u_des)
    assert l[Nh * K + niq] == 0 and u[Nh * K + niq] == np.inf
    # first barrier rows: sensitivity is the identity at t = 0
    f0, B = f_u(x0, pbm.u_des)
    hv, _, dh = h(0.0, x0)
    assert np.allclose(A[:Nh, :Nu], dh @ B) and np.allclose(l[:Nh], -1.0 * hv - dh @ f0)


def test_asif_step_schedule_vehicle():
    # asif_func.hpp:139-143,170-176 with T = 2.5, K = 200, dt = 0.01 (mpc_asif_vehicle.cpp:105-123)
    tk, steps = tr.asif_step_schedule(2.5, 200, 0.01)
    assert len(tk) == 200 and tk[0] == 0.0
    assert steps[0] == [0.01, 0.01]  # dt_act is fixed before the while loop: the first interval overshoots tau = 0.0125
    n = sum(len(s) for s in steps)
    assert 200 <= n <= 400
    assert abs(sum(sum(s) for s in steps) - 2.5) < 0.011  # the trajectory ends within one step of T


def test_vehicle_asif_qp_and_solve(oracle):
    # BASELINE configs[4]: n = 3, m = 203, polish off (mpc_asif_vehicle.cpp:127)
    t0, x0 = tr.sample_vehicle_states(24, seed=7)
    rng = np.random.default_rng(7)
    u_des = rng.uniform(-0.5, 0.5, (24, 2))
    P, q, A, l, u = tr.vehicle_asif_qp_batch(x0, u_des)
    assert P.shape == (24, 3, 3) and A.shape == (24, 203, 3)
    assert np.allclose(P[0], np.diag([20.0, 1.0, 100.0])) and not q.any()
    assert np.isinf(u[:, :200]).all() and np.isinf(u[:, 202]).all() and (A[:, :200, 2] == 1).all()
    o = oracle.qp_solve_batch(P, q, A, l, u, params=oracle.default_params(max_iter=4000, polish=0), nthreads=4)
    assert (o.status == 0).all(), o.status
    # the filtered input respects the input bounds up to the solver tolerance (rows 200, 201)
    uf = u_des + o.x[:, :2]
    assert (uf[:, 0] <= 0.5 + 5e-3).all() and (uf[:, 0] >= -0.2 - 5e-3).all()
    assert (np.abs(uf[:, 1]) <= 0.5 + 5e-3).all()
    # far from the obstacle nothing is filtered
    far = np.array([[10.0, 10.0, 0.0, 1.0, 1.0, 0.0, 0.0]])
    Pf, qf, Af, lf, uf_ = tr.vehicle_asif_qp_batch(far, np.array([[0.1, 0.1]]))
    of = oracle.qp_solve_batch(Pf, qf, Af, lf, uf_, params=oracle.default_params(max_iter=4000, polish=0))
    assert of.status[0] == 0 and np.abs(of.x[0]).max() < 1e-3


def test_ocp_to_qp_basic_known_trajectory():
    # tests/test_ocp_to_qp.cpp:41-107: double integrator on R^2, Mesh<5,5> refined with refine_ph(0, 10), tf = 2
    g = tr.Rn(2)
    mesh = tr.Mesh(1, 5)
    mesh.refine_ph(0, 10)
    assert mesh.N_ivals() == 2 and mesh.N_colloc() == 10
    tf = 2.0
    xl = lambda t: (np.array([0.05 * t * t, 0.1 * t]), np.array([0.1 * t, 0.1]))
    ul = lambda t: np.array([0.1])
    f = lambda t, x, u: (np.array([x[1], u[0]]), np.array([[0.0, 1.0], [0.0, 0.0]]), np.array([[0.0], [1.0]]))
    cr = lambda t, x, u: (np.array([u[0]]), np.zeros((1, 2)), np.ones((1, 1)))
    ce = lambda tf_, x0, xf: (xf, None, np.eye(2), None, np.ones((2, 2), bool))
    d = tr.ocp_dims(mesh, 2, 1, 1, 2)
    qp = tr.TripletQP(d["Nvar"], d["Ncon"])
    tr.ocp_to_qp_update_dyn(qp, g, mesh, tf, xl, ul, f, 2, 1)
    tr.ocp_to_qp_update_cr(qp, mesh, tf, xl, ul, cr, np.array([-1.0]), np.array([1.0]), 2, 1)
    tr.ocp_to_qp_update_ce(qp, mesh, tf, xl, ce, np.array([-5.0, -5.0]), np.array([5.0, 5.0]), 2, 1, 1)
    _, A = qp.dense()
    x0, v0, u0 = 3.0, -0.3, 0.1
    N = mesh.N_colloc()
    nodes = mesh.all_nodes()
    X = np.array([[x0 + v0 * tf * t + u0 * (tf * t) ** 2 / 2, v0 + u0 * tf * t] for t in nodes])
    var = np.concatenate([X.reshape(-1), np.full(N, u0)])
    assert (A @ var - qp.l).min() >= -1e-8 and (qp.u - A @ var).min() >= -1e-8
    assert np.abs((A @ var)[: 2 * N]).max() < 1e-8  # collocation rows hold with equality


def _se2_test_mpc(xrand):
    # the model of tests/test_mpc.cpp:12-44: SE(2), f = (u0, 0, u1), cr = u in [-1, 1], K = 10 (default) -> n = m = 63
    g = tr.SE2()
    f = lambda t, x, u: (np.array([u[0], 0.0, u[1]]), np.zeros((3, 3)), np.array([[1.0, 0], [0, 0], [0, 1.0]]))
    xdes = lambda t: (g.identity(), np.zeros(3))
    udes = lambda t: np.ones(2)
    return tr.MPCRestated(g, 2, f, [-1, -1], [1, 1], xdes, udes, K=10, tf=1.0)


def _solve_triplet(oracle, qp, warm=None):
    P, A = qp.dense()
    kw = {} if warm is None else dict(warm_x=warm[0][None], warm_y=warm[1][None])
    return oracle.qp_solve_batch(P[None], qp.q[None], A[None], qp.l[None], qp.u[None],
                                 params=oracle.default_params(max_iter=20000), **kw)


def test_mpc_api_se2(oracle):
    # tests/test_mpc.cpp:73-118: Optimal, and u(cold) == u(warm) (isApprox, 1e-12 relative in Eigen)
    g = tr.SE2()
    x = g.rplus(g.identity(), np.random.default_rng(5).normal(size=3))
    mpc = _se2_test_mpc(x)
    assert mpc.dims["Nvar"] == 63 and mpc.dims["Ncon"] == 63
    qp = mpc.transcribe(2.0, x)
    o1 = _solve_triplet(oracle, qp)
    assert o1.status[0] == 0
    u1 = mpc.input_from_primal(2.0, o1.x[0])
    qp = mpc.transcribe(3.0, x)
    o2 = _solve_triplet(oracle, qp, warm=(o1.x[0], o1.y[0]))
    assert o2.status[0] == 0 and o2.iter[0] == 2
    u2 = mpc.input_from_primal(3.0, o2.x[0])
    assert np.allclose(u1, u2, rtol=1e-9, atol=1e-12)
    # P holds the upper triangle only and is diagonal for identity weights; x_N carries no cost (see module docstring)
    P, _ = qp.dense()
    assert np.array_equal(P, np.triu(P)) and np.count_nonzero(P - np.diag(np.diag(P))) == 0
    assert not P[3 * 12: 3 * 13, 3 * 12: 3 * 13].any()


def test_vehicle_mpc_cfg3(oracle):
    # BASELINE configs[2]: SE(2) x R^3 bus, K = 50 -> 13 intervals x 4 nodes -> n = m = 422
    mpc = tr.vehicle_mpc()
    assert mpc.dims["Nvar"] == 422 and mpc.dims["Ncon"] == 422
    t0, x0 = tr.sample_vehicle_states(3, seed=11)
    us = []
    for b in range(3):
        qp = mpc.transcribe(t0[b], x0[b])
        o = _solve_triplet(oracle, qp)
        assert o.status[0] == 0
        us.append(mpc.input_from_primal(t0[b], o.x[0]))
        _, A = qp.dense()
        # equality rows (dynamics, initial state) hold at the polished solution; inputs inside their box
        r = A @ o.x[0]
        eq = np.isclose(qp.l, qp.u)
        assert np.abs(r - qp.l)[eq].max() < 1e-7
        assert (r <= qp.u + 1e-6).all() and (r >= qp.l - 1e-6).all()
    rp, ci, _ = qp.csr_A()
    cp, ri, pv = qp.csc_P()
    assert rp[-1] == 312 * 12 + 104 * 8 + 10 and cp[-1] == 6 * 52 + 2 * 52  # nnz(A), nnz(P)
    assert np.abs(np.array(us)).max() <= 0.5 + 1e-9


def test_mpc_trajectory_outputs(oracle):
    # mpc.hpp:493-507: u_traj[i] = udes(t + tf tau_i) + u-segment i, x_traj[i] = xdes(t + tf tau_i) (+) x-segment i.  The
    # reference's tests never read them; pinned here by what the construction guarantees: u_traj[0] is the applied input
    # (:518), x_traj[0] is the measured state (end constraint ce = x_0 (-) x0_fix = 0), the trajectory satisfies the input
    # box, and a zero primal returns the desired trajectory itself.
    mpc = tr.vehicle_mpc()
    t0, x0 = tr.sample_vehicle_states(2, seed=3)
    tau = mpc.mesh.all_nodes()
    assert len(tau) == 53 and tau[0] == 0.0 and tau[-1] == 1.0 and (np.diff(tau) > 0).all()
    for b in range(2):
        o = _solve_triplet(oracle, mpc.transcribe(t0[b], x0[b]))
        assert o.status[0] == 0
        ut, xt = mpc.trajectories(t0[b], o.x[0])
        assert ut.shape == (52, 2) and xt.shape == (53, 7)
        assert np.array_equal(ut[0], mpc.input_from_primal(t0[b], o.x[0]))
        assert np.abs(xt[0] - x0[b]).max() < 1e-6
        assert np.abs(ut).max() <= 0.5 + 1e-6
        assert np.abs(np.hypot(xt[:, 2], xt[:, 3]) - 1).max() < 1e-12
    ut, xt = mpc.trajectories(1.5, np.zeros(mpc.dims["Nvar"]))
    assert not ut.any()
    for i in (0, 17, 52):
        assert np.allclose(xt[i], tr.vehicle_xdes(1.5 + mpc.tf * tau[i])[0], rtol=0, atol=1e-15)

