"""GPU parity tests of the on-device MPC (sfb_mpc_fleet_*) against the numpy restatement of the reference's transcription
(oracle/transcribe.py: ocp_to_qp.hpp:40-400, mpc.hpp:405-519) followed by the CPU QP oracle, for the vehicle family of
examples/mpc_asif_vehicle.cpp -- BASELINE.json configs[2]: K = 50 -> n = m = 422, sparse.

Bar: pattern identical to the restatement's, values within 1e-12; status / iteration counts exact on well-posed instances;
the applied input u within 1e-6 relative in fp64 (1e-3 in fp32); warm-start retention rule of mpc.hpp:510-516 step by step.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sfb():
    import smooth_feedback_b200 as s

    return s


def _dense(pat, Pv, Av):
    from smooth_feedback_b200.generators import sparse_to_dense

    return sparse_to_dense(pat, Pv, Av)


def test_transcription_matches_restatement(sfb):
    from workloads import vehicle_mpc_batch

    B = 48
    pat, Pv, q, Av, l, u, mpc, t0, x0 = vehicle_mpc_batch(B, seed=3)
    fleet = sfb.MPCVehicleFleet(B)
    assert (fleet.n, fleet.m, fleet.nnzP, fleet.nnzA) == (422, 422, len(pat["P_rowidx"]), len(pat["A_colidx"]))
    fp = fleet.pattern()
    for k in ("P_colptr", "P_rowidx", "A_rowptr", "A_colidx"):
        assert np.array_equal(fp[k], pat[k]), k
    P2, q2, A2, l2, u2 = fleet.to_qp(t0, x0)
    assert np.abs(P2 - Pv).max() <= 1e-14 and not q2.any() and not q.any()
    assert np.abs(A2 - Av).max() <= 1e-12 * np.abs(Av).max()
    assert np.abs(l2 - l).max() <= 1e-12 and np.abs(u2 - u).max() <= 1e-12
    # other mesh sizes / weights: K = 10 (tests/test_mpc.cpp's mesh: 3 intervals) and zero weights that shrink the pattern
    from oracle import transcribe as tr

    prm = sfb.MPCVehicleParams(K=10, tf=2.0, Q=(1, 2, 0, 1, 1, 3), R=(0.5, 2.0), Qtf=(1, 0, 1, 2, 1, 1))
    small = sfb.MPCVehicleFleet(4, prm)
    model = tr.VehicleModel()
    f = lambda t, x, uu: (model.f(x, uu), model.df_dx(), model.f_u(x, uu)[1])
    ref = tr.MPCRestated(model.group, 2, f, [-0.5, -0.5], [0.5, 0.5], tr.vehicle_xdes, lambda t: np.zeros(2), K=10, tf=2.0,
                         Q=np.diag(prm.Q), R=np.diag(prm.R), Qtf=np.diag(prm.Qtf))
    ts, xs = tr.sample_vehicle_states(4, seed=8)
    P3, q3, A3, l3, u3 = small.to_qp(ts, xs)
    sp = small.pattern()
    for b in range(4):
        qp = ref.transcribe(ts[b], xs[b])
        rp, ci, av = qp.csr_A(); cp, ri, pv = qp.csc_P()
        assert np.array_equal(sp["A_rowptr"], rp) and np.array_equal(sp["A_colidx"], ci)
        assert np.array_equal(sp["P_colptr"], cp) and np.array_equal(sp["P_rowidx"], ri)
        assert np.abs(P3[b] - pv).max() <= 1e-14 and np.abs(A3[b] - av).max() <= 1e-12 * np.abs(av).max()
        assert np.abs(l3[b] - qp.l).max() <= 1e-12 and np.abs(u3[b] - qp.u).max() <= 1e-12


def _oracle_step(oracle, mpc, pat, t, x, warm, max_iter=4000):
    """MPC::operator() restated for a batch: transcribe + dense oracle solve (cold or warm per agent)."""
    B = len(t)
    Pv, Av, ls, us = [], [], [], []
    for b in range(B):
        qp = mpc.transcribe(t[b], x[b])
        _, _, av = qp.csr_A(); _, _, pv = qp.csc_P()
        Pv.append(pv); Av.append(av); ls.append(qp.l.copy()); us.append(qp.u.copy())
    Pv, Av, ls, us = (np.stack(v) for v in (Pv, Av, ls, us))
    P, A = _dense(pat, Pv, Av)
    q = np.zeros((B, pat["n"]))
    prm = oracle.default_params(max_iter=max_iter)
    kw = {} if warm is None else dict(warm_x=warm[0], warm_y=warm[1])
    o = oracle.qp_solve_batch(P, q, A, ls, us, params=prm, nthreads=8, **kw)
    o2 = oracle.qp_solve_batch(P, q, A, ls, us, params=prm, nthreads=8, fast=True, **kw)
    return o, o2


def test_closed_loop_parity_with_resident_warm_starts(sfb, oracle):
    """Several control steps of a small fleet: the device keeps each agent's warm start (retained if Optimal / MaxTime /
    MaxIterations, mpc.hpp:510-516); the host loop does the same with the oracle.  States advance with the ORACLE's input so
    that both sides see identical problems at every step."""
    from oracle import transcribe as tr
    from workloads import vehicle_mpc_batch

    B, steps, dt = 12, 4, 0.025
    pat, _, _, _, _, _, mpc, t0, x0 = vehicle_mpc_batch(B, seed=4)
    fleet = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)))
    g, model = tr.BundleSE2Rk(3), tr.VehicleModel()
    t, x = t0.copy(), x0.copy()
    warm = None
    xvar = mpc.dims["xvar_L"]
    for k in range(steps):
        u, st, it = fleet(t, x)
        o, o2 = _oracle_step(oracle, mpc, pat, t, x, warm)
        wp = (o.status == o2.status) & (o.iter == o2.iter) & (o.active == o2.active).all(axis=1)
        assert wp.mean() >= 0.75
        assert np.array_equal(st[wp], o.status[wp]) and np.array_equal(it[wp], o.iter[wp]), (k, st, o.status, it, o.iter)
        uo = o.x[:, xvar:xvar + 2]
        assert np.abs(u - uo)[wp].max() <= 1e-6 * max(1.0, np.abs(uo).max())
        if k > 0:  # the warm start is used (same iterates as the oracle's warm solve), whatever it buys on a moved problem
            assert np.array_equal(it[wp], o.iter[wp])
        keep = np.isin(o.status, [0, 4, 5])
        if warm is None:
            warm = (np.zeros_like(o.x), np.zeros_like(o.y))
        warm[0][keep] = o.x[keep]; warm[1][keep] = o.y[keep]
        x = np.stack([g.rplus(x[b], dt * model.f(x[b], uo[b])) for b in range(B)])
        t = t + dt
    # reset_warmstart: the next step is a cold solve again
    fleet.reset_warmstart()
    u, st, it = fleet(t0, x0)
    o, o2 = _oracle_step(oracle, mpc, pat, t0, x0, None)
    wp = (o.status == o2.status) & (o.iter == o2.iter)
    assert np.array_equal(it[wp], o.iter[wp]) and (it > 2).all()


def test_step_fp32_and_device_path(sfb, oracle):
    import torch

    from workloads import vehicle_mpc_batch

    B = 64
    pat, Pv, q, Av, l, u, mpc, t0, x0 = vehicle_mpc_batch(B, seed=6)
    f32 = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)), dtype=np.float32)
    u32, st32, it32 = f32(t0.astype(np.float32), x0.astype(np.float32))
    t32 = t0.astype(np.float32).astype(np.float64); x32 = x0.astype(np.float32).astype(np.float64)
    o, o2 = _oracle_step(oracle, mpc, pat, t32, x32, None)
    xvar = mpc.dims["xvar_L"]
    uo = o.x[:, xvar:xvar + 2]
    assert np.array_equal(st32, o.status)
    same = it32 == o.iter
    assert same.mean() >= 0.9
    assert np.abs(u32 - uo)[same].max() <= 1e-3 * max(1.0, np.abs(uo).max())
    # device tensors in, device tensors out; equal to the host path bit for bit
    f64 = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)))
    uh, sth, ith = f64(t0, x0)
    f64.reset_warmstart()
    td = torch.from_numpy(t0).cuda(); xd = torch.from_numpy(x0).cuda()
    ud, std, itd, px, py = f64(td, xd, return_solution=True)
    torch.cuda.synchronize()
    assert np.array_equal(ud.cpu().numpy(), uh) and np.array_equal(std.cpu().numpy(), sth)
    assert np.array_equal(px.cpu().numpy()[:, xvar:xvar + 2], uh)


def test_trajectory_outputs(sfb):
    """The optional outputs of MPC::operator() (mpc.hpp:493-507): u_traj / x_traj of the fleet against the restatement applied
    to the SAME primal solution (1e-12: pure post-processing), the mesh nodes against Mesh::all_nodes(), the properties the
    reference's construction guarantees (u_traj[0] is the applied input; x_traj[0] is the measured state: the end constraint
    pins x_0), host path == device path, fp32 within 1e-5."""
    import torch

    from workloads import vehicle_mpc_batch

    B = 24
    pat, _, _, _, _, _, mpc, t0, x0 = vehicle_mpc_batch(B, seed=8)
    fleet = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)))
    N, tau = fleet.nodes()
    ref_tau = mpc.mesh.all_nodes()
    assert N == len(ref_tau) - 1 == 52 and np.abs(tau - ref_tau).max() <= 1e-15
    u, st, it, px, py = fleet(t0, x0, return_solution=True)
    ut, xt = fleet.trajectories(t0)
    assert ut.shape == (B, N, 2) and xt.shape == (B, N + 1, 7)
    for b in range(B):
        uo, xo = mpc.trajectories(t0[b], px[b])
        assert np.abs(ut[b] - uo).max() <= 1e-12 * max(1.0, np.abs(uo).max())
        assert np.abs(xt[b] - xo).max() <= 1e-12 * max(1.0, np.abs(xo).max())
    assert np.array_equal(ut[:, 0, :], u)
    ok = st == 0
    assert ok.mean() >= 0.9
    assert np.abs(xt[ok, 0, :] - x0[ok]).max() <= 1e-4           # ce = x_0 (-) x0_fix = 0 up to the solver's tolerance
    assert np.abs(np.hypot(xt[..., 2], xt[..., 3]) - 1).max() <= 1e-12   # SE(2) elements stay on the group
    # device tensors: same bits
    td = torch.from_numpy(t0).cuda()
    utd, xtd = fleet.trajectories(td)
    torch.cuda.synchronize()
    assert np.array_equal(utd.cpu().numpy(), ut) and np.array_equal(xtd.cpu().numpy(), xt)
    # fp32 fleet
    f32 = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)), dtype=np.float32)
    t32 = t0.astype(np.float32)
    u32, _, _, p32, _ = f32(t32, x0.astype(np.float32), return_solution=True)
    ut32, xt32 = f32.trajectories(t32)
    for b in range(0, B, 5):
        uo, xo = mpc.trajectories(float(t32[b]), p32[b].astype(np.float64))
        assert np.abs(ut32[b] - uo).max() <= 1e-5 * max(1.0, np.abs(uo).max())
        assert np.abs(xt32[b] - xo).max() <= 1e-5 * max(1.0, np.abs(xo).max())


def test_full_size_closed_loop_properties(sfb):
    """BASELINE configs[2] at full size (8192 agents, fp32) for 5 control steps on device tensors: every solve Optimal, warm
    steps exit at the first check, inputs inside their box, replicas of 256 distinct agents identical."""
    import torch

    from oracle import transcribe as tr

    base, B = 256, 8192
    t0, x0 = tr.sample_vehicle_states(base, seed=12)
    rep = B // base
    t = torch.from_numpy(np.tile(t0, rep)).to("cuda:0", dtype=torch.float32).contiguous()
    x = torch.from_numpy(np.tile(x0, (rep, 1))).to("cuda:0", dtype=torch.float32).contiguous()
    fleet = sfb.MPCVehicleFleet(B, sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000)), dtype=np.float32)
    for k in range(5):
        u, st, it = fleet(t, x)
        torch.cuda.synchronize()
        un, stn, itn = u.cpu().numpy(), st.cpu().numpy(), it.cpu().numpy()
        assert (stn == 0).all()
        assert np.array_equal(un.reshape(rep, base, 2), np.broadcast_to(un[:base], (rep, base, 2)))
        assert (np.abs(un) <= 0.5 + 1e-4).all()
        if k > 0:
            assert (itn == 2).mean() > 0.9  # the same problem again: the resident warm start is its solution
