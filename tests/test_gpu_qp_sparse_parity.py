"""GPU parity tests for the sparse (shared-pattern) QP path: sfb_qp_solve_sparse_batch_* vs the CPU oracle.

The oracle for QuadraticProgramSparse problems is the dense restatement applied to the densified problem: the reference's
sparse branches run the same algorithm on the same numbers (scale / check_stopping walk the stored entries, only col >= row
of P enters the KKT matrix); they differ from its dense branch only in the factorisation routine (SimplicialLDLT instead of
the pivoted dense LDLT), i.e. in rounding -- "parity unpinned" for both.
"""
import numpy as np
import pytest

from test_gpu_qp_parity import REL_F32, REL_F64, _assert_fp32, _assert_parity as _assert_parity_dense, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sfb():
    import smooth_feedback_b200 as s

    return s


_HANDLES = {}
_KERNEL = "onchip"


def _handle(sfb, kernel):
    """One handle per sparse kernel: the on-chip kernel (qp_sparse_cta.cuh, default whenever the instance fits in shared memory)
    and the HBM-tiled kernel (qp_sparse_tiled.cuh, the fallback), selected when the handle is created."""
    import os

    if kernel not in _HANDLES:
        os.environ["SFB_SPARSE_KERNEL"] = "tiled" if kernel == "tiled" else "cta"
        try:
            _HANDLES[kernel] = sfb.Handle(0)
        finally:
            os.environ.pop("SFB_SPARSE_KERNEL", None)
    return _HANDLES[kernel]


@pytest.fixture(autouse=True, params=["onchip", "tiled"])
def sparse_kernel(request):
    """Every test of this module runs against both sparse kernels."""
    global _KERNEL
    _KERNEL = request.param
    yield request.param
    _KERNEL = "onchip"


def _pattern(sfb, pat, **kw):
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL), **kw)
    return sp


def _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, prm_kw=None, max_iter=4000, warm=None, dtype=np.float64):
    from smooth_feedback_b200.generators import sparse_to_dense

    prm_kw = dict(prm_kw or {})
    sp = _pattern(sfb, pat)
    prm = sfb.QPSolverParams(max_iter=max_iter, **prm_kw)
    c = lambda t: None if t is None else np.ascontiguousarray(t, dtype=dtype)
    wx, wy = (None, None) if warm is None else warm
    r = sfb.solve_sparse_batch(sp, c(Pv), c(q), c(Av), c(l), c(u), prm, c(wx), c(wy))
    P, A = sparse_to_dense(pat, Pv, Av)
    okw = {k: (int(v) if isinstance(v, bool) else v) for k, v in prm_kw.items()}
    op = oracle.default_params(max_iter=max_iter, **okw)
    kw = {} if warm is None else dict(warm_x=wx, warm_y=wy)
    o = oracle.qp_solve_batch(P, q, A, l, u, params=op, nthreads=8, **kw)
    o2 = oracle.qp_solve_batch(P, q, A, l, u, params=op, nthreads=8, fast=True, **kw)
    wp = (o.status == o2.status) & (o.iter == o2.iter) & (o.active == o2.active).all(axis=1)
    _solve_both.last_fast = o2
    _solve_both.last_problem = (P, q, A, l, u)
    return sp, r, o, wp


def _assert_parity(r, o, wp, rel, min_well_posed=0.97):
    # the dense helper, fed with THIS module's second oracle build and densified problem (instances outside the
    # well-posed mask must match one of the two builds or be finite KKT points -- never unchecked)
    _assert_parity_dense(r, o, wp, rel, min_well_posed, o2=_solve_both.last_fast, problem=_solve_both.last_problem)


def _assert_discrete_and_conditioned(r, o, o_fast, wp, min_well_posed=0.9):
    """Exact status / iteration / active-set parity on the well-posed instances; continuous outputs within the north-star
    1e-6 OR within 10x the oracle's own FMA-vs-no-FMA disagreement on that instance (n = 12 random sparse instances
    include polish systems with nearly dependent active rows, where the duals are determined to ~1e-5 only)."""
    assert wp.mean() >= min_well_posed
    assert np.array_equal(r.status[wp], o.status[wp]) and np.array_equal(r.iter[wp], o.iter[wp])
    assert np.array_equal(r.active[wp], o.active[wp])
    ok = wp & (o.status == 0)
    tol_x = np.maximum(REL_F64, 10 * rel_err(o_fast.x[ok], o.x[ok]))
    tol_y = np.maximum(REL_F64, 10 * rel_err(o_fast.y[ok], o.y[ok]))
    assert (rel_err(r.x[ok], o.x[ok]) <= tol_x).all() and (rel_err(r.y[ok], o.y[ok]) <= tol_y).all()
    assert (rel_err(r.x[ok], o.x[ok]) <= REL_F64).mean() >= 0.95 and (rel_err(r.y[ok], o.y[ok]) <= REL_F64).mean() >= 0.95


@pytest.mark.parametrize("n,m,density", [(10, 20, 0.3), (40, 60, 0.15), (60, 30, 0.1)])
def test_parity_random_sparse(sfb, oracle, n, m, density):
    # benchmarks/bench_types.hpp recipe at density < 1 (the reference's sparse benchmark arm), shared mask
    from smooth_feedback_b200.generators import random_sparse_qp_numpy

    pat, Pv, q, Av, l, u = random_sparse_qp_numpy(200, n, m, density=density, seed=n + m)
    sp, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u)
    assert sp.nnzL <= n * (n - 1) // 2
    if m >= n:
        _assert_parity(r, o, wp, REL_F64, min_well_posed=0.95)
    else:
        # m < n at low density leaves directions constrained only through a nearly singular P (|x| ~ 1e3): knife-edge
        # |y_i| ~ 100 eps active-set decisions are more frequent; demand exact status / iteration parity, and exact
        # active sets + 1e-6 solutions on >= 97 % of the well-posed instances
        assert wp.mean() >= 0.9
        assert np.array_equal(r.status[wp], o.status[wp]) and np.array_equal(r.iter[wp], o.iter[wp])
        same = wp & (r.active == o.active).all(axis=1) & (o.status == 0)
        assert same.sum() >= 0.97 * (wp & (o.status == 0)).sum()
        assert rel_err(r.x[same], o.x[same]).max() <= REL_F64 and rel_err(r.y[same], o.y[same]).max() <= REL_F64


def test_parity_random_sparse_no_polish_tight(sfb, oracle):
    from smooth_feedback_b200.generators import random_sparse_qp_numpy

    pat, Pv, q, Av, l, u = random_sparse_qp_numpy(128, 30, 45, density=0.2, seed=3)
    _, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, prm_kw=dict(eps_abs=1e-6, eps_rel=1e-6, polish=False), max_iter=20000)
    _assert_parity(r, o, wp, REL_F64, min_well_posed=0.95)


def test_parity_infeasible_mix_and_no_scaling(sfb, oracle):
    from smooth_feedback_b200.generators import random_sparse_qp_numpy

    pat, Pv, q, Av, l, u = random_sparse_qp_numpy(128, 12, 24, density=0.4, seed=11, feasible=False)
    _, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, max_iter=5000)
    assert (o.status == 2).any() and (o.status == 0).any()
    _assert_discrete_and_conditioned(r, o, _solve_both.last_fast, wp)
    _, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, prm_kw=dict(scaling=False), max_iter=5000)
    _assert_discrete_and_conditioned(r, o, _solve_both.last_fast, wp)


def test_parity_mpc_structured_small(sfb, oracle):
    # MPC-shaped QP (ocp_to_qp.hpp pattern: upper-triangular P, equality dynamics rows, explicit zeros) on a small mesh:
    # Nx=3, Nu=2, 3 intervals of 4 nodes -> n = 63, m = 63 = tests/test_mpc.cpp's problem size
    from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern

    pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
    assert pat["n"] == 63 and pat["m"] == 63
    Pv, q, Av, l, u = mpc_structured_batch(pat, 96, seed=2)
    sp, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u)
    assert (o.status == 0).all()
    _assert_parity(r, o, wp, REL_F64, min_well_posed=0.9)
    # warm re-solve: exits at the first check, like Mpc.Api's u(cold) == u(warm) (tests/test_mpc.cpp:73-118)
    _, r2, o2, wp2 = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, warm=(o.x, o.y))
    assert (r2.status == 0).all() and (r2.iter[wp2] == o2.iter[wp2]).all()
    same = wp2 & (r2.active == o2.active).all(axis=1)
    assert same.mean() > 0.9 and rel_err(r2.x[same], o2.x[same]).max() <= REL_F64


def test_parity_mpc_cfg3_real_workload(sfb, oracle):
    """BASELINE.json configs[2] as the reference builds it: SE(2) x R^3 bus, K = 50 -> 13 intervals x 4 nodes, n = m = 422
    (SURVEY D5), QPs transcribed by the restated ocp_to_qp / MPC (oracle/transcribe.py) for 256 agents sampled around the
    desired trajectory.  fp64: status / iterations / active sets exact, x and y within 1e-6 on >= 95 % well posed."""
    from workloads import vehicle_mpc_batch

    pat, Pv, q, Av, l, u, mpc, t0, x0 = vehicle_mpc_batch(256, seed=5)
    assert pat["n"] == 422 and pat["m"] == 422
    sp, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u)
    assert sp.nnzL < 16000  # minimum-degree fill (dense would be 88831)
    assert (o.status == 0).all()
    _assert_parity(r, o, wp, REL_F64, min_well_posed=0.95)
    # the control input the MPC applies (mpc.hpp:518) agrees too
    xvar = mpc.dims["xvar_L"]
    assert np.abs(r.x[:, xvar:xvar + 2] - o.x[:, xvar:xvar + 2])[wp].max() <= 1e-8
    # warm re-solve (mpc.hpp:491,510-516): exits at the first stop check with the same input
    _, r2, o2, wp2 = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, warm=(o.x, o.y))
    assert (r2.status == 0).all() and np.array_equal(r2.iter[wp2], o2.iter[wp2]) and (r2.iter == 2).mean() > 0.95
    same = wp2 & (r2.active == o2.active).all(axis=1)
    assert same.mean() > 0.9 and rel_err(r2.x[same], o2.x[same]).max() <= REL_F64


def test_parity_mpc_cfg3_synthetic_ltv(sfb, oracle):
    # same shape with the linear time-varying surrogate generator (random Jacobians per agent and node): a harder numeric
    # factorisation than the vehicle's time-invariant linearisation
    from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern

    pat = mpc_structured_pattern()
    assert pat["n"] == 422 and pat["m"] == 422
    Pv, q, Av, l, u = mpc_structured_batch(pat, 64, seed=5)
    sp, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u)
    assert (o.status == 0).all()
    _assert_parity(r, o, wp, REL_F64, min_well_posed=0.85)


def test_fp32_against_fp64_oracle(sfb, oracle):
    """BASELINE configs[2] is quoted in fp32 (new functionality): ADMM iterations in fp32, polish_qp as a mixed-
    precision fp64 pass.  MAX relative error <= 1e-3 on every instance whose discrete outcomes agree with the fp64 oracle."""
    from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern
    from workloads import vehicle_mpc_batch

    f32r = lambda t: np.asarray(t, dtype=np.float32).astype(np.float64)
    pat, Pv, q, Av, l, u, _, _, _ = vehicle_mpc_batch(128, seed=6)
    Pv, q, Av, l, u = (f32r(t) for t in (Pv, q, Av, l, u))
    _, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, dtype=np.float32)
    assert (r.flags[r.status == 0] == 1).all()
    _assert_fp32(r, o, _solve_both.last_fast, min_same=0.9)
    # every Optimal instance, whatever its active set: the fp32 result is the fp64 one at the 1e-3 level
    opt = (o.status == 0) & (r.status == 0)
    assert rel_err(r.x[opt], o.x[opt]).max() <= 10 * REL_F32
    pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
    Pv, q, Av, l, u = (f32r(t) for t in mpc_structured_batch(pat, 64, seed=4))
    _, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, dtype=np.float32)
    _assert_fp32(r, o, _solve_both.last_fast, min_same=0.9)


def test_device_path_equals_host_path_and_errors(sfb):
    import torch

    from smooth_feedback_b200.generators import random_sparse_qp_numpy

    pat, Pv, q, Av, l, u = random_sparse_qp_numpy(70, 20, 30, density=0.2, seed=8)
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    prm = sfb.QPSolverParams(max_iter=4000)
    rh = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, prm)
    t = lambda a: torch.from_numpy(a).cuda()
    rd = sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm)
    torch.cuda.synchronize()
    assert np.array_equal(rh.status, rd.status.cpu().numpy()) and np.array_equal(rh.iter, rd.iter.cpu().numpy().astype(np.uint32))
    assert np.array_equal(rh.x, rd.x.cpu().numpy()) and np.array_equal(rh.y, rd.y.cpu().numpy())
    with pytest.raises(sfb.SfbError):  # column index out of range
        bad = pat["A_colidx"].copy(); bad[0] = pat["n"]
        sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], bad)
    with pytest.raises(sfb.SfbError):  # host / device pointers mixed
        sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm, out=rh)


def test_full_size_properties_cfg3(sfb, oracle):
    """BASELINE.json configs[2] at full size (vehicle MPC n = m = 422, batch 8192; fp64 and fp32): size-independent
    properties.  The batch tiles 128 distinct agents, so replicas must agree bit for bit wherever they sit in the batch
    (tile / lane independence), every instance must be Optimal with the KKT conditions of the ORIGINAL problem satisfied,
    and a warm re-solve must exit at the first check (iter 2) with the same solution."""
    import torch

    from smooth_feedback_b200.generators import sparse_to_dense
    from workloads import vehicle_mpc_batch

    base, B = 128, 8192
    pat, Pv, q, Av, l, u, _, _, _ = vehicle_mpc_batch(base, seed=9)
    rep = B // base
    t = lambda a, dt=torch.float64: torch.from_numpy(np.tile(a, (rep, 1))).to("cuda:0", dtype=dt).contiguous()
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    prm = sfb.QPSolverParams(max_iter=4000)
    r = sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm)
    torch.cuda.synchronize()
    st, it = r.status.cpu().numpy(), r.iter.cpu().numpy()
    x, y = r.x.cpu().numpy(), r.y.cpu().numpy()
    assert (st == 0).all() and ((r.flags.cpu().numpy() & 1) == 1).all()
    # replicas identical
    assert np.array_equal(x.reshape(rep, base, -1), np.broadcast_to(x[:base], (rep, base, x.shape[1])))
    assert np.array_equal(it.reshape(rep, base), np.broadcast_to(it[:base], (rep, base)))
    # KKT of the original problem on the distinct agents: stationarity with sym(triu P), primal feasibility (no sign property
    # for the duals: polish solves an equality-constrained QP on the guessed active set and, like the reference's, does not
    # re-check multiplier signs)
    P, A = sparse_to_dense(pat, Pv, Av)
    Ps = np.triu(P) + np.transpose(np.triu(P, 1), (0, 2, 1))
    xb, yb = x[:base], y[:base]
    stat = np.einsum("bij,bj->bi", Ps, xb) + q + np.einsum("bji,bj->bi", A, yb)
    assert np.abs(stat).max() <= 1e-8 * max(1.0, np.abs(yb).max())
    Ax = np.einsum("bij,bj->bi", A, xb)
    # polish (like the reference's) enforces the rows it found ACTIVE exactly; a row the eps = 1e-3 ADMM iterate left
    # inactive may end up violated at that level, so feasibility is an eps-level property, equality rows are exact
    eq = np.isclose(l, u)
    assert np.abs(Ax - u)[eq].max() <= 1e-7 * (1.0 + np.abs(u[eq]).max())
    assert (Ax <= u + 2e-2).all() and (Ax >= l - 2e-2).all(), (np.max(Ax - u), np.max(l - Ax))
    # parity with the oracle on the distinct agents
    o = oracle.qp_solve_batch(P, q, A, l, u, params=oracle.default_params(max_iter=4000), nthreads=8)
    assert np.array_equal(st[:base], o.status) and np.array_equal(it[:base].astype(np.uint32), o.iter)
    assert rel_err(xb, o.x).max() <= REL_F64
    # warm re-solve
    r2 = sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm, warm_x=r.x, warm_y=r.y)
    torch.cuda.synchronize()
    assert (r2.status.cpu().numpy() == 0).all() and (r2.iter.cpu().numpy() == 2).mean() > 0.95
    assert np.quantile(rel_err(r2.x.cpu().numpy(), x), 0.95) <= 1e-6  # a few re-solves pick a different active set in the polish
    # fp32 (BASELINE configs[2] is quoted in fp32), polish on as MPC runs it: MAX error against the fp64 result
    f32 = torch.float32
    r32 = sfb.solve_sparse_batch(sp, t(Pv, f32), t(q, f32), t(Av, f32), t(l, f32), t(u, f32), prm)
    torch.cuda.synchronize()
    assert (r32.status.cpu().numpy() == 0).all() and (r32.flags.cpu().numpy() == 1).all()
    same = ((r32.active == r.active).all(dim=1) & (r32.iter == r.iter)).cpu().numpy()
    assert same.mean() >= 0.9
    e32 = rel_err(r32.x.cpu().numpy().astype(np.float64), x)
    assert e32[same].max() <= REL_F32 and e32.max() <= 10 * REL_F32


def test_csc_ingestion_matches_csr(sfb):
    """OSQP-style ingestion (compat/osqp.hpp:36-49 hands OSQP the constraint matrix in CSC): the same problems through
    sfb_qp_sparse_analyze_csc / sfb_qp_solve_sparse_batch_csc_f64 give bit-identical results to the CSR entry point, on host
    and on device arrays."""
    import torch

    from smooth_feedback_b200.generators import random_sparse_qp_numpy, sparse_to_dense

    pat, Pv, q, Av, l, u = random_sparse_qp_numpy(96, 30, 45, density=0.2, seed=21)
    n, m = pat["n"], pat["m"]
    prm = sfb.QPSolverParams(max_iter=4000)
    sp = sfb.SparsePattern(n, m, pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    r = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, prm)
    _, A = sparse_to_dense(pat, Pv, Av)
    mask = np.zeros((m, n), bool)
    mask[np.repeat(np.arange(m), np.diff(pat["A_rowptr"])), pat["A_colidx"]] = True
    cc, cr = np.nonzero(mask.T)  # column-major order: (col, row)
    A_colptr = np.concatenate([[0], np.cumsum(np.bincount(cc, minlength=n))]).astype(np.int32)
    Av_csc = np.ascontiguousarray(A[:, cr, cc])
    spc = sfb.SparsePattern(n, m, pat["P_colptr"], pat["P_rowidx"], A_colptr, cr.astype(np.int32), a_csc=True, handle=_handle(sfb, _KERNEL))
    rc = sfb.solve_sparse_batch(spc, Pv, q, Av_csc, l, u, prm)
    assert np.array_equal(r.x, rc.x) and np.array_equal(r.y, rc.y) and np.array_equal(r.status, rc.status) and np.array_equal(r.iter, rc.iter)
    t = lambda a: torch.from_numpy(a).cuda()
    rd = sfb.solve_sparse_batch(spc, t(Pv), t(q), t(Av_csc), t(l), t(u), prm)
    torch.cuda.synchronize()
    assert np.array_equal(rd.x.cpu().numpy(), r.x) and np.array_equal(rd.iter.cpu().numpy().astype(np.uint32), r.iter)


def _case_as_sparse(case):
    """A known-answer case as QuadraticProgramSparse: the stored patterns are what Eigen's sparseView keeps (non-zeros)."""
    from qp_cases import as_batch

    P, q, A, l, u = as_batch(case)
    n, m = P.shape[1], A.shape[1]
    pc, pr = np.nonzero(P[0].T)   # column-major order
    ar, ac = np.nonzero(A[0])
    pat = dict(n=n, m=m, P_colptr=np.concatenate([[0], np.cumsum(np.bincount(pc, minlength=n))]).astype(np.int32),
               P_rowidx=pr.astype(np.int32), A_rowptr=np.concatenate([[0], np.cumsum(np.bincount(ar, minlength=m))]).astype(np.int32),
               A_colidx=ac.astype(np.int32))
    return pat, np.ascontiguousarray(P[:, pr, pc]), q, np.ascontiguousarray(A[:, ar, ac]), l, u


from qp_cases import CASES, OPTIMAL, is_approx  # noqa: E402


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_known_answers_through_sparse_c_abi(sfb, case):
    # the reference's own unit tests (tests/test_qp.cpp: BasicSparse :100-122, PortfolioOptimizationSparse :274-312 and the
    # dense cases as sparse problems, cf. TwoDimensional :314-336 dense == sparse) against the sparse CUDA path, including
    # empty rows (Unconstrained: A = 0 has no stored entry) and the warm re-solve of every Optimal case
    pat, Pv, q, Av, l, u = _case_as_sparse(case)
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    r = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u)
    assert r.status[0] == case["status"]
    if case["x"] is not None:
        assert is_approx(r.x[0], case["x"], case["x_rtol"])
    if case["obj"] is not None:
        assert abs(r.obj[0] - case["obj"]) <= case["obj_atol"]
    if case["status"] == OPTIMAL:
        r2 = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, warm_x=r.x, warm_y=r.y)
        assert r2.status[0] == OPTIMAL and r2.iter[0] == 2
        assert is_approx(r2.x[0], case["x"], case["x_rtol"])


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_known_answers_sparse_match_oracle(sfb, oracle, case):
    from qp_cases import as_batch

    pat, Pv, q, Av, l, u = _case_as_sparse(case)
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    r = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u)
    P, q_, A, l_, u_ = as_batch(case)
    o = oracle.qp_solve_batch(P, q_, A, l_, u_)
    assert r.status[0] == o.status[0] and r.iter[0] == o.iter[0]
    assert np.array_equal(r.active, o.active)
    if case["status"] == OPTIMAL:
        assert rel_err(r.x, o.x).max() <= REL_F64 and rel_err(r.y, o.y).max() <= REL_F64


def test_edge_patterns(sfb, oracle):
    """m = 0 (no constraint rows at all), n = 1, an empty row in A, and P stored with BOTH triangles vs upper only."""
    rng = np.random.default_rng(4)
    # unconstrained, diagonal + one off-diagonal pair stored in both triangles
    n, B = 5, 33
    P = np.zeros((B, n, n)); d = 1.0 + rng.random((B, n)); P[:, np.arange(n), np.arange(n)] = d
    P[:, 0, 3] = P[:, 3, 0] = 0.3
    q = rng.uniform(-1, 1, (B, n))
    pc, pr = np.nonzero(P[0].T)
    pat = dict(n=n, m=0, P_colptr=np.concatenate([[0], np.cumsum(np.bincount(pc, minlength=n))]).astype(np.int32),
               P_rowidx=pr.astype(np.int32), A_rowptr=np.zeros(1, np.int32), A_colidx=np.zeros(0, np.int32))
    sp = sfb.SparsePattern(n, 0, pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, _KERNEL))
    z = np.zeros((B, 0))
    r = sfb.solve_sparse_batch(sp, np.ascontiguousarray(P[:, pr, pc]), q, z, z, z, sfb.QPSolverParams(max_iter=4000))
    o = oracle.qp_solve_batch(P, q, np.zeros((B, 0, n)), z, z, params=oracle.default_params(max_iter=4000))
    assert np.array_equal(r.status, o.status) and np.array_equal(r.iter, o.iter)
    assert rel_err(r.x, o.x).max() <= REL_F64 and rel_err(r.x, -np.linalg.solve(P, q[..., None])[..., 0]).max() <= 1e-3
    # n = 1 with three rows, the middle one structurally empty
    P1 = 1.0 + rng.random((B, 1, 1)); q1 = rng.uniform(-1, 1, (B, 1))
    A1 = np.zeros((B, 3, 1)); A1[:, 0, 0] = 1.0; A1[:, 2, 0] = -2.0
    l1 = np.tile([-0.1, -np.inf, -0.3], (B, 1)); u1 = np.tile([0.1, np.inf, 0.3], (B, 1))
    sp1 = sfb.SparsePattern(1, 3, np.array([0, 1], np.int32), np.array([0], np.int32), np.array([0, 1, 1, 2], np.int32), np.array([0, 0], np.int32), handle=_handle(sfb, _KERNEL))
    r1 = sfb.solve_sparse_batch(sp1, P1[:, :, 0], q1, np.ascontiguousarray(A1[:, [0, 2], 0]), l1, u1, sfb.QPSolverParams(max_iter=4000))
    o1 = oracle.qp_solve_batch(P1, q1, A1, l1, u1, params=oracle.default_params(max_iter=4000))
    assert np.array_equal(r1.status, o1.status) and np.array_equal(r1.iter, o1.iter) and np.array_equal(r1.active, o1.active)
    assert np.abs(r1.x - o1.x).max() <= 1e-9



def test_onchip_kernel_is_the_one_that_runs_and_agrees_with_the_tiled_kernel(sfb, sparse_kernel):
    """The real vehicle MPC pattern (n = m = 422) fits in shared memory in both precisions: the default handle must run the
    on-chip kernel (a silent fallback to the tiled kernel would hide it from every other test), a handle created with
    SFB_SPARSE_KERNEL=tiled must not; both kernels return the same discrete outcomes and solutions to rounding."""
    import torch
    from workloads import vehicle_mpc_batch

    pat, Pv, q, Av, l, u = vehicle_mpc_batch(24)[:6]
    outs = {}
    for kind in ("onchip", "tiled"):
        sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=_handle(sfb, kind))
        for sb in (4, 8):
            used, info = sp.uses_onchip(sb)
            assert used == (kind == "onchip"), (kind, sb, info)
            if kind == "onchip":
                assert info["levels"] == 4 and info["sweep_stages"] == 15 and 0 < info["smem_bytes"] <= 227 * 1024, info
        outs[kind] = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, sfb.QPSolverParams(max_iter=4000))
    a, b = outs["onchip"], outs["tiled"]
    assert np.array_equal(a.status, b.status) and np.array_equal(a.iter, b.iter) and np.array_equal(a.active, b.active)
    assert (a.status == 0).all() and (a.flags & 1).all()
    assert rel_err(a.x, b.x).max() <= 1e-9 and rel_err(a.y, b.y).max() <= 1e-9
    # a default handle (no environment override) picks the on-chip kernel
    sp0 = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
    assert sp0.uses_onchip(8)[0] and sp0.uses_onchip(4)[0]


def test_parity_wide_supernodes(sfb, oracle, sparse_kernel):
    """Nearly dense patterns: the factor collapses into one or two supernodes of 100+ columns (more than 32 vector column groups
    per supernode: the on-chip factorisation walks them in strides; the sweeps run more outputs than threads per stage)."""
    from smooth_feedback_b200.generators import random_sparse_qp_numpy

    for (n, m, dens, seed) in [(140, 90, 0.5, 11), (100, 160, 0.9, 12)]:
        pat, Pv, q, Av, l, u = random_sparse_qp_numpy(48, n, m, density=dens, seed=seed)
        sp, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u)
        if sparse_kernel == "onchip":  # (the second pattern's fp64 working set exceeds shared memory: that solve falls back to the tiled kernel)
            used64, info = sp.uses_onchip(8)
            used32, _ = sp.uses_onchip(4)
            assert info["largest_supernode"] > 32 * 2 and used32 and (used64 or n == 100), (info, used32, used64)
        _assert_parity(r, o, wp, REL_F64, min_well_posed=0.9)
        r32 = sfb.solve_sparse_batch(sp, *(np.ascontiguousarray(a, dtype=np.float32) for a in (Pv, q, Av, l, u)), sfb.QPSolverParams(max_iter=4000))
        same = (r32.status == o.status) & (o.status == 0)
        assert same.mean() >= 0.9 and rel_err(r32.x[same].astype(np.float64), o.x[same]).max() <= 5e-2  # fp32 on |x| ~ 1e2 problems: sanity, not parity


def test_max_time_and_iteration_budget(sfb, sparse_kernel):
    """qp_solver.hpp:504-508 / :449: a time budget that is already spent ends the solve at the first stop check with MaxTime, an
    iteration budget below the first successful check with MaxIterations; both return finite iterates (both sparse kernels)."""
    from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern

    pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
    Pv, q, Av, l, u = mpc_structured_batch(pat, 40, seed=7)
    sp = _pattern(sfb, pat)
    r = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, sfb.QPSolverParams(max_time=1e-9))
    assert (r.status == int(sfb.QPSolutionStatus.MaxTime)).all() and (r.iter == 2).all() and np.isfinite(r.x).all()
    r = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, sfb.QPSolverParams(max_iter=20))
    assert (r.status == int(sfb.QPSolutionStatus.MaxIterations)).all() and (r.iter == 20).all() and np.isfinite(r.x).all()
