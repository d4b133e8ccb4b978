"""CPU tests of the sparse path's host logic: workload generators, the densified-oracle equivalence they rely on, and the
symbolic analysis (sfb_qp_sparse_symbolic: ordering + fill) checked against a numeric Cholesky."""
import numpy as np
import pytest

import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import (_lgr_diff_matrix, mpc_structured_batch, mpc_structured_pattern,
                                             random_sparse_qp_numpy, sparse_to_dense)


def test_lgr_differentiation_matrix_is_exact_on_polynomials():
    # collocation/mesh.hpp: K LGR nodes + end point; D differentiates polynomials of degree <= K exactly
    for K in (2, 4, 7):
        D, qw, tau = _lgr_diff_matrix(K)
        assert abs(tau[0] + 1) < 1e-14 and abs(tau[-1] - 1) < 1e-14 and abs(qw.sum() - 2) < 1e-12
        for p in range(K + 1):
            assert np.abs(D @ tau ** p - p * tau[:K] ** max(p - 1, 0) * (p > 0)).max() < 1e-10


def test_mpc_pattern_matches_ocp_to_qp_sizes():
    # ocp_to_qp.hpp:58-96 for the SE(2) x R^3 bus, K = 50 -> 13 intervals x 4 nodes: n = m = 422
    pat = mpc_structured_pattern()
    Nx, Nu, N, Ki = 6, 2, 52, 4
    assert pat["n"] == Nx * (N + 1) + Nu * N == 422 and pat["m"] == Nx * N + Nu * N + Nx == 422
    rows = np.diff(pat["A_rowptr"])
    assert (rows[: Nx * N] == Nx + Ki + Nu).all()                 # dynamics rows, :77-80
    assert (rows[Nx * N: Nx * N + Nu * N] == Nx + Nu).all()       # running constraints, :81
    assert (rows[-Nx:] == 2 * Nx).all()                           # end constraints, :82
    cols = np.diff(pat["P_colptr"])                               # upper-triangular P, :86-96
    assert list(cols[:Nx]) == [1, 2, 3, 4, 5, 6] and list(cols[Nx * N: Nx * (N + 1)]) == [7, 8, 9, 10, 11, 12]
    assert list(cols[Nx * (N + 1): Nx * (N + 1) + Nu]) == [7, 8]
    pc = np.repeat(np.arange(pat["n"]), cols)
    assert (pat["P_rowidx"] <= pc).all()                          # nothing below the diagonal


def test_mpc_surrogate_is_solved_by_the_oracle(oracle):
    # the synthetic MPC workload must be a problem the reference algorithm solves (Optimal, ~50 iterations)
    pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
    Pv, q, Av, l, u = mpc_structured_batch(pat, 24, seed=1)
    P, A = sparse_to_dense(pat, Pv, Av)
    o = oracle.qp_solve_batch(P, q, A, l, u, params=oracle.default_params(max_iter=4000), nthreads=4)
    assert (o.status == 0).all() and o.iter.max() <= 202
    x = o.x
    eq = np.isclose(l, u)
    Ax = np.einsum("bij,bj->bi", A, x)
    assert np.abs(Ax - u)[eq].max() < 1e-6                         # dynamics / initial state satisfied
    assert (Ax <= u + 1e-6).all() and (Ax >= l - 1e-6).all()


def test_upper_only_P_quirk_with_non_diagonal_weights(oracle):
    # check_stopping multiplies with P as stored (upper triangle only for MPC), so a non-diagonal
    # weight never passes the dual residual test -> MaxIterations.  Reproduced, not fixed.
    pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=2, Ki=3)
    Pv, q, Av, l, u = mpc_structured_batch(pat, 4, seed=1)
    kinds = [t[0] for t in pat["P_terms"]]
    offdiag = np.array([k == "Q" and t[2] != t[3] for k, t in zip(kinds, pat["P_terms"])])
    Pv2 = Pv.copy(); Pv2[:, offdiag] = 0.05
    P, A = sparse_to_dense(pat, Pv2, Av)
    o = oracle.qp_solve_batch(P, q, A, l, u, params=oracle.default_params(max_iter=600), nthreads=4)
    assert (o.status == 4).all()


@pytest.mark.parametrize("which", ["mpc", "random"])
def test_symbolic_analysis_covers_the_numeric_factor(which):
    if which == "mpc":
        pat = mpc_structured_pattern()
        Pv, q, Av, l, u = mpc_structured_batch(pat, 1, seed=3)
    else:
        pat, Pv, q, Av, l, u = random_sparse_qp_numpy(1, 60, 80, density=0.08, seed=2)
    n, m = pat["n"], pat["m"]
    sym = sfb.sparse_symbolic(n, m, pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
    perm = sym["perm"]
    assert sorted(perm.tolist()) == list(range(n))
    P, A = sparse_to_dense(pat, np.abs(Pv) + 1.0, np.abs(Av) + 1.0)   # no accidental cancellation in the pattern
    Pu = np.triu(P[0]); M = Pu + Pu.T + A[0].T @ A[0] + n * 100.0 * np.eye(n)
    Mp = M[np.ix_(perm, perm)]
    L = np.linalg.cholesky(Mp)
    numeric = [int((np.abs(L[j + 1:, j]) > 1e-14).sum()) for j in range(n)]
    symbolic = np.diff(sym["L_colptr"])
    assert (np.array(numeric) <= symbolic).all()                      # every numeric nonzero has a slot
    assert sym["nnzL"] == symbolic.sum() and sym["nnzL"] <= 1.05 * sum(numeric) + 8
    dense_lower = n * (n - 1) // 2
    assert sym["nnzL"] < (0.12 if which == "mpc" else 0.9) * dense_lower
    assert sym["factor_flops"] == int((symbolic * (symbolic + 1) // 2).sum())


def test_symbolic_rejects_bad_patterns():
    pat, *_ = random_sparse_qp_numpy(1, 8, 6, density=0.4, seed=1)
    bad = pat["A_colidx"].copy(); bad[0] = 99
    with pytest.raises(sfb.SfbError):
        sfb.sparse_symbolic(8, 6, pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], bad)


@pytest.mark.parametrize("n,m,density,seed", [(1, 1, 1.0, 0), (7, 3, 0.5, 1), (30, 80, 0.1, 2), (64, 64, 0.05, 3), (40, 5, 1.0, 4),
                                               (90, 200, 0.03, 5), (50, 100, 1.0, 6)])
def test_schedules_are_self_consistent(n, m, density, seed):
    # sfb_qp_sparse_symbolic validates every schedule the device kernel relies on (qp_sparse_host.hpp::sparse_validate):
    # each factor slot streamed once per sweep, each A entry once per padded stream, targets in range, pairs inside their
    # row / column, distinct targets per column update.  Dense rows (> 32 entries) switch the padded SpMV streams off.
    pat, *_ = random_sparse_qp_numpy(1, n, m, density=density, seed=seed)
    sym = sfb.sparse_symbolic(n, m, pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
    assert sorted(sym["perm"].tolist()) == list(range(n))
    assert 0 <= sym["nnzL"] <= n * (n - 1) // 2


def test_malformed_patterns_are_rejected_not_dereferenced():
    """sfb_qp_sparse_symbolic validates pointer arrays (monotone, starting at 0, nnz >= 0, non-NULL index arrays) before it
    copies or dereferences anything: a malformed pattern is SFB_ERR_INVALID_ARGUMENT, not undefined behaviour."""
    import ctypes as C

    from smooth_feedback_b200 import _lib

    L = _lib.lib()
    i32 = lambda a: np.asarray(a, np.int32)
    ptr = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)

    def rc(n, m, pc, pr, ar, ac):
        keep = [i32(t) if t is not None else None for t in (pc, pr, ar, ac)]
        return L.sfb_qp_sparse_symbolic(n, m, ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), ptr(keep[3]), None, None, None, None)

    assert rc(2, 1, [0, 1, 2], [0, 1], [0, 2], [0, 1]) == 0                    # well formed
    assert rc(2, 1, [0, 2, 1], [0, 1], [0, 2], [0, 1]) == 1                    # P_colptr not monotone
    assert rc(2, 1, [0, 1, -5], [0, 1], [0, 2], [0, 1]) == 1                   # negative nnz
    assert rc(2, 1, [1, 1, 2], [0, 1], [0, 2], [0, 1]) == 1                    # does not start at 0
    assert rc(2, 1, [0, 1, 2], None, [0, 2], [0, 1]) == 1                      # NULL row indices with nnz > 0
    assert rc(2, 1, [0, 1, 2], [0, 1], [0, 2], None) == 1                      # NULL column indices with nnz > 0
    assert rc(2, 1, [0, 1, 2], [0, 1], [0, 2], [0, 2]) == 1                    # column index out of range
    assert rc(2, 1, [0, 1, 2], [0, 1], [0, 2], [1, 1]) == 1                    # duplicate column in a row
    assert rc(2, 0, [0, 0, 0], None, None, None) == 0                          # empty P, no rows


# ---- on-chip kernel (qp_sparse_cta.cuh): its schedules are executed on the host, table by table, against a dense solve ----
def _onchip_patterns():
    pats = [("mpc63", mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)), ("mpc422", mpc_structured_pattern())]
    for (n, m, dens, seed) in [(20, 30, 0.2, 3), (7, 5, 0.5, 4), (60, 30, 0.1, 5), (33, 80, 0.08, 6)]:
        pats.append((f"rand{n}x{m}", random_sparse_qp_numpy(2, n, m, density=dens, seed=seed)[0]))
    n = 5  # no constraints at all, diagonal P: every column its own supernode
    pats.append(("m0", dict(n=n, m=0, P_colptr=np.arange(n + 1), P_rowidx=np.arange(n), A_rowptr=np.zeros(1, int), A_colidx=np.zeros(0, int))))
    pats.append(("n1", dict(n=1, m=2, P_colptr=np.array([0, 1]), P_rowidx=np.array([0]), A_rowptr=np.array([0, 1, 2]), A_colidx=np.array([0, 0]))))
    dn = 9  # dense P and A: one supernode
    pats.append(("dense", dict(n=dn, m=4, P_colptr=np.arange(dn + 1) * dn, P_rowidx=np.tile(np.arange(dn), dn),
                               A_rowptr=np.arange(5) * dn, A_colidx=np.tile(np.arange(dn), 4))))
    return pats


@pytest.mark.parametrize("ordering", [-1, 0, 1])
def test_onchip_schedules_solve_the_reduced_system(ordering):
    """sfb_qp_sparse_cta_selfcheck: assembly (coloured pairs), level-synchronous supernodal factorisation, in-place inversion
    of the diagonal blocks and the staged sweeps -- run on the host from the kernel's own tables (fp32 and fp64 layouts) --
    reproduce a dense Cholesky solve of the same random SPD matrix."""
    for name, pat in _onchip_patterns():
        info, err = sfb.sparse_onchip_selfcheck(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], ordering)
        assert err < 1e-10, (name, ordering, err, info)
        assert info["nnzL"] <= info["factor_slots"] and info["supernodes"] >= 1 and info["levels"] >= 1, (name, info)
        assert info["sweep_stages"] <= 4 * info["levels"], (name, info)


def test_onchip_ordering_shortens_the_elimination_tree_of_the_mpc_problem():
    """The point of the on-chip analysis: minimum degree orders the REAL K = 50 vehicle MPC problem (the pattern ocp_to_qp
    produces, n = m = 422) along the time axis (a chain of 13 supernodes, 51 sweep stages); nested dissection + amalgamation
    gives 4 levels / 15 stages, and the cost model picks it."""
    from workloads import vehicle_mpc_batch

    pat = vehicle_mpc_batch(1)[0]
    args = (pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
    md, e0 = sfb.sparse_onchip_selfcheck(*args, 0)
    nd, e1 = sfb.sparse_onchip_selfcheck(*args, 1)
    auto, _ = sfb.sparse_onchip_selfcheck(*args, -1)
    assert max(e0, e1) < 1e-10
    assert md["levels"] == 13 and nd["levels"] == 4 and nd["sweep_stages"] == 15 and md["sweep_stages"] == 51
    assert auto["ordering"] == 1 and auto["levels"] == nd["levels"]
    assert nd["factor_slots"] < 1.3 * md["factor_slots"]  # the shorter tree costs some fill, not a multiple
