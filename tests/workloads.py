"""Real (not synthetic-surrogate) MPC / ASIF workloads for the parity tests, built with the host restatement of the
reference's transcriptions (oracle/transcribe.py): SURVEY 8(d) cfg3 / cfg5 sampling, x0 = xdes(t0) (+) xi, t0 ~ U(0, 30)."""
import numpy as np


def vehicle_mpc_batch(B: int, seed: int = 5, K: int = 50, tf: float = 5.0):
    """BASELINE configs[2]: the QP MPC<Time, Bundle<SE2, R^3>, R^2, F, CR>::operator() builds (examples/mpc_asif_vehicle.cpp:
    42-89 at K = 50 -> n = m = 422), one per agent, with a shared sparsity pattern.

    -> (pat dict, P_vals [B,nnzP], q [B,n], A_vals [B,nnzA], l [B,m], u [B,m], mpc, t0, x0)"""
    from oracle import transcribe as tr

    mpc = tr.vehicle_mpc(K=K, tf=tf)
    t0, x0 = tr.sample_vehicle_states(B, seed=seed)
    Pv, qs, Av, ls, us = [], [], [], [], []
    pat = None
    for b in range(B):
        qp = mpc.transcribe(t0[b], x0[b])
        rp, ci, av = qp.csr_A()
        cp, ri, pv = qp.csc_P()
        if pat is None:
            pat = dict(n=qp.n, m=qp.m, P_colptr=cp, P_rowidx=ri, A_rowptr=rp, A_colidx=ci)
        else:
            assert np.array_equal(rp, pat["A_rowptr"]) and np.array_equal(ci, pat["A_colidx"])
        Pv.append(pv); Av.append(av); qs.append(qp.q.copy()); ls.append(qp.l.copy()); us.append(qp.u.copy())
    return (pat,) + tuple(np.stack(t) for t in (Pv, qs, Av, ls, us)) + (mpc, t0, x0)
