"""Dev tool: mismatch statistics of the sparse QP path vs the oracle on the parity-test workloads."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import smooth_feedback_b200 as sfb
from oracle import oracle
from smooth_feedback_b200.generators import random_sparse_qp_numpy
from test_gpu_qp_sparse_parity import _solve_both
from test_gpu_qp_parity import rel_err

def report(name, r, o, wp):
    st = r.status != o.status; it = r.iter != o.iter; ac = (r.active != o.active).any(1)
    print(f"{name}: B={len(wp)} well-posed {wp.mean():.3f}; mismatches on well-posed: status {st[wp].sum()} iter {it[wp].sum()} active {ac[wp].sum()}")
    bad = np.nonzero((st | it | ac) & wp)[0]
    for b in bad[:6]:
        print(f"   inst {b}: status {r.status[b]} vs {o.status[b]}, iter {r.iter[b]} vs {o.iter[b]}, active diff {(r.active[b] != o.active[b]).sum()}, flags {r.flags[b]}, relx {rel_err(r.x[b], o.x[b]):.2e}")
    ok = (o.status == 0) & wp & ~st & ~it & ~ac
    if ok.any():
        print(f"   on agreeing Optimal: max rel x {rel_err(r.x[ok], o.x[ok]).max():.2e}, y {rel_err(r.y[ok], o.y[ok]).max():.2e}, flags {np.bincount(r.flags[ok])}")

pat, Pv, q, Av, l, u = random_sparse_qp_numpy(200, 60, 30, density=0.1, seed=90)
_, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u); report("rand 60x30", r, o, wp)
pat, Pv, q, Av, l, u = random_sparse_qp_numpy(128, 12, 24, density=0.4, seed=11, feasible=False)
_, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, max_iter=5000); report("infeasible mix", r, o, wp)
_, r, o, wp = _solve_both(sfb, oracle, pat, Pv, q, Av, l, u, prm_kw=dict(scaling=False), max_iter=5000); report("no scaling", r, o, wp)
