"""Dev tool: phase split of the sparse QP kernel (setup / ADMM iterations / polish) by toggling parameters."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
pat = mpc_structured_pattern()
Pv, q, Av, l, u = mpc_structured_batch(pat, 512, seed=5)
rep = (B + 511) // 512
t = lambda x: torch.from_numpy(np.tile(x, (rep, 1))[:B]).cuda().contiguous()
Pv, q, Av, l, u = t(Pv), t(q), t(Av), t(l), t(u)
sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
def run(**kw):
    prm = sfb.QPSolverParams(**kw)
    out = None
    for _ in range(2):
        out = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, prm, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = sfb.solve_sparse_batch(sp, Pv, q, Av, l, u, prm, out=out); e1.record(); e1.synchronize()
    return e0.elapsed_time(e1), out.iter.double().mean().item()
res = {}
res["setup_noscale (max_iter=0, scaling off)"] = run(max_iter=0, scaling=False, polish=False)
res["setup (max_iter=0)"] = run(max_iter=0, polish=False)
res["setup + 26 iters no check (max_iter=26, stop_check_iter=1000)"] = run(max_iter=26, polish=False, stop_check_iter=1000)
res["full no polish"] = run(max_iter=4000, polish=False)
res["full"] = run(max_iter=4000)
for k, v in res.items():
    print(f"{k:70s} {v[0]:9.2f} ms   mean iter {v[1]:.1f}")
