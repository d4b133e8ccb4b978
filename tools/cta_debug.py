"""Dev tool: small problems through the on-chip sparse kernel vs the tiled kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import mpc_structured_batch, mpc_structured_pattern, random_sparse_qp_numpy

def handle(kind):
    os.environ["SFB_SPARSE_KERNEL"] = kind
    h = sfb.Handle(0)
    os.environ.pop("SFB_SPARSE_KERNEL", None)
    return h
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0", dtype=dt)
cases = []
pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
cases.append(("mpc63", pat) + tuple(mpc_structured_batch(pat, 64, seed=2)))
p2, Pv, q, Av, l, u = random_sparse_qp_numpy(64, 20, 30, density=0.2, seed=3)
cases.append(("rand20x30", p2, Pv, q, Av, l, u))
p3, Pv, q, Av, l, u = random_sparse_qp_numpy(64, 7, 5, density=0.5, seed=4)
cases.append(("rand7x5", p3, Pv, q, Av, l, u))
for name, pat, Pv, q, Av, l, u in cases:
    for dt in (torch.float64, torch.float32):
        outs = {}
        for kind in ("tiled", "cta"):
            h = handle(kind)
            sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=h)
            outs[kind] = sfb.solve_sparse_batch(sp, t(Pv, dt), t(q, dt), t(Av, dt), t(l, dt), t(u, dt), sfb.QPSolverParams(max_iter=2000))
            torch.cuda.synchronize()
        a, b = outs["tiled"], outs["cta"]
        same = (a.status == b.status) & (a.iter == b.iter)
        dx = ((a.x - b.x).abs().amax(dim=1) / a.x.abs().amax(dim=1).clamp_min(1e-30))
        print(name, dt, "status tiled", torch.bincount(a.status, minlength=7).tolist(), "cta", torch.bincount(b.status, minlength=7).tolist(),
              "same", int(same.sum()), "/", len(same), "max dx", float(dx[same].max()) if same.any() else None,
              "iters", a.iter[:6].tolist(), b.iter[:6].tolist(), "flags", b.flags[:4].tolist())
