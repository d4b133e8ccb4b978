"""Dev tool: the on-chip sparse kernel on the real vehicle MPC workload (fp64 solve, fp32 solve + its fp64 polish pass, twice
each).  Default: per-phase cycle counts from the in-kernel clock64 accounting (SFB_CTA_PROF=1, printed by libsfb on stderr);
--plain: no instrumentation (the target of `ncu -k regex:qp_sparse_cta_kernel -s 1 -c 3`).

    python tools/cta_prof.py [batch] [--plain]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if "--plain" in sys.argv:
    sys.argv.remove("--plain")
else:
    os.environ["SFB_CTA_PROF"] = "1"
import numpy as np, torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import vehicle_fleet_numpy
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
base = 512
t0, x0, _ = vehicle_fleet_numpy(base, seed=5)
fl = sfb.MPCVehicleFleet(base)
pat = fl.pattern()
Pv, q, Av, l, u = fl.to_qp(t0, x0)
fl.close()
rep = (batch + base - 1) // base
for dt in (torch.float64, torch.float32):
    t = lambda a: torch.from_numpy(np.tile(a, (rep, 1))[:batch]).to("cuda:0", dtype=dt).contiguous()
    d = [t(a) for a in (Pv, q, Av, l, u)]
    h = sfb.Handle(0)
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=h)
    for _ in range(2):
        out = sfb.solve_sparse_batch(sp, *d, sfb.QPSolverParams(max_iter=4000))
    torch.cuda.synchronize()
