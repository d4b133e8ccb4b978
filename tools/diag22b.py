import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smooth_feedback_b200 as sfb
from smooth_feedback_b200 import _lib
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy
cm = sfb.to_colmajor
n, m = 2, 2
P, q, A, l, u = random_qp_numpy(64, n, m, seed=n * 1000 + m)
o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=1)
o2 = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=1, fast=True)
h = sfb.Handle(0)
for force in (0, 1):
    h.set_option(_lib.OPT_FORCE_POLISH_SCRATCH, force)
    r = sfb.solve_dense_batch(cm(P), q, cm(A), l, u, sfb.QPSolverParams(max_iter=4000), handle=h)
    ex = np.linalg.norm(r.x - o.x, axis=1) / np.linalg.norm(o.x, axis=1)
    bad = np.nonzero(ex > 1e-4)[0]
    print("force", force, "bad", bad, "flags", r.flags[bad], "na", (o.active[bad] != 0).sum(1), "gpu act", r.active[bad].tolist(), "orc act", o.active[bad].tolist())
    for b in bad:
        print("  inst", b, "x gpu", r.x[b], "x orc", o.x[b], "x fast", o2.x[b], "y gpu", r.y[b], "y orc", o.y[b], "iter", r.iter[b], o.iter[b], "status", r.status[b], o.status[b])
        print("   P", P[b].tolist(), "q", q[b].tolist(), "A", A[b].tolist(), "u", u[b].tolist())
