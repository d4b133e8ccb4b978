"""Diagnostics: status parity against the oracle with the dx != 0 guard of the dual-infeasibility test on / off."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smooth_feedback_b200 as sfb
from smooth_feedback_b200 import _lib
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy
cm = sfb.to_colmajor
hs = {}
for generic in (0, 1):
    if generic: os.environ["SFB_DENSE_FORCE_GENERIC"] = "1"
    hs[generic] = sfb.Handle(0)
    os.environ.pop("SFB_DENSE_FORCE_GENERIC", None)
for (n, m) in [(1, 5), (2, 40), (3, 64), (3, 203), (4, 129), (4, 256), (10, 20)]:
    B = 512
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=7000 + 10 * n + m + 1)
    prm = sfb.QPSolverParams(max_iter=5000, polish=False)
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=5000, polish=0), nthreads=8)
    o2 = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=5000, polish=0), nthreads=8, fast=True)
    line = f"n={n} m={m}: oracle DualInf {(o.status == 3).sum()} (fast {(o2.status == 3).sum()}, both {((o.status == 3) & (o2.status == 3)).sum()})"
    for generic in (0, 1):
        for guard in (1, 0):
            hs[generic].set_option(_lib.OPT_DUAL_INF_DX_GUARD, guard)
            r = sfb.solve_dense_batch(cm(P), q, cm(A), l, u, prm, handle=hs[generic])
            line += f" | {'gen' if generic else 'skn'} guard={guard}: DualInf {(r.status == 3).sum()} status!=oracle {(r.status != o.status).sum()} iter!= {(r.iter != o.iter).sum()}"
    print(line)
