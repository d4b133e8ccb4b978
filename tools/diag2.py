"""GPU diagnostic #2: ragged-shape mismatches and fp32 behaviour (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import smooth_feedback_b200 as sfb
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy
cm = sfb.to_colmajor
def run(B, n, m, seed, dtype=np.float64, **kw):
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=seed)
    prm = sfb.QPSolverParams(max_iter=4000, **kw)
    c = lambda t: np.ascontiguousarray(t, dtype=dtype)
    r = sfb.solve_dense_batch(c(cm(P)), c(q), c(cm(A)), c(l), c(u), prm)
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=16)
    o2 = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=16, fast=True)
    return r, o, o2
r, o, o2 = run(64, 3, 203, 3203)
wp = (o.status == o2.status) & (o.iter == o2.iter) & (o.active == o2.active).all(1)
bad = np.nonzero((r.status != o.status) | (r.iter != o.iter))[0]
print("n=3 m=203: wp", wp.mean(), "bad", bad)
for b in bad:
    print("  inst", b, "gpu", r.status[b], r.iter[b], "oracle", o.status[b], o.iter[b], "fast", o2.status[b], o2.iter[b], "wp", wp[b])
# fp32
r, o, o2 = run(256, 10, 20, 41, dtype=np.float32)
ex = np.linalg.norm(r.x - o.x, axis=1) / np.linalg.norm(o.x, axis=1)
same = (r.active == o.active).all(1)
print("fp32 n=10 m=20: status hist", np.bincount(r.status, minlength=7), "iter equal frac", (r.iter == o.iter).mean(), "active equal frac", same.mean(), "flags", np.bincount(r.flags, minlength=8))
print("  relx quantiles (all)", np.quantile(ex, [0.5, 0.9, 0.99, 1.0]), " (same active)", np.quantile(ex[same], [0.5, 0.9, 0.99, 1.0]) if same.any() else None)
r2, _, _ = run(256, 10, 20, 41, dtype=np.float32, polish=False)
op = orc.qp_solve_batch(*random_qp_numpy(256, 10, 20, seed=41), params=orc.default_params(max_iter=4000, polish=0), nthreads=16)
ex2 = np.linalg.norm(r2.x - op.x, axis=1) / np.linalg.norm(op.x, axis=1)
print("  no polish: relx quantiles vs unpolished oracle", np.quantile(ex2, [0.5, 0.9, 0.99, 1.0]), "iter equal", (r2.iter == op.iter).mean())
na_g = (r.active != 0).sum(1); na_o = (o.active != 0).sum(1)
print("  na gpu mean", na_g.mean(), "na oracle mean", na_o.mean(), " worst inst:", ex.argmax(), "na", na_g[ex.argmax()], na_o[ex.argmax()], "y gpu", r.y[ex.argmax()][:8], "y or", o.y[ex.argmax()][:8])
