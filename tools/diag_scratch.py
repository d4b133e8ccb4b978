"""Diagnostics: dense polish with the Schur block on chip vs in the global workspace (SFB_OPT_FORCE_POLISH_SCRATCH)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import smooth_feedback_b200 as sfb
from smooth_feedback_b200 import _lib
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy
cm = sfb.to_colmajor
h = sfb.Handle(0)
rel = lambda a, b: np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-9)
for (n, m) in [(2, 2), (3, 2), (2, 3), (4, 4), (10, 20), (50, 100), (64, 64), (70, 40), (33, 31)]:
    B = 128
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=n * 1000 + m)
    prm = sfb.QPSolverParams(max_iter=4000)
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=8)
    o2 = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=8, fast=True)
    res = {}
    for force in (0, 1):
        h.set_option(_lib.OPT_FORCE_POLISH_SCRATCH, force)
        res[force] = sfb.solve_dense_batch(cm(P), q, cm(A), l, u, prm, handle=h)
    r0, r1 = res[0], res[1]
    ok = (o.status == 0) & (r0.status == 0) & (r0.iter == o.iter) & (r0.active == o.active).all(1)
    na = (o.active != 0).sum(1)
    print(f"n={n} m={m}: flags on-chip {np.bincount(r0.flags.astype(int), minlength=16)[[1, 2, 4, 9]]} forced {np.bincount(r1.flags.astype(int), minlength=16)[[1, 2, 4, 9]]} "
          f"max|x0-x1| rel {rel(r1.x, r0.x)[ok].max():.2e} y {rel(r1.y, r0.y)[ok].max():.2e} | vs oracle: on-chip {rel(r0.x, o.x)[ok].max():.2e} forced {rel(r1.x, o.x)[ok].max():.2e} "
          f"oracle self (fma) {rel(o2.x, o.x)[ok].max():.2e}; y: on-chip {rel(r0.y, o.y)[ok].max():.2e} forced {rel(r1.y, o.y)[ok].max():.2e} self {rel(o2.y, o.y)[ok].max():.2e} na max {na.max()}")
