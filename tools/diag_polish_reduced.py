"""Diagnostics: dense polish with the Schur complement S = delta I + Aa Kinv Aa^T (primal block eliminated first: default when
na <= n) against the REDUCED form N = Pbar + delta I + Aa^T Aa / delta (duals eliminated first: one n x n inverse, no S) on
every instance -- accuracy against the oracle and time at the headline shape."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200 import _lib
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy, random_qp_torch
cm = sfb.to_colmajor
h = sfb.Handle(0)
rel = lambda a, b: np.linalg.norm(a - b, axis=1) / np.maximum(np.linalg.norm(b, axis=1), 1e-9)
for (n, m, seed) in [(50, 100, 5), (10, 20, 7), (33, 31, 2), (64, 64, 3), (20, 90, 4)]:
    B = 512
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=seed)
    prm = sfb.QPSolverParams(max_iter=4000)
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=os.cpu_count())
    o2 = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=os.cpu_count(), fast=True)
    res = {}
    for mode in (1, 0, 2):
        h.set_option(_lib.OPT_POLISH_FORM, mode)
        res[mode] = sfb.solve_dense_batch(cm(P), q, cm(A), l, u, prm, handle=h)
    r0, r2, rd = res[1], res[2], res[0]
    ok = (o.status == 0) & (r0.status == 0) & (r0.iter == o.iter) & (r0.active == o.active).all(1)
    na = (o.active != 0).sum(1)
    print(f"n={n} m={m} ok={ok.sum()}/{B} na<= {na.max()}: status equal {np.array_equal(r0.status, r2.status)} iter equal {np.array_equal(r0.iter, r2.iter)} "
          f"polished {int((r0.flags & 1).sum())}/{int((r2.flags & 1).sum())} | x vs oracle: schur {rel(r0.x, o.x)[ok].max():.2e} reduced {rel(r2.x, o.x)[ok].max():.2e} "
          f"oracle self {rel(o2.x, o.x)[ok].max():.2e} | y: schur {rel(r0.y, o.y)[ok].max():.2e} reduced {rel(r2.y, o.y)[ok].max():.2e} self {rel(o2.y, o.y)[ok].max():.2e} "
          f"| obj: schur {np.abs(r0.obj - o.obj)[ok].max():.2e} reduced {np.abs(r2.obj - o.obj)[ok].max():.2e}"
          f" || DEFAULT (reduced + check + fallback): x {rel(rd.x, o.x)[ok].max():.2e} y {rel(rd.y, o.y)[ok].max():.2e} obj {np.abs(rd.obj - o.obj)[ok].max():.2e}"
          f" reduced kept on {int(((rd.flags & 16) != 0).sum())}/{B}", flush=True)
# time at the headline shape
dev = torch.device("cuda:0")
Pc, q, Ac, l, u = random_qp_torch(65536, 50, 100, seed=5, device=dev)
prm = sfb.QPSolverParams(max_iter=4000)
for mode in (1, 0, 2, 1, 0, 2):
    h.set_option(_lib.OPT_POLISH_FORM, mode)
    h.set_stream(torch.cuda.current_stream(dev).cuda_stream)
    out = None
    for _ in range(2):
        out = sfb.solve_dense_batch(Pc, q, Ac, l, u, prm, handle=h, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        out = sfb.solve_dense_batch(Pc, q, Ac, l, u, prm, handle=h, out=out)
    e1.record(); e1.synchronize()
    print(f"mode {mode}: {e0.elapsed_time(e1) / 5:.3f} ms per 65536 solves, polished {int((out.flags & 1).sum())} reduced {int(((out.flags & 16) != 0).sum())}", flush=True)
