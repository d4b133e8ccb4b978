"""Diagnostics: fp32 entry points against the fp64 oracle (fractions of matching discrete outcomes, error distribution)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import smooth_feedback_b200 as sfb
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy

def rel_err(a, ref):
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return np.linalg.norm(a - ref, axis=-1) / np.maximum(np.linalg.norm(ref, axis=-1), 1e-9)

def report(name, r, o, o2):
    wp = (o.status == o2.status) & (o.iter == o2.iter) & (o.active == o2.active).all(1)
    opt = wp & (o.status == 0)
    same_it = r.iter == o.iter
    same_act = (r.active == o.active).all(1)
    same = opt & same_it & same_act
    ex, ey = rel_err(r.x, o.x), rel_err(r.y, o.y)
    q = lambda e, m: np.quantile(e[m], [0.5, 0.95, 1.0]) if m.any() else None
    print(f"{name}: wp {wp.mean():.3f} status_eq {(r.status == o.status).mean():.3f} iter_eq {same_it[opt].mean():.3f} act_eq {same_act[opt].mean():.3f} "
          f"same {same.sum()}/{opt.sum()} flags {np.bincount(r.flags.astype(int), minlength=5)[:5]}")
    print(f"    ex(same) {q(ex, same)} ey(same) {q(ey, same)}  ex(opt) {q(ex, opt)} ey(opt) {q(ey, opt)}")

cm = sfb.to_colmajor
for (B, n, m, seed, pol) in [(256, 10, 20, 41, True), (256, 50, 100, 42, True), (256, 3, 203, 43, False), (256, 3, 203, 43, True), (128, 33, 31, 44, True)]:
    P, q, A, l, u = (np.asarray(t, np.float32).astype(np.float64) for t in random_qp_numpy(B, n, m, seed=seed))
    prm = sfb.QPSolverParams(max_iter=4000, polish=pol)
    c = lambda t: np.ascontiguousarray(t, dtype=np.float32)
    r = sfb.solve_dense_batch(c(cm(P)), c(q), c(cm(A)), c(l), c(u), prm)
    op = orc.default_params(max_iter=4000, polish=int(pol))
    o = orc.qp_solve_batch(P, q, A, l, u, params=op, nthreads=8)
    o2 = orc.qp_solve_batch(P, q, A, l, u, params=op, nthreads=8, fast=True)
    report(f"dense n={n} m={m} polish={pol}", r, o, o2)

# real vehicle workloads
from oracle import transcribe as tr
from smooth_feedback_b200.generators import sparse_to_dense
mpc = tr.vehicle_mpc()
B = int(os.environ.get("DIAG_MPC_B", "64"))
t0, x0 = tr.sample_vehicle_states(B, seed=5)
Pv, qs, Av, ls, us = [], [], [], [], []
for b in range(B):
    qp = mpc.transcribe(t0[b], x0[b])
    rp, ci, av = qp.csr_A(); cp, ri, pv = qp.csc_P()
    Pv.append(pv); Av.append(av); qs.append(qp.q.copy()); ls.append(qp.l.copy()); us.append(qp.u.copy())
pat = dict(n=qp.n, m=qp.m, P_colptr=cp, P_rowidx=ri, A_rowptr=rp, A_colidx=ci)
Pv, qs, Av, ls, us = (np.stack(t) for t in (Pv, qs, Av, ls, us))
sp = sfb.SparsePattern(pat["n"], pat["m"], cp, ri, rp, ci)
print("vehicle MPC pattern nnzL", sp.nnzL)
Pd, Ad = sparse_to_dense(pat, Pv, Av)
for dt, pol in [(np.float64, True), (np.float32, True), (np.float32, False)]:
    f = lambda t: np.ascontiguousarray(np.asarray(t, dt))
    r = sfb.solve_sparse_batch(sp, f(Pv), f(qs), f(Av), f(ls), f(us), sfb.QPSolverParams(max_iter=4000, polish=pol))
    g = lambda t: np.asarray(t, dt).astype(np.float64)
    op = orc.default_params(max_iter=4000, polish=int(pol))
    o = orc.qp_solve_batch(g(Pd), g(qs), g(Ad), g(ls), g(us), params=op, nthreads=8)
    o2 = orc.qp_solve_batch(g(Pd), g(qs), g(Ad), g(ls), g(us), params=op, nthreads=8, fast=True)
    report(f"vehicle MPC sparse {np.dtype(dt).name} polish={pol}", r, o, o2)
    print("    iters", np.bincount(o.iter)[np.bincount(o.iter) > 0], np.unique(o.iter), "gpu", np.unique(r.iter))
