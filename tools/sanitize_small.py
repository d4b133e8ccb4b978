"""Dev tool: tiny invocations of every kernel family for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import (mpc_structured_batch, mpc_structured_pattern, random_ekf_numpy, random_qp_numpy,
                                             random_sparse_qp_numpy)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
cm = sfb.to_colmajor
prm = sfb.QPSolverParams(max_iter=200)
for (n, m, B) in [(50, 100, 24), (10, 20, 40), (3, 203, 16), (2, 2, 8)]:
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=1)
    r = sfb.solve_dense_batch(t(cm(P)), t(q), t(cm(A)), t(l), t(u), prm)
for (n, m, B) in [(3, 203, 40), (4, 256, 9), (1, 1, 5), (2, 33, 17)]:  # tall-skinny register kernel (polish off), fp64 + fp32
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=2)
    prm_np = sfb.QPSolverParams(max_iter=200, polish=False)
    sfb.solve_dense_batch(t(cm(P)), t(q), t(cm(A)), t(l), t(u), prm_np)
    f = lambda a: t(a).float()
    sfb.solve_dense_batch(f(cm(P)), f(q), f(cm(A)), f(l), f(u), prm_np)
Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(200, 6, 3, seed=1)
sfb.ekf_step_batch(t(cm(Pk)), t(cm(Ak)), t(cm(Qk)), 0.1, t(cm(Hk)), t(cm(Rk)), t(innov))
sfb.ekf_predict_batch(t(cm(Pk)), t(cm(Ak)), t(cm(Qk)), 0.1, stepper="rk4", dt=0.05)
pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
Pv, q, Av, l, u = mpc_structured_batch(pat, 13, seed=2)
for tw in (4, 8, 32):
    os.environ["SFB_SPARSE_TW"] = str(tw)
    h = sfb.Handle(0)
    sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=h)
    sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm)
    f = lambda a: t(a).float()
    sfb.solve_sparse_batch(sp, f(Pv), f(q), f(Av), f(l), f(u), prm)
os.environ.pop("SFB_SPARSE_TW", None)
# the on-chip sparse kernel (default handle): MPC structure (several supernode levels, fp64 + fp32 with its fp64 polish pass),
# warm start, a random pattern, no constraints at all
h = sfb.Handle(0)
sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=h)
assert sp.uses_onchip(8)[0] and sp.uses_onchip(4)[0]
r0 = sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm)
sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm, warm_x=r0.x, warm_y=r0.y)
f = lambda a: t(a).float()
sfb.solve_sparse_batch(sp, f(Pv), f(q), f(Av), f(l), f(u), prm)
pat2, Pv, q, Av, l, u = random_sparse_qp_numpy(9, 20, 30, density=0.2, seed=3)
sp = sfb.SparsePattern(pat2["n"], pat2["m"], pat2["P_colptr"], pat2["P_rowidx"], pat2["A_rowptr"], pat2["A_colidx"], handle=h)
sfb.solve_sparse_batch(sp, t(Pv), t(q), t(Av), t(l), t(u), prm)
# round 2: fp32 + polish (mixed-precision second pass), the ASIF / MPC fleets (transcription, fused filter, epilogue) and the
# library's own all-gather (world size 1)
from smooth_feedback_b200.generators import vehicle_fleet_numpy
P, q, A, l, u = random_qp_numpy(24, 10, 20, seed=4)
f = lambda a: t(a).float()
sfb.solve_dense_batch(f(cm(P)), f(q), f(cm(A)), f(l), f(u), sfb.QPSolverParams(max_iter=200))
t0, x0, ud = vehicle_fleet_numpy(70, seed=3)
for dt in (np.float64, np.float32):
    fl = sfb.ASIFVehicleFleet(70, sfb.ASIFVehicleParams(qp=sfb.QPSolverParams(polish=False, max_iter=120)), dtype=dt)
    fl(x0.astype(dt), ud.astype(dt)); fl(x0.astype(dt), ud.astype(dt))
    if dt == np.float64:
        fl.to_qp(x0, ud)
    fl.close()
    mf = sfb.MPCVehicleFleet(9, sfb.MPCVehicleParams(K=10, tf=2.0, qp=sfb.QPSolverParams(max_iter=120)), dtype=dt)
    mf(t0[:9].astype(dt), x0[:9].astype(dt)); mf(t0[:9].astype(dt), x0[:9].astype(dt))
    mf.close()
hc = sfb.Handle(0)
comm = sfb.Communicator(hc, 1, 0, sfb.Communicator.unique_id())
r = sfb.solve_dense_batch(t(cm(P)), t(q), t(cm(A)), t(l), t(u), prm, handle=hc)
comm.all_gather([r.x, r.y, r.status]); comm.wait(host=True); comm.close()
torch.cuda.synchronize()
print("sanitize workload done")
