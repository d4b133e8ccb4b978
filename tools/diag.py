"""GPU diagnostic: parity statistics vs the oracle and quick timings. Run on the GPU box (dev tool)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import smooth_feedback_b200 as sfb
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy, random_qp_torch

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
cm = sfb.to_colmajor
print(torch.cuda.get_device_name(0), "cpu cores", os.cpu_count(), flush=True)

def parity(B, n, m, seed, feasible=True, **kw):
    P, q, A, l, u = random_qp_numpy(B, n, m, seed=seed, feasible=feasible)
    prm = sfb.QPSolverParams(max_iter=4000, **kw)
    r = sfb.solve_dense_batch(t(cm(P)), t(q), t(cm(A)), t(l), t(u), prm)
    torch.cuda.synchronize()
    okw = {k: (int(v) if isinstance(v, bool) else v) for k, v in kw.items()}
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000, **okw), nthreads=os.cpu_count())
    st = r.status.cpu().numpy(); it = r.iter.cpu().numpy().astype(np.uint32); act = r.active.cpu().numpy()
    x = r.x.cpu().numpy(); y = r.y.cpu().numpy(); fl = r.flags.cpu().numpy()
    ok = (o.status == 0) & (st == 0)
    ex = np.linalg.norm(x - o.x, axis=1) / np.maximum(np.linalg.norm(o.x, axis=1), 1e-9)
    ey = np.linalg.norm(y - o.y, axis=1) / np.maximum(np.linalg.norm(o.y, axis=1), 1e-9)
    print(f"[parity B={B} n={n} m={m} feas={feasible} {kw}] status_mismatch={(st != o.status).sum()} iter_mismatch={(it != o.iter).sum()} "
          f"active_mismatch={(act != o.active).any(1).sum()} relx_max={ex[ok].max() if ok.any() else -1:.3e} rely_max={ey[ok].max() if ok.any() else -1:.3e} "
          f"status_hist={np.bincount(o.status, minlength=7)} gpu_hist={np.bincount(st, minlength=7)} flags_hist={np.bincount(fl, minlength=8)} "
          f"iter_mean={o.iter.mean():.1f} na_mean={(o.active != 0).sum(1).mean():.1f}", flush=True)
    if (st != o.status).any() or (it != o.iter).any():
        bad = np.nonzero((st != o.status) | (it != o.iter))[0][:5]
        for b in bad:
            print("   inst", b, "gpu", st[b], it[b], "oracle", o.status[b], o.iter[b], "relx", ex[b])

for args in [(64, 2, 2, 1), (256, 10, 20, 5), (256, 50, 100, 5), (64, 3, 203, 5), (64, 7, 13, 5)]:
    try:
        parity(*args)
    except Exception as e:
        print("parity", args, "FAILED:", repr(e), flush=True)
try:
    parity(256, 10, 20, 11, feasible=False)
    parity(128, 10, 20, 7, eps_abs=1e-6, eps_rel=1e-6, polish=False)
except Exception as e:
    print("FAILED:", repr(e), flush=True)

def timeit(B, n, m, dtype=torch.float64, reps=3, **kw):
    P_cm, q, A_cm, l, u = random_qp_torch(B, n, m, seed=5, device=dev, dtype=dtype)
    prm = sfb.QPSolverParams(max_iter=4000, **kw)
    out = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, prm)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, prm, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    it = out.iter.double().mean().item()
    print(f"[time B={B} n={n} m={m} {dtype} {kw}] {best:.2f} ms -> {B / best * 1e3:.3e} solves/s  mean_iter={it:.1f} optimal={(out.status == 0).double().mean().item():.4f}", flush=True)

for args in [(16384, 50, 100), (65536, 50, 100), (65536, 10, 20), (32768, 3, 203)]:
    try:
        timeit(*args)
    except Exception as e:
        print("time", args, "FAILED:", repr(e), flush=True)
try:
    timeit(65536, 50, 100, polish=False)
    timeit(65536, 50, 100, max_iter=None) if False else None
    timeit(65536, 50, 100, dtype=torch.float32)
except Exception as e:
    print("FAILED:", repr(e), flush=True)

# ---- full-size property statistics (what fraction of polished solutions violate inactive rows?) ----
try:
    B, n, m = 65536, 50, 100
    P_cm, q, A_cm, l, u = random_qp_torch(B, n, m, seed=5)
    r = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, sfb.QPSolverParams(max_iter=4000))
    torch.cuda.synchronize()
    A = A_cm.transpose(1, 2)
    viol = (torch.einsum("bij,bj->bi", A, r.x) - u).clamp(min=0).max(dim=1).values
    stat = (torch.einsum("bij,bj->bi", P_cm, r.x) + q + torch.einsum("bij,bi->bj", A, r.y)).abs().max(dim=1).values
    print(f"[fullsize] viol>1e-7: {(viol > 1e-7).sum().item()} max {viol.max().item():.3e}; stat>1e-6: {(stat > 1e-6).sum().item()} max {stat.max().item():.3e}; "
          f"miny {r.y.min().item():.3e}; flags {torch.bincount(r.flags, minlength=8).tolist()}; na max {(r.active != 0).sum(1).max().item()}", flush=True)
    bad = torch.nonzero(viol > 1e-7).flatten()[:64]
    if len(bad):
        cpu = lambda t_: t_.cpu().numpy()
        Pb = np.swapaxes(cpu(P_cm[bad]), 1, 2); Ab = np.swapaxes(cpu(A_cm[bad]), 1, 2)
        o = orc.qp_solve_batch(Pb, cpu(q[bad]), Ab, cpu(l[bad]), cpu(u[bad]), params=orc.default_params(max_iter=4000), nthreads=os.cpu_count())
        ov = np.clip(np.einsum("bij,bj->bi", Ab, o.x) - cpu(u[bad]), 0, None).max(1)
        ex = np.linalg.norm(cpu(r.x[bad]) - o.x, axis=1) / np.linalg.norm(o.x, axis=1)
        print(f"   oracle on the {len(bad)} violating instances: oracle viol max {ov.max():.3e} (>1e-7: {(ov > 1e-7).sum()}), relx vs gpu max {ex.max():.3e}, "
              f"iter equal {(cpu(r.iter[bad]).astype(np.uint32) == o.iter).all()}, active equal {(cpu(r.active[bad]) == o.active).all()}", flush=True)
except Exception as e:
    import traceback; traceback.print_exc()

# ---- EKF large batch vs oracle on slices ----
try:
    from smooth_feedback_b200.generators import random_ekf_numpy
    B, d, ny = 200000, 6, 3
    Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(B, d, ny, seed=5)
    Pp = sfb.ekf_predict_batch(t(cm(Pk)), t(cm(Ak)), t(cm(Qk)), 0.1)
    delta, Pu = sfb.ekf_update_batch(Pp, t(cm(Hk)), t(cm(Rk)), t(innov))
    torch.cuda.synchronize()
    oPp = orc.ekf_predict_batch(Pk, Ak, Qk, 0.1, nthreads=os.cpu_count())
    od, oPu = orc.ekf_update_batch(oPp, Hk, Rk, innov, nthreads=os.cpu_count())
    eP = np.abs(np.swapaxes(Pp.cpu().numpy(), 1, 2) - oPp).reshape(B, -1).max(1)
    eU = np.abs(np.swapaxes(Pu.cpu().numpy(), 1, 2) - oPu).reshape(B, -1).max(1)
    eD = np.abs(delta.cpu().numpy() - od).max(1)
    print(f"[ekf B={B}] predict err max {eP.max():.3e} update err max {eU.max():.3e} (argmax {eU.argmax()}) delta err max {eD.max():.3e}; n bad(>1e-9) {(eU > 1e-9).sum()}", flush=True)
    # same check against the textbook formula in float64 numpy
    S = Hk @ oPp @ np.swapaxes(Hk, 1, 2) + Rk
    K = oPp @ np.swapaxes(Hk, 1, 2) @ np.linalg.inv(S)
    tb = (np.eye(d) - K @ Hk) @ oPp
    eT = np.abs(oPu - tb).reshape(B, -1).max(1)
    print(f"   oracle vs textbook err max {eT.max():.3e}, cond(S) max {np.linalg.cond(S).max():.3e}", flush=True)
except Exception as e:
    import traceback; traceback.print_exc()
