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
