"""Phase cost estimate for the dense QP kernel at cfg2 via parameter variations (dev tool, GPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import random_qp_torch
B, n, m = 65536, 50, 100
if len(sys.argv) > 3: B, n, m = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
P_cm, q, A_cm, l, u = random_qp_torch(B, n, m, seed=5)
def t(**kw):
    prm = sfb.QPSolverParams(**kw)
    out = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, prm)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); out = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, prm, out=out); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, out.iter.double().mean().item()
full, it = t(max_iter=4000)
nopol, _ = t(max_iter=4000, polish=False)
setup, _ = t(max_iter=1, polish=False)
it51, _ = t(max_iter=51, polish=False, stop_check_iter=1000)   # 50 more iterations, no checks
noscale, _ = t(max_iter=1, polish=False, scaling=False)
print(f"B={B} n={n} m={m}: full {full:.2f} ms ({B/full*1e3:.3e}/s, mean iter {it:.1f}) | no polish {nopol:.2f} | setup+1it {setup:.2f} (no scaling {noscale:.2f}) | 51 it no checks {it51:.2f}"
      f" -> per-iteration {(it51-setup)/50*1e3:.1f} us/batch, polish {full-nopol:.2f}, checks+rest {nopol-setup-(it51-setup)/50*(it-1):.2f}")
