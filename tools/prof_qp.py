"""Tiny driver for ncu: a few launches of the dense QP kernel at cfg2 shape (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import random_qp_torch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, m = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (50, 100)
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
polish = (sys.argv[5] != "nopolish") if len(sys.argv) > 5 else True
P_cm, q, A_cm, l, u = random_qp_torch(B, n, m, seed=5)
prm = sfb.QPSolverParams(max_iter=4000, polish=polish)
out = None
for _ in range(reps):
    out = sfb.solve_dense_batch(P_cm, q, A_cm, l, u, prm, out=out)
torch.cuda.synchronize()
print("mean iter", out.iter.double().mean().item(), "optimal", (out.status == 0).double().mean().item())
