"""Tiny driver for ncu: the tall-skinny dense QP kernel at BASELINE configs[4] shape (n=3, m=203, fp32, batch 32768)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import random_qp_torch
P, q, A, l, u = random_qp_torch(32768, 3, 203, seed=5, device="cuda", dtype=torch.float32)
prm = sfb.QPSolverParams(max_iter=4000, polish=False)
r = None
for _ in range(3):
    r = sfb.solve_dense_batch(P, q, A, l, u, prm, out=r)
torch.cuda.synchronize()
print("mean iter", r.iter.double().mean().item(), "optimal", (r.status == 0).double().mean().item())
