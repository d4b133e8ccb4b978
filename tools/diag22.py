"""Dev tool: dense-kernel polish parity over tiny shapes (which (n, m, na) combinations disagree with the oracle)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import smooth_feedback_b200 as sfb
from oracle import oracle as orc
from smooth_feedback_b200.generators import random_qp_numpy
from test_gpu_qp_parity import gpu_solve, rel_err
for (n, m) in [(2, 2), (2, 3), (3, 2), (3, 3), (4, 4), (5, 5), (6, 6), (8, 8), (4, 6), (6, 4), (10, 10), (16, 16), (20, 20)]:
    P, q, A, l, u = random_qp_numpy(128, n, m, seed=n * 1000 + m)
    r = gpu_solve(sfb, P, q, A, l, u, sfb.QPSolverParams(max_iter=4000))
    o = orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(max_iter=4000), nthreads=4)
    ok = (o.status == 0) & (r.status == 0) & (r.active == o.active).all(1)
    ex = rel_err(r.x, o.x); na = (o.active != 0).sum(1)
    bad = ok & (ex > 1e-6)
    ldA = m + ((2 - m % 4) + 4) % 4
    print(f"n={n} m={m} ldA={ldA}: bad {bad.sum()}/{ok.sum()}  na of bad {sorted(set(na[bad]))}  na present {sorted(set(na[ok]))}  (2na>ldA means S in global scratch)")
