"""Dev tool: a handful of headline-shape dense QPs (n = 50, m = 100: the register-resident ADMM loop and the register-blocked
Gauss-Jordan inverse) for `compute-sanitizer --tool racecheck`; 60 iterations so that the stop check, the warp-synchronised
end of an iteration and the polish all run."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import random_qp_numpy

t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
P, q, A, l, u = random_qp_numpy(6, 50, 100, seed=1)
r = sfb.solve_dense_batch(t(sfb.to_colmajor(P)), t(q), t(sfb.to_colmajor(A)), t(l), t(u), sfb.QPSolverParams(max_iter=60))
torch.cuda.synchronize()
print("iter", r.iter.tolist(), "status", r.status.tolist())
