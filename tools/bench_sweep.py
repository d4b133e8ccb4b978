#!/usr/bin/env python
"""The reference's own benchmark protocol (benchmarks/bench.cpp:140-221) on the engine: sizes x densities, m = n, the solver
settings of bench.cpp:144-153 (eps_abs = eps_rel = 1e-6, polish on, max_iter 10000, scaling OFF), generator
bench_types.hpp:19-41 (here with delta ~ U(0,1) so that instances are feasible; the literal recipe makes about half of them
infeasible at m = n too).  Per (size, density): average time per solve over the instances that end Optimal, for
  dense   sfb_qp_solve_dense_batch_f64   (every density: a dense kernel does not care; sizes that fit in shared memory)
  sparse  sfb_qp_solve_sparse_batch_f64  (density < 1, ONE Bernoulli mask per size shared by the batch -- the engine's sparse
          path needs a shared pattern; the reference draws a new mask per instance)
so that the numbers can be laid over media/qp_benchmarks.png (smooth-dense ~0.07 s at n = m = 500, <= 0.01 s at 100).

    python tools/bench_sweep.py [--batch 256] [--sizes 20,40,60,80,100] [--sparse-sizes 100,200,300]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch

    import smooth_feedback_b200 as sfb
    from smooth_feedback_b200.generators import random_qp_numpy, random_sparse_qp_numpy

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--sizes", default="20,40,60,80,100")
    ap.add_argument("--sparse-sizes", default="100,200,300")
    ap.add_argument("--densities", default="0.05,0.3")
    a = ap.parse_args()
    prm = sfb.QPSolverParams(eps_abs=1e-6, eps_rel=1e-6, polish=True, max_iter=10000, scaling=False)  # bench.cpp:144-153
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    cm = sfb.to_colmajor

    def timed(fn):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); e1.synchronize()
        return r, e0.elapsed_time(e1) * 1e-3

    rows = []
    for n in [int(s) for s in a.sizes.split(",") if s]:
        P, q, A, l, u = random_qp_numpy(a.batch, n, n, seed=5)
        args = (t(cm(P)), t(q), t(cm(A)), t(l), t(u))
        try:
            r, dt = timed(lambda: sfb.solve_dense_batch(*args, prm))
            st = r.status.cpu().numpy()
            rows.append({"solver": "dense", "n": n, "m": n, "density": 1.0, "batch": a.batch, "s_per_solve_batched": dt / a.batch,
                         "optimal_frac": float((st == 0).mean()), "mean_iter": float(r.iter.double().mean().item()),
                         "status_hist": np.bincount(st, minlength=7).tolist()})
        except sfb.SfbError as e:
            rows.append({"solver": "dense", "n": n, "m": n, "error": str(e)})
    for dens in [float(s) for s in a.densities.split(",") if s]:
        for n in [int(s) for s in a.sparse_sizes.split(",") if s]:
            pat, Pv, q, Av, l, u = random_sparse_qp_numpy(a.batch, n, n, density=dens, seed=5)
            sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"])
            args = (t(Pv), t(q), t(Av), t(l), t(u))
            r, dt = timed(lambda: sfb.solve_sparse_batch(sp, *args, prm))
            st = r.status.cpu().numpy()
            rows.append({"solver": "sparse", "n": n, "m": n, "density": dens, "batch": a.batch, "nnzA": sp.nnzA, "nnzL": sp.nnzL,
                         "s_per_solve_batched": dt / a.batch, "optimal_frac": float((st == 0).mean()),
                         "mean_iter": float(r.iter.double().mean().item()), "status_hist": np.bincount(st, minlength=7).tolist()})
    print(json.dumps({"protocol": "benchmarks/bench.cpp:140-221 (eps 1e-6, polish, max_iter 1e4, scaling off; m = n)", "rows": rows}), flush=True)


if __name__ == "__main__":
    main()
