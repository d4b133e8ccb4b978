#!/usr/bin/env python
"""Regenerates the committed golden fixtures of tests/golden/.

The reference (pettni/smooth_feedback) cannot be built in this image (Eigen / Boost / smooth absent), so the fixtures are
produced by the CPU oracle (oracle/, a restatement pinned by the reference's own known-answer tests) on seeded inputs:

    python tests/golden/make_golden.py

  qp_known_answers.json   the reference's known-answer cases (tests/test_qp.cpp:54-336) with the oracle's status / iteration
                          count / solution next to the expected values of the reference test
  qp_seeded.npz           seeded batches (dense n=10,m=20 and n=50,m=100, tall-skinny n=3,m=203 without polish, MPC-structured
                          sparse n=m=63): inputs are regenerated from the seed by the tests, outputs are stored here
  ekf_seeded.npz          seeded EKF predict + update (d=6, ny=3)
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as orc  # noqa: E402
from qp_cases import CASES, as_batch  # noqa: E402
from smooth_feedback_b200.generators import (mpc_structured_batch, mpc_structured_pattern, random_ekf_numpy,  # noqa: E402
                                             random_qp_numpy, sparse_to_dense)

SPECS = {  # name -> (kind, args, params)
    "dense_n10_m20": ("dense", dict(B=64, n=10, m=20, seed=101), dict(max_iter=4000)),
    "dense_n50_m100": ("dense", dict(B=16, n=50, m=100, seed=102), dict(max_iter=4000)),
    "skinny_n3_m203_nopolish": ("dense", dict(B=32, n=3, m=203, seed=103), dict(max_iter=4000, polish=0)),
    "sparse_mpc_n63": ("mpc", dict(B=16, seed=104), dict(max_iter=4000)),
}


def solve_spec(name):
    kind, args, prm = SPECS[name]
    if kind == "dense":
        P, q, A, l, u = random_qp_numpy(args["B"], args["n"], args["m"], seed=args["seed"])
    else:
        pat = mpc_structured_pattern(Nx=3, Nu=2, nivals=3, Ki=4)
        Pv, q, Av, l, u = mpc_structured_batch(pat, args["B"], seed=args["seed"])
        P, A = sparse_to_dense(pat, Pv, Av)
    return orc.qp_solve_batch(P, q, A, l, u, params=orc.default_params(**prm), nthreads=4)


def main():
    orc.build()
    known = []
    for c in CASES:
        o = orc.qp_solve_batch(*as_batch(c))
        known.append(dict(name=c["name"], expected_status=int(c["status"]), expected_x=None if c["x"] is None else list(map(float, c["x"])),
                          oracle_status=int(o.status[0]), oracle_iter=int(o.iter[0]), oracle_x=o.x[0].tolist(), oracle_y=o.y[0].tolist(),
                          oracle_obj=float(o.obj[0]), oracle_active=o.active[0].tolist()))
    json.dump(known, open(os.path.join(HERE, "qp_known_answers.json"), "w"), indent=1)
    out = {}
    for name in SPECS:
        o = solve_spec(name)
        out[name + "/x"] = o.x; out[name + "/y"] = o.y; out[name + "/obj"] = o.obj
        out[name + "/status"] = o.status; out[name + "/iter"] = o.iter; out[name + "/active"] = o.active
    np.savez_compressed(os.path.join(HERE, "qp_seeded.npz"), **out)
    Pk, Ak, Qk, Hk, Rk, innov = random_ekf_numpy(128, 6, 3, seed=105)
    Pp = orc.ekf_predict_batch(Pk, Ak, Qk, 0.1)
    d, Pu = orc.ekf_update_batch(Pp, Hk, Rk, innov)
    np.savez_compressed(os.path.join(HERE, "ekf_seeded.npz"), Pp=Pp, delta=d, Pu=Pu)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
