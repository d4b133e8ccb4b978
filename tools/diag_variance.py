"""Dev diagnostic: run-to-run spread of the on-chip sparse kernel (fp64 / fp32) on the real vehicle MPC workload, every timed
repetition printed next to the SM clock / power / throttle reasons nvidia-smi reported while it ran.

  python tools/diag_variance.py [--reps 12] [--batch 8192]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=12)
    ap.add_argument("--batch", type=int, default=8192)
    a = ap.parse_args()
    import numpy as np
    import torch

    import smooth_feedback_b200 as sfb
    from smooth_feedback_b200.generators import vehicle_fleet_numpy

    rows = []
    q = "clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active"
    proc = subprocess.Popen(["nvidia-smi", "--id=0", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

    def pump():
        for line in proc.stdout:
            rows.append((time.monotonic(), line.strip()))

    threading.Thread(target=pump, daemon=True).start()

    nb = 512
    t0, x0, _ = vehicle_fleet_numpy(nb, seed=5)
    fl = sfb.MPCVehicleFleet(nb)
    pat = fl.pattern()
    Pv, qv, Av, l, u = fl.to_qp(t0, x0)
    fl.close()
    rep = (a.batch + nb - 1) // nb
    out_lines = []
    for dtype in ("f64", "f32", "f64"):
        dt = torch.float64 if dtype == "f64" else torch.float32
        t = lambda x: torch.from_numpy(np.tile(x, (rep, 1))[:a.batch]).to("cuda:0", dtype=dt).contiguous()
        dP, dq, dA, dl, du = t(Pv), t(qv), t(Av), t(l), t(u)
        handle = sfb.Handle(0)
        sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=handle)
        prm = sfb.QPSolverParams(max_iter=4000)
        out = None
        for _ in range(2):
            out = sfb.solve_sparse_batch(sp, dP, dq, dA, dl, du, prm, out=out)
        torch.cuda.synchronize()
        for r in range(a.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0 = time.monotonic()
            e0.record()
            th0 = time.monotonic()
            out = sfb.solve_sparse_batch(sp, dP, dq, dA, dl, du, prm, out=out)
            th1 = time.monotonic()
            e1.record()
            e1.synchronize()
            w1 = time.monotonic()
            smi = [s for ts, s in rows if w0 <= ts <= w1 + 0.02]
            out_lines.append({"dtype": dtype, "rep": r, "ms_events": round(e0.elapsed_time(e1), 2), "ms_wall": round((w1 - w0) * 1e3, 2),
                              "ms_host_call": round((th1 - th0) * 1e3, 3), "smi": smi[:3] + smi[-1:] if len(smi) > 4 else smi})
        del handle, sp
    proc.terminate()
    for ln in out_lines:
        print(json.dumps(ln))


if __name__ == "__main__":
    main()
