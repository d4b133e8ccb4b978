#!/usr/bin/env python
"""Fleet benchmarks on one GPU: the on-device ASIFilter (BASELINE.json configs[4]: SE(2) vehicle, 1-dim barrier, n = 3,
m = 203, batch 32768, fp32) and the on-device MPC (configs[2]: K = 50 -> n = m = 422 sparse, 8192 agents, fp32), on the REAL
workloads (states sampled around the desired trajectory) instead of random QPs of the same shape.

    python tools/bench_fleet.py asif|mpc [--batch N] [--dtype f32|f64] [--steps K] [--closed-loop S]

Each prints one JSON line:
  cold      every agent solves from scratch (fresh ASIFilter / MPC objects): agent-steps/s, device-resident (t, x, u_des)
  closed_loop  S control steps of the whole fleet THROUGH HOST BUFFERS (pinned): per step the states go H2D, the inputs come
            back D2H, the plant is integrated on the host (not timed); warm starts stay resident on the device
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _events(fn, steps, warmup=2):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return {"ms_min": ms[0], "ms_median": ms[len(ms) // 2], "ms_all": ms}


def run_asif(batch=32768, dtype="f32", steps=5, closed_loop=50, device=0):
    import numpy as np
    import torch

    import smooth_feedback_b200 as sfb
    from smooth_feedback_b200.generators import vehicle_dynamics_numpy, vehicle_fleet_numpy, vehicle_rplus_numpy

    npdt = np.float32 if dtype == "f32" else np.float64
    tdt = torch.float32 if dtype == "f32" else torch.float64
    dev = torch.device("cuda", device)
    t0, x0, ud = vehicle_fleet_numpy(batch, seed=5)
    handle = sfb.Handle(device)
    prm = sfb.ASIFVehicleParams(qp=sfb.QPSolverParams(polish=False, max_iter=4000))
    fleet = sfb.ASIFVehicleFleet(batch, prm, dtype=npdt, handle=handle)
    xd = torch.from_numpy(x0).to(dev, dtype=tdt).contiguous()
    udd = torch.from_numpy(ud).to(dev, dtype=tdt).contiguous()
    res = {"workload": f"ASIFilter vehicle fleet (mpc_asif_vehicle.cpp:96-129: K=200, nh=1 -> n=3 m=203, polish off) batch={batch} {dtype}, "
                       f"states x0 = xdes(t0) (+) N(0, 0.1^2), u_des ~ U(-0.5, 0.5)^2"}
    out = {}

    def cold():
        fleet.reset_warmstart()
        out["r"] = fleet(xd, udd)

    c = _events(cold, steps)
    u, st, it = out["r"]
    res["cold"] = dict(c, steps_per_s=batch / (c["ms_median"] * 1e-3), mean_iter=float(it.double().mean().item()),
                       status_hist=torch.bincount(st, minlength=7).tolist(), filtered_frac=float(((u - udd).abs().amax(dim=1) > 1e-3).double().mean().item()))
    w = _events(lambda: fleet(xd, udd), steps)  # same problem again from the resident warm start
    res["warm_same_state"] = dict(w, steps_per_s=batch / (w["ms_median"] * 1e-3))
    # closed loop through host buffers: x (7), u_des (2) in; u (2), status, iter out, per agent and step
    if closed_loop:
        fleet.reset_warmstart()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=npdt)).pin_memory().numpy()
        x, udh = x0.copy(), pin(ud)
        spent, iters, nonopt = 0.0, [], 0
        for k in range(closed_loop):
            xh = pin(x)
            t1 = time.perf_counter()
            u, st, it = fleet(xh, udh)  # returns with the results on the host
            spent += time.perf_counter() - t1
            iters.append(float(it.mean())); nonopt += int((st != 0).sum())
            x = vehicle_rplus_numpy(x, 0.025 * vehicle_dynamics_numpy(x, u.astype(np.float64)))  # plant: explicit Euler, 25 ms
        per = 9 * npdt().itemsize, 2 * npdt().itemsize + 8
        res["closed_loop"] = {"control_steps": closed_loop, "agent_steps_per_s": batch * closed_loop / spent, "ms_per_step": 1e3 * spent / closed_loop,
                              "h2d_bytes_per_agent_step": per[0], "d2h_bytes_per_agent_step": per[1], "mean_iter_first": iters[0],
                              "mean_iter_last": iters[-1], "non_optimal": nonopt}
    res["gpu_launches"] = handle.launch_count()
    return res


def run_mpc(batch=8192, dtype="f32", steps=3, closed_loop=50, device=0):
    import numpy as np
    import torch

    import smooth_feedback_b200 as sfb
    from smooth_feedback_b200.generators import vehicle_dynamics_numpy, vehicle_fleet_numpy, vehicle_rplus_numpy

    npdt = np.float32 if dtype == "f32" else np.float64
    tdt = torch.float32 if dtype == "f32" else torch.float64
    dev = torch.device("cuda", device)
    t0, x0, _ = vehicle_fleet_numpy(batch, seed=5)
    handle = sfb.Handle(device)
    prm = sfb.MPCVehicleParams(qp=sfb.QPSolverParams(max_iter=4000))
    tc = time.perf_counter()
    fleet = sfb.MPCVehicleFleet(batch, prm, dtype=npdt, handle=handle)
    t_create = time.perf_counter() - tc
    td = torch.from_numpy(t0).to(dev, dtype=tdt).contiguous()
    xd = torch.from_numpy(x0).to(dev, dtype=tdt).contiguous()
    res = {"workload": f"MPC vehicle fleet (mpc_asif_vehicle.cpp:42-89 at K=50: n={fleet.n} m={fleet.m} nnzA={fleet.nnzA} nnzP={fleet.nnzP} "
                       f"nnzL={fleet.nnzL}, polish on) batch={batch} {dtype}, x0 = xdes(t0) (+) N(0, 0.1^2), t0 ~ U(0, 30)",
           "create_s": t_create}
    out = {}

    def cold():
        fleet.reset_warmstart()
        out["r"] = fleet(td, xd)

    c = _events(cold, steps, warmup=1)
    u, st, it = out["r"]
    res["cold"] = dict(c, steps_per_s=batch / (c["ms_median"] * 1e-3), mean_iter=float(it.double().mean().item()),
                       status_hist=torch.bincount(st, minlength=7).tolist())
    if closed_loop:
        fleet.reset_warmstart()
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=npdt)).pin_memory().numpy()
        t, x = t0.copy(), x0.copy()
        spent, iters, nonopt = 0.0, [], 0
        for k in range(closed_loop):
            th, xh = pin(t), pin(x)
            t1 = time.perf_counter()
            u, st, it = fleet(th, xh)
            spent += time.perf_counter() - t1
            iters.append(float(it.mean())); nonopt += int((st != 0).sum())
            x = vehicle_rplus_numpy(x, 0.025 * vehicle_dynamics_numpy(x, u.astype(np.float64)))
            t = t + 0.025
        res["closed_loop"] = {"control_steps": closed_loop, "agent_steps_per_s": batch * closed_loop / spent, "ms_per_step": 1e3 * spent / closed_loop,
                              "h2d_bytes_per_agent_step": 8 * npdt().itemsize, "d2h_bytes_per_agent_step": 2 * npdt().itemsize + 8,
                              "mean_iter_first": iters[0], "mean_iter_last": iters[-1], "non_optimal": nonopt,
                              "tracking_error_final": float(np.abs(x[:, 4:7] - np.array([1.0, 0.0, 0.4])).max())}
    res["gpu_launches"] = handle.launch_count()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("which", choices=["asif", "mpc"])
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--dtype", default="f32", choices=["f64", "f32"])
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--closed-loop", type=int, default=50)
    a = ap.parse_args()
    if a.which == "asif":
        print(json.dumps(run_asif(a.batch or 32768, a.dtype, a.steps, a.closed_loop)), flush=True)
    else:
        print(json.dumps(run_mpc(a.batch or 8192, a.dtype, a.steps, a.closed_loop)), flush=True)
