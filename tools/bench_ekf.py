#!/usr/bin/env python
"""EKF covariance cycle throughput (BASELINE.json configs[3] shape: d=6, ny=3, batch 2^20, fp64) on one GPU.

    python tools/bench_ekf.py [--batch N] [--d 6] [--ny 3] [--steps 20]

Prints one JSON line: fused TMA kernel (sfb_ekf_step_batch_f64) vs the two generic kernels back to back, each as
filter cycles/s and as achieved HBM GB/s over the algorithmic bytes of the step
    in  8*(3 d^2 + ny d + ny^2 + ny)   P, A, Q, H, R, innov
    out 8*(d^2 + d)                    P', delta
(SURVEY 8(d)'s B_ekf minus the group element and f, which stay with the caller).  Inputs exceed L2 (1.16 GB at 2^20).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def ekf_bytes(d, ny):
    return 8 * (3 * d * d + ny * d + ny * ny + ny), 8 * (d * d + d)


def run(batch, d, ny, steps, warmup=3):
    import torch

    import smooth_feedback_b200 as sfb

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    M = torch.rand(batch, d, d, generator=g, device=dev, dtype=torch.float64) * 2 - 1
    P = M @ M.transpose(1, 2) + 0.1 * torch.eye(d, device=dev, dtype=torch.float64)
    P = (0.5 * (P + P.transpose(1, 2))).contiguous()
    A = torch.randn(batch, d, d, generator=g, device=dev, dtype=torch.float64)
    Q = (0.01 * torch.eye(d, device=dev, dtype=torch.float64)).expand(batch, d, d).contiguous()
    H = torch.randn(batch, d, ny, generator=g, device=dev, dtype=torch.float64)
    R = (0.01 * torch.eye(ny, device=dev, dtype=torch.float64)).expand(batch, ny, ny).contiguous()
    innov = torch.randn(batch, ny, generator=g, device=dev, dtype=torch.float64)
    outP = torch.empty_like(P)
    outd = torch.empty(batch, d, device=dev, dtype=torch.float64)
    tmpP = torch.empty_like(P)

    h_fused = sfb.Handle(0)
    os.environ["SFB_EKF_FORCE_GENERIC"] = "1"
    h_gen = sfb.Handle(0)
    os.environ.pop("SFB_EKF_FORCE_GENERIC")

    def fused():
        sfb.ekf_step_batch(P, A, Q, 0.1, H, R, innov, handle=h_fused, out_delta=outd, out_P=outP)

    def generic():
        sfb.ekf_predict_batch(P, A, Q, 0.1, handle=h_gen, out=tmpP)
        sfb.ekf_update_batch(tmpP, H, R, innov, handle=h_gen, out_delta=outd, out_P=outP)

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        best, tot = 1e30, 0.0
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1)
            best = min(best, ms); tot += ms
        return best, tot / steps

    bi, bo = ekf_bytes(d, ny)
    peak = 6650.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    res = {"workload": f"EKF covariance predict(euler)+update, d={d} ny={ny} batch={batch} fp64", "bytes_in": bi, "bytes_out": bo,
           "hbm_peak_gbs": peak}
    fused(); ref_P, ref_d = outP.clone(), outd.clone()
    generic()
    res["fused_vs_generic_max_abs_diff"] = float(max((outP - ref_P).abs().max().item(), (outd - ref_d).abs().max().item()))
    for name, fn, nbytes in (("fused_tma", fused, bi + bo), ("generic_two_kernels", generic, bi + bo + 16 * d * d)):
        best, mean = timed(fn)
        res[name] = {"ms_best": best, "ms_mean": mean, "cycles_per_s": batch / (mean * 1e-3),
                     "achieved_gbs": batch * nbytes / (mean * 1e-3) / 1e9, "frac_hbm": batch * nbytes / (mean * 1e-3) / 1e9 / peak,
                     "bytes_per_cycle": nbytes}
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1 << 20)
    ap.add_argument("--d", type=int, default=6)
    ap.add_argument("--ny", type=int, default=3)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    print(json.dumps(run(a.batch, a.d, a.ny, a.steps)), flush=True)
