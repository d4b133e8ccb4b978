"""Dev tool: on-chip sparse kernel (qp_sparse_cta.cuh) against the HBM-tiled kernel on the real vehicle MPC workload:
discrete outcomes, max solution difference, timing (CUDA events, median of 5)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import smooth_feedback_b200 as sfb
from smooth_feedback_b200.generators import vehicle_fleet_numpy

def handle(kind):
    os.environ["SFB_SPARSE_KERNEL"] = kind
    h = sfb.Handle(0)
    os.environ.pop("SFB_SPARSE_KERNEL", None)
    return h

def main(batch=8192, base=512):
    t0, x0, _ = vehicle_fleet_numpy(base, seed=5)
    fl = sfb.MPCVehicleFleet(base)
    pat = fl.pattern()
    Pv, q, Av, l, u = fl.to_qp(t0, x0)
    fl.close()
    rep = (batch + base - 1) // base
    res = {}
    for dt, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        t = lambda a: torch.from_numpy(np.tile(a, (rep, 1))[:batch]).to("cuda:0", dtype=dt).contiguous()
        d = [t(a) for a in (Pv, q, Av, l, u)]
        outs = {}
        for kind in ("tiled", "cta"):
            h = handle(kind)
            sp = sfb.SparsePattern(pat["n"], pat["m"], pat["P_colptr"], pat["P_rowidx"], pat["A_rowptr"], pat["A_colidx"], handle=h)
            prm = sfb.QPSolverParams(max_iter=4000)
            out = None
            for _ in range(2):
                out = sfb.solve_sparse_batch(sp, *d, prm, out=out)
            torch.cuda.synchronize()
            ms = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); out = sfb.solve_sparse_batch(sp, *d, prm, out=out); e1.record(); e1.synchronize()
                ms.append(e0.elapsed_time(e1))
            outs[kind] = out
            res[f"{name}_{kind}_ms"] = sorted(ms)[2]
            res[f"{name}_{kind}_solves_per_s"] = batch / (sorted(ms)[2] * 1e-3)
            res[f"{name}_{kind}_status"] = torch.bincount(out.status, minlength=7).tolist()
            res[f"{name}_{kind}_iter_mean"] = out.iter.double().mean().item()
            res[f"{name}_{kind}_polished"] = float((out.flags & 1).double().mean().item())
        a, b = outs["tiled"], outs["cta"]
        res[f"{name}_status_mismatch"] = int((a.status != b.status).sum().item())
        res[f"{name}_iter_mismatch"] = int((a.iter != b.iter).sum().item())
        same = (a.status == b.status) & (a.iter == b.iter)
        dx = (a.x - b.x).abs().amax(dim=1) / a.x.abs().amax(dim=1).clamp_min(1e-30)
        dy = (a.y - b.y).abs().amax(dim=1) / a.y.abs().amax(dim=1).clamp_min(1e-30)
        res[f"{name}_max_rel_dx"] = dx[same].max().item()
        res[f"{name}_max_rel_dy"] = dy[same].max().item()
        res[f"{name}_nan"] = int(torch.isnan(b.x).any(dim=1).sum().item())
    print(json.dumps(res, indent=1))

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8192)
