"""Aggregate ncu source-level samples of the QP kernel into algorithm phases by source-line ranges (dev tool)."""
import csv, io, subprocess, sys, re
rep = sys.argv[1]
src_file = "smooth_feedback_b200/csrc/qp_dense_group.cuh"
lines = open(src_file).read().split("\n")
# phase boundaries: find the line numbers of function starts
marks = [("helpers", 1)]
pat = [("load", "__device__ void load("), ("scale", "__device__ void scale("), ("gj_reg", "__device__ bool gj_invert_reg("),
       ("gj_misc", "__device__ __forceinline__ bool gj_invert_at("), ("colpass", "__device__ __forceinline__ void colpass("),
       ("colpass_skinny", "__device__ void colpass_skinny("), ("At_vec", "__device__ void At_vec("), ("rowdot", "T rowdot("),
       ("check", "__device__ int check_stopping("), ("pbar", "T pbar("), ("polish", "__device__ unsigned polish("),
       ("form_kkt", "__device__ void form_reduced_kkt("), ("setup", "__device__ int setup("), ("admm_loop(smem)", "__device__ int admm_loop("),
       ("admm_loop_reg", "__device__ int admm_loop_reg("), ("solve(load..store)", "__device__ void solve("),
       ("stages", "// out-of-line stages\n")]
for name, p in pat:
    for i, l in enumerate(lines):
        if p.strip() in l and "template <typename T, int G> __device__" not in l and "__noinline__" not in l:
            marks.append((name, i + 1)); break
marks.sort(key=lambda t: t[1])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]; ci_smp = h.index("# Samples"); ci_ins = h.index("Instructions Executed"); ci_wf = h.index("L1 Wavefronts Shared")
agg = {}
for r in rows[hi + 1:]:
    if len(r) <= ci_wf or not r[0].strip(): continue
    try: ln = int(r[0]); sm = float(r[ci_smp] or 0); ins = float(r[ci_ins] or 0); wf = float(r[ci_wf] or 0)
    except ValueError: continue
    name = [m for m in marks if m[1] <= ln][-1][0]
    a = agg.setdefault(name, [0, 0, 0]); a[0] += sm; a[1] += ins; a[2] += wf
ts = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values()); tw = sum(a[2] for a in agg.values())
print(f"{'phase':18s} samples%  inst%  smemwf%")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]): print(f"{k:18s} {100*a[0]/ts:7.2f} {100*a[1]/ti:6.2f} {100*a[2]/tw:7.2f}")
