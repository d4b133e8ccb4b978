"""Summarise an .ncu-rep (first kernel): key metrics, stall breakdown, hottest source lines. Dev tool (runs without a GPU)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = dict(zip(hdr, vals))
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__cycles_elapsed.avg", "smsp__average_warp_latency_per_inst_issued.ratio"]
for k in keys:
    if k in d: print(f"{k:80s} {d[k]} {units[hdr.index(k)]}")
st = {h[len('smsp__pcsamp_warps_issue_stalled_'):]: float(v) for h, v in d.items() if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('not_issued')}
tot = sum(st.values()) or 1
print("stalls:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
h = rows[hi]
ci_smp = h.index("# Samples"); ci_ins = h.index("Instructions Executed")
ci_wf = h.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in h else ci_ins  # kernels without shared memory
ci_wfi = h.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in h else ci_ins
lines = []
for r in rows[hi + 1:]:
    if len(r) <= ci_wf or not r[0].strip(): continue
    try: lines.append((int(r[0]), r[1].strip(), float(r[ci_smp] or 0), float(r[ci_ins] or 0), float(r[ci_wf] or 0), float(r[ci_wfi] or 0)))
    except ValueError: pass
ts = sum(l[2] for l in lines) or 1; ti = sum(l[3] for l in lines) or 1; tw = sum(l[4] for l in lines) or 1
print(f"source lines: total samples {ts:.0f}, inst {ti:.3e}, smem wavefronts {tw:.3e}")
print("  line  samples%   inst%  smemwf% (excess)  source")
for ln, txt, sm, ins, wf, wfi in sorted(lines, key=lambda l: -l[2])[:topn]:
    print(f"{ln:6d}  {100*sm/ts:6.2f}  {100*ins/ti:6.2f}  {100*wf/tw:6.2f} ({100*(wf-wfi)/max(wf,1):3.0f}%)  {txt[:110]}")
