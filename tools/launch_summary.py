"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: launches, total time and share per kernel."""
import collections, csv, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    a = agg.setdefault(r[ki].split("(")[0][:100], [0, 0.0]); a[0] += 1; a[1] += float(r[vi])
tot = sum(v[1] for v in agg.values())
print("# kernel, launches, total_ms, share_of_gpu_time")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k}, {v[0]}, {v[1] / 1e6:.3f}, {100 * v[1] / tot:.2f}%")
