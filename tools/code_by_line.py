"""Instruction count per source line for one kernel of libsfb.so (dev tool). usage: code_by_line.py <substr of mangled name> [top]"""
import re, collections, subprocess, sys, os, tempfile
lib = "smooth_feedback_b200/lib/libsfb.so"
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
sub = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
per = collections.defaultdict(collections.Counter); func = None; cur = None
for line in dis.splitlines():
    if line.startswith("\t.section") and ".text." in line:
        func = line.split(".text.")[1].split(",")[0]; cur = None; continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = int(m.group(2)); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): per[func][cur] += 1
src = open("smooth_feedback_b200/csrc/qp_dense_group.cuh").read().splitlines()
for f, c in per.items():
    if sub not in f: continue
    tot = sum(c.values()); print(f, f"{tot*16/1024:.1f} KB")
    for k, v in c.most_common(top): print(f"  {v:6d}  L{k}: {src[k-1].strip()[:100] if k and k <= len(src) else ''}")
