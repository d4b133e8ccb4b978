"""SASS size per device function inside one kernel (dev tool). usage: code_by_func.py <substr of kernel name>"""
import re, collections, subprocess, sys, os, tempfile
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath("smooth_feedback_b200/lib/libsfb.so")], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout
cnt = collections.Counter(); cur = None; sec = None
for line in dis.splitlines():
    if line.startswith('\t.section') and '.text.' in line:
        sec = line.split('.text.')[1].split(',')[0]; cur = "<kernel body>"; continue
    m = re.match(r'^(\S+):\s*$', line)
    if m and not m.group(1).startswith('.L'): cur = m.group(1).split('$')[-1]
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and sec and sys.argv[1] in sec: cnt[cur] += 1
for k, v in cnt.most_common(14): print(f'{v*16/1024:8.1f} KB {k[:120]}')
