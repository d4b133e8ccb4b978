"""Print SASS code size per kernel of libsfb.so (dev tool)."""
import sys, re, collections, subprocess
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1] if len(sys.argv) > 1 else "smooth_feedback_b200/lib/libsfb.so"], capture_output=True, text=True).stdout
cnt = collections.Counter(); name = None
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m: name = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line): cnt[name] += 1
for n, c in sorted(cnt.items(), key=lambda kv: kv[1]): print(f'{c*16/1024:8.1f} KB  {n[:110]}')
