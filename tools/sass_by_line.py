"""Static SASS instruction count per CUDA source line of one kernel (needs -lineinfo):

    python tools/sass_by_line.py <mangled-kernel-substring> [lib.so] [n_top]

For straight-line per-candidate code the static count is the dynamic count per candidate, so this
shows where the instructions outside the hot loop come from without a GPU.  Inlined device
functions are attributed to the innermost line (the callee's)."""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

kern = sys.argv[1]
so = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else "f1tenth_planning_b200/lib/libf1l.so")
n_top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False
cur = ("?", 0)
per_line = defaultdict(Counter)
order = []
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m:
        per_line[cur][m.group(1)] += 1
        order.append(cur)
tot = sum(sum(c.values()) for c in per_line.values())
print("%d SASS instructions in *%s*" % (tot, kern))
srcs = {}
for loc, c in sorted(per_line.items(), key=lambda x: -sum(x[1].values()))[:n_top]:
    f = loc[0]
    if f not in srcs:
        p = os.path.join(os.path.dirname(so), "..", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][loc[1] - 1].strip()[:70] if 0 < loc[1] <= len(srcs[f]) else ""
    print("%5d  %s:%-4d %-70s | %s" % (sum(c.values()), f, loc[1], text,
                                       " ".join("%s %d" % kv for kv in c.most_common(6))))
