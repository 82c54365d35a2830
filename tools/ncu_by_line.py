"""Attribute an `ncu --page source --csv` (SASS view) dump to CUDA source lines.

    python tools/ncu_by_line.py <src.csv> <mangled-kernel-substring> [lib.so] [n_top]

The SASS rows of the dump are in address order, the same order nvdisasm prints the function in,
so row k pairs with the k-th instruction of `nvdisasm -c -g` (which carries //## File/line marks;
needs -lineinfo at build time)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

src_csv, kern = sys.argv[1], sys.argv[2]
so = sys.argv[3] if len(sys.argv) > 3 else "f1tenth_planning_b200/lib/libf1l.so"
n_top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
so = os.path.abspath(so)
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
lines = []
cur = ("?", 0)
inside = False
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
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
if len(data) != len(lines):
    print("warning: %d SASS rows in the profile vs %d in the cubin (different build?)" % (len(data), len(lines)))
agg = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r, loc in zip(data, lines):
    smp = float(r[ix["# Samples"]] or 0)
    n = float(r[ix["Instructions Executed"]] or 0)
    a = agg[loc]
    a[0] += smp
    a[1] += n
    for c in stall_cols:
        a[2][c[6:]] += float(r[ix[c]] or 0)
tot_s = sum(a[0] for a in agg.values())
tot_n = sum(a[1] for a in agg.values())
srcs = {}
print("total samples %.0f, warp-instructions %.4g" % (tot_s, tot_n))
for loc, (smp, n, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:n_top]:
    f = loc[0]
    if f not in srcs:
        p = os.path.join(os.path.dirname(so), "..", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = srcs[f][loc[1] - 1].strip()[:86] if 0 < loc[1] <= len(srcs[f]) else ""
    top = ", ".join("%s %.0f%%" % (k, 100 * v / max(smp, 1)) for k, v in sorted(st.items(), key=lambda x: -x[1])[:3])
    print("%5.2f%% smp %5.2f%% inst  %s:%-4d %-86s | %s" % (100 * smp / tot_s, 100 * n / tot_n, f, loc[1], text, top))
