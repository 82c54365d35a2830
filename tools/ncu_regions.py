"""Stage breakdown of eval_kernel from an `ncu --page source --csv` dump: every SASS instruction is
attributed to the outermost call site inside the kernel body (nvdisasm -gi inline chains), and the
kernel body is cut into stages at its `// ----` marker comments.

    python tools/ncu_regions.py <src.csv> <mangled-kernel-substring> [lib.so]
"""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src_csv, kern = sys.argv[1], sys.argv[2]
so = os.path.abspath(sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "f1tenth_planning_b200/lib/libf1l.so"))
SRC = os.path.join(ROOT, "f1tenth_planning_b200/csrc/f1l_lattice.cuh")
text = open(SRC).read().splitlines()


def find(pat, start=0):
    for i in range(start, len(text)):
        if pat in text[i]:
            return i + 1
    raise SystemExit("marker not found: " + pat)


k0 = find("eval_kernel(EvalArgs a)")
marks = [("prologue (window table)", k0),
         ("goals + LUT seeds + Newton, four candidates per warp", find("// A warp takes `item` consecutive", k0)),
         ("solution hand-over (or single-candidate Newton)", find("// ---- goal, seed, Newton", k0)),
         ("arc samples", find("// ---- arc samples", k0)),
         ("curvature / validity / slab", find("// ---- curvature terms", k0)),
         ("similarity + collision", find("// ---- similarity", k0)),
         ("deviation: setup + pruning bound", find("// ---- raceline deviation", k0)),
         ("deviation: segment loop", find("uint32_t ta = cbase + L::C_TAB + (k_begin + ggi) * 32;", k0)),
         ("deviation: reduce", find("float rows[S];", k0) - 3),
         ("cost + output + next candidate", find("if (!(flags & (F1L_FLAG_COLLIDE_OPP", k0)),
         ("end", find("// K5: select", k0))]

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-c", "-gi", cubin], capture_output=True, text=True).stdout.splitlines()
locs, chain, inside = [], [], False
for ln in dis:
    if ln.startswith("//---") and ".text." in ln:
        inside = kern in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        chain.append(m.groups())
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        if chain:
            cur = chain
            chain = []
        # outermost site: the last chain entry's "inlined at" if present, else its own line
        f, l, fi, li = cur[-1]
        site = (os.path.basename(fi), int(li)) if fi else (os.path.basename(f), int(l))
        locs.append(site)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > ix["# Samples"]]
if len(data) != len(locs):
    print("warning: %d SASS rows in the profile vs %d in the cubin (different build?)" % (len(data), len(locs)))
agg = defaultdict(lambda: [0.0, 0.0])
for r, (f, l) in zip(data, locs):
    name = "other (%s)" % f
    if f == "f1l_lattice.cuh":
        for (n, a), (_, b) in zip(marks[:-1], marks[1:]):
            if a <= l < b:
                name = n
    agg[name][0] += float(r[ix["# Samples"]] or 0)
    agg[name][1] += float(r[ix["Instructions Executed"]] or 0)
ts = sum(v[0] for v in agg.values())
ti = sum(v[1] for v in agg.values())
print("| stage | stall samples | warp-instructions |")
print("|---|---|---|")
order = [n for n, _ in marks[:-1]] + sorted(k for k in agg if k.startswith("other"))
for n in order:
    if n in agg:
        print("| %s | %.1f %% | %.1f %% |" % (n, 100 * agg[n][0] / ts, 100 * agg[n][1] / ti))
