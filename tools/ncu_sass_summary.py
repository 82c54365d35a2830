"""Summarise an `ncu --page source --csv` dump: executed warp-instructions by opcode and by
execution-count band (the inner loops), plus the stall samples per band."""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot = 0
by_op = defaultdict(float)
by_band = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
for r in data:
    try:
        n = float(r[ix["Instructions Executed"]])
    except (ValueError, IndexError):
        continue
    sass = r[ix["Source"]].strip()
    op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
    op = op.split(".")[0]
    tot += n
    by_op[op] += n
    smp = float(r[ix["# Samples"]] or 0)
    band = int(n)
    b = by_band[band]
    b[0] += n
    b[1] += smp
    b[2][op] += n
print("total warp-instructions executed: %.4g" % tot)
print("-- by opcode")
for op, n in sorted(by_op.items(), key=lambda x: -x[1])[:28]:
    print("  %-10s %12.4g  %5.1f%%" % (op, n, 100 * n / tot))
print("-- by execution-count band (instructions executed the same number of times = one loop body)")
for band, (n, smp, ops) in sorted(by_band.items(), key=lambda x: -x[1][0])[:12]:
    top = ", ".join("%s %.0f%%" % (o, 100 * c / n) for o, c in sorted(ops.items(), key=lambda x: -x[1])[:8])
    print("  count=%-10d share %5.1f%%  samples %6.0f  [%s]" % (band, 100 * n / tot, smp, top))
