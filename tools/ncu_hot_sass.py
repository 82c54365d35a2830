"""Top SASS instructions by stall samples from an `ncu --page source --csv` dump."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n_top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
out = []
for ln, r in enumerate(rows[2:]):
    try:
        smp = float(r[ix["# Samples"]] or 0)
        n = float(r[ix["Instructions Executed"]] or 0)
    except (ValueError, IndexError):
        continue
    st = sorted(((float(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    out.append((smp, ln, n, r[ix["Source"]].strip()[:70], ", ".join("%s %.0f" % (c, v) for v, c in st if v > 0)))
tot = sum(o[0] for o in out)
print("total samples", tot)
for smp, ln, n, sass, st in sorted(out, reverse=True)[:n_top]:
    print("%6.0f %5.2f%%  #%-5d x%-9.0f %-70s | %s" % (smp, 100 * smp / tot, ln, n, sass, st))
