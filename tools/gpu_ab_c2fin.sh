#!/bin/bash
# A/B of K1 finish-kernel variants: per-kernel durations (ncu launch list) of config 2
for v in "$@"; do
  export F1L_LIB=$PWD/f1tenth_planning_b200/lib/variants/libf1l_$v.so
  ncu --metrics gpu__time_duration.sum --clock-control none -c 8 --csv --log-file gpurun_out/k1_$v.csv python tools/run_c2.py >/dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/k1_$v.csv")))
h=[r for r in rows if "Kernel Name" in r][0]
t={}
for r in rows:
    if len(r)==len(h) and r!=h:
        d=dict(zip(h,r)); t.setdefault(d["Kernel Name"][:16],[]).append(float(d["Metric Value"]))
print("$v", {k: round(sum(x)/len(x)/1e3,2) for k,x in t.items()})
PY
  python tools/run_c2.py
done
