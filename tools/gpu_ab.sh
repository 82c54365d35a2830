#!/bin/bash
# A/B bench of kernel-variant builds on the GPU box:  tools/gpu_ab.sh <tag> <variant> ...
# (variants built by tools/build_variants.py; "default" = the product library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset F1L_LIB; else export F1L_LIB=$PWD/f1tenth_planning_b200/lib/variants/libf1l_$v.so; fi
  python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ab_${tag}_$v.json 2> gpurun_out/ab_${tag}_$v.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/ab_${tag}_$v.json"))
    print("$v", "value %.4g" % d["value"], "eval_ms %.3f" % d["kernels_ms"]["eval"], "frac %.3f" % d["roofline"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("$v failed", e)
PY
done
