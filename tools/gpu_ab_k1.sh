#!/bin/bash
# A/B of K1 (config 2) variant builds on the GPU box:  tools/gpu_ab_k1.sh <variant> ...
# timing by tools/run_c2.py, parity by the K1 GPU tests
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = default ]; then unset F1L_LIB; else export F1L_LIB=$PWD/f1tenth_planning_b200/lib/variants/libf1l_$v.so; fi
  echo -n "$v: "; python tools/run_c2.py 2>&1 | tail -1
  echo -n "$v small: "; python tools/run_c2.py 4096 2>&1 | tail -1
  python -m pytest tests/test_gpu_pure_pursuit.py -x -q -m gpu 2>&1 | tail -1
done
