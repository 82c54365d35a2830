#!/bin/bash
# quick A/B of kernel-variant builds (tools/build_variants.py) on the GPU box, device-resident arm only:
#   tools/gpu_ab_quick.sh <variant> ...      ("default" = the product library)
for v in "$@"; do
  if [ "$v" = default ]; then unset F1L_LIB; else export F1L_LIB=$PWD/f1tenth_planning_b200/lib/variants/libf1l_$v.so; fi
  python bench.py --quick --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('%-10s eval %.3f ms  step %.3f ms  %s  checksum %.6f' % ('$v', d['kernels_ms']['eval'], d['ms_per_step'], d['kernel'], d['checksum']))"
done
