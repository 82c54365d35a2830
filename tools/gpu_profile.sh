#!/bin/bash
# ncu captures for profiles/:  tools/gpu_profile.sh <tag>
#   launch list (gpu__time_duration) of the device-resident bench step, one --set full capture each of
#   eval_kernel (bench workload, full scan and pruned), pp_scan_kernel (config 2)
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval_$tag \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/ncu_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval_pruned_$tag \
    python bench.py --steps 2 --warmup 1 --quick --prune 1 > gpurun_out/ncu_pruned_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pp_scan_kernel -s 2 -c 1 -f -o gpurun_out/prof_pp_$tag \
    python tools/run_c2.py > gpurun_out/ncu_pp_$tag.log 2>&1
ls -la gpurun_out/*.ncu-rep
