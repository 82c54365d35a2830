#!/bin/bash
# ncu captures of one bench command:  tools/gpu_profile.sh <tag> [extra bench.py arguments]
#   launch list (gpu__time_duration) + one --set full capture of eval_kernel
tag=$1; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --scenarios 100000 --no-extras --no-cpu-baseline "$@" > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 1 -c 1 -f -o gpurun_out/prof_eval_$tag \
    python bench.py --steps 2 --warmup 1 --scenarios 100000 --no-extras --no-cpu-baseline "$@" > gpurun_out/ncu_$tag.log 2>&1
