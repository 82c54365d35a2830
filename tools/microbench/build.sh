#!/bin/bash
# Builds the two microbenchmarks DESIGN.md section 5 rests on (run them on the B200 with gpurun):
#   rf_bench       issue rate of FFMA / FFMA2 / FMNMX as a function of register-operand count
#   devloop_bench  the raceline-deviation segment loop in isolation, several formulations
cd "$(dirname "$0")"
for b in rf_bench devloop_bench mma_bench; do
  nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o $b $b.cu || exit 1
done
