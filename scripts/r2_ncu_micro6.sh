#!/bin/bash
# ncu of the staged multisplit prototype (scripts/micro/split_micro.cu): both levels, first variant
set -u
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/split_micro scripts/micro/split_micro.cu || exit 1
timeout 280 ncu --set full --clock-control none --import-source on -k regex:split_kernel -c 2 -f -o gpurun_out/r2_micro6 /tmp/split_micro prof > gpurun_out/r2_ncu_micro6.log 2>&1
ncu -i gpurun_out/r2_micro6.ncu-rep --page raw --csv > gpurun_out/r2_micro6_raw.csv 2>/dev/null
tail -5 gpurun_out/r2_ncu_micro6.log
