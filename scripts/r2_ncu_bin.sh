#!/bin/bash
# ncu --set full of the (k,mu) binning kernel at nmesh 1024 (few particles: the kernel only depends on the mesh)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 ncu --set full --clock-control none --import-source on -k regex:power_bin -c 1 -f -o gpurun_out/prof_r2d_bin python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --nparticles 20000000 > gpurun_out/ncu_bin.log 2>&1
ncu -i gpurun_out/prof_r2d_bin.ncu-rep --page raw --csv > gpurun_out/prof_r2d_bin_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r2d_bin.ncu-rep --page source --csv > gpurun_out/prof_r2d_bin_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_bin.log | cut -c1-300
