#!/bin/bash
# round-2 GPU call 3: ncu full capture of the walk deposit kernel (1/8 of config 3 at equal density)
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
timeout 200 python scripts/exp_deposit.py --n 512 --N 125000000 --reps 3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tsc_tile_walk -s 4 -c 2 -f -o gpurun_out/prof_r2a_walk \
    python scripts/exp_deposit.py --n 512 --N 125000000 --reps 1 2>&1 | tail -5
ls -la gpurun_out/*.ncu-rep
} 2>&1 | tee gpurun_out/r2_call3.log
