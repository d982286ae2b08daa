#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for N in 100000000 1000000000; do echo N=$N; timeout 200 python scripts/exp_deposit.py --n 1024 --N $N --reps 3 2>&1 | tail -3; done
timeout 600 python -m pytest tests/test_gpu_tsc.py -m gpu -x -q 2>&1 | tail -2
} 2>&1 | tee gpurun_out/r2_quick8.log
