#!/bin/bash
# quick A/B on the GPU box: deposit timings at 1/8 of config 3 (same density), then a few GPU parity tests
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
timeout 200 python scripts/exp_deposit.py --n 512 --N 125000000 --reps 5
timeout 300 python -m pytest tests/test_gpu_tsc.py -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_quick.log
