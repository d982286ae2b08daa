#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
timeout 900 python -m pytest tests/test_gpu_power.py tests/test_gpu_zz_kppi.py -m gpu -x -q 2>&1 | tail -3
} 2>&1 | tee gpurun_out/r2_quick11.log
