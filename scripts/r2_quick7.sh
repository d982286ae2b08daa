#!/bin/bash
# walk kernel with the cross-warp row exchange: deposit timings vs particle count, parity tests, bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for N in 100000000 600000000 1000000000; do echo N=$N; timeout 200 python scripts/exp_deposit.py --n 1024 --N $N --reps 3 2>&1 | tail -3; done
timeout 600 python -m pytest tests/test_gpu_tsc.py tests/test_gpu_power.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
} 2>&1 | tee gpurun_out/r2_quick7.log
