#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for b in 0 1; do
ABK_CONCURRENT_DEPOSITS=$b timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('concurrent_deposits=$b', round(d['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
done
} 2>&1 | tee gpurun_out/r2_quick4.log
