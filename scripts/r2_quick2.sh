#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
timeout 200 python scripts/exp_deposit.py --n 512 --N 125000000 --reps 5
echo "== bench default (1 device segment)"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
echo "== bench 14 device segments"
ABK_DEVICE_SEGMENTS=14 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
timeout 300 python -m pytest tests/test_gpu_tsc.py -m gpu -x -q 2>&1 | tail -2
} 2>&1 | tee gpurun_out/r2_quick2.log
