#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for t in 2 3 4; do for g in 1 2; do
ABK_TAIL_SEGMENTS=$t ABK_EARLY_GROUPS=$g timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tail=$t groups=$g', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done; done
} 2>&1 | tee gpurun_out/r2_call7.log
