#!/bin/bash
# clustered host input: early deposit groups / tail segments
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for cfg in "1 0" "2 4" "2 5"; do
set -- $cfg
if [ "$2" != "0" ]; then export ABK_TAIL_SEGMENTS=$2; else unset ABK_TAIL_SEGMENTS; fi
ABK_EARLY_GROUPS=$1 timeout 300 python bench.py --clustered 0.5 --no-cpu --steps 2 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('clustered groups=$1 tail=$2', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done
} 2>&1 | tee gpurun_out/r2_quick12.log
