#!/bin/bash
# host input: early groups / tail segments / a short last chunk
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for cfg in "1 0 1" "2 2 1" "2 1 0.25" "3 1 0.25" "2 1 0.5" "2 2 0.5"; do
set -- $cfg
if [ "$2" != "0" ]; then export ABK_TAIL_SEGMENTS=$2; else unset ABK_TAIL_SEGMENTS; fi
ABK_EARLY_GROUPS=$1 ABK_LAST_CHUNK_FRAC=$3 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('groups=$1 tail=$2 lastfrac=$3', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done
} 2>&1 | tee gpurun_out/r2_quick9.log
