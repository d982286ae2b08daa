#!/bin/bash
# device-resident input: early tile-deposit groups on the auxiliary stream while later segments are still being bucketed
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
for cfg in "0 0" "1 0" "2 0" "3 0" "1 7" "2 4" "3 2"; do
set -- $cfg
if [ "$2" != "0" ]; then export ABK_TAIL_SEGMENTS=$2; else unset ABK_TAIL_SEGMENTS; fi
ABK_DEVICE_EARLY=$1 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('device_early=$1 tail=$2', round(d['value'],2), {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"
done
} 2>&1 | tee gpurun_out/r2_quick5.log
