#!/bin/bash
# round-2 GPU call 2: first run of the walk deposit kernel: parity tests, then A/B timing against the round-1 kernel
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== gpu tests (tsc + power)"
timeout 900 python -m pytest tests/test_gpu_tsc.py tests/test_gpu_power.py -m gpu -x -q 2>&1 | tail -8
echo "== bench walk kernel"
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['config2_tsc'])"
echo "== bench round-1 kernel"
ABK_TILE_KNOBS=0x10000 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()}, d['config2_tsc'])"
echo "== done"
} 2>&1 | tee gpurun_out/r2_call2.log
