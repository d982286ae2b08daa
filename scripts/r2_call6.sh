#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== 64-bit indexing test"
timeout 600 python -m pytest tests/test_gpu_tsc.py -m gpu -x -q -k "64bit" 2>&1 | tail -4
echo "== ingest kernels"
timeout 300 python scripts/ingest_bench.py | tee gpurun_out/r2_ingest_bench.json
echo "== early groups (host input)"
for g in 1 2 3; do
ABK_EARLY_GROUPS=$g timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('early_groups=$g', round(d['value'],2), 'e2e', round(d['e2e']['value'],2))"
done
} 2>&1 | tee gpurun_out/r2_call6.log
