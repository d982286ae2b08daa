#!/bin/bash
# round-2 GPU call 4: whole GPU suite with the new tests, bench with the reference arm, ncu of the top kernels on bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== gpu suite"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== bench (ours)"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 3000 gpurun_out/r2_bench.json; tail -5 gpurun_out/r2_bench.err
echo "== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; cat gpurun_out/r2_bench_ref.json; tail -5 gpurun_out/r2_bench_ref.err
echo "== done"
} 2>&1 | tee gpurun_out/r2_call4.log
