#!/bin/bash
# One gpurun call for the start of a GPU session (1 GPU):  gpurun --timeout 1500 -- 'bash scripts/gpu_session.sh'
# Everything is wrapped in its own `timeout`; outputs go to gpurun_out/ (merged back by gpurun).
#   1. the newest, not-yet-device-run tests first, then the whole -m gpu suite
#   2. micro-benchmarks the pending knobs depend on
#   3. A/B timing of every experiment knob (device-resident and pinned-host input)
#   4. the bench line, its ncu launch list, and one full capture of the top kernel
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."

echo "== new tests" | tee gpurun_out/session.log
timeout 300 python -m pytest tests/test_gpu_zz_kppi.py tests/test_gpu_zzz_ingest.py -m gpu -q -x 2>&1 | tail -5 | tee -a gpurun_out/session.log
echo "== gpu suite" | tee -a gpurun_out/session.log
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee -a gpurun_out/session.log

echo "== micro" | tee -a gpurun_out/session.log
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atoms_micro scripts/micro/atoms_micro.cu && timeout 60 /tmp/atoms_micro 2>&1 | tee -a gpurun_out/session.log

echo "== knobs" | tee -a gpurun_out/session.log
timeout 900 python scripts/exp_knobs.py --host --variants 2>&1 | tee gpurun_out/knobs.log | tail -40 | tee -a gpurun_out/session.log

echo "== bench" | tee -a gpurun_out/session.log
timeout 400 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 600 gpurun_out/bench.json | tee -a gpurun_out/session.log
timeout 300 python bench.py --steps 1 --warmup 1 --no-cpu --packed pack9 > gpurun_out/bench_packed.json 2>> gpurun_out/bench.err

echo "== ncu launch list" | tee -a gpurun_out/session.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: tile deposit" | tee -a gpurun_out/session.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_tile_deposit -s 2 -c 2 -o gpurun_out/prof_deposit \
    python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_deposit.log 2>&1
echo "== done" | tee -a gpurun_out/session.log
