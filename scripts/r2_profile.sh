#!/bin/bash
# evidence for profiles/: bench line, ncu launch list of the same command, ncu --set full of the top kernels
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== bench"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 400 gpurun_out/r2_final_bench.json; tail -3 gpurun_out/r2_final_bench.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_tile_walk -s 2 -c 2 -f -o gpurun_out/prof_r2c_walk python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"tsc_bucket|power_bin" -s 30 -c 4 -f -o gpurun_out/prof_r2c_bucket_bin python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bucket.log 2>&1
ls -la gpurun_out/prof_r2c*
echo "== done"
} 2>&1 | tee gpurun_out/r2_profile.log
