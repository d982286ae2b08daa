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
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_tile_walk -s 2 -c 2 -f -o gpurun_out/prof_r2e_walk python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:tsc_bucket -s 30 -c 4 -f -o gpurun_out/prof_r2e_bucket python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bucket.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:power_bin -s 1 -c 1 -f -o gpurun_out/prof_r2e_bin python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bin.log 2>&1
python scripts/ncu_extract.py gpurun_out/prof_r2e_walk.ncu-rep gpurun_out/prof_r2e_bucket.ncu-rep gpurun_out/prof_r2e_bin.ncu-rep
cp profiles/r2_ncu_metrics.json profiles/prof_r2e_*_raw.csv gpurun_out/ 2>/dev/null
echo "== bench again (with the fresh ncu metrics)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; tail -c 300 gpurun_out/r2_final_bench.json
echo "== done"
} 2>&1 | tee gpurun_out/r2_profile.log
