#!/bin/bash
# round-2 GPU call 5: evidence runs -- full-size CPU reference check, clustered input, packed input, ncu captures of bench
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== reference arm with one full-size run"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 --cpu-full > gpurun_out/r2_bench_ref_full.json 2> gpurun_out/r2_bench_ref_full.err; cat gpurun_out/r2_bench_ref_full.json; tail -3 gpurun_out/r2_bench_ref_full.err
echo "== clustered 0.5"
timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --clustered 0.5 > gpurun_out/r2_bench_clustered.json 2> gpurun_out/r2_bench_clustered.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_clustered.json').read()); print(d['value'], d['e2e']['value'], {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})"; tail -3 gpurun_out/r2_bench_clustered.err
echo "== packed"
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --packed pack9 > gpurun_out/r2_bench_pack9.json 2> gpurun_out/r2_bench_pack9.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_pack9.json').read()); print(d['value'], d['e2e'], d['e2e_packed'])"; tail -3 gpurun_out/r2_bench_pack9.err
timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --packed rvint > gpurun_out/r2_bench_rvint.json 2> gpurun_out/r2_bench_rvint.err; python -c "import sys,json; d=json.loads(open('gpurun_out/r2_bench_rvint.json').read()); print(d['value'], d['e2e_packed'])"; tail -3 gpurun_out/r2_bench_rvint.err
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
echo "== ncu full: walk, bucket, bin"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tsc_tile_walk -s 2 -c 2 -f -o gpurun_out/prof_r2b_walk python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_walk.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"tsc_bucket|power_bin" -s 30 -c 4 -f -o gpurun_out/prof_r2b_bucket_bin python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bucket.log 2>&1
ls -la gpurun_out/*.ncu-rep
echo "== done"
} 2>&1 | tee gpurun_out/r2_call5.log
