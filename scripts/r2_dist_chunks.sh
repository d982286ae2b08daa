#!/bin/bash
# sharded path at world size $1: number of exchange chunks per rank
set -u
G=$1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for ch in 4 8; do
ABK_DIST_CHUNKS=$ch ABK_BENCH_NO_PARITY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 3 --warmup 2 --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('chunks=$ch N=$G value', round(d['value'],2), {k:round(v,2) for k,v in d.get('phases_ms_rank0',{}).items()}, {k:round(v['ms_per_step'],2) for k,v in d['stages'].items() if 'bucket' in k})"
done 2>&1 | tee gpurun_out/r2_dist_chunks_n$G.log
