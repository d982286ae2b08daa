#!/bin/bash
# multi-GPU bench at the world size given as $1 (phases + parity block in the JSON line)
set -u
G=$1
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $G --steps 5 --warmup 3 > gpurun_out/r2_bench_n$G.json 2> gpurun_out/r2_bench_n$G.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$G.json').read().strip().splitlines()[-1])
print('N=$G value', round(d['value'],2), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, 'parity ok', d['parity']['ok'] if d.get('parity') else None)
print(' stages', {k:round(v['ms_per_step'],2) for k,v in d['stages'].items()})
print(' phases', {k:round(v,2) for k,v in d.get('phases_ms_rank0',{}).items()})
PY
tail -3 gpurun_out/r2_bench_n$G.err | cut -c1-300
