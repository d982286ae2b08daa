#!/bin/bash
# round-2 GPU call 1: micro-benchmarks behind the deposit redesign, host facts, pending round-1 knobs
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
{
echo "== host"; nproc; free -g | head -2; python -c "import numba, scipy; print('numba', numba.__version__, 'scipy', scipy.__version__)" 2>&1 | tail -1
lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket' 
echo "== micro"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/deposit_micro scripts/micro/deposit_micro.cu && timeout 180 /tmp/deposit_micro
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/atoms_micro scripts/micro/atoms_micro.cu && timeout 60 /tmp/atoms_micro
echo "== knobs"
timeout 600 python scripts/exp_knobs.py --variants 2>&1 | tail -40
echo "== done"
} 2>&1 | tee gpurun_out/r2_call1.log
