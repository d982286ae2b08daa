"""Experiment (GPU box): the sharded pipeline at the per-GPU footprint of nmesh=4096 on 8 GPUs,
emulated on ONE GPU as nmesh=2048 with world_size 1 (slab = 34.5 GB, 1.25e9 particles)."""
import argparse
import socket
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from abacusutils_b200 import dist as abk_dist
from abacusutils_b200._lib import Engine

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=2048)
ap.add_argument('--N', type=int, default=1_250_000_000)
ap.add_argument('--interlaced', type=int, default=0)
args = ap.parse_args()

with socket.socket() as s:
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
torch.cuda.set_device(0)
dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{port}', rank=0, world_size=1, device_id=torch.device('cuda', 0))
L = 4000.0
g = torch.Generator(device='cuda'); g.manual_seed(6)
pos = torch.rand((args.N, 3), device='cuda', generator=g) * L
for it in range(2):
    torch.cuda.reset_peak_memory_stats()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t = abk_dist.calc_power(pos, L, kbins=100, mubins=10, nmesh=args.n, compensated=True, interlaced=bool(args.interlaced), poles=[0, 2, 4])
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'iter {it}: n={args.n} N={args.N} interlaced={args.interlaced}: {dt*1e3:.1f} ms, peak {torch.cuda.max_memory_allocated()/2**30:.1f} GiB allocated,'
          f' {torch.cuda.max_memory_reserved()/2**30:.1f} GiB reserved; N_mode total {int(np.asarray(t["N_mode"]).sum())}, P[5]={np.asarray(t["power"])[5][:3]}')
    Engine.get(0).release_scratch() if it == 0 else None
dist.destroy_process_group()
