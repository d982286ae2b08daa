"""BASELINE configs[3] and configs[4] on 8 GPUs (torchrun):
  configs[3]: cross-power of two weighted catalogues (5e8 each), nmesh=2048, interlaced, slab-sharded
  configs[4]: nmesh=4096 auto-power of 1e10 particles (1.25e9 per GPU), x-slab sharded, non-interlaced
Prints time, peak memory, mode-count total and a shot-noise sanity check (uniform random particles:
P(k) ~ L^3 * <w^2> / (N <w>^2 ...) -- for unit weights L^3/N)."""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from abacusutils_b200 import dist as abk_dist

which = sys.argv[1] if len(sys.argv) > 1 else '4'
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
rank, world = dist.get_rank(), dist.get_world_size()
g = torch.Generator(device='cuda')
g.manual_seed(100 + rank)


def run(tag, n, L, N_total, interlaced, cross, weighted, reps=2):
    nl = N_total // world
    pos = torch.rand((nl, 3), device='cuda', generator=g) * L
    w = torch.rand((nl,), device='cuda', generator=g) if weighted else None
    pos2 = w2 = None
    if cross:
        pos2 = torch.rand((nl, 3), device='cuda', generator=g) * L
        w2 = torch.rand((nl,), device='cuda', generator=g) if weighted else None
    for it in range(reps):
        torch.cuda.reset_peak_memory_stats()
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        t = abk_dist.calc_power(pos, L, kbins=100, mubins=10, nmesh=n, compensated=True, interlaced=interlaced,
                                w=w, pos2=pos2, w2=w2, poles=[0, 2, 4])
        torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
        peak = torch.tensor([torch.cuda.max_memory_allocated() / 2**30], device='cuda')
        dist.all_reduce(peak, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(f'{tag} iter {it}: nmesh={n} N={N_total} x{2 if cross else 1} G={world} interlaced={interlaced}: '
                  f'{dt*1e3:.1f} ms, peak {peak.item():.1f} GiB/GPU, N_mode total {int(np.asarray(t["N_mode"]).sum())}, '
                  f'P[k-bin 50]={np.asarray(t["power"])[50][:3]}, L^3/N={L**3/N_total:.4g}', flush=True)
    del pos, w, pos2, w2
    torch.cuda.empty_cache()


if which in ('3', 'both'):
    run('configs[3]', 2048, 2000.0, 500_000_000, True, True, True)
if which in ('4', 'both'):
    run('configs[4]', 4096, 4000.0, 10_000_000_000, False, False, False)
dist.barrier()
dist.destroy_process_group()
