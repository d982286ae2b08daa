"""A/B timing of the pending experiment knobs on ONE B200 (run under gpurun, wrapped in `timeout`):

    timeout 600 python scripts/exp_knobs.py [--N 1000000000] [--nmesh 1024] [--reps 3] [--host]

Times device-resident (and with --host, pinned-host) calc_power at config 3 for every combination of
  ABK_FUSED_NORMALIZE   0 / 1    normalize_field folded into the deposit (default 1)
  ABK_DEVICE_SEGMENTS   - / 1 / 4 number of bucket segments for device-resident input (default: 14)
  ABK_SCATTER           1 / 2     one-level scattered stores vs two-level multisplit with coalesced runs (default 1)
  ABK_EARLY_GROUPS      1 / 2 / 3 early deposit groups while host chunks are still arriving (--host, default 1)
and, with the best of those, the deposit kernel variants (abk_ctx_set_tile_capacity bits 16-18).
Prints one line per configuration: CUDA-event ms per step (median of reps after 2 warm-ups) and the per-kernel stage times.
"""
import argparse
import itertools
import os
import statistics
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from abacusutils_b200._lib import Engine, check
from abacusutils_b200.analysis.power_spectrum import calc_power

ap = argparse.ArgumentParser()
ap.add_argument('--N', type=int, default=1_000_000_000)
ap.add_argument('--nmesh', type=int, default=1024)
ap.add_argument('--L', type=float, default=2000.0)
ap.add_argument('--reps', type=int, default=3)
ap.add_argument('--host', action='store_true')
ap.add_argument('--variants', action='store_true', help='also sweep the deposit kernel variants')
args = ap.parse_args()

torch.cuda.set_device(0)
eng = Engine.get(0)
gen = torch.Generator(device='cuda')
gen.manual_seed(3)
pos = torch.rand((args.N, 3), device='cuda', dtype=torch.float32, generator=gen) * args.L
kw = dict(kbins=100, mubins=10, nmesh=args.nmesh, compensated=True, interlaced=True, poles=[0, 2, 4])
host = None
if args.host:
    host = torch.empty((args.N, 3), dtype=torch.float32, pin_memory=True)
    host.copy_(pos)


def measure(src):
    for _ in range(2):
        calc_power(src, args.L, **kw)
    torch.cuda.synchronize()
    eng.profile(True)
    eng.profile_collect()
    times = []
    for _ in range(args.reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        res = calc_power(src, args.L, **kw)
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b))
    prof = eng.profile_collect()
    eng.profile(False)
    stages = ' '.join(f'{k}={v[0] / args.reps:.1f}' for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])[:7])
    return statistics.median(times), stages, float(res['power'][5][0])


best = None
for fused, segs, scatter in itertools.product(('1', '0'), (None, '1', '4'), ('1', '2')):
    os.environ['ABK_FUSED_NORMALIZE'] = fused
    os.environ['ABK_SCATTER'] = scatter
    if segs is None:
        os.environ.pop('ABK_DEVICE_SEGMENTS', None)
    else:
        os.environ['ABK_DEVICE_SEGMENTS'] = segs
    eng.release_scratch()
    ms, stages, p5 = measure(pos)
    print(f'device fused_norm={fused} segments={segs or "default"} scatter={scatter}: {ms:.1f} ms   P[5,0]={p5:.6g}   {stages}', flush=True)
    if best is None or ms < best[0]:
        best = (ms, fused, segs, scatter)
    if host is not None and segs is None and scatter == '1':
        for groups in ('1', '2', '3'):
            os.environ['ABK_EARLY_GROUPS'] = groups
            ms_h, stages, _ = measure(host)
            print(f'host   fused_norm={fused} early_groups={groups}: {ms_h:.1f} ms   {stages}', flush=True)
        os.environ.pop('ABK_EARLY_GROUPS', None)

print(f'best: {best[0]:.1f} ms with fused_norm={best[1]} segments={best[2] or "default"} scatter={best[3]}')
if args.variants:
    os.environ['ABK_FUSED_NORMALIZE'] = best[1]
    os.environ['ABK_SCATTER'] = best[3]
    if best[2]:
        os.environ['ABK_DEVICE_SEGMENTS'] = best[2]
    for variant in ():
        check(eng.lib.abk_ctx_set_tile_capacity(eng.ctx, variant << 16))
        ms, stages, _ = measure(pos)
        print(f'deposit variant {variant}: {ms:.1f} ms   {stages}', flush=True)
    check(eng.lib.abk_ctx_set_tile_capacity(eng.ctx, 1 << 20))
    ms, stages, _ = measure(pos)
    print(f'deposit variant 0 + vector flush (bit 20): {ms:.1f} ms   {stages}', flush=True)
    check(eng.lib.abk_ctx_set_tile_capacity(eng.ctx, 1 << 21))
    ms, stages, _ = measure(pos)
    print(f'deposit variant 0, 64-register build forced (bit 21): {ms:.1f} ms   {stages}', flush=True)
    check(eng.lib.abk_ctx_set_tile_capacity(eng.ctx, 0))
