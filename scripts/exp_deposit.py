"""Experiment driver (GPU box): per-kernel timings of bucketing + deposit for one particle set."""
import argparse
import ctypes as C
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

from abacusutils_b200._lib import Engine, check, ptr
from abacusutils_b200.analysis.power_spectrum import _Painter

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=512)
ap.add_argument('--N', type=int, default=100_000_000)
ap.add_argument('--cap', type=int, default=0)
ap.add_argument('--reps', type=int, default=3)
args = ap.parse_args()

eng = Engine.get()
L = 1000.0
g = torch.Generator(device='cuda'); g.manual_seed(1)
pos = torch.rand((args.N, 3), device='cuda', generator=g) * L
if args.cap:
    check(eng.lib.abk_ctx_set_tile_capacity(eng.ctx, args.cap))
P = _Painter(eng, args.n, L)
d = 0.5 * L / args.n
for offs in ([0.0], [d], [0.0, d]):
    for _ in range(2):
        P.paint(pos, None, offs)
    torch.cuda.synchronize()
    eng.profile(True); eng.profile_collect()
    for _ in range(args.reps):
        grids = P.paint(pos, None, offs)
    prof = eng.profile_collect(); eng.profile(False)
    print('offsets', offs, {k: round(v[0] / args.reps, 3) for k, v in prof.items()},
          'sum', [float(x.sum(dtype=torch.float64)) for x in grids])
