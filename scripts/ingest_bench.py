"""Kernel-level timing of the device-side particle decoders (GPU box): abk_unpack_rvint, abk_pack9_count + abk_pack9_decode,
with their algorithmic bytes against the measured HBM peak (MEASURED_PEAKS.json).   python scripts/ingest_bench.py [--n 200000000]"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch

from abacusutils_b200._lib import Engine
from abacusutils_b200.data import bitpacked, pack9

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=200_000_000)
args = ap.parse_args()
N = args.n
eng = Engine.get(0)
try:
    peak = float(json.loads((ROOT / 'MEASURED_PEAKS.json').read_text())['hbm_gbs'])
except Exception:
    peak = 6650.0
gen = torch.Generator(device='cuda')
gen.manual_seed(5)


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
rv = (torch.randint(-500000, 500000, (N, 3), device='cuda', dtype=torch.int32, generator=gen) << 12) | \
    torch.randint(0, 4096, (N, 3), device='cuda', dtype=torch.int32, generator=gen)
ms = timeit(lambda: bitpacked.unpack_rvint(rv, 1000.0, velout=False))
out['unpack_rvint (pos only, f32)'] = {'ms': ms, 'algorithmic_bytes': N * 24, 'gbs': N * 24 / ms / 1e6, 'frac_of_hbm': N * 24 / ms / 1e6 / peak}
ms = timeit(lambda: bitpacked.unpack_rvint(rv, 1000.0))
out['unpack_rvint (pos + vel, f32)'] = {'ms': ms, 'algorithmic_bytes': N * 36, 'gbs': N * 36 / ms / 1e6, 'frac_of_hbm': N * 36 / ms / 1e6 / peak}
del rv
raw = torch.randint(0, 255, (N, 9), device='cuda', dtype=torch.uint8, generator=gen)
h = torch.arange(0, N, 32, device='cuda')
f = torch.stack([torch.zeros_like(h), torch.full_like(h, 1000 - 2000 + 2048), torch.full_like(h, 2048)] +
                [torch.randint(-2000 + 2048, 1000 - 2000 + 2048, h.shape, device='cuda', generator=gen) for _ in range(3)], 1)
hb = torch.empty((len(h), 9), dtype=torch.uint8, device='cuda')
for q in range(3):
    a_, b_ = f[:, 2 * q], f[:, 2 * q + 1]
    hb[:, 3 * q] = (a_ >> 4) & 0xFF
    hb[:, 3 * q + 1] = ((a_ & 0xF) | (((b_ >> 8) & 0xF) << 4)).to(torch.uint8)
    hb[:, 3 * q + 2] = (b_ & 0xFF).to(torch.uint8)
hb[:, 0] = 0xFF
raw[h] = hb
ms = timeit(lambda: pack9.unpack_pack9(raw, 1000.0, 1.0, velout=False))
np_ = N - len(h)
out['unpack_pack9 (count + scan + decode, pos only)'] = {'ms': ms, 'algorithmic_bytes': N * 9 * 2 + np_ * 12, 'gbs': (N * 18 + np_ * 12) / ms / 1e6,
                                                         'frac_of_hbm': (N * 18 + np_ * 12) / ms / 1e6 / peak}
print(json.dumps({'n_records': N, 'hbm_peak_gbs': peak, 'kernels': out}, indent=1))
