"""
Golden outputs of the UNMODIFIED reference for float64 inputs (run in the BUILD CONTAINER only):
tsc_parallel with every combination of float32 / float64 positions, weights and grid (tsc.py:155-165, :400: arithmetic in
the dtype of the positions, accumulation in the dtype of the grid) and calc_power / get_field_fft with dtype=float64 on the
non-interlaced branch (power_spectrum.py:1053-1069).  Inputs are regenerated from seeds by cases.py; outputs go to
tests/golden/reference_f64.npz.

Usage:  NUMBA_NUM_THREADS=2 python tests/golden/make_golden_f64.py
"""

import os
import sys
import warnings
from pathlib import Path

os.environ.setdefault('NUMBA_NUM_THREADS', '2')
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import numpy as np  # noqa: E402

import cases  # noqa: E402
from oracle import ref_shim  # noqa: E402


def main():
    tsc, ps = ref_shim.load(num_threads=2)
    out = {}
    pos, w = cases.f64_inputs()
    box, shape = 77.0, (20, 24, 28)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for pd, gd, wd in (('f8', 'f8', 'f8'), ('f4', 'f8', None), ('f8', 'f4', 'f4'), ('f4', 'f8', 'f8')):
            p = pos.astype(pd)
            grid = np.zeros(shape, dtype=gd)
            tsc.tsc_parallel(p, grid, box, weights=None if wd is None else w.astype(wd), nthread=2, offset=0.125)
            out[f'tsc/{pd}_{gd}_{wd}'] = grid
            if (pd, gd) == ('f8', 'f8'):
                out['tsc/wrapped_pos_f8'] = p      # _wrap_inplace in float64: only out-of-range entries change
        c = cases.F64_POWER
        rng = np.random.default_rng(c['seed'])
        p32 = rng.random((c['N'], 3), dtype='f4') * np.float32(c['L'])
        w32 = rng.random(c['N'], dtype='f4')
        p2 = rng.random((c['N'] // 2, 3), dtype='f4') * np.float32(c['L'])
        kw = dict(kbins=c['kbins'], mubins=c['mubins'], nmesh=c['nmesh'], compensated=True, interlaced=False, poles=c['poles'],
                  dtype=np.float64, nthread=2)
        for name, args in (('auto_w', dict(w=w32)), ('cross', dict(pos2=p2))):
            t = ps.calc_power(p32.copy(), c['L'], **kw, **args)
            for k in t.keys():
                out[f'power/{name}/{k}'] = np.asarray(t[k])
        W = ps.get_W_compensated(c['L'], c['nmesh'], 'TSC', False)
        f = ps.get_field_fft(p32.copy(), c['L'], c['nmesh'], 'TSC', w32, W, True, False, nthread=2, dtype=np.float64)
        out['field/f64'] = f
        p64 = (rng.random((5000, 3)) * c['L'])
        f = ps.get_field_fft(p64.copy(), c['L'], 16, 'TSC', None, None, False, False, nthread=2, dtype=np.float64)
        out['field/f64_pos64'] = f
    np.savez_compressed(HERE / 'reference_f64.npz', **out)
    for k, v in out.items():
        print(k, v.dtype, v.shape)


if __name__ == '__main__':
    main()
