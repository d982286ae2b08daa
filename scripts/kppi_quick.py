"""Torch-free check of abk_bin_kppi against tests/golden/reference_kppi.npz (ctypes + libcudart only)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

t0 = time.time()
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
import cases  # noqa: E402

rt = C.CDLL('/usr/local/cuda/lib64/libcudart.so')
lib = C.CDLL(str(ROOT / 'abacusutils_b200' / 'libabk.so'))
lib.abk_last_error.restype = C.c_char_p
vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
lib.abk_bin_kppi.argtypes = [vp, vp, i32, i32, i64, vp, i32, vp, i32, i32, vp, vp]
rt.cudaMalloc.argtypes = [C.POINTER(vp), C.c_size_t]
rt.cudaMemcpy.argtypes = [vp, vp, C.c_size_t, i32]
rt.cudaMemset.argtypes = [vp, i32, C.c_size_t]


def dev(a):
    p = vp()
    assert rt.cudaMalloc(C.byref(p), max(a.nbytes, 16)) == 0
    assert rt.cudaMemcpy(p, a.ctypes.data_as(vp), a.nbytes, 1) == 0
    return p


def host(p, shape, dt):
    a = np.empty(shape, dtype=dt)
    assert rt.cudaMemcpy(a.ctypes.data_as(vp), p, a.nbytes, 2) == 0
    return a


ctx = vp()
assert lib.abk_ctx_create(0, C.byref(ctx)) == 0, lib.abk_last_error()
g = np.load(ROOT / 'tests' / 'golden' / 'reference_kppi.npz')
bad = 0
for name, c in cases.KPPI_CASES.items():
    w, kedges, pimax = cases.kppi_inputs(c)
    dt = np.dtype(c['dtype']).type
    n, Nk, Npi = c['n'], c['Nk'], c['Npi']
    dk = 2 * np.pi / c['L'] if c['fourier'] else c['L'] / n
    ke = ((kedges / dk) ** 2).astype(dt).astype(np.float64)
    pe = ((np.linspace(0.0, pimax, Npi + 1) / dk) ** 2).astype(dt).astype(np.float64)
    dw, dke, dpe = dev(np.ascontiguousarray(w)), dev(ke), dev(pe)
    dc, ds = dev(np.zeros(Nk * Npi, dtype=np.uint64)), dev(np.zeros(Nk * Npi, dtype=np.float64))
    rc = lib.abk_bin_kppi(ctx, dw, int(dt is np.float64), n, w.shape[2], dke, Nk, dpe, Npi, int(dt is np.float32), dc, ds)
    assert rc == 0, lib.abk_last_error()
    assert lib.abk_ctx_sync(ctx) == 0, lib.abk_last_error()
    cnt = host(dc, (Nk, Npi), np.int64)
    s = host(ds, (Nk, Npi), np.float64)
    nz = cnt != 0
    s[nz] /= cnt[nz]
    want, wc = g[f'kppi/{name}/mean'], g[f'kppi/{name}/counts']
    ok_c = np.array_equal(cnt, wc)
    err = np.abs(s - want).max() / np.abs(want).max()
    print(name, 'counts', 'OK' if ok_c else 'MISMATCH', 'rel err %.2e' % err, flush=True)
    bad += (not ok_c) or err > 1e-4
print('elapsed %.1fs' % (time.time() - t0), 'FAIL' if bad else 'ALL OK')
sys.exit(1 if bad else 0)
