"""
The real kernel sources of libabk executed through the raw C ABI on the CPU by the fiber-per-CUDA-thread emulator of
tests/emu (TEST INFRASTRUCTURE ONLY): the launch plumbing -- grid/block decomposition, block
scans, ballots, barriers, shared-memory staging, header look-ups, warp-aggregated reductions -- runs unmodified and is
compared with the oracle and the reference's golden arrays.  This is how kernels written without GPU time are checked
before they ever reach a B200; the `-m gpu` tests repeat the comparison on the device.
"""

import ctypes as C
import sys
from pathlib import Path

import numpy as np
import pytest

import cases

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / 'tests' / 'golden'
vp = C.c_void_p


@pytest.fixture(scope='module')
def emu(emu_build_dir):
    import build_emu

    lib = C.CDLL(str(build_emu.build(emu_build_dir)))
    lib.abk_last_error.restype = C.c_char_p
    ctx = vp()
    assert lib.abk_ctx_create(0, C.byref(ctx)) == 0, lib.abk_last_error()

    class Emu:
        pass

    e = Emu()
    e.lib, e.ctx = lib, ctx

    def call(name, *args):
        rc = getattr(lib, name)(*args)
        assert rc == 0, (name, lib.abk_last_error())

    e.call = call
    yield e
    lib.abk_ctx_destroy(ctx)


def P(a):
    return vp(0) if a is None else vp(a.ctypes.data)


def aligned(nbytes, align=256):
    buf = np.zeros(nbytes + align, np.uint8)
    off = (-buf.ctypes.data) % align
    return buf, vp(buf.ctypes.data + off)


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_emu_rvint_and_pids(emu, oracle, dt):
    g = np.load(GOLD / 'ref_ingest.npz')
    f64 = int(dt == np.float64)
    iv = np.ascontiguousarray(g['rvint/in'])
    pos, vel = np.empty(iv.shape, dt), np.empty(iv.shape, dt)
    emu.call('abk_unpack_rvint', emu.ctx, P(iv), C.c_int64(len(iv)), C.c_double(float(g['rvint/box'])), P(pos), P(vel), f64)
    opos, ovel = oracle.unpack_rvint(iv, float(g['rvint/box']), float_dtype=dt)
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)
    packed = np.ascontiguousarray(g['pids/in'])
    n = len(packed)
    out = dict(pid=np.empty(n, np.int64), lagr_pos=np.empty((n, 3), dt), lagr_idx=np.empty((n, 3), np.int16),
               tagged=np.empty(n, np.uint8), density=np.empty(n, dt))
    emu.call('abk_unpack_pids', emu.ctx, P(packed), C.c_int64(n), C.c_double(float(g['pids/box'])), C.c_int64(int(g['pids/ppd'])),
             P(out['pid']), P(out['lagr_pos']), P(out['lagr_idx']), P(out['tagged']), P(out['density']), f64)
    want = oracle.unpack_pids(packed, box=float(g['pids/box']), ppd=float(g['pids/ppd']), float_dtype=dt,
                              **{k: True for k in out})
    for k in out:
        np.testing.assert_array_equal(out[k], want[k], err_msg=k)
        if dt == np.float32:
            np.testing.assert_array_equal(out[k], g[f'pids/{k}'], err_msg=k)


PACK9_CASES = [
    dict(seed=41, nrec=5003, hdr_frac=0.05, first=True),
    dict(seed=42, nrec=1024, hdr_frac=0.4, first=True),        # dense headers, exact multiple of the block
    dict(seed=43, nrec=9001, hdr_frac=0.0005, first=True),     # headers many blocks apart
    dict(seed=44, nrec=3000, hdr_frac=0.01, first=False),      # particles before any header -> NaN
    dict(seed=45, nrec=255, hdr_frac=0.1, first=True),
    dict(seed=46, nrec=1, hdr_frac=0.0, first=True),           # a lone header
]


def run_pack9(emu, d, box, velz, dt):
    nrec = len(d)
    nb = C.c_size_t()
    emu.call('abk_pack9_scratch_bytes', C.c_int64(nrec), C.byref(nb))
    keep, scratch = aligned(nb.value)
    nh = C.c_int64()
    emu.call('abk_pack9_count', emu.ctx, P(d), C.c_int64(nrec), scratch, C.c_size_t(nb.value), C.byref(nh))
    npart = nrec - nh.value
    tab = np.zeros((max(nh.value, 1), 5), dt)
    pos, vel = np.full((npart, 3), -7, dt), np.full((npart, 3), -7, dt)
    emu.call('abk_pack9_decode', emu.ctx, P(d), C.c_int64(nrec), C.c_double(box), C.c_double(velz), scratch, P(tab),
             C.c_int64(nh.value), P(pos), P(vel), int(dt == np.float64))
    del keep
    return pos, vel


@pytest.mark.parametrize('dt', [np.float32, np.float64])
@pytest.mark.parametrize('case', PACK9_CASES, ids=lambda c: f"n{c['nrec']}")
def test_emu_pack9_pipeline(emu, oracle, dt, case):
    d = cases.pack9_inputs(case['seed'], case['nrec'], hdr_frac=case['hdr_frac'], first_header=case['first'])
    pos, vel = run_pack9(emu, d, 2000.0, 1234.5678, dt)
    opos, ovel = oracle.unpack_pack9(d, 2000.0, 1234.5678, float_dtype=dt)
    assert pos.shape == opos.shape
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)


def test_emu_pack9_reference_fixture(emu):
    g = np.load(GOLD / 'ref_ingest.npz')
    d = np.ascontiguousarray(g['pack9/in'][:6000])
    pos, vel = run_pack9(emu, d, float(g['pack9/box']), float(g['pack9/velz']), np.float32)
    np.testing.assert_array_equal(pos, g['pack9/pos'][:len(pos)])
    np.testing.assert_array_equal(vel, g['pack9/vel'][:len(vel)])


@pytest.mark.parametrize('name', ['f16', 'f16_break', 'f15_odd', 'r24', 'f20_f64', 'f40_1bin'])
def test_emu_bin_kppi(emu, name):
    g = np.load(GOLD / 'reference_kppi.npz')
    c = cases.KPPI_CASES[name]
    w, kedges, pimax = cases.kppi_inputs(c)
    dt = np.dtype(c['dtype']).type
    n, Nk, Npi = c['n'], c['Nk'], c['Npi']
    dk = 2 * np.pi / c['L'] if c['fourier'] else c['L'] / n
    ke = ((kedges / dk) ** 2).astype(dt).astype(np.float64)
    pe = ((np.linspace(0.0, pimax, Npi + 1) / dk) ** 2).astype(dt).astype(np.float64)
    w = np.ascontiguousarray(w)
    cnt, sm = np.zeros((Nk, Npi), np.int64), np.zeros((Nk, Npi), np.float64)
    emu.call('abk_bin_kppi', emu.ctx, P(w), int(dt == np.float64), n, C.c_int64(w.shape[2]), P(ke), Nk, P(pe), Npi,
             int(dt == np.float32), P(cnt), P(sm))
    np.testing.assert_array_equal(cnt, g[f'kppi/{name}/counts'])
    nz = cnt != 0
    sm[nz] /= cnt[nz]
    want = g[f'kppi/{name}/mean']
    np.testing.assert_allclose(sm, want, rtol=1e-5, atol=1e-5 * np.abs(want).max())


@pytest.mark.parametrize('name', list(cases.KFIELD_CASES))
def test_emu_kfield_helpers(emu, name):
    from abacusutils_b200.analysis.power_spectrum import legendre_coefficients

    g = np.load(GOLD / 'reference_kfields.npz')
    c = cases.KFIELD_CASES[name]
    n = c['n']
    delta, k_ell, P_ell = cases.kfield_inputs(c)
    delta = np.ascontiguousarray(delta)
    out = np.empty_like(delta)
    emu.call('abk_delta_mu2', emu.ctx, P(delta), P(out), n)
    np.testing.assert_allclose(out, g[f'kf/{name}/delta_mu2'], rtol=1e-6, atol=1e-7)
    sm = np.empty((n, n, n // 2 + 1), np.float32)
    emu.call('abk_smoothing', emu.ctx, P(sm), n, C.c_double(c['L']), C.c_double(c['R']))
    np.testing.assert_allclose(sm, g[f'kf/{name}/smoothing'], rtol=2e-6, atol=1e-30)
    poles = np.asarray(c['poles'], dtype=np.int32)
    coef = legendre_coefficients(poles).astype(np.float64) / (2 * poles[:, None] + 1)   # plain P_l, as the wrapper passes it
    coef = np.ascontiguousarray(coef, dtype=np.float32)
    k32, P32 = np.ascontiguousarray(k_ell, np.float32), np.ascontiguousarray(P_ell, np.float32)
    ex = np.empty((n, n, n // 2 + 1), np.float32)
    emu.call('abk_expand_poles_to_3d', emu.ctx, P(ex), n, C.c_double(c['L']), P(k32), P(P32), len(k32),
             poles.ctypes.data_as(C.POINTER(C.c_int32)), len(poles), P(coef))
    want = g[f'kf/{name}/expand']
    np.testing.assert_allclose(ex, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


@pytest.mark.parametrize('nranks,n', [(1, 12), (2, 12), (3, 10)])
def test_emu_transpose_scatter_p2p(emu, nranks, n):
    """abk_transpose_scatter_p2p (the default FFT transpose at N > 1): every 'rank' is a slab + a pencil buffer on this
    one device, the peer pointers are plain pointers -- the kernel must leave rank r's pencil = full[:, y-range r, :]."""
    rng = np.random.default_rng(n + nranks)
    nzc = n // 2 + 1
    full = (rng.standard_normal((n, n, nzc)) + 1j * rng.standard_normal((n, n, nzc))).astype(np.complex64)
    xs = [r * n // nranks for r in range(nranks)] + [n]
    js = [0] + sorted(rng.choice(np.arange(1, n), size=nranks - 1, replace=False).tolist()) + [n]      # uneven y split
    nyl_max = max(js[r + 1] - js[r] for r in range(nranks))
    pencils = [np.full(n * nyl_max * nzc, np.nan + 0j, np.complex64) for _ in range(nranks)]
    peers = (vp * nranks)(*[p.ctypes.data for p in pencils])
    jsp = (C.c_int64 * (nranks + 1))(*js)
    for r in range(nranks):
        slab = np.ascontiguousarray(full[xs[r]:xs[r + 1]])
        emu.call('abk_transpose_scatter_p2p', emu.ctx, P(slab), peers, C.c_int64(xs[r + 1] - xs[r]), C.c_int64(n), C.c_int64(nzc),
                 nranks, jsp, C.c_int64(xs[r]))
    for r in range(nranks):
        nyl = js[r + 1] - js[r]
        got = pencils[r][: n * nyl * nzc].reshape(n, nyl, nzc)
        np.testing.assert_array_equal(got, full[:, js[r]:js[r + 1], :])


@pytest.mark.parametrize('n,j0,j1,cross', [(12, 3, 9, False), (9, 0, 4, True), (7, 2, 7, False), (16, 0, 16, True)])
def test_emu_pencil_pair_binning_matches_one_mode_kernel(emu, n, j0, j1, cross):
    """power_bin_pair_kernel (pencil layout, +-i' mirror pairs share the bin arithmetic) against power_bin2_kernel (one mode at
    a time) on the same pencil: mode counts identical, sums equal to float32 round-off -- even and odd meshes (the unpaired
    i' = -n/2 or -(n+1)/2 plane), auto and cross spectra, interlaced + compensated."""
    from abacusutils_b200._lib import BinRequest, KMesh
    from abacusutils_b200.analysis.power_spectrum import get_W_compensated, legendre_coefficients

    rng = np.random.default_rng(100 * n + j0)
    nzc, nyl, L = n // 2 + 1, j1 - j0, 100.0
    shape = (n, nyl, nzc)

    def mesh():
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)

    f1, f1s = mesh(), mesh()
    f2, f2s = (mesh(), mesh()) if cross else (None, None)
    W = np.ascontiguousarray(get_W_compensated(L, n, 'TSC', True), np.float32)
    assert np.array_equal(W[1:], W[:0:-1])
    Nk, Nmu, poles = 6, 3, np.array([0, 2, 4], np.int64)
    dk = 2 * np.pi / L
    kedges2 = ((np.linspace(0.0, np.pi * n / L, Nk + 1) / dk) ** 2).astype(np.float32)
    muedges2 = (np.linspace(0.0, 1.0, Nmu + 1) ** 2).astype(np.float32)
    tables = np.concatenate([kedges2, muedges2, legendre_coefficients(poles).reshape(-1)]).astype(np.float32)
    nb = C.c_size_t()
    emu.call('abk_power_bin_scratch_bytes', Nk, Nmu, len(poles), C.byref(nb))
    out = {}
    for sym in (1, 0):
        sums = np.zeros(3 * Nk * Nmu + len(poles) * Nk, np.float64)
        scratch, sptr = aligned(nb.value)
        req = BinRequest()
        req.mesh = KMesh(n=n, nzc=nzc, i0=0, i1=n, j0=j0, j1=j1, stride_i=nyl * nzc, stride_j=nzc)
        req.f1, req.f1s = f1.ctypes.data, f1s.ctypes.data
        req.f2, req.f2s = (f2.ctypes.data, f2s.ctypes.data) if cross else (None, None)
        req.real_in, req.W = None, W.ctypes.data
        req.scale, req.finish, req.w_symmetric = 0.5 / n**3, 1, sym
        base = tables.ctypes.data
        req.kedges2, req.muedges2, req.pole_coef = base, base + 4 * (Nk + 1), base + 4 * (Nk + 1 + Nmu + 1)
        for ip, ell in enumerate(poles):
            req.pole_ell[ip] = int(ell)
        req.Nk, req.Nmu, req.Np = Nk, Nmu, len(poles)
        sb = sums.ctypes.data
        req.counts, req.sum_p, req.sum_k, req.sum_poles = sb, sb + 8 * Nk * Nmu, sb + 16 * Nk * Nmu, sb + 24 * Nk * Nmu
        req.scratch, req.scratch_bytes = sptr.value, nb.value
        emu.call('abk_power_bin', emu.ctx, C.byref(req))
        emu.call('abk_ctx_sync', emu.ctx)
        out[sym] = sums
    nbin = Nk * Nmu
    np.testing.assert_array_equal(out[1][:nbin].view(np.int64), out[0][:nbin].view(np.int64))
    assert out[0][:nbin].view(np.int64).sum() > 0
    scale = np.abs(out[0][nbin:]).max()
    np.testing.assert_allclose(out[1][nbin:], out[0][nbin:], rtol=2e-5, atol=2e-6 * scale)
