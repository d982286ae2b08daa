"""
The product's Python layer on top of the real kernel sources, end to end on the CPU (tests/emu/emu_engine.py):
tsc_parallel, calc_power (auto/cross, interlaced, compensated, log-k), get_field_fft, calc_pk_from_deltak, pk_to_xi,
CIC, partition_parallel, 2-D grids and the ingest entry points, compared with the outputs of the unmodified reference
(tests/golden/*.npz) exactly like the `-m gpu` tests do on the device.  Only cuFFT is substituted (scipy).
"""

import numpy as np
import pytest

import cases
from common import assert_int_exact, compare_power_tables


@pytest.fixture
def emu(monkeypatch, emu_build_dir):
    import emu_engine

    return emu_engine.install(monkeypatch, emu_build_dir)


@pytest.mark.parametrize('name', ['cube24_w_off', 'aniso', 'unwrapped', 'tiny10'])
def test_tsc_parallel(emu, golden, name):
    from abacusutils_b200.analysis import tsc

    c = cases.TSC_CASES[name]
    pos, w = cases.tsc_inputs(c)
    dens = np.zeros(c['shape'], dtype=np.float32)
    assert tsc.tsc_parallel(pos, dens, c['box'], weights=w, offset=c['offset'], wrap=c.get('wrap', True)) is None
    np.testing.assert_allclose(dens, golden[f'tsc/{name}'], rtol=1e-4, atol=1e-5)
    out = tsc.tsc_parallel(pos, c['shape'], c['box'], weights=w, offset=c['offset'], wrap=c.get('wrap', True))
    np.testing.assert_allclose(out, golden[f'tsc/{name}'], rtol=1e-4, atol=1e-5)


def test_tile_capacity_passes_and_variants(emu, oracle):
    """Clustered input that overflows the per-tile record capacity (several passes per tile) against the oracle."""
    from abacusutils_b200._lib import check
    from abacusutils_b200.analysis import tsc

    rng = np.random.default_rng(3)
    box, n = 50.0, 20
    pos = np.concatenate([rng.normal(25.0, 1.0, size=(6000, 3)), rng.random((1500, 3)) * box]).astype(np.float32) % np.float32(box)
    w = rng.random(len(pos), dtype=np.float32)
    want = np.zeros((n, n, n), np.float32)
    oracle.tsc_parallel(pos.copy(), want, box, weights=w, nthread=1)
    try:
        for cap in (256, 512):      # 7500 particles in a handful of tiles: many passes per tile
            check(emu.lib.abk_ctx_set_tile_capacity(emu.ctx, cap))
            got = tsc.tsc_parallel(pos.copy(), (n, n, n), box, weights=w)
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5, err_msg=f'capacity {cap}')
    finally:
        check(emu.lib.abk_ctx_set_tile_capacity(emu.ctx, 0))


@pytest.mark.parametrize('name', ['n32_ci', 'n32_cross_ci', 'n32_raw', 'n36_nopoles_mu'])
def test_calc_power(emu, golden, name):
    from abacusutils_b200.analysis import power_spectrum as ps

    c = cases.POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                      nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'], w=w, pos2=pos2, w2=w2,
                      poles=c['poles'])
    want = {k[len(f'power/{name}/'):]: golden[k] for k in golden.files if k.startswith(f'power/{name}/')}
    compare_power_tables(t, want)


def test_calc_power_cic(emu):
    from abacusutils_b200.analysis import power_spectrum as ps

    g = np.load(cases.__file__.replace('cases.py', 'reference_cic.npz'))
    name = next(iter(cases.CIC_POWER_CASES))
    c = cases.CIC_POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], logk=c['logk'], paste='CIC', nmesh=c['nmesh'],
                      compensated=c['compensated'], interlaced=c['interlaced'], w=w, pos2=pos2, w2=w2, poles=c['poles'])
    want = {k[len(f'power/{name}/'):]: g[k] for k in g.files if k.startswith(f'power/{name}/')}
    compare_power_tables(t, want)


def test_field_fft_and_deltak_binning(emu, golden):
    from abacusutils_b200.analysis import power_spectrum as ps

    name = next(n for n, c in cases.FIELD_CASES.items() if c['interlaced'] and c['compensated'])
    c = cases.FIELD_CASES[name]
    pos, w, _, _ = cases.power_inputs(c)
    W = ps.get_W_compensated(c['L'], c['nmesh'], 'TSC', True)
    f = ps.get_field_fft(pos, c['L'], c['nmesh'], 'TSC', w, W, True, True)
    want = golden[f'field/{name}']
    assert f.dtype == want.dtype and f.shape == want.shape
    assert np.abs(f - want).max() <= 2e-5 * np.sqrt(np.mean(np.abs(want) ** 2))
    name = next(iter(cases.DELTAK_CASES))
    c = cases.DELTAK_CASES[name]
    f1, f2, raw = cases.deltak_inputs(c)
    kedges, muedges = cases.count_edges(c)
    poles = np.asarray(c['poles'], dtype=np.int64)
    P = ps.calc_pk_from_deltak(f1, c['L'], kedges, muedges, field2_fft=f2, poles=poles)
    assert_int_exact(P['N_mode'], golden[f'deltak/{name}/N_mode'], 'N_mode')
    np.testing.assert_allclose(P['power'], golden[f'deltak/{name}/power'], rtol=1e-4,
                               atol=1e-5 * np.abs(golden[f'deltak/{name}/power']).max())


def test_pk_to_xi(emu):
    from abacusutils_b200.analysis import power_spectrum as ps

    g = np.load(cases.__file__.replace('cases.py', 'reference_xi.npz'))
    name = next(iter(cases.XI_CASES))
    c = cases.XI_CASES[name]
    Pk, r_bins = cases.xi_inputs(c)
    r_binc, binned_poles, Npoles = ps.pk_to_xi(Pk.copy(), c['L'], r_bins, poles=c['poles'])
    assert_int_exact(Npoles, g[f'xi/{name}/Npoles'], 'Npoles')
    want = g[f'xi/{name}/binned_poles']
    np.testing.assert_allclose(binned_poles, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


def test_partition_parallel(emu, golden):
    from abacusutils_b200.analysis import tsc

    name = next(iter(cases.PARTITION_CASES))
    c = cases.PARTITION_CASES[name]
    pos, w = cases.tsc_inputs(c)
    ppart, starts, wpart = tsc.partition_parallel(pos, c['npartition'], c['box'], weights=w)
    np.testing.assert_array_equal(starts, golden[f'partition/{name}/starts'])
    ref = golden[f'partition/{name}/ppart']
    for i in range(c['npartition']):                       # same multiset per stripe
        a, b = starts[i], starts[i + 1]
        np.testing.assert_array_equal(np.sort(ppart[a:b].view('f4,f4,f4'), axis=0), np.sort(ref[a:b].view('f4,f4,f4'), axis=0))
    # and the same ORDER as the unmodified reference: its output is the stable partition, restored from the source indices
    np.testing.assert_array_equal(ppart, ref)
    if f'partition/{name}/wpart' in golden.files:
        np.testing.assert_array_equal(wpart, golden[f'partition/{name}/wpart'])


def test_ingest_through_the_product_wrappers(emu, oracle):
    from abacusutils_b200.data import bitpacked, pack9

    g = np.load(cases.__file__.replace('cases.py', 'ref_ingest.npz'))
    pos, vel = pack9.unpack_pack9(g['pack9/in'][:5000], float(g['pack9/box']), float(g['pack9/velz']))
    np.testing.assert_array_equal(pos, g['pack9/pos'][:len(pos)])
    np.testing.assert_array_equal(vel, g['pack9/vel'][:len(vel)])
    pos, vel = bitpacked.unpack_rvint(g['rvint/in'], float(g['rvint/box']))
    np.testing.assert_array_equal(pos, g['rvint/pos'])
    got = bitpacked.unpack_pids(g['pids/in'], box=float(g['pids/box']), ppd=float(g['pids/ppd']), pid=True, lagr_pos=True,
                                density=True)
    for k in got:
        np.testing.assert_array_equal(got[k], g[f'pids/{k}'], err_msg=k)


@pytest.mark.parametrize('groups', ['1', '2', '3'])
def test_multi_segment_early_deposit_groups(emu, golden, monkeypatch, groups):
    """Host input cut into 14 chunks (one bucket segment each) with 1-3 early deposit groups on the auxiliary
    stream: the orchestration of abacusutils_b200.analysis.power_spectrum._Painter, result unchanged."""
    from abacusutils_b200.analysis import power_spectrum as ps

    monkeypatch.setenv('ABK_CHUNK_MIN', '64')
    monkeypatch.setenv('ABK_EARLY_GROUPS', groups)
    name = 'n32_ci'
    c = cases.POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    P = ps._Painter.__new__(ps._Painter)
    assert len(P.chunk_plan(len(pos))) == 14
    t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                      nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'], w=w, pos2=pos2, w2=w2,
                      poles=c['poles'])
    want = {k[len(f'power/{name}/'):]: golden[k] for k in golden.files if k.startswith(f'power/{name}/')}
    compare_power_tables(t, want)


def test_ragged_and_odd_grids(emu, golden, oracle):
    """Padded FFT grids (32 | n or not) and grids with odd row strides / a partial last z tile: rows that wrap in z inside a
    warp, z-halo cells that wrap, meshes smaller than a tile."""
    from abacusutils_b200.analysis import power_spectrum as ps
    from abacusutils_b200.analysis import tsc

    for name in ('n32_ci', 'n48_log'):
        c = cases.POWER_CASES[name]
        pos, w, pos2, w2 = cases.power_inputs(c)
        t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                          nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'], w=w, pos2=pos2,
                          w2=w2, poles=c['poles'])
        want = {k[len(f'power/{name}/'):]: golden[k] for k in golden.files if k.startswith(f'power/{name}/')}
        compare_power_tables(t, want)
    rng = np.random.default_rng(4)
    for shape in ((16, 16, 64), (9, 11, 33), (8, 8, 32), (12, 10, 70), (5, 40, 3), (33, 7, 2)):
        pos = (rng.random((3000, 3), dtype=np.float32) * np.float32(80.0)).astype(np.float32)
        got, want = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
        tsc.tsc_parallel(pos.copy(), got, 80.0, offset=0.3)
        oracle.tsc_parallel(pos.copy(), want, 80.0, offset=0.3, nthread=1)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5, err_msg=str(shape))


@pytest.mark.parametrize('kind', ['pack9', 'rvint'])
def test_calc_power_from_packed_records(emu, oracle, monkeypatch, kind):
    """calc_power(PackedParticles(...)): records cross to the device packed, are decoded there chunk by chunk (pack9
    chunks cut at cell headers) and painted; same table as decoding on the host first."""
    from abacusutils_b200.analysis import power_spectrum as ps
    from abacusutils_b200.data.packed import PackedParticles

    monkeypatch.setenv('ABK_CHUNK_MIN', '700')          # ~14 chunks
    L = 1000.0
    if kind == 'pack9':
        raw = cases.pack9_inputs(77, 9000, hdr_frac=0.03, cpd=405)
        pos, _ = oracle.unpack_pack9(raw, L, 1.0)
        src = PackedParticles(raw, L, velzspace_to_kms=1.0)
        assert len(src) == 9000 and src.n_particles is None
        plan = src.chunk_plan(14, 700)
        assert plan[0][0] == 0 and plan[-1][1] == 9000 and all(raw[a, 0] == 0xFF for a, _ in plan)
    else:
        rng = np.random.default_rng(78)
        raw = ((rng.integers(-500000, 500000, size=(8000, 3)).astype(np.int32) << 12) | rng.integers(0, 4096, size=(8000, 3)).astype(np.int32))
        pos, _ = oracle.unpack_rvint(raw, L)
        src = PackedParticles(raw, L)
    kw = dict(kbins=12, mubins=3, nmesh=24, poles=[0, 2, 4])
    got = ps.calc_power(src, L, **kw)
    want = ps.calc_power(pos.copy(), L, **kw)
    assert got.meta['N_pos'] == len(pos) == src.n_particles
    assert_int_exact(got['N_mode'], want['N_mode'], 'N_mode')
    np.testing.assert_allclose(got['power'], want['power'], rtol=2e-5, atol=2e-6 * np.abs(np.asarray(want['power'])).max())
    np.testing.assert_allclose(got['poles'], want['poles'], rtol=1e-4, atol=2e-5 * np.abs(np.asarray(want['poles'])).max())
    # a second call on the same object takes the fused-normalisation path (the particle count is known by now): it must
    # normalise by the particle count, not by the record count (pack9 records include cell headers)
    again = ps.calc_power(src, L, **kw)
    assert again.meta['N_pos'] == len(pos)
    assert_int_exact(again['N_mode'], want['N_mode'], 'N_mode (second call)')
    np.testing.assert_allclose(again['power'], got['power'], rtol=2e-5, atol=2e-6 * np.abs(np.asarray(want['power'])).max())
    np.testing.assert_allclose(again['poles'], got['poles'], rtol=1e-4, atol=2e-5 * np.abs(np.asarray(want['poles'])).max())
    with pytest.raises(ValueError):
        ps.calc_power(src, L, w=np.ones(len(src), np.float32), **kw)


def test_cic_serial(emu):
    """abacusutils_b200.analysis.cic.cic_serial against get_field(paste='CIC') of the unmodified reference."""
    from abacusutils_b200.analysis.cic import cic_serial

    g = np.load(cases.__file__.replace('cases.py', 'reference_cic.npz'))
    for name, c in cases.CIC_FIELD_CASES.items():
        if c['d'] != 0.0:
            continue
        pos, w = cases.cic_field_inputs(c)
        n = c['nmesh']
        dens = np.zeros((n, n, n), dtype=np.float32)
        assert cic_serial(pos, dens, c['L'], weights=w) is None
        field = dens * np.float32(n**3 / len(pos)) - 1       # get_field normalises by len(pos) (power_spectrum.py:856)
        np.testing.assert_allclose(field, g[f'field/{name}'], rtol=1e-4, atol=1e-5)


def test_wrap_writes_back_only_changed_entries_in_callers_dtype(emu):
    """tsc_parallel(wrap=True) mutates the caller's positions like _wrap_inplace (tsc.py:219-226): only out-of-range
    entries change, in the array's own dtype -- a float64 catalogue is not rounded through float32 (advisor finding)."""
    from abacusutils_b200.analysis import tsc

    rng = np.random.default_rng(8)
    box = 100.0
    with pytest.warns(UserWarning):
        pos = rng.random((500, 3)) * box                      # float64, values not representable in float32
        pos[7, 1] += box
        pos[11, 2] -= box
        before = pos.copy()
        tsc.tsc_parallel(pos, 16, box)
    changed = pos != before
    assert changed.sum() == 2 and changed[7, 1] and changed[11, 2]
    assert pos[7, 1] == before[7, 1] - box and pos[11, 2] == before[11, 2] + box
    # float32 input: wrapped values are float32(float64(v) -+ box), untouched entries bit-identical
    p32 = (rng.random((300, 3)) * box).astype(np.float32)
    p32[3, 0] += np.float32(box)
    b32 = p32.copy()
    tsc.tsc_parallel(p32, 16, box)
    assert (p32 != b32).sum() == 1 and p32[3, 0] == np.float32(np.float64(b32[3, 0]) - box)
    # 2-D grid with an (N, 3) array: the unused third column is wrapped too, nothing else moves
    p2 = (rng.random((200, 3)) * box).astype(np.float32)
    p2[5, 2] = np.float32(-3.0)
    b2 = p2.copy()
    tsc.tsc_parallel(p2, (12, 12), box)
    assert (p2 != b2).sum() == 1 and p2[5, 2] == np.float32(97.0)


@pytest.mark.parametrize('pd,gd,wd', [('f8', 'f8', 'f8'), ('f4', 'f8', None), ('f8', 'f4', 'f4'), ('f4', 'f8', 'f8')])
def test_tsc_parallel_float64(emu, pd, gd, wd):
    import f64_checks

    f64_checks.check_tsc_parallel(pd, gd, wd)


@pytest.mark.parametrize('name', ['auto_w', 'cross'])
def test_calc_power_float64(emu, name):
    import f64_checks

    f64_checks.check_calc_power(name)


def test_shift_field_fft_any_shift(emu, oracle):
    """shift_field_fft (power_spectrum.py:904-948) for the interlacing shift and for arbitrary d, against the oracle and --
    where the live reference is importable -- against the reference itself."""
    from abacusutils_b200.analysis import power_spectrum as ps
    from oracle import ref_shim

    c = cases.DELTAK_CASES['d32']
    f1, f2, _ = cases.deltak_inputs(c)
    ref = ref_shim.load(2)[1] if ref_shim.available() else None
    for d in (c['L'] / c['n'], 0.37 * c['L'] / c['n'], 2.5 * c['L'] / c['n'], -0.8):
        a, b = f1.copy(), f1.copy()
        ps.shift_field_fft(a, f2, c['n'], c['L'], d)
        oracle.shift_field_fft(b, f2, c['n'], c['L'], d, nthread=2)
        assert np.abs(a - b).max() < 2e-6 * np.abs(b).max(), d
        if ref is not None:
            r = f1.copy()
            ref.shift_field_fft(r, f2, c['n'], c['L'], d)
            assert np.abs(a - r).max() < 2e-6 * np.abs(r).max(), d
