"""
GPU parity of the power-spectrum chain (FFT via cuFFT, abk_kspace.cu) against outputs of the
unmodified reference (tests/golden/reference_runs.npz) and the CPU oracle:
  * N_mode / N_mode_poles bit-exact (north_star), incl. the all-ones mode-count cases;
  * P(k,mu), multipoles, k_avg within relative 1e-4 per bin, with the amplitude-scaled absolute
    term of tests/common.py where a quantity crosses zero; the DC-only bin is skipped (rounding
    noise in the reference as well);
  * delta(k) of get_field_fft within 2e-5 of the rms mode amplitude;
  * the reference's own identity test (tests/test_power.py:58-61).
"""

import numpy as np
import pytest

import cases
from common import assert_close_scaled, assert_int_exact, compare_power_tables

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ps():
    from abacusutils_b200.analysis import power_spectrum as m

    return m


@pytest.mark.parametrize('name', list(cases.COUNT_CASES))
def test_mode_counts_bit_exact(ps, golden, name):
    c = cases.COUNT_CASES[name]
    n = c['n']
    kedges, muedges = cases.count_edges(c)
    ones = np.ones((n, n, n // 2 + 1), dtype=np.float32)
    poles = np.asarray(c['poles'], dtype=np.int64)
    wc, cnt, wcp, cntp, wk = ps.bin_kmu(n, c['L'], kedges, muedges, ones, poles=poles)
    assert cnt.dtype == np.int64 and cntp.dtype == np.int64 and wc.dtype == np.float32
    assert wcp.shape == (len(poles), c['Nk'])
    assert_int_exact(cnt, golden[f'counts/{name}/N_mode'], 'N_mode')
    assert_int_exact(cntp, golden[f'counts/{name}/N_mode_poles'], 'N_mode_poles')
    assert_close_scaled(wk, golden[f'counts/{name}/k_avg'], rtol=2e-5, what='k_avg')
    assert_close_scaled(wc, golden[f'counts/{name}/power'], rtol=2e-5, what='power')
    assert_close_scaled(wcp, golden[f'counts/{name}/poles'], scale=1.0, rtol=1e-4, what='poles')


@pytest.mark.parametrize('n,Nk,Nmu', [(256, 100, 10), (200, 200, 1), (96, 2000, 30)])
def test_mode_counts_vs_oracle_large(ps, oracle, n, Nk, Nmu):
    """Bigger meshes / bin tables (incl. one that does not fit the shared-memory tables) vs the oracle."""
    L = 1000.0
    kedges = np.linspace(0, np.pi * n / L, Nk + 1)
    muedges = np.linspace(0, 1, Nmu + 1)
    rng = np.random.default_rng(n)
    wts = rng.random((n, n, n // 2 + 1), dtype='f4')
    poles = np.array([0, 2, 4])
    got = ps.bin_kmu(n, L, kedges, muedges, wts, poles=poles)
    want = oracle.bin_kmu(n, L, kedges, muedges, wts, poles=poles, acc64=True)
    assert_int_exact(got[1], want[1])
    assert_int_exact(got[3], want[3])
    assert_close_scaled(got[0], want[0], rtol=1e-5, what='power')
    assert_close_scaled(got[4], want[4], rtol=1e-5, what='k_avg')
    assert_close_scaled(got[2], want[2], scale=0.5, rtol=1e-5, what='poles')


@pytest.mark.parametrize('name', list(cases.DELTAK_CASES))
def test_deltak_binning(ps, golden, name):
    c = cases.DELTAK_CASES[name]
    f1, f2, raw = cases.deltak_inputs(c)
    kedges, muedges = cases.count_edges(c)
    poles = np.asarray(c['poles'], dtype=np.int64)
    f1c, f2c = f1.copy(), f2.copy()
    P = ps.calc_pk_from_deltak(f1, c['L'], kedges, muedges, field2_fft=f2, poles=poles)
    assert np.array_equal(f1, f1c) and np.array_equal(f2, f2c)  # pure
    pre = f'deltak/{name}/'
    for key in ('power', 'N_mode', 'binned_poles', 'N_mode_poles', 'k_avg'):
        assert P[key].shape == golden[pre + key].shape and P[key].dtype == golden[pre + key].dtype, key
    assert_int_exact(P['N_mode'], golden[pre + 'N_mode'])
    assert_int_exact(P['N_mode_poles'], golden[pre + 'N_mode_poles'])
    scale = np.abs(golden[pre + 'binned_poles'][0]) if 0 in c['poles'] else 1.0
    assert_close_scaled(P['power'], golden[pre + 'power'], rtol=1e-4, what='power')
    assert_close_scaled(P['k_avg'], golden[pre + 'k_avg'], rtol=1e-4, what='k_avg')
    assert_close_scaled(P['binned_poles'], golden[pre + 'binned_poles'], scale=scale, rtol=1e-4, what='poles')
    bp, Npo = ps.project_3d_to_poles(kedges, raw, c['L'], poles)
    assert_int_exact(Npo, golden[pre + 'proj_N'])
    assert_close_scaled(bp, golden[pre + 'proj_poles'], scale=np.abs(golden[pre + 'proj_poles']).max(), rtol=1e-4,
                        what='proj_poles')
    with pytest.raises(AssertionError):
        ps.project_3d_to_poles(kedges, raw, c['L'], [0, 12])


def test_auto_power_of_deltak_and_raw_power(ps, oracle):
    c = cases.DELTAK_CASES['d32']
    f1, f2, _ = cases.deltak_inputs(c)
    np.testing.assert_allclose(ps.get_raw_power(f1), np.abs(f1) ** 2, rtol=1e-6)
    np.testing.assert_allclose(ps.get_raw_power(f1, f2), (np.conj(f1) * f2).real, rtol=1e-5, atol=1e-6)
    kedges, muedges = cases.count_edges(c)
    got = ps.calc_pk_from_deltak(f1, c['L'], kedges, muedges, poles=np.array([0, 2]))
    want = oracle.calc_pk_from_deltak(f1, c['L'], kedges, muedges, poles=np.array([0, 2]), acc64=True)
    assert_int_exact(got['N_mode'], want['N_mode'])
    assert_close_scaled(got['power'], want['power'], rtol=1e-5)
    assert_close_scaled(got['binned_poles'], want['binned_poles'], scale=np.abs(want['binned_poles'][0]), rtol=1e-5)


@pytest.mark.parametrize('name', list(cases.FIELD_CASES))
def test_field_fft(ps, golden, name):
    c = cases.FIELD_CASES[name]
    pos, w, _, _ = cases.power_inputs(c)
    W = ps.get_W_compensated(c['L'], c['nmesh'], 'TSC', c['interlaced']) if c['compensated'] else None
    f = ps.get_field_fft(pos, c['L'], c['nmesh'], 'TSC', w, W, c['compensated'], c['interlaced'])
    want = golden[f'field/{name}']
    assert f.dtype == want.dtype and f.shape == want.shape
    rms = np.sqrt((np.abs(want[1:]) ** 2).mean())
    assert np.abs(f - want).max() < 2e-5 * rms + 1e-6 * np.abs(want).max()


def test_get_field_and_normalize(ps, oracle):
    c = cases.FIELD_CASES['f24_c']
    pos, w, _, _ = cases.power_inputs(c)
    got = ps.get_field(pos.copy(), c['L'], c['nmesh'], 'TSC', w=w)
    want = oracle.get_field(pos.copy(), c['L'], c['nmesh'], 'TSC', w=w, nthread=1)
    assert got.shape == want.shape and got.dtype == np.float32
    assert np.allclose(got, want, rtol=1e-4, atol=1e-5)
    raw = np.random.default_rng(0).random((12, 12, 12), dtype='f4')
    np.testing.assert_allclose(ps.normalize_field(raw), raw / raw.mean(dtype='f8') - 1, rtol=2e-5, atol=2e-6)
    r2 = raw.copy()
    ps.normalize_field(r2, tot_weight=50.0, inplace=True)
    np.testing.assert_allclose(r2, raw * np.float32(raw.size / 50.0) - 1, rtol=1e-6, atol=1e-6)


def test_shift_field_fft(ps, oracle):
    c = cases.DELTAK_CASES['d32']
    f1, f2, _ = cases.deltak_inputs(c)
    a, b = f1.copy(), f1.copy()
    ps.shift_field_fft(a, f2, c['n'], c['L'], c['L'] / c['n'])
    oracle.shift_field_fft(b, f2, c['n'], c['L'], c['L'] / c['n'], nthread=2)
    assert np.abs(a - b).max() < 1e-6 * np.abs(b).max()
    for d in (0.37 * c['L'] / c['n'], 2.5 * c['L'] / c['n'], -0.8):      # any shift, not only the interlacing one
        a, b = f1.copy(), f1.copy()
        ps.shift_field_fft(a, f2, c['n'], c['L'], d)
        oracle.shift_field_fft(b, f2, c['n'], c['L'], d, nthread=2)
        assert np.abs(a - b).max() < 2e-6 * np.abs(b).max(), d


@pytest.mark.parametrize('name', list(cases.POWER_CASES))
def test_calc_power_vs_reference(ps, golden, name):
    c = cases.POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                      paste='TSC', nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'],
                      w=w, pos2=pos2, w2=w2, poles=c['poles'])
    pre = f'power/{name}/'
    want = {k[len(pre):]: golden[k] for k in golden.files if k.startswith(pre)}
    assert set(want) == set(t.keys())
    for k in want:
        assert np.asarray(t[k]).shape == want[k].shape, k
        assert np.asarray(t[k]).dtype == want[k].dtype, k
    compare_power_tables(t, want)
    assert t.meta['N_pos'] == len(pos) and t.meta['nmesh'] == c['nmesh']


def test_calc_power_device_inputs_match_host_inputs(ps):
    import torch

    c = cases.POWER_CASES['n32_ci']
    pos, w, _, _ = cases.power_inputs(c)
    kw = dict(kbins=16, mubins=4, nmesh=32, poles=[0, 2, 4])
    a = ps.calc_power(pos, c['L'], w=w, **kw)
    b = ps.calc_power(torch.from_numpy(pos).cuda(), c['L'], w=torch.from_numpy(w).cuda(), **kw)
    d = ps.calc_power(torch.from_numpy(pos).pin_memory(), c['L'], w=torch.from_numpy(w).pin_memory(), **kw)
    for other in (b, d):
        assert_int_exact(a['N_mode'], other['N_mode'])
        np.testing.assert_allclose(a['power'], other['power'], rtol=2e-5)


def test_monopole_identity(ps):
    """tests/test_power.py:58-61 of the reference."""
    c = cases.POWER_CASES['n32_ci']
    pos, w, _, _ = cases.power_inputs(c)
    t = ps.calc_power(pos, c['L'], kbins=16, mubins=4, nmesh=32, w=w, poles=[0, 2, 4])
    P = (t['power'] * t['N_mode']).sum(axis=1) / t['N_mode'].sum(axis=1)
    assert np.allclose(np.nan_to_num(P), t['poles'][:, 0], rtol=1e-5)


def test_errors(ps):
    pos = np.zeros((10, 3), dtype=np.float32)
    with pytest.raises(ValueError):
        ps.calc_power(pos, 100.0, paste='NGP', nmesh=8)
    with pytest.raises(AssertionError):
        ps.calc_power(pos, 100.0, nmesh=8, w=np.ones(3, dtype=np.float32))
    with pytest.raises(AssertionError):
        ps.get_field_fft(pos, 100.0, 8, 'TSC', None, None, True, False)


def test_config1_full_vs_oracle(ps, oracle):
    """BASELINE config 1 at full size: 1e6 uniform particles, L=1000, nmesh=128, TSC, compensated,
    non-interlaced, poles 0/2/4, default bins -- GPU vs the CPU oracle (float64 bin sums)."""
    rng = np.random.default_rng(12345)
    pos = rng.random((1000000, 3), dtype='f4') * np.float32(1000.0)
    kw = dict(nmesh=128, compensated=True, interlaced=False, poles=[0, 2, 4])
    got = ps.calc_power(pos.copy(), 1000.0, **kw)
    want = oracle.calc_power(pos.copy(), 1000.0, acc64=True, **kw)
    compare_power_tables(got, want)


@pytest.mark.parametrize('name', list(cases.XI_CASES))
def test_pk_to_xi(ps, name):
    """xi(r) multipoles: C2R FFT + real-space binning (fourier=False) vs the unmodified reference."""
    g = np.load(cases.__file__.replace('cases.py', 'reference_xi.npz'))
    c = cases.XI_CASES[name]
    Pk, r_bins = cases.xi_inputs(c)
    r_binc, bp, Npo = ps.pk_to_xi(Pk, c['L'], r_bins, poles=c['poles'])
    want = g[f'xi/{name}/binned_poles']
    np.testing.assert_allclose(r_binc, g[f'xi/{name}/r_binc'])
    assert bp.dtype == want.dtype and bp.shape == want.shape
    assert_int_exact(Npo, g[f'xi/{name}/Npoles'])
    assert_close_scaled(bp, want, scale=np.abs(want).max(), rtol=1e-4, what='xi poles')


def test_bin_kmu_configuration_space(ps, oracle):
    """bin_kmu(fourier=False) on a real-space mesh (n,n,n): only r_z <= n/2 is visited, dk = L/n."""
    n, L = 28, 140.0
    rng = np.random.default_rng(5)
    xi = rng.standard_normal((n, n, n)).astype(np.float32)
    r_bins = np.linspace(0, 60.0, 13)
    mu = np.linspace(0, 1, 4)
    got = ps.bin_kmu(n, L, r_bins, mu, xi, poles=np.array([0, 2]), fourier=False)
    want = oracle.bin_kmu(n, L, r_bins, mu, xi, poles=np.array([0, 2]), fourier=False, acc64=True)
    assert_int_exact(got[1], want[1])
    assert_int_exact(got[3], want[3])
    assert_close_scaled(got[0], want[0], scale=0.05, rtol=1e-5)
    assert_close_scaled(got[4], want[4], rtol=1e-5)
    assert_close_scaled(got[2], want[2], scale=0.05, rtol=1e-5)


@pytest.mark.parametrize('name', list(cases.CIC_POWER_CASES))
def test_cic_calc_power(ps, name):
    """paste='CIC' (analysis/cic.py through power_spectrum.py:846-853) vs the unmodified reference."""
    g = np.load(cases.__file__.replace('cases.py', 'reference_cic.npz'))
    c = cases.CIC_POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], logk=c['logk'], paste='CIC', nmesh=c['nmesh'],
                      compensated=c['compensated'], interlaced=c['interlaced'], w=w, pos2=pos2, w2=w2, poles=c['poles'])
    pre = f'power/{name}/'
    want = {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
    assert set(want) == set(t.keys())
    compare_power_tables(t, want)


@pytest.mark.parametrize('name', list(cases.CIC_FIELD_CASES))
def test_cic_field(ps, name):
    g = np.load(cases.__file__.replace('cases.py', 'reference_cic.npz'))
    c = cases.CIC_FIELD_CASES[name]
    pos, w = cases.cic_field_inputs(c)
    f = ps.get_field(pos, c['L'], c['nmesh'], 'CIC', w=w, d=c['d'])
    np.testing.assert_allclose(f, g[f'field/{name}'], rtol=1e-4, atol=1e-5)
    # and TSC still is TSC afterwards (the scheme is per call, not sticky)
    f2 = ps.get_field(pos, c['L'], c['nmesh'], 'TSC', w=w, d=c['d'])
    assert np.abs(f2 - f).max() > 1e-3


@pytest.mark.parametrize('name', list(cases.KFIELD_CASES))
def test_kfield_helpers(ps, name):
    """get_delta_mu2 / get_smoothing / expand_poles_to_3d (ZCV helpers) vs the unmodified reference."""
    g = np.load(cases.__file__.replace('cases.py', 'reference_kfields.npz'))
    c = cases.KFIELD_CASES[name]
    delta, k_ell, P_ell = cases.kfield_inputs(c)
    got = ps.get_delta_mu2(delta, c['n'])
    want = g[f'kf/{name}/delta_mu2']
    assert got.dtype == want.dtype and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    got = ps.get_smoothing(c['n'], c['L'], c['R'])
    want = g[f'kf/{name}/smoothing']
    assert got.dtype == want.dtype and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-30)
    got = ps.expand_poles_to_3d(k_ell, P_ell, c['n'], c['L'], np.asarray(c['poles']))
    want = g[f'kf/{name}/expand']
    assert got.dtype == want.dtype and got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


def test_fused_normalisation_matches_separate_pass(ps, monkeypatch):
    """normalize_field folded into the deposit (grid starts at -1, weights scaled by n^3/N; the default) against the
    separate normalisation kernel (ABK_FUSED_NORMALIZE=0): same spectrum to float32 round-off, same mode counts."""
    c = cases.POWER_CASES['n32_ci']
    pos, w, _, _ = cases.power_inputs(c)
    kw = dict(kbins=16, mubins=4, nmesh=32, poles=[0, 2, 4], w=w)
    monkeypatch.setenv('ABK_FUSED_NORMALIZE', '1')
    a = ps.calc_power(pos, c['L'], **kw)
    monkeypatch.setenv('ABK_FUSED_NORMALIZE', '0')
    b = ps.calc_power(pos, c['L'], **kw)
    assert_int_exact(a['N_mode'], b['N_mode'])
    np.testing.assert_allclose(a['power'], b['power'], rtol=2e-5, atol=2e-6 * np.abs(np.asarray(b['power'])).max())
    np.testing.assert_allclose(a['poles'], b['poles'], rtol=1e-4, atol=2e-5 * np.abs(np.asarray(b['poles'])).max())


def test_cic_serial_drop_in():
    """abacusutils_b200.analysis.cic.cic_serial (cic.py:13-125) against the reference's CIC field."""
    from abacusutils_b200.analysis.cic import cic_serial

    g = np.load(cases.__file__.replace('cases.py', 'reference_cic.npz'))
    name = 'cicf24'
    c = cases.CIC_FIELD_CASES[name]
    pos, w = cases.cic_field_inputs(c)
    n = c['nmesh']
    dens = np.full((n, n, n), 2.0, dtype=np.float32)                   # accumulates into the caller's grid
    assert cic_serial(pos, dens, c['L'], weights=w) is None
    field = (dens - 2.0) * np.float32(n**3 / len(pos)) - 1
    np.testing.assert_allclose(field, g[f'field/{name}'], rtol=1e-4, atol=2e-5)


def clustered_anisotropic(N, L, seed, frac=0.7, nblob=200, sigma=0.02, squash=0.2):
    """Gaussian blobs squashed along z (sigma_z = squash * sigma_xy) on a uniform background: the quadrupole and
    hexadecapole are O(P0), so the relative 1e-4 bites on every multipole; dense cells exercise the capacity passes."""
    rng = np.random.default_rng(seed)
    nb = int(N * frac)
    centers = rng.random((nblob, 3)) * L
    d = rng.standard_normal((nb, 3)) * (sigma * L) * np.array([1.0, 1.0, squash])
    blob = (centers[rng.integers(0, nblob, nb)] + d) % L
    pos = np.concatenate([blob, rng.random((N - nb, 3)) * L]).astype(np.float32)
    return np.minimum(pos, np.nextafter(np.float32(L), np.float32(0)))


def test_large_mesh_interlaced_vs_oracle(ps, oracle):
    """1e7 uniform particles on a 512^3 mesh (0.075 per cell; 17 z-tiles with a ragged last one), TSC, compensated +
    interlaced, 100 x 10 bins, poles 0/2/4 -- the bench configuration at 1/100 of its size -- vs the CPU oracle."""
    rng = np.random.default_rng(31)
    L = 1000.0
    pos = rng.random((10_000_000, 3), dtype='f4') * np.float32(L)
    kw = dict(kbins=100, mubins=10, nmesh=512, compensated=True, interlaced=True, poles=[0, 2, 4])
    got = ps.calc_power(pos.copy(), L, **kw)
    want = oracle.calc_power(pos.copy(), L, acc64=True, **kw)
    compare_power_tables(got, want)


@pytest.mark.parametrize('weighted', [False, True])
def test_clustered_anisotropic_vs_oracle(ps, oracle, weighted):
    """Clustered + anisotropic catalogue (blobs squashed along z): multipoles of order P0, dense tiles (several capacity
    passes), weights.  Power and multipoles within a relative 1e-4 (+ 2e-5 P0), counts bit-exact."""
    L, N, n = 500.0, 3_000_000, 256
    pos = clustered_anisotropic(N, L, 5)
    w = np.random.default_rng(6).random(N, dtype='f4') if weighted else None
    kw = dict(kbins=64, mubins=8, nmesh=n, compensated=True, interlaced=True, poles=[0, 2, 4], w=w)
    got = ps.calc_power(pos.copy(), L, **kw)
    want = oracle.calc_power(pos.copy(), L, acc64=True, **kw)
    # the input is strongly anisotropic: |P2| is a sizeable fraction of P0 on most scales
    p0, p2 = np.abs(want['poles'][2:40, 0]), np.abs(want['poles'][2:40, 1])
    assert np.median(p2 / p0) > 0.2
    compare_power_tables(got, want)
    # where a multipole is O(P0) the pure relative criterion holds on its own
    big = np.abs(want['poles']) > 0.05 * np.abs(want['poles'][:, :1])
    big[0] = False
    rel = np.abs(np.asarray(got['poles'], 'f8') - want['poles'])[big] / np.abs(want['poles'])[big]
    assert rel.max() < 1e-4, rel.max()


def test_meta_carries_kernel_timings(ps, monkeypatch):
    """ABK_META_TIMINGS=1: per-kernel CUDA-event times of the call in meta['kernel_ms']."""
    c = cases.POWER_CASES['n32_ci']
    pos, w, _, _ = cases.power_inputs(c)
    monkeypatch.setenv('ABK_META_TIMINGS', '1')
    t = ps.calc_power(pos, c['L'], kbins=16, mubins=4, nmesh=32, w=w, poles=[0, 2, 4])
    km = t.meta['kernel_ms']
    assert {'tsc_bucket_scatter', 'tsc_tile_deposit', 'cufft', 'power_bin'} <= set(km)
    assert all(v['ms'] > 0 and v['launches'] >= 1 for v in km.values())
    monkeypatch.delenv('ABK_META_TIMINGS')
    assert 'kernel_ms' not in ps.calc_power(pos, c['L'], kbins=16, mubins=4, nmesh=32, w=w).meta
