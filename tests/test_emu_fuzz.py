"""
Randomised differential tests on the CPU emulator (tests/emu): the product's tsc_parallel / calc_power with random
mesh shapes, offsets, wrapping, weights, clustering, kernel variants, capacity overflow, chunking and the experiment
knobs, against the CPU oracle.  Seeds are fixed; raise the trial counts with ABK_FUZZ_TRIALS for a longer soak.
"""

import os
import warnings

import numpy as np
import pytest

from common import compare_power_tables

TRIALS = int(os.environ.get('ABK_FUZZ_TRIALS', '12'))


@pytest.fixture
def emu(monkeypatch, emu_build_dir):
    import emu_engine

    return emu_engine.install(monkeypatch, emu_build_dir)


def test_fuzz_tsc_parallel(emu, oracle):
    from abacusutils_b200._lib import check
    from abacusutils_b200.analysis import tsc

    rng = np.random.default_rng(1)
    try:
        for trial in range(TRIALS):
            shape = tuple(int(x) for x in rng.integers(3, 60, size=3)) if rng.random() < 0.5 else (int(rng.integers(3, 60)),) * 3
            N = int(rng.integers(0, 3000))
            box = float(rng.uniform(1, 500))
            off = float(rng.choice([0.0, rng.uniform(-box / shape[0], box / shape[0])]))
            weighted, wrap = bool(rng.integers(0, 2)), bool(rng.integers(0, 2))
            lo, hi = (-0.99 * box, 1.99 * box) if wrap else (0.0, float(np.nextafter(np.float32(box), np.float32(0))))
            pos = (rng.random((N, 3), dtype=np.float32) * np.float32(hi - lo) + np.float32(lo)).astype(np.float32)
            if N > 3 and rng.random() < 0.5:
                k = N // 2
                pos[:k] = (pos[0] + rng.normal(0, box / 200, size=(k, 3))).astype(np.float32)
            pos = np.clip(pos, lo, hi).astype(np.float32)
            w = rng.random(N, dtype=np.float32) if weighted else None
            cap = int(rng.choice([0, 256, 512]))
            check(emu.lib.abk_ctx_set_tile_capacity(emu.ctx, cap))
            got, want = np.zeros(shape, np.float32), np.zeros(shape, np.float32)
            tsc.tsc_parallel(pos.copy(), got, box, weights=w, offset=off, wrap=wrap)
            oracle.tsc_parallel(pos.copy(), want, box, weights=w, offset=off, wrap=wrap, nthread=1)
            np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5 + 1e-5 * np.abs(want).max(),
                                       err_msg=f'trial {trial}: shape={shape} N={N} off={off} wrap={wrap} cap={cap}')
    finally:
        check(emu.lib.abk_ctx_set_tile_capacity(emu.ctx, 0))


def test_fuzz_calc_power(emu, oracle, monkeypatch):
    from abacusutils_b200.analysis import power_spectrum as ps

    rng = np.random.default_rng(2)
    for trial in range(max(TRIALS * 2 // 3, 1)):
        n = int(rng.choice([8, 12, 16, 20, 24, 27, 32, 36]))
        N, L = int(rng.integers(200, 5000)), float(rng.uniform(50, 2000))
        pos = (rng.random((N, 3), dtype=np.float32) * np.float32(L)).astype(np.float32)
        w = rng.random(N, dtype=np.float32) if rng.random() < 0.5 else None
        cross = rng.random() < 0.3
        pos2 = (rng.random((N // 2 + 5, 3), dtype=np.float32) * np.float32(L)).astype(np.float32) if cross else None
        w2 = rng.random(len(pos2), dtype=np.float32) if (cross and rng.random() < 0.5) else None
        kw = dict(kbins=int(rng.integers(1, 30)), mubins=(None if rng.random() < 0.3 else int(rng.integers(1, 8))),
                  logk=bool(rng.integers(0, 2)), nmesh=n, compensated=bool(rng.integers(0, 2)),
                  interlaced=bool(rng.integers(0, 2)),
                  poles=[[], [0, 2, 4], [0, 1, 2, 3], [2], [0, 2, 4, 6, 8, 10]][int(rng.integers(0, 5))],
                  paste=str(rng.choice(['TSC', 'TSC', 'CIC'])))
        if rng.random() < 0.5:
            kw['k_max'] = float(rng.uniform(0.3, 1.7)) * np.pi * n / L
        monkeypatch.setenv('ABK_CHUNK_MIN', str(int(rng.choice([1 << 25, 300, 1000]))))
        monkeypatch.setenv('ABK_SCATTER', str(rng.choice(['1', '2'])))
        monkeypatch.setenv('ABK_EARLY_GROUPS', str(rng.choice(['1', '2', '3'])))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            got = ps.calc_power(pos.copy(), L, w=w, pos2=None if pos2 is None else pos2.copy(), w2=w2, **kw)
            want = oracle.calc_power(pos.copy(), L, w=w, pos2=None if pos2 is None else pos2.copy(), w2=w2, nthread=2, **kw)
        amp = None
        if cross:      # tolerance of a cross-spectrum is set by the auto-spectra, not by its own near-zero value
            with warnings.catch_warnings():
                warnings.simplefilter('ignore')
                a1 = oracle.calc_power(pos.copy(), L, w=w, nthread=2, **{**kw, 'poles': []})
                a2 = oracle.calc_power(pos2.copy(), L, w=w2, nthread=2, **{**kw, 'poles': []})
            p1, p2 = np.abs(np.asarray(a1['power'], 'f8')), np.abs(np.asarray(a2['power'], 'f8'))
            amp = np.sqrt(p1 * p2) if p1.ndim == 1 else np.sqrt(p1 * p2).mean(axis=1)
        try:
            compare_power_tables(got, {k: np.asarray(want[k]) for k in want.keys()}, poles=kw['poles'], amp=amp)
        except AssertionError as e:
            raise AssertionError(f'trial {trial}: n={n} N={N} {kw} cross={cross}: {e}') from e


def test_fuzz_binning_entry_points(emu, oracle):
    """calc_pk_from_deltak (auto / cross, arbitrary user edges incl. mu edges that stop below 1 are avoided),
    project_3d_to_poles and bin_kppi on random meshes against the oracle: integer counts exact."""
    from abacusutils_b200.analysis import power_spectrum as ps

    rng = np.random.default_rng(3)
    for trial in range(max(TRIALS * 2 // 3, 1)):
        n = int(rng.choice([6, 8, 9, 12, 15, 16, 20, 24]))
        L = float(rng.uniform(20, 3000))
        shp = (n, n, n // 2 + 1)
        f1 = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64)
        f2 = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64) if rng.random() < 0.5 else None
        k_ny = np.pi * n / L
        Nk, Nmu = int(rng.integers(1, 20)), int(rng.integers(1, 6))
        kedges = np.sort(rng.uniform(0, 1.8 * k_ny, Nk + 1)) if rng.random() < 0.5 else np.linspace(0, float(rng.uniform(0.4, 1.8)) * k_ny, Nk + 1)
        muedges = np.linspace(0, 1, Nmu + 1)
        poles = np.asarray([[], [0, 2, 4], [0, 1, 2, 3, 5], [4, 10]][int(rng.integers(0, 4))], dtype=np.int64)
        got = ps.calc_pk_from_deltak(f1, L, kedges, muedges, field2_fft=f2, poles=poles)
        want = oracle.calc_pk_from_deltak(f1, L, kedges, muedges, field2_fft=f2, poles=poles, nthread=2, acc64=True)
        msg = f'trial {trial}: n={n} Nk={Nk} Nmu={Nmu} poles={list(poles)} cross={f2 is not None}'
        np.testing.assert_array_equal(got['N_mode'], want['N_mode'], err_msg=msg)
        np.testing.assert_array_equal(got['N_mode_poles'], want['N_mode_poles'], err_msg=msg)
        scale = np.abs(want['power']).max() + 1e-30
        np.testing.assert_allclose(got['power'], want['power'], rtol=2e-4, atol=2e-5 * scale, err_msg=msg)
        np.testing.assert_allclose(got['k_avg'], want['k_avg'], rtol=2e-5, atol=1e-6 * k_ny, err_msg=msg)
        if len(poles):
            np.testing.assert_allclose(got['binned_poles'], want['binned_poles'], rtol=5e-4, atol=2e-4 * scale, err_msg=msg)
        # (k_perp, pi) binning of a real weight mesh
        w = rng.standard_normal(shp).astype(np.float32)
        pimax, Npi = float(rng.uniform(0.2, 1.3)) * k_ny, int(rng.integers(1, 12))
        kp = np.linspace(float(rng.uniform(0, 0.2)) * k_ny, float(rng.uniform(0.3, 1.6)) * k_ny, Nk + 1)
        m1, c1 = ps.bin_kppi(n, L, kp, pimax, Npi, w)
        m2, c2 = oracle.bin_kppi(n, L, kp, pimax, Npi, w)
        np.testing.assert_array_equal(c1, c2, err_msg=msg)
        np.testing.assert_allclose(m1, m2, rtol=1e-4, atol=1e-5 * (np.abs(m2).max() + 1e-30), err_msg=msg)
