"""Shared parity helpers.

Tolerances (BASELINE.json north_star): integer outputs bit-exact; float32 spectra within a
relative 1e-4 per bin.  A pure rtol is ill-posed where a quantity crosses zero (l=2,4 multipoles,
cross-spectra, the DC bin): the reference itself moves by ~1e-6 of the monopole scale there between
thread counts, so float comparisons use  |a-b| <= rtol*|b| + rtol*scale  with the
scale given by the caller (the monopole / auto-power amplitude of the same k-bin).
"""

import numpy as np

RTOL = 1e-4


def assert_int_exact(a, b, what=''):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), f'{what}: integer mismatch'


def assert_close_scaled(a, b, scale=None, rtol=RTOL, what=''):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if scale is None:
        scale = 0.0
    tol = rtol * np.abs(b) + rtol * np.abs(scale)
    bad = np.abs(a - b) > tol
    if bad.any():
        i = np.argmax(np.abs(a - b) - tol)
        raise AssertionError(
            f'{what}: {bad.sum()} of {bad.size} bins differ; worst flat idx {i}: got {a.ravel()[i]!r} '
            f'want {b.ravel()[i]!r} tol {np.broadcast_to(tol, a.shape).ravel()[i]!r}')


def power_scale(table_like, key_power='power', key_nmode='N_mode'):
    """Per-k-bin amplitude scale: mode-weighted mean |P| of the bin row (monopole-like)."""
    p = np.abs(np.asarray(table_like[key_power], dtype=np.float64))
    n = np.asarray(table_like[key_nmode], dtype=np.float64)
    if p.ndim == 1:
        return p
    tot = n.sum(axis=1)
    m = np.where(tot > 0, (p * n).sum(axis=1) / np.maximum(tot, 1), 0.0)
    return m


def compare_power_tables(got, want, rtol=RTOL, skip_dc=True):
    """Compare two calc_power results (dict-like with the reference's column names)."""
    assert_int_exact(got['N_mode'], want['N_mode'], 'N_mode')
    scale = power_scale(want)
    # a global floor: the smallest meaningful amplitude is ~1e-6 of the typical power (f32 noise)
    floor = 1e-2 * np.median(scale[scale > 0]) if (scale > 0).any() else 0.0
    pw, pg = np.asarray(want['power'], 'f8'), np.asarray(got['power'], 'f8')
    sc = scale if pw.ndim == 1 else scale[:, None]
    sl = slice(None)
    if skip_dc:
        # the bin that contains only the k=0 mode holds rounding noise in the reference too
        dc = (np.asarray(want['N_mode']).reshape(len(pw), -1).sum(axis=1) <= 1)
        pw, pg = pw[~dc], pg[~dc]
        sc = sc[~dc]
        sl = ~dc
    assert_close_scaled(pg, pw, scale=0.1 * (sc + floor), rtol=rtol, what='power')
    assert_close_scaled(np.asarray(got['k_avg'])[sl], np.asarray(want['k_avg'])[sl], rtol=rtol, what='k_avg')
    for key in ('k_min', 'k_max', 'k_mid', 'mu_min', 'mu_max', 'mu_mid'):
        if key in want:
            np.testing.assert_allclose(np.asarray(got[key]), np.asarray(want[key]), rtol=1e-12, atol=0)
    if 'poles' in want:
        assert_int_exact(got['N_mode_poles'], want['N_mode_poles'], 'N_mode_poles')
        wp, gp = np.asarray(want['poles'], 'f8')[sl], np.asarray(got['poles'], 'f8')[sl]
        scp = (scale[sl] + floor)[:, None] * 11.0  # (2l+1) <= 11 for l <= 5; generous for higher l
        assert_close_scaled(gp, wp, scale=scp * 0.1, rtol=rtol, what='poles')
