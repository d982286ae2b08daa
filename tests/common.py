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


def legendre_conditioning(ell):
    """(2l+1) * sum_k |c_k| of P_l written as a polynomial in mu: how much float32 round-off of the reference's power-sum
    evaluation (power_spectrum.py:121-147) is amplified.  76 for l=4, 22911 for l=10."""
    from math import comb

    ell = int(ell)
    return (2 * ell + 1) * sum(comb(ell, k) * comb(2 * ell - 2 * k, ell) for k in range(ell // 2 + 1)) / 2.0**ell


def compare_power_tables(got, want, rtol=RTOL, skip_dc=True, poles=None, amp=None):
    """Compare two calc_power results (dict-like with the reference's column names).

    ``poles`` (optional): the multipole orders of the ``poles`` columns.  High orders are evaluated by the reference
    in float32 as an alternating power sum, whose own round-off grows with the conditioning of the polynomial (measured:
    for l=10 the reference-style evaluation is 2.4e-4 of the monopole away from the exact float64 value, the GPU's Horner
    evaluation 4e-5); the absolute tolerance of a column grows accordingly.
    ``amp`` (optional): per-k-bin amplitude that sets the absolute tolerance instead of |power| itself -- needed for
    cross-spectra of independent catalogues, whose value is a near-zero residual of sqrt(P11 P22)."""
    assert_int_exact(got['N_mode'], want['N_mode'], 'N_mode')
    scale = power_scale(want) if amp is None else np.maximum(power_scale(want), np.asarray(amp, dtype=np.float64))
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
        # absolute term: 2e-5 of the monopole of the k-bin (the reference itself moves by ~1e-6 of P0 between thread
        # counts); where a multipole is O(P0) (anisotropic input) the relative 1e-4 is what binds
        scp = (scale[sl] + floor)[:, None] * 2.0
        if poles is not None and len(poles) == gp.shape[1]:
            # float32 eps * conditioning of the reference's power-sum evaluation; only matters from l = 8 on
            scp = scp * np.maximum(1.0, np.array([legendre_conditioning(l) for l in poles]) * 1.2e-3 / 2.0)[None, :]
        assert_close_scaled(gp, wp, scale=scp * 0.1, rtol=rtol, what='poles')
