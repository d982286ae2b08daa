"""CPU tests of host-side logic of the product package (no GPU, no kernels)."""

import numpy as np
import pytest


def test_slab_plan_properties():
    from abacusutils_b200.dist import TILE_X, SlabPlan

    rng = np.random.default_rng(1)
    for _ in range(300):
        world = int(rng.integers(1, 9))
        n = int(rng.integers(2 * world, 4200))
        p = SlabPlan(n, world)
        for split in (p.xsplit, p.jsplit):
            assert split[0] == 0 and split[-1] == n and len(split) == world + 1
            assert all(b - a >= 1 for a, b in zip(split[:-1], split[1:]))
        assert all(p.nxl(r) >= 2 for r in range(world))
        if p.aligned:
            assert all(x % TILE_X == 0 for x in p.xsplit[:-1])
        for r in range(world):
            send, recv = p.transpose_splits(r)
            assert sum(send) == p.nxl(r) * n * p.nzc and sum(recv) == n * p.nyl(r) * p.nzc
        # what rank r sends to q is what q expects from r
        for r in range(world):
            for q in range(world):
                assert p.transpose_splits(r)[0][q] == p.transpose_splits(q)[1][r]
        for ix in (0, n - 1, n, -1, n // 2):
            o = p.owner_of_plane(ix)
            assert p.xsplit[o] <= ix % n < p.xsplit[o + 1]
    with pytest.raises(ValueError):
        SlabPlan(7, 4)


def test_chunk_plan_covers_everything_within_limits():
    from abacusutils_b200._lib import ABK_MAX_SEGMENTS, SEGMENT_MAX
    from abacusutils_b200.analysis.power_spectrum import _Painter

    P = _Painter.__new__(_Painter)
    for N in (1, 1000, (1 << 25) - 1, 1 << 25, 10**8, 10**9, 5 * 10**9, 12 * 10**9):
        chunks = P.chunk_plan(N)
        assert chunks[0][0] == 0 and chunks[-1][1] == N
        assert all(a2 == b1 for (_, b1), (a2, _) in zip(chunks[:-1], chunks[1:]))
        assert all(0 < b - a <= SEGMENT_MAX for a, b in chunks)
        assert len(chunks) <= ABK_MAX_SEGMENTS


def test_chunk_plan_device_segment_knob(monkeypatch):
    from abacusutils_b200._lib import SEGMENT_MAX
    from abacusutils_b200.analysis.power_spectrum import _Painter

    P = _Painter.__new__(_Painter)
    monkeypatch.delenv('ABK_DEVICE_SEGMENTS', raising=False)
    assert P.chunk_plan(10**9, host=False) == P.chunk_plan(10**9, host=True)      # default: same plan
    monkeypatch.setenv('ABK_DEVICE_SEGMENTS', '1')
    assert P.chunk_plan(10**9, host=False) == [(0, 10**9)]
    assert P.chunk_plan(10**9, host=True) == _Painter.chunk_plan(P, 10**9)          # host plan untouched
    big = P.chunk_plan(3 * 10**9, host=False)                                       # still bounded by SEGMENT_MAX
    assert big[-1][1] == 3 * 10**9 and all(b - a <= SEGMENT_MAX for a, b in big)
    monkeypatch.setenv('ABK_DEVICE_SEGMENTS', '4')
    assert len(P.chunk_plan(10**9, host=False)) == 4


def test_legendre_coefficients_against_numpy():
    from numpy.polynomial import legendre as npl

    from abacusutils_b200.analysis.power_spectrum import legendre_coefficients

    mu = np.linspace(0, 1, 41)
    co = legendre_coefficients(list(range(11))).astype(np.float64)
    for ell in range(11):
        want = (2 * ell + 1) * npl.legval(mu, [0] * ell + [1])
        got = sum(co[ell, m] * mu**m for m in range(11))
        np.testing.assert_allclose(got, want, rtol=2e-6, atol=2e-5)


def test_table_fallback_and_paste_check():
    from abacusutils_b200.analysis import power_spectrum as ps

    t = ps.Table({'a': np.arange(3)}, meta={'x': 1})
    assert t.colnames == ['a'] and t.meta['x'] == 1 and t['a'][2] == 2
    assert ps._check_paste('tsc') == 'TSC' and ps._check_paste('Cic') == 'CIC'
    with pytest.raises(ValueError):
        ps._check_paste('NGP')


def test_finalize_bins_matches_reference_tail():
    """Host tail of bin_kmu (power_spectrum.py:276-293): l=0 pole from the wedge sums, divide where count != 0."""
    import torch

    from abacusutils_b200.analysis.power_spectrum import _finalize_bins

    Nk, Nmu = 4, 3
    poles = np.array([0, 2])
    counts = np.array([[1, 0, 2], [0, 0, 0], [4, 4, 4], [2, 0, 0]], dtype=np.int64)
    sp = np.arange(12, dtype=np.float64).reshape(4, 3) + 1
    sk = 2 * sp
    spl = np.array([[9, 9, 9, 9], [1, 2, 3, 4]], dtype=np.float64)
    sums = torch.from_numpy(np.concatenate([counts.view(np.float64).ravel(), sp.ravel(), sk.ravel(), spl.ravel()]))
    sw, c, p, cp, k = _finalize_bins(sums, Nk, Nmu, poles, dk=0.5)
    assert np.array_equal(c, counts) and np.array_equal(cp, counts.sum(1))
    assert sw[1].tolist() == [4.0, 5.0, 6.0]            # empty bins keep their (zero-count) raw value, never NaN
    assert sw[0, 0] == 1.0 and sw[0, 2] == 1.5 and sw[2, 1] == 2.0
    assert k[2, 0] == pytest.approx(14 * 0.5 / 4)
    np.testing.assert_allclose(p[0], [(1 + 2 + 3) / 3, 15, (7 + 8 + 9) / 12, (10 + 11 + 12) / 2])  # row l=0 = sum over mu / counts
    np.testing.assert_allclose(p[1], [1 / 3, 2, 3 / 12, 4 / 2])


def test_bin_kppi_wrapper_marshalling(monkeypatch):
    """The Python side of bin_kppi (squared-edge tables, dtype flags, row stride, final division) against the
    reference's outputs, with the kernel replaced by a NumPy stand-in that consumes exactly what the wrapper passes."""
    import ctypes as C

    import torch

    import cases
    from abacusutils_b200.analysis import power_spectrum as ps

    def as_np(p, dtype, count):
        addr = p.value if hasattr(p, 'value') else p
        return np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype)

    class Lib:
        def abk_bin_kppi(self, ctx, w, w_f64, n, ldz, ke2, Nk, pe2, Npi, kperp_f32, counts, sums):
            wt = as_np(w, np.float64 if w_f64 else np.float32, n * n * ldz).reshape(n, n, ldz)
            ke, pe = as_np(ke2, np.float64, Nk + 1), as_np(pe2, np.float64, Npi + 1)
            cnt, sm = as_np(counts, np.int64, Nk * Npi).reshape(Nk, Npi), as_np(sums, np.float64, Nk * Npi).reshape(Nk, Npi)
            fold = np.where(np.arange(n) < n // 2, np.arange(n), np.arange(n) - n).astype(np.int64)
            kz2 = np.arange(n // 2 + 1, dtype=np.float64) ** 2
            use = kz2 < pe[-1]
            bpi = np.searchsorted(pe[1:], kz2[use], side='left')
            mult = np.where(np.arange(n // 2 + 1) == 0, 1, 2)[use]
            for i in range(n):
                kp2 = fold[i] ** 2 + fold ** 2
                kp2 = kp2.astype(np.float32).astype(np.float64) if kperp_f32 else kp2.astype(np.float64)
                over = np.flatnonzero(kp2 >= ke[-1])
                jend = over[0] if len(over) else n
                for j in np.flatnonzero(kp2[:jend] >= ke[0]):
                    bk = np.searchsorted(ke[1:], kp2[j], side='left')
                    np.add.at(cnt[bk], bpi, mult)
                    np.add.at(sm[bk], bpi, wt[i, j, :n // 2 + 1][use].astype(np.float64) * mult)
            return 0

    class Eng:
        lib, ctx = Lib(), None

        def bind_stream(self):
            pass

        def to_device(self, arr, dtype=None):
            t = arr if isinstance(arr, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(arr))
            return (t.to(dtype) if dtype is not None and t.dtype != dtype else t).contiguous()

        def zeros(self, shape, dtype):
            return torch.zeros(shape, dtype=dtype)

    monkeypatch.setattr(ps.Engine, 'get', classmethod(lambda cls, device=None: Eng()))
    g = np.load(cases.__file__.replace('cases.py', 'reference_kppi.npz'))
    for name in ('f16', 'f16_break', 'f15_odd', 'r24', 'f20_f64', 'f40_1bin'):
        c = cases.KPPI_CASES[name]
        w, kedges, pimax = cases.kppi_inputs(c)
        mean, cnt = ps.bin_kppi(c['n'], c['L'], kedges, pimax, c['Npi'], w, dtype=np.dtype(c['dtype']).type, fourier=c['fourier'])
        want = g[f'kppi/{name}/mean']
        assert mean.dtype == want.dtype and cnt.dtype == np.int64
        np.testing.assert_array_equal(cnt, g[f'kppi/{name}/counts'], err_msg=name)
        np.testing.assert_allclose(mean, want, rtol=1e-5, atol=2e-6, err_msg=name)


def test_scalar_helpers_match_reference(golden):
    """factorial / n_choose_k / P_n / linear_interp (power_spectrum.py:58-147, 509-536) against the reference's P_n table."""
    import cases
    from abacusutils_b200.analysis import power_spectrum as ps

    assert ps.factorial(0) == 1 and ps.factorial(20) == 2432902008176640000 and ps.factorial_slow(5) == 120
    with pytest.raises(ValueError):
        ps.factorial(21)
    assert ps.n_choose_k(10, 3) == 120 and ps.n_choose_k(20, 10) == 184756
    for ell in range(0, 11):
        got = np.array([ps.P_n(np.float32(x), ell) for x in cases.PN_X], dtype=np.float32)
        np.testing.assert_allclose(got, golden[f'P_n/{ell}'], rtol=2e-5, atol=2e-5 * max(1.0, np.abs(golden[f'P_n/{ell}']).max()))
    x, y = np.linspace(1.0, 3.0, 5), np.array([0.0, 1.0, 4.0, 9.0, 16.0])
    assert ps.linear_interp(0.5, x, y) == 0.0 and ps.linear_interp(3.5, x, y) == 16.0
    assert np.isclose(ps.linear_interp(1.75, x, y), 2.5)
