"""
GPU parity of bin_kppi (abk_kfields.cu) against outputs of the unmodified reference
(tests/golden/reference_kppi.npz) and against the CPU oracle at a larger mesh: mode counts bit-exact,
means within relative 1e-4 of the largest mean.
"""

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ps():
    from abacusutils_b200.analysis import power_spectrum as m

    return m


@pytest.mark.parametrize('name', list(cases.KPPI_CASES))
def test_bin_kppi_vs_reference(ps, name):
    g = np.load(cases.__file__.replace('cases.py', 'reference_kppi.npz'))
    c = cases.KPPI_CASES[name]
    w, kedges, pimax = cases.kppi_inputs(c)
    mean, cnt = ps.bin_kppi(c['n'], c['L'], kedges, pimax, c['Npi'], w, dtype=np.dtype(c['dtype']).type,
                            fourier=c['fourier'])
    want = g[f'kppi/{name}/mean']
    assert mean.dtype == want.dtype and mean.shape == want.shape and cnt.dtype == np.int64
    np.testing.assert_array_equal(cnt, g[f'kppi/{name}/counts'])
    np.testing.assert_allclose(mean, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


def test_bin_kppi_device_input_and_oracle(ps, oracle):
    """A 160^3 half-spectrum resident on the device (torch tensor in), checked against the oracle."""
    import torch

    n, L = 160, 500.0
    rng = np.random.default_rng(5)
    w = rng.standard_normal((n, n, n // 2 + 1)).astype(np.float32)
    kedges = np.linspace(0.0, 0.8 * np.pi * n / L, 31)
    pimax = 0.5 * np.pi * n / L
    mean, cnt = ps.bin_kppi(n, L, kedges, pimax, 17, torch.from_numpy(w).cuda())
    omean, ocnt = oracle.bin_kppi(n, L, kedges, pimax, 17, w, raw=True)
    np.testing.assert_array_equal(cnt, ocnt)
    np.testing.assert_allclose(mean, omean, rtol=1e-4, atol=1e-4 * np.abs(omean).max())


def test_bin_kppi_rejects_bad_shape(ps):
    with pytest.raises(ValueError):
        ps.bin_kppi(16, 100.0, np.linspace(0, 1, 4), 0.5, 3, np.zeros((16, 16, 5), dtype=np.float32))
