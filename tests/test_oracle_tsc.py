"""
Pin the CPU oracle's TSC restatement (oracle/abk_oracle.c) against the reference's own tests and
golden vectors: tests/test_tsc.py of the reference (analytic single-particle KAT :25-90, mass
conservation :120, golden grids :128-159, partition invariants :162-208, return contract :211-230)
and grids produced by the unmodified reference (tests/golden/reference_runs.npz).
"""

import numpy as np
import pytest

import cases


@pytest.mark.parametrize('ngrid', [10, 256])
@pytest.mark.parametrize('nthread', [1, 4], ids=['serial', 'parallel'])
def test_single(oracle, ngrid, nthread):
    box = 123.0
    cen = np.array([5, 6, 7])
    single = (cen / ngrid * box).astype('f4').reshape(1, -1)
    dens = oracle.tsc_parallel(single, ngrid, box, nthread=nthread)
    assert (dens == 0).sum() == ngrid**3 - 27
    assert np.isclose(dens.sum(), 1.0)
    cube = dens[cen[0] - 1:cen[0] + 2, cen[1] - 1:cen[1] + 2, cen[2] - 1:cen[2] + 2]
    nface = (np.indices((3, 3, 3)) == 1).sum(axis=0)  # how many coordinates are central
    assert np.allclose(cube[nface == 0], 0.5**9)
    assert np.allclose(cube[nface == 1], 0.5**6 * 0.75)
    assert np.allclose(cube[nface == 2], 0.5**3 * 0.75**2)
    assert np.allclose(cube[nface == 3], 0.75**3)


@pytest.mark.parametrize('nthread', [1, 4], ids=['serial', 'parallel'])
def test_multi_vs_reference_golden_ngrid10(oracle, nthread):
    pos, weights, box = cases.ref_tsc_inputs()
    g = np.load(cases.__file__.replace('cases.py', 'ref_tsc_ngrid10.npz'))
    dens = oracle.tsc_parallel(pos, 10, box, nthread=nthread, weights=weights)
    assert np.isclose(dens.sum(dtype='f8'), weights.sum(dtype='f8'))
    assert np.allclose(dens, g['pydens'], rtol=1e-4, atol=1e-5)
    assert np.allclose(dens, g['nbodykit'], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('nthread', [1, 4], ids=['serial', 'parallel'])
def test_multi_vs_reference_golden_ngrid256(oracle, nthread):
    pos, weights, box = cases.ref_tsc_inputs()
    g = np.load(cases.__file__.replace('cases.py', 'ref_tsc_ngrid256.npz'))
    dens = oracle.tsc_parallel(pos, 256, box, nthread=nthread, weights=weights)
    assert np.isclose(dens.sum(dtype='f8'), weights.sum(dtype='f8'))
    assert np.isclose(dens.sum(dtype='f8'), float(g['own_sum']), rtol=1e-6)
    assert np.isclose((dens.astype('f8') ** 2).sum(), float(g['own_sumsq']), rtol=1e-5)
    assert (dens != 0).sum() == int(g['own_nnz'])
    assert np.allclose(dens.sum(axis=(1, 2), dtype='f8'), g['own_xsum'], rtol=1e-5, atol=1e-5)
    assert np.allclose(dens.sum(axis=(0, 1), dtype='f8'), g['own_zsum'], rtol=1e-5, atol=1e-5)
    sub = dens[g['planes']].reshape(-1)
    for tag in ('own', 'nbk'):
        want = np.zeros_like(sub)
        want[g[f'{tag}_idx']] = g[f'{tag}_val']
        assert np.allclose(sub, want, rtol=1e-4, atol=1e-5), tag


@pytest.mark.parametrize('name', list(cases.TSC_CASES))
def test_vs_reference_runs(oracle, golden, name):
    c = cases.TSC_CASES[name]
    pos, w = cases.tsc_inputs(c)
    want = golden[f'tsc/{name}']
    # serial restatement
    dens = np.zeros(c['shape'], dtype=np.float32)
    p = pos.copy()
    oracle.tsc_parallel(p, dens, c['box'], weights=w, nthread=1, offset=c['offset'])
    assert np.allclose(dens, want, rtol=1e-5, atol=1e-6)
    # striped restatement
    dens2 = np.zeros(c['shape'], dtype=np.float32)
    oracle.tsc_parallel(pos.copy(), dens2, c['box'], weights=w, nthread=4, offset=c['offset'])
    assert np.allclose(dens2, want, rtol=1e-5, atol=1e-6)


def test_wrap_inplace_mutates(oracle):
    c = cases.TSC_CASES['unwrapped']
    pos, w = cases.tsc_inputs(c)
    assert (pos < 0).any() and (pos >= c['box']).any()
    oracle.tsc_parallel(pos, 20, c['box'], weights=w, nthread=2)
    assert pos.min() >= 0 and pos.max() <= c['box']


@pytest.mark.parametrize('name', list(cases.PARTITION_CASES))
def test_partition_vs_reference(oracle, golden, name):
    c = cases.PARTITION_CASES[name]
    pos, w = cases.tsc_inputs(c)
    ppart, starts, wpart = oracle.partition_parallel(pos, c['npartition'], c['box'], weights=w, nthread=3)
    assert np.array_equal(starts, golden[f'partition/{name}/starts'])
    # the reference's partition is the stable one for any thread count
    assert np.array_equal(ppart, golden[f'partition/{name}/ppart'])
    assert np.array_equal(wpart, golden[f'partition/{name}/wpart'])
    keys = np.minimum((pos[:, 0] * np.float32(c['npartition'] / c['box'])).astype(np.int32), c['npartition'] - 1)
    np_starts = np.r_[0, np.bincount(keys, minlength=c['npartition']).cumsum()]
    assert np.array_equal(starts, np_starts)


def test_returns(oracle):
    rng = np.random.default_rng(123)
    box, ngrid = 123.0, 10
    pos = rng.random((100, 3), dtype='f4') * box
    dens = oracle.tsc_parallel(pos, ngrid, box)
    assert dens.shape == (ngrid, ngrid, ngrid)
    dens_allocated = np.zeros((ngrid, ngrid, ngrid), dtype=np.float32)
    assert oracle.tsc_parallel(pos, dens_allocated, box) is None
    np.testing.assert_allclose(dens_allocated, dens)
