"""
GPU parity of the particle decoders (abk_ingest.cu) -- bit-exact against
  * the reference's own fixtures (tests/golden/ref_ingest.npz: inputs from tests/Mini_N64_L32, outputs from
    tests/ref_data, the arrays tests/test_data.py compares with), and
  * the CPU oracle on seeded streams (cell headers at block/warp boundaries, no leading header, empty input),
and the packed -> calc_power path with the positions never leaving the device.

This file sorts last on purpose: the kernels were written after the round-1 GPU budget was spent.  Their per-record
arithmetic is pinned on the CPU (tests/test_ingest_host.py) and the unmodified kernel sources run under the CPU
emulator (tests/test_emu_kernels.py); this is their first execution on a device.
"""

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = cases.__file__.replace('cases.py', 'ref_ingest.npz')


@pytest.fixture(scope='module')
def mods():
    from abacusutils_b200.data import bitpacked, pack9

    return bitpacked, pack9


def test_rvint_reference_fixture(mods):
    bitpacked, _ = mods
    g = np.load(GOLD)
    pos, vel = bitpacked.unpack_rvint(g['rvint/in'], float(g['rvint/box']))
    assert pos.dtype == np.float32 and pos.shape == g['rvint/pos'].shape
    np.testing.assert_array_equal(pos, g['rvint/pos'])
    np.testing.assert_array_equal(vel, g['rvint/vel'])


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_rvint_vs_oracle(mods, oracle, dt):
    bitpacked, _ = mods
    iv = cases.rvint_inputs(31, 300001)
    opos, ovel = oracle.unpack_rvint(iv, 1185.0, float_dtype=dt)
    pos, vel = bitpacked.unpack_rvint(iv, 1185.0, float_dtype=dt)
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)
    buf = np.zeros((300001, 3), dtype=dt)
    assert bitpacked.unpack_rvint(iv, 1185.0, float_dtype=dt, posout=False, velout=buf) == (0, 300001)
    np.testing.assert_array_equal(buf, ovel)


PID_KEYS = ('pid', 'lagr_pos', 'lagr_idx', 'tagged', 'density')


def test_pids_reference_fixture(mods):
    bitpacked, _ = mods
    g = np.load(GOLD)
    got = bitpacked.unpack_pids(g['pids/in'], box=float(g['pids/box']), ppd=float(g['pids/ppd']), **{k: True for k in PID_KEYS})
    for k in PID_KEYS:
        assert got[k].dtype == g[f'pids/{k}'].dtype, k
        np.testing.assert_array_equal(got[k], g[f'pids/{k}'], err_msg=k)


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_pids_vs_oracle(mods, oracle, dt):
    import torch

    bitpacked, _ = mods
    rng = np.random.default_rng(61)
    packed = rng.integers(0, 2**63, size=200001, dtype=np.int64).astype(np.uint64) | (rng.integers(0, 2, size=200001).astype(np.uint64) << np.uint64(63))
    want = oracle.unpack_pids(packed, box=2000.0, ppd=6912, float_dtype=dt, **{k: True for k in PID_KEYS})
    got = bitpacked.unpack_pids(packed, box=2000.0, ppd=6912, float_dtype=dt, **{k: True for k in PID_KEYS})
    for k in PID_KEYS:
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    dev = bitpacked.unpack_pids(torch.from_numpy(packed.view(np.int64)).cuda(), density=True, float_dtype=dt)
    assert dev.keys() == {'density'} and dev['density'].is_cuda
    np.testing.assert_array_equal(dev['density'].cpu().numpy(), want['density'])


def test_pack9_reference_fixture(mods):
    _, pack9 = mods
    g = np.load(GOLD)
    pos, vel = pack9.unpack_pack9(g['pack9/in'], float(g['pack9/box']), float(g['pack9/velz']))
    assert pos.dtype == np.float32 and pos.shape == g['pack9/pos'].shape
    np.testing.assert_array_equal(pos, g['pack9/pos'])
    np.testing.assert_array_equal(vel, g['pack9/vel'])


@pytest.mark.parametrize('dt', [np.float32, np.float64])
@pytest.mark.parametrize('case', [
    dict(seed=41, nrec=200003, hdr_frac=0.05, cpd=875, first=True),
    dict(seed=42, nrec=65536, hdr_frac=0.4, cpd=1701, first=True),      # dense headers, exact multiple of the block
    dict(seed=43, nrec=70001, hdr_frac=0.0005, cpd=405, first=True),    # headers many blocks apart
    dict(seed=44, nrec=5000, hdr_frac=0.01, cpd=875, first=False),      # particles before any header -> NaN
    dict(seed=45, nrec=255, hdr_frac=0.1, cpd=875, first=True),
    dict(seed=46, nrec=1, hdr_frac=0.0, cpd=875, first=True),           # a lone header: zero particles
])
def test_pack9_vs_oracle(mods, oracle, dt, case):
    _, pack9 = mods
    d = cases.pack9_inputs(case['seed'], case['nrec'], hdr_frac=case['hdr_frac'], cpd=case['cpd'],
                           first_header=case['first'])
    opos, ovel = oracle.unpack_pack9(d, 2000.0, 1234.5678, float_dtype=dt)
    pos, vel = pack9.unpack_pack9(d, 2000.0, 1234.5678, float_dtype=dt)
    assert pos.shape == opos.shape and pos.dtype == dt
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)


def test_pack9_empty_and_outputs(mods, oracle):
    _, pack9 = mods
    pos, vel = pack9.unpack_pack9(np.zeros((0, 9), np.uint8), 100.0, 1.0)
    assert pos.shape == (0, 3) and vel.shape == (0, 3)
    d = cases.pack9_inputs(47, 9000)
    opos, _ = oracle.unpack_pack9(d, 100.0, 1.0)
    buf = np.zeros((9000, 3), dtype=np.float32)
    assert pack9.unpack_pack9(d.view(np.int8), 100.0, 1.0, posout=buf, velout=False) == (len(opos), 0)
    np.testing.assert_array_equal(buf[:len(opos)], opos)


def test_packed_to_power_on_device(mods, oracle):
    """RVint copied to the GPU still packed -> unpack on the device -> calc_power on the device tensor; the
    spectrum equals the one computed from host-decoded positions."""
    import torch

    from abacusutils_b200.analysis.power_spectrum import calc_power

    bitpacked, _ = mods
    rng = np.random.default_rng(51)
    N, L = 200000, 500.0
    iv = (rng.integers(-500000, 500000, size=(N, 3)).astype(np.int32) << 12) | rng.integers(0, 4096, size=(N, 3)).astype(np.int32)
    pos_d, zero = bitpacked.unpack_rvint(torch.from_numpy(iv).cuda(), L, velout=False)
    assert zero == 0 and pos_d.is_cuda and pos_d.dtype == torch.float32
    opos, _ = oracle.unpack_rvint(iv, L)
    np.testing.assert_array_equal(pos_d.cpu().numpy(), opos)
    kw = dict(kbins=16, mubins=4, nmesh=64, poles=[0, 2])
    a = calc_power(pos_d, L, **kw)
    b = calc_power(opos.copy(), L, **kw)
    np.testing.assert_array_equal(np.asarray(a['N_mode']), np.asarray(b['N_mode']))
    np.testing.assert_allclose(np.asarray(a['power']), np.asarray(b['power']), rtol=1e-4, atol=1e-4 * np.abs(np.asarray(b['power'])).max())


def test_read_asdf_to_device(tmp_path, oracle):
    """File -> host decompression -> packed bytes to the GPU -> decode there -> columns stay on the device."""
    from asdf_writer import write_asdf

    from abacusutils_b200.data.read_abacus import read_asdf

    g = np.load(GOLD)
    hdr = {'BoxSize': float(g['pack9/box']), 'VelZSpace_to_kms': float(g['pack9/velz']), 'ppd': float(g['pids/ppd'])}
    fn = tmp_path / 'p9.asdf'
    write_asdf(fn, {'pack9': g['pack9/in'].view(np.int8)}, hdr, shuffle='bitshuffle', pad=700)
    t = read_asdf(fn, device=True, verbose=False)
    assert t['pos'].is_cuda and t['vel'].is_cuda and t.meta['BoxSize'] == hdr['BoxSize']
    np.testing.assert_array_equal(t['pos'].cpu().numpy(), g['pack9/pos'])
    np.testing.assert_array_equal(t['vel'].cpu().numpy(), g['pack9/vel'])
    t = read_asdf(fn, load=('pos',))
    np.testing.assert_array_equal(t['pos'], g['pack9/pos'])
    fn = tmp_path / 'rv.asdf'
    write_asdf(fn, {'rvint': g['rvint/in']}, hdr)
    t = read_asdf(fn)
    np.testing.assert_array_equal(t['pos'], g['rvint/pos'])
    np.testing.assert_array_equal(t['vel'], g['rvint/vel'])
    fn = tmp_path / 'pid.asdf'
    write_asdf(fn, {'packedpid': g['pids/in']}, hdr)
    t = read_asdf(fn, load=('aux', 'pid', 'lagr_pos', 'tagged', 'density', 'lagr_idx'))
    np.testing.assert_array_equal(t['aux'], g['pids/in'])
    for k in PID_KEYS:
        np.testing.assert_array_equal(t[k], g[f'pids/{k}'], err_msg=k)


@pytest.mark.parametrize('kind', ['pack9', 'rvint'])
def test_calc_power_from_packed_records(oracle, kind, monkeypatch):
    """calc_power(PackedParticles(...)): packed records over PCIe, decoded on the device inside the painter pipeline."""
    from abacusutils_b200.analysis import power_spectrum as ps
    from abacusutils_b200.data.packed import PackedParticles

    monkeypatch.setenv('ABK_CHUNK_MIN', '20000')        # several chunks, early deposit group included
    L = 1000.0
    if kind == 'pack9':
        raw = cases.pack9_inputs(77, 300000, hdr_frac=0.03, cpd=405)
        pos, _ = oracle.unpack_pack9(raw, L, 1.0)
        src = PackedParticles(raw, L, velzspace_to_kms=1.0)
    else:
        rng = np.random.default_rng(78)
        raw = ((rng.integers(-500000, 500000, size=(250000, 3)).astype(np.int32) << 12) | rng.integers(0, 4096, size=(250000, 3)).astype(np.int32))
        pos, _ = oracle.unpack_rvint(raw, L)
        src = PackedParticles(raw, L)
    kw = dict(kbins=24, mubins=4, nmesh=64, poles=[0, 2, 4])
    got = ps.calc_power(src, L, **kw)
    want = ps.calc_power(pos.copy(), L, **kw)
    assert got.meta['N_pos'] == len(pos) == src.n_particles
    np.testing.assert_array_equal(np.asarray(got['N_mode']), np.asarray(want['N_mode']))
    np.testing.assert_allclose(got['power'], want['power'], rtol=2e-5, atol=2e-6 * np.abs(np.asarray(want['power'])).max())
