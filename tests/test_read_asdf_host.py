"""
Host side of the file ingest (abacusutils_b200/data/{asdf_container,read_abacus}.py; reference
abacusnbody/data/read_abacus.py:34-212 and data/asdf.py):
  * the container reader against the reference's real files (tests/Mini_N64_L32, skipped where the reference tree
    is absent) and against files written by tests/asdf_writer.py (zstd frames, byte- and bit-shuffle, stored frames,
    uncompressed blocks, trailing garbage beyond the declared shape);
  * read_asdf end to end through the fake engine of test_ingest_host.py (host build of the kernels' arithmetic),
    compared with the reference's golden arrays.
"""

from pathlib import Path

import numpy as np
import pytest

import cases
from asdf_writer import write_asdf
from test_ingest_host import GOLD, PID_KEYS, fake_engine, hc  # noqa: F401  (fixtures)

from oracle import ref_shim

REF_SIM = Path(ref_shim.REF_ROOT) / 'tests' / 'Mini_N64_L32'
needs_ref = pytest.mark.skipif(not ref_shim.available(), reason='reference tree not present')


@needs_ref
def test_container_reads_reference_files():
    from abacusutils_b200.data.asdf_container import ArrayRef, AsdfFile

    g = np.load(GOLD)
    af = AsdfFile(REF_SIM / 'halos' / 'z0.000' / 'field_rv_A' / 'field_rv_A_000.asdf')
    ref = af.tree['data']['rvint']
    assert isinstance(ref, ArrayRef) and ref.shape == (1646, 3)
    np.testing.assert_array_equal(af.read(ref), g['rvint/in'])
    assert af.tree['header']['BoxSize'] == float(g['rvint/box'])
    af = AsdfFile(REF_SIM / 'slices' / 'z0.000' / 'L0_pack9' / 'slab000.L0.pack9.asdf')     # bit-shuffle flag, odd tail
    d = af.read(af.tree['data']['pack9'])
    assert d.shape == (int(g['pack9/nrec_full']), 9) and d.dtype == np.int8
    np.testing.assert_array_equal(d.view(np.uint8)[:cases.PACK9_GOLDEN_RECORDS], g['pack9/in'])
    assert af.tree['header']['VelZSpace_to_kms'] == float(g['pack9/velz'])
    af = AsdfFile(REF_SIM / 'halos' / 'z0.000' / 'field_pid_A' / 'field_pid_A_000.asdf')
    np.testing.assert_array_equal(af.read(af.tree['data']['packedpid']), g['pids/in'])


@needs_ref
def test_read_asdf_reference_files(fake_engine):  # noqa: F811
    """tests/test_data.py:258-325 of the reference, with the decoding done by the kernels' host build."""
    from abacusutils_b200.data.read_abacus import read_asdf

    g = np.load(GOLD)
    t = read_asdf(REF_SIM / 'halos' / 'z0.000' / 'field_rv_A' / 'field_rv_A_000.asdf', load=('pos', 'vel'), verbose=False)
    np.testing.assert_array_equal(t['pos'], g['rvint/pos'])
    np.testing.assert_array_equal(t['vel'], g['rvint/vel'])
    assert t.meta['BoxSize'] == 32.0
    t = read_asdf(REF_SIM / 'slices' / 'z0.000' / 'L0_pack9' / 'slab000.L0.pack9.asdf', dtype=np.float32)
    assert sorted(t.keys() if not hasattr(t, 'colnames') else t.colnames) == ['pos', 'vel']
    assert len(t['pos']) == int(g['pack9/npart_full'])
    n = len(g['pack9/pos'])
    np.testing.assert_array_equal(t['pos'][:n], g['pack9/pos'])
    np.testing.assert_array_equal(t['vel'][:n], g['pack9/vel'])
    t = read_asdf(REF_SIM / 'halos' / 'z0.000' / 'field_pid_A' / 'field_pid_A_000.asdf',
                  load=('aux', 'pid', 'lagr_pos', 'tagged', 'density', 'lagr_idx'))
    np.testing.assert_array_equal(t['aux'], g['pids/in'])
    for k in PID_KEYS:
        np.testing.assert_array_equal(t[k], g[f'pids/{k}'], err_msg=k)
    t = read_asdf(REF_SIM / 'halos' / 'z0.000' / 'field_pid_A' / 'field_pid_A_000.asdf')
    assert list(t.keys() if not hasattr(t, 'colnames') else t.colnames) == ['pid']


@pytest.mark.parametrize('shuffle', ['shuffle', 'bitshuffle', None])
@pytest.mark.parametrize('compression', ['blsc', None])
def test_container_roundtrip_of_written_files(tmp_path, shuffle, compression):
    from abacusutils_b200.data.asdf_container import AsdfFile

    g = np.load(GOLD)
    arrays = {'rvint': g['rvint/in'], 'pack9': g['pack9/in'][:4000].view(np.int8), 'packedpid': g['pids/in']}
    fn = tmp_path / 'f.asdf'
    write_asdf(fn, arrays, {'BoxSize': 32.0, 'ppd': 64.0, 'Name': 'x'}, compression=compression, shuffle=shuffle, pad=4099)
    af = AsdfFile(fn)
    assert af.tree['header'] == {'BoxSize': 32.0, 'ppd': 64.0, 'Name': 'x'}
    for name, a in arrays.items():
        got = af.read(af.tree['data'][name])
        assert got.dtype == a.dtype and got.shape == a.shape
        np.testing.assert_array_equal(got, a)


def test_read_asdf_written_files(tmp_path, fake_engine, oracle):  # noqa: F811
    from abacusutils_b200.data.read_abacus import read_asdf

    g = np.load(GOLD)
    hdr = {'BoxSize': float(g['pack9/box']), 'VelZSpace_to_kms': float(g['pack9/velz']), 'ppd': 64.0}
    fn = tmp_path / 'p9.asdf'
    write_asdf(fn, {'pack9': g['pack9/in'].view(np.int8)}, hdr, shuffle='bitshuffle', pad=700)
    t = read_asdf(fn, load=('pos',), verbose=False)
    assert list(t.keys() if not hasattr(t, 'colnames') else t.colnames) == ['pos']
    np.testing.assert_array_equal(t['pos'], g['pack9/pos'])
    t = read_asdf(fn, dtype=np.float64)
    opos, ovel = oracle.unpack_pack9(g['pack9/in'], hdr['BoxSize'], hdr['VelZSpace_to_kms'], float_dtype=np.float64)
    np.testing.assert_array_equal(t['pos'], opos)
    np.testing.assert_array_equal(t['vel'], ovel)
    fn2 = tmp_path / 'two.asdf'
    write_asdf(fn2, {'rvint': g['rvint/in'], 'pack9': g['pack9/in'][:100].view(np.int8)}, hdr)
    with pytest.raises(ValueError, match='More than one key'):
        read_asdf(fn2)
    t = read_asdf(fn2, colname='rvint', load=('vel',))
    np.testing.assert_array_equal(t['vel'], g['rvint/vel'])
    fn3 = tmp_path / 'none.asdf'
    write_asdf(fn3, {'other': g['rvint/in']}, hdr)
    with pytest.raises(ValueError, match='Could not find'):
        read_asdf(fn3)
    with pytest.warns(FutureWarning):
        t = read_asdf(fn2, colname='rvint', load_pos=True)
    assert list(t.keys() if not hasattr(t, 'colnames') else t.colnames) == ['pos']
