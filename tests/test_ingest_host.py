"""
CPU checks of the particle decoders (reference abacusnbody/data/{bitpacked,pack9}.py):

  * the oracle restatements against the reference's own fixtures (tests/golden/ref_ingest.npz: packed inputs of
    tests/Mini_N64_L32 and the decoded arrays of tests/ref_data the reference test-suite compares against);
  * the per-record arithmetic of the CUDA kernels (abacusutils_b200/csrc/abk_ingest.cuh), compiled for the HOST
    by g++ -ffp-contract=off (tests/hostcheck/ingest_host.cpp), bit for bit against the same fixtures and the
    oracle on seeded streams -- every rounding step of the kernels is pinned here without a GPU;
  * the Python wrappers' argument/return conventions, driven through a fake engine that runs the host build.
"""

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import cases

ROOT = Path(__file__).resolve().parent.parent
GOLD = ROOT / 'tests' / 'golden' / 'ref_ingest.npz'


@pytest.fixture(scope='module')
def hc(tmp_path_factory):
    out = tmp_path_factory.mktemp('hostcheck') / 'libhc.so'
    subprocess.run(['/usr/bin/g++' if Path('/usr/bin/g++').exists() else 'g++', '-O2', '-ffp-contract=off', '-fPIC',
                    '-shared', f'-I{ROOT / "abacusutils_b200" / "csrc"}', '-o', str(out),
                    str(ROOT / 'tests' / 'hostcheck' / 'ingest_host.cpp')], check=True)
    lib = C.CDLL(str(out))
    for f in ('hc_pack9_f32', 'hc_pack9_f64'):
        getattr(lib, f).restype = C.c_int64
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    for f in ('hc_rvint_f32', 'hc_rvint_f64'):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
    for f in ('hc_pids_f32', 'hc_pids_f64'):
        getattr(lib, f).argtypes = [C.c_void_p, C.c_int64, C.c_double, C.c_int64] + [C.c_void_p] * 5
    return lib


def hc_rvint(lib, iv, box, dt):
    iv = np.ascontiguousarray(iv)
    pos, vel = np.empty(iv.shape, dt), np.empty(iv.shape, dt)
    fn = lib.hc_rvint_f32 if dt == np.float32 else lib.hc_rvint_f64
    fn(iv.ctypes.data, len(iv), box, pos.ctypes.data, vel.ctypes.data)
    return pos, vel


def hc_pack9(lib, d, box, velz, dt):
    d = np.ascontiguousarray(d)
    pos, vel = np.empty((len(d), 3), dt), np.empty((len(d), 3), dt)
    fn = lib.hc_pack9_f32 if dt == np.float32 else lib.hc_pack9_f64
    n = fn(d.ctypes.data, len(d), box, velz, pos.ctypes.data, vel.ctypes.data)
    return pos[:n], vel[:n]


def hc_pids(lib, packed, box, ppd, dt):
    packed = np.ascontiguousarray(packed, dtype=np.uint64)
    n = len(packed)
    out = dict(pid=np.empty(n, np.int64), lagr_pos=np.empty((n, 3), dt), lagr_idx=np.empty((n, 3), np.int16),
               tagged=np.empty(n, np.uint8), density=np.empty(n, dt))
    fn = lib.hc_pids_f32 if dt == np.float32 else lib.hc_pids_f64
    fn(packed.ctypes.data, n, box, ppd, *(out[k].ctypes.data for k in ('pid', 'lagr_pos', 'lagr_idx', 'tagged', 'density')))
    return out


PID_KEYS = ('pid', 'lagr_pos', 'lagr_idx', 'tagged', 'density')


def random_packed_pids(seed, n):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 2**63, size=n, dtype=np.int64).astype(np.uint64) | (rng.integers(0, 2, size=n).astype(np.uint64) << np.uint64(63))


def test_oracle_pids_vs_reference_fixture(oracle):
    g = np.load(GOLD)
    got = oracle.unpack_pids(g['pids/in'], box=float(g['pids/box']), ppd=float(g['pids/ppd']), **{k: True for k in PID_KEYS})
    for k in PID_KEYS:
        assert got[k].dtype == g[f'pids/{k}'].dtype, k
        np.testing.assert_array_equal(got[k], g[f'pids/{k}'], err_msg=k)
    assert oracle.unpack_pids(g['pids/in'], pid=True).keys() == {'pid'}


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_kernel_arithmetic_pids(hc, oracle, dt):
    g = np.load(GOLD)
    got = hc_pids(hc, g['pids/in'], float(g['pids/box']), int(g['pids/ppd']), dt)
    if dt == np.float32:
        for k in PID_KEYS:
            np.testing.assert_array_equal(got[k], g[f'pids/{k}'], err_msg=k)
    for seed, box, ppd in ((61, 2000.0, 6912), (62, 500.0, 1728), (63, 296.0, 1000)):
        packed = random_packed_pids(seed, 40000)
        got = hc_pids(hc, packed, box, ppd, dt)
        want = oracle.unpack_pids(packed, box=box, ppd=ppd, float_dtype=dt, **{k: True for k in PID_KEYS})
        for k in PID_KEYS:
            np.testing.assert_array_equal(got[k], want[k], err_msg=k)


def test_oracle_rvint_vs_reference_fixture(oracle):
    g = np.load(GOLD)
    pos, vel = oracle.unpack_rvint(g['rvint/in'], float(g['rvint/box']))
    assert pos.dtype == np.float32
    np.testing.assert_array_equal(pos, g['rvint/pos'])
    np.testing.assert_array_equal(vel, g['rvint/vel'])
    # return conventions (bitpacked.py:85-99)
    buf = np.full((len(pos) + 3, 3), -1, dtype=np.float32)
    assert oracle.unpack_rvint(g['rvint/in'], float(g['rvint/box']), posout=buf, velout=False) == (len(pos), 0)
    np.testing.assert_array_equal(buf[:len(pos)], pos)


def test_oracle_pack9_vs_reference_fixture(oracle):
    g = np.load(GOLD)
    pos, vel = oracle.unpack_pack9(g['pack9/in'], float(g['pack9/box']), float(g['pack9/velz']))
    assert pos.shape == g['pack9/pos'].shape and pos.dtype == np.float32
    np.testing.assert_array_equal(pos, g['pack9/pos'])
    np.testing.assert_array_equal(vel, g['pack9/vel'])


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_kernel_arithmetic_rvint(hc, oracle, dt):
    g = np.load(GOLD)
    pos, vel = hc_rvint(hc, g['rvint/in'], float(g['rvint/box']), dt)
    if dt == np.float32:
        np.testing.assert_array_equal(pos, g['rvint/pos'])
        np.testing.assert_array_equal(vel, g['rvint/vel'])
    for seed, box in ((1, 2000.0), (2, 296.0), (3, 1185.0)):
        iv = cases.rvint_inputs(seed, 50000)
        pos, vel = hc_rvint(hc, iv, box, dt)
        opos, ovel = oracle.unpack_rvint(iv, box, float_dtype=dt)
        np.testing.assert_array_equal(pos, opos)
        np.testing.assert_array_equal(vel, ovel)


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_kernel_arithmetic_pack9(hc, oracle, dt):
    g = np.load(GOLD)
    pos, vel = hc_pack9(hc, g['pack9/in'], float(g['pack9/box']), float(g['pack9/velz']), dt)
    if dt == np.float32:
        np.testing.assert_array_equal(pos, g['pack9/pos'])
        np.testing.assert_array_equal(vel, g['pack9/vel'])
    for seed, nrec, cpd, first in ((11, 40000, 875, True), (12, 20000, 1701, True), (13, 3000, 405, False)):
        d = cases.pack9_inputs(seed, nrec, cpd=cpd, first_header=first)
        pos, vel = hc_pack9(hc, d, 2000.0, 1234.5678, dt)
        opos, ovel = oracle.unpack_pack9(d, 2000.0, 1234.5678, float_dtype=dt)
        assert pos.shape == opos.shape
        np.testing.assert_array_equal(pos, opos)   # NaN == NaN for the records before the first header
        np.testing.assert_array_equal(vel, ovel)
        if not first:
            assert np.isnan(opos[0]).all()


# ---- Python wrappers through a fake engine that executes the host build of the kernels' arithmetic ----------
class _FakeLib:
    def __init__(self, hc):
        self.hc = hc

    @staticmethod
    def _addr(p):
        return (p.value if hasattr(p, 'value') else p) or 0

    def abk_unpack_rvint(self, ctx, data, N, box, pos, vel, f64):
        fn = self.hc.hc_rvint_f64 if f64 else self.hc.hc_rvint_f32
        fn(self._addr(data), N, box, self._addr(pos) or None, self._addr(vel) or None)
        return 0

    def abk_unpack_pids(self, ctx, packed, N, box, ppd, pid, lagr_pos, lagr_idx, tagged, density, f64):
        fn = self.hc.hc_pids_f64 if f64 else self.hc.hc_pids_f32
        fn(self._addr(packed), N, box, ppd, *(self._addr(p) or None for p in (pid, lagr_pos, lagr_idx, tagged, density)))
        return 0

    def abk_pack9_scratch_bytes(self, nrec, out):
        out._obj.value = 1024
        return 0

    def abk_pack9_count(self, ctx, data, nrec, scratch, nbytes, out):
        if nrec == 0:
            out._obj.value = 0
            return 0
        raw = np.ctypeslib.as_array(C.cast(self._addr(data), C.POINTER(C.c_uint8)), shape=(max(nrec, 1), 9))
        out._obj.value = int((raw[:nrec, 0] == 0xFF).sum())
        return 0

    def abk_pack9_decode(self, ctx, data, nrec, box, velz, scratch, tab, nhdr, pos, vel, f64):
        if nrec == 0:
            return 0
        dt = np.float64 if f64 else np.float32
        raw = np.ctypeslib.as_array(C.cast(self._addr(data), C.POINTER(C.c_uint8)), shape=(max(nrec, 1), 9))[:nrec]
        p, v = hc_pack9(self.hc, raw, box, velz, dt)
        for dst, src in ((pos, p), (vel, v)):
            if self._addr(dst):
                C.memmove(self._addr(dst), src.ctypes.data, src.nbytes)
        return 0


class _FakeEngine:
    def __init__(self, hc):
        self.lib, self.ctx = _FakeLib(hc), None

    def bind_stream(self):
        pass

    def to_device(self, arr, dtype=None):
        import torch

        return torch.from_numpy(np.ascontiguousarray(arr).copy())

    def empty(self, shape, dtype):
        import torch

        return torch.empty(shape, dtype=dtype)

    def scratch(self, key, nbytes):
        import torch

        return torch.empty(max(int(nbytes), 256), dtype=torch.uint8)


@pytest.fixture
def fake_engine(hc, monkeypatch):
    from abacusutils_b200.data import _common

    eng = _FakeEngine(hc)
    monkeypatch.setattr(_common.Engine, 'get', classmethod(lambda cls, device=None: eng))
    return eng


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_wrapper_unpack_rvint_conventions(fake_engine, oracle, dt):
    from abacusutils_b200.data.bitpacked import unpack_rvint

    iv = cases.rvint_inputs(5, 777)
    opos, ovel = oracle.unpack_rvint(iv, 500.0, float_dtype=dt)
    pos, vel = unpack_rvint(iv.reshape(-1), 500.0, float_dtype=dt)          # flat input is reshaped (bitpacked.py:59)
    assert isinstance(pos, np.ndarray) and pos.dtype == dt and pos.shape == (777, 3)
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)
    buf = np.zeros((800, 3), dtype=dt)
    assert unpack_rvint(iv, 500.0, float_dtype=dt, posout=buf, velout=False) == (777, 0)
    np.testing.assert_array_equal(buf[:777], opos)
    assert not buf[777:].any()
    assert unpack_rvint(iv, 500.0, float_dtype=dt, posout=False)[0] == 0
    with pytest.raises(AssertionError):
        unpack_rvint(iv.astype(np.int64), 500.0)
    with pytest.raises(ValueError):
        unpack_rvint(iv, 500.0, float_dtype=np.float16)


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_wrapper_unpack_pack9_conventions(fake_engine, oracle, dt):
    from abacusutils_b200.data.pack9 import unpack_pack9

    d = cases.pack9_inputs(21, 5000)
    opos, ovel = oracle.unpack_pack9(d, 1000.0, 321.0, float_dtype=dt)
    pos, vel = unpack_pack9(d, 1000.0, 321.0, float_dtype=dt)
    assert pos.dtype == dt and pos.shape == opos.shape
    np.testing.assert_array_equal(pos, opos)
    np.testing.assert_array_equal(vel, ovel)
    # int8 input, as stored in the ASDF files (datatype: int8), and user-supplied Nmax-row outputs (pack9.py:22-55)
    pbuf, vbuf = np.zeros((5000, 3), dtype=dt), np.zeros((5000, 3), dtype=dt)
    npos, nvel = unpack_pack9(d.view(np.int8), 1000.0, 321.0, float_dtype=dt, posout=pbuf, velout=vbuf)
    assert npos == nvel == len(opos)
    np.testing.assert_array_equal(pbuf[:npos], opos)
    np.testing.assert_array_equal(vbuf[:nvel], ovel)
    assert unpack_pack9(d, 1000.0, 321.0, float_dtype=dt, posout=False, velout=False) == (0, 0)
    # empty stream
    pos, vel = unpack_pack9(np.zeros((0, 9), np.uint8), 1000.0, 321.0, float_dtype=dt)
    assert pos.shape == (0, 3) and vel.shape == (0, 3)


@pytest.mark.parametrize('dt', [np.float32, np.float64])
def test_wrapper_unpack_pids_conventions(fake_engine, oracle, dt):
    from abacusutils_b200.data.bitpacked import unpack_pids

    packed = random_packed_pids(71, 3001)
    want = oracle.unpack_pids(packed, box=720.0, ppd=1440, float_dtype=dt, **{k: True for k in PID_KEYS})
    got = unpack_pids(packed, box=720.0, ppd=1440.0, float_dtype=dt, **{k: True for k in PID_KEYS})
    assert got.keys() == want.keys()
    for k in PID_KEYS:
        assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
        np.testing.assert_array_equal(got[k], want[k], err_msg=k)
    only = unpack_pids(packed, tagged=True, density=True, float_dtype=dt)   # no box / ppd needed (bitpacked.py:200-205)
    assert only.keys() == {'tagged', 'density'}
    np.testing.assert_array_equal(only['density'], want['density'])
    with pytest.raises(ValueError):
        unpack_pids(packed, lagr_pos=True)
    with pytest.raises(ValueError):
        unpack_pids(packed, box=720.0, ppd=12.5, lagr_pos=True)
