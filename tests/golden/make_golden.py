"""
Generate the committed golden fixtures under tests/golden/ (run in the BUILD CONTAINER only).

Two sources, both the reference's own:
  1. the checked-in golden grids of the reference test-suite, tests/ref_tsc/*.asdf
     (/root/reference/tests/test_tsc.py:128-159), decoded with asdf_blsc.py;
  2. outputs of the UNMODIFIED reference modules imported from /root/reference via oracle/ref_shim.py
     on seeded inputs (the inputs are re-generated from the seeds by tests/golden/cases.py, only
     outputs are stored).

Usage:  NUMBA_NUM_THREADS=2 python tests/golden/make_golden.py
(2 threads keeps every reference run on the race-free stripe branch for nmesh >= 12)
"""

import os
import sys
from pathlib import Path

os.environ.setdefault('NUMBA_NUM_THREADS', '2')
HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import numpy as np  # noqa: E402

import cases  # noqa: E402
from asdf_blsc import read_single_array  # noqa: E402
from oracle import ref_shim  # noqa: E402

REF_TESTS = Path(ref_shim.REF_ROOT) / 'tests'


def sparse_planes(a, planes):
    sub = a[planes]
    idx = np.flatnonzero(sub)
    return idx.astype(np.int32), sub.reshape(-1)[idx]


def golden_ref_tsc():
    for n in (10, 256):
        own = read_single_array(REF_TESTS / 'ref_tsc' / f'tsc_ngrid{n}.asdf', np.float32, (n, n, n))
        nbk = read_single_array(REF_TESTS / 'ref_tsc' / f'nbodykit_tsc_ngrid{n}.asdf', np.float32, (n, n, n))
        if n == 10:
            np.savez_compressed(HERE / 'ref_tsc_ngrid10.npz', pydens=own, nbodykit=nbk)
        else:
            planes = cases.REF_TSC_256_PLANES
            oi, ov = sparse_planes(own, planes)
            ni, nv = sparse_planes(nbk, planes)
            np.savez_compressed(
                HERE / 'ref_tsc_ngrid256.npz', planes=planes, own_idx=oi, own_val=ov, nbk_idx=ni, nbk_val=nv,
                own_sum=own.sum(dtype='f8'), own_sumsq=(own.astype('f8') ** 2).sum(), own_nnz=(own != 0).sum(),
                nbk_sum=nbk.sum(dtype='f8'), nbk_sumsq=(nbk.astype('f8') ** 2).sum(),
                own_xsum=own.sum(axis=(1, 2), dtype='f8'), own_zsum=own.sum(axis=(0, 1), dtype='f8'))
        print('ref_tsc', n, own.sum(dtype='f8'))


def golden_reference_runs():
    tsc, ps = ref_shim.load(num_threads=2)
    out = {}

    # A. tsc_parallel on small grids (full grids stored)
    for name, c in cases.TSC_CASES.items():
        pos, w = cases.tsc_inputs(c)
        dens = np.zeros(c['shape'], dtype=np.float32)
        tsc.tsc_parallel(pos, dens, c['box'], weights=w, nthread=c.get('nthread', 2), offset=c['offset'],
                         wrap=c.get('wrap', True))
        out[f'tsc/{name}'] = dens
        print('tsc', name, dens.sum(dtype='f8'))

    # partition_parallel
    for name, c in cases.PARTITION_CASES.items():
        pos, w = cases.tsc_inputs(c)
        ppart, starts, wpart = tsc.partition_parallel(pos, c['npartition'], c['box'], weights=w, nthread=2)
        out[f'partition/{name}/starts'] = starts
        out[f'partition/{name}/ppart'] = ppart
        out[f'partition/{name}/wpart'] = wpart

    # B. bin_kmu on an all-ones mesh: data-independent mode counts
    for name, c in cases.COUNT_CASES.items():
        n, L = c['n'], c['L']
        kedges, muedges = cases.count_edges(c)
        ones = np.ones((n, n, n // 2 + 1), dtype=np.float32)
        poles = np.asarray(c['poles'], dtype=np.int64)
        wc, cnt, wcp, cntp, wk = ps.bin_kmu(n, L, kedges, muedges, ones, poles=poles, nthread=2)
        out[f'counts/{name}/N_mode'] = cnt
        out[f'counts/{name}/N_mode_poles'] = cntp
        out[f'counts/{name}/k_avg'] = wk
        out[f'counts/{name}/power'] = wc
        out[f'counts/{name}/poles'] = wcp
        print('counts', name, cnt.sum())

    # C. calc_power end to end
    for name, c in cases.POWER_CASES.items():
        pos, w, pos2, w2 = cases.power_inputs(c)
        t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'),
                          logk=c['logk'], paste='TSC', nmesh=c['nmesh'], compensated=c['compensated'],
                          interlaced=c['interlaced'], w=w, pos2=pos2, w2=w2, poles=c['poles'], nthread=2)
        for key in t:
            out[f'power/{name}/{key}'] = np.asarray(t[key])
        print('power', name, np.asarray(t['power']).ravel()[:3])

    # D. get_field_fft (materialised delta(k))
    for name, c in cases.FIELD_CASES.items():
        pos, w, _, _ = cases.power_inputs(c)
        W = ps.get_W_compensated(c['L'], c['nmesh'], 'TSC', c['interlaced']) if c['compensated'] else None
        f = ps.get_field_fft(pos, c['L'], c['nmesh'], 'TSC', w, W, c['compensated'], c['interlaced'], nthread=2)
        out[f'field/{name}'] = f
        if W is not None:
            out[f'field/{name}/W'] = W

    # E/F. binning of supplied arrays
    for name, c in cases.DELTAK_CASES.items():
        f1, f2, raw = cases.deltak_inputs(c)
        kedges, muedges = cases.count_edges(c)
        poles = np.asarray(c['poles'], dtype=np.int64)
        P = ps.calc_pk_from_deltak(f1, c['L'], kedges, muedges, field2_fft=f2, poles=poles, nthread=2)
        for key, v in P.items():
            out[f'deltak/{name}/{key}'] = np.asarray(v)
        bp, Np_ = ps.project_3d_to_poles(kedges, raw, c['L'], poles)
        out[f'deltak/{name}/proj_poles'] = bp
        out[f'deltak/{name}/proj_N'] = Np_

    # window tables and edges (host formulas)
    for n, L, inter in [(8, 100.0, True), (8, 100.0, False), (72, 1000.0, True), (128, 1000.0, False),
                        (250, 2000.0, True)]:
        out[f'W/{n}_{int(inter)}'] = ps.get_W_compensated(L, n, 'TSC', inter)
        out[f'Wcic/{n}_{int(inter)}'] = ps.get_W_compensated(L, n, 'CIC', inter)
    for ell in range(0, 11):
        xs = cases.PN_X
        out[f'P_n/{ell}'] = np.array([ps.P_n(np.float32(x), ell) for x in xs], dtype=np.float32)

    np.savez_compressed(HERE / 'reference_runs.npz', **out)
    print('wrote', HERE / 'reference_runs.npz', len(out), 'arrays')


def golden_xi():
    _, ps = ref_shim.load(num_threads=2)
    out = {}
    for name, c in cases.XI_CASES.items():
        Pk, r_bins = cases.xi_inputs(c)
        r_binc, binned_poles, Npoles = ps.pk_to_xi(Pk.copy(), c['L'], r_bins, poles=c['poles'])
        out[f'xi/{name}/r_binc'] = r_binc
        out[f'xi/{name}/binned_poles'] = binned_poles
        out[f'xi/{name}/Npoles'] = Npoles
        print('xi', name, binned_poles[0][:3])
    np.savez_compressed(HERE / 'reference_xi.npz', **out)


def golden_cic():
    import warnings

    _, ps = ref_shim.load(num_threads=2)
    out = {}
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        for name, c in cases.CIC_POWER_CASES.items():
            pos, w, pos2, w2 = cases.power_inputs(c)
            t = ps.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], logk=c['logk'], paste='CIC',
                              nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'], w=w,
                              pos2=pos2, w2=w2, poles=c['poles'], nthread=2)
            for key in t:
                out[f'power/{name}/{key}'] = np.asarray(t[key])
            print('cic power', name, np.asarray(t['power']).ravel()[:3])
        for name, c in cases.CIC_FIELD_CASES.items():
            pos, w = cases.cic_field_inputs(c)
            out[f'field/{name}'] = ps.get_field(pos, c['L'], c['nmesh'], 'CIC', w=w, d=c['d'], nthread=2)
    np.savez_compressed(HERE / 'reference_cic.npz', **out)


def golden_kfields():
    _, ps = ref_shim.load(num_threads=2)
    out = {}
    for name, c in cases.KFIELD_CASES.items():
        delta, k_ell, P_ell = cases.kfield_inputs(c)
        out[f'kf/{name}/delta_mu2'] = ps.get_delta_mu2(delta, c['n'])
        out[f'kf/{name}/smoothing'] = ps.get_smoothing(c['n'], c['L'], c['R'])
        out[f'kf/{name}/expand'] = ps.expand_poles_to_3d(k_ell, P_ell, c['n'], c['L'], np.asarray(c['poles']))
    np.savez_compressed(HERE / 'reference_kfields.npz', **out)
    print('kfields done')


def golden_kppi():
    _, ps = ref_shim.load(num_threads=2)
    out = {}
    for name, c in cases.KPPI_CASES.items():
        w, kedges, pimax = cases.kppi_inputs(c)
        mean, cnt = ps.bin_kppi(c['n'], c['L'], kedges, pimax, c['Npi'], w, dtype=np.dtype(c['dtype']).type,
                                fourier=c['fourier'], nthread=2)
        out[f'kppi/{name}/mean'] = mean
        out[f'kppi/{name}/counts'] = cnt
        print('kppi', name, cnt.sum(), mean.dtype)
    np.savez_compressed(HERE / 'reference_kppi.npz', **out)


def golden_ingest():
    """The reference's own fixtures for the particle decoders (tests/test_data.py:258-325): packed inputs from
    tests/Mini_N64_L32 and the decoded outputs the reference test-suite compares against (tests/ref_data)."""
    import re

    from asdf_blsc import read_asdf_blocks

    def header_value(path, key):
        raw = open(path, 'rb').read()
        y = raw[: raw.find(b'\n...\n')].decode('latin1')
        return float(re.search(rf'\n  {key}: (\S+)', y).group(1))

    out = {}
    sim = REF_TESTS / 'Mini_N64_L32'
    # RVint: halos/z0.000/field_rv_A/field_rv_A_000.asdf -> ref_data/test_read_asdf.asdf (blocks 0, 1 = pos, vel)
    fn = sim / 'halos' / 'z0.000' / 'field_rv_A' / 'field_rv_A_000.asdf'
    rv = np.frombuffer(read_asdf_blocks(fn)[0], dtype='<i4').reshape(-1, 3)
    gb = read_asdf_blocks(REF_TESTS / 'ref_data' / 'test_read_asdf.asdf')
    out['rvint/in'] = rv
    out['rvint/box'] = header_value(fn, 'BoxSize')
    out['rvint/pos'] = np.frombuffer(gb[0], dtype='<f4').reshape(-1, 3)[: len(rv)]
    out['rvint/vel'] = np.frombuffer(gb[1], dtype='<f4').reshape(-1, 3)[: len(rv)]
    # packed PIDs: halos/z0.000/field_pid_A/field_pid_A_000.asdf -> ref_data/test_read_asdf.asdf blocks 2..7
    # (aux, pid, lagr_pos, lagr_idx, tagged, density; tests/test_data.py:303-318)
    fn = sim / 'halos' / 'z0.000' / 'field_pid_A' / 'field_pid_A_000.asdf'
    packed = np.frombuffer(read_asdf_blocks(fn)[0], dtype='<u8')
    n = len(packed)
    out['pids/in'] = packed
    out['pids/box'] = header_value(fn, 'BoxSize')
    out['pids/ppd'] = header_value(fn, 'ppd')
    assert np.array_equal(np.frombuffer(gb[2], dtype='<u8')[:n], packed)
    out['pids/pid'] = np.frombuffer(gb[3], dtype='<i8')[:n]
    out['pids/lagr_pos'] = np.frombuffer(gb[4], dtype='<f4').reshape(-1, 3)[:n]
    out['pids/lagr_idx'] = np.frombuffer(gb[5], dtype='<i2').reshape(-1, 3)[:n]
    out['pids/tagged'] = np.frombuffer(gb[6], dtype='u1')[:n]
    out['pids/density'] = np.frombuffer(gb[7], dtype='<f4')[:n]
    # pack9: slices/z0.000/L0_pack9/slab000.L0.pack9.asdf -> ref_data/test_pack9.asdf; keep the first records only
    fn = sim / 'slices' / 'z0.000' / 'L0_pack9' / 'slab000.L0.pack9.asdf'
    raw = open(fn, 'rb').read()
    nrec = int(re.search(rb'shape: \[(\d+), 9\]', raw).group(1))
    d = np.frombuffer(read_asdf_blocks(fn)[0][: nrec * 9], dtype=np.uint8).reshape(-1, 9)
    gb = read_asdf_blocks(REF_TESTS / 'ref_data' / 'test_pack9.asdf')
    keep = cases.PACK9_GOLDEN_RECORDS
    npart = int((d[:keep, 0] != 0xFF).sum())
    out['pack9/in'] = d[:keep]
    out['pack9/box'] = header_value(fn, 'BoxSize')
    out['pack9/velz'] = header_value(fn, 'VelZSpace_to_kms')
    out['pack9/pos'] = np.frombuffer(gb[0], dtype='<f4').reshape(-1, 3)[:npart]
    out['pack9/vel'] = np.frombuffer(gb[1], dtype='<f4').reshape(-1, 3)[:npart]
    out['pack9/npart_full'] = int((d[:, 0] != 0xFF).sum())
    out['pack9/nrec_full'] = nrec
    np.savez_compressed(HERE / 'ref_ingest.npz', **out)
    print('ingest: rvint', rv.shape, 'pack9', d[:keep].shape, '->', npart, 'particles')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'ingest':
        golden_ingest()
    elif len(sys.argv) > 1 and sys.argv[1] == 'kfields':
        golden_kfields()
    elif len(sys.argv) > 1 and sys.argv[1] == 'kppi':
        golden_kppi()
    elif len(sys.argv) > 1 and sys.argv[1] == 'xi':
        golden_xi()
    elif len(sys.argv) > 1 and sys.argv[1] == 'cic':
        golden_cic()
    else:
        golden_ref_tsc()
        golden_reference_runs()
        golden_xi()
        golden_cic()
        golden_kfields()
        golden_kppi()
        golden_ingest()
