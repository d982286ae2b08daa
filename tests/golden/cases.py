"""
Seeded input definitions shared by make_golden.py (which stores the reference's OUTPUTS) and the
parity tests (which re-create the same inputs and compare the oracle / the CUDA path with them).
"""

import numpy as np

# planes of the 256^3 reference golden grid kept (sparsely) in ref_tsc_ngrid256.npz
REF_TSC_256_PLANES = np.r_[0:12, 122:134, 250:256]

PN_X = np.linspace(0.0, 1.0, 33, dtype=np.float32)


def ref_tsc_inputs(dtype='f4'):
    """Inputs of the reference's test_multi (tests/test_tsc.py:101-109)."""
    rng = np.random.default_rng(234)
    N, box = 10000, 123.0
    pos = rng.random((N, 3), dtype='f4').astype(dtype) * box
    weights = rng.random((N,), dtype='f4').astype(dtype)
    return pos, weights, box


# ---------------------------------------------------------------- TSC
TSC_CASES = {
    'cube24_w': dict(seed=11, N=5000, box=100.0, shape=(24, 24, 24), weighted=True, offset=0.0),
    'cube24_w_off': dict(seed=11, N=5000, box=100.0, shape=(24, 24, 24), weighted=True, offset=0.5 * 100.0 / 24),
    'cube40_now': dict(seed=12, N=20000, box=250.0, shape=(40, 40, 40), weighted=False, offset=0.0),
    'aniso': dict(seed=13, N=4000, box=77.0, shape=(12, 18, 30), weighted=True, offset=0.0),
    'unwrapped': dict(seed=14, N=3000, box=50.0, shape=(20, 20, 20), weighted=True, offset=0.0, spill=True),
    'tiny10': dict(seed=15, N=500, box=123.0, shape=(10, 10, 10), weighted=False, offset=0.0, nthread=1),
}

PARTITION_CASES = {
    'p16': dict(seed=21, N=10000, box=123.0, weighted=True, npartition=16),
    'p1000': dict(seed=22, N=10000, box=123.0, weighted=True, npartition=1000),
}


def tsc_inputs(c):
    rng = np.random.default_rng(c['seed'])
    pos = rng.random((c['N'], 3), dtype='f4') * np.float32(c['box'])
    if c.get('spill'):
        # positions in [-box/2, 3box/2): exercises the one-shot periodic wrap
        pos = pos * np.float32(2.0) - np.float32(0.5 * c['box'])
    w = rng.random(c['N'], dtype='f4') if c['weighted'] else None
    return pos, w


# ---------------------------------------------------------------- mode counts
COUNT_CASES = {
    'n16_k8': dict(n=16, L=100.0, Nk=8, Nmu=1, logk=False, poles=[0, 2, 4]),
    'n32_k16_mu4': dict(n=32, L=500.0, Nk=16, Nmu=4, logk=False, poles=[0, 2, 4]),
    'n48_k12_mu3_log': dict(n=48, L=750.0, Nk=12, Nmu=3, logk=True, poles=[0, 2]),
    'n64_default': dict(n=64, L=1000.0, Nk=64, Nmu=1, logk=False, poles=[]),
    'n72_testpower': dict(n=72, L=1000.0, Nk=36, Nmu=4, logk=False, poles=[0, 2, 4],
                          k_max=np.pi * 72 / 1000.0 + 1e-6, k_lo=1e-6),
    'n128_k100_mu10': dict(n=128, L=1000.0, Nk=100, Nmu=10, logk=False, poles=[0, 2, 4]),
    'n90_k45_mu5': dict(n=90, L=333.0, Nk=45, Nmu=5, logk=False, poles=[2]),
    'n50_halfk': dict(n=50, L=200.0, Nk=20, Nmu=2, logk=False, poles=[0, 4], k_max=0.5 * np.pi * 50 / 200.0),
    'n40_log_mu7': dict(n=40, L=1234.5, Nk=25, Nmu=7, logk=True, poles=[0, 1, 2, 3]),
}


def count_edges(c):
    n, L = c['n'], c['L']
    k_max = c.get('k_max', np.pi * n / L)
    if c['logk']:
        k_min = (1.0 - 1.0e-4) * 2.0 * np.pi / L
        kedges = np.geomspace(k_min, k_max, c['Nk'] + 1)
    else:
        kedges = np.linspace(c.get('k_lo', 0.0), k_max, c['Nk'] + 1)
    muedges = np.linspace(0.0, 1.0, c['Nmu'] + 1)
    return kedges, muedges


# ---------------------------------------------------------------- calc_power / get_field_fft
def _pc(seed, N, L, nmesh, kbins, mubins, poles, compensated, interlaced, weighted=True, cross=False,
        logk=False, **kw):
    return dict(seed=seed, N=N, L=L, nmesh=nmesh, kbins=kbins, mubins=mubins, poles=poles,
                compensated=compensated, interlaced=interlaced, weighted=weighted, cross=cross, logk=logk, **kw)


POWER_CASES = {
    'n32_ci': _pc(31, 20000, 500.0, 32, 16, 4, [0, 2, 4], True, True),
    'n32_c': _pc(31, 20000, 500.0, 32, 16, 4, [0, 2, 4], True, False),
    'n32_i': _pc(31, 20000, 500.0, 32, 16, 4, [0, 2, 4], False, True),
    'n32_raw': _pc(31, 20000, 500.0, 32, 16, 4, [0, 2, 4], False, False),
    'n32_cross_ci': _pc(32, 15000, 500.0, 32, 16, 4, [0, 2, 4], True, True, cross=True),
    'n32_cross_c': _pc(32, 15000, 500.0, 32, 16, 4, [0, 2, 4], True, False, cross=True),
    'n48_log': _pc(33, 30000, 750.0, 48, 12, 3, [0, 2], True, True, logk=True),
    'n40_defaults': _pc(34, 25000, 1000.0, 40, None, None, None, True, True, weighted=False),
    'n36_nopoles_mu': _pc(35, 10000, 300.0, 36, 18, 5, None, True, False, weighted=False),
    # BASELINE config 1 scaled down 8x in particle count (same L/nmesh/options)
    'cfg1_small': _pc(12345, 125000, 1000.0, 128, None, None, [0, 2, 4], True, False, weighted=False),
}

FIELD_CASES = {
    'f24_ci': _pc(41, 3000, 100.0, 24, 0, 0, None, True, True),
    'f24_c': _pc(41, 3000, 100.0, 24, 0, 0, None, True, False),
    'f24_i': _pc(41, 3000, 100.0, 24, 0, 0, None, False, True, weighted=False),
    'f24_raw': _pc(41, 3000, 100.0, 24, 0, 0, None, False, False, weighted=False),
}


def power_inputs(c):
    rng = np.random.default_rng(c['seed'])
    L = np.float32(c['L'])
    pos = rng.random((c['N'], 3), dtype='f4') * L
    w = rng.random(c['N'], dtype='f4') if c['weighted'] else None
    pos2 = w2 = None
    if c['cross']:
        N2 = c['N'] // 2 + 17
        pos2 = rng.random((N2, 3), dtype='f4') * L
        # correlate the second catalogue with the first so the cross-spectrum is not pure noise
        pos2[: N2 // 2] = np.mod(pos[: N2 // 2] + rng.normal(0, 0.01 * c['L'], (N2 // 2, 3)).astype('f4'), L)
        pos2 = np.ascontiguousarray(pos2, dtype=np.float32)
        pos2[pos2 >= L] = 0.0
        w2 = rng.random(N2, dtype='f4') if c['weighted'] else None
    return pos, w, pos2, w2


# ---------------------------------------------------------------- binning of supplied arrays
DELTAK_CASES = {
    'd32': dict(seed=51, n=32, L=400.0, Nk=14, Nmu=4, logk=False, poles=[0, 2, 4]),
    'd30_allpoles': dict(seed=52, n=30, L=90.0, Nk=10, Nmu=3, logk=False, poles=[0, 1, 2, 3, 4, 6, 8, 10]),
    'd20_log': dict(seed=53, n=20, L=50.0, Nk=9, Nmu=1, logk=True, poles=[0, 2]),
}


def deltak_inputs(c):
    rng = np.random.default_rng(c['seed'])
    n = c['n']
    shp = (n, n, n // 2 + 1)
    f1 = (rng.standard_normal(shp, dtype='f4') + 1j * rng.standard_normal(shp, dtype='f4')).astype(np.complex64)
    f2 = (0.5 * f1 + rng.standard_normal(shp, dtype='f4') + 1j * rng.standard_normal(shp, dtype='f4')).astype(
        np.complex64)
    raw = (rng.random(shp, dtype='f4') * 10).astype(np.float32)
    return f1, f2, raw


# ---------------------------------------------------------------- xi(r) multipoles (pk_to_xi)
XI_CASES = {
    'x24': dict(seed=61, n=24, L=120.0, Nr=10, r_max=50.0, poles=[0, 2, 4]),
    'x30': dict(seed=62, n=30, L=300.0, Nr=14, r_max=150.0, poles=[0, 2]),
}


def xi_inputs(c):
    """A physical P(k) mesh: |rfftn(real field)|^2, so P(-k) = P(k) holds on the k_z = 0 / Nyquist planes
    (irfftn of a mesh without that symmetry is implementation-defined)."""
    rng = np.random.default_rng(c['seed'])
    n = c['n']
    field = rng.standard_normal((n, n, n)).astype(np.float32)
    dk = np.fft.rfftn(field.astype(np.float64)) / n**3
    Pk = (np.abs(dk) ** 2).astype(np.float32)
    r_bins = np.linspace(0.0, c['r_max'], c['Nr'] + 1)
    return Pk, r_bins


# ---------------------------------------------------------------- CIC paste (analysis/cic.py via paste='CIC')
CIC_POWER_CASES = {
    'cic32_ci': _pc(71, 20000, 500.0, 32, 16, 4, [0, 2, 4], True, True),
    'cic32_c': _pc(71, 20000, 500.0, 32, 16, 4, [0, 2, 4], True, False),
    'cic32_raw': _pc(72, 20000, 500.0, 32, 16, 4, [0, 2], False, False, weighted=False),
    'cic40_cross_i': _pc(73, 15000, 400.0, 40, 10, 2, [0, 2, 4], False, True, cross=True),
}
CIC_FIELD_CASES = {
    'cicf24': dict(seed=74, N=4000, L=100.0, nmesh=24, weighted=True, d=0.0),
    'cicf24_off': dict(seed=74, N=4000, L=100.0, nmesh=24, weighted=True, d=0.5 * 100.0 / 24),
}


def cic_field_inputs(c):
    rng = np.random.default_rng(c['seed'])
    # stay inside [0, L - d): cic_serial applies no periodic wrap
    pos = rng.random((c['N'], 3), dtype='f4') * np.float32(c['L'] - 2 * c['L'] / c['nmesh'])
    w = rng.random(c['N'], dtype='f4') if c['weighted'] else None
    return pos, w


# ---------------------------------------------------------------- ZCV k-space helpers
KFIELD_CASES = {
    'k16': dict(seed=81, n=16, L=200.0, R=7.5, Nk=12, poles=[0, 2, 4]),
    'k18': dict(seed=82, n=18, L=90.0, R=3.0, Nk=20, poles=[0, 1, 2]),
}


def kfield_inputs(c):
    rng = np.random.default_rng(c['seed'])
    n = c['n']
    shp = (n, n, n // 2 + 1)
    delta = (rng.standard_normal(shp, dtype='f4') + 1j * rng.standard_normal(shp, dtype='f4')).astype(np.complex64)
    k_ny = np.pi * n / c['L']
    k_ell = np.linspace(0.1 * k_ny, 1.2 * k_ny, c['Nk'])   # starts above 0 and ends below the mesh corner: both clamps hit
    P_ell = (rng.random((len(c['poles']), c['Nk'])) * 100).astype(np.float32)
    return delta, k_ell, P_ell


# ---------------------------------------------------------------- 2-D TSC (tsc.py:57-62, 452-468)
TSC2D_CASES = {
    'sq36': dict(seed=91, N=3000, box=60.0, shape=(36, 36), weighted=True, offset=0.0, ncol=3),
    'rect_off': dict(seed=92, N=2500, box=45.0, shape=(20, 48), weighted=False, offset=0.4, ncol=3),
}


def tsc2d_inputs(c):
    rng = np.random.default_rng(c['seed'])
    pos = rng.random((c['N'], c['ncol']), dtype='f4') * np.float32(c['box'])
    w = rng.random(c['N'], dtype='f4') if c['weighted'] else None
    return pos, w


# ---------------------------------------------------------------- bin_kppi (power_spectrum.py:303-412)
# kmax / pimax are in units of the mesh Nyquist scale (pi n / L in Fourier space, L / 2 in real space)
KPPI_CASES = {
    'f16': dict(seed=101, n=16, L=100.0, k0=0.0, kmax=1.0, Nk=5, pimax=0.6, Npi=4, fourier=True, dtype='f4'),
    'f16_break': dict(seed=102, n=16, L=100.0, k0=0.0, kmax=0.6, Nk=5, pimax=1.2, Npi=7, fourier=True, dtype='f4'),
    'f15_odd': dict(seed=103, n=15, L=50.0, k0=0.05, kmax=0.8, Nk=6, pimax=0.5, Npi=3, fourier=True, dtype='f4'),
    'f72': dict(seed=104, n=72, L=300.0, k0=0.02, kmax=0.9, Nk=11, pimax=0.7, Npi=9, fourier=True, dtype='f4'),
    'f72_wide': dict(seed=105, n=72, L=300.0, k0=0.0, kmax=1.5, Nk=40, pimax=1.01, Npi=37, fourier=True, dtype='f4'),
    'r24': dict(seed=106, n=24, L=200.0, k0=0.0, kmax=1.1, Nk=8, pimax=0.6, Npi=6, fourier=False, dtype='f4'),
    'f20_f64': dict(seed=107, n=20, L=50.0, k0=0.1, kmax=0.9, Nk=6, pimax=1.0, Npi=3, fourier=True, dtype='f8'),
    'f40_1bin': dict(seed=108, n=40, L=80.0, k0=0.0, kmax=2.0, Nk=1, pimax=1.01, Npi=1, fourier=True, dtype='f4'),
}


def kppi_inputs(c):
    rng = np.random.default_rng(c['seed'])
    n = c['n']
    shape = (n, n, n // 2 + 1) if c['fourier'] else (n, n, n)
    w = rng.standard_normal(shape).astype(c['dtype']) + np.dtype(c['dtype']).type(0.25)
    scale = np.pi * n / c['L'] if c['fourier'] else c['L'] / 2
    kedges = np.linspace(c['k0'] * scale, c['kmax'] * scale, c['Nk'] + 1)
    return w, kedges, c['pimax'] * scale


# ---------------------------------------------------------------- particle decoders (data/bitpacked.py, data/pack9.py)
PACK9_GOLDEN_RECORDS = 12000   # prefix of the reference's slab000.L0.pack9 fixture kept in ref_ingest.npz


def rvint_inputs(seed, N):
    rng = np.random.default_rng(seed)
    return rng.integers(-2**31, 2**31, size=(N, 3), dtype=np.int64).astype(np.int32)


def pack9_inputs(seed, nrec, hdr_frac=0.05, cpd=875, first_header=True):
    """A synthetic pack9 stream: random particle records with cell headers (first byte 0xFF, fields 1..5 =
    cpd, velocity scale, cell i/j/k, each stored + 2048 in 12 bits) sprinkled in."""
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, size=(nrec, 9), dtype=np.int64).astype(np.uint8)
    d[d[:, 0] == 0xFF, 0] = 0xFE
    is_hdr = rng.random(nrec) < hdr_frac
    if nrec:
        is_hdr[0] = first_header
    idx = np.flatnonzero(is_hdr)

    def setfield(f, val):
        val = (val + 2048).astype(np.int64)
        b = 3 * (f // 2)
        if f % 2 == 0:
            d[idx, b] = (val >> 4) & 0xFF
            d[idx, b + 1] = (d[idx, b + 1] & 0xF0) | (val & 0xF)
        else:
            d[idx, b + 1] = (d[idx, b + 1] & 0x0F) | (((val >> 8) & 0xF) << 4)
            d[idx, b + 2] = val & 0xFF

    nh = len(idx)
    setfield(1, np.full(nh, cpd - 2000))
    setfield(2, rng.integers(-1500, 2000, nh))
    for f in (3, 4, 5):
        setfield(f, rng.integers(-2000, cpd - 2000, nh))
    d[idx, 0] = 0xFF
    return d


# ---------------------------------------------------------------- float64 (tests/golden/make_golden_f64.py)
def f64_inputs(seed=71, N=6000, box=77.0):
    """float64 positions whose low bits matter (not representable in float32), some outside [0, box); float64 weights."""
    rng = np.random.default_rng(seed)
    pos = rng.random((N, 3)) * box
    pos[:40] += box * rng.integers(-1, 2, size=(40, 3))
    pos = np.clip(pos, -0.999 * box, 1.999 * box)
    w = rng.random(N) + 0.5
    return pos, w


F64_POWER = dict(seed=72, N=30000, L=600.0, nmesh=36, kbins=14, mubins=3, poles=[0, 2, 4])
