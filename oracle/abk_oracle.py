"""
CPU oracle for the TSC + power-spectrum path (TEST INFRASTRUCTURE ONLY).

Python driver over ``oracle/abk_oracle.c`` (C + OpenMP) and ``scipy.fft.rfftn`` (the same
third-party FFT the reference calls, power_spectrum.py:980,986,1059).  The function names and
signatures mirror the reference (``abacusnbody.analysis.tsc`` / ``.power_spectrum``) so parity
tests read like the reference's own tests.  Each function cites the reference lines it restates
(paths relative to /root/reference/abacusnbody/analysis/).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  The product package (``abacusutils_b200``) never does.

Parity status: PINNED (see the header of abk_oracle.c and tests/test_oracle_*.py).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
import warnings
from pathlib import Path

import numpy as np
from scipy.fft import rfftn

_HERE = Path(__file__).resolve().parent
_LIB = None

MAX_THREADS = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else os.cpu_count()


def build(force=False):
    """Compile libabk_oracle.so with the committed Makefile (gcc + OpenMP)."""
    so = _HERE / 'libabk_oracle.so'
    src = _HERE / 'abk_oracle.c'
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(['make', '-C', str(_HERE), '-B', 'libabk_oracle.so'], check=True,
                       capture_output=True)
    return so


def _lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = _HERE / 'libabk_oracle.so'
    if not so.exists():
        build()
    L = C.CDLL(str(so))
    fp, dp, ip = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_int64)
    vp = C.c_void_p
    L.abko_max_threads.restype = C.c_int
    L.abko_wrap_inplace_f32.argtypes = [vp, C.c_int64, C.c_double, C.c_int]
    L.abko_partition_f32.argtypes = [vp, vp, C.c_int64, C.c_int, C.c_double, C.c_int, vp, vp, vp, C.c_int]
    L.abko_partition_f32.restype = C.c_int
    L.abko_tsc_scatter_f32.argtypes = [vp, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    L.abko_tsc_stripes_f32.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int]
    L.abko_cic_serial_f32.argtypes = [vp, vp, C.c_int64, vp, C.c_int, C.c_int, C.c_int, C.c_double]
    L.abko_normalize_field_f32.argtypes = [vp, C.c_int64, C.c_double, C.c_int]
    L.abko_scale_c64.argtypes = [vp, C.c_int64, C.c_float, C.c_int]
    L.abko_shift_field_fft.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double, C.c_int]
    L.abko_compensate.argtypes = [vp, vp, C.c_int, C.c_int]
    L.abko_raw_power.argtypes = [vp, vp, vp, C.c_int64, C.c_int]
    L.abko_P_n.argtypes = [C.c_float, C.c_int]
    L.abko_P_n.restype = C.c_float
    L.abko_bin_kmu.argtypes = [C.c_int, C.c_double, vp, C.c_int, vp, C.c_int, vp, C.c_int64, vp, C.c_int,
                               C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
    L.abko_bin_kmu.restype = C.c_int
    del fp, dp, ip
    _LIB = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data


def _nthread(nthread):
    return MAX_THREADS if (nthread is None or nthread < 0) else int(nthread)


class Table(dict):
    """Minimal stand-in for astropy.table.Table (power_spectrum.py:1318)."""

    def __init__(self, d, meta=None):
        super().__init__(d)
        self.meta = meta or {}


# ----------------------------------------------------------------------------- TSC
def choose_npartition(n1d, nthread):
    """Stripe count rule of tsc.py:126-139, restricted to the race-free branch.

    The reference may pick ``n1d//2`` stripes (2 cells wide), which has a lost-update race
    (SURVEY.md section 5); the oracle never does: it caps at ``n1d//3``.
    """
    if nthread <= 1:
        return 1
    npart = min(n1d // 3, 2 * nthread)
    npart = 2 * (npart // 2)
    return max(npart, 1)


def partition_parallel(pos, npartition, boxsize, weights=None, coord=0, nthread=-1, sort=False):
    """tsc.py:259-384.  Returns (psort, starts int64[npartition+1], wsort|None)."""
    assert pos.shape[1] == 3
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
    psort = np.empty_like(pos)
    wsort = None if w is None else np.empty_like(w)
    starts = np.empty(npartition + 1, dtype=np.int64)
    rc = _lib().abko_partition_f32(_p(pos), _p(w), len(pos), int(npartition), float(boxsize), int(coord),
                                   _p(psort), _p(wsort), _p(starts), _nthread(nthread))
    assert rc == 0
    if sort:
        for i in range(npartition):
            sl = slice(starts[i], starts[i + 1])
            iord = psort[sl, coord].argsort()
            psort[sl] = psort[sl][iord]
            if wsort is not None:
                wsort[sl] = wsort[sl][iord]
    return psort, starts, wsort


def tsc_scatter_serial(pos, dens, box, weights=None, offset=0.0):
    """Serial 27-point scatter, the restatement of ``_tsc_scatter`` (tsc.py:394-507)."""
    pos = np.ascontiguousarray(pos, dtype=np.float32)
    w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float32)
    assert dens.dtype == np.float32 and dens.flags.c_contiguous and dens.ndim == 3
    gx, gy, gz = dens.shape
    _lib().abko_tsc_scatter_f32(_p(pos), _p(w), len(pos), _p(dens), gx, gy, gz, float(box), float(offset))


def tsc_parallel(pos, densgrid, box, weights=None, nthread=-1, wrap=True, npartition=None, sort=False,
                 coord=0, verbose=False, offset=0.0):
    """tsc.py:10-206 (float32, 3-D, coord=0 only)."""
    nthread = _nthread(nthread)
    if isinstance(densgrid, (int, np.integer)):
        densgrid = (densgrid, densgrid, densgrid)
    if isinstance(densgrid, tuple):
        densgrid = np.zeros(densgrid, dtype=np.float32)
        user_supplied_grid = False
    else:
        user_supplied_grid = True
    assert pos.dtype == np.float32 and pos.flags.c_contiguous, 'oracle is float32-only'
    assert densgrid.dtype == np.float32 and densgrid.flags.c_contiguous
    assert coord == 0
    L = _lib()
    if wrap:
        L.abko_wrap_inplace_f32(_p(pos), len(pos), float(box), nthread)
    gx, gy, gz = densgrid.shape
    if not npartition:
        npartition = choose_npartition(gx, nthread)
    if npartition > 1:
        psort, starts, wsort = partition_parallel(pos, npartition, box, weights=weights, nthread=nthread,
                                                  coord=coord, sort=sort)
        L.abko_tsc_stripes_f32(_p(psort), _p(wsort), _p(starts), npartition, _p(densgrid), gx, gy, gz,
                               float(box), float(offset), nthread)
    else:
        tsc_scatter_serial(pos, densgrid, box, weights=weights, offset=offset)
    if user_supplied_grid:
        return None
    return densgrid


# ----------------------------------------------------------------------------- k-space helpers
def get_k_mu_edges(Lbox, k_max, kbins, mubins, logk):
    """power_spectrum.py:663-704."""
    if isinstance(kbins, (int, np.integer)):
        if logk:
            k_min = (1.0 - 1.0e-4) * 2.0 * np.pi / Lbox
            kbins = np.geomspace(k_min, k_max, kbins + 1)
        else:
            kbins = np.linspace(0.0, k_max, kbins + 1)
    if isinstance(mubins, (int, np.integer)):
        mubins = np.linspace(0.0, 1.0, mubins + 1)
    return kbins, mubins


def get_W_compensated(Lbox, nmesh, paste, interlaced):
    """power_spectrum.py:1081-1128 (TSC and CIC windows; float32 wavenumbers)."""
    d = Lbox / nmesh
    kN = np.pi / d
    k = (np.fft.fftfreq(nmesh, d=d) * 2.0 * np.pi).astype(np.float32)
    paste = paste.upper()
    if paste not in ('TSC', 'CIC'):
        raise ValueError(f'Unknown pasting method {paste}')
    if interlaced:
        p = 3.0 if paste == 'TSC' else 2.0
        W = np.sinc(0.5 * k / kN) ** p
    else:
        s = np.sin(0.5 * np.pi * k / kN) ** 2
        if paste == 'TSC':
            W = (1 - s + 2.0 / 15 * s**2) ** 0.5
        else:
            W = (1 - 2.0 / 3 * s) ** 0.5
    return W


def normalize_field(field, tot_weight=None, inplace=False, nthread=MAX_THREADS):
    """power_spectrum.py:860-901."""
    if tot_weight is None:
        tot_weight = field.sum()
    if not inplace:
        field = field.copy()
    assert field.dtype == np.float32 and field.flags.c_contiguous
    _lib().abko_normalize_field_f32(_p(field), field.size, float(tot_weight), _nthread(nthread))
    return field


def get_field(pos, Lbox, nmesh, paste, w=None, d=0.0, nthread=MAX_THREADS, dtype=np.float32):
    """power_spectrum.py:808-857.  Normalised by len(pos), not sum(w) (:856)."""
    if w is not None:
        assert pos.shape[0] == len(w)
    field = np.zeros((nmesh, nmesh, nmesh), dtype=np.float32)
    if paste.upper() == 'CIC':
        # power_spectrum.py:846-853: cic_serial(pos + d, ...) -- a float32 sum, no periodic wrap
        p = np.ascontiguousarray(pos + d if d != 0.0 else pos, dtype=np.float32)
        wc = None if w is None else np.ascontiguousarray(w, dtype=np.float32)
        _lib().abko_cic_serial_f32(_p(p), _p(wc), len(p), _p(field), nmesh, nmesh, nmesh, float(Lbox))
    elif paste.upper() == 'TSC':
        tsc_parallel(pos, field, Lbox, weights=w, nthread=nthread, offset=d)
    else:
        raise ValueError(f'Unknown pasting method: {paste}')
    normalize_field(field, inplace=True, tot_weight=len(pos), nthread=nthread)
    return field


def shift_field_fft(field_fft, field_shift_fft, n1d, L, d, nthread=MAX_THREADS):
    """power_spectrum.py:904-948, in place on field_fft."""
    assert field_fft.dtype == np.complex64 and field_shift_fft.dtype == np.complex64
    _lib().abko_shift_field_fft(_p(field_fft), _p(field_shift_fft), int(n1d), float(L), float(d),
                                _nthread(nthread))


def get_interlaced_field_fft(pos, Lbox, nmesh, paste, w, nthread=MAX_THREADS, verbose=False):
    """power_spectrum.py:951-998."""
    d = Lbox / nmesh
    field = get_field(pos, Lbox, nmesh, paste, w, nthread=nthread)
    field_fft = rfftn(field, workers=nthread)
    del field
    field_shift = get_field(pos, Lbox, nmesh, paste, w, d=0.5 * d, nthread=nthread)
    field_shift_fft = rfftn(field_shift, workers=nthread)
    del field_shift
    shift_field_fft(field_fft, field_shift_fft, nmesh, Lbox, d, nthread=nthread)
    return field_fft


def get_field_fft(pos, Lbox, nmesh, paste, w, W, compensated, interlaced, nthread=MAX_THREADS, verbose=False,
                  dtype=np.float32):
    """power_spectrum.py:1001-1070."""
    nthread = _nthread(nthread)
    if interlaced:
        field_fft = get_interlaced_field_fft(pos, Lbox, nmesh, paste, w, nthread=nthread)
    else:
        field = get_field(pos, Lbox, nmesh, paste, w, nthread=nthread)
        inv_size = np.float32(1 / field.size)
        field_fft = rfftn(field, overwrite_x=True, workers=nthread)
        _lib().abko_scale_c64(_p(field_fft), field_fft.size, inv_size, nthread)
    if compensated:
        assert W is not None
        Wf = np.ascontiguousarray(W, dtype=np.float32)
        _lib().abko_compensate(_p(field_fft), _p(Wf), int(nmesh), nthread)
    return field_fft


def get_raw_power(field_fft, field2_fft=None, nthread=MAX_THREADS):
    """power_spectrum.py:707-727."""
    out = np.empty(field_fft.shape, dtype=np.float32)
    f1 = np.ascontiguousarray(field_fft, dtype=np.complex64)
    f2 = None if field2_fft is None else np.ascontiguousarray(field2_fft, dtype=np.complex64)
    _lib().abko_raw_power(_p(f1), _p(f2), _p(out), f1.size, _nthread(nthread))
    return out


def P_n(x, n):
    """power_spectrum.py:121-147."""
    return float(_lib().abko_P_n(float(x), int(n)))


def bin_kmu(n1d, L, kedges, muedges, weights, poles=np.empty(0, 'i8'), dtype=np.float32, fourier=True,
            nthread=MAX_THREADS, acc64=False, raw=False):
    """power_spectrum.py:150-300.

    Returns (weighted_counts f32 (Nk,Nmu), counts i64, weighted_counts_poles f32 (Np,Nk),
    counts_poles i64 (Nk,), weighted_counts_k f32 (Nk,Nmu)).  ``acc64`` accumulates in double
    (tighter checker); ``raw=True`` returns the float64 means before the float32 cast.
    """
    kedges = np.ascontiguousarray(kedges, dtype=np.float64)
    muedges = np.ascontiguousarray(muedges, dtype=np.float64)
    weights = np.ascontiguousarray(weights, dtype=np.float32)
    poles = np.ascontiguousarray(poles, dtype=np.int64)
    Nk, Nmu, Np = len(kedges) - 1, len(muedges) - 1, len(poles)
    assert weights.ndim == 3 and weights.shape[0] == n1d and weights.shape[1] == n1d
    counts = np.zeros((Nk, Nmu), dtype=np.int64)
    sw = np.zeros((Nk, Nmu), dtype=np.float64)
    sk = np.zeros((Nk, Nmu), dtype=np.float64)
    sp = np.zeros((Np, Nk), dtype=np.float64)
    rc = _lib().abko_bin_kmu(int(n1d), float(L), _p(kedges), Nk, _p(muedges), Nmu, _p(weights),
                             weights.shape[2], _p(poles), Np, int(bool(fourier)), int(bool(acc64)),
                             _nthread(nthread), _p(counts), _p(sw), _p(sk), _p(sp))
    assert rc == 0
    counts_poles = counts.sum(axis=1)
    for ip, pole in enumerate(poles):
        if pole == 0:
            sp[ip] = sw.sum(axis=1)
    nz = counts != 0
    sw[nz] /= counts[nz]
    sk[nz] /= counts[nz]
    nzp = counts_poles != 0
    sp[:, nzp] /= counts_poles[nzp]
    if raw:
        return sw, counts, sp, counts_poles, sk
    return sw.astype(np.float32), counts, sp.astype(np.float32), counts_poles, sk.astype(np.float32)


def bin_kppi(n1d, L, kedges, pimax, Npi, weights, dtype=np.float32, fourier=True, nthread=MAX_THREADS, raw=False):
    """power_spectrum.py:303-412 (NumPy restatement, row-vectorised).

    Reference behaviour kept on purpose:
      * kperp^2 = dtype(i'^2 + j'^2) with (lo, hi] bins, skipped below kedges2[0] (:371-386);
      * the j loop *breaks* at the first j whose kperp^2 >= kedges2[-1] (:379-380), so for a row i
        whose peak value (at j = n//2) is out of range every j >= that first failing j is dropped,
        including the mirror modes j > n//2 that would be back in range;
      * kz^2 stays an integer and is compared with the float edges after promotion to float64
        (:388-395); modes with kz^2 >= piedges2[-1] are dropped; multiplicity 1 for k == 0, else 2.
    Returns (weighted_counts dtype (Nk,Npi), counts i64); ``raw=True`` keeps the float64 means.
    """
    dtype = np.dtype(dtype).type
    weights = np.asarray(weights)
    kzlen = n1d // 2 + 1
    Nk = len(kedges) - 1
    dk = 2.0 * np.pi / L if fourier else L / n1d
    kedges2 = ((np.asarray(kedges, dtype=np.float64) / dk) ** 2).astype(dtype)
    piedges2 = ((np.linspace(0.0, pimax, Npi + 1) / dk) ** 2).astype(dtype)
    counts = np.zeros((Nk, Npi), dtype=np.int64)
    sums = np.zeros((Nk, Npi), dtype=np.float64)

    kz2 = (np.arange(kzlen, dtype=np.int64) ** 2).astype(np.float64)
    pie = piedges2.astype(np.float64)
    kuse = kz2 < pie[-1]
    bpi = np.searchsorted(pie[1:], kz2[kuse], side='left')
    mult = np.where(np.arange(kzlen) == 0, 1, 2)[kuse]
    fold = np.where(np.arange(n1d) < n1d // 2, np.arange(n1d), np.arange(n1d) - n1d).astype(np.int64)
    for i in range(n1d):
        kp2 = (fold[i] ** 2 + fold ** 2).astype(dtype)
        over = np.flatnonzero(kp2 >= kedges2[-1])
        jend = over[0] if len(over) else n1d
        js = np.flatnonzero(kp2[:jend] >= kedges2[0])
        if len(js) == 0 or len(bpi) == 0:
            continue
        bk = np.searchsorted(kedges2[1:], kp2[js], side='left')
        flat = (bk[:, None] * Npi + bpi[None, :]).ravel()
        wrow = weights[i, js][:, :kzlen][:, kuse].astype(np.float64) * mult[None, :]
        counts += np.bincount(flat, weights=np.broadcast_to(mult, (len(js), len(mult))).ravel(),
                              minlength=Nk * Npi).astype(np.int64).reshape(Nk, Npi)
        sums += np.bincount(flat, weights=wrow.ravel(), minlength=Nk * Npi).reshape(Nk, Npi)
    nz = counts != 0
    sums[nz] /= counts[nz]
    return (sums if raw else sums.astype(dtype)), counts


def project_3d_to_poles(k_bin_edges, raw_p3d, Lbox, poles):
    """power_spectrum.py:415-447."""
    assert np.max(poles) <= 10, 'numba implementation works up to ell = 10'
    nmesh = raw_p3d.shape[0]
    poles = np.asarray(poles)
    raw_p3d = np.asarray(raw_p3d)
    muedges = np.array([0.0, 1.0])
    _, _, binned_poles, Npoles, _ = bin_kmu(nmesh, Lbox, k_bin_edges, muedges, raw_p3d, poles=poles)
    binned_poles *= Lbox**3
    return binned_poles, Npoles


def _kgrid(n1d):
    f = np.fft.fftfreq(n1d, 1.0 / n1d).astype(np.int64)
    fi = np.where(np.arange(n1d) < n1d // 2, np.arange(n1d), np.arange(n1d) - n1d)
    k = np.arange(n1d // 2 + 1)
    kmag2 = (fi[:, None, None] ** 2 + fi[None, :, None] ** 2 + k[None, None, :] ** 2).astype(np.float32)
    with np.errstate(divide='ignore', invalid='ignore'):
        mu2 = np.where(kmag2 > 0, (k[None, None, :] ** 2).astype(np.float32) / kmag2, np.float32(0)).astype(np.float32)
    del f
    return kmag2, mu2


def get_delta_mu2(delta, n1d):
    """power_spectrum.py:577-617 (NumPy restatement)."""
    _, mu2 = _kgrid(n1d)
    return (delta * mu2).astype(np.complex64)


def get_smoothing(n1d, L, R):
    """power_spectrum.py:539-574 (NumPy restatement)."""
    kmag2, _ = _kgrid(n1d)
    dk = np.float32(2.0 * np.pi / L)
    dk2 = np.float32(dk**2)
    R2 = np.float32(R**2)
    t = (-kmag2 * dk2) * R2
    return np.exp(t.astype(np.float64) / 2.0).astype(np.float32)


def expand_poles_to_3d(k_ell, P_ell, n1d, L, poles):
    """power_spectrum.py:450-536 (NumPy restatement; P_l evaluated with abko_P_n)."""
    kmag2, mu2 = _kgrid(n1d)
    dk = np.float32(2.0 * np.pi / L)
    k_ell = np.asarray(k_ell, dtype=np.float32)
    P_ell = np.asarray(P_ell, dtype=np.float32)
    xd = np.sqrt(kmag2) * dk
    dx = k_ell[1] - k_ell[0]
    f = (xd - k_ell[0]) / dx
    fl = np.clip(f.astype(np.int64), 0, len(k_ell) - 2)
    out = np.zeros_like(kmag2)
    pn = np.vectorize(lambda x, ell: P_n(x, ell), otypes=[np.float32])
    for ip, ell in enumerate(poles):
        y = P_ell[ip]
        yd = y[fl] + (f - fl).astype(np.float32) * (y[fl + 1] - y[fl])
        yd = np.where(xd <= k_ell[0], y[0], np.where(xd >= k_ell[-1], y[-1], yd)).astype(np.float32)
        out += yd if ell == 0 else yd * pn(mu2, int(ell))
    return out


def pk_to_xi(Pk, Lbox, r_bins, poles=[0, 2, 4]):
    """power_spectrum.py:620-660."""
    from scipy.fft import irfftn

    Xi = irfftn(Pk, workers=-1).real
    r_bins = np.asarray(r_bins)
    r_binc = (r_bins[1:] + r_bins[:-1]) * 0.5
    nmesh = Xi.shape[0]
    poles = np.asarray(poles)
    muedges = np.array([0.0, 1.0])
    _, _, binned_poles, Npoles, _ = bin_kmu(nmesh, Lbox, r_bins, muedges, Xi, poles=poles, fourier=False)
    binned_poles *= nmesh**3
    return r_binc, binned_poles, Npoles


def calc_pk_from_deltak(field_fft, Lbox, k_bin_edges, mu_bin_edges, field2_fft=None, poles=np.empty(0, 'i8'),
                        squeeze_mu_axis=True, nthread=MAX_THREADS, acc64=False):
    """power_spectrum.py:730-805."""
    raw_p3d = get_raw_power(field_fft, field2_fft, nthread=nthread)
    nmesh = raw_p3d.shape[0]
    power, N_mode, binned_poles, N_mode_poles, k_avg = bin_kmu(nmesh, Lbox, k_bin_edges, mu_bin_edges, raw_p3d,
                                                               poles, nthread=nthread, acc64=acc64)
    power *= Lbox**3
    if len(poles) > 0:
        binned_poles *= Lbox**3
    if squeeze_mu_axis and len(mu_bin_edges) == 2:
        power = power[:, 0]
        N_mode = N_mode[:, 0]
        k_avg = k_avg[:, 0]
    return dict(power=power, N_mode=N_mode, binned_poles=binned_poles, N_mode_poles=N_mode_poles, k_avg=k_avg)


def calc_power(pos, Lbox, kbins=None, mubins=None, k_max=None, logk=False, paste='TSC', nmesh=128,
               compensated=True, interlaced=True, w=None, pos2=None, w2=None, poles=None, squeeze_mu_axis=True,
               nthread=MAX_THREADS, dtype=np.float32, acc64=False):
    """power_spectrum.py:1131-1319."""
    if kbins is None:
        kbins = nmesh
    if k_max is None:
        k_max = np.pi * nmesh / Lbox
    return_mubins = mubins is not None
    if mubins is None:
        mubins = 1
    meta = dict(Lbox=Lbox, logk=logk, paste=paste, nmesh=nmesh, compensated=compensated, interlaced=interlaced,
                poles=poles, nthread=nthread, N_pos=len(pos), is_weighted=w is not None, field_dtype=dtype,
                squeeze_mu_axis=squeeze_mu_axis)
    if pos2 is not None:
        meta['N_pos2'] = len(pos2)
        meta['is_weighted2'] = w2 is not None
    W = get_W_compensated(Lbox, nmesh, paste, interlaced) if compensated else None
    field_fft = get_field_fft(pos, Lbox, nmesh, paste, w, W, compensated, interlaced, nthread=nthread)
    field2_fft = None
    if pos2 is not None:
        field2_fft = get_field_fft(pos2, Lbox, nmesh, paste, w2, W, compensated, interlaced, nthread=nthread)
    poles = np.asarray(poles or [], dtype=np.int64)
    kbins, mubins = get_k_mu_edges(Lbox, k_max, kbins, mubins, logk)
    P = calc_pk_from_deltak(field_fft, Lbox, kbins, mubins, field2_fft=field2_fft, poles=poles,
                            squeeze_mu_axis=squeeze_mu_axis, nthread=nthread, acc64=acc64)
    k_binc = (kbins[1:] + kbins[:-1]) * 0.5
    mu_binc = (mubins[1:] + mubins[:-1]) * 0.5
    res = dict(k_min=kbins[:-1], k_max=kbins[1:], k_mid=k_binc, k_avg=P['k_avg'], power=P['power'],
               N_mode=P['N_mode'])
    if len(poles) > 0:
        res.update(poles=P['binned_poles'].T, N_mode_poles=P['N_mode_poles'])
    if return_mubins:
        res.update(mu_min=np.broadcast_to(mubins[:-1], res['power'].shape),
                   mu_max=np.broadcast_to(mubins[1:], res['power'].shape),
                   mu_mid=np.broadcast_to(mu_binc, res['power'].shape))
    return Table(res, meta=meta)


__all__ = ['tsc_parallel', 'partition_parallel', 'calc_power', 'pk_to_xi', 'calc_pk_from_deltak', 'project_3d_to_poles',
           'get_k_mu_edges', 'get_W_compensated', 'get_field_fft', 'get_field', 'normalize_field',
           'get_interlaced_field_fft', 'shift_field_fft', 'get_raw_power', 'bin_kmu', 'P_n',
           'tsc_scatter_serial', 'Table', 'build']
_ = warnings


# ---------------------------------------------------------------------------------------------
# Particle-format decoders: NumPy restatements, checked bit for bit against the
# unmodified reference (tests/test_oracle_vs_live_reference.py) and against the reference's own golden files
# (tests/golden/ref_ingest.npz).

def unpack_rvint(intdata, boxsize, float_dtype=np.float32, posout=None, velout=None):
    """abacusnbody/data/bitpacked.py:29-120.  int32 op uint32 promotes to int64 there; the products are
    float64 and are rounded once on the store."""
    iv = np.asarray(intdata).reshape(-1, 3)
    assert iv.dtype == np.int32
    v = iv.astype(np.int64)
    ret = []
    for spec, val in ((posout, (v >> 12) * (boxsize / 1e6)), (velout, ((v & 0xFFF) - 2048) * (6000.0 / 2048))):
        if spec is None:
            ret.append(val.astype(float_dtype))
        elif spec is False:
            ret.append(0)
        else:
            out = spec.view()
            out.shape = (-1, 3)
            out[: len(iv)] = val
            ret.append(len(iv))
    return tuple(ret)


def _pack9_fields(d):
    c = d.astype(np.int64)
    s = np.empty((len(d), 6), dtype=np.int64)
    s[:, 0] = (c[:, 1] & 0x0F) | (c[:, 0] << 4)
    s[:, 1] = ((c[:, 1] & 0xF0) << 4) | c[:, 2]
    s[:, 2] = (c[:, 4] & 0x0F) | (c[:, 3] << 4)
    s[:, 3] = ((c[:, 4] & 0xF0) << 4) | c[:, 5]
    s[:, 4] = (c[:, 7] & 0x0F) | (c[:, 6] << 4)
    s[:, 5] = ((c[:, 7] & 0xF0) << 4) | c[:, 8]
    return s - 2048


def unpack_pack9(data, boxsize, velzspace_to_kms, float_dtype=np.float32, posout=None, velout=None):
    """abacusnbody/data/pack9.py:16-123, vectorised: the header that governs a particle record is the last
    record with first byte 0xFF at or before it; particles before any header decode to NaN (the reference's
    initial header state).  Every rounding of the serial reference is kept (T = float_dtype):
    invcpd = T(1/(cpd)), csize = T*T, vscale = T(f64) * T * T, cell origin and pscale computed in float64 and
    rounded to T, particle = T(int16) * pscale + origin with two roundings in T."""
    T = np.dtype(float_dtype).type
    d = np.asanyarray(data, dtype=np.ubyte).reshape(-1, 9)
    s = _pack9_fields(d)
    is_hdr = d[:, 0] == 0xFF
    box, velz = T(boxsize), T(velzspace_to_kms)
    halfbox = np.float64(box) / 2
    hs = s[is_hdr].astype(np.float64)
    with np.errstate(divide='ignore', invalid='ignore'):
        invcpd = (1.0 / (hs[:, 1] + 2000)).astype(T)
        csize = (box * invcpd).astype(T)
        vscale = (((hs[:, 2] + 2000) * 0.0005).astype(T) * invcpd).astype(T) * velz
        cell = ((hs[:, 3:6] + 2000.5) * csize.astype(np.float64)[:, None] - halfbox).astype(T)
        pscale = (0.0005 * csize.astype(np.float64)).astype(T)
    # per-record header number (-1: none yet)
    hnum = np.cumsum(is_hdr) - 1
    part = ~is_hdr
    hn = hnum[part]
    nan = T(np.nan)
    def take(a):
        out = np.full((len(hn),) + a.shape[1:], nan, dtype=T)
        ok = hn >= 0
        out[ok] = a[hn[ok]]
        return out
    sp = s[part]
    with np.errstate(invalid='ignore'):
        pos = (sp[:, 0:3].astype(T) * take(pscale)[:, None]).astype(T) + take(cell)
        vel = (sp[:, 3:6].astype(T) * take(vscale)[:, None]).astype(T)
    npart = int(part.sum())
    ret = []
    for spec, val in ((posout, pos), (velout, vel)):
        if spec is None:
            ret.append(val.astype(T))
        elif spec is False:
            ret.append(0)
        else:
            spec[:npart] = val
            ret.append(npart)
    return tuple(ret)


def unpack_pids(packed, box=None, ppd=None, pid=False, lagr_pos=False, tagged=False, density=False, lagr_idx=False,
                float_dtype=np.float32):
    """abacusnbody/data/bitpacked.py:123-311.  uint64 * float promotes to float64 in the reference's kernel; the
    scale and the half-box are rounded to ``float_dtype`` first (:286-287)."""
    T = np.dtype(float_dtype).type
    p = np.asanyarray(packed, dtype=np.uint64).reshape(-1)
    ppd = 1 if ppd is None else int(round(ppd))
    box = 1.0 if box is None else float(box)
    inv_ppd, half = np.float64(T(box / ppd)), np.float64(T(box / 2))
    ix = (p & np.uint64(0x7FFF)).astype(np.int64)
    iy = ((p & np.uint64(0x7FFF0000)) >> np.uint64(16)).astype(np.int64)
    iz = ((p & np.uint64(0x7FFF00000000)) >> np.uint64(32)).astype(np.int64)
    out = {}
    if pid is True:
        out['pid'] = (p & np.uint64(0x7FFF7FFF7FFF)).astype(np.int64)
    if lagr_pos is True:
        out['lagr_pos'] = (np.stack([ix, iy, iz], axis=1).astype(np.float64) * inv_ppd - half).astype(T)
    if lagr_idx is True:
        out['lagr_idx'] = np.stack([ix, iy, iz], axis=1).astype(np.int16)
    if tagged is True:
        out['tagged'] = ((p >> np.uint64(48)) & np.uint64(1)).astype(np.uint8)
    if density is True:
        r = ((p & np.uint64(0x07FE000000000000)) >> np.uint64(49)).astype(np.int64)
        out['density'] = (r * r).astype(T)
    return out
