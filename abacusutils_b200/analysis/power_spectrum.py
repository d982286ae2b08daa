r"""
Density-field power spectra on B200 -- drop-in for ``abacusnbody.analysis.power_spectrum``.

Same public names, argument order, defaults and return types as the reference
(/root/reference/abacusnbody/analysis/power_spectrum.py); the heavy lifting (TSC painting, FFT,
interlacing, window compensation, (k,mu) binning, Legendre multipoles) runs in libabk.so
(abacusutils_b200/csrc) through ctypes.  Host-side pieces that are O(nmesh) or O(Nk) -- bin edges,
the 1-D window table, the final divisions and the result table -- stay in NumPy, written so the
float32/float64 tables are identical to the reference's.

``calc_power`` never materialises delta(k) or P(k): the two painted grids are transformed in place
and a single fused kernel applies 1/n^3, the interlacing phase and the window, forms |delta|^2 (or
the cross spectrum) and bins it.  ``get_field_fft`` / ``calc_pk_from_deltak`` expose the same stages
separately for callers that want delta(k) (the ZCV modules of the reference).

Inputs may be NumPy arrays, torch CPU tensors (pinned memory is copied asynchronously) or torch CUDA
tensors; array-valued results follow the kind of the input (NumPy in -> NumPy out).
"""

from __future__ import annotations

import ctypes as C
import math
import os
import warnings

import numpy as np

from .._lib import (ABK_MAX_POLES, ABK_MAX_SEGMENTS, ABK_POLE_NCOEF, SEGMENT_MAX, AbkError, BinRequest, Engine, KMesh,
                    check, is_torch_tensor, ptr)
from ..data.packed import PackedParticles
from .tsc import padded_ldz, tsc_parallel

__all__ = ['calc_power', 'calc_pk_from_deltak', 'pk_to_xi', 'project_3d_to_poles', 'get_k_mu_edges']

MAX_THREADS = 1  # the reference's thread knob (numba.config.NUMBA_NUM_THREADS); ignored on the GPU

try:  # the reference returns an astropy Table (power_spectrum.py:1318); use it when available
    from astropy.table import Table as _AstropyTable
except Exception:  # pragma: no cover - astropy is not installed in the build image
    _AstropyTable = None


class Table(dict):
    """Fallback result container with the part of astropy.table.Table the callers use
    (column access by name, ``.meta``, ``.colnames``)."""

    def __init__(self, d, meta=None):
        super().__init__(d)
        self.meta = meta or {}

    @property
    def colnames(self):
        return list(self.keys())


def _make_table(res, meta):
    if _AstropyTable is not None:
        return _AstropyTable(res, meta=meta)
    return Table(res, meta=meta)


# ---------------------------------------------------------------------------------------------
# host-side O(n) pieces
def get_k_mu_edges(Lbox, k_max, kbins, mubins, logk):
    r"""Bin edges of k and mu (reference: power_spectrum.py:663-704).

    ``kbins`` / ``mubins`` may be ints (number of bins) or arrays (returned unchanged).  Linear k bins
    run from 0 to ``k_max``; logarithmic ones from :math:`(1-10^{-4})\,2\pi/L` to ``k_max``.
    """
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
    """1-D mass-assignment window over ``fftfreq`` order (reference: power_spectrum.py:1081-1128).

    Interlaced: ``sinc(k/2kN)**p`` with p=3 (TSC) / 2 (CIC); otherwise the first-order aliasing
    correction ``sqrt(1 - s + 2/15 s^2)`` (TSC) / ``sqrt(1 - 2/3 s)`` (CIC), ``s = sin^2(pi k / 2 kN)``.
    The wavenumbers are cast to float32 first, as in the reference.
    """
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
        W = (1 - s + 2.0 / 15 * s**2) ** 0.5 if paste == 'TSC' else (1 - 2.0 / 3 * s) ** 0.5
    return W


# ---------------------------------------------------------------------------------------------
# small scalar helpers of the reference module (host side; the kernels evaluate the same polynomials on the device)
def factorial(n):
    """n! for 0 <= n <= 20 (power_spectrum.py:58-77)."""
    if n > 20 or n < 0:
        raise ValueError
    return np.int64(math.factorial(int(n)))


def factorial_slow(x):
    """n! by repeated multiplication (power_spectrum.py:80-98)."""
    return math.factorial(int(x)) if x >= 0 else 1


def n_choose_k(n, k):
    """Binomial coefficient (power_spectrum.py:101-119)."""
    return factorial(n) // (factorial(k) * factorial(n - k))


def P_n(x, n, dtype=np.float32):
    """Legendre polynomial of order ``n`` for the SQUARED argument ``x = mu^2`` (power_spectrum.py:122-147), evaluated
    like the reference: an alternating power sum in ``dtype``; valid up to n = 10."""
    dtype = np.dtype(dtype).type
    x = dtype(x)
    total = dtype(0.0)
    for k in range(n // 2 + 1):
        factor = dtype(int(n_choose_k(n, k)) * int(n_choose_k(2 * n - 2 * k, n)))
        term = factor * x ** dtype(0.5 * (n - 2 * k))
        total = dtype(total + term) if k % 2 == 0 else dtype(total - term)
    return dtype(total * dtype(0.5**n))


def linear_interp(xd, x, y):
    """Linear interpolation on an equidistant grid, clamped to ``y[0]`` / ``y[-1]`` outside (power_spectrum.py:509-536)."""
    if xd <= x[0]:
        return y[0]
    elif xd >= x[-1]:
        return y[-1]
    dx = x[1] - x[0]
    f = (xd - x[0]) / dx
    fl = np.int64(f)
    return y[fl] + (f - fl) * (y[fl + 1] - y[fl])


def legendre_coefficients(poles):
    """(2l+1) P_l(mu) as polynomial coefficients in mu (degree <= 10), float32[Np][11].

    Closed form used by the reference's ``P_n`` (power_spectrum.py:121-147):
    P_l(mu) = 2^-l sum_k (-1)^k C(l,k) C(2l-2k,l) mu^(l-2k).
    """
    out = np.zeros((len(poles), ABK_POLE_NCOEF), dtype=np.float64)
    for ip, ell in enumerate(poles):
        ell = int(ell)
        for k in range(ell // 2 + 1):
            out[ip, ell - 2 * k] += (-1) ** k * math.comb(ell, k) * math.comb(2 * ell - 2 * k, ell) * 0.5**ell
        out[ip] *= 2 * ell + 1
    return out.astype(np.float32)


# ---------------------------------------------------------------------------------------------
# device plumbing
def _as_source(a, dtype_np):
    """Normalise an input array: returns (kind, obj) with kind in {'cuda', 'host'}; host objects are
    torch CPU tensors sharing memory with the caller's array (no copy for float32 C-contiguous)."""
    import torch

    if is_torch_tensor(a):
        if a.is_cuda:
            return 'cuda', a
        return 'host', a.contiguous()
    arr = np.ascontiguousarray(a)
    if arr.dtype != dtype_np:
        arr = arr.astype(dtype_np)
    if not arr.flags.writeable:
        arr = arr.copy()
    return 'host', torch.from_numpy(arr)


class _Painter:
    """Stages a particle set on the device, buckets it for one or two offsets, deposits, normalises and
    transforms in place.  Host inputs are streamed in chunks on a copy stream so the bucketing of
    chunk i overlaps the host->device copy of chunk i+1; every chunk becomes one bucket segment."""

    def __init__(self, eng, nmesh, Lbox, paste='TSC'):
        self.eng = eng
        self.n = int(nmesh)
        self.L = float(Lbox)
        self.ldz = padded_ldz(self.n)
        # 'CIC' is the reference's cic_serial (analysis/cic.py) -- same 27-cell update, other weights; it
        # applies no periodic wrap to the positions (power_spectrum.py:846-853)
        self.paste = _check_paste(paste)

    def chunk_plan(self, N, host=True):
        """Chunks of the particle set; every chunk becomes one bucket segment.  Host inputs want many chunks (they
        are the unit of copy/compute overlap).  Device-resident inputs only need them below SEGMENT_MAX particles."""
        max_seg = ABK_MAX_SEGMENTS - 2
        if not host and os.environ.get('ABK_DEVICE_SEGMENTS'):
            # testing knob.  Default for device-resident input: the host plan -- measured on B200 at config 3, 14 segments
            # beat one (bucket scatter 30.2 vs 32.4 ms: a segment's open write frontier is smaller), step 86.8 vs 87.7 ms
            max_seg = max(1, min(max_seg, int(os.environ['ABK_DEVICE_SEGMENTS'])))
            chunk = -(-N // max_seg)
        else:
            # ABK_CHUNK_MIN: testing knob, lets small inputs exercise the multi-segment / early-deposit machinery
            chunk = max(int(os.environ.get('ABK_CHUNK_MIN', 1 << 25)), -(-N // max_seg))
        chunk = min(max(chunk, 1), SEGMENT_MAX)
        return [(a, min(a + chunk, N)) for a in range(0, N, chunk)]

    def paint(self, pos, w, offsets, wrap=True, tag='', fft_weight=None):
        """Returns one padded device grid (n, n, ldz) float32 per offset.  With ``fft_weight=None`` the grids
        hold the raw deposit; otherwise they are normalised by ``fft_weight`` and transformed in place
        (the transform of the first grid then overlaps the deposit of the second on an auxiliary stream).

        Host inputs are streamed in chunks: bucketing of chunk i overlaps the copy of chunk i+1, and the
        tile deposit of the first ~70% of the chunks runs on the auxiliary stream while the remaining
        chunks are still arriving over PCIe, so only the last chunks' deposit is left for the tail."""
        import torch

        eng, n, ldz = self.eng, self.n, self.ldz
        lib = eng.lib
        eng.set_scheme(self.paste)
        if self.paste == 'CIC':
            wrap = False
        packed = pos if isinstance(pos, PackedParticles) else None
        if packed is not None:
            if w is not None:
                raise ValueError('packed particle records carry no weights')
            kind, psrc, N = 'host', None, len(packed)      # N: records, an upper bound of the particle count
        else:
            kind, psrc = _as_source(pos, np.float32)
            N = int(psrc.shape[0])
        wsrc = None
        if w is not None:
            wkind, wsrc = _as_source(w, np.float32)
            if wkind != kind:
                wsrc = wsrc.to(psrc.device)
        compute = eng.bind_stream()
        if N == 0:
            if fft_weight is not None:
                raise ValueError('cannot normalise an empty particle set')
            return [eng.zeros((n, n, ldz), torch.float32) for _ in offsets]
        # normalize_field (rho * n^3/N - 1, power_spectrum.py:860-901) folded into the deposit: the grid starts at -1
        # and every weight is scaled by n^3/N as the bucket records are written -- one read+write pass less per grid
        fused_norm = fft_weight is not None and os.environ.get('ABK_FUSED_NORMALIZE', '1') != '0'
        if packed is not None:
            if packed.n_particles is None:
                fused_norm = False      # pack9: the particle count is only known once the last chunk has been decoded
            elif fft_weight is not None:
                # len(packed) counts RECORDS (pack9: cell headers included); the field is normalised by the
                # number of particles, known from an earlier decode of the same object
                fft_weight = packed.n_particles
                if fft_weight == 0:
                    raise ValueError('cannot normalise an empty particle set')
        # the grids are initialised on the auxiliary stream: 4.3 GB each at nmesh 1024, 0.7 ms that would otherwise sit
        # in front of the (latency-bound) bucketing; every deposit waits for `self._grids_ready`
        grids = [eng.empty((n, n, ldz), torch.float32) for _ in offsets]
        aux = eng.aux_stream()
        aux.wait_stream(compute)      # the allocator handed the blocks out in the order of the compute stream
        with torch.cuda.stream(aux):
            for g in grids:
                g.fill_(-1.0 if fused_norm else 0.0)
            self._grids_ready = torch.cuda.Event()
            self._grids_ready.record(aux)
        if fused_norm:
            eng.set_weight_scale(float(np.float32(float(n) ** 3 / float(fft_weight))))
        try:
            return self._paint(psrc, wsrc, kind, N, offsets, wrap, tag, fft_weight, fused_norm, grids, compute, packed)
        finally:
            if fused_norm:
                eng.set_weight_scale(1.0)

    def _paint(self, psrc, wsrc, kind, N, offsets, wrap, tag, fft_weight, fused_norm, grids, compute, packed=None):
        import torch

        eng, n, ldz = self.eng, self.n, self.ldz
        lib = eng.lib
        if psrc is not None and psrc.dtype != torch.float32:
            psrc = psrc.to(torch.float32)
        if wsrc is not None and wsrc.dtype != torch.float32:
            wsrc = wsrc.to(torch.float32)

        ntiles = C.c_int64()
        check(lib.abk_tsc_num_tiles(n, n, n, C.byref(ntiles)))
        ntiles = ntiles.value
        if packed is not None:
            chunks = packed.chunk_plan(ABK_MAX_SEGMENTS - 2, int(os.environ.get('ABK_CHUNK_MIN', 1 << 25)))
        else:
            chunks = self.chunk_plan(N, host=(kind == 'host'))
        nseg = len(chunks)
        counts = [b - a for a, b in chunks]      # particles per segment (packed input: filled in as chunks are decoded)
        bucket_fn = lib.abk_tsc_bucket
        nb = C.c_size_t()
        check(lib.abk_tsc_bucket_scratch_bytes(chunks[0][1] - chunks[0][0], n, n, n, C.byref(nb)))
        scan_buf = eng.scratch('bucket_scan', nb.value + 256)
        scan_ptr = C.c_void_p((scan_buf.data_ptr() + 255) & ~255)
        starts_stride = (ntiles + 1 + 63) // 64 * 64
        host = kind == 'host'
        # Device-resident input: bucket ONCE (tile of the cell at the first offset); the deposit of the
        # half-cell-shifted grid reuses the records with the widened tile domain (abk_tsc_deposit_tiles,
        # bucket_offset) -- one histogram+scatter less.  Host input: the run is bound by the PCIe copy and
        # bucketing hides behind it, so bucket per offset and keep the (faster) plain tile kernel in the
        # un-overlapped tail.
        nbuck = len(offsets) if host else 1
        records = [eng.scratch(f'records{tag}{o}', N * 16) for o in range(nbuck)]
        starts = [eng.scratch(f'starts{tag}{o}', nseg * starts_stride * 4) for o in range(nbuck)]
        aux = eng.aux_stream()

        def deposit(lo, hi, o, stream):
            """Enqueue the tile deposit of segments [lo, hi) for offset o on `stream`."""
            ob = o if nbuck > 1 else 0
            m = hi - lo
            recs = (C.c_void_p * m)(*[records[ob].data_ptr() + chunks[s][0] * 16 for s in range(lo, hi)])
            sts = (C.c_void_p * m)(*[starts[ob].data_ptr() + s * starts_stride * 4 for s in range(lo, hi)])
            cnts = (C.c_int64 * m)(*[counts[s] for s in range(lo, hi)])
            stream.wait_event(self._grids_ready)
            with torch.cuda.stream(stream):
                eng.bind_stream()
                check(lib.abk_tsc_deposit_tiles(eng.ctx, m, recs, sts, cnts, ptr(grids[o]), n, n, n, ldz, self.L,
                                                float(offsets[o]), float(offsets[ob]), 0, 0, n))
            eng.bind_stream()

        # Early deposits (host input): segments [0, split) are deposited on the auxiliary stream while the remaining
        # chunks are still arriving; `cuts` are the ends of the early groups.  Every group costs one full per-cell pass
        # over the mesh (7.4 ms at nmesh 1024 however few particles it carries).  One group with a 5-of-14-segment tail is
        # the default because it is the setting without a cliff: two groups save 3-4 ms on a uniform catalogue with a
        # 3-segment tail (250 vs 254 ms end to end at config 3) but cost 30 ms as soon as the last early group runs into
        # the tail -- measured for a clustered catalogue with a 3-segment tail (283 ms; 4 segments: 250.6; one group:
        # 252.0) and for the uniform one with a 4-segment tail (282 ms): the early deposits run on the high-priority
        # stream, the bucket kernels behind them starve, the staging ring fills and the PCIe copy stalls.
        # ABK_EARLY_GROUPS / ABK_TAIL_SEGMENTS are experiment knobs.
        split, cuts = 0, []
        # (Device-resident input does not get early groups: measured at config 3, every extra deposit launch costs its
        # per-tile walk again -- 7.4 ms per grid however few particles it carries -- and the bucket kernels do not
        # speed up next to it: 101.6 / 119.0 / 136.8 ms per step with 1 / 2 / 3 early groups against 82.0.)
        if host and nseg >= 6:
            groups = max(1, int(os.environ.get('ABK_EARLY_GROUPS', '1')))
            tail = max(3, -(-3 * nseg // 10)) if groups == 1 else max(2, -(-(2.8 if groups == 2 else 1.5) * nseg // 10))
            if os.environ.get('ABK_TAIL_SEGMENTS'):      # experiment knob: segments left for the un-overlapped tail
                tail = max(1, min(nseg - 1, int(os.environ['ABK_TAIL_SEGMENTS'])))
            split = nseg - int(tail)
            cuts = sorted({max(1, round(split * (j + 1) / groups)) for j in range(groups)})
        early_done = None

        if host:
            csize = max(b - a for a, b in chunks)
            copy_stream = torch.cuda.Stream(device=eng.device)
            # a ring of staging buffers: deep enough that PCIe keeps streaming while the bucket kernels share
            # the SMs with the early tile deposits
            NSTAGE = 4
            stage_p = [eng.scratch(f'stage_p{i}', csize * 12) for i in range(NSTAGE)]
            stage_raw = [eng.scratch(f'stage_raw{i}', csize * packed.rec_bytes) for i in range(NSTAGE)] if packed is not None else None
            stage_w = [eng.scratch(f'stage_w{i}', csize * 4) for i in range(NSTAGE)] if wsrc is not None else None
            ready = [torch.cuda.Event() for _ in range(NSTAGE)]
            done = [torch.cuda.Event() for _ in range(NSTAGE)]
            for ev in done:
                ev.record(compute)

        staged = {}

        def issue_copy(si):
            """Enqueue the host->device copy of chunk si on the copy stream (waits, in stream order, for its staging slot)."""
            a_, b_ = chunks[si]
            m_, sl = b_ - a_, si % NSTAGE
            copy_stream.wait_event(done[sl])
            with torch.cuda.stream(copy_stream):
                wd_ = None
                if packed is not None:
                    pd_ = stage_raw[sl][: m_ * packed.rec_bytes]
                    pd_.copy_(packed.raw(a_, b_), non_blocking=True)
                else:
                    pd_ = stage_p[sl][: m_ * 12].view(torch.float32).view(m_, 3)
                    pd_.copy_(psrc[a_:b_], non_blocking=True)
                    if wsrc is not None:
                        wd_ = stage_w[sl][: m_ * 4].view(torch.float32)
                        wd_.copy_(wsrc[a_:b_], non_blocking=True)
                ready[sl].record(copy_stream)
            staged[si] = (pd_, wd_)

        serial = os.environ.get('ABK_NO_OVERLAP') == '1'      # measurement knob: every kernel alone on one stream
        nbs = min(nseg, max(1, int(os.environ.get('ABK_BUCKET_STREAMS', '2')))) if not (host or serial) else 1
        bstreams, scan_ptrs = [], []
        if nbs > 1:
            if len(getattr(eng, '_bstreams', [])) < nbs - 1:
                eng._bstreams = [torch.cuda.Stream(device=eng.device) for _ in range(nbs - 1)]
            bstreams = eng._bstreams[: nbs - 1]
            for i, bs in enumerate(bstreams):
                sb = eng.scratch(f'bucket_scan_side{i}', nb.value + 256)
                scan_ptrs.append(C.c_void_p((sb.data_ptr() + 255) & ~255))
                bs.wait_stream(compute)
        issued = 0
        for s, (a, b) in enumerate(chunks):
            m = b - a
            if host:
                slot = s % NSTAGE
                # packed input: the header count of chunk s blocks the host, so chunk s+1 is put on the wire first
                ahead = 1 if (packed is not None and NSTAGE > 1) else 0
                while issued <= min(s + ahead, nseg - 1):
                    issue_copy(issued)
                    issued += 1
                pd, wd = staged.pop(s)
                compute.wait_event(ready[slot])
                if packed is not None:      # decode on the device; m becomes the number of PARTICLES of this chunk
                    raw_d = pd
                    pd = stage_p[slot][: m * 12].view(torch.float32).view(m, 3)
                    m = packed.decode(eng, raw_d, m, pd)
                    counts[s] = m
                    pd = pd[:m]
            else:
                pd = psrc[a:b]
                wd = None if wsrc is None else wsrc[a:b]
                if not pd.is_contiguous():
                    pd = pd.contiguous()
            # Device-resident input: the segments are independent, so odd segments are bucketed on a second stream -- the
            # histogram pass of one segment (bound by the L2 reduction rate) then shares the GPU with the scatter pass of
            # another (bound by store / translation latency at 12 % issue utilisation) instead of running after it.
            side = (s % nbs) if nbs > 1 else 0
            if side:
                with torch.cuda.stream(bstreams[side - 1]):
                    eng.bind_stream()
                    for o in range(nbuck):
                        check(bucket_fn(eng.ctx, ptr(pd), ptr(wd), m, n, n, n, self.L, float(offsets[o]), int(bool(wrap)),
                                        C.c_void_p(records[o].data_ptr() + a * 16), C.c_void_p(starts[o].data_ptr() + s * starts_stride * 4),
                                        scan_ptrs[side - 1], nb.value))
                eng.bind_stream()
            else:
                for o in range(nbuck):
                    rec_ptr = records[o].data_ptr() + a * 16
                    st_ptr = starts[o].data_ptr() + s * starts_stride * 4
                    check(bucket_fn(eng.ctx, ptr(pd), ptr(wd), m, n, n, n, self.L, float(offsets[o]), int(bool(wrap)),
                                    C.c_void_p(rec_ptr), C.c_void_p(st_ptr), scan_ptr, nb.value))
            if host:
                done[slot].record(compute)
            if (s + 1) in cuts:
                lo = ([0] + cuts)[cuts.index(s + 1)]
                ev = torch.cuda.Event()
                ev.record(compute)
                aux.wait_event(ev)
                for o in range(len(offsets)):
                    deposit(lo, s + 1, o, aux)
                early_done = torch.cuda.Event()
                early_done.record(aux)

        for bs in bstreams:
            compute.wait_stream(bs)
        if packed is not None:
            packed.n_particles = int(sum(counts))
            if fft_weight is not None:
                fft_weight = packed.n_particles      # normalise by the particles actually decoded
                if fft_weight == 0:
                    raise ValueError('cannot normalise an empty particle set')
        # ---- remaining deposits; FFT of grid o overlaps the deposit of grid o+1 ----------------------------
        fft_prev = None
        for o in range(len(offsets)):
            deposit(split, nseg, o, compute)
            if fft_weight is None:
                continue
            last = o == len(offsets) - 1
            if not last and not serial:
                ev = torch.cuda.Event()
                ev.record(compute)
                aux.wait_event(ev)  # aux already holds the early deposits of this grid, in order
                with torch.cuda.stream(aux):
                    self.normalize_fft(grids[o], None if fused_norm else fft_weight)
                    fft_prev = torch.cuda.Event()
                    fft_prev.record(aux)
                eng.bind_stream()
            else:
                if early_done is not None:
                    compute.wait_event(early_done)
                if fft_prev is not None:
                    compute.wait_event(fft_prev)  # one FFT work area: transforms run one after the other
                self.normalize_fft(grids[o], None if fused_norm else fft_weight)
        if early_done is not None:
            compute.wait_event(early_done)
        if fft_prev is not None:
            compute.wait_event(fft_prev)
        return grids

    def normalize_fft(self, grid, tot_weight):
        """In-place FFT of a painted grid; ``tot_weight=None`` means the grid already holds the normalised field."""
        eng, n, ldz = self.eng, self.n, self.ldz
        eng.bind_stream()
        if tot_weight is not None:
            check(eng.lib.abk_normalize_field(eng.ctx, ptr(grid), n, n, n, ldz, float(n) ** 3, float(tot_weight)))
        eng.rfft3_inplace(grid, n, n, n)


def _kmesh(n, row_len=None):
    nzc = n // 2 + 1
    row_len = nzc if row_len is None else int(row_len)
    return KMesh(n=n, nzc=nzc, i0=0, i1=n, j0=0, j1=n, stride_i=n * row_len, stride_j=row_len)


def _complex_view(grid_padded, n):
    """(n, n, ldz) float32 in-place FFT buffer -> (n, n, n//2+1) complex64 view (same memory)."""
    import torch

    return torch.view_as_complex(grid_padded.view(n, n, n // 2 + 1, 2))


def _check_paste(paste):
    paste_u = str(paste).upper()
    if paste_u not in ('TSC', 'CIC'):
        raise ValueError(f'Unknown pasting method: {paste}')
    return paste_u


# ---------------------------------------------------------------------------------------------
# real-space field
def normalize_field(field, tot_weight=None, inplace=False, nthread=MAX_THREADS):
    """``field / field.mean() - 1`` as ``field * f32(size/tot_weight) - 1`` (power_spectrum.py:860-901)."""
    import torch

    on_device = is_torch_tensor(field) and field.is_cuda
    eng = Engine.get(field.device if on_device else None)
    eng.bind_stream()
    fd = field if on_device else eng.to_device(field, torch.float32)
    if tot_weight is None:
        tot_weight = float(fd.sum(dtype=torch.float32).item())
    if on_device and not inplace:
        fd = fd.clone()
    if not fd.is_contiguous() or fd.dtype != torch.float32:
        raise AbkError('normalize_field: device field must be contiguous float32')
    nz = fd.shape[-1]
    rows = fd.numel() // nz
    check(eng.lib.abk_normalize_field(eng.ctx, ptr(fd), 1, rows, nz, nz, float(fd.numel()), float(tot_weight)))
    if on_device:
        return fd
    out = fd.cpu().numpy()
    if inplace:
        np.copyto(field, out.astype(field.dtype, copy=False))
        return field
    return out


def get_field(pos, Lbox, nmesh, paste, w=None, d=0.0, nthread=MAX_THREADS, dtype=np.float32):
    """Paint + normalise: returns the overdensity field (nmesh,)*3 float32 (power_spectrum.py:808-857).
    Normalisation uses ``len(pos)``, not ``sum(w)``, like the reference (:856)."""
    if w is not None:
        assert pos.shape[0] == len(w)
    paste = _check_paste(paste)
    on_device = is_torch_tensor(pos) and pos.is_cuda
    eng = Engine.get(pos.device if on_device else None)
    P = _Painter(eng, nmesh, Lbox, paste)
    (grid,) = P.paint(pos, w, [d])
    n = int(nmesh)
    check(eng.lib.abk_normalize_field(eng.ctx, ptr(grid), n, n, n, P.ldz, float(n) ** 3, float(len(pos))))
    field = grid[:, :, :n].contiguous()
    return field if on_device else field.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# Fourier-space fields
def shift_field_fft(field_fft, field_shift_fft, n1d, L, d, dtype=np.float32):
    """In place ``f <- (f + fs * exp(i*0.5*d*(kx+ky+kz))) * 0.5/n^3`` for any shift ``d`` (power_spectrum.py:904-948;
    the reference's own callers pass ``d = L/n1d``, the half-cell interlacing shift)."""
    import torch

    on_device = is_torch_tensor(field_fft) and field_fft.is_cuda
    eng = Engine.get(field_fft.device if on_device else None)
    eng.bind_stream()
    n = int(n1d)
    f = field_fft if on_device else eng.to_device(field_fft, torch.complex64)
    fs = eng.to_device(field_shift_fft, torch.complex64)
    mesh = _kmesh(n)
    check(eng.lib.abk_shift_field_fft(eng.ctx, C.byref(mesh), ptr(f), ptr(fs), float(d) / float(L), np.float32(0.5 / n**3)))
    if not on_device:
        np.copyto(field_fft, f.cpu().numpy())


def _field_fft_device(eng, pos, Lbox, nmesh, w, interlaced, tag='', paste='TSC'):
    """Paint, normalise and FFT; returns the in-place padded grids [A] or [A, B(shifted)] (unscaled)."""
    n = int(nmesh)
    P = _Painter(eng, n, Lbox, paste)
    offsets = [0.0, 0.5 * (float(Lbox) / n)] if interlaced else [0.0]
    return P.paint(pos, w, offsets, tag=tag, fft_weight=len(pos))


def _is_f64(dtype):
    return np.dtype(dtype) == np.float64


def _field_fft_f64(eng, pos, Lbox, nmesh, w, paste='TSC'):
    """The non-interlaced branch of get_field_fft in float64 (power_spectrum.py:1053-1060 with dtype=float64): paint onto
    a float64 grid (arithmetic in the dtype of ``pos``, tsc.py:400), normalise by len(pos), D2Z transform.  Returns the
    UNSCALED spectrum as a device tensor complex128 (n, n, n//2+1) (a view of the padded in-place buffer)."""
    import torch

    from .tsc import _float_kind

    n = int(nmesh)
    ldz = padded_ldz(n)
    N = len(pos)
    if N == 0:
        raise ValueError('cannot normalise an empty particle set')
    pos_dt = _float_kind(pos)
    pos_d = eng.to_device(pos, torch.float64 if pos_dt == 'f8' else torch.float32)
    w_dt = None if w is None else _float_kind(w)
    w_d = None if w is None else eng.to_device(w, torch.float64 if w_dt == 'f8' else torch.float32)
    grid = eng.zeros((n, n, ldz), torch.float64)
    eng.bind_stream()
    eng.set_scheme(paste)
    # TSC wraps the positions once (tsc.py:171-173); cic_serial does not (power_spectrum.py:846-853)
    check(eng.lib.abk_tsc_deposit_typed(eng.ctx, ptr(pos_d), int(pos_dt == 'f8'), ptr(w_d), int(w_dt == 'f8'), N, ptr(grid), 1,
                                        n, n, n, ldz, float(Lbox), 0.0, int(paste == 'TSC')))
    check(eng.lib.abk_normalize_field_f64(eng.ctx, ptr(grid), n, n, n, ldz, float(n) ** 3, float(N)))
    check(eng.lib.abk_rfft3_f64(eng.ctx, ptr(grid), n, n, n))
    return torch.view_as_complex(grid.view(n, n, n // 2 + 1, 2))


def get_interlaced_field_fft(pos, Lbox, nmesh, paste, w, nthread=MAX_THREADS, verbose=False):
    """Interlaced delta(k), complex64 (n, n, n//2+1) (power_spectrum.py:951-998)."""
    return get_field_fft(pos, Lbox, nmesh, paste, w, None, False, True, nthread=nthread, verbose=verbose)


def get_field_fft(pos, Lbox, nmesh, paste, w, W, compensated, interlaced, nthread=MAX_THREADS, verbose=False,
                  dtype=np.float32):
    """delta(k) of a particle set, complex64 (n, n, n//2+1), NumPy rfftn conventions
    (power_spectrum.py:1001-1070): paint (+ half-cell-shifted paint), normalise by len(pos), FFT,
    scale by 1/n^3 (0.5/n^3 and the interlacing phase when ``interlaced``), divide by the window."""
    import torch

    paste = _check_paste(paste)
    if w is not None:
        assert pos.shape[0] == len(w)
    if compensated:
        assert W is not None
    on_device = is_torch_tensor(pos) and pos.is_cuda
    eng = Engine.get(pos.device if on_device else None)
    n = int(nmesh)
    if _is_f64(dtype) and not interlaced:
        # the reference honours dtype on the non-interlaced branch only (power_spectrum.py:1053-1069; the interlaced
        # paints are always float32, :979,985): float64 field, float64 transform, complex128 result
        f = _field_fft_f64(eng, pos, Lbox, n, w, paste)
        f *= 1.0 / float(n) ** 3
        if compensated:
            Wt = eng.to_device(np.asarray(W, dtype=np.float32), torch.float32)
            f /= ((Wt[:, None, None] * Wt[None, :, None]) * Wt[None, None, : n // 2 + 1])
        return f if on_device else f.cpu().numpy()
    grids = _field_fft_device(eng, pos, Lbox, n, w, interlaced, paste=paste)
    W_d = eng.to_device(np.asarray(W, dtype=np.float32), torch.float32) if compensated else None
    mesh = _kmesh(n)
    scale = np.float32(0.5 / n**3) if interlaced else np.float32(1 / n**3)
    eng.bind_stream()
    check(eng.lib.abk_field_fft_finish(eng.ctx, C.byref(mesh), ptr(grids[0]), ptr(grids[1]) if interlaced else None,
                                       ptr(W_d), scale))
    out = _complex_view(grids[0], n)
    if on_device:
        return out
    return out.cpu().numpy()


def get_raw_power(field_fft, field2_fft=None):
    """``|f|^2`` or ``Re(conj(f) f2)`` as a float32 array (power_spectrum.py:707-727)."""
    import torch

    on_device = is_torch_tensor(field_fft) and field_fft.is_cuda
    eng = Engine.get(field_fft.device if on_device else None)
    eng.bind_stream()
    f1 = eng.to_device(field_fft, torch.complex64)
    f2 = None if field2_fft is None else eng.to_device(field2_fft, torch.complex64)
    out = eng.empty(tuple(f1.shape), torch.float32)
    check(eng.lib.abk_raw_power(eng.ctx, ptr(f1), ptr(f2), ptr(out), f1.numel()))
    return out if on_device else out.cpu().numpy()


# ---------------------------------------------------------------------------------------------
# binning
def _bin_device(eng, n, L, kedges, muedges, poles, fourier, *, f1=None, f1s=None, f2=None, f2s=None, real_in=None,
                row_len=None, W_d=None, scale=1.0, finish=False, W_sym=True):
    """Run abk_power_bin and return the reference's five arrays as float64/int64 NumPy
    (means, not yet cast): (weighted_counts, counts, weighted_counts_poles, counts_poles, weighted_counts_k)."""
    import torch

    kedges = np.asarray(kedges, dtype=np.float64)
    muedges = np.asarray(muedges, dtype=np.float64)
    poles = np.asarray(poles, dtype=np.int64).reshape(-1)
    Nk, Nmu, Np = len(kedges) - 1, len(muedges) - 1, len(poles)
    if Nk < 1 or Nmu < 1:
        raise ValueError('need at least one k bin and one mu bin')
    if Np > ABK_MAX_POLES:
        raise ValueError(f'at most {ABK_MAX_POLES} multipoles per call')
    if Np and (poles.max() > 10 or poles.min() < 0):
        raise ValueError('multipoles must be in [0, 10]')
    dk = 2.0 * np.pi / L if fourier else L / n
    # power_spectrum.py:217-218: square in float64, THEN cast to float32
    kedges2 = ((kedges / dk) ** 2).astype(np.float32)
    muedges2 = (muedges**2).astype(np.float32)

    Nb = Nk * Nmu
    tables = np.concatenate([kedges2, muedges2, legendre_coefficients(poles).reshape(-1)]).astype(np.float32)
    tables_d = eng.to_device(tables, torch.float32)
    sums = eng.zeros((3 * Nb + Np * Nk,), torch.float64)

    req = BinRequest()
    req.mesh = _kmesh(n, row_len)
    for name, t in (('f1', f1), ('f1s', f1s), ('f2', f2), ('f2s', f2s), ('real_in', real_in), ('W', W_d)):
        setattr(req, name, None if t is None else t.data_ptr())
    req.scale = float(scale)
    req.finish = int(bool(finish))
    base = tables_d.data_ptr()
    req.kedges2 = base
    req.muedges2 = base + 4 * (Nk + 1)
    req.pole_coef = base + 4 * (Nk + 1 + Nmu + 1)
    for ip in range(Np):
        req.pole_ell[ip] = int(poles[ip])
    req.Nk, req.Nmu, req.Np = Nk, Nmu, Np
    req.w_symmetric = int(W_sym)
    sb = sums.data_ptr()
    req.counts = sb
    req.sum_p = sb + 8 * Nb
    req.sum_k = sb + 16 * Nb
    req.sum_poles = sb + 24 * Nb
    nb = C.c_size_t()
    check(eng.lib.abk_power_bin_scratch_bytes(Nk, Nmu, Np, C.byref(nb)))
    if nb.value <= 64 << 20:
        rep = eng.scratch('bin_replicas', nb.value)
        req.scratch = rep.data_ptr()
        req.scratch_bytes = rep.numel()
    eng.bind_stream()
    check(eng.lib.abk_power_bin(eng.ctx, C.byref(req)))
    return _finalize_bins(sums, Nk, Nmu, poles, dk)


def _finalize_bins(sums, Nk, Nmu, poles, dk):
    """Host tail of bin_kmu (power_spectrum.py:276-293): pole l=0 from the wedge sums, divisions by
    the mode counts where non-zero (empty bins stay 0)."""
    Np, Nb = len(poles), Nk * Nmu
    h = sums.cpu().numpy()
    counts = h[:Nb].view(np.int64).reshape(Nk, Nmu).copy()
    sw = h[Nb:2 * Nb].reshape(Nk, Nmu).copy()
    sk = h[2 * Nb:3 * Nb].reshape(Nk, Nmu) * dk
    sp = h[3 * Nb:].reshape(Np, Nk).copy()
    counts_poles = counts.sum(axis=1)
    for ip, pole in enumerate(poles):
        if pole == 0:
            sp[ip] = sw.sum(axis=1)
    nz = counts != 0
    sw[nz] /= counts[nz]
    sk[nz] /= counts[nz]
    nzp = counts_poles != 0
    if Np:
        sp[:, nzp] /= counts_poles[nzp]
    return sw, counts, sp, counts_poles, sk


def bin_kmu(n1d, L, kedges, muedges, weights, poles=np.empty(0, 'i8'), dtype=np.float32, fourier=True,
            nthread=MAX_THREADS):
    """Mean and mode count of a 3-D mesh in (k, mu) bins, plus Legendre multipoles
    (power_spectrum.py:150-300).  ``weights`` is real, shape (n1d, n1d, n1d//2+1) (rfft layout) or
    (n1d, n1d, n1d) (real-space mesh, ``fourier=False``).  Bins are right-closed, Nyquist-plane modes
    count twice, empty bins are 0 -- all as in the reference."""
    import torch

    on_device = is_torch_tensor(weights) and weights.is_cuda
    eng = Engine.get(weights.device if on_device else None)
    wd = eng.to_device(weights, torch.float32)
    n = int(n1d)
    if wd.ndim != 3 or wd.shape[0] != n or wd.shape[1] != n or wd.shape[2] < n // 2 + 1:
        raise ValueError(f'weights has shape {tuple(wd.shape)}, expected ({n},{n},>={n // 2 + 1})')
    sw, counts, sp, counts_poles, sk = _bin_device(eng, n, float(L), kedges, muedges, poles, fourier, real_in=wd,
                                                   row_len=wd.shape[2])
    dt = np.dtype(dtype)
    return sw.astype(dt), counts, sp.astype(dt), counts_poles, sk.astype(dt)


def project_3d_to_poles(k_bin_edges, raw_p3d, Lbox, poles):
    """Project a 3-D power array (rfftn layout) onto multipoles (power_spectrum.py:415-447).
    Returns ``(binned_poles * Lbox**3  (Np, Nk) float32, N_mode_poles (Nk,) int64)``."""
    assert np.max(poles) <= 10, 'numba implementation works up to ell = 10'
    nmesh = raw_p3d.shape[0]
    poles = np.asarray(poles)
    muedges = np.array([0.0, 1.0])
    _, _, binned_poles, Npoles, _ = bin_kmu(nmesh, Lbox, k_bin_edges, muedges, raw_p3d, poles=poles)
    binned_poles *= Lbox**3
    return binned_poles, Npoles


def _package_pk(binned, Lbox, mu_bin_edges, poles, squeeze_mu_axis):
    sw, counts, sp, counts_poles, sk = binned
    power = sw.astype(np.float32)
    k_avg = sk.astype(np.float32)
    binned_poles = sp.astype(np.float32)
    power *= Lbox**3  # power_spectrum.py:790
    if len(poles) > 0:
        binned_poles *= Lbox**3
    N_mode = counts
    if squeeze_mu_axis and len(mu_bin_edges) == 2:
        power = power[:, 0]
        N_mode = N_mode[:, 0]
        k_avg = k_avg[:, 0]
    return dict(power=power, N_mode=N_mode, binned_poles=binned_poles, N_mode_poles=counts_poles, k_avg=k_avg)


def calc_pk_from_deltak(field_fft, Lbox, k_bin_edges, mu_bin_edges, field2_fft=None, poles=np.empty(0, 'i8'),
                        squeeze_mu_axis=True, nthread=MAX_THREADS):
    """Bin ``|delta(k)|^2`` (or ``Re(conj(d1) d2)``) of supplied Fourier fields in (k, mu) and multipoles
    (power_spectrum.py:730-805).  Inputs are not modified.  Returns a dict with ``power``, ``N_mode``,
    ``binned_poles`` (Np, Nk), ``N_mode_poles``, ``k_avg``; ``power`` and the poles are multiplied by
    ``Lbox**3``."""
    import torch

    on_device = is_torch_tensor(field_fft) and field_fft.is_cuda
    eng = Engine.get(field_fft.device if on_device else None)
    f1 = eng.to_device(field_fft, torch.complex64)
    f2 = None if field2_fft is None else eng.to_device(field2_fft, torch.complex64)
    n = int(f1.shape[0])
    if tuple(f1.shape) != (n, n, n // 2 + 1) or (f2 is not None and f2.shape != f1.shape):
        raise ValueError(f'field_fft must have shape (n, n, n//2+1), got {tuple(f1.shape)}')
    poles = np.asarray(poles, dtype=np.int64)
    binned = _bin_device(eng, n, float(Lbox), k_bin_edges, mu_bin_edges, poles, True, f1=f1, f2=f2)
    return _package_pk(binned, Lbox, mu_bin_edges, poles, squeeze_mu_axis)


def calc_power(pos, Lbox, kbins=None, mubins=None, k_max=None, logk=False, paste='TSC', nmesh=128, compensated=True,
               interlaced=True, w=None, pos2=None, w2=None, poles=None, squeeze_mu_axis=True, nthread=MAX_THREADS,
               dtype=np.float32):
    r"""
    3-D power spectrum of a particle set in (k, mu) wedges and Legendre multipoles
    (reference: power_spectrum.py:1131-1319; same parameters and returned columns).

    Returns a table with columns ``k_min, k_max, k_mid`` (float64), ``k_avg, power`` (float32),
    ``N_mode`` (int64), and when ``poles`` is given ``poles`` (Nk, Np) float32 and ``N_mode_poles``;
    ``mu_min, mu_max, mu_mid`` when ``mubins`` is given.  ``meta`` records the run parameters.
    With ``pos2`` (and ``w2``) the cross spectrum of the two catalogues is returned.
    """
    import torch

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
    paste_u = _check_paste(paste)
    if w is not None:
        assert pos.shape[0] == len(w)
    if pos2 is not None and w2 is not None:
        assert pos2.shape[0] == len(w2)

    on_device = is_torch_tensor(pos) and pos.is_cuda
    eng = Engine.get(pos.device if on_device else None)
    n = int(nmesh)
    # ABK_META_TIMINGS=1: per-kernel CUDA-event times of THIS call in meta['kernel_ms'] (the reference prints stage times
    # when verbose, tsc.py:175-201; here they travel with the result)
    timings = os.environ.get('ABK_META_TIMINGS') == '1'
    if timings:
        eng.profile_collect()
        eng.profile(True)
    try:
        return _calc_power_impl(eng, on_device, pos, Lbox, kbins, mubins, k_max, logk, paste, paste_u, n, compensated, interlaced, w, pos2,
                                w2, poles, squeeze_mu_axis, dtype, return_mubins, meta)
    finally:
        if timings:
            meta['kernel_ms'] = {k: {'ms': v[0], 'launches': v[1]} for k, v in eng.profile_collect().items()}
            eng.profile(False)


def _calc_power_impl(eng, on_device, pos, Lbox, kbins, mubins, k_max, logk, paste, paste_u, n, compensated, interlaced, w, pos2, w2,
                     poles, squeeze_mu_axis, dtype, return_mubins, meta):
    import torch

    W = get_W_compensated(Lbox, n, paste, interlaced) if compensated else None
    W_d = eng.to_device(np.asarray(W, dtype=np.float32), torch.float32) if compensated else None

    if _is_f64(dtype) and not interlaced and not isinstance(pos, PackedParticles):
        return _calc_power_f64(eng, pos, Lbox, kbins, mubins, k_max, logk, paste_u, n, W, w, pos2, w2, poles, squeeze_mu_axis,
                               return_mubins, meta)
    g1 = _field_fft_device(eng, pos, Lbox, n, w, interlaced, tag='a', paste=paste_u)
    g2 = _field_fft_device(eng, pos2, Lbox, n, w2, interlaced, tag='a', paste=paste_u) if pos2 is not None else None
    if isinstance(pos, PackedParticles):      # pack9: the particle count is known once the records are decoded
        meta['N_pos'] = pos.n_particles
    if isinstance(pos2, PackedParticles):
        meta['N_pos2'] = pos2.n_particles

    poles_arr = np.asarray(poles or [], dtype=np.int64)
    kbins, mubins = get_k_mu_edges(Lbox, k_max, kbins, mubins, logk)
    scale = np.float32(0.5 / n**3) if interlaced else np.float32(1 / n**3)
    binned = _bin_device(eng, n, float(Lbox), kbins, mubins, poles_arr, True, f1=g1[0],
                         f1s=g1[1] if interlaced else None, f2=None if g2 is None else g2[0],
                         f2s=g2[1] if (g2 is not None and interlaced) else None, W_d=W_d, scale=scale, finish=True,
                         W_sym=(W is None) or bool(np.array_equal(np.asarray(W)[1:], np.asarray(W)[:0:-1])))
    P = _package_pk(binned, Lbox, mubins, poles_arr, squeeze_mu_axis)

    kbins = np.asarray(kbins)
    mubins = np.asarray(mubins)
    k_binc = (kbins[1:] + kbins[:-1]) * 0.5
    mu_binc = (mubins[1:] + mubins[:-1]) * 0.5
    res = dict(k_min=kbins[:-1], k_max=kbins[1:], k_mid=k_binc, k_avg=P['k_avg'], power=P['power'],
               N_mode=P['N_mode'])
    if len(poles_arr) > 0:
        res.update(poles=P['binned_poles'].T, N_mode_poles=P['N_mode_poles'])
    if return_mubins:
        res.update(mu_min=np.broadcast_to(mubins[:-1], res['power'].shape),
                   mu_max=np.broadcast_to(mubins[1:], res['power'].shape),
                   mu_mid=np.broadcast_to(mu_binc, res['power'].shape))
    return _make_table(res, meta)


def _calc_power_f64(eng, pos, Lbox, kbins, mubins, k_max, logk, paste, n, W, w, pos2, w2, poles, squeeze_mu_axis, return_mubins, meta):
    """calc_power(dtype=float64, interlaced=False): float64 field and transform (power_spectrum.py:1053-1069), raw power
    |f|^2 in float64 (:707-727), handed to the binning kernel as the float32 mesh whose sums the reference also keeps in
    float32 (bin_kmu is always called with dtype=float32, :781-783)."""
    import torch

    f1 = _field_fft_f64(eng, pos, Lbox, n, w, paste)
    f2 = _field_fft_f64(eng, pos2, Lbox, n, w2, paste) if pos2 is not None else None
    W_d = None if W is None else eng.to_device(np.asarray(W, dtype=np.float32), torch.float32)
    nzc = n // 2 + 1
    power = eng.empty((n, n, nzc), torch.float32)
    eng.bind_stream()
    check(eng.lib.abk_power_from_f64(eng.ctx, ptr(f1), ptr(f2), ptr(W_d), n, 1.0 / float(n) ** 3, ptr(power)))
    del f1, f2
    poles_arr = np.asarray(poles or [], dtype=np.int64)
    kbins, mubins = get_k_mu_edges(Lbox, k_max, kbins, mubins, logk)
    binned = _bin_device(eng, n, float(Lbox), kbins, mubins, poles_arr, True, real_in=power, row_len=nzc)
    P = _package_pk(binned, Lbox, mubins, poles_arr, squeeze_mu_axis)
    kbins, mubins = np.asarray(kbins), np.asarray(mubins)
    res = dict(k_min=kbins[:-1], k_max=kbins[1:], k_mid=(kbins[1:] + kbins[:-1]) * 0.5, k_avg=P['k_avg'], power=P['power'],
               N_mode=P['N_mode'])
    if len(poles_arr) > 0:
        res.update(poles=P['binned_poles'].T, N_mode_poles=P['N_mode_poles'])
    if return_mubins:
        mu_binc = (mubins[1:] + mubins[:-1]) * 0.5
        res.update(mu_min=np.broadcast_to(mubins[:-1], res['power'].shape),
                   mu_max=np.broadcast_to(mubins[1:], res['power'].shape),
                   mu_mid=np.broadcast_to(mu_binc, res['power'].shape))
    return _make_table(res, meta)


def pk_to_xi(Pk, Lbox, r_bins, poles=[0, 2, 4]):
    r"""
    Transform a 3-D power spectrum (rfftn layout, real, shape (N, N, N//2+1)) into correlation-function
    multipoles (reference: power_spectrum.py:620-660): ``Xi = irfftn(Pk)``, then ``bin_kmu(fourier=False)``
    over separations ``r = |index| * Lbox/N`` with a single mu bin.

    Returns ``(r_binc, binned_poles * N**3  (Np, Nr) float32, Npoles (Nr,) int64)``.
    """
    import torch

    on_device = is_torch_tensor(Pk) and Pk.is_cuda
    eng = Engine.get(Pk.device if on_device else None)
    r_bins = np.asarray(r_bins)
    pk_d = eng.to_device(Pk, torch.float32)
    n = int(pk_d.shape[0])
    if tuple(pk_d.shape) != (n, n, n // 2 + 1):
        raise ValueError(f'Pk must have shape (N, N, N//2+1), got {tuple(pk_d.shape)}')
    ldz = padded_ldz(n)
    buf = eng.empty((n, n, ldz), torch.float32)
    eng.bind_stream()
    check(eng.lib.abk_real_to_complex(eng.ctx, ptr(pk_d), ptr(buf), pk_d.numel()))
    plan, wb = eng.rfft3_plan(n, n, n)
    # the C2R plan is created on first use and may want a larger work area than the R2C one
    work = eng.scratch('fftwork', max(wb, 2 * buf.numel() * 4))
    check(eng.lib.abk_irfft3_exec(eng.ctx, plan, ptr(buf), ptr(work), work.numel()))
    r_binc = (r_bins[1:] + r_bins[:-1]) * 0.5
    poles = np.asarray(poles)
    muedges = np.array([0.0, 1.0])
    sw, counts, sp, counts_poles, sk = _bin_device(eng, n, float(Lbox), r_bins, muedges, poles, False, real_in=buf,
                                                   row_len=ldz)
    # the C2R transform is unnormalised: Xi = out / N^3 (scipy's irfftn normalises), power_spectrum.py:645
    binned_poles = (sp / float(n) ** 3).astype(np.float32)
    binned_poles *= n**3  # power_spectrum.py:659
    return r_binc, binned_poles, counts_poles


def get_delta_mu2(delta, n1d, dtype_c=np.complex64, dtype_f=np.float32):
    """``delta(k) * mu^2`` for a Fourier field of shape (n1d, n1d, n1d//2+1), complex64
    (reference: power_spectrum.py:577-617)."""
    import torch

    on_device = is_torch_tensor(delta) and delta.is_cuda
    eng = Engine.get(delta.device if on_device else None)
    eng.bind_stream()
    d = eng.to_device(delta, torch.complex64)
    n = int(n1d)
    if tuple(d.shape) != (n, n, n // 2 + 1):
        raise ValueError(f'delta must have shape ({n},{n},{n // 2 + 1}), got {tuple(d.shape)}')
    out = eng.empty(tuple(d.shape), torch.complex64)
    check(eng.lib.abk_delta_mu2(eng.ctx, ptr(d), ptr(out), n))
    return out if on_device else out.cpu().numpy()


def get_smoothing(n1d, L, R, dtype=np.float32):
    """Gaussian smoothing kernel ``exp(-k^2 R^2 / 2)`` on the (n1d, n1d, n1d//2+1) mesh, float32 NumPy array
    (reference: power_spectrum.py:539-574)."""
    import torch

    eng = Engine.get(None)
    eng.bind_stream()
    n = int(n1d)
    out = eng.empty((n, n, n // 2 + 1), torch.float32)
    check(eng.lib.abk_smoothing(eng.ctx, ptr(out), n, float(L), float(R)))
    return out.cpu().numpy().astype(dtype, copy=False)


def expand_poles_to_3d(k_ell, P_ell, n1d, L, poles, dtype=np.float32):
    """Expand multipoles P_l(k) to a 3-D mesh (n1d, n1d, n1d//2+1): ``sum_l interp(k_ell, P_l)(|k|) P_l(mu)``
    with linear interpolation on the uniform ``k_ell`` grid, clamped at both ends
    (reference: power_spectrum.py:450-536)."""
    import torch

    k_ell = np.asarray(k_ell)
    P_ell = np.asarray(P_ell)
    poles = np.asarray(poles, dtype=np.int64).reshape(-1)
    assert np.abs((k_ell[1] - k_ell[0]) - (k_ell[-1] - k_ell[-2])) < 1.0e-6
    if P_ell.ndim != 2 or P_ell.shape != (len(poles), len(k_ell)):
        raise ValueError(f'P_ell must have shape (len(poles), len(k_ell)), got {P_ell.shape}')
    eng = Engine.get(None)
    eng.bind_stream()
    n = int(n1d)
    coef = legendre_coefficients(poles).astype(np.float64)
    coef /= (2 * poles[:, None] + 1)  # plain P_l here, no (2l+1)
    tables = eng.to_device(np.concatenate([k_ell.astype(np.float32), P_ell.astype(np.float32).reshape(-1),
                                           coef.astype(np.float32).reshape(-1)]), torch.float32)
    Nk, Np = len(k_ell), len(poles)
    base = tables.data_ptr()
    out = eng.empty((n, n, n // 2 + 1), torch.float32)
    ph = (C.c_int32 * Np)(*[int(p) for p in poles])
    check(eng.lib.abk_expand_poles_to_3d(eng.ctx, ptr(out), n, float(L), C.c_void_p(base), C.c_void_p(base + 4 * Nk), Nk,
                                         ph, Np, C.c_void_p(base + 4 * Nk * (1 + Np))))
    return out.cpu().numpy().astype(dtype, copy=False)


def bin_kppi(n1d, L, kedges, pimax, Npi, weights, dtype=np.float32, fourier=True, nthread=MAX_THREADS):
    """Mean and mode count in (k_perp, pi) bins of a (n1d, n1d, n1d//2+1) half-spectrum, or of a real
    (n1d, n1d, n1d) mesh with ``fourier=False`` (reference: power_spectrum.py:303-412).

    Returns ``(weighted_counts, counts)``: the per-bin mean, ``dtype`` (Nk, Npi), and the int64 mode count.
    Bins are (lo, hi]; ``pi`` runs over ``linspace(0, pimax, Npi + 1)``.  As in the reference, a row of the
    mesh stops contributing at its first ``j`` whose k_perp reaches ``kedges[-1]``.  ``nthread`` is ignored.
    """
    import torch

    dtype = np.dtype(dtype).type
    if dtype not in (np.float32, np.float64):
        raise ValueError('dtype must be float32 or float64')
    on_device = is_torch_tensor(weights) and weights.is_cuda
    eng = Engine.get(weights.device if on_device else None)
    eng.bind_stream()
    n = int(n1d)
    wt = eng.to_device(weights, torch.float64 if dtype is np.float64 else torch.float32)
    if wt.ndim != 3 or wt.shape[0] != n or wt.shape[1] != n or wt.shape[2] < n // 2 + 1:
        raise ValueError(f'weights must have shape ({n},{n},>={n // 2 + 1}), got {tuple(wt.shape)}')
    Nk, Npi = len(kedges) - 1, int(Npi)
    dk = 2.0 * np.pi / L if fourier else L / n
    # squared edges rounded to the compute dtype (:365-366), handed over as float64 (exact)
    kedges2 = ((np.asarray(kedges, dtype=np.float64) / dk) ** 2).astype(dtype).astype(np.float64)
    piedges2 = ((np.linspace(0.0, pimax, Npi + 1) / dk) ** 2).astype(dtype).astype(np.float64)
    ke = eng.to_device(kedges2)
    pe = eng.to_device(piedges2)
    counts = eng.zeros((Nk, Npi), torch.int64)
    sums = eng.zeros((Nk, Npi), torch.float64)
    check(eng.lib.abk_bin_kppi(eng.ctx, ptr(wt), int(dtype is np.float64), n, int(wt.shape[2]), ptr(ke), Nk, ptr(pe), Npi,
                               int(dtype is np.float32), ptr(counts), ptr(sums)))
    c = counts.cpu().numpy()
    m = sums.cpu().numpy()
    nz = c != 0
    m[nz] /= c[nz]
    return m.astype(dtype), c


_ = (warnings, tsc_parallel)
