"""
Packed particle records as a direct input of the painter: ``calc_power(PackedParticles(raw, boxsize, ...), Lbox, ...)``.

End-to-end runs from host memory are bound by the PCIe copy of the positions (12 bytes per particle as float32).  Abacus
time slices are stored as pack9 (9 bytes per particle) and halo / subsample particles as RVint (12 bytes for position
*and* velocity); handing the painter the records still packed means they cross PCIe packed and are decoded on the GPU
(``abk_pack9_count`` / ``abk_pack9_decode`` / ``abk_unpack_rvint``), chunk by chunk, overlapped with the copy exactly like
plain positions.  pack9 chunks are cut at cell headers (first byte 0xFF), so every chunk decodes on its own.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from .._lib import check, ptr

__all__ = ['PackedParticles']


class PackedParticles:
    """Host-resident pack9 or RVint records.

    ``data``: pack9 -- uint8/int8, 9 bytes per record; RVint -- int32, 3 per particle (NumPy array or CPU torch tensor;
    pinned memory is copied asynchronously).  ``boxsize`` (and ``velzspace_to_kms`` for pack9) as for
    ``unpack_pack9`` / ``unpack_rvint``.  ``len()`` is the number of records, an upper bound of the number of particles
    for pack9 (some records are cell headers); ``n_particles`` is set once the records have been painted.
    """

    def __init__(self, data, boxsize, kind=None, velzspace_to_kms=1.0):
        import torch

        t = data if isinstance(data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asanyarray(data)))
        if t.is_cuda:
            raise ValueError('PackedParticles wraps host memory; decode device-resident records with unpack_pack9 / unpack_rvint')
        if kind is None:
            kind = 'rvint' if t.dtype == torch.int32 else 'pack9'
        if kind == 'pack9':
            if t.dtype == torch.int8:
                t = t.view(torch.uint8)
            if t.dtype != torch.uint8 or t.numel() % 9:
                raise ValueError('pack9 data must be uint8/int8 with 9 bytes per record')
            t = t.contiguous().view(-1, 9)
        elif kind == 'rvint':
            if t.dtype != torch.int32 or t.numel() % 3:
                raise ValueError('rvint data must be int32 with 3 values per particle')
            t = t.contiguous().view(-1, 3)
        else:
            raise ValueError(f'unknown packed format {kind!r}')
        self.kind, self.data = kind, t
        self.boxsize, self.velz = float(boxsize), float(velzspace_to_kms)
        self.rec_bytes = 9 if kind == 'pack9' else 12
        self.n_particles = int(t.shape[0]) if kind == 'rvint' else None
        self.is_cuda = False

    def __len__(self):
        return int(self.data.shape[0])

    @property
    def shape(self):
        return (len(self), 3)

    def chunk_plan(self, max_seg, chunk_min):
        """Record ranges of at most ``max_seg`` chunks; pack9 boundaries are moved back onto a cell header."""
        n = len(self)
        chunk = max(int(chunk_min), -(-n // max_seg), 1)
        cuts = list(range(0, n, chunk)) + [n]
        if self.kind == 'pack9':
            first = self.data[:, 0]
            for i in range(1, len(cuts) - 1):
                c = cuts[i]
                while c > cuts[i - 1] and int(first[c]) != 0xFF:
                    c -= 1
                cuts[i] = c
        return [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    def raw(self, a, b):
        """Flat byte view of records [a, b) (host tensor)."""
        import torch

        return self.data[a:b].reshape(-1).view(torch.uint8)

    def decode(self, eng, raw_dev, nrec, pos_dev):
        """Decode ``nrec`` records in ``raw_dev`` (device uint8) into ``pos_dev`` (device float32 (>=nrec, 3)); returns the
        number of particles.  pack9 synchronises the stream once (the header count)."""
        if nrec == 0:
            return 0
        if self.kind == 'rvint':
            check(eng.lib.abk_unpack_rvint(eng.ctx, ptr(raw_dev), nrec, self.boxsize, ptr(pos_dev), None, 0))
            return nrec
        nb = C.c_size_t()
        check(eng.lib.abk_pack9_scratch_bytes(nrec, C.byref(nb)))
        scratch = eng.scratch('pack9', nb.value + 256)
        sptr = C.c_void_p((scratch.data_ptr() + 255) & ~255)
        nhdr = C.c_int64()
        check(eng.lib.abk_pack9_count(eng.ctx, ptr(raw_dev), nrec, sptr, nb.value, C.byref(nhdr)))
        tab = eng.scratch('pack9_tab', max(nhdr.value, 1) * 20)
        check(eng.lib.abk_pack9_decode(eng.ctx, ptr(raw_dev), nrec, self.boxsize, self.velz, sptr, ptr(tab), nhdr.value,
                                       ptr(pos_dev), None, 0))
        return nrec - nhdr.value
