"""
pack9 decoding on the GPU (reference: abacusnbody/data/pack9.py:16-123).

A pack9 stream mixes 9-byte particle records with 9-byte cell headers (first byte 0xFF); the reference walks it
serially.  Here the headers are counted and scanned on the device (``abk_pack9_count``), decoded into a table, and
every particle record looks its header up by the number of headers before it (``abk_pack9_decode``).  NumPy in ->
NumPy out; a torch CUDA uint8 tensor in -> results stay on the device.
"""

import ctypes as C

import numpy as np

from .._lib import check, ptr
from ._common import Output, engine_for, torch_float

__all__ = ['unpack_pack9']


def unpack_pack9(data, boxsize, velzspace_to_kms, float_dtype=np.float32, posout=None, velout=None):
    """Unpack pack9 records into pos and vel (pack9.py:16-55).

    ``data``: uint8 (Nmax, 9); some records are cell headers.  Returns ``(pos, vel)`` with ``npart`` rows each, or
    ``npart`` in place of an array that was supplied (it is filled in its first ``npart`` rows), or 0 for
    ``False``.  Values are bit-identical to the reference for ``float_dtype`` float32 and float64.
    """
    import torch

    eng, on_device = engine_for(data)
    tdtype, f64 = torch_float(float_dtype)
    if on_device:
        d = data
        if d.dtype == torch.int8:
            d = d.view(torch.uint8)
        if d.dtype != torch.uint8:
            raise ValueError('pack9 data must be uint8')
        d = d.contiguous().view(-1, 9)
        if d.data_ptr() % 4:
            d = d.clone()
    else:
        d = eng.to_device(np.ascontiguousarray(np.asanyarray(data).view(np.uint8) if np.asanyarray(data).dtype == np.int8
                                               else np.asanyarray(data, dtype=np.ubyte)).reshape(-1, 9))
    nrec = int(d.shape[0])
    nb = C.c_size_t()
    check(eng.lib.abk_pack9_scratch_bytes(nrec, C.byref(nb)))
    scratch = eng.scratch('pack9', nb.value + 256)
    sptr = C.c_void_p((scratch.data_ptr() + 255) & ~255)   # the scan wants 256-byte alignment whatever the allocator gives
    nhdr = C.c_int64()
    check(eng.lib.abk_pack9_count(eng.ctx, ptr(d), nrec, sptr, nb.value, C.byref(nhdr)))
    nhdr = nhdr.value
    npart = nrec - nhdr
    hdr_tab = eng.empty((max(nhdr, 1), 5), tdtype)
    pos = Output(eng, posout, npart, tdtype, on_device)
    vel = Output(eng, velout, npart, tdtype, on_device)
    check(eng.lib.abk_pack9_decode(eng.ctx, ptr(d), nrec, float(boxsize), float(velzspace_to_kms), sptr,
                                   ptr(hdr_tab), nhdr, pos.pointer(), vel.pointer(), f64))
    return pos.result(), vel.result()
