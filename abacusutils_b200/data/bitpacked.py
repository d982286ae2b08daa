"""
RVint decoding on the GPU (reference: abacusnbody/data/bitpacked.py:29-120).

``unpack_rvint`` keeps the reference's signature and return convention.  NumPy in -> NumPy out; a torch CUDA
tensor in -> the unpacked arrays stay on the device, ready for ``abacusutils_b200.analysis`` (so a catalogue can
be copied to the GPU still packed, 12 bytes per particle for positions *and* velocities).
"""

import numpy as np

from .._lib import check, ptr
from ._common import Output, engine_for, torch_float

__all__ = ['unpack_rvint', 'unpack_pids']

# the names of the bit-packed PID fields that can be requested (bitpacked.py:26)
PID_FIELDS = ['pid', 'lagr_pos', 'tagged', 'density', 'lagr_idx', 'packedpid']


def unpack_rvint(intdata, boxsize, float_dtype=np.float32, posout=None, velout=None):
    """Unpack rvint data into pos and vel (bitpacked.py:29-98).

    ``intdata``: int32, any shape with 3*N elements.  ``posout`` / ``velout``: ``None`` (allocate and return),
    ``False`` (do not unpack, returns 0) or an array to fill (returns N).  Values are bit-identical to the
    reference: ``(x >> 12) * (boxsize / 1e6)`` and ``((x & 0xfff) - 2048) * (6000 / 2048)`` in float64, rounded
    once to ``float_dtype``.
    """
    import torch

    eng, on_device = engine_for(intdata)
    tdtype, f64 = torch_float(float_dtype)
    if on_device:
        if intdata.dtype != torch.int32:
            raise AssertionError('intdata must be int32')
        d = intdata.contiguous().view(-1, 3)
    else:
        a = np.asarray(intdata)
        assert a.dtype == np.int32
        d = eng.to_device(a.reshape(-1, 3))
    N = int(d.shape[0])
    pos = Output(eng, posout, N, tdtype, on_device)
    vel = Output(eng, velout, N, tdtype, on_device)
    check(eng.lib.abk_unpack_rvint(eng.ctx, ptr(d), N, float(boxsize), pos.pointer(), vel.pointer(), f64))
    return pos.result(), vel.result()


def unpack_pids(packed, box=None, ppd=None, pid=False, lagr_pos=False, tagged=False, density=False, lagr_idx=False,
                float_dtype=np.float32):
    """Extract fields from the bit-packed PIDs / aux words (bitpacked.py:123-220).

    Returns a dict with the requested arrays: ``pid`` int64 (N,), ``lagr_pos`` ``float_dtype`` (N,3) (needs ``box``
    and ``ppd``), ``lagr_idx`` int16 (N,3), ``tagged`` uint8 (N,), ``density`` ``float_dtype`` (N,).  NumPy in ->
    NumPy out; a torch CUDA tensor (int64 or uint64) in -> CUDA tensors out.
    """
    import torch

    if lagr_pos is not False:
        if box is None:
            raise ValueError('Must supply `box` if requesting `lagr_pos`')
        if ppd is None:
            raise ValueError('Must supply `ppd` if requesting `lagr_pos`')
    if ppd is not None:
        if not np.isclose(ppd, int(round(ppd))):
            raise ValueError(f'ppd "{ppd}" not valid int?')
        ppd = int(round(ppd))
    else:
        ppd = 1
    if box is None:
        box = 1.0

    eng, on_device = engine_for(packed)
    tdtype, f64 = torch_float(float_dtype)
    if on_device:
        d = packed
        if d.dtype == torch.uint64:
            d = d.view(torch.int64)
        if d.dtype != torch.int64:
            raise ValueError('packed PIDs must be a 64-bit integer tensor')
        d = d.contiguous().view(-1)
    else:
        d = eng.to_device(np.ascontiguousarray(np.asanyarray(packed, dtype=np.uint64)).view(np.int64).reshape(-1))
    N = int(d.shape[0])
    arr = {}
    if pid is True:
        arr['pid'] = eng.empty((N,), torch.int64)
    if lagr_pos is True:
        arr['lagr_pos'] = eng.empty((N, 3), tdtype)
    if lagr_idx is True:
        arr['lagr_idx'] = eng.empty((N, 3), torch.int16)
    if tagged is True:
        arr['tagged'] = eng.empty((N,), torch.uint8)
    if density is True:
        arr['density'] = eng.empty((N,), tdtype)
    check(eng.lib.abk_unpack_pids(eng.ctx, ptr(d), N, float(box), ppd, ptr(arr.get('pid')), ptr(arr.get('lagr_pos')),
                                  ptr(arr.get('lagr_idx')), ptr(arr.get('tagged')), ptr(arr.get('density')), f64))
    if on_device:
        return arr
    return {k: v.cpu().numpy() for k, v in arr.items()}
