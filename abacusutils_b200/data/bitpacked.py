"""
RVint decoding on the GPU (reference: abacusnbody/data/bitpacked.py:29-120).

``unpack_rvint`` keeps the reference's signature and return convention.  NumPy in -> NumPy out; a torch CUDA
tensor in -> the unpacked arrays stay on the device, ready for ``abacusutils_b200.analysis`` (so a catalogue can
be copied to the GPU still packed, 12 bytes per particle for positions *and* velocities).
"""

import numpy as np

from .._lib import check, ptr
from ._common import Output, engine_for, torch_float

__all__ = ['unpack_rvint']


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
