"""
Cloud-in-cell painting on the GPU -- drop-in for ``abacusnbody.analysis.cic.cic_serial`` (cic.py:13-125).

The reference function is serial; here it is the same bucket + tile deposit as TSC with the CIC weights
``(max(d, 0), 1 - |d|, max(-d, 0))`` of cic.py:43-67 (``abk_ctx_set_scheme(ctx, 1)``).
"""

import numpy as np

from .._lib import Engine, is_torch_tensor
from .tsc import deposit_device

__all__ = ['cic_serial']


def cic_serial(positions, density, boxsize, weights=None):
    """Accumulate the CIC density of ``positions`` (N, 3) into ``density`` (3-D, any shape; a last dimension of 1 makes
    it 2-D and ignores z), in place, with optional ``weights`` (N,).  No periodic wrap is applied to the positions
    (cic.py:29-42); cell indices wrap like the reference's."""
    import torch

    on_device = is_torch_tensor(positions) and positions.is_cuda
    eng = Engine.get(positions.device if on_device else None)
    eng.bind_stream()
    if density.ndim != 3:
        raise ValueError('density must be a 3-D array (use a last dimension of 1 for a 2-D grid)')
    shape = tuple(int(s) for s in density.shape)
    pos_d = eng.to_device(positions, torch.float32)
    if shape[2] == 1:
        pos_d = pos_d.clone()
        pos_d[:, 2] = 0.0
    w_d = None if weights is None else eng.to_device(weights, torch.float32)
    grid_is_cuda = is_torch_tensor(density) and density.is_cuda
    direct = grid_is_cuda and density.dtype == torch.float32 and density.is_contiguous()
    grid_d = density if direct else eng.zeros(shape, torch.float32)
    deposit_device(eng, pos_d, w_d, grid_d, shape, shape[2], boxsize, 0.0, wrap=False, scheme='CIC')
    if not direct:
        if is_torch_tensor(density):
            density += grid_d.to(density.device, density.dtype)
        else:
            density += grid_d.cpu().numpy().astype(density.dtype, copy=False)
    else:
        eng.sync()
