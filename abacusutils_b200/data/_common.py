"""Shared argument handling of the device-side decoders."""

import numpy as np

from .._lib import Engine, is_torch_tensor


def torch_float(float_dtype):
    import torch

    dt = np.dtype(float_dtype)
    if dt == np.float32:
        return torch.float32, 0
    if dt == np.float64:
        return torch.float64, 1
    raise ValueError(f'float_dtype must be float32 or float64, got {dt}')


def engine_for(data):
    on_device = is_torch_tensor(data) and data.is_cuda
    eng = Engine.get(data.device if on_device else None)
    eng.bind_stream()
    return eng, on_device


class Output:
    """One of ``posout`` / ``velout``: ``None`` -> allocate and return the array, ``False`` -> skip and return 0,
    an array -> fill its first rows and return the row count (bitpacked.py:63-99, pack9.py:22-55)."""

    def __init__(self, eng, spec, nrows, tdtype, on_device):
        self.eng, self.spec, self.nrows, self.on_device = eng, spec, nrows, on_device
        self.skip = spec is False
        self.user = None if (spec is None or spec is False) else spec
        self.dev = None
        if self.skip:
            return
        if self.user is not None and is_torch_tensor(self.user) and self.user.is_cuda:
            u = self.user
            if u.dtype != tdtype or not u.is_contiguous() or u.numel() < 3 * nrows:
                raise ValueError('output tensor must be contiguous, of float_dtype, with at least 3*N elements')
            self.dev = u.view(-1)[: 3 * nrows].view(nrows, 3)
            self.direct = True
        else:
            if self.user is not None:
                a = np.asarray(self.user)   # shares memory with NumPy arrays and CPU tensors
                if a.size < 3 * nrows:
                    raise ValueError('output array too small')
                self.user = a
            self.dev = eng.empty((nrows, 3), tdtype)
            self.direct = False

    def pointer(self):
        from .._lib import ptr

        return ptr(None if self.skip else self.dev)

    def result(self):
        if self.skip:
            return 0
        if self.user is None:
            return self.dev if self.on_device else self.dev.cpu().numpy()
        if not self.direct:
            host = self.dev.cpu().numpy()
            out = self.user.view()
            out.shape = (-1, 3)
            out[: self.nrows] = host
        return self.nrows
