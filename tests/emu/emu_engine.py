"""
Run the PRODUCT's Python layer (abacusutils_b200.analysis.*, .data.*) on the CPU emulator build of the kernels
(TEST INFRASTRUCTURE ONLY).  `install(monkeypatch)` swaps three things and nothing else:

  * `Engine.get` returns an engine whose library is libabk_emu.so (the unmodified kernel sources, one fiber per CUDA
    thread, see include/cuda_runtime.h) and whose buffers are CPU torch tensors;
  * the cuFFT entry points, which are not kernels of ours, are answered by scipy on the same in-place padded layout;
  * torch.cuda.Stream / Event / stream(), used by the painter for copy/compute overlap, become no-ops.

Everything else -- bucketing, tile deposit, normalisation, interlacing, binning, the host-side orchestration -- is the
code that runs on the B200.
"""

import contextlib
import ctypes as C
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))


def _view(ptr, dtype, count):
    addr = ptr.value if hasattr(ptr, 'value') else int(ptr)
    raw = np.ctypeslib.as_array(C.cast(addr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,))
    return raw.view(dtype)


class EmuLib:
    """libabk_emu.so with the product's ctypes signatures; FFT entry points answered by scipy."""

    def __init__(self, so_path):
        from abacusutils_b200 import _lib

        self._cdll = C.CDLL(str(so_path))
        for name, (res, args) in _lib.SIGNATURES.items():
            if hasattr(self._cdll, name):
                fn = getattr(self._cdll, name)
                fn.restype, fn.argtypes = res, args
        self._plans = {}

    def __getattr__(self, name):
        return getattr(self._cdll, name)

    # -- cuFFT stand-ins (abk_fft.cu is a thin wrapper over a vendor library, not a kernel of this repo) --------
    def abk_rfft3_plan_create(self, ctx, nx, ny, nz, plan_out, work_bytes_out):
        h = len(self._plans) + 1
        self._plans[h] = (nx, ny, nz)
        plan_out._obj.value = h
        work_bytes_out._obj.value = 256
        return 0

    def _shape(self, plan):
        return self._plans[plan.value if hasattr(plan, 'value') else int(plan)]

    def abk_rfft3_exec(self, ctx, plan, grid, work, work_bytes):
        from scipy.fft import rfftn

        nx, ny, nz = self._shape(plan)
        ldz = 2 * (nz // 2 + 1)
        real = _view(grid, np.float32, nx * ny * ldz).reshape(nx, ny, ldz)
        spec = rfftn(real[:, :, :nz].astype(np.float32)).astype(np.complex64)
        _view(grid, np.complex64, nx * ny * (nz // 2 + 1)).reshape(nx, ny, nz // 2 + 1)[...] = spec
        return 0

    def abk_rfft3_f64(self, ctx, grid, nx, ny, nz):
        from scipy.fft import rfftn

        ldz = 2 * (nz // 2 + 1)
        real = _view(grid, np.float64, nx * ny * ldz).reshape(nx, ny, ldz)
        spec = rfftn(real[:, :, :nz].copy())
        _view(grid, np.complex128, nx * ny * (nz // 2 + 1)).reshape(nx, ny, nz // 2 + 1)[...] = spec
        return 0

    def abk_irfft3_exec(self, ctx, plan, grid, work, work_bytes):
        from scipy.fft import irfftn

        nx, ny, nz = self._shape(plan)
        ldz = 2 * (nz // 2 + 1)
        spec = _view(grid, np.complex64, nx * ny * (nz // 2 + 1)).reshape(nx, ny, nz // 2 + 1).copy()
        real = irfftn(spec, s=(nx, ny, nz)).astype(np.float32) * np.float32(nx * ny * nz)   # cuFFT C2R is unnormalised
        _view(grid, np.float32, nx * ny * ldz).reshape(nx, ny, ldz)[:, :, :nz] = real
        return 0

    def abk_fft_yz_plan_create(self, ctx, nplanes, ny, nz, plan_out, work_bytes_out):
        h = len(self._plans) + 1
        self._plans[h] = ('yz', int(nplanes), int(ny), int(nz))
        plan_out._obj.value = h
        work_bytes_out._obj.value = 256
        return 0

    def abk_fft_x_plan_create(self, ctx, nx, nrows, nzc, plan_out, work_bytes_out):
        h = len(self._plans) + 1
        self._plans[h] = ('x', int(nx), int(nrows), int(nzc))
        plan_out._obj.value = h
        work_bytes_out._obj.value = 256
        return 0

    def abk_fft_exec_generic(self, ctx, plan, data, work, work_bytes):
        from scipy.fft import fft, rfft2

        kind, a, b, c = self._shape(plan)
        if kind == 'yz':      # 2-D R2C over (y, z) on a planes, in place, ldz = 2 (nz/2 + 1)
            ldz = 2 * (c // 2 + 1)
            real = _view(data, np.float32, a * b * ldz).reshape(a, b, ldz)
            spec = rfft2(real[:, :, :c].astype(np.float32), axes=(1, 2)).astype(np.complex64)
            _view(data, np.complex64, a * b * (c // 2 + 1)).reshape(a, b, c // 2 + 1)[...] = spec
        else:                 # 1-D C2C along x of [x][row][nzc]
            arr = _view(data, np.complex64, a * b * c).reshape(a, b, c)
            arr[...] = fft(arr.copy(), axis=0).astype(np.complex64)
        return 0

    def abk_fft_plan_destroy(self, plan):
        return 0


class _NullStream:
    cuda_stream = 0

    def wait_event(self, ev):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass


class _NullEvent:
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass

    def wait(self, stream=None):
        pass


_STATE = {}


def emu_library(build_dir):
    import build_emu

    if 'lib' not in _STATE:
        _STATE['lib'] = EmuLib(build_emu.build(build_dir))
    return _STATE['lib']


def install(monkeypatch, build_dir):
    """Route the product's Python layer to the emulator for the duration of a test."""
    import torch

    from abacusutils_b200 import _lib

    lib = emu_library(build_dir)

    class EmuEngine(_lib.Engine):
        def __init__(self):
            self.lib = lib
            self.index = 0
            self.device = torch.device('cpu')
            h = C.c_void_p()
            assert lib.abk_ctx_create(0, C.byref(h)) == 0
            self.ctx = h
            self._plans = {}
            self._bufs = {}

        def bind_stream(self):
            return _NullStream()

        def aux_stream(self):
            return _NullStream()

        def to_device(self, arr, dtype=None):
            # a host->device copy never aliases the caller's array; keep that property on the CPU
            t = super().to_device(arr, dtype)
            return t.clone() if isinstance(arr, torch.Tensor) or t.data_ptr() == np.asarray(arr).ctypes.data else t

    if 'engine' not in _STATE:
        _STATE['engine'] = EmuEngine()
    eng = _STATE['engine']
    monkeypatch.setattr(_lib.Engine, 'get', classmethod(lambda cls, device=None: eng))
    monkeypatch.setattr(_lib, '_lib', lib)                      # check() reads abk_last_error() from here
    monkeypatch.setattr(torch.cuda, 'Stream', lambda *a, **k: _NullStream())
    monkeypatch.setattr(torch.cuda, 'Event', _NullEvent)
    monkeypatch.setattr(torch.cuda, 'stream', lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_device', lambda: 0)
    return eng
