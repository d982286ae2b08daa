"""
ctypes binding of libabk.so (include/abk.h) and the per-device engine state.

PyTorch is used here only as the owner of device memory and streams (``torch.empty`` on the CUDA
device, ``torch.cuda.current_stream()``); every computation on the path goes through the C ABI.
There is no CPU fallback: if the library or a CUDA device is missing, calls raise.
"""

from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / 'libabk.so'

ABK_MAX_SEGMENTS = 16
ABK_MAX_POLES = 16
ABK_POLE_NCOEF = 11
SEGMENT_MAX = 1 << 30


class AbkError(RuntimeError):
    pass


class KMesh(C.Structure):
    _fields_ = [('n', C.c_int32), ('nzc', C.c_int32), ('i0', C.c_int32), ('i1', C.c_int32), ('j0', C.c_int32),
                ('j1', C.c_int32), ('stride_i', C.c_int64), ('stride_j', C.c_int64)]


class BinRequest(C.Structure):
    _fields_ = [('mesh', KMesh), ('f1', C.c_void_p), ('f1s', C.c_void_p), ('f2', C.c_void_p), ('f2s', C.c_void_p),
                ('real_in', C.c_void_p), ('W', C.c_void_p), ('scale', C.c_float), ('finish', C.c_int32),
                ('kedges2', C.c_void_p), ('muedges2', C.c_void_p), ('Nk', C.c_int32), ('Nmu', C.c_int32),
                ('Np', C.c_int32), ('pole_coef', C.c_void_p), ('pole_ell', C.c_int32 * ABK_MAX_POLES), ('counts', C.c_void_p),
                ('sum_p', C.c_void_p), ('sum_k', C.c_void_p), ('sum_poles', C.c_void_p), ('scratch', C.c_void_p),
                ('scratch_bytes', C.c_size_t), ('w_symmetric', C.c_int32)]


_vp, _i32, _i64, _dbl, _flt, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_float, C.c_size_t
_psz = C.POINTER(C.c_size_t)

# name -> (restype, argtypes); mirrors include/abk.h one to one (tests/test_abi.py checks that)
SIGNATURES = {
    'abk_version': (_i32, []),
    'abk_last_error': (C.c_char_p, []),
    'abk_ctx_create': (_i32, [_i32, C.POINTER(_vp)]),
    'abk_ctx_destroy': (_i32, [_vp]),
    'abk_ctx_set_stream': (_i32, [_vp, _vp]),
    'abk_ctx_sync': (_i32, [_vp]),
    'abk_ctx_launch_count': (_i64, [_vp]),
    'abk_ctx_set_tile_capacity': (_i32, [_vp, _i32]),
    'abk_ctx_set_scheme': (_i32, [_vp, _i32]),
    'abk_ctx_set_weight_scale': (_i32, [_vp, _dbl]),
    'abk_ctx_profile_enable': (_i32, [_vp, _i32]),
    'abk_ctx_profile_collect': (_i32, [_vp, C.POINTER(C.c_double), C.POINTER(_i64)]),
    'abk_kernel_count': (_i32, []),
    'abk_kernel_name': (C.c_char_p, [_i32]),
    'abk_wrap_inplace': (_i32, [_vp, _vp, _i64, _dbl, _vp]),
    'abk_partition_scratch_bytes': (_i32, [_i64, _i32, _psz]),
    'abk_partition': (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _i32, _vp, _vp, _vp, _vp, _vp, _sz]),
    'abk_tsc_num_tiles': (_i32, [_i32, _i32, _i32, C.POINTER(_i64)]),
    'abk_bench_red_rate': (_i32, [_vp, _vp, _i64, _i32, C.POINTER(C.c_double)]),
    'abk_tsc_tile_shape': (_i32, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    'abk_tsc_bucket_scratch_bytes': (_i32, [_i64, _i32, _i32, _i32, _psz]),
    'abk_tsc_bucket': (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _dbl, _dbl, _i32, _vp, _vp, _vp, _sz]),
    'abk_tsc_bucket_slab': (_i32, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _dbl, _dbl, _i32, _i32, _i32, _vp, _vp,
                                   _vp, _sz, C.POINTER(C.c_ulonglong)]),
    'abk_route_particles': (_i32, [_vp, _vp, _vp, _i64, _i32, _dbl, _i32, _i32, C.POINTER(C.c_int32), _vp,
                                   C.POINTER(_i64)]),
    'abk_tsc_deposit_tiles': (_i32, [_vp, _i32, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(_i64), _vp, _i32, _i32, _i32,
                                     _i64, _dbl, _dbl, _dbl, _i32, _i32, _i32]),
    'abk_tsc_deposit_scratch_bytes': (_i32, [_i64, _i32, _i32, _i32, _psz]),
    'abk_tsc_deposit': (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i64, _dbl, _dbl, _i32, _vp, _sz]),
    'abk_tsc_deposit_naive': (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i64, _dbl, _dbl, _i32]),
    'abk_tsc_deposit_typed': (_i32, [_vp, _vp, _i32, _vp, _i32, _i64, _vp, _i32, _i32, _i32, _i32, _i64, _dbl, _dbl, _i32]),
    'abk_normalize_field_f64': (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl]),
    'abk_rfft3_f64': (_i32, [_vp, _vp, _i64, _i64, _i64]),
    'abk_power_from_f64': (_i32, [_vp, _vp, _vp, _vp, _i32, _dbl, _vp]),
    'abk_normalize_field': (_i32, [_vp, _vp, _i64, _i64, _i64, _i64, _dbl, _dbl]),
    'abk_rfft3_plan_create': (_i32, [_vp, _i64, _i64, _i64, C.POINTER(_vp), _psz]),
    'abk_rfft3_exec': (_i32, [_vp, _vp, _vp, _vp, _sz]),
    'abk_irfft3_exec': (_i32, [_vp, _vp, _vp, _vp, _sz]),
    'abk_fft_plan_destroy': (_i32, [_vp]),
    'abk_fft_backend': (_i32, [C.c_char_p, _i32, C.POINTER(C.c_int)]),
    'abk_fft_yz_plan_create': (_i32, [_vp, _i64, _i64, _i64, C.POINTER(_vp), _psz]),
    'abk_fft_x_plan_create': (_i32, [_vp, _i64, _i64, _i64, C.POINTER(_vp), _psz]),
    'abk_fft_exec_generic': (_i32, [_vp, _vp, _vp, _vp, _sz]),
    'abk_field_fft_finish': (_i32, [_vp, C.POINTER(KMesh), _vp, _vp, _vp, _flt]),
    'abk_shift_field_fft': (_i32, [_vp, C.POINTER(KMesh), _vp, _vp, _dbl, _flt]),
    'abk_raw_power': (_i32, [_vp, _vp, _vp, _vp, _i64]),
    'abk_real_to_complex': (_i32, [_vp, _vp, _vp, _i64]),
    'abk_delta_mu2': (_i32, [_vp, _vp, _vp, _i32]),
    'abk_smoothing': (_i32, [_vp, _vp, _i32, _dbl, _dbl]),
    'abk_expand_poles_to_3d': (_i32, [_vp, _vp, _i32, _dbl, _vp, _vp, _i32, C.POINTER(C.c_int32), _i32, _vp]),
    'abk_bin_kppi': (_i32, [_vp, _vp, _i32, _i32, _i64, _vp, _i32, _vp, _i32, _i32, _vp, _vp]),
    'abk_unpack_rvint': (_i32, [_vp, _vp, _i64, _dbl, _vp, _vp, _i32]),
    'abk_unpack_pids': (_i32, [_vp, _vp, _i64, _dbl, _i64, _vp, _vp, _vp, _vp, _vp, _i32]),
    'abk_pack9_scratch_bytes': (_i32, [_i64, C.POINTER(C.c_size_t)]),
    'abk_pack9_count': (_i32, [_vp, _vp, _i64, _vp, C.c_size_t, C.POINTER(C.c_int64)]),
    'abk_pack9_decode': (_i32, [_vp, _vp, _i64, _dbl, _dbl, _vp, _vp, _i64, _vp, _vp, _i32]),
    'abk_power_bin_scratch_bytes': (_i32, [_i32, _i32, _i32, _psz]),
    'abk_power_bin': (_i32, [_vp, C.POINTER(BinRequest)]),
    'abk_add_planes': (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i64]),
    'abk_transpose_pack': (_i32, [_vp, _vp, _vp, _i64, _i64, _i64, _i32, C.POINTER(_i64)]),
    'abk_transpose_scatter_p2p': (_i32, [_vp, _vp, C.POINTER(_vp), _i64, _i64, _i64, _i32, C.POINTER(_i64), _i64]),
}

_lib = None
_lib_lock = threading.Lock()


def load_library():
    """Load libabk.so; raises AbkError if it has not been built (no fallback)."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise AbkError(f'{LIB_PATH} not found: build it with `python -m abacusutils_b200._build` '
                           '(or __graft_entry__.build()); there is no CPU fallback')
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        raise AbkError(f'libabk error {rc}: {load_library().abk_last_error().decode(errors="replace")}')


def _torch():
    import torch

    return torch


class Engine:
    """Per-device state: library context, cached FFT plans and reusable scratch buffers.  Not thread-safe: the context
    carries ONE current stream (``bind_stream``); use one Engine per host thread or serialise the calls."""

    _instances: dict = {}
    _lock = threading.Lock()

    @classmethod
    def get(cls, device=None):
        torch = _torch()
        if not torch.cuda.is_available():
            raise AbkError('no CUDA device visible: abacusutils_b200 has no CPU fallback')
        if device is None:
            device = torch.cuda.current_device()
        dev = torch.device(device) if not isinstance(device, int) else torch.device('cuda', device)
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        with cls._lock:
            if idx not in cls._instances:
                cls._instances[idx] = cls(idx)
            return cls._instances[idx]

    def __init__(self, index):
        torch = _torch()
        self.lib = load_library()
        self.index = index
        self.device = torch.device('cuda', index)
        h = C.c_void_p()
        check(self.lib.abk_ctx_create(index, C.byref(h)))
        self.ctx = h
        self._plans = {}
        self._bufs = {}
        knob = os.environ.get('ABK_TILE_KNOBS')      # experiments: raw argument of abk_ctx_set_tile_capacity
        if knob:
            check(self.lib.abk_ctx_set_tile_capacity(self.ctx, int(knob, 0)))

    # -- stream / memory plumbing ---------------------------------------------------------------
    def bind_stream(self):
        torch = _torch()
        s = torch.cuda.current_stream(self.device)
        check(self.lib.abk_ctx_set_stream(self.ctx, C.c_void_p(s.cuda_stream)))
        return s

    def aux_stream(self):
        """A second CUDA stream of this device (overlap of independent stages)."""
        if getattr(self, '_aux', None) is None:
            # high priority: its kernels (FFT of the previous grid, transposes) must get SMs while a long tile deposit
            # that fills the whole GPU is running on the main stream
            self._aux = _torch().cuda.Stream(device=self.device, priority=-1)
        return self._aux

    def sync(self):
        check(self.lib.abk_ctx_sync(self.ctx))

    def launch_count(self):
        return int(self.lib.abk_ctx_launch_count(self.ctx))

    def set_scheme(self, paste):
        """Select the mass-assignment scheme ('TSC' / 'CIC') of the following bucket / deposit calls."""
        check(self.lib.abk_ctx_set_scheme(self.ctx, 1 if str(paste).upper() == 'CIC' else 0))

    def set_weight_scale(self, scale):
        """Factor applied to every particle weight by the following bucket calls (1.0 = off)."""
        check(self.lib.abk_ctx_set_weight_scale(self.ctx, float(scale)))

    def profile(self, on):
        check(self.lib.abk_ctx_profile_enable(self.ctx, int(bool(on))))

    def profile_collect(self):
        """{kernel name: (total ms, launches)} since the last collect (synchronises the stream)."""
        nk = self.lib.abk_kernel_count()
        ms = (C.c_double * nk)()
        cnt = (C.c_int64 * nk)()
        self.bind_stream()
        check(self.lib.abk_ctx_profile_collect(self.ctx, ms, cnt))
        return {self.lib.abk_kernel_name(i).decode(): (ms[i], cnt[i]) for i in range(nk) if cnt[i]}

    def red_rate(self, nfloats=1 << 28):
        """Measured float-reduction rates of this GPU (1e9 adds/s): {'coalesced': rows of 32 floats, 'scattered': ...}."""
        torch = _torch()
        buf = torch.zeros(nfloats, dtype=torch.float32, device=self.device)
        out = {}
        self.bind_stream()
        for mode, name in ((0, 'coalesced'), (1, 'scattered')):
            r = C.c_double()
            check(self.lib.abk_bench_red_rate(self.ctx, C.c_void_p(buf.data_ptr()), nfloats, mode, C.byref(r)))
            out[name] = r.value
        return out

    def empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self.device)

    def zeros(self, shape, dtype):
        return _torch().zeros(shape, dtype=dtype, device=self.device)

    def scratch(self, key, nbytes):
        """A reusable uint8 device buffer of at least nbytes (grown on demand, kept per key)."""
        torch = _torch()
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes:
            if buf is not None and buf.is_cuda:
                # scratch is shared between the compute, copy and auxiliary streams without allocator bookkeeping: before
                # a buffer is handed back (regrowth, rare) nothing may still be using it on any stream
                torch.cuda.synchronize(self.device)
            self._bufs.pop(key, None)
            buf = None
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=self.device)
            self._bufs[key] = buf
        return buf

    def release_scratch(self):
        if self._bufs and next(iter(self._bufs.values())).is_cuda:
            _torch().cuda.synchronize(self.device)
        self._bufs.clear()

    def to_device(self, arr, dtype=None):
        """numpy / torch (any device) -> contiguous torch tensor on this device."""
        torch = _torch()
        if isinstance(arr, torch.Tensor):
            t = arr
        else:
            a = np.ascontiguousarray(arr)
            t = torch.from_numpy(a)
        if dtype is not None and t.dtype != dtype:
            t = t.to(dtype)
        return t.to(self.device, non_blocking=True).contiguous()

    # -- FFT plans ----------------------------------------------------------------------------------
    def rfft3_plan(self, nx, ny, nz):
        key = ('r3', nx, ny, nz)
        if key not in self._plans:
            h, wb = C.c_void_p(), C.c_size_t()
            check(self.lib.abk_rfft3_plan_create(self.ctx, nx, ny, nz, C.byref(h), C.byref(wb)))
            self._plans[key] = (h, int(wb.value))
        return self._plans[key]

    def rfft3_inplace(self, grid_padded, nx, ny, nz):
        plan, wb = self.rfft3_plan(nx, ny, nz)
        work = self.scratch('fftwork', wb)
        self.bind_stream()
        check(self.lib.abk_rfft3_exec(self.ctx, plan, C.c_void_p(grid_padded.data_ptr()), C.c_void_p(work.data_ptr()),
                                      work.numel()))


def is_torch_tensor(x):
    try:
        import torch
    except ImportError:  # pragma: no cover
        return False
    return isinstance(x, torch.Tensor)


def ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())
