"""
Triangular-shaped-cloud mass assignment on B200 -- drop-in for ``abacusnbody.analysis.tsc``.

Same public names, argument order, defaults and error behaviour as the reference
(/root/reference/abacusnbody/analysis/tsc.py:10-22, :260-268); the work is done by libabk.so
(abacusutils_b200/csrc/abk_tsc.cu) through ctypes.  Inputs may be NumPy arrays (host; copied to the
device) or ``torch`` CUDA tensors (used in place, zero-copy).  Results come back in the kind of the
input: NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out.

Differences from the reference, all deliberate:
  * ``nthread``, ``npartition``, ``sort`` (for ``tsc_parallel``) are accepted and validated like the
    reference but do not influence the result: the GPU schedule (tile bucketing) replaces the
    x-stripe schedule they tune.
  * float64 positions / grids / weights are honoured like the reference does (same warning, tsc.py:155-165):
    arithmetic in the dtype of the positions (tsc.py:400), accumulation in the dtype of the grid.  Only
    float32 positions on a float32 grid take the tuned path (tile bucketing + walk kernel); every other
    combination runs the typed one-thread-per-particle kernel of csrc/abk_f64.cu.
  * the sum order differs (per-cell register sums, float reductions into the grid), so grids agree
    with the reference to float32 round-off, not bit for bit -- as between two thread counts of the
    reference itself (its kernels are ``fastmath=True``).
"""

from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from .._lib import ABK_MAX_SEGMENTS, SEGMENT_MAX, AbkError, Engine, check, is_torch_tensor, ptr

__all__ = ['tsc_parallel', 'partition_parallel']


def _check_dtype(a, name):
    itemsize = a.element_size() if is_torch_tensor(a) else a.itemsize
    if itemsize > 4:
        warnings.warn(f'{name}.dtype={a.dtype} instead of np.float32. float32 is recommended for performance.')


def _validate_npartition(npartition, n1d, nthread):
    """tsc.py:126-147 -- kept so that callers see the same ValueErrors."""
    if not npartition:
        return
    if npartition > n1d // 3 and npartition != n1d // 2 and nthread > 1:
        raise ValueError(f'npartition {npartition} must be less than ngrid//3 = {n1d // 3} or equal to '
                         f'ngrid//2 = {n1d // 2}')
    if npartition > 1 and npartition % 2 != 0 and nthread > 1:
        raise ValueError(f'npartition {npartition} not divisible by 2')


def _float_kind(a):
    """'f8' for float64 arrays / tensors, 'f4' otherwise (other dtypes are converted to float32, as before)."""
    if is_torch_tensor(a):
        import torch

        return 'f8' if a.dtype == torch.float64 else 'f4'
    return 'f8' if np.asarray(a).dtype == np.float64 else 'f4'


def _wrap_caller_array(pos, box):
    """One-shot periodic wrap of the caller's array (NumPy or torch, any float dtype), changed entries only:
    ``>= box -> -= box``, ``< 0 -> += box``, evaluated in float64 and stored in the array's dtype (tsc.py:219-226)."""
    if is_torch_tensor(pos):
        import torch

        hi, lo = pos >= box, pos < 0
        if bool(hi.any()):
            pos[hi] = (pos[hi].to(torch.float64) - box).to(pos.dtype)
        if bool(lo.any()):
            pos[lo] = (pos[lo].to(torch.float64) + box).to(pos.dtype)
        return
    hi, lo = pos >= box, pos < 0
    if hi.any():
        pos[hi] = (pos[hi].astype(np.float64) - box).astype(pos.dtype)
    if lo.any():
        pos[lo] = (pos[lo].astype(np.float64) + box).astype(pos.dtype)


def padded_ldz(nz):
    """Row length (floats) of the in-place R2C layout."""
    return 2 * (nz // 2 + 1)


def deposit_device(eng, pos_d, w_d, grid_d, shape, ldz, box, offset, wrap, scheme='TSC'):
    """Bucket + deposit device-resident particles into a device grid (accumulating)."""
    import torch

    lib = eng.lib
    N = int(pos_d.shape[0])
    if N == 0:
        return
    nx, ny, nz = shape
    nb = C.c_size_t()
    check(lib.abk_tsc_deposit_scratch_bytes(N, nx, ny, nz, C.byref(nb)))
    scratch = eng.scratch('deposit', nb.value)
    eng.bind_stream()
    eng.set_scheme(scheme)
    assert pos_d.dtype == torch.float32 and pos_d.is_contiguous()
    check(lib.abk_tsc_deposit(eng.ctx, ptr(pos_d), ptr(w_d), N, ptr(grid_d), nx, ny, nz, ldz, float(box),
                              float(offset), int(bool(wrap)), ptr(scratch), scratch.numel()))


def tsc_parallel(pos, densgrid, box, weights=None, nthread=-1, wrap=True, npartition=None, sort=False, coord=0,
                 verbose=False, offset=0.0):
    """
    TSC-paint particles onto a 3-D grid (reference: tsc.py:10-206).

    Parameters are those of the reference.  ``densgrid`` may be an int (cubic grid), a tuple
    (grid shape), a NumPy array or a torch CUDA tensor; an existing grid is ACCUMULATED into, never
    zeroed (tsc.py:45-50).  Returns the new grid if one was allocated, else ``None`` (tsc.py:204-206).
    With ``wrap=True`` positions outside ``[0, box)`` are wrapped once, IN PLACE, like the
    reference's ``_wrap_inplace`` (for a host array the write-back happens only if a value changed).
    """
    import torch

    if nthread is None or nthread < 0:
        nthread = 2  # only used to mirror the reference's argument validation
    on_device = is_torch_tensor(pos) and pos.is_cuda
    eng = Engine.get(pos.device if on_device else None)

    if isinstance(densgrid, (int, np.integer)):
        densgrid = (int(densgrid),) * 3
    user_supplied_grid = not isinstance(densgrid, tuple)
    shape = tuple(int(s) for s in (densgrid if not user_supplied_grid else densgrid.shape))
    if len(shape) not in (2, 3):
        raise ValueError(f'densgrid must be 2-D or 3-D, got shape {shape}')
    two_d = len(shape) == 2
    if two_d:
        # tsc.py:452-468: a 2-D grid gets the 9-point stencil with w_z = 1.  Here it is painted as a
        # (nx, ny, 1) grid with z = 0: the three z-weights of the 27-point stencil all land on the single
        # plane and sum to 1 (to float32 round-off).
        shape3 = shape + (1,)
    else:
        shape3 = shape
    if coord != 0:
        # the partition coordinate only steers the reference's CPU schedule; results do not depend on it
        if coord not in (1, 2):
            raise ValueError(f'coord {coord} out of range')
    _validate_npartition(npartition, shape[coord], nthread)
    if pos.ndim != 2 or pos.shape[1] not in ((2, 3) if two_d else (3,)):
        raise ValueError(f'pos must have shape (N, 3), got {tuple(pos.shape)}')
    if weights is not None and len(weights) != len(pos):
        raise ValueError('weights and pos have different lengths')
    if max(shape) > 32767:
        raise ValueError('grid dimensions are limited to 32767 (int16 cell indices in the reference)')

    _check_dtype(pos, 'pos')
    if user_supplied_grid:
        _check_dtype(densgrid, 'densgrid')
    if weights is not None:
        _check_dtype(weights, 'weights')

    N = len(pos)
    stream = eng.bind_stream()
    pos_dt = _float_kind(pos)
    grid_dt = _float_kind(densgrid) if user_supplied_grid else 'f4'      # a new grid is float32 (tsc.py:118-120)
    w_dt = None if weights is None else _float_kind(weights)
    typed = not (pos_dt == 'f4' and grid_dt == 'f4' and w_dt in (None, 'f4'))
    tpos = torch.float64 if pos_dt == 'f8' else torch.float32
    # ---- particles on the device ------------------------------------------------------------------
    if on_device:
        pos_d = pos if (pos.dtype == tpos and pos.is_contiguous()) else pos.to(tpos).contiguous()
    else:
        pos_d = eng.to_device(pos, tpos)
    w_d = None if weights is None else eng.to_device(weights, torch.float64 if w_dt == 'f8' else torch.float32)
    pos_in = pos_d
    if two_d:
        # (x, y, 0) copy for the 3-D kernels; the in-place wrap below is applied to the caller's columns
        pos_d = torch.zeros((N, 3), dtype=tpos, device=eng.device)
        pos_d[:, :2] = pos_in[:, :2]

    if wrap and N > 0:
        # tsc.py:171-173: the caller's array is wrapped in place.  The device copy is wrapped by the kernel (that is what
        # gets painted); the CALLER's array is then updated entry by entry, in its own dtype and only where a value lies
        # outside [0, box) -- like _wrap_inplace (tsc.py:219-226), which never touches in-range entries.
        if typed:
            changed = int(((pos_in >= box) | (pos_in < 0)).sum().item())
        else:
            flag = eng.zeros((1,), torch.int64)
            check(eng.lib.abk_wrap_inplace(eng.ctx, ptr(pos_d), N, float(box), ptr(flag)))
            changed = int(flag.item())
            if two_d and pos.shape[1] == 3:
                # the reference wraps every column of the caller's array, also the third one a 2-D grid never reads
                zc = pos_in[:, 2]
                changed += int(((zc >= box) | (zc < 0)).sum().item())
        if changed and (typed or pos_d is not pos):
            _wrap_caller_array(pos, float(box))

    # ---- grid on the device -------------------------------------------------------------------------
    tgrid = torch.float64 if grid_dt == 'f8' else torch.float32
    grid_is_cuda = user_supplied_grid and is_torch_tensor(densgrid) and densgrid.is_cuda
    if grid_is_cuda and densgrid.dtype == tgrid and densgrid.is_contiguous():
        grid_d = densgrid
    else:
        grid_d = eng.zeros(shape, tgrid)
    if typed:
        if N > 0:
            eng.bind_stream()
            eng.set_scheme('TSC')
            check(eng.lib.abk_tsc_deposit_typed(eng.ctx, ptr(pos_d), int(pos_dt == 'f8'), ptr(w_d), int(w_dt == 'f8'), N, ptr(grid_d),
                                                int(grid_dt == 'f8'), shape3[0], shape3[1], shape3[2], shape3[2], float(box),
                                                float(offset), int(bool(wrap))))
    else:
        deposit_device(eng, pos_d, w_d, grid_d, shape3, shape3[2], box, offset, wrap=False)

    if user_supplied_grid:
        if grid_d is not densgrid:
            if is_torch_tensor(densgrid):
                densgrid += grid_d.to(densgrid.device, densgrid.dtype)
            else:
                densgrid += grid_d.cpu().numpy().astype(densgrid.dtype, copy=False)
        else:
            stream.synchronize()
        return None
    if on_device:
        return grid_d
    return grid_d.cpu().numpy()


def partition_parallel(pos, npartition, boxsize, weights=None, coord=0, nthread=-1, sort=False):
    """
    Partition particles into ``npartition`` stripes along ``coord`` (reference: tsc.py:259-384).

    Returns ``(partitioned, part_starts int64[npartition+1], wpart or None)``.  Inside a stripe the
    particles keep their input order (the reference's stable partition); with ``sort=True`` every stripe
    is sorted on ``coord``.
    """
    import torch

    assert pos.shape[1] == 3
    on_device = is_torch_tensor(pos) and pos.is_cuda
    eng = Engine.get(pos.device if on_device else None)
    eng.bind_stream()
    N = len(pos)
    if N >= 1 << 32:
        raise AbkError('partition_parallel: more than 2^32-1 particles in one call')
    in_dtype = pos.dtype
    pos_d = eng.to_device(pos, torch.float32)
    w_d = None if weights is None else eng.to_device(weights, torch.float32)
    out_pos = eng.empty((N, 3), torch.float32)
    out_w = None if w_d is None else eng.empty((N,), torch.float32)
    starts = eng.empty((npartition + 1,), torch.int64)
    nb = C.c_size_t()
    check(eng.lib.abk_partition_scratch_bytes(N, int(npartition), C.byref(nb)))
    scratch = eng.scratch('partition', nb.value)
    src = eng.empty((N,), torch.int32)
    check(eng.lib.abk_partition(eng.ctx, ptr(pos_d), ptr(w_d), N, int(npartition), float(boxsize), int(coord),
                                ptr(out_pos), ptr(out_w), ptr(starts), ptr(src), ptr(scratch), scratch.numel()))
    if N > 1 and not sort:
        # The reference's output is the STABLE partition (input order kept inside a stripe, tsc.py:338-376); the kernel's
        # atomic scatter is not.  Stripes are contiguous row ranges, so ordering the rows of every stripe by their source
        # index restores it: one sort of (stripe, source index) keys.
        stripe = torch.searchsorted(starts[1:].contiguous(), torch.arange(N, device=starts.device), right=True)
        order = torch.argsort((stripe << 32) | (src.to(torch.int64) & 0xffffffff))
        out_pos = out_pos[order].contiguous()
        if out_w is not None:
            out_w = out_w[order].contiguous()
    if sort and N > 0:
        # tsc.py:361-367, :378-382 sort each stripe on the coordinate.  The stripe key is monotone in
        # pos[:, coord], so one global stable sort on the coordinate leaves the stripes where they are
        # (same part_starts) and orders each of them.
        order = torch.argsort(out_pos[:, coord], stable=True)
        out_pos = out_pos[order].contiguous()
        if out_w is not None:
            out_w = out_w[order].contiguous()
    if on_device:
        return out_pos.to(in_dtype), starts, (None if out_w is None else out_w.to(weights.dtype))
    psort = out_pos.cpu().numpy().astype(in_dtype, copy=False)
    wsort = None if out_w is None else out_w.cpu().numpy().astype(weights.dtype, copy=False)
    return psort, starts.cpu().numpy(), wsort


_ = (ABK_MAX_SEGMENTS, SEGMENT_MAX)
