"""
``read_asdf`` with the particle decoding on the GPU (reference: abacusnbody/data/read_abacus.py:34-187).

The container is parsed and decompressed on the host (:mod:`.asdf_container`); the packed column (``rvint``,
``pack9``, ``packedpid`` / ``pid``) is copied to the device still packed and decoded there.  With ``device=True``
the columns stay on the GPU as torch tensors, ready for ``abacusutils_b200.analysis.power_spectrum.calc_power``.
"""

from __future__ import annotations

import warnings
from os.path import basename

import numpy as np

from .asdf_container import ArrayRef, AsdfFile
from .bitpacked import unpack_pids, unpack_rvint
from .pack9 import unpack_pack9

__all__ = ['read_asdf']

ASDF_DATA_KEY = 'data'
ASDF_HEADER_KEY = 'header'


def _table(columns, meta):
    from ..analysis.power_spectrum import Table, _AstropyTable

    if _AstropyTable is not None and all(isinstance(v, np.ndarray) for v in columns.values()):
        return _AstropyTable(columns, meta=meta, copy=False)
    return Table(columns, meta=meta)


def read_asdf(fn, load=None, colname=None, dtype=np.float32, verbose=True, device=False, **kwargs):
    """Read an Abacus ASDF file into a table (read_abacus.py:34-187).

    ``load``: columns to produce, from ``'pos', 'vel', 'pid', 'lagr_pos', 'tagged', 'density', 'lagr_idx', 'aux'``
    (default: ``pos`` and ``vel`` for rvint / pack9 files, ``pid`` for PID files).  ``colname``: the packed column
    in the file (auto-detected).  ``dtype``: float type of the unpacked columns.  ``device=True`` (extension) keeps
    the unpacked columns on the GPU as torch tensors.  The table's ``meta`` is the file's header.
    """
    data_key = kwargs.pop('data_key', ASDF_DATA_KEY)
    header_key = kwargs.pop('header_key', ASDF_HEADER_KEY)

    af = AsdfFile(fn)
    tree_data = af.tree[data_key]
    if colname is None:
        _colnames = ['rvint', 'pack9', 'packedpid', 'pid']
        for cn in _colnames:
            if cn in tree_data:
                if colname is not None:
                    raise ValueError(f'More than one key of {_colnames} found in asdf file {fn}. Need to specify colname!')
                colname = cn
        if colname is None:
            raise ValueError(f'Could not find any of {_colnames} in asdf file {fn}. Need to specify colname!')

    load = _resolve_columns(colname, load, kwargs)
    header = dict(af.tree[header_key])
    ref = tree_data[colname]
    if not isinstance(ref, ArrayRef):
        raise ValueError(f'column {colname!r} of {fn} is not a block-backed array')
    data = af.read(ref)

    if header.get('OutputType', None) == 'LightCone' and header.get('SimSet') == 'AbacusSummit':
        header['SubsampleFraction'] = header['ParticleSubsampleA'] + header['ParticleSubsampleB']
        if verbose:
            print(f'Loading "{basename(str(fn))}", which contains the A and B subsamples '
                  f'({int(header["SubsampleFraction"] * 100):d}% total)')

    def on_device(a):
        import torch

        return torch.from_numpy(np.array(a, order='C')).cuda()      # a copy: the block buffer is read-only

    columns = {}
    want_pos, want_vel = 'pos' in load, 'vel' in load
    if colname in ('rvint', 'pack9'):
        if want_pos or want_vel:
            src = on_device(data.view(np.uint8) if colname == 'pack9' else data) if device else data
            if colname == 'rvint':
                pos, vel = unpack_rvint(src, header['BoxSize'], float_dtype=dtype, posout=None if want_pos else False,
                                        velout=None if want_vel else False)
            else:
                pos, vel = unpack_pack9(src, header['BoxSize'], header['VelZSpace_to_kms'], float_dtype=dtype,
                                        posout=None if want_pos else False, velout=None if want_vel else False)
            if want_pos:
                columns['pos'] = pos
            if want_vel:
                columns['vel'] = vel
    elif 'pid' in colname:
        if 'aux' in load:
            columns['aux'] = on_device(data.view(np.int64)) if device else data
        ppd = kwargs.get('ppd', int(round(header['ppd'])))
        fields = {k: (k in load) for k in ('pid', 'lagr_pos', 'tagged', 'density', 'lagr_idx')}
        if any(fields.values()):
            src = on_device(data.view(np.int64)) if device else data
            columns.update(unpack_pids(src, box=header['BoxSize'], ppd=ppd, float_dtype=dtype, **fields))
    else:
        raise ValueError(f'unknown packed column {colname!r}')
    return _table(columns, header)


def _resolve_columns(colname, load, kwargs):
    """Which columns to produce; honours the deprecated load_pos / load_vel keywords (read_abacus.py:190-212)."""
    load_pos = kwargs.pop('load_pos', None)
    load_vel = kwargs.pop('load_vel', None)
    if load_pos is not None or load_vel is not None:
        if load is None:
            warnings.warn('`load_pos` and `load_vel` are deprecated; use `load=("pos","vel")` instead.', FutureWarning)
            load = []
            if load_pos or (load_pos is None and load_vel is False):
                load += ['pos']
            if load_vel or (load_vel is None and load_pos is False):
                load += ['vel']
        else:
            warnings.warn('`load` and deprecated `load_pos` or `load_vel` specified. Ignoring deprecated parameters.')
    if load is None:
        load = []
        if colname in ('pack9', 'rvint'):
            load += ['pos', 'vel']
        if 'pid' in colname:
            load += ['pid']
    return tuple(load)
