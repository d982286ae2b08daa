"""
Minimal reader for the ASDF containers Abacus writes (host side of the particle ingest).

The reference opens its files with the ``asdf`` package plus its own 'blsc' compression extension
(abacusnbody/data/asdf.py:24-170, read_abacus.py:76-110); neither ``asdf`` nor ``blosc`` is a dependency here.
What is read:

  * the YAML tree (``yaml`` SafeLoader; ASDF tags are dropped, ``core/ndarray`` nodes become :class:`ArrayRef`),
  * binary blocks (ASDF standard 1.5: magic ``\\xd3BLK``, u16 header size, flags / compression / allocated /
    used / data sizes / checksum),
  * block compression ``blsc`` -- the reference's framing ``[u32 big-endian length][blosc-1 frame]...`` --
    plus ``zlib``, ``bzp2`` and uncompressed blocks,
  * blosc-1 frames: 16-byte header, ``bstarts``, per block one stream (flag 0x10) or ``typesize`` streams of
    ``[i32 size][payload]``; codecs zstd / lz4 / zlib (via pyarrow or the standard library), memcpy'd frames;
    byte-shuffle (0x1) and bit-shuffle (0x4) filters undone with NumPy.  BloscLZ and Snappy are not supported.

Decompression is host work by nature (entropy coding); the particle decoding that follows runs on the GPU
(``bitpacked.unpack_rvint`` / ``pack9.unpack_pack9``).
"""

from __future__ import annotations

import bz2
import struct
import zlib

import numpy as np

__all__ = ['AsdfFile', 'ArrayRef']

_BLOCK_MAGIC = b'\xd3BLK'
_DTYPES = {'int8': 'i1', 'uint8': 'u1', 'int16': 'i2', 'uint16': 'u2', 'int32': 'i4', 'uint32': 'u4', 'int64': 'i8',
           'uint64': 'u8', 'float32': 'f4', 'float64': 'f8', 'bool8': 'b1'}


class ArrayRef:
    """An ``!core/ndarray`` node: which block holds the bytes and how to view them."""

    def __init__(self, node):
        if 'source' not in node or not isinstance(node['source'], int):
            raise ValueError('only block-backed ndarrays are supported')
        if not isinstance(node.get('datatype'), str) or node['datatype'] not in _DTYPES:
            raise ValueError(f'unsupported ASDF datatype {node.get("datatype")!r}')
        self.source = node['source']
        order = '>' if node.get('byteorder', 'little') == 'big' else '<'
        self.dtype = np.dtype(order + _DTYPES[node['datatype']])
        self.shape = tuple(int(s) for s in node['shape'])
        self.offset = int(node.get('offset', 0))
        if 'strides' in node:
            raise ValueError('strided ndarrays are not supported')

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return f'ArrayRef(block={self.source}, dtype={self.dtype}, shape={self.shape})'


def _zstd(buf, size):
    import pyarrow as pa

    return pa.Codec('zstd').decompress(buf, decompressed_size=size).to_pybytes()


def _lz4(buf, size):
    import pyarrow as pa

    return pa.Codec('lz4_raw').decompress(buf, decompressed_size=size).to_pybytes()


_CODECS = {1: _lz4, 3: lambda buf, size: zlib.decompress(buf), 4: _zstd}


def _unshuffle(buf, typesize):
    n = len(buf) // typesize
    body = np.frombuffer(buf, dtype=np.uint8, count=n * typesize).reshape(typesize, n).T
    return np.ascontiguousarray(body).tobytes() + bytes(buf[n * typesize:])


def _bitunshuffle(buf, typesize):
    # c-blosc 1.x (what python-blosc 1.x -- the reference's dependency -- wraps; shuffle.c, blosc_internal_bitshuffle) only
    # bit-shuffles a block whose element count is a multiple of 8 and stores any other block as is.  (c-blosc2 differs: it
    # shuffles the first n - n % 8 elements and copies the rest; blosc-2 frames are not what Abacus files contain and are
    # rejected by their header version before they get here.)
    # layout [byte in element][bit][n/8], element 8j in the least significant bit of byte j
    n = len(buf) // typesize
    if n == 0 or n % 8:
        return bytes(buf)
    rows = np.frombuffer(buf, dtype=np.uint8, count=n * typesize).reshape(typesize, 8, n // 8)
    bits = np.unpackbits(rows, axis=2, bitorder='little')
    elems = np.packbits(bits.transpose(2, 0, 1), axis=2, bitorder='little')
    return elems.reshape(-1).tobytes() + bytes(buf[n * typesize:])


def blosc1_decompress(frame):
    """Decompress one blosc-1 frame (the format python-blosc 1.x / c-blosc 1.x writes)."""
    frame = memoryview(frame)
    if len(frame) < 16:
        raise ValueError('truncated blosc frame')
    _version, _versionlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack('<BBBBIII', frame[:16])
    if _version > 2:      # c-blosc 1.x writes format version 2; blosc-2 frames (>= 3) differ (header, bitshuffle leftovers)
        raise NotImplementedError(f'blosc format version {_version}: only blosc-1 frames (version 2) are supported')
    if cbytes != len(frame):
        raise ValueError(f'blosc frame length {len(frame)} does not match its header ({cbytes})')
    if nbytes == 0:
        return b''
    if flags & 0x2:  # stored uncompressed
        return bytes(frame[16:16 + nbytes])
    codec = _CODECS.get(flags >> 5)
    if codec is None:
        raise NotImplementedError(f'blosc codec {flags >> 5} is not supported (zstd, lz4 and zlib are)')
    shuffle, bitshuffle, dont_split = bool(flags & 0x1), bool(flags & 0x4), bool(flags & 0x10)
    nblocks = (nbytes + blocksize - 1) // blocksize
    bstarts = struct.unpack(f'<{nblocks}i', frame[16:16 + 4 * nblocks])
    out = []
    for b in range(nblocks):
        bsize = min(blocksize, nbytes - b * blocksize)
        leftover = bsize != blocksize
        split = (not dont_split) and typesize <= 16 and bsize // typesize >= 128 and not leftover
        nstreams = typesize if split else 1
        neblock = bsize // nstreams
        p = bstarts[b]
        parts = []
        for _ in range(nstreams):
            (csize,) = struct.unpack('<i', frame[p:p + 4])
            p += 4
            parts.append(bytes(frame[p:p + csize]) if csize == neblock else codec(frame[p:p + csize], neblock))
            p += csize
        blk = b''.join(parts)
        if shuffle and typesize > 1:
            blk = _unshuffle(blk, typesize)
        elif bitshuffle:
            blk = _bitunshuffle(blk, typesize)
        out.append(blk)
    res = b''.join(out)
    if len(res) != nbytes:
        raise ValueError('blosc frame decompressed to the wrong size')
    return res


def _decompress_block(comp, body, data_size):
    if comp == b'\0\0\0\0':
        raw = bytes(body)
    elif comp == b'blsc':  # abacusnbody/data/asdf.py:72-84: [u32 BE length][blosc frame] repeated
        out, p = [], 0
        while p < len(body):
            (n,) = struct.unpack('!I', body[p:p + 4])
            out.append(blosc1_decompress(body[p + 4:p + 4 + n]))
            p += 4 + n
        raw = b''.join(out)
    elif comp == b'zlib':
        raw = zlib.decompress(body)
    elif comp == b'bzp2':
        raw = bz2.decompress(body)
    else:
        raise NotImplementedError(f'unsupported ASDF block compression {comp!r}')
    if len(raw) != data_size:
        raise ValueError(f'ASDF block decompressed to {len(raw)} bytes, header says {data_size}')
    return raw


def _load_tree(text):
    import yaml

    base = getattr(yaml, 'CSafeLoader', yaml.SafeLoader)

    class Loader(base):
        pass

    def generic(loader, suffix, node):
        if isinstance(node, yaml.MappingNode):
            value = loader.construct_mapping(node, deep=True)
            if 'core/ndarray' in suffix:
                return ArrayRef(value)
            return value
        if isinstance(node, yaml.SequenceNode):
            return loader.construct_sequence(node, deep=True)
        return loader.construct_scalar(node)

    Loader.add_multi_constructor('!', generic)
    Loader.add_multi_constructor('tag:', generic)
    return yaml.load(text, Loader=Loader)


class AsdfFile:
    """``AsdfFile(path).tree`` is the YAML tree; ``read(ref)`` returns the NumPy array of an :class:`ArrayRef`."""

    def __init__(self, path):
        self.path = str(path)
        with open(self.path, 'rb') as f:
            data = f.read()
        if not data.startswith(b'#ASDF'):
            raise ValueError(f'{self.path} is not an ASDF file')
        end = data.find(b'\n...')
        if end < 0:
            raise ValueError('ASDF file without a YAML document end marker')
        self.tree = _load_tree(data[:end].decode('utf-8'))
        self._data = data
        self._blocks = []   # (offset of the payload, compression, used, data_size)
        pos = end
        while True:
            i = data.find(_BLOCK_MAGIC, pos)
            if i < 0:
                break
            (hsize,) = struct.unpack('>H', data[i + 4:i + 6])
            flags, comp, alloc, used, dsize = struct.unpack('>I4sQQQ', data[i + 6:i + 6 + 32])
            if flags & 1:
                raise NotImplementedError('streamed ASDF blocks are not supported')
            self._blocks.append((i + 6 + hsize, comp, used, dsize))
            pos = i + 6 + hsize + alloc
        self._cache = {}

    def block(self, index):
        if index not in self._cache:
            off, comp, used, dsize = self._blocks[index]
            self._cache[index] = _decompress_block(comp, memoryview(self._data)[off:off + used], dsize)
        return self._cache[index]

    def read(self, ref):
        raw = self.block(ref.source)
        count = int(np.prod(ref.shape, dtype=np.int64))
        a = np.frombuffer(raw, dtype=ref.dtype, count=count, offset=ref.offset).reshape(ref.shape)
        return a.astype(ref.dtype.newbyteorder('='), copy=False)
