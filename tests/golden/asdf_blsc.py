"""
Minimal reader for the reference's blosc-compressed ASDF golden files (tests/ref_tsc/*.asdf).

Neither ``asdf`` nor ``blosc`` is installed here, so this decodes the three layers by hand:
  1. ASDF binary block: magic ``\\xd3BLK``, u16 header size, then flags/compression/sizes/checksum
     (ASDF standard 1.5, "Block header").
  2. the reference's own framing of a 'blsc' block: a sequence of ``[u32 big-endian length][blosc frame]``
     (/root/reference/abacusnbody/data/asdf.py:72-84).
  3. blosc-1 frame (byte-shuffle flag 0x1 or bit-shuffle flag 0x4, undone per block): 16-byte header (version, versionlz, flags, typesize, nbytes, blocksize, cbytes),
     ``bstarts`` table, then per block either one stream (flag 0x10 "don't split") or ``typesize``
     streams, each ``[i32 csize][payload]``; byte-shuffle (flag 0x1) undone per block.
     Only the zstd codec (flags >> 5 == 4; what the golden files use) and memcpy'd frames (flag 0x2)
     are supported; zstd comes from pyarrow.

Used only by make_golden.py (build container).
"""

import struct

import numpy as np
import pyarrow as pa


def _unshuffle(buf, typesize):
    n = len(buf) // typesize
    main = np.frombuffer(buf[: n * typesize], dtype=np.uint8).reshape(typesize, n).T.reshape(-1)
    return main.tobytes() + bytes(buf[n * typesize:])


def _bitunshuffle(buf, typesize):
    """Undo blosc-1's bitshuffle.  c-blosc only bit-shuffles a block whose element count is a multiple of 8
    (otherwise the block is stored unshuffled, shuffle.c `blosc_internal_bitshuffle`); a shuffled block is laid
    out as [byte-in-element][bit][n/8] rows (element 8j in the least significant bit of byte j), and bytes
    beyond the last whole element are copied verbatim."""
    n = len(buf) // typesize
    if n == 0 or n % 8:
        return bytes(buf)
    rows = np.frombuffer(buf[: n * typesize], dtype=np.uint8).reshape(typesize, 8, n // 8)
    bits = np.unpackbits(rows, axis=2, bitorder='little')          # [typesize][8][n] : bit k of byte b of element e
    elems = np.packbits(bits.transpose(2, 0, 1), axis=2, bitorder='little')  # [n][typesize][1]
    return elems.reshape(-1).tobytes() + bytes(buf[n * typesize:])


def blosc1_decompress(frame):
    version, versionlz, flags, typesize, nbytes, blocksize, cbytes = struct.unpack('<BBBBIII', frame[:16])
    assert cbytes == len(frame), (cbytes, len(frame))
    if flags & 0x2:  # memcpy'd
        return bytes(frame[16:16 + nbytes])
    codec = flags >> 5
    assert codec == 4, f'only zstd blosc frames supported, got codec {codec}'
    zstd = pa.Codec('zstd')
    doshuffle = bool(flags & 0x1)
    dobitshuffle = bool(flags & 0x4)
    dont_split = bool(flags & 0x10)
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
            if csize == neblock:
                parts.append(bytes(frame[p:p + csize]))
            else:
                parts.append(zstd.decompress(frame[p:p + csize], decompressed_size=neblock).to_pybytes())
            p += csize
        blk = b''.join(parts)
        if doshuffle and typesize > 1:
            blk = _unshuffle(blk, typesize)
        elif dobitshuffle:
            blk = _bitunshuffle(blk, typesize)
        out.append(blk)
    res = b''.join(out)
    assert len(res) == nbytes
    return res


def read_asdf_blocks(path):
    """Yield the decompressed bytes of every binary block in an ASDF file."""
    data = open(path, 'rb').read()
    pos = 0
    blocks = []
    while True:
        i = data.find(b'\xd3BLK', pos)
        if i < 0:
            break
        (hsize,) = struct.unpack('>H', data[i + 4:i + 6])
        flags, comp, alloc, used, dsize = struct.unpack('>I4sQQQ', data[i + 6:i + 6 + 32])
        body = data[i + 6 + hsize:i + 6 + hsize + used]
        if comp == b'blsc':
            out, p = [], 0
            while p < len(body):
                (n,) = struct.unpack('!I', body[p:p + 4])
                out.append(blosc1_decompress(body[p + 4:p + 4 + n]))
                p += 4 + n
            raw = b''.join(out)
        elif comp == b'\0\0\0\0':
            raw = bytes(body)
        else:
            raise ValueError(f'unsupported ASDF block compression {comp!r}')
        assert len(raw) == dsize, (len(raw), dsize)
        blocks.append(raw)
        pos = i + 6 + hsize + alloc
    return blocks


def read_single_array(path, dtype, shape):
    (raw,) = read_asdf_blocks(path)
    return np.frombuffer(raw, dtype=dtype).reshape(shape).copy()
