"""Test-side writer of Abacus-style ASDF files (YAML tree + 'blsc' blocks of blosc-1 frames), used to build
fixtures for the container reader where the reference tree (and its real files) is absent, e.g. on the GPU box."""

import struct

import numpy as np
import pyarrow as pa
import yaml

_NAMES = {'i1': 'int8', 'u1': 'uint8', 'i4': 'int32', 'u8': 'uint64', 'f4': 'float32'}


def blosc1_frame(raw, typesize, shuffle='shuffle', blocksize=1 << 16, store=False):
    """One blosc-1 frame with zstd streams, 'don't split' layout (flag 0x10)."""
    nbytes = len(raw)
    flags = 0x10 | (4 << 5)
    if shuffle == 'shuffle':
        flags |= 0x1
    elif shuffle == 'bitshuffle':
        flags |= 0x4
    if store:
        body = bytes(raw)
        return struct.pack('<BBBBIII', 2, 1, 0x2, typesize, nbytes, nbytes, 16 + len(body)) + body
    blocksize -= blocksize % typesize
    nblocks = max(1, -(-nbytes // blocksize))
    zstd = pa.Codec('zstd')
    parts = []
    for b in range(nblocks):
        blk = raw[b * blocksize:(b + 1) * blocksize]
        n = len(blk) // typesize
        if shuffle == 'shuffle' and typesize > 1:
            body = np.frombuffer(blk, np.uint8, n * typesize).reshape(n, typesize).T
            blk = np.ascontiguousarray(body).tobytes() + blk[n * typesize:]
        elif shuffle == 'bitshuffle' and n and n % 8 == 0:
            el = np.frombuffer(blk, np.uint8, n * typesize).reshape(n, typesize, 1)
            bits = np.unpackbits(el, axis=2, bitorder='little')                      # [n][typesize][8]
            rows = np.packbits(bits.transpose(1, 2, 0), axis=2, bitorder='little')   # [typesize][8][n/8]
            blk = rows.tobytes() + blk[n * typesize:]
        c = zstd.compress(blk).to_pybytes()
        if len(c) >= len(blk):
            c = blk   # incompressible stream: stored verbatim, recognised by csize == block size
        parts.append(struct.pack('<i', len(c)) + c)
    bstarts, p = [], 16 + 4 * nblocks
    for part in parts:
        bstarts.append(p)
        p += len(part)
    header = struct.pack('<BBBBIII', 2, 1, flags, typesize, nbytes, blocksize, p)
    return header + struct.pack(f'<{nblocks}i', *bstarts) + b''.join(parts)


def write_asdf(path, arrays, header, compression='blsc', shuffle='shuffle', frame_bytes=1 << 18, pad=0):
    """arrays: {name: ndarray} -> data/<name>; header -> header/.  ``pad`` appends garbage bytes to each block's
    payload before compression (the Abacus slab files carry such a tail beyond the declared shape)."""
    tree = {'data': {}, 'header': header}
    blocks = []
    for i, (name, a) in enumerate(arrays.items()):
        a = np.ascontiguousarray(a)
        tree['data'][name] = {'__nd__': i, 'datatype': _NAMES[a.dtype.str[1:]], 'byteorder': 'little', 'shape': list(a.shape)}
        raw = a.tobytes() + bytes(range(7)) * (pad // 7)
        if compression == 'blsc':
            step = frame_bytes - frame_bytes % a.itemsize
            body = b''
            for o in range(0, max(len(raw), 1), step):
                fr = blosc1_frame(raw[o:o + step], a.itemsize if a.ndim == 1 or name != 'pack9' else 9, shuffle)
                body += struct.pack('!I', len(fr)) + fr
            comp = b'blsc'
        else:
            body, comp = raw, b'\0\0\0\0'
        blocks.append((comp, body, len(raw)))
    text = yaml.safe_dump(tree, default_flow_style=False)
    # turn the placeholder mappings into ndarray-tagged nodes
    out = []
    for line in text.splitlines():
        if line.strip().startswith('__nd__:'):
            out.append(line.replace('__nd__:', 'source:'))
        else:
            out.append(line)
    text = '\n'.join(out)
    for name in arrays:
        text = text.replace(f'  {name}:\n', f'  {name}: !core/ndarray-1.0.0\n', 1)
    with open(path, 'wb') as f:
        f.write(b'#ASDF 1.0.0\n#ASDF_STANDARD 1.5.0\n%YAML 1.1\n%TAG ! tag:stsci.edu:asdf/\n--- !core/asdf-1.1.0\n')
        f.write(text.encode() + b'\n...\n')
        for comp, body, dsize in blocks:
            f.write(b'\xd3BLK' + struct.pack('>H', 48) + struct.pack('>I4sQQQ', 0, comp, len(body), len(body), dsize) + b'\0' * 16)
            f.write(body)
