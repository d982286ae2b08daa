"""
Build recipe for libabk.so (nvcc, sm_100a only, in-tree so the .so travels with the repo snapshot).

    python -m abacusutils_b200._build          # rebuild if sources are newer than the library
"""

import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / 'csrc'
LIB = PKG / 'libabk.so'
SOURCES = ['abk_ctx.cu', 'abk_tsc.cu', 'abk_fft.cu', 'abk_kspace.cu', 'abk_kfields.cu', 'abk_ingest.cu', 'abk_f64.cu']
CUDA_HOME = os.environ.get('CUDA_HOME', '/usr/local/cuda')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '--shared', '--extended-lambda',
    # IEEE division/sqrt and no FMA contraction surprises in the parity-critical paths:
    # the kernels use explicit intrinsics where rounding matters; fast-math stays OFF.
    '-Xptxas', '-v',
]


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [CSRC / 'abk_common.cuh', CSRC / 'abk_ingest.cuh', ROOT / 'include' / 'abk.h']
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.path.join(CUDA_HOME, 'bin', 'nvcc')
    cmd = [nvcc, *NVCC_FLAGS, '-ccbin', '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++',
           '-I', str(ROOT / 'include'), '-I', str(CSRC),
           *[str(CSRC / s) for s in SOURCES], '-o', str(LIB),
           '-L', os.path.join(CUDA_HOME, 'lib64'), '-ldl']      # cuFFT is opened at run time (csrc/abk_fft.cu), not linked
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libabk.so')
    (PKG / 'build_ptxas.log').write_text(res.stdout + res.stderr)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True)
    print(LIB)
