"""
Build libabk_emu.so: the UNMODIFIED kernel sources of libabk (everything but the cuFFT wrapper) compiled for the CPU against
tests/emu/include/cuda_runtime.h, which runs every CUDA thread as an OS thread (TEST INFRASTRUCTURE ONLY).

The only source transformation is syntactic: `kernel<<<grid, block, smem, stream>>>(args)` becomes
`emu::launch(grid, block, smem, stream)(kernel)(args)` and `extern __shared__ T name[];` becomes a pointer to the
block's dynamic shared memory.  Lets the `not gpu` suite execute the real launch plumbing (block scans, ballots,
barriers, header look-ups) of kernels, bit for bit, in a container without a GPU.
"""

import os
import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / 'abacusutils_b200' / 'csrc'
SOURCES = ['abk_ctx.cu', 'abk_ingest.cu', 'abk_kfields.cu', 'abk_kspace.cu', 'abk_tsc.cu', 'abk_f64.cu']

LAUNCH = re.compile(r'([A-Za-z_]\w*(?:<[^<>;()]*>)?)\s*<<<(.*?)>>>\s*\(', re.S)
PTX_HINT = re.compile(r'asm\s+volatile\s*\(\s*"prefetch[^;]*;[^;]*;')   # the PTX string itself ends in ';'
DYN_SMEM = re.compile(r'extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w:<> ]+?)\s+(\w+)\s*\[\s*\]\s*;')


def translate(text):
    text = LAUNCH.sub(lambda m: f'emu::launch({m.group(2)})({m.group(1)})(', text)
    text = PTX_HINT.sub('(void)0;', text)   # cache hints have no functional effect
    text = DYN_SMEM.sub(lambda m: f'{m.group(1)} *{m.group(2)} = ({m.group(1)} *)emu::dyn_smem();', text)
    return text


def build(outdir):
    outdir = Path(outdir)
    outdir.mkdir(parents=True, exist_ok=True)
    so = outdir / 'libabk_emu.so'
    deps = [CSRC / s for s in SOURCES] + list(CSRC.glob('*.cuh')) + [HERE / 'include' / 'cuda_runtime.h', ROOT / 'include' / 'abk.h']
    if so.exists() and all(so.stat().st_mtime >= d.stat().st_mtime for d in deps):
        return so
    cpps = []
    for s in SOURCES:
        dst = outdir / (Path(s).stem + '_emu.cpp')
        dst.write_text(translate((CSRC / s).read_text()))
        cpps.append(str(dst))
    gxx = '/usr/bin/g++' if Path('/usr/bin/g++').exists() else 'g++'
    # ABK_EMU_ASAN=1: AddressSanitizer build (out-of-bounds global / shared accesses of every kernel); run the tests with
    #   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
    asan = ['-fsanitize=address', '-fno-omit-frame-pointer'] if os.environ.get('ABK_EMU_ASAN') == '1' else []
    cmd = [gxx, '-std=c++20', '-O1', '-g', *asan, '-pthread', '-fPIC', '-shared', '-ffp-contract=off', '-Wno-attributes', '-D__CUDACC__',
           f'-I{HERE / "include"}', f'-I{CSRC}', f'-I{ROOT / "include"}', '-o', str(so), *cpps]
    subprocess.run(cmd, check=True)
    return so


if __name__ == '__main__':
    import sys

    print(build(sys.argv[1] if len(sys.argv) > 1 else '/tmp/abk_emu'))
