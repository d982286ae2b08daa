"""
Recipe for ``oracle/_ref/``: stage the UNMODIFIED reference modules of the hot path so that the reference's own Numba
implementation can be timed on the GPU box's host cores (``bench.py --impl reference``, ``cpu_baseline.kind = "reference"``).

    python oracle/make_ref.py            # also run by __graft_entry__.build() when /root/reference is present

The reference is a pure-Python package; "building" it for this path means placing the three modules the path imports
(``abacusnbody/analysis/{tsc,power_spectrum,cic}.py``) byte for byte under ``oracle/_ref/abacusnbody/analysis/``.
``oracle/_ref/`` is git-ignored (the reference sources never enter this repository's history) but not gpurun-ignored, so
it travels to the GPU box like the built ``.so`` files.  ``oracle/ref_shim.py`` imports the modules from there (or from
/root/reference when that exists) with the two shims described in its header; nothing is patched.
TEST / MEASUREMENT INFRASTRUCTURE ONLY: the product package never imports anything under oracle/.
"""

import hashlib
import json
import shutil
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = Path('/root/reference/abacusnbody/analysis')
DST = HERE / '_ref' / 'abacusnbody' / 'analysis'
FILES = ('tsc.py', 'power_spectrum.py', 'cic.py')


def stage(verbose=False):
    """Copy the modules if the reference tree is present; returns True when oracle/_ref is usable afterwards."""
    if SRC.is_dir():
        DST.mkdir(parents=True, exist_ok=True)
        manifest = {}
        for f in FILES:
            shutil.copyfile(SRC / f, DST / f)
            manifest[f] = hashlib.sha256((DST / f).read_bytes()).hexdigest()
        (HERE / '_ref' / 'MANIFEST.json').write_text(json.dumps({'source': str(SRC), 'sha256': manifest}, indent=1))
        if verbose:
            print(f'staged {len(FILES)} reference modules under {DST}')
    return all((DST / f).is_file() for f in FILES)


if __name__ == '__main__':
    sys.exit(0 if stage(verbose=True) else 1)
