import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / 'tests' / 'golden', ROOT / 'tests' / 'emu'):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

# keep the live reference (when present) on its race-free stripe branch; must precede numba import
os.environ.setdefault('NUMBA_NUM_THREADS', '2')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')
    config.addinivalue_line('markers', 'reference: needs the live reference tree under /root/reference')


@pytest.fixture(scope='session')
def golden():
    import numpy as np

    return np.load(ROOT / 'tests' / 'golden' / 'reference_runs.npz')


@pytest.fixture(scope='session')
def oracle():
    from oracle import abk_oracle

    abk_oracle.build()
    return abk_oracle


@pytest.fixture(scope='session')
def emu_build_dir(tmp_path_factory):
    """Where tests/emu/build_emu.py puts libabk_emu.so (built once per session)."""
    return tmp_path_factory.mktemp('abk_emu')
