"""pytest plugin (opt-in: `-p pytest_emu_plugin`): run the `-m gpu` tests on the CPU emulator instead of a device.
Tests that need real CUDA tensors fail/skip; everything that goes NumPy-in / NumPy-out runs the real kernel sources.
Usage:  python -m pytest tests -m gpu -p pytest_emu_plugin --timeout 300"""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent))


@pytest.fixture(autouse=True)
def _abk_emulator(monkeypatch, emu_build_dir):
    import emu_engine

    emu_engine.install(monkeypatch, emu_build_dir)
    yield
