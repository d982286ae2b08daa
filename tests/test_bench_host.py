"""Host-side pieces of bench.py that the driver depends on (no GPU): the algorithmic byte counts behind the roofline, the
mode count the binning kernel must read, the parity block, and the reference arm's behaviour on ranks other than 0."""

import importlib.util
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location('bench_module', ROOT / 'bench.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_used_entries_matches_brute_force():
    b = _bench()
    for n, L in ((8, 10.0), (12, 7.0), (9, 3.0)):
        kmax = np.pi * n / L
        dk = 2 * np.pi / L
        lim = np.float32((kmax / dk) ** 2)
        i = np.fft.fftfreq(n, 1.0 / n).astype(np.int64)
        k = np.arange(n // 2 + 1, dtype=np.int64)
        k2 = i[:, None, None] ** 2 + i[None, :, None] ** 2 + k[None, None, :] ** 2
        assert b.used_entries(n, L, kmax) == int((k2 < lim).sum())


def test_algorithmic_bytes_follow_survey_8d():
    b = _bench()
    cfg = dict(N=1000, nmesh=16)
    assert b.algorithmic_bytes('tsc_tile_deposit', cfg, 2, 0) == 1000 * 12 + 4 * 16**3      # 12 N + 4 n^3
    assert b.algorithmic_bytes('tsc_bucket_hist', cfg, 2, 0) == 1000 / 2 * 12
    assert b.algorithmic_bytes('tsc_bucket_scatter', cfg, 2, 0) == 1000 / 2 * 28
    assert b.algorithmic_bytes('power_bin', cfg, 2, 77) == 77 * 16
    assert b.algorithmic_bytes('scan', cfg, 2, 0) is None


def test_parity_block_flags_differences():
    b = _bench()
    rng = np.random.default_rng(0)
    Nk, Nmu = 5, 3
    want = dict(N_mode=rng.integers(1, 50, (Nk, Nmu)), power=rng.random((Nk, Nmu)) + 1.0,
                poles=rng.standard_normal((Nk, 3)), N_mode_poles=rng.integers(1, 100, Nk))      # table layout: one row per k bin
    want['poles'][:, 0] = np.abs(want['poles'][:, 0]) + 1.0
    got = {k: np.array(v, copy=True) for k, v in want.items()}
    ok = b.parity_block(got, want)
    assert ok['n_mode_exact'] and ok['max_rel_power'] == 0.0
    got['power'][2, 1] *= 1.0 + 3e-4
    got['N_mode'][0, 0] += 1
    bad = b.parity_block(got, want)
    assert not bad['n_mode_exact'] and bad['max_rel_power'] > 2e-4


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ''
