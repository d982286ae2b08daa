"""
The sharded pipeline (abacusutils_b200/dist.py) end to end on the CPU: world_size 2 and 3 over gloo, every rank
running the real kernel sources under the emulator (tests/emu): routing with tile bucketing, the bucketed all-to-all,
slab deposit with ghost planes, 2-D FFT / transpose pack + all-to-all / 1-D FFT, pencil binning and the all-reduce.
The result must equal the reference's single-process table (tests/golden/reference_runs.npz).
"""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import cases

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, build_dir, name, out_dir, force_reroute, scatter):
    os.environ['OMP_NUM_THREADS'] = '1'
    os.environ['ABK_SCATTER'] = scatter
    os.environ['ABK_NO_P2P'] = '1'          # NVLink peer memory does not exist here: pack + all-to-all path
    for p in (ROOT, ROOT / 'tests', ROOT / 'tests' / 'golden', ROOT / 'tests' / 'emu'):
        sys.path.insert(0, str(p))
    import emu_engine

    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    mpatch = pytest.MonkeyPatch()
    try:
        emu_engine.install(mpatch, build_dir)
        from abacusutils_b200 import dist as abk_dist

        c = cases.POWER_CASES[name]
        pos, w, pos2, w2 = cases.power_inputs(c)
        sl = slice(rank, None, world)       # every rank holds an arbitrary share of the catalogue
        t = abk_dist.calc_power(pos[sl], c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                                nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'],
                                w=None if w is None else w[sl], pos2=None if pos2 is None else pos2[sl],
                                w2=None if w2 is None else w2[sl], poles=c['poles'], force_reroute=force_reroute)
        if rank == 0:
            np.savez(Path(out_dir) / 'table.npz', **{k: np.asarray(t[k]) for k in t.keys()})
    finally:
        mpatch.undo()
        dist.destroy_process_group()


@pytest.mark.parametrize('world,name,force_reroute,scatter', [(2, 'n32_ci', False, '1'), (3, 'n48_log', False, '1'),
                                                             (2, 'n32_cross_ci', True, '1'), (2, 'n32_ci', False, '2')])
def test_dist_calc_power_on_emulator(tmp_path, emu_build_dir, golden, world, name, force_reroute, scatter):
    import build_emu
    from common import compare_power_tables

    build_emu.build(emu_build_dir)          # once, before the ranks start
    mp.spawn(_worker, args=(world, _free_port(), str(emu_build_dir), name, str(tmp_path), force_reroute, scatter), nprocs=world, join=True)
    got = dict(np.load(tmp_path / 'table.npz'))
    want = {k[len(f'power/{name}/'):]: golden[k] for k in golden.files if k.startswith(f'power/{name}/')}
    compare_power_tables(got, want)
