"""
GPU tests of the x-slab sharded path (abacusutils_b200/dist.py): slab-mode bucketing/deposit with
ghost planes, 2-D + 1-D FFT with the pack/transpose, pencil-layout binning, all-reduce.
  * world_size 1 (always runs): the sharded pipeline on one rank (ghost planes fold onto itself) must
    reproduce the reference's calc_power goldens;
  * world_size 2 and 4 through torchrun when the box has that many GPUs.
"""

import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import cases
from common import compare_power_tables

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.fixture(scope='module')
def single_rank_group():
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        torch.cuda.set_device(0)
        dist.init_process_group('nccl', init_method=f'tcp://127.0.0.1:{_free_port()}', rank=0, world_size=1,
                                device_id=torch.device('cuda', 0))
    yield
    dist.destroy_process_group()


@pytest.mark.parametrize('reroute', [False, True], ids=['bucketed-exchange', 'route+rebucket'])
@pytest.mark.parametrize('name', ['n32_ci', 'n32_c', 'n32_raw', 'n32_cross_ci', 'n48_log', 'n40_defaults', 'cfg1_small'])
def test_sharded_world1_vs_reference(single_rank_group, golden, name, reroute):
    from abacusutils_b200 import dist as abk_dist

    c = cases.POWER_CASES[name]
    pos, w, pos2, w2 = cases.power_inputs(c)
    t = abk_dist.calc_power(pos, c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'],
                            paste='TSC', nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'],
                            w=w, pos2=pos2, w2=w2, poles=c['poles'], force_reroute=reroute)
    pre = f'power/{name}/'
    want = {k[len(pre):]: golden[k] for k in golden.files if k.startswith(pre)}
    assert set(want) == set(t.keys())
    compare_power_tables(t, want)


@pytest.mark.parametrize('reroute', ['0', '1'], ids=['bucketed-exchange', 'route+rebucket'])
@pytest.mark.parametrize('world', [2, 4])
@pytest.mark.parametrize('name', ['n32_ci', 'n32_cross_ci', 'n40_defaults'])
def test_sharded_multi_rank_vs_reference(golden, tmp_path, world, name, reroute):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f'needs {world} GPUs')
    out = tmp_path / 'res.npz'
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={world}',
           '--master-addr', '127.0.0.1', '--master-port', str(_free_port()), str(ROOT / 'tests' / 'dist_worker.py'),
           str(out), name, reroute]
    env = dict(os.environ)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    got = dict(np.load(out))
    pre = f'power/{name}/'
    want = {k[len(pre):]: golden[k] for k in golden.files if k.startswith(pre)}
    compare_power_tables(got, want)


@pytest.mark.parametrize('nranks,n', [(1, 24), (2, 24), (3, 20), (8, 64)])
def test_transpose_scatter_p2p_virtual_ranks(nranks, n):
    """The fused pack + peer-store transpose kernel (abk_transpose_scatter_p2p, the default at N > 1) on ONE GPU: every
    virtual rank owns an x-slab and a pencil buffer on this device, so the 'peer' stores are local -- the indexing (uneven
    y ranges, x offsets, row strides) is exactly what runs over NVLink."""
    import ctypes as C

    import torch

    from abacusutils_b200._lib import Engine, check

    eng = Engine.get(0)
    eng.bind_stream()
    g = torch.Generator(device='cuda')
    g.manual_seed(n * 10 + nranks)
    nzc = n // 2 + 1
    full = torch.view_as_complex(torch.randn((n, n, nzc, 2), device='cuda', generator=g))
    xs = [r * n // nranks for r in range(nranks)] + [n]
    rng = np.random.default_rng(nranks)
    js = [0] + sorted(rng.choice(np.arange(1, n), size=nranks - 1, replace=False).tolist()) + [n]
    nyl_max = max(js[r + 1] - js[r] for r in range(nranks))
    pencils = [torch.full((n * nyl_max * nzc,), float('nan'), dtype=torch.complex64, device='cuda') for _ in range(nranks)]
    peers = (C.c_void_p * nranks)(*[p.data_ptr() for p in pencils])
    jsp = (C.c_int64 * (nranks + 1))(*js)
    for r in range(nranks):
        slab = full[xs[r]:xs[r + 1]].contiguous()
        check(eng.lib.abk_transpose_scatter_p2p(eng.ctx, C.c_void_p(slab.data_ptr()), peers, xs[r + 1] - xs[r], n, nzc, nranks, jsp,
                                                xs[r]))
    torch.cuda.synchronize()
    for r in range(nranks):
        nyl = js[r + 1] - js[r]
        got = pencils[r][: n * nyl * nzc].view(n, nyl, nzc)
        assert torch.equal(got, full[:, js[r]:js[r + 1], :]), f'rank {r}'
