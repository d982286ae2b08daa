"""
Host-side logic of the sharded path (abacusutils_b200/dist.py) on CPU with the gloo backend,
world_size 2 and 3: slab plan, particle routing exchange, ghost-plane ring exchange, and the
slab->pencil transpose.  The compute kernels are replaced by NumPy / the CPU oracle here (they need a
GPU); what is checked is that the exchanges put the right data in the right place, by comparing
against the un-sharded result.
"""

import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ['OMP_NUM_THREADS'] = '1'
    dist.init_process_group('gloo', init_method=f'tcp://127.0.0.1:{port}', rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


def test_slab_plan():
    from abacusutils_b200.dist import SlabPlan

    p = SlabPlan(10, 3)
    assert p.xsplit == [0, 3, 6, 10] and p.nzc == 6
    assert [p.nxl(r) for r in range(3)] == [3, 3, 4]
    assert p.owner_of_plane(0) == 0 and p.owner_of_plane(5) == 1 and p.owner_of_plane(9) == 2
    assert p.owner_of_plane(10) == 0 and p.owner_of_plane(-1) == 2
    send, recv = p.transpose_splits(2)
    assert send == [4 * 3 * 6, 4 * 3 * 6, 4 * 4 * 6] and recv == [3 * 4 * 6, 3 * 4 * 6, 4 * 4 * 6]
    with pytest.raises(ValueError):
        SlabPlan(4, 3)


def _owner(pos, n, box, plan):
    cell = np.rint(pos[:, 0] * np.float32(n / box)).astype(np.int64) % n
    return np.searchsorted(np.asarray(plan.xsplit), cell, side='right') - 1


def _routing(rank, world, n, box):
    from abacusutils_b200.dist import SlabPlan, exchange_counts, exchange_rows

    plan = SlabPlan(n, world)
    rng = np.random.default_rng(100 + rank)
    N = 5000 + 300 * rank
    pos = rng.random((N, 3), dtype='f4') * np.float32(box)
    w = rng.random(N, dtype='f4')
    own = _owner(pos, n, box, plan)
    order = np.argsort(own, kind='stable')
    rows = torch.from_numpy(np.c_[pos, w][order].astype(np.float32))
    send_counts = np.bincount(own, minlength=world).tolist()
    recv_counts = exchange_counts(send_counts)
    got = exchange_rows(rows, send_counts, recv_counts).numpy()
    assert got.shape == (sum(recv_counts), 4)
    assert np.all(_owner(got[:, :3], n, box, plan) == rank)
    tot = torch.tensor([got.shape[0], N], dtype=torch.int64)
    dist.all_reduce(tot)
    assert tot[0] == tot[1]
    csum = torch.tensor([float(got[:, 3].sum(dtype='f8')), float(w.sum(dtype='f8'))], dtype=torch.float64)
    dist.all_reduce(csum)
    assert abs(csum[0] - csum[1]) < 1e-6 * csum[1]


@pytest.mark.parametrize('world', [2, 3])
def test_routing_exchange(world):
    _run(_routing, world, 24, 100.0)


def _ghosts(rank, world, n, box, shifted):
    """Each rank TSC-deposits (CPU oracle) the particles it owns into an extended local grid, the ghost
    exchange folds the planes, and the owned planes must equal the slab of the global deposit."""
    from abacusutils_b200.dist import SlabPlan, exchange_ghost_planes
    from oracle import abk_oracle as O

    plan = SlabPlan(n, world)
    rng = np.random.default_rng(7)  # same catalogue on every rank
    N = 4000
    pos = rng.random((N, 3), dtype='f4') * np.float32(box)
    w = rng.random(N, dtype='f4')
    off = 0.5 * box / n if shifted else 0.0
    ref = np.zeros((n, n, n), dtype=np.float32)
    O.tsc_scatter_serial(pos, ref, box, weights=w, offset=off)

    x_lo, x_hi = plan.x_range(rank)
    nxl = x_hi - x_lo
    mine = _owner(pos, n, box, plan) == rank
    # deposit into a periodic scratch grid, then cut out planes x_lo-1 .. x_lo+nxl+1 (what the slab kernel fills)
    scratch = np.zeros((n, n, n), dtype=np.float32)
    O.tsc_scatter_serial(pos[mine], scratch, box, weights=w[mine], offset=off)
    planes = [(x_lo - 1 + p) % n for p in range(nxl + 3)]
    assert scratch.sum(dtype='f8') == pytest.approx(scratch[sorted(set(planes))].sum(dtype='f8'), rel=1e-6)
    local = torch.from_numpy(np.ascontiguousarray(scratch[planes]))
    if world == 1 or n <= nxl + 3:
        pytest.skip('degenerate')

    def add_planes(dst, src):
        dst += src

    exchange_ghost_planes(local, nxl, add_planes)
    np.testing.assert_allclose(local[1:nxl + 1].numpy(), ref[x_lo:x_hi], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('world', [2, 3])
@pytest.mark.parametrize('shifted', [False, True])
def test_ghost_exchange(world, shifted):
    _run(_ghosts, world, 24, 100.0, shifted)


def _transpose(rank, world, n):
    """Distributed rfftn = local rfft2 -> pack -> all-to-all -> local fft along x, vs np.fft.rfftn."""
    from abacusutils_b200.dist import SlabPlan, transpose_slab_to_pencil

    plan = SlabPlan(n, world)
    rng = np.random.default_rng(3)
    field = rng.standard_normal((n, n, n)).astype(np.float32)
    x_lo, x_hi = plan.x_range(rank)
    j0, j1 = plan.jsplit[rank], plan.jsplit[rank + 1]
    slab = np.fft.rfft2(field[x_lo:x_hi].astype(np.float64), axes=(1, 2)).astype(np.complex64)  # [nxl][n][nzc]
    # the layout abk_transpose_pack produces: per destination q the block [nxl][nyl_q][nzc]
    packed = np.concatenate([slab[:, plan.jsplit[q]:plan.jsplit[q + 1], :].reshape(-1) for q in range(world)])
    pencil = transpose_slab_to_pencil(torch.from_numpy(packed), plan, rank).numpy()
    assert pencil.shape == (n, j1 - j0, n // 2 + 1)
    got = np.fft.fft(pencil.astype(np.complex128), axis=0)
    want = np.fft.rfftn(field.astype(np.float64))[:, j0:j1, :]
    np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4 * np.abs(want).max())


@pytest.mark.parametrize('world,n', [(2, 12), (3, 10)])
def test_transpose(world, n):
    _run(_transpose, world, n)


def _tile_ids(pos, n, box):
    c = (np.rint(pos * np.float32(n / box)).astype(np.int64)) % n
    nty, ntz = -(-n // 8), -(-n // 32)
    return ((c[:, 0] // 8) * nty + c[:, 1] // 8) * ntz + c[:, 2] // 32, (-(-n // 8)) * nty * ntz


def _bucketed(rank, world, n, box):
    """exchange_bucketed: tile-bucketed records + tile-offset slices arrive so that, per source rank and
    local tile, the slice [starts[t]-base, starts[t+1]-base) holds exactly that source's particles of the tile."""
    from abacusutils_b200.dist import SlabPlan, exchange_bucketed

    plan = SlabPlan(n, world)
    assert plan.aligned
    nty, ntz = -(-n // 8), -(-n // 32)
    per_col = nty * ntz
    rng = np.random.default_rng(500 + rank)
    N = 3000 + 500 * rank
    pos = rng.random((N, 3), dtype='f4') * np.float32(box)
    w = rng.random(N, dtype='f4')
    tid, ntiles = _tile_ids(pos, n, box)
    order = np.argsort(tid, kind='stable')
    rows = torch.from_numpy(np.c_[pos, w][order].astype(np.float32))
    starts = torch.from_numpy(np.searchsorted(tid[order], np.arange(ntiles + 1), side='left').astype(np.int32))
    t0 = [(plan.xsplit[r] // 8) * per_col for r in range(world)] + [ntiles]
    parts = exchange_bucketed(rows, starts, t0)
    assert len(parts) == world
    # what every source holds for my tiles: gather all catalogues (small) and recompute
    sizes = [3000 + 500 * q for q in range(world)]
    got_total = 0
    for q, (recs, st, base) in enumerate(parts):
        rq = np.random.default_rng(500 + q)
        pq = rq.random((sizes[q], 3), dtype='f4') * np.float32(box)
        wq = rq.random(sizes[q], dtype='f4')
        tq, _ = _tile_ids(pq, n, box)
        recs, st = recs.numpy(), st.numpy().astype(np.int64)
        assert len(st) == t0[rank + 1] - t0[rank] + 1
        assert st[0] == base and st[-1] - base == len(recs)
        for t in range(0, t0[rank + 1] - t0[rank], 7):  # every 7th tile
            sl = recs[st[t] - base: st[t + 1] - base]
            mine = tq == t0[rank] + t
            want = np.c_[pq[mine], wq[mine]]
            assert sl.shape == want.shape
            assert np.array_equal(sl[np.lexsort(sl.T[::-1])], want[np.lexsort(want.T[::-1])])
        got_total += len(recs)
    tot = torch.tensor([got_total], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot) == sum(sizes)


@pytest.mark.parametrize('world,n', [(2, 32), (3, 40)])
def test_bucketed_exchange(world, n):
    _run(_bucketed, world, n, 100.0)
