"""
Multi-GPU power spectrum: the mesh is sharded by x-planes over the ranks of a torch.distributed
process group (one process per GPU, NCCL over NVLink/NVSwitch).  The reference has no distributed path
(/root/reference/docs/tutorials/analysis/tsc.ipynb:21 "can't readily scale to multiple nodes"); this
module extends ``calc_power`` to a sharded mesh with the four exchanges SURVEY.md 8(e) lists:

  1. particle routing   -- all-to-all of (x,y,z,w) records to the rank owning the centre cell's x-plane
  2. ghost planes       -- each rank deposits into its planes plus 1 ghost plane below and 2 above (the
                           half-cell-shifted cloud reaches one plane further), sends them to its ring
                           neighbours, which add them (abk_add_planes)
  3. FFT transpose      -- local 2-D R2C over (y,z), slab -> pencil all-to-all (packed by
                           abk_transpose_pack), local 1-D C2C along x; the spectrum stays in the pencil
                           layout [x][y_local][kz]: the binning kernel takes the layout as strides
  4. bin all-reduce     -- float64 / int64 sums of the (k,mu) bins and multipoles

``calc_power(pos_local, ...)`` has the reference's signature; ``pos_local`` is THIS rank's share of the
catalogue (any particles, they are routed).  Every rank returns the same table.

The data-movement helpers take the process group and plain tensors, so the host logic is exercised on
CPU with the gloo backend (tests/test_dist_cpu.py) -- the kernels themselves need a GPU.
"""

from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np

from ._lib import ABK_MAX_SEGMENTS, BinRequest, Engine, KMesh, check, is_torch_tensor, ptr
from .analysis import power_spectrum as ps
from .analysis.tsc import padded_ldz


TILE_X, TILE_Y, TILE_Z = 8, 8, 30  # deposit tile (abk_common.cuh: ABK_TX, ABK_TY, ABK_TZ; checked in tests/test_abi.py)


# ------------------------------------------------------------------------------------------- plan
class SlabPlan:
    """x-plane ownership (slabs) and y-row ownership (pencils) of an n^3 mesh over `world` ranks.
    Uneven splits are allowed (n need not be a multiple of world); every rank needs >= 2 planes."""

    def __init__(self, n, world):
        if world < 1 or n // world < 2:
            raise ValueError(f'mesh size {n} is too small for {world} ranks (need >= 2 planes per rank)')
        self.n = int(n)
        self.world = int(world)
        self.nzc = n // 2 + 1
        self.jsplit = [r * n // world for r in range(world + 1)]
        # x-planes: cut on multiples of the deposit tile width (8 planes) when every rank can get at least one
        # tile column; then particles bucketed by GLOBAL tile are already grouped by owner and the tile-bucketed
        # records can be exchanged as they are (route_bucketed).  Otherwise plain even splits (route + re-bucket).
        ncol = -(-n // TILE_X)
        self.aligned = ncol >= world
        if self.aligned:
            self.xsplit = [min(n, (r * ncol // world) * TILE_X) for r in range(world)] + [n]
            if min(self.xsplit[r + 1] - self.xsplit[r] for r in range(world)) < 2:
                self.aligned = False
        if not self.aligned:
            self.xsplit = list(self.jsplit)

    def x_range(self, r):
        return self.xsplit[r], self.xsplit[r + 1]

    def nxl(self, r):
        return self.xsplit[r + 1] - self.xsplit[r]

    def nyl(self, r):
        return self.jsplit[r + 1] - self.jsplit[r]

    def owner_of_plane(self, ix):
        return int(np.searchsorted(self.xsplit, ix % self.n, side='right') - 1)

    def transpose_splits(self, r):
        """(input_split_sizes, output_split_sizes) in complex elements for rank r's slab->pencil all-to-all."""
        nxl, nyl, nzc = self.nxl(r), self.nyl(r), self.nzc
        send = [nxl * self.nyl(q) * nzc for q in range(self.world)]
        recv = [self.nxl(q) * nyl * nzc for q in range(self.world)]
        return send, recv


# ------------------------------------------------------------------------------------------- exchanges
def _dist():
    import torch.distributed as dist

    return dist


def exchange_counts(send_counts, group=None, device=None):
    """all-to-all of per-destination counts -> per-source counts (lists of ints)."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    t = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    out = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(out, t, group=group)
    return [int(v) for v in out.tolist()]


def exchange_rows(rows, send_counts, recv_counts, group=None):
    """Variable all-to-all of the rows of a 2-D tensor grouped by destination rank."""
    import torch

    dist = _dist()
    out = torch.empty((int(sum(recv_counts)), rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(out, rows, output_split_sizes=list(recv_counts), input_split_sizes=list(send_counts),
                           group=group)
    return out


def exchange_bucketed(rows, starts, t0, group=None):
    """Exchange tile-bucketed records between ranks (aligned slab plans).

    rows   : (N, 4) float32 records sorted by GLOBAL tile id (abk_tsc_bucket)
    starts : (ntiles + 1,) int32 exclusive tile offsets into `rows`
    t0     : list of world + 1 global tile ids; rank r owns tiles [t0[r], t0[r+1])
    Returns a list with one entry per source rank q: (records (M_q, 4), starts_slice (ntiles_me + 1,), base_q)
    where starts_slice holds the SENDER's offsets: records of local tile t are
    records[starts_slice[t] - base_q : starts_slice[t + 1] - base_q].
    """
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    idx = torch.tensor(t0, dtype=torch.int64, device=starts.device)
    b = [int(v) for v in starts[idx].to(torch.int64).tolist()]  # record offset where each owner's tiles begin
    send_counts = [b[r + 1] - b[r] for r in range(world)]
    # counts and base offsets (value of starts[] at the first tile of the destination) in one exchange
    meta = torch.tensor([[send_counts[r], b[r]] for r in range(world)], dtype=torch.int64, device=starts.device)
    meta_in = torch.empty_like(meta)
    dist.all_to_all_single(meta_in, meta, group=group)
    recv_counts = [int(v) for v in meta_in[:, 0].tolist()]
    bases = [int(v) for v in meta_in[:, 1].tolist()]
    recv = exchange_rows(rows, send_counts, recv_counts, group)
    # tile-offset slices: to owner r goes starts[t0[r] .. t0[r+1]] (inclusive end)
    sizes_out = [t0[r + 1] - t0[r] + 1 for r in range(world)]
    send_st = torch.cat([starts[t0[r]: t0[r + 1] + 1] for r in range(world)])
    mine = sizes_out[rank]
    recv_st = torch.empty(mine * world, dtype=torch.int32, device=starts.device)
    dist.all_to_all_single(recv_st, send_st, output_split_sizes=[mine] * world, input_split_sizes=sizes_out, group=group)
    out, off = [], 0
    for q in range(world):
        out.append((recv[off: off + recv_counts[q]], recv_st[mine * q: mine * (q + 1)], bases[q]))
        off += recv_counts[q]
    return out


def exchange_ghost_planes(grid, nxl, add_planes, group=None):
    """Ring exchange of the ghost planes of a slab grid laid out as
        plane 0          : ghost  x_lo - 1        -> added to the LEFT neighbour's last owned plane
        planes 1..nxl    : owned
        planes nxl+1,+2  : ghosts x_lo+nxl, +1    -> added to the RIGHT neighbour's first two owned planes
    `add_planes(dst_view, src)` performs dst += src (abk_add_planes on the GPU)."""
    import torch

    dist = _dist()
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    lo = grid[0:1]
    hi = grid[nxl + 1:nxl + 3]
    if world == 1:
        add_planes(grid[nxl:nxl + 1], lo.clone())
        add_planes(grid[1:3], hi.clone())
        return
    left, right = (rank - 1) % world, (rank + 1) % world
    from_right = torch.empty_like(lo)   # the right neighbour's low ghost lands on my last owned plane
    from_left = torch.empty_like(hi)    # the left neighbour's high ghosts land on my first two planes
    gl = dist.get_global_rank(group, left) if group is not None else left
    gr = dist.get_global_rank(group, right) if group is not None else right
    ops = [dist.P2POp(dist.isend, lo.contiguous(), gl, group=group), dist.P2POp(dist.isend, hi.contiguous(), gr, group=group),
           dist.P2POp(dist.irecv, from_right, gr, group=group), dist.P2POp(dist.irecv, from_left, gl, group=group)]
    if world == 2:
        # both neighbours are the same peer: one message each way carrying all three planes [lo | hi, hi+1]
        out = torch.cat([lo, hi])
        inc = torch.empty_like(out)
        for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, out, gr, group=group), dist.P2POp(dist.irecv, inc, gr, group=group)]):
            req.wait()
        from_right, from_left = inc[0:1], inc[1:3]
    else:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    add_planes(grid[nxl:nxl + 1], from_right)
    add_planes(grid[1:3], from_left)


def transpose_slab_to_pencil(packed, plan, rank, group=None):
    """packed: complex64 1-D tensor holding, destination by destination, the blocks [nxl][nyl_q][nzc]
    (abk_transpose_pack).  Returns the pencil [n][nyl][nzc] of this rank."""
    import torch

    dist = _dist()
    send, recv = plan.transpose_splits(rank)
    out = torch.empty(plan.n * plan.nyl(rank) * plan.nzc, dtype=packed.dtype, device=packed.device)
    a, b = torch.view_as_real(packed), torch.view_as_real(out)
    dist.all_to_all_single(b, a, output_split_sizes=recv, input_split_sizes=send, group=group)
    return out.view(plan.n, plan.nyl(rank), plan.nzc)


# ------------------------------------------------------------------------------------------- peer memory
_SYMM = {'ok': {}, 'bufs': {}}      # per process group: symmetric memory usable?, cached buffers

# optional phase timing (bench.py's multi-GPU arm): CUDA-event pairs on the current stream around the phases of the step
_PHASES = {'on': False, 'recs': []}


class _phase:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _PHASES['on']:
            import torch

            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PHASES['on']:
            self.b.record()
            _PHASES['recs'].append((self.name, self.a, self.b))
        return False


def phase_times(reset=True):
    """{phase: total ms} of the recorded phases (synchronises)."""
    import torch

    torch.cuda.synchronize()
    out = {}
    for name, a, b in _PHASES['recs']:
        out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    if reset:
        _PHASES['recs'] = []
    return out



def _symm_pencil(group, nelem, slot, device):
    """A symmetric-memory (peer-mapped over NVLink) complex64 buffer of `nelem` elements, cached per slot.
    Returns (tensor, handle) or None when symmetric memory is unavailable (then NCCL all-to-all is used)."""
    import os

    import torch

    key = (id(group), slot)
    if os.environ.get('ABK_NO_P2P') == '1' or _SYMM['ok'].get(id(group)) is False:
        return None
    ent = _SYMM['bufs'].get(key)
    if ent is not None and ent[0].numel() >= nelem:
        return ent
    dist = _dist()
    g = group if group is not None else dist.group.WORLD
    got, err = None, None
    try:
        import torch.distributed._symmetric_memory as symm_mem

        t = symm_mem.empty(int(nelem), dtype=torch.complex64, device=device)
        hdl = symm_mem.rendezvous(t, g)
        got = (t, hdl)
    except Exception as e:  # pragma: no cover - depends on the platform
        err = e
    # every rank must take the same path (peer stores + device barriers vs NCCL all-to-all): agree on the outcome
    flag = torch.tensor([1 if got is not None else 0], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if int(flag.item()) == 0:
        import warnings

        why = f'{type(err).__name__}: {err}' if err is not None else 'failed on another rank'
        warnings.warn(f'symmetric memory unavailable ({why}); the FFT transpose uses NCCL all-to-all')
        _SYMM['ok'][id(group)] = False
        return None
    _SYMM['ok'][id(group)] = True
    _SYMM['bufs'][key] = got
    return got


# ------------------------------------------------------------------------------------------- pipeline
class DistEngine:
    def __init__(self, group=None):
        import torch

        dist = _dist()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.eng = Engine.get(torch.cuda.current_device())
        self.device = self.eng.device

    # -- 1. routing -----------------------------------------------------------------------------------
    def route(self, pos, w, plan, Lbox, paste='TSC'):
        """Returns the (M,4) float32 records of the particles whose centre cell this rank owns."""
        import torch

        eng = self.eng
        eng.bind_stream()
        eng.set_scheme(paste)
        pos_d = eng.to_device(pos, torch.float32)
        w_d = None if w is None else eng.to_device(w, torch.float32)
        N = int(pos_d.shape[0])
        xs = (C.c_int32 * (self.world + 1))(*plan.xsplit)
        counts = (C.c_int64 * self.world)()
        out = eng.scratch('route_out', max(N, 1) * 16)
        check(eng.lib.abk_route_particles(eng.ctx, ptr(pos_d), ptr(w_d), N, plan.n, float(Lbox),
                                          0 if str(paste).upper() == 'CIC' else 1, self.world, xs,
                                          ptr(out), counts))
        send_counts = [int(c) for c in counts]
        rows = out[: N * 16].view(torch.float32).view(N, 4)
        if self.world == 1:
            return rows
        recv_counts = exchange_counts(send_counts, self.group, self.device)
        return exchange_rows(rows, send_counts, recv_counts, self.group)

    def route_bucketed(self, pos, w, plan, Lbox, paste='TSC'):
        """Aligned plans: bucket the local particles by GLOBAL tile (one histogram + scatter), then exchange
        the tile-bucketed records and the matching slices of the tile-offset table.  Every rank ends up with
        one bucket segment per (source rank, local chunk of <= 2^30 particles), which the tile kernel consumes
        directly: no second bucketing.  Returns (segments, M, keep-alive) with segments = [(record_ptr,
        starts_ptr, count)]."""
        import torch

        eng, n = self.eng, plan.n
        lib = eng.lib
        eng.bind_stream()
        eng.set_scheme(paste)
        # host inputs (CPU torch tensors / NumPy arrays of float32): copied chunk by chunk on a copy stream further down, so
        # that the bucketing of chunk c overlaps the host->device copy of chunk c+1; anything else is moved up front
        host_src = None
        if not (is_torch_tensor(pos) and pos.is_cuda):
            src = pos if is_torch_tensor(pos) else torch.from_numpy(np.ascontiguousarray(pos))
            wsrc = None if w is None else (w if is_torch_tensor(w) else torch.from_numpy(np.ascontiguousarray(w)))
            if src.dtype == torch.float32 and src.is_contiguous() and src.device.type == 'cpu' and \
                    (wsrc is None or (wsrc.dtype == torch.float32 and wsrc.is_contiguous() and wsrc.device.type == 'cpu')):
                host_src = (src, wsrc)
        if host_src is not None:
            pos_d = eng.empty(tuple(host_src[0].shape), torch.float32)
            w_d = None if w is None else eng.empty((int(host_src[0].shape[0]),), torch.float32)
        else:
            pos_d = eng.to_device(pos, torch.float32)
            w_d = None if w is None else eng.to_device(w, torch.float32)
        N = int(pos_d.shape[0])
        nty, ntz = -(-n // TILE_Y), -(-n // TILE_Z)
        per_col = nty * ntz
        ntiles = -(-n // TILE_X) * per_col
        t0 = [(plan.xsplit[r] // TILE_X) * per_col for r in range(self.world)] + [ntiles]
        CH = 1 << 30
        nchunk_local = max(1, -(-N // CH))
        # every rank must take part in the same number of exchanges
        nchunk = nchunk_local
        if self.world > 1:
            t = torch.tensor([nchunk_local], dtype=torch.int64, device=self.device)
            _dist().all_reduce(t, op=_dist().ReduceOp.MAX, group=self.group)
            nchunk = int(t.item())
            # pipelining: more, smaller chunks so that the exchange of chunk c (NCCL, its own stream) runs while chunk c+1
            # is being bucketed; the tile kernel takes at most 16 segments = chunks x ranks
            nchunk = max(nchunk, min(ABK_MAX_SEGMENTS // self.world, int(os.environ.get('ABK_DIST_CHUNKS', '4'))))
        if nchunk * self.world > ABK_MAX_SEGMENTS:
            raise NotImplementedError(f'{nchunk} local chunks x {self.world} ranks exceed the {ABK_MAX_SEGMENTS} bucket segments of the tile kernel')
        csize = max(1, -(-N // nchunk))
        nb = C.c_size_t()
        check(lib.abk_tsc_bucket_scratch_bytes(max(min(N, csize), 1), n, n, n, C.byref(nb)))
        scan_buf = eng.scratch('bucket_scan', nb.value + 256)
        scan_ptr = C.c_void_p((scan_buf.data_ptr() + 255) & ~255)
        wrap = 0 if str(paste).upper() == 'CIC' else 1
        segs, keep, total = [], [], 0
        compute = eng.bind_stream()

        copied = {}
        if host_src is not None and N > 0:
            copy_stream = self._copy_stream()
            copy_stream.wait_stream(compute)
            with torch.cuda.stream(copy_stream):
                for c in range(nchunk):
                    a, bnd = min(c * csize, N), min((c + 1) * csize, N)
                    if bnd > a:
                        pos_d[a:bnd].copy_(host_src[0][a:bnd], non_blocking=True)
                        if w_d is not None:
                            w_d[a:bnd].copy_(host_src[1][a:bnd], non_blocking=True)
                    copied[c] = torch.cuda.Event()
                    copied[c].record(copy_stream)

        # odd chunks are bucketed on a second stream: the histogram pass of chunk c+1 (L2-reduction-bound) shares the GPU with
        # the scatter pass of chunk c (latency-bound), as in the single-GPU painter
        side = None
        if nchunk > 1 and os.environ.get('ABK_BUCKET_STREAMS', '2') != '1':
            side = self._side_stream()
            side.wait_stream(compute)
            scan_side = eng.scratch('bucket_scan_side0', nb.value + 256)
            scan_ptr_side = C.c_void_p((scan_side.data_ptr() + 255) & ~255)

        def bucket(c):
            a, bnd = min(c * csize, N), min((c + 1) * csize, N)
            m = bnd - a
            st = side if (side is not None and (c & 1)) else compute
            if c in copied:
                st.wait_event(copied[c])
            if self.world == 1:
                rec = eng.scratch(f'route_out{c}', max(m, 1) * 16)
                starts = eng.scratch(f'route_starts{c}', (ntiles + 1) * 4).view(torch.int32)[: ntiles + 1]
            else:  # send-side buffers are dead after the exchange: plain tensors, returned to the allocator
                rec = torch.empty(max(m, 1) * 16, dtype=torch.uint8, device=self.device)
                starts = torch.empty(ntiles + 1, dtype=torch.int32, device=self.device)
                if rec.is_cuda and st is not compute:
                    rec.record_stream(st)
                    starts.record_stream(st)
            with torch.cuda.stream(st):
                eng.bind_stream()
                check(lib.abk_tsc_bucket(eng.ctx, ptr(pos_d[a:bnd]) if m else None, ptr(w_d[a:bnd]) if (w_d is not None and m) else None,
                                         m, n, n, n, float(Lbox), 0.0, wrap, ptr(rec), ptr(starts),
                                         scan_ptr_side if st is not compute else scan_ptr, nb.value))
                ev = torch.cuda.Event()
                ev.record(st)
            eng.bind_stream()
            return rec[: m * 16].view(torch.float32).view(m, 4), starts, m, ev

        if self.world == 1:
            for c in range(nchunk):
                rows, starts, m, _ = bucket(c)
                segs.append((rows.data_ptr(), starts.data_ptr(), m))
                keep += [rows, starts]
                total += m
            if side is not None:
                compute.wait_stream(side)
            return segs, total, keep

        comm = self._comm_stream()
        nxt = bucket(0)
        for c in range(nchunk):
            rows, starts, m, ev = nxt
            if c + 1 < nchunk:
                nxt = bucket(c + 1)      # queued before this chunk's exchange blocks the host on its sizes
            comm.wait_event(ev)
            with torch.cuda.stream(comm):
                parts = exchange_bucketed(rows, starts, t0, self.group)
                if rows.is_cuda:      # allocator bookkeeping: tensors cross between the compute and the exchange stream
                    for recs_q, st_q, _ in parts:
                        recs_q.record_stream(compute)
                        st_q.record_stream(compute)
                    rows.record_stream(comm)
                    starts.record_stream(comm)
            off = 0
            for recs_q, st_q, base_q in parts:
                # the slice holds the SENDER's record offsets: bias the record pointer instead of rewriting it
                segs.append((recs_q.data_ptr() - 16 * base_q, st_q.data_ptr(), int(recs_q.shape[0])))
                off += int(recs_q.shape[0])
            keep.append(parts)
            total += off
            del rows, starts
        compute.wait_stream(comm)
        if side is not None:
            compute.wait_stream(side)
        return segs, total, keep

    def _side_stream(self):
        import torch

        if getattr(self.eng, '_side', None) is None:
            self.eng._side = torch.cuda.Stream(device=self.device)
        return self.eng._side

    def _copy_stream(self):
        import torch

        if getattr(self.eng, '_copy', None) is None:
            self.eng._copy = torch.cuda.Stream(device=self.device)
        return self.eng._copy

    def _comm_stream(self):
        import torch

        if getattr(self.eng, '_comm', None) is None:
            self.eng._comm = torch.cuda.Stream(device=self.device, priority=-1)
        return self.eng._comm

    def paint_segments(self, segs, plan, Lbox, offsets, paste='TSC', bucket_offset=None, init=0.0, ghosts=True):
        """Tile deposit of pre-bucketed segments (route_bucketed) into slab grids, then the ghost exchange.  `init`: start
        value of the OWNED planes (-1 when the field normalisation is folded into the deposit, power_spectrum.py:860-901)."""
        import torch

        eng, n = self.eng, plan.n
        lib = eng.lib
        eng.set_scheme(paste)
        x_lo, x_hi = plan.x_range(self.rank)
        nxl = x_hi - x_lo
        ldz = padded_ldz(n)
        live = [sg for sg in segs if sg[2] > 0]
        grids = []
        for off in offsets:
            grid = eng.zeros((nxl + 3, n, ldz), torch.float32)
            if init != 0.0:
                grid[1:nxl + 1].fill_(init)
            if live:
                m = len(live)
                recs = (C.c_void_p * m)(*[sg[0] for sg in live])
                sts = (C.c_void_p * m)(*[sg[1] for sg in live])
                cnts = (C.c_int64 * m)(*[sg[2] for sg in live])
                eng.bind_stream()
                check(lib.abk_tsc_deposit_tiles(eng.ctx, m, recs, sts, cnts, ptr(grid), n, n, n, ldz, float(Lbox), float(off),
                                                float(offsets[0] if bucket_offset is None else bucket_offset), 1, x_lo, nxl))

            if ghosts:
                self.fold_ghosts(grid, plan)
            grids.append(grid)
        return grids

    def fold_ghosts(self, grid, plan):
        """Ring exchange of the three ghost planes of a slab grid; the received planes are added to the owned ones."""
        eng, n = self.eng, plan.n
        x_lo, x_hi = plan.x_range(self.rank)
        ldz = padded_ldz(n)

        def add_planes(dst, src):
            eng.bind_stream()
            check(eng.lib.abk_add_planes(eng.ctx, ptr(dst), ptr(src.contiguous()), dst.shape[0], n, n, ldz))

        exchange_ghost_planes(grid, x_hi - x_lo, add_planes, self.group)

    # -- 2. deposit + ghosts ----------------------------------------------------------------------------
    def paint_slab(self, records, plan, Lbox, offsets, paste='TSC', bucket_offset=None):
        """Deposit this rank's records for every offset into slab grids [(nxl+3), n, ldz] and fold the ghosts."""
        import torch

        eng, n = self.eng, plan.n
        lib = eng.lib
        eng.set_scheme(paste)
        x_lo, x_hi = plan.x_range(self.rank)
        nxl = x_hi - x_lo
        ldz = padded_ldz(n)
        M = int(records.shape[0])
        grids = []
        # bucket once by the unshifted centre cell (tiles cover the nxl owned planes); the shifted deposit
        # reuses the records with the widened tile domain (abk_tsc_deposit_tiles, bucket_offset)
        rec = starts = None
        if M > 0:
            if M > 1 << 30:
                raise NotImplementedError('more than 2^30 particles per rank')
            ntiles = C.c_int64()
            check(lib.abk_tsc_num_tiles(nxl, n, n, C.byref(ntiles)))
            nb = C.c_size_t()
            check(lib.abk_tsc_bucket_scratch_bytes(M, nxl, n, n, C.byref(nb)))
            scan_tmp = eng.scratch('bucket_scan', nb.value)
            rec = eng.scratch('records_slab', M * 16)
            starts = eng.scratch('starts_slab', (ntiles.value + 1) * 4)
            dropped = C.c_ulonglong(0)
            eng.bind_stream()
            boff = float(offsets[0] if bucket_offset is None else bucket_offset)
            check(lib.abk_tsc_bucket_slab(eng.ctx, ptr(records), None, M, 1, n, n, n, float(Lbox), boff, 0,
                                          x_lo, nxl, ptr(rec), ptr(starts), ptr(scan_tmp), scan_tmp.numel(),
                                          C.byref(dropped)))
            if dropped.value:
                raise RuntimeError(f'rank {self.rank}: {dropped.value} routed particles fall outside the slab')
        for off in offsets:
            grid = eng.zeros((nxl + 3, n, ldz), torch.float32)
            if M > 0:
                recs = (C.c_void_p * 1)(rec.data_ptr())
                sts = (C.c_void_p * 1)(starts.data_ptr())
                cnts = (C.c_int64 * 1)(M)
                eng.bind_stream()
                check(lib.abk_tsc_deposit_tiles(eng.ctx, 1, recs, sts, cnts, ptr(grid), n, n, n, ldz, float(Lbox),
                                                float(off), float(offsets[0] if bucket_offset is None else bucket_offset), 1,
                                                x_lo, nxl))

            def add_planes(dst, src):
                eng.bind_stream()
                check(lib.abk_add_planes(eng.ctx, ptr(dst), ptr(src.contiguous()), dst.shape[0], n, n, ldz))

            exchange_ghost_planes(grid, nxl, add_planes, self.group)
            grids.append(grid)
        return grids

    # -- 3. distributed FFT -------------------------------------------------------------------------------
    def _plan(self, kind, *dims):
        eng = self.eng
        key = (kind,) + dims
        if key not in eng._plans:
            h, wb = C.c_void_p(), C.c_size_t()
            fn = eng.lib.abk_fft_yz_plan_create if kind == 'yz' else eng.lib.abk_fft_x_plan_create
            check(fn(eng.ctx, *dims, C.byref(h), C.byref(wb)))
            eng._plans[key] = (h, int(wb.value))
        return eng._plans[key]

    def _exec(self, plan_h, wb, data_ptr):
        eng = self.eng
        work = eng.scratch('fftwork', wb)
        eng.bind_stream()
        check(eng.lib.abk_fft_exec_generic(eng.ctx, plan_h, C.c_void_p(data_ptr), ptr(work), work.numel()))

    def fft_slab(self, grid_box, plan, n_total, slot=0):
        """Normalise the owned planes, 2-D FFT, transpose, 1-D FFT.  Returns the pencil [n][nyl][nzc] complex64.
        `grid_box` is a one-element list holding the slab; it is emptied as soon as the slab has been packed, so
        the slab's memory is free again before the pencil is allocated (35 GB each at nmesh 4096 on 8 GPUs)."""
        import torch

        grid = grid_box[0]
        eng, n = self.eng, plan.n
        nxl, nyl, nzc = plan.nxl(self.rank), plan.nyl(self.rank), plan.nzc
        ldz = padded_ldz(n)
        owned = grid[1:nxl + 1]
        eng.bind_stream()
        if n_total is not None:      # None: the deposit already produced rho * n^3/N - 1
            check(eng.lib.abk_normalize_field(eng.ctx, ptr(owned), nxl, n, n, ldz, float(n) ** 3, float(n_total)))
        h, wb = self._plan('yz', nxl, n, n)
        self._exec(h, wb, owned.data_ptr())
        js = (C.c_int64 * (self.world + 1))(*plan.jsplit)
        sym = None
        if self.world > 1:
            # every rank allocates the same (largest) pencil size: symmetric memory needs identical shapes
            sym = _symm_pencil(self.group, n * max(plan.nyl(r) for r in range(self.world)) * nzc, slot, self.device)
        if sym is not None:
            # fused pack + transfer: rows go straight into the owners' pencil buffers over NVLink peer memory
            buf, hdl = sym
            hdl.barrier(channel=0)  # nobody still reads the buffer's previous contents
            peers = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
            eng.bind_stream()
            check(eng.lib.abk_transpose_scatter_p2p(eng.ctx, ptr(owned), peers, nxl, n, nzc, self.world, js,
                                                    plan.xsplit[self.rank]))
            hdl.barrier(channel=0)  # every rank's stores have landed
            del owned, grid
            grid_box.clear()
            pencil = buf[: n * nyl * nzc].view(n, nyl, nzc)
        else:
            packed = eng.empty((nxl * n * nzc,), torch.complex64)
            check(eng.lib.abk_transpose_pack(eng.ctx, ptr(owned), ptr(packed), nxl, n, nzc, self.world, js))
            del owned, grid
            grid_box.clear()
            if self.world == 1:
                pencil = packed.view(n, nyl, nzc)
            else:
                pencil = transpose_slab_to_pencil(packed, plan, self.rank, self.group)
        h, wb = self._plan('x', n, nyl, nzc)
        self._exec(h, wb, pencil.data_ptr())
        return pencil

    # -- 4. binning + all-reduce ---------------------------------------------------------------------------
    def bin_pencils(self, plan, Lbox, kedges, muedges, poles, f1, f1s, f2, f2s, W_d, scale, W_sym=False):
        import torch

        eng, n = self.eng, plan.n
        dist = _dist()
        kedges = np.asarray(kedges, dtype=np.float64)
        muedges = np.asarray(muedges, dtype=np.float64)
        poles = np.asarray(poles, dtype=np.int64).reshape(-1)
        Nk, Nmu, Np = len(kedges) - 1, len(muedges) - 1, len(poles)
        dk = 2.0 * np.pi / Lbox
        kedges2 = ((kedges / dk) ** 2).astype(np.float32)
        muedges2 = (muedges**2).astype(np.float32)
        Nb = Nk * Nmu
        tables = np.concatenate([kedges2, muedges2, ps.legendre_coefficients(poles).reshape(-1)]).astype(np.float32)
        tables_d = eng.to_device(tables, torch.float32)
        sums = eng.zeros((3 * Nb + Np * Nk,), torch.float64)
        j0, j1 = plan.jsplit[self.rank], plan.jsplit[self.rank + 1]
        nyl, nzc = j1 - j0, plan.nzc
        req = BinRequest()
        req.mesh = KMesh(n=n, nzc=nzc, i0=0, i1=n, j0=j0, j1=j1, stride_i=nyl * nzc, stride_j=nzc)
        for name, t in (('f1', f1), ('f1s', f1s), ('f2', f2), ('f2s', f2s), ('W', W_d)):
            setattr(req, name, None if t is None else t.data_ptr())
        req.real_in = None
        req.scale = float(scale)
        req.finish = 1
        req.w_symmetric = int(W_sym)
        base = tables_d.data_ptr()
        req.kedges2 = base
        req.muedges2 = base + 4 * (Nk + 1)
        req.pole_coef = base + 4 * (Nk + 1 + Nmu + 1)
        for ip in range(Np):
            req.pole_ell[ip] = int(poles[ip])
        req.Nk, req.Nmu, req.Np = Nk, Nmu, Np
        sb = sums.data_ptr()
        req.counts, req.sum_p, req.sum_k, req.sum_poles = sb, sb + 8 * Nb, sb + 16 * Nb, sb + 24 * Nb
        nb = C.c_size_t()
        check(eng.lib.abk_power_bin_scratch_bytes(Nk, Nmu, Np, C.byref(nb)))
        rep = eng.scratch('bin_replicas', nb.value)
        req.scratch, req.scratch_bytes = rep.data_ptr(), rep.numel()
        eng.bind_stream()
        check(eng.lib.abk_power_bin(eng.ctx, C.byref(req)))
        if self.world > 1:
            # ONE all-reduce for counts and sums: the integer mode counts travel as float64, which is exact (every count
            # and every partial sum of counts is far below 2^53), and are turned back into int64 bit patterns afterwards
            cnt = sums[:Nb].view(torch.int64)
            sums[:Nb] = cnt.to(torch.float64)
            dist.all_reduce(sums, group=self.group)
            cnt.copy_(sums[:Nb].round().to(torch.int64))
        return ps._finalize_bins(sums, Nk, Nmu, poles, dk)


def calc_power(pos, Lbox, kbins=None, mubins=None, k_max=None, logk=False, paste='TSC', nmesh=128, compensated=True,
               interlaced=True, w=None, pos2=None, w2=None, poles=None, squeeze_mu_axis=True, nthread=1,
               dtype=np.float32, group=None, force_reroute=False):
    """Sharded ``calc_power`` (same parameters and result table as the single-GPU / reference function,
    power_spectrum.py:1131-1319).  ``pos``/``w`` (and ``pos2``/``w2``) are this rank's share of the catalogue."""
    import torch

    dist = _dist()
    de = DistEngine(group)
    eng = de.eng
    n = int(nmesh)
    plan = SlabPlan(n, de.world)
    paste = ps._check_paste(paste)
    if kbins is None:
        kbins = nmesh
    if k_max is None:
        k_max = np.pi * nmesh / Lbox
    return_mubins = mubins is not None
    if mubins is None:
        mubins = 1

    def total(npart):
        t = torch.tensor([npart], dtype=torch.int64, device=de.device)
        if de.world > 1:
            dist.all_reduce(t, group=group)
        return int(t.item())

    def field(p, wt, slot_base=0):
        ntot = total(len(p))
        offsets = [0.0, 0.5 * (float(Lbox) / n)] if interlaced else [0.0]
        # one grid at a time (paint -> ghosts -> FFT -> drop the slab): at nmesh 4096 on 8 GPUs a slab is 35 GB
        pencils = []
        if plan.aligned and not force_reroute:
            # normalize_field folded into the deposit: weights scaled by n^3/N as the records are written, owned planes
            # start at -1 (ABK_FUSED_NORMALIZE=0 restores the separate pass)
            fused = os.environ.get('ABK_FUSED_NORMALIZE', '1') != '0' and ntot > 0
            with _phase('bucket + particle exchange'):
                if fused:
                    eng.set_weight_scale(float(np.float32(float(n) ** 3 / float(ntot))))
                try:
                    segs, _, keep = de.route_bucketed(p, wt, plan, Lbox, paste)
                finally:
                    eng.set_weight_scale(1.0)
            # the FFT + transpose of grid i runs on an auxiliary stream while grid i+1 is being painted; two slabs are
            # alive at once then, so only where they fit comfortably (a slab is 35 GB at nmesh 4096 on 8 GPUs)
            overlap = de.world > 1 and len(offsets) > 1 and plan.n <= 2048 and os.environ.get('ABK_DIST_OVERLAP', '1') != '0'
            compute = eng.bind_stream()
            aux = eng.aux_stream()
            done = None
            for io, off in enumerate(offsets):
                with _phase('deposit' if overlap else 'deposit + ghost planes'):
                    gb = de.paint_segments(segs, plan, Lbox, [off], paste, bucket_offset=offsets[0], init=-1.0 if fused else 0.0,
                                           ghosts=not overlap)
                if io == len(offsets) - 1:
                    del keep, segs  # the routed records are dead once the last grid has been painted
                if overlap:
                    # ghost planes, FFT and transpose of this grid all go to the auxiliary stream: the compute stream
                    # paints the grids back to back (all ranks issue their collectives in the same order: only `aux` does)
                    if gb[0].is_cuda:
                        gb[0].record_stream(aux)
                    aux.wait_stream(compute)
                    with torch.cuda.stream(aux):
                        de.fold_ghosts(gb[0], plan)
                        pencils.append(de.fft_slab(gb, plan, None if fused else ntot, slot=slot_base + io))
                        done = torch.cuda.Event()
                        done.record(aux)
                    eng.bind_stream()
                else:
                    with _phase('fft + transpose'):
                        pencils.append(de.fft_slab(gb, plan, None if fused else ntot, slot=slot_base + io))
            if done is not None:
                with _phase('ghost planes + fft + transpose (exposed tail)'):
                    compute.wait_event(done)
        else:
            rec = de.route(p, wt, plan, Lbox, paste)
            for io, off in enumerate(offsets):
                gb = de.paint_slab(rec, plan, Lbox, [off], paste, bucket_offset=offsets[0])
                pencils.append(de.fft_slab(gb, plan, ntot, slot=slot_base + io))
        return pencils, ntot

    g1, N1 = field(pos, w)
    g2, N2 = (field(pos2, w2, slot_base=2) if pos2 is not None else (None, None))
    meta = dict(Lbox=Lbox, logk=logk, paste=paste, nmesh=nmesh, compensated=compensated, interlaced=interlaced,
                poles=poles, nthread=nthread, N_pos=N1, is_weighted=w is not None, field_dtype=dtype,
                squeeze_mu_axis=squeeze_mu_axis, n_ranks=de.world)
    if pos2 is not None:
        meta['N_pos2'] = N2
        meta['is_weighted2'] = w2 is not None
    W = ps.get_W_compensated(Lbox, n, paste, interlaced) if compensated else None
    W_d = eng.to_device(np.asarray(W, dtype=np.float32), torch.float32) if compensated else None
    poles_arr = np.asarray(poles or [], dtype=np.int64)
    kbins, mubins = ps.get_k_mu_edges(Lbox, k_max, kbins, mubins, logk)
    scale = np.float32(0.5 / n**3) if interlaced else np.float32(1 / n**3)
    with _phase('bin + all-reduce'):
        binned = de.bin_pencils(plan, float(Lbox), kbins, mubins, poles_arr, g1[0], g1[1] if interlaced else None,
                                None if g2 is None else g2[0], g2[1] if (g2 is not None and interlaced) else None, W_d,
                                scale, W_sym=(W is None) or bool(np.array_equal(np.asarray(W)[1:], np.asarray(W)[:0:-1])))
    P = ps._package_pk(binned, Lbox, mubins, poles_arr, squeeze_mu_axis)
    kbins, mubins = np.asarray(kbins), np.asarray(mubins)
    res = dict(k_min=kbins[:-1], k_max=kbins[1:], k_mid=(kbins[1:] + kbins[:-1]) * 0.5, k_avg=P['k_avg'],
               power=P['power'], N_mode=P['N_mode'])
    if len(poles_arr) > 0:
        res.update(poles=P['binned_poles'].T, N_mode_poles=P['N_mode_poles'])
    if return_mubins:
        mu_binc = (mubins[1:] + mubins[:-1]) * 0.5
        res.update(mu_min=np.broadcast_to(mubins[:-1], res['power'].shape),
                   mu_max=np.broadcast_to(mubins[1:], res['power'].shape),
                   mu_mid=np.broadcast_to(mu_binc, res['power'].shape))
    return ps._make_table(res, meta)


# ------------------------------------------------------------------------------------------- bench (N > 1)
def parity_at_world(parity_block, world, rank, device):
    """Parity evidence AT THIS WORLD SIZE for bench.py's multi-GPU line: the sharded pipeline (particle exchange, ghost
    planes, peer-memory transpose, pencil binning, all-reduce) on (a) three committed outputs of the unmodified reference
    (tests/golden/reference_runs.npz: weighted auto, cross, default bins) and (b) 1e7 uniform particles on a 256^3 mesh
    against the single-GPU pipeline on rank 0.  Every rank passes a strided share of the catalogue."""
    import sys
    from pathlib import Path

    import torch

    from .analysis import power_spectrum as single

    root = Path(__file__).resolve().parent.parent
    out = {'world': world, 'cases': {}}
    gold_dir = root / 'tests' / 'golden'
    if (gold_dir / 'reference_runs.npz').exists():
        if str(gold_dir) not in sys.path:
            sys.path.insert(0, str(gold_dir))
        import cases

        gold = np.load(gold_dir / 'reference_runs.npz')
        for name in ('n32_ci', 'n32_cross_ci', 'n40_defaults'):
            c = cases.POWER_CASES[name]
            pos, w, pos2, w2 = cases.power_inputs(c)
            sh = lambda a: None if a is None else np.ascontiguousarray(a[rank::world])      # noqa: E731
            t = calc_power(sh(pos), c['L'], kbins=c['kbins'], mubins=c['mubins'], k_max=c.get('k_max'), logk=c['logk'], paste='TSC',
                           nmesh=c['nmesh'], compensated=c['compensated'], interlaced=c['interlaced'], w=sh(w), pos2=sh(pos2),
                           w2=sh(w2), poles=c['poles'])
            pre = f'power/{name}/'
            want = {k[len(pre):]: gold[k] for k in gold.files if k.startswith(pre)}
            amp = None
            if pos2 is not None:      # cross-spectrum of independent catalogues: a near-zero residual of the auto-power scale
                amp = np.full(np.asarray(want['power']).shape[0], float(np.abs(np.asarray(want['power'])).max()))
            out['cases'][name + ' (vs reference golden)'] = parity_block(t, want, amp=amp)
    # (b) a larger mesh: every plane / pencil split is ragged for world = 3, 5, 6, 7 and even for 2, 4, 8
    N, L, n = 10_000_000, 1000.0, 256
    gen = torch.Generator(device='cuda')
    gen.manual_seed(77)
    full = torch.rand((N, 3), device=device, dtype=torch.float32, generator=gen) * L      # same catalogue on every rank
    kw = dict(kbins=100, mubins=10, nmesh=n, compensated=True, interlaced=True, poles=[0, 2, 4])
    t = calc_power(full[rank::world].contiguous(), L, **kw)
    if rank == 0:
        ref = single.calc_power(full, L, **kw)
        out['cases']['1e7 particles, nmesh 256 (vs single-GPU pipeline on rank 0)'] = parity_block(t, ref)
    del full
    oks = [v['n_mode_exact'] and v['max_rel_power'] < 1e-4 and v['max_abs_poles_over_P0'] < 1e-4 for v in out['cases'].values()]
    out['ok'] = bool(all(oks)) if oks else None
    return out


def bench_main(args, cfg, metric, workload, ClockSampler, peaks, parity_block=None):
    """bench.py's multi-GPU arm: strong scaling of the configs[2] workload over the ranks of one node."""
    import os

    import torch

    dist = _dist()
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    # NCCL on a high-priority stream: the particle exchange of chunk c must get SMs while chunk c+1 is being bucketed
    try:
        opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank), pg_options=opts)
    except Exception:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    world, rank = dist.get_world_size(), dist.get_rank()
    eng = Engine.get(local_rank)
    N, L, n = cfg['N'], cfg['L'], cfg['nmesh']
    n_local = N // world + (1 if rank < N % world else 0)
    gen = torch.Generator(device='cuda')
    gen.manual_seed(cfg['seed'] + 1000 * rank)
    pos = torch.rand((n_local, 3), device='cuda', dtype=torch.float32, generator=gen)
    pos *= L
    kw = dict(kbins=cfg['kbins'], mubins=cfg['mubins'], nmesh=n, compensated=True, interlaced=True, poles=cfg['poles'])

    def step(p):
        return calc_power(p, L, **kw)

    for _ in range(args.warmup):
        res = step(pos)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = eng.launch_count()
    eng.profile(True)
    eng.profile_collect()
    _PHASES['on'], _PHASES['recs'] = True, []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        res = step(pos)
    ev1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device='cuda', dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    prof = eng.profile_collect()
    eng.profile(False)
    phases = {k: v / args.steps for k, v in phase_times().items()}
    _PHASES['on'] = False
    launches = eng.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None

    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((n_local, 3), dtype=torch.float32, pin_memory=True)
            pinned = True
        except Exception:
            host = torch.empty((n_local, 3), dtype=torch.float32)
            pinned = False
        host.copy_(pos)
        torch.cuda.synchronize()
        del pos
        step(host)
        e2e_steps = max(1, min(args.steps, 3))
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            res_h = step(host)
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([(time.perf_counter() - t0) / e2e_steps * 1e3], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d2h = sum(np.asarray(res_h[k]).nbytes for k in ('power', 'N_mode', 'k_avg', 'poles', 'N_mode_poles'))
        e2e = {'value': float(t.item()), 'unit': 'ms', 'h2d_bytes_per_step': int(N * 12), 'd2h_bytes_per_step': int(d2h),
               'host_memory': 'pinned' if pinned else 'pageable', 'steps': e2e_steps}
    parity = None
    if parity_block is not None:
        try:
            pos = None
            torch.cuda.empty_cache()
            parity = parity_at_world(parity_block, world, rank, eng.device)
        except Exception as e:      # evidence must not be lost silently: the failure goes into the line
            parity = {'world': world, 'ok': False, 'error': repr(e)}
    if rank == 0:
        peak, peak_src = peaks()
        stages = {k: {'ms_per_step': v[0] / args.steps, 'launches_per_step': v[1] / args.steps} for k, v in prof.items()}
        kernel_ms = sum(v['ms_per_step'] for v in stages.values())
        stages['nccl_and_host (step - kernels, rank 0)'] = {'ms_per_step': float(ms.item()) - kernel_ms}
        line = {
            'metric': metric, 'value': float(ms.item()), 'unit': 'ms', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': float(ms.item()), 'higher_is_better': False, 'scaling': 'strong',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': workload, 'parallelism': f'x-slab mesh sharding over {world} GPUs',
                       'l2': 'inputs are larger than the 126 MB L2'},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches * world),
            'roofline': {'bound': 'hbm', 'achieved': None, 'peak': peak, 'unit': 'GB/s', 'frac': None, 'traffic': None,
                         'peak_source': peak_src, 'note': 'per-kernel roofline is reported by the N=1 run'},
            'cpu_baseline': None, 'parity': parity, 'stages': stages, 'phases_ms_rank0': phases, 'mpart_per_s': N / float(ms.item()) / 1e3,
            'N_mode_total': int(np.asarray(res['N_mode']).sum()),
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    return 0
