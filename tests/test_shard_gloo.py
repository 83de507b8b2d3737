"""N > 1 host logic on CPU: world_size-2 gloo run of the unit partitioning and the variable-length
gather that bench.py / the recursion sharding use over NCCL on the GPUs."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from reveal_b200 import shard


def test_partition_is_balanced_and_deterministic():
    sizes = [100, 7, 93, 12, 50, 49, 1, 1, 30]
    parts = shard.partition(sizes, 4)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert max(loads) <= 100 and max(loads) - min(loads) <= 30
    assert parts == shard.partition(sizes, 4)
    assert shard.partition([], 3) == [[], [], []]


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # each rank "builds" its own units and holds a different number of MUM rows (rank 1: none)
    units = shard.partition([40, 10, 30, 20], world)[rank]
    k = 0 if rank == 1 else 5
    rows = torch.arange(k * 3, dtype=torch.int64).reshape(k, 3) + 1000 * rank
    got = shard.gather_rows(rows, dst=0)
    if rank == 0:
        q.put((units, [g.tolist() for g in got]))
    else:
        assert got is None
        q.put((units, None))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gathered = [r for r in res if r[1] is not None][0][1]
    assert gathered[0] == (torch.arange(15).reshape(5, 3)).tolist()
    assert gathered[1] == []
    all_units = sorted(u for r in res for u in r[0])
    assert all_units == [0, 1, 2, 3]


def _anchor_worker(rank, world, port, q, lib_path):
    import numpy as np
    import oracle.port as P
    from reveal_b200 import _native
    from util import random_related
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _native.bind(lib_path)  # the emulated kernels stand in for the GPUs of the box
    rng = np.random.default_rng(42)  # same units on every rank
    units = []
    for length, ns in ((1800, 2), (600, 3), (1200, 2), (900, 4), (300, 2)):  # pair units and multi-sample units mixed
        T, nsep, _ = P.assemble(random_related(rng, ns, length, 4))
        units.append((T, np.asarray(nsep, dtype=np.int64), ns))
    Tp, nsepp, _ = P.assemble(random_related(rng, 2, 1000, 4))
    units += shard.fwd_rc_units(Tp, np.asarray(nsepp, dtype=np.int64))   # the forward / reverse-complement pair of `finish`
    got = shard.anchor_units(units, minl=8, lib=L)
    if rank == 0:
        ok = True
        for u, res in zip(units, got):
            T, nsep, ns = u[:3]
            rc = u[3] if len(u) > 3 else 0
            o = P.Index(T, nsep, ns, rc)
            if ns == 2:
                ok = ok and np.array_equal(res, o.getmums(8, rem=not rc))
            else:  # header rows AND the member rows their `first` column indexes (positions of every multi-MUM)
                oh, om = o.getmultimums(8, 2)
                ok = ok and np.array_equal(res[0], oh) and np.array_equal(res[1], om) and len(om) > 0
        q.put(("rank0", ok, [len(r) if not isinstance(r, tuple) else len(r[0]) for r in got]))
    else:
        q.put(("rank%d" % rank, got is None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_anchor_units_sharded_world2_gloo(emu_lib):
    """Independent index builds sharded over two ranks (emulated kernels), MUM rows gathered to rank 0 and
    compared with the oracle unit by unit."""
    import sys
    lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "_build", "libreveal_emu.so")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29850 + os.getpid() % 100
    procs = [ctx.Process(target=_anchor_worker, args=(r, 2, port, q, lib_path)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for name, ok, _ in res:
        assert ok, name
    counts = [r[2] for r in res if r[2] is not None][0]
    assert len(counts) == 7 and sum(counts) > 0


def _fixed_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = shard.FixedGather(6 + 2 * rank, 3, torch.device("cpu"))  # ranks ask for different capacities: they must agree on 8
    assert g.cap == 8
    for step in range(3):
        k = 3 + rank * 2 + step
        rows = torch.arange(k * 3, dtype=torch.int64).reshape(k, 3) + 100 * rank
        g.gather(rows)
    parts = g.check()
    parts = None if parts is None else [p.tolist() for p in parts]  # views into the receive buffers: copy before the next gather
    overflow = False
    g.gather(torch.zeros((20, 3), dtype=torch.int64))  # more rows than the capacity
    try:
        g.check()
    except OverflowError:
        overflow = True
    q.put((rank, parts, overflow))
    dist.barrier()
    dist.destroy_process_group()


def test_fixed_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 100
    procs = [ctx.Process(target=_fixed_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r for r in [q.get(timeout=120) for _ in range(2)]}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[1][1] is None
    parts = res[0][1]
    assert len(parts[0]) == 5 and len(parts[1]) == 7 and parts[1][0] == [100, 101, 102]
    assert res[0][2] is True


class _BrokenOpen(object):
    """The C-ABI library with a failing rv_peer_open (what a box without CUDA IPC would look like)."""

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        return getattr(self._lib, name)

    def rv_peer_open(self, handle, out):
        return -2


def _peer_worker(rank, world, port, q, lib_path):
    import ctypes

    import numpy as np
    import oracle.port as P
    from reveal_b200 import _native
    from util import random_related
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = _native.bind(lib_path)
    cpu = torch.device("cpu")
    # a box where the mapping fails on one rank: every rank gets the same RuntimeError (collective fallback decision)
    failed = False
    try:
        shard.PeerGather(16, 3, _BrokenOpen(L) if rank == 1 else L, cpu)
    except RuntimeError:
        failed = True
    rng = np.random.default_rng(7)
    units = []  # [step][rank]
    for step in range(3):
        units.append([P.assemble(random_related(rng, 2, 500 + 200 * r + 100 * step, 4)) for r in range(world)])
    h = ctypes.c_void_p()
    _native.check(L, L.rv_index_create(ctypes.byref(h), None))
    g = shard.PeerGather(40 + 10 * rank, 3, L, cpu, depth=2)  # different requests: the ranks agree on the largest
    cap_ok = g.cap == 50
    cnt = ctypes.c_int64()
    for step in range(3):  # three steps through a ring of two: the first block is overwritten
        T, nsep, _ = units[step][rank]
        nsep = np.asarray(nsep, dtype=np.int64)
        _native.check(L, L.rv_build(h, T.ctypes.data, len(T), nsep.ctypes.data, 2, 0))
        _native.check(L, L.rv_mums_pair_count(h, 8, 1, ctypes.byref(cnt)))
        _native.check(L, L.rv_result_pack_device(h, ctypes.c_void_p(g.slot()), g.cap))
        g.advance()
    parts = g.check(expect_seq=3)
    ok = None
    if rank == 0:
        ok = len(parts) == world
        for r in range(world):
            T, nsep, _ = units[2][r]
            want = np.asarray(P.Index(T, nsep, 2).getmums(8, rem=True), dtype=np.int64).reshape(-1, 3)
            ok = ok and np.array_equal(parts[r][0], want) and len(want) > 0
    stale = False
    try:
        g.check(expect_seq=2)  # a stale block would carry another sequence number
    except RuntimeError:
        stale = True
    g.close()
    # more rows than the capacity: only min(count, cap) rows are written, check() reports it on dst
    g2 = shard.PeerGather(2, 3, L, cpu)
    _native.check(L, L.rv_result_pack_device(h, ctypes.c_void_p(g2.slot()), g2.cap))
    g2.advance()
    overflow = False
    try:
        g2.check()
    except OverflowError:
        overflow = True
    g2.close()
    L.rv_index_free(h)
    q.put((rank, failed, cap_ok, ok, stale, overflow))
    dist.barrier()
    dist.destroy_process_group()


def test_peer_gather_world2_gloo(emu_lib):
    """One-sided gather through mapped peer blocks between two real processes (emulated kernels; POSIX shared
    memory stands in for CUDA IPC): rows of the last step of both ranks, ring reuse, sequence numbers, overflow,
    and the collective failure that lets callers fall back to FixedGather."""
    lib_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu", "_build", "libreveal_emu.so")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 40
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q, lib_path)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r[0]: r for r in [q.get(timeout=300) for _ in range(2)]}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank in (0, 1):
        assert res[rank][1] is True, "broken mapping must raise on every rank"
        assert res[rank][2] is True
    assert res[0][3] is True and res[1][3] is None
    assert res[0][4] is True and res[0][5] is True       # dst detects stale blocks and overflow
    assert res[1][4] is False and res[1][5] is False     # the other ranks only take part in the barrier
