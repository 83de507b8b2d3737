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
    for length in (1800, 600, 1200, 900, 300):
        T, nsep, _ = P.assemble(random_related(rng, 2, length, 4))
        units.append((T, np.asarray(nsep, dtype=np.int64), 2))
    got = shard.anchor_units(units, minl=8, lib=L)
    if rank == 0:
        ok = True
        for (T, nsep, ns), rows in zip(units, got):
            o = P.Index(T, nsep, ns)
            ok = ok and np.array_equal(rows, o.getmums(8, rem=True))
        q.put(("rank0", ok, [len(r) for r in got]))
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
    assert len(counts) == 5 and sum(counts) > 0


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
