"""Sharding of independent index builds over the GPUs of one box.

The root build is one global sort (replicas only), but the path shards one level up as independent
units: recursion sub-intervals, `--order=sequential --chunksize` jobs (reveal/align.py:39-53), the
forward / reverse-complement builds of `finish` (transformold.py:142-143).  Every rank builds its
own units; the only collective on the path is the gather of the MUM records to rank 0.
Host-side logic only: works with NCCL (device tensors) and with gloo (CPU tensors, tests)."""
import torch
import torch.distributed as dist


def partition(sizes, world):
    """Greedy longest-first bin packing of units (by size) onto `world` ranks.
    Returns a list of `world` lists of unit indices; deterministic."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    return out


def gather_rows(rows, dst=0, group=None):
    """Variable-length gather of [k, c] int64 row blocks to `dst`.

    all_gather of the counts, then a gather of blocks padded to the largest count.
    Returns the list of per-rank row tensors on `dst`, None elsewhere."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    k = rows.shape[0]
    cols = rows.shape[1]
    counts = [torch.zeros(1, dtype=torch.int64, device=rows.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([k], dtype=torch.int64, device=rows.device), group=group)
    counts = [int(c.item()) for c in counts]
    mx = max(counts) if counts else 0
    pad = torch.zeros((mx, cols), dtype=torch.int64, device=rows.device)
    if k:
        pad[:k] = rows
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [bufs[r][:counts[r]] for r in range(world)]
