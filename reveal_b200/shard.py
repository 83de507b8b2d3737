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


class FixedGather(object):
    """Variable-length gather in ONE collective and without a host synchronisation: every rank sends a block of
    fixed capacity whose first row carries its row count; `dst` checks the counts afterwards (`check`).  For
    steady pipelines (same-sized units step after step) this replaces the count all_gather + padded gather.

    On CUDA the collective runs on its own stream with two send/receive buffer sets, so the gather of step i
    overlaps the index build of step i+1: `submit` only enqueues, `wait_previous` makes the compute stream wait
    for the gather of the step before (whose buffers are about to be reused), `wait_all` for everything.
    All collectives are issued by the calling thread, in program order -- the same order on every rank."""

    def __init__(self, capacity, cols, device, group=None, dst=0):
        self.cols, self.group, self.dst = int(cols), group, dst
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        # every rank must send a block of the SAME size: agree on the largest requested capacity (one collective,
        # at construction time)
        c = torch.tensor([int(capacity)], dtype=torch.int64, device=self.device)
        dist.all_reduce(c, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(c.item())
        self.cuda = self.device.type == "cuda"
        nbuf = 2 if self.cuda else 1
        self.send = [torch.zeros((self.cap + 1, self.cols), dtype=torch.int64, device=device) for _ in range(nbuf)]
        self.recv = [[torch.empty_like(self.send[0]) for _ in range(self.world)] if self.rank == dst else None for _ in range(nbuf)]
        self.step = 0
        self.last = 0
        if self.cuda:
            self.comm = torch.cuda.Stream(device=device)
            self.done = [None] * nbuf

    def close(self):
        pass

    def drain(self):
        pass

    def next_send(self):
        """The send block of the coming gather, for callers that fill it themselves (row 0 = count, rows 1.. = data,
        e.g. rv_result_pack_device) and then call submit()."""
        return self.send[self.step % len(self.send)]

    def gather(self, rows):
        send = self.next_send()
        k = rows.shape[0]
        send[0, 0] = k                            # device-side write, no sync
        m = min(k, self.cap)
        if m:
            send[1:m + 1] = rows[:m]
        self.submit()

    def submit(self):
        b = self.step % len(self.send)
        send = self.send[b]
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ready)
                dist.gather(send, self.recv[b], dst=self.dst, group=self.group)
                self.done[b] = torch.cuda.Event()
                self.done[b].record(self.comm)
        else:
            dist.gather(send, self.recv[b], dst=self.dst, group=self.group)
        self.last = b
        self.step += 1

    def wait_previous(self):
        """The current stream waits for the gather issued one step before the last one (buffer about to be reused)."""
        if self.cuda and self.step >= 2:
            ev = self.done[(self.step - 2) % len(self.send)]
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def wait_all(self):
        if self.cuda:
            for ev in self.done:
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)

    def check(self):
        """On dst: per-rank row tensors of the LAST gather (views into its receive buffers); raises if a rank had
        more rows than the capacity."""
        if self.cuda:
            self.comm.synchronize()
        if self.rank != self.dst:
            return None
        out = []
        recv = self.recv[self.last]
        for r in range(self.world):
            k = int(recv[r][0, 0].item())
            if k > self.cap:
                raise OverflowError("rank %d produced %d rows, gather capacity %d" % (r, k, self.cap))
            out.append(recv[r][1:k + 1])
        return out


class PeerGather(object):
    """One-sided variant of FixedGather for ranks that share one box: `dst` allocates `depth` rings of one
    fixed-capacity block per rank in ITS device memory, every other rank maps that allocation (CUDA IPC, C-ABI
    rv_peer_*), and rv_result_pack_device(handle, slot(), cap) writes a rank's block straight into dst's HBM over
    NVLink / NVSwitch.  Nothing on the data path is a collective: no rendezvous between ranks, no NCCL launch, no
    host work besides the pack launch itself -- ranks only meet in the constructor, in check() and in close().

    A block written at step s is overwritten at step s + depth: dst has to check() at least every `depth` steps if
    it wants every step's rows (a steady pipeline that only consumes the last step, like bench.py, never waits).
    Construction is collective; it raises RuntimeError on EVERY rank when the mapping failed on any of them, so
    callers can fall back to FixedGather together."""

    def __init__(self, capacity, cols, lib, device, group=None, dst=0, depth=2):
        import ctypes

        import numpy as np
        self._ct, self._np = ctypes, np
        self.lib, self.group, self.dst, self.depth, self.cols = lib, group, dst, int(depth), int(cols)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device(device)
        c = torch.tensor([int(capacity)], dtype=torch.int64, device=self.device)
        dist.all_reduce(c, op=dist.ReduceOp.MAX, group=group)
        self.cap = int(c.item())
        self.block = ((self.cap + 1) * self.cols * 8 + 255) // 256 * 256   # bytes per (ring, rank) block
        self.base = None
        self.step = 0
        handle = torch.zeros(64, dtype=torch.uint8)
        ok = 1
        if self.rank == dst:
            p = ctypes.c_void_p()
            hb = (ctypes.c_uint8 * 64)()
            if lib.rv_peer_alloc(self.block * self.world * self.depth, ctypes.byref(p), hb) == 0:
                self.base = p.value
                handle = torch.tensor(list(hb), dtype=torch.uint8)
            else:
                ok = 0
        handle = handle.to(self.device)
        dist.broadcast(handle, src=dst, group=group)
        if self.rank != dst:
            hb = (ctypes.c_uint8 * 64)(*handle.cpu().tolist())
            p = ctypes.c_void_p()
            if any(hb) and lib.rv_peer_open(hb, ctypes.byref(p)) == 0:
                self.base = p.value
            else:
                ok = 0
        flag = torch.tensor([ok], dtype=torch.int64, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            err = lib.rv_last_error().decode() if not ok else "another rank failed"
            self._release()
            raise RuntimeError("peer blocks unavailable: " + err)

    def slot(self):
        """Device address (in dst's memory) of this rank's block for the coming step."""
        return self.base + ((self.step % self.depth) * self.world + self.rank) * self.block

    def advance(self):
        self.step += 1

    def check(self, expect_seq=None):
        """Collective.  Call after the device work of the last step is complete on the calling rank
        (stream / device synchronised): waits for every rank, then returns on dst the per-rank (rows, seq) of the
        LAST step -- int64 numpy arrays [k, cols] copied out of the ring -- and None elsewhere.  Raises on dst if
        a rank had more rows than the capacity, or if `expect_seq` is given and a block carries another sequence
        number (= it was not written by that rank's latest pack)."""
        np, ctypes = self._np, self._ct
        dist.barrier(group=self.group)
        if self.rank != self.dst or self.step == 0:
            return None
        ring = (self.step - 1) % self.depth
        out = []
        for r in range(self.world):
            blk = np.empty(self.block // 8, dtype=np.int64)
            src = self.base + (ring * self.world + r) * self.block
            if self.lib.rv_peer_read(ctypes.c_void_p(src), blk.ctypes.data, self.block) != 0:
                raise RuntimeError(self.lib.rv_last_error().decode())
            k, seq = int(blk[0]), int(blk[1])
            if k > self.cap:
                raise OverflowError("rank %d produced %d rows, gather capacity %d" % (r, k, self.cap))
            if expect_seq is not None and seq != expect_seq:
                raise RuntimeError("rank %d: block carries pack #%d, expected #%d" % (r, seq, expect_seq))
            out.append((blk[self.cols:self.cols * (k + 1)].reshape(k, self.cols), seq))
        return out

    def _release(self):
        if self.base is not None:
            if self.rank == self.dst:
                self.lib.rv_peer_free(self._ct.c_void_p(self.base))
            else:
                self.lib.rv_peer_close(self._ct.c_void_p(self.base))
            self.base = None

    def close(self):
        """Collective: the mappings go first, then dst frees the allocation."""
        if self.rank != self.dst:
            self._release()
        dist.barrier(group=self.group)
        if self.rank == self.dst:
            self._release()


def fwd_rc_units(T, nsep):
    """The two index builds `finish` / `transform` need for one reference + contigs pair (transformold.py:142-143,
    transform.py:255,271): the text as it is, and with the second sample reverse-complemented -- two independent units,
    i.e. a clean 2-GPU job for anchor_units (BASELINE configs[3])."""
    return [(T, nsep, 2, 0), (T, nsep, 2, 1)]


def anchor_units(units, minl=20, minn=2, group=None, device=None, lib=None):
    """Index build + MUM sweep of independent units, sharded over the ranks of `group`.

    units: list of (T uint8 array, nsep int64 array, nsamples[, rc]) -- e.g. the sub-intervals of a recursion frontier
    rebuilt as independent indexes, the jobs of `--order=sequential --chunksize`, or the forward / reverse-complement
    pair of `finish` / `transform` (rc = 1: construct(rc=1), the second sample is reverse-complemented on the device first,
    interface.c:168-172; see fwd_rc_units).  Every rank builds the units `partition` assigns to it on its own GPU; the MUM records are
    gathered to rank 0 (two tagged variable-length gathers: record rows, member rows).  Returns on rank 0 a list with
    one entry per unit -- pair rows (l, a, b) for two samples; for more samples the tuple (hdr, members) exactly as
    rv_mums_multi_fetch delivers it: header rows (l, n, first) and the (sample, position) member rows `first` indexes
    -- and None on the other ranks.
    `lib` is the C-ABI library object (default: the CUDA library; tests inject the emulated one)."""
    import ctypes

    import numpy as np

    from . import _native
    L = lib if lib is not None else _native.lib()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = partition([len(u[0]) for u in units], world)[rank]
    h = ctypes.c_void_p()
    _native.check(L, L.rv_index_create(ctypes.byref(h), None))
    blocks, mblocks = [], []
    try:
        for uid in mine:
            T, nsep, ns = units[uid][:3]
            rc = int(units[uid][3]) if len(units[uid]) > 3 else 0
            T = np.ascontiguousarray(T, dtype=np.uint8)
            nsep = np.ascontiguousarray(nsep, dtype=np.int64)
            _native.check(L, L.rv_build(h, T.ctypes.data, len(T), nsep.ctypes.data if len(nsep) else None, int(ns), rc))
            if ns == 2:
                c = ctypes.c_int64()
                # pair rows as getmums reports them for the root index (reveal.c:55-116): with rc the second coordinate is already
                # mapped back to the forward strand (reveal.c:98-100)
                _native.check(L, L.rv_mums_pair_count(h, int(minl), 0 if rc else 1, ctypes.byref(c)))
                rows = np.empty((c.value, 3), dtype=np.int64)
                _native.check(L, L.rv_mums_pair_fetch(h, rows.ctypes.data, c.value))
            else:
                nr, nm = ctypes.c_int64(), ctypes.c_int64()
                _native.check(L, L.rv_mums_multi_count(h, int(minl), int(minn), ctypes.byref(nr), ctypes.byref(nm)))
                rows = np.empty((nr.value, 3), dtype=np.int64)
                mem = np.empty((nm.value, 2), dtype=np.int64)
                _native.check(L, L.rv_mums_multi_fetch(h, rows.ctypes.data, nr.value, mem.ctypes.data, nm.value))
                # the member rows travel too (tagged like the records): `first` of a header row indexes the unit's own
                # member rows, whose order a tagged gather keeps
                mblocks.append(np.concatenate([np.full((len(mem), 1), uid, dtype=np.int64), mem], axis=1))
            # tag every row with its unit so that one gather carries all units of the rank
            tagged = np.concatenate([np.full((len(rows), 1), uid, dtype=np.int64), rows], axis=1)
            blocks.append(tagged)
    finally:
        L.rv_index_free(h)
    any_multi = any(u[2] != 2 for u in units)  # known to every rank: the second gather is collective

    def collect(parts_local, cols):
        local = np.concatenate(parts_local, axis=0) if parts_local else np.zeros((0, cols), dtype=np.int64)
        t = torch.from_numpy(local)
        if device is not None:
            t = t.to(device)
        if world == 1:
            return t.cpu().numpy()
        parts = gather_rows(t, dst=0, group=group)
        return torch.cat(parts, dim=0).cpu().numpy() if rank == 0 else None

    allrows = collect(blocks, 4)
    allmem = collect(mblocks, 3) if any_multi else None
    if rank != 0:
        return None
    out = []
    for uid, u in enumerate(units):
        rows = allrows[allrows[:, 0] == uid][:, 1:]
        out.append(rows if u[2] == 2 else (rows, allmem[allmem[:, 0] == uid][:, 1:]))
    return out
