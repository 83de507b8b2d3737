#!/usr/bin/env python
"""bench.py -- aligned bases/sec of the rem anchor phase (index build + MUM sweep).

A "step" is one pass of the hot path over one batch of synthetic input:
construct() = suffix array + inverse + barrier-aware LCP (+ sample array), then
the SA/LCP sweep that emits the (multi-)MUMs at the reference's `rem` defaults
(-m 20 -n 2).  Default workload = BASELINE.json configs[1]: 2 synthetic 5 Mbp
genomes (1 % SNP, 0.1 % indel).  "bases" = characters of the concatenated text
(genome bases + one sentinel per sequence) that the step anchors.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4] [--impl reference]

N > 1 (torchrun, one rank per GPU): every rank builds its own independent index
(the rem recursion / --chunksize jobs shard as independent index builds, SURVEY
8e), MUM records are gathered to rank 0 with NCCL, nothing else crosses GPUs.

`--impl reference` times the reference's own CPU implementation of the same step
(its unmodified extension compiled into oracle/_ref: divsufsort + compute_lcp +
getmums/getmultimums through its Python API) on the host, single-threaded like
the reference is.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (genomes, length, description)
    "c2": (2, 5_000_000, "2 synthetic 5 Mbp genomes (1% SNP, 0.1% indel), rem -m 20 -n 2 (BASELINE configs[1])"),
    "c3": (5, 5_000_000, "5 synthetic 5 Mbp genomes, simultaneous rem anchor (BASELINE configs[2])"),
    "c4": (2, 100_000_000, "2 synthetic 100 Mbp genomes (BASELINE configs[3], root build + sweep)"),
    "tiny": (2, 200_000, "2 synthetic 200 kbp genomes (plumbing)"),
    # repeat-bearing inputs (SURVEY 8d "realistic repeat-bearing check"): the doubling rounds and the LCP fallback run here
    "real": (2, 0, "reference fixtures tests/123a.fa + 123b.fa (3 + 3 Aspergillus niger contigs, n = 10 754 553; committed compact copy tests/golden/real/asp_niger.npz), rem -m 20 -n 2"),
    "repeats": (2, 5_000_000, "2 synthetic 5 Mbp genomes on a repeat-bearing ancestor (interspersed families, tandem arrays, segmental duplications, N runs), rem -m 20 -n 2"),
    "graph": (2, 10_000_000, "graph-like text: 2 samples of 10 Mbp cut into 5e5 contigs each (one '$' per node, SURVEY 8 C5)"),
}
MINL, MINN = 20, 2
METRIC = "aligned bases/sec (rem anchor phase: index build + MUM sweep)"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_traffic(kernel, workload):
    """dram bytes per launch of `kernel` from the newest committed ncu capture of this workload (profiles/*/traffic.json)."""
    import glob
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "*", "traffic.json")), reverse=True):
        try:
            d = json.load(open(path))
            if d.get("workload") == workload and kernel in d:
                return int(d[kernel]["dram_read_bytes"]) + int(d[kernel]["dram_write_bytes"]), os.path.relpath(path, ROOT)
        except Exception:
            pass
    return None, None


def peaks():
    """HBM roofline denominator: the driver-measured copy bandwidth when MEASURED_PEAKS.json exists, else the fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))

            def find(obj):
                if isinstance(obj, dict):
                    for k in ("hbm_gbs", "hbm_gbps", "hbm_GBs"):
                        if isinstance(obj.get(k), (int, float)):
                            return float(obj[k])
                    for k, v in obj.items():  # tolerate other spellings / nesting: first numeric "hbm*" entry in GB/s range
                        if isinstance(v, (int, float)) and "hbm" in k.lower() and 500 < float(v) < 20000:
                            return float(v)
                    for v in obj.values():
                        r = find(v)
                        if r:
                            return r
                return None

            v = find(d)
            if v:
                return v, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons during the timed region: through NVML (nvidia_ml_py, a sample every 5 ms --
    the timed region of a default run is a few milliseconds long) or, if that is not usable, through nvidia-smi."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}  # nvml.h

    def __init__(self, gpu):
        threading.Thread.__init__(self, daemon=True)
        self.gpu, self.rows, self.stop_flag, self.source = gpu, [], False, "nvidia-smi"
        self.sm, self.mx, self.mask = [], [], 0

    def _nvml_loop(self):
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
        reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
        self.mx.append(int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)))
        self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))   # raises here if NVML is unusable
        self.source = "nvml"
        while not self.stop_flag:
            self.sm.append(int(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
            self.mask |= int(reasons(h))
            time.sleep(0.005)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            self.source = "nvidia-smi"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        self.join(timeout=6)
        if self.source == "nvml":
            sm, mx = self.sm, self.mx
            reasons = [nm for nm in self.NAMES if self.mask & self.BITS[nm]]
            n = len(sm)
        else:
            sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
            mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
            reasons = [nm for i, nm in enumerate(self.NAMES) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
            n = len(self.rows)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": n,
                "source": self.source}


def make_workload(name, seed):
    from reveal_b200 import synth
    g, length, desc = WORKLOADS[name]
    if name == "real":
        a, b, _ = synth.load_packed_fixture(os.path.join(ROOT, "tests", "golden", "real", "asp_niger.npz"))
        T, nsep = synth.concat([a, b])
        return T, nsep, 2, desc
    if name == "repeats":
        T, nsep, ns = synth.repeat_workload(g, length, seed=seed)
    elif name == "graph":
        T, nsep, ns = synth.graph_like_workload(500_000, 20, seed=seed)
    else:
        T, nsep, ns = synth.workload(g, length, seed=seed)
    return T, nsep, ns, desc


def config_block(desc, bases_per_unit, mums, units):
    """The `config` object of the JSON line: the SAME keys in both arms (the driver compares them key by key)."""
    return {"workload": desc, "bases_per_step_per_gpu": int(bases_per_unit), "mums_per_step": int(mums), "minl": MINL, "minn": MINN,
            "units": int(units), "l2": "flushed between timed steps (256 MiB write)",
            "sharding": "single unit" if units == 1 else "one independent index build per unit (GPU rank / host process); MUM records to rank 0"}


# ------------------------------------------------------------------------------------------
_REF_UNIT = {}


def _ref_unit_step(job):
    """One index build + sweep of the unmodified reference extension on one unit (worker process or in-process).
    job = (workload, seed, fraction of every genome to use)."""
    import oracle.ref as R
    if job not in _REF_UNIT:
        name, seed, frac = job
        T, nsep, ns, _ = make_workload(name, seed=seed)
        n = len(T)
        bounds = [0] + [int(x) + 1 for x in nsep] + [n]
        seqs = [T[bounds[k]:bounds[k + 1] - 1].tobytes().decode("ascii") for k in range(ns)]
        if frac < 1.0:
            seqs = [s[:max(1000, int(len(s) * frac))] for s in seqs]
        _REF_UNIT.clear()
        _REF_UNIT[job] = (seqs, ns)
    seqs, ns = _REF_UNIT[job]
    t0 = time.perf_counter()   # the timed region is the user's call sequence, as in the B200 arm's e2e leg
    idx = R.module(32).index()
    for k, s in enumerate(seqs):
        idx.addsample("g%d" % k)
        idx.addsequence(s)
    idx.construct()
    mums = idx.getmums(MINL) if ns == 2 else idx.getmultimums(minlength=MINL, minn=MINN)
    return time.perf_counter() - t0, len(mums), sum(len(s) + 1 for s in seqs)


def run_reference(args, rank, world):
    """The reference's own CPU path (unmodified extension in oracle/_ref), rank 0 only.  At --gpus N the job is N
    independent units (what the N ranks of the B200 arm build): the reference gets one host process per unit, all
    running at once -- every host thread this path can use (one index build is single-threaded in the reference)."""
    if rank != 0:
        return
    import oracle.ref as R
    if not R.available():
        emit({"impl": "reference", "unavailable": "oracle/_ref not built (reference tree absent at build time)"})
        return
    units = max(1, args.gpus)
    cores = min(units, os.cpu_count() or 1)
    desc = WORKLOADS[args.workload][2]
    pool = None
    if units > 1:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(cores)

    def step(frac):
        jobs = [(args.workload, 1 + r, frac) for r in range(units)]
        t0 = time.perf_counter()
        res = pool.map(_ref_unit_step, jobs, chunksize=1) if pool else [_ref_unit_step(jobs[0])]
        wall = time.perf_counter() - t0
        # one unit: its own timed region (index() .. getmums()); several units: wall time of the concurrent batch
        return (res[0][0] if not pool else wall), res[0][1], sum(r[2] for r in res), res[0][2]

    # bounded sample: if the full workload would not finish K + W steps in about four minutes, every step uses a prefix of
    # each genome (throughput per base is what is reported)
    frac = 1.0
    dt, nm, bases, bases0 = step(frac)
    full_bases0, full_nm = bases0, nm   # unit 0 of the workload the config names; a bounded sample of it may be timed below
    budget = float(os.environ.get("RV_REF_BUDGET_S", "240"))
    planned = args.steps + max(args.warmup, 1) - 1
    if dt * planned > budget:
        frac = max(0.02, budget / (dt * planned))
        dt, nm, bases, _ = step(frac)
    for _ in range(max(args.warmup, 1) - 1):
        step(frac)
    tot = 0.0
    for _ in range(args.steps):
        dt, nm, bases, _ = step(frac)
        tot += dt
    if pool:
        pool.close()
    value = bases * args.steps / tot
    sample = "the full workload per step" if frac >= 1.0 else "the first %.0f %% of every genome per step" % (100 * frac)
    sample += " (index() + addsample/addsequence + construct + getmums%s through the reference extension's Python API" % ("" if WORKLOADS[args.workload][0] == 2 else "/getmultimums")
    sample += "; %d independent units in %d processes at once)" % (units, cores) if units > 1 else ")"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": config_block(desc, full_bases0, full_nm, units),
            "cpu_baseline": {"value": value, "unit": "bases/s", "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": "bases/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------
class DevArray:
    """Exposes a raw device pointer to torch through __cuda_array_interface__."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from reveal_b200 import _native
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = _native.lib()
    _native.check(L, L.rv_set_device(local_rank))
    if world > 1:
        import datetime
        # a short collective timeout: a mismatched or stuck gather must abort the run, not hold 8 GPUs for minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=90))

    T, nsep, ns, desc = make_workload(args.workload, seed=1 + rank)
    n = len(T)
    nsep = np.ascontiguousarray(nsep, dtype=np.int64)
    hT = torch.from_numpy(T).pin_memory()
    dT = hT.to(dev)
    stream = torch.cuda.Stream(device=dev)
    h = ctypes.c_void_p()
    _native.check(L, L.rv_index_create(ctypes.byref(h), ctypes.c_void_p(stream.cuda_stream)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    cnt = ctypes.c_int64()
    nmem = ctypes.c_int64()
    host_rows = torch.empty((1 << 22, 3), dtype=torch.int64).pin_memory()

    def sweep():
        if ns == 2:
            _native.check(L, L.rv_mums_pair_count(h, MINL, 1, ctypes.byref(cnt)))
        else:
            _native.check(L, L.rv_mums_multi_count(h, MINL, MINN, ctypes.byref(cnt), ctypes.byref(nmem)))
        return cnt.value

    def result_rows():
        """Rows of the last sweep on the host (pair rows, or the hdr rows of a multi sweep)."""
        k = cnt.value
        out = np.empty((k, 3), np.int64)
        if ns == 2:
            _native.check(L, L.rv_mums_pair_fetch(h, out.ctypes.data, k))
        else:
            mem = np.empty((nmem.value, 2), np.int64)
            _native.check(L, L.rv_mums_multi_fetch(h, out.ctypes.data, k, mem.ctypes.data, nmem.value))
        return out

    gatherer = {}

    def gather_results():
        """N > 1: MUM records of every rank to rank 0, no host synchronisation.  Preferred transport: mapped peer
        blocks (reveal_b200.shard.PeerGather) -- the pack kernel stores each rank's rows straight into rank 0's HBM
        over NVLink / NVSwitch, nothing on the data path is a collective.  If the box cannot map peer memory every
        rank falls back together to one NCCL gather of fixed-capacity blocks per step (shard.FixedGather)."""
        if world == 1:
            return
        from reveal_b200 import shard
        with torch.cuda.stream(stream):  # same stream as the sweep kernels that produced the rows
            if "g" not in gatherer:
                cap = max(4096, 2 * cnt.value)
                try:
                    if os.environ.get("RV_BENCH_GATHER", "peer") != "peer":
                        raise RuntimeError("RV_BENCH_GATHER asks for the NCCL gather")
                    gatherer["g"] = shard.PeerGather(cap, 3, L, dev, depth=2)
                    gatherer["kind"] = "peer"
                except RuntimeError as e:
                    if rank == 0:
                        print("[bench] peer blocks not used (%s): NCCL gather" % e, file=sys.stderr)
                    gatherer["g"] = shard.FixedGather(cap, 3, dev)
                    gatherer["kind"] = "nccl"
            g = gatherer["g"]
            if gatherer["kind"] == "peer":
                _native.check(L, L.rv_result_pack_device(h, ctypes.c_void_p(g.slot()), g.cap))  # count + rows into rank 0's ring
                g.advance()
                return
            send = g.next_send()
            _native.check(L, L.rv_result_pack_device(h, ctypes.c_void_p(send.data_ptr()), g.cap))  # count + rows, async on `stream`
            g.submit()            # enqueued on the communication stream: overlaps the next step's index build
            g.wait_previous()     # ... but a step does not end before the gather of the step before it has landed

    gather_evs = []

    def step_resident():
        _native.check(L, L.rv_build_device(h, ctypes.c_void_p(dT.data_ptr()), n, nsep.ctypes.data, ns, 0))
        k = sweep()
        if world > 1:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(stream)
            gather_results()
            g1.record(stream)
            gather_evs.append((g0, g1))
        return k

    # the drop-in surface: what a user of the reference calls (reveal/rem.py:560-575 feeds the index like this)
    from reveal_b200 import reveallib
    bounds = [0] + [int(x) + 1 for x in nsep] + [n]
    seqs = [T[bounds[k]:bounds[k + 1] - 1].tobytes().decode("ascii") for k in range(ns)]

    def step_api():
        """index() -> addsample/addsequence -> construct() -> getmums()/getmultimums() with its Python list: host strings in,
        Python objects out.  A new index object per step, like one `rem` run."""
        idx = reveallib.index()
        for k, sq in enumerate(seqs):
            idx.addsample("g%d" % k)
            idx.addsequence(sq)
        idx.construct()
        mums = idx.getmums(MINL) if ns == 2 else idx.getmultimums(minlength=MINL, minn=MINN)
        k = len(mums)
        d2h = k * 24 if ns == 2 else k * 24 + sum(len(m[2]) for m in mums) * 16
        return k, d2h

    def step_e2e():
        _native.check(L, L.rv_build(h, ctypes.c_void_p(hT.data_ptr()), n, nsep.ctypes.data, ns, 0))
        k = sweep()
        if ns == 2:
            _native.check(L, L.rv_mums_pair_fetch(h, ctypes.c_void_p(host_rows.data_ptr()), min(k, host_rows.shape[0])))
            d2h = k * 24
        else:
            hdr = np.empty((k, 3), np.int64)
            mem = np.empty((nmem.value, 2), np.int64)
            _native.check(L, L.rv_mums_multi_fetch(h, hdr.ctypes.data, k, mem.ctypes.data, nmem.value))
            d2h = k * 24 + nmem.value * 16
        return k, d2h

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t_cold = time.perf_counter()
    k_cold, _ = step_api()      # the very first call of the process through the drop-in: CUDA context, module load, first allocations
    torch.cuda.synchronize()
    cold_first_call_s = time.perf_counter() - t_cold
    # Warm-up steps with every kernel group bracketed by event pairs: they name the dominant kernel (largest share of the step).
    # The event pairs cost a few microseconds of a step themselves (C2: 0.94 ms with all of them, 0.90 ms without), so the
    # TIMED steps bracket only that kernel -- its launch durations for `roofline` are measured live inside the timed region, as
    # the contract asks -- and the table of the other kernels comes from extra, untimed steps afterwards.
    _native.check(L, L.rv_profile(h, 1))
    for _ in range(max(args.warmup, 3)):
        step_resident()
    profw = _native.KernelProfile()
    _native.check(L, L.rv_get_profile(h, ctypes.byref(profw)))
    top_slot = max(range(len(profw.SLOTS)), key=lambda k: profw.ms[k] if profw.launches[k] else -1.0)
    # ---- timed: device-resident input -----------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    _native.check(L, L.rv_profile(h, 1 | (2 << top_slot)))   # mask: only the dominant kernel's slot
    prof0 = _native.KernelProfile()
    _native.check(L, L.rv_get_profile(h, ctypes.byref(prof0)))
    barrier()
    evs = []
    for _ in range(args.steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)  # evict L2 between timed iterations (inputs are smaller than L2); on the library's stream, so it is ordered before e0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        nm = step_resident()
        if world > 1 and _ == args.steps - 1 and gatherer["kind"] == "nccl":
            with torch.cuda.stream(stream):
                gatherer["g"].wait_all()  # the last step's gather is inside the timed region
            # (peer blocks: the pack kernel's stores ARE the transfer; they are complete when e1 is reached)
        e1.record(stream)
        evs.append((e0, e1))
    barrier()
    gather_check = None
    if world > 1:
        # rank 0: every rank's rows of the last step arrived.  A failed check is REPORTED in the JSON line ("gather_check") and on
        # stderr instead of raised: rank 0 dying here would leave the other ranks waiting in the collectives that follow.
        try:
            if gatherer["kind"] == "peer":
                mine = result_rows()  # every rank: (row count, wrapping sum of its rows) of the last step, to compare on rank 0
                sig = torch.tensor([mine.shape[0], int(mine.sum(dtype=np.int64))], dtype=torch.int64, device=dev)
                sigs = [torch.zeros_like(sig) for _ in range(world)]
                dist.all_gather(sigs, sig)
                parts = gatherer["g"].check(expect_seq=gatherer["g"].step)   # barrier inside; raises on overflow / stale blocks (rank 0)
                if rank == 0:
                    assert len(parts) == world and all(len(rows) > 0 for rows, _ in parts), "a rank delivered no rows"
                    assert np.array_equal(parts[0][0], mine), "rank 0's own block differs from its sweep result"
                    for r in range(world):
                        got = [parts[r][0].shape[0], int(parts[r][0].sum(dtype=np.int64))]
                        assert got == sigs[r].tolist(), "rank %d: rows in rank 0's ring differ from what the rank produced" % r
            else:
                parts = gatherer["g"].check()
                if rank == 0:
                    assert len(parts) == world and all(len(x) > 0 for x in parts), "a rank delivered no rows"
            gather_check = "ok: rows of every rank verified on rank 0 after the timed region"
        except Exception as e:
            gather_check = "FAILED: %s: %s" % (type(e).__name__, e)
            print("[bench] rank %d: gather check %s" % (rank, gather_check), file=sys.stderr)
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    gather_ms = sum(a.elapsed_time(b) for a, b in gather_evs[-args.steps:]) / args.steps if gather_evs else 0.0
    prof = _native.KernelProfile()
    _native.check(L, L.rv_get_profile(h, ctypes.byref(prof)))
    times = _native.Times()
    _native.check(L, L.rv_get_times(h, ctypes.byref(times)))
    # the other kernels: extra steps after the timed region, every group bracketed
    _native.check(L, L.rv_profile(h, 1))
    extra_steps = min(args.steps, 3)
    extra_evs = []
    for _ in range(extra_steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _native.check(L, L.rv_build_device(h, ctypes.c_void_p(dT.data_ptr()), n, nsep.ctypes.data, ns, 0))
        sweep()   # (no gather: rank 0 may still be reading the blocks of the timed steps)
        e1.record(stream)
        extra_evs.append((e0, e1))
    torch.cuda.synchronize()
    extra_ms = sum(a.elapsed_time(b) for a, b in extra_evs)
    prof_all = _native.KernelProfile()
    _native.check(L, L.rv_get_profile(h, ctypes.byref(prof_all)))
    _native.check(L, L.rv_profile(h, 0))
    # ---- timed: end to end from pinned host memory through the C-ABI (secondary: e2e_cabi) -----------------------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        step_e2e()
    barrier()
    cabi_s = time.perf_counter() - t0
    # ---- timed: end to end through the drop-in extension (the headline e2e): host strings in, Python list out --------------
    for _ in range(2):
        step_api()
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for _ in range(args.steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        k, d2h = step_api()
    barrier()
    e2e_s = time.perf_counter() - t0
    assert k == nm == k_cold, "the drop-in returned %d MUMs, the C-ABI sweep %d" % (k, nm)
    clocks = sampler.summary() if sampler else None

    tmax = torch.tensor([dev_ms, e2e_s * 1e3, cabi_s * 1e3], dtype=torch.float64, device=dev)
    ntot = torch.tensor([float(n)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(ntot, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, cabi_ms_max = float(tmax[0]), float(tmax[1]), float(tmax[2])
    total_bases = float(ntot[0])

    sharded = None
    if world > 1 and (args.workload in ("c2", "tiny") or args.rem):
        sharded = rem_sharded(args.workload, rank, world, dev)
    if rank == 0:
        peak, peak_src = peaks()
        value = total_bases * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_bases * args.steps / (e2e_ms_max * 1e-3)
        launches = int(prof.launches_total - prof0.launches_total)
        def rows_of(pr, total_ms, where):
            out = []
            for k, name in enumerate(pr.SLOTS):
                if pr.launches[k] and pr.ms[k] > 0:
                    gbs = (pr.bytes[k] / 1e9) / (pr.ms[k] * 1e-3)
                    out.append({"kernel": name, "launches": int(pr.launches[k]), "avg_launch_ms": pr.ms[k] / pr.launches[k],
                                "achieved": gbs, "frac": gbs / peak, "share_of_step": pr.ms[k] / total_ms,
                                "algorithmic_bytes_per_launch": pr.bytes[k] / pr.launches[k], "measured": where})
                    if name == "text passes":  # (10.7 us when ncu runs it alone: profiles/)
                        out[-1]["note"] = "side stream, beside the digit passes of the sort: elapsed time under contention, not on the critical path"

            out.sort(key=lambda d: -d["share_of_step"])
            return out
        timed = rows_of(prof, dev_ms, "inside the timed region (the only kernel bracketed there)")
        top = timed[0] if timed else {}
        kernels = timed + [r for r in rows_of(prof_all, extra_ms, "%d extra steps after the timed region, every kernel group bracketed" % extra_steps)
                           if r["kernel"] != top.get("kernel")]
        traffic, traffic_src = measured_traffic(top.get("kernel"), args.workload)
        bpb = 35 if ns == 2 else 39
        line = {"metric": METRIC, "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32", "data": "synthetic",
                "config": config_block(desc, n, nm, world),
                "transport": ("mapped peer blocks (NVLink stores, no collective)" if gatherer.get("kind") == "peer" else "one NCCL gather per step") if world > 1 else None,
                "e2e": {"value": e2e_value, "unit": "bases/s", "h2d_bytes_per_step": int(n + 8 * len(nsep)), "d2h_bytes_per_step": int(d2h + 32),
                        "ms_per_step": e2e_ms_max / args.steps,
                        "what": "reveallib.index(): addsample/addsequence (host str) -> construct() -> getmums()/getmultimums() incl. its Python list, a new index object per step",
                        "cold_first_call_ms": cold_first_call_s * 1e3},
                "e2e_cabi": {"value": total_bases * args.steps / (cabi_ms_max * 1e-3), "unit": "bases/s", "ms_per_step": cabi_ms_max / args.steps,
                             "what": "rv_build from pinned host memory + sweep + rows into a pinned buffer on a warm handle (no Python objects)"},
                "gpu_launches": launches, "gather_ms_per_step": gather_ms, "gather_check": gather_check,
                # the dominant kernel = largest measured share of the step; algorithmic bytes per slot: include/reveal_b200.h, DESIGN.md
                "roofline": {"bound": "hbm", "kernel": top.get("kernel"), "achieved": top.get("achieved"), "peak": peak, "unit": "GB/s",
                             "frac": top.get("frac"), "peak_source": peak_src, "traffic": traffic, "traffic_source": traffic_src,
                             "algorithmic_bytes_per_launch": top.get("algorithmic_bytes_per_launch"), "launches": top.get("launches"),
                             "avg_launch_ms": top.get("avg_launch_ms"), "share_of_step": top.get("share_of_step"),
                             "kernels": kernels,
                             "path": {"algorithmic_bytes_per_base": bpb, "achieved": bpb * n * args.steps / (dev_ms * 1e-3) / 1e9,
                                      "frac": bpb * n * args.steps / (dev_ms * 1e-3) / 1e9 / peak, "unit": "GB/s"}},
                "phases_ms_last_step": {k: (round(v, 4) if isinstance(v, float) else v) for k, v in times.as_dict().items()},
                "clocks": clocks}
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(T, nsep, ns)
            line["rem_configs0"] = rem_configs0()
            if args.workload in ("c2", "tiny") or args.rem:
                line["rem_end_to_end"] = rem_end_to_end(args.workload)
        if world > 1:
            line["rem_sharded"] = sharded
        emit(line)
    L.rv_index_free(h)
    if world > 1:
        if "g" in gatherer:
            gatherer["g"].close()
        dist.destroy_process_group()


def cpu_baseline(T, nsep, ns):
    """The reference's CPU code on this host, one core (it is single-threaded on this path)."""
    import oracle.ref as R
    n = len(T)
    if R.available():
        SA, SAi, LCP, (t_sa, t_isa, t_lcp) = R.raw_build(T)
        import oracle.port as P
        t0 = time.perf_counter()
        if ns == 2:
            k = len(P.getmums(T, SA, LCP, int(nsep[0]), MINL, rem=True))
        else:
            SO = P.build_so(nsep, ns, n)
            k = len(P.getmulti(T, SA, LCP, SO, int(nsep[0]), ns, MINL, MINN)[0])
        t_sw = time.perf_counter() - t0
        tot = t_sa + t_isa + t_lcp + t_sw
        return {"value": n / tot, "unit": "bases/s", "cores": 1, "kind": "reference", "host_cores": os.cpu_count(),
                "sample": "the full workload once: reference divsufsort %.2fs + ISA %.2fs + compute_lcp %.2fs (oracle/_ref objects) + sweep %.2fs (oracle port)" % (t_sa, t_isa, t_lcp, t_sw),
                "mums": int(k)}
    import oracle.port as P
    t0 = time.perf_counter()
    o = P.Index(T, nsep, ns)
    k = len(o.getmums(MINL, rem=True)) if ns == 2 else len(o.getmultimums(MINL, MINN)[0])
    tot = time.perf_counter() - t0
    return {"value": n / tot, "unit": "bases/s", "cores": 1, "kind": "port", "host_cores": os.cpu_count(),
            "sample": "the full workload once through the oracle port (SA-IS + Kasai + sweep)", "mums": int(k)}


def _write_genome_files(workload, seed, tmp):
    """FASTA files of a synthetic workload (one per genome, 100 columns), or of the 1a/1b reference pair for 'c1'."""
    from reveal_b200 import synth
    g, length, _ = WORKLOADS[workload]
    files = []
    for k, seq in enumerate(synth.genomes(g, length, seed=seed)):
        files.append(os.path.join(tmp, "g%d.fa" % k))
        text = seq.tobytes().decode()
        with open(files[-1], "w") as f:
            f.write(">g%d\n" % k)
            f.write("\n".join(text[i:i + 100] for i in range(0, len(text), 100)))
            f.write("\n")
    return files


def _rem_once(files, module, shard=None):
    """One `rem`: FASTA files -> alignment graph (+ prune for more than two genomes); seconds, aligned bases, where the time went."""
    from reveal_b200 import rem
    t0 = time.perf_counter()
    G, idx = rem.align_genomes(rem.rem_args(files), index_module=module, shard=shard)
    t1 = time.perf_counter()
    out = {"align_genomes_s": t1 - t0, "align_stats": module.align_stats() if hasattr(module, "align_stats") else None,
           "shard_stats": rem.align_genomes.last_shard_stats}
    if shard is None or shard[0] == 0:
        if len(G.graph["paths"]) > 2:
            rem.prune_nodes(G, T=idx.T)
        bases, total, nodes = rem.aligned_bases(G, idx)
        out.update({"aligned_bases": bases, "total_bases": total, "aligned_nodes": nodes, "nodes": G.number_of_nodes()})
    out["seconds"] = time.perf_counter() - t0
    return out


def rem_end_to_end(workload):
    """First-class secondary figure: `rem` of the bench workload end to end through the driver (FASTA files -> alignment graph,
    SURVEY 8d(ii)), on the B200 library and -- same driver, same host -- on the reference's own compiled extension (CPU)."""
    import tempfile
    try:
        from reveal_b200 import reveallib
        with tempfile.TemporaryDirectory() as tmp:
            files = _write_genome_files(workload, 1, tmp)
            _rem_once(files, reveallib)   # warm-up of the extension's own handle
            out = {"workload": "rem " + WORKLOADS[workload][2] + ": FASTA files -> alignment graph", "b200": _rem_once(files, reveallib)}
            out["b200"]["aligned_bases_per_s"] = out["b200"]["aligned_bases"] / out["b200"]["seconds"]
            import oracle.ref as R
            if R.available() and WORKLOADS[workload][0] * WORKLOADS[workload][1] <= 12_000_000:
                ref = _rem_once(files, R.module(32))
                ref["aligned_bases_per_s"] = ref["aligned_bases"] / ref["seconds"]
                out["reference_extension_same_driver"] = ref
                out["identical_aligned_bases"] = ref["aligned_bases"] == out["b200"]["aligned_bases"] and ref["nodes"] == out["b200"]["nodes"]
            return out
    except Exception as e:  # a secondary figure must never cost the bench line
        return {"error": "%s: %s" % (type(e).__name__, e)}


def rem_sharded(workload, rank, world, dev):
    """STRONG scaling of one alignment (row N1, BASELINE configs[2]/[4]): the same `rem` on 1 GPU (rank 0 alone) and with its
    recursion sharded over the `world` ranks (reveal_b200.rem.align_genomes(shard=...)): every rank builds the index and runs
    the tree above the cut, the units below it are shared out, rank 0 collects the graph.  Reports what limits it."""
    import tempfile

    import torch
    import torch.distributed as dist
    from reveal_b200 import reveallib
    try:
        with tempfile.TemporaryDirectory() as tmp:
            files = _write_genome_files(workload, 1, tmp)
            one = _rem_once(files, reveallib)     # the whole alignment on 1 GPU: rank 0's run is the one reported, the others' their warm-up
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            mine = _rem_once(files, reveallib, shard=(rank, world))
            torch.cuda.synchronize()
            dist.barrier()
            wall = time.perf_counter() - t0
            tl = torch.tensor([wall, mine["align_stats"]["total_s"], float(mine["align_stats"]["steps"])], dtype=torch.float64, device=dev)
            rows = [torch.zeros_like(tl) for _ in range(world)]
            dist.all_gather(rows, tl)
            if rank != 0:
                return None
            per_rank = [{"rank": r, "wall_s": float(x[0]), "align_s": float(x[1]), "steps": int(x[2])} for r, x in enumerate(rows)]
            n_wall = max(p["wall_s"] for p in per_rank)
            return {"workload": "rem " + WORKLOADS[workload][2] + ": ONE alignment, recursion sharded over %d GPUs" % world, "scaling": "strong",
                    "seconds_1gpu": one["seconds"], "seconds_ngpu": n_wall, "speedup": one["seconds"] / n_wall,
                    "aligned_bases": mine["aligned_bases"], "aligned_bases_per_s": mine["aligned_bases"] / n_wall,
                    "identical_to_1gpu": mine["aligned_bases"] == one["aligned_bases"] and mine["nodes"] == one["nodes"] and mine["aligned_nodes"] == one["aligned_nodes"],
                    "per_rank": per_rank, "rank0_1gpu": one, "rank0_sharded": mine,
                    "limits": "every rank reads the inputs, builds the root index and runs the tree above the cut (replicated, not divided); "
                              "the units below the cut divide; rank 0 then unpickles and merges every rank's part of the graph (serial)"}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}


def rem_configs0():
    """Secondary figure: BASELINE configs[0] -- `rem` of the reference's 1a.fa + 1b.fa -- end to end
    through the REM driver (reveal_b200/rem.py): FASTA files -> alignment graph, aligned bases as the reference
    counts them per second of wall time, on the B200 library and, with the same driver, on the reference's own
    compiled extension (CPU, oracle/_ref).  The recursion is callback-bound, so this moves with the driver."""
    import gzip
    import tempfile
    try:
        from reveal_b200 import rem, reveallib
        inputs = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "rem", "inputs.json.gz")).read())
        with tempfile.TemporaryDirectory() as tmp:
            files = []
            for fn in ("1a.fa", "1b.fa"):
                files.append(os.path.join(tmp, fn))
                with open(files[-1], "w") as f:
                    for name, seq in inputs[fn]:
                        f.write(">%s\n%s\n" % (name, seq))

            def run(module):
                t0 = time.perf_counter()
                G, idx = rem.align_genomes(rem.rem_args(files), index_module=module)
                dt = time.perf_counter() - t0
                bases, total, nodes = rem.aligned_bases(G, idx)
                return {"seconds": dt, "aligned_bases": bases, "total_bases": total, "aligned_nodes": nodes, "aligned_bases_per_s": bases / dt}

            run(reveallib)
            out = {"workload": "rem 1a.fa 1b.fa (-m 20 -n 2), FASTA files -> alignment graph", "b200": run(reveallib)}
            import oracle.ref as R
            if R.available():
                out["reference_extension_same_driver"] = run(R.module(32))
            return out
    except Exception as e:  # a secondary figure must never cost the bench line
        return {"error": "%s: %s" % (type(e).__name__, e)}


_REAL_STDOUT = None


def emit(line):
    """The one JSON line goes to the real stdout; everything else any library prints (NCCL banner ...) was
    redirected to stderr at start-up."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--rem", action="store_true", help="also run `rem` end to end on this workload (default for c2): 1 GPU vs the reference extension, N GPUs sharded")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
