#!/usr/bin/env python
"""Times index build + sweep of ONE build of the CUDA library given by path (kernel tuning experiments: several builds with
different RV_* macros are compared in one gpurun call).  Not the bench.py metric.

usage: variant_bench.py <lib.so> [workload] [steps]"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from reveal_b200 import _native  # noqa: E402


def main():
    path = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "c2"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    L = _native.bind(path)
    T, nsep, ns, desc = bench.make_workload(workload, seed=1)
    n = len(T)
    nsep = np.ascontiguousarray(nsep, dtype=np.int64)
    dT = torch.from_numpy(T).cuda()
    stream = torch.cuda.Stream()
    h = ctypes.c_void_p()
    _native.check(L, L.rv_index_create(ctypes.byref(h), ctypes.c_void_p(stream.cuda_stream)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cnt, nmem = ctypes.c_int64(), ctypes.c_int64()

    def step():
        _native.check(L, L.rv_build_device(h, ctypes.c_void_p(dT.data_ptr()), n, nsep.ctypes.data, ns, 0))
        if ns == 2:
            _native.check(L, L.rv_mums_pair_count(h, 20, 1, ctypes.byref(cnt)))
        else:
            _native.check(L, L.rv_mums_multi_count(h, 20, 2, ctypes.byref(cnt), ctypes.byref(nmem)))

    for _ in range(3):
        step()
    if not os.environ.get("RV_BENCH_NO_PROFILE"):   # (per-kernel event pairs cost a few microseconds of the step themselves)
        _native.check(L, L.rv_profile(h, 1))
    torch.cuda.synchronize()
    ms = 0.0
    for _ in range(steps):
        with torch.cuda.stream(stream):
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
    prof = _native.KernelProfile()
    _native.check(L, L.rv_get_profile(h, ctypes.byref(prof)))
    out = {"lib": os.path.basename(path), "workload": workload, "ms_per_step": ms / steps, "mums": cnt.value}
    for k, name in enumerate(prof.SLOTS):
        if prof.launches[k]:
            out[name] = round(prof.ms[k] / steps, 4)
    print(json.dumps(out))
    L.rv_index_free(h)


if __name__ == "__main__":
    main()
