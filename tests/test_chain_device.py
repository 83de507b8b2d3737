"""rv_chain_batch (csrc/rv_chain.cu): the chaining recurrence of the mumpicker on the device, many lists per launch, against
the host recurrence rv_chain_dp (csrc/ext/chain_dp.h, exposed as reveallib.chain_dp) -- links and scores bit-identical, ties
included.  CPU tier: the emulated kernels; gpu tier: the CUDA library."""
import ctypes

import numpy as np
import pytest

from reveal_b200 import _native, reveallib


def random_lists(rng, nlists, maxm, ties=False):
    lists = []
    for _ in range(nlists):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(1, maxm))
        start = np.sort(rng.integers(0, 300 if ties else 5000, size=(m + 1, 1)), axis=0) + rng.integers(-30, 30, size=(m + 1, k))
        start[0] = -100
        start[m] = 10000
        length = rng.integers(1, 40, size=m + 1)
        length[0] = length[m] = 0
        gain = rng.integers(0, 50, size=m + 1) * (0 if ties else 1)
        lists.append((np.ascontiguousarray(start, np.int64), np.ascontiguousarray(length, np.int64), np.ascontiguousarray(gain, np.int64)))
    return lists


def check(L, lists, wpen, model):
    h = ctypes.c_void_p()
    _native.check(L, L.rv_index_create(ctypes.byref(h), None))
    try:
        row_off = np.zeros(len(lists) + 1, np.int64)
        start_off = np.zeros(len(lists) + 1, np.int64)
        kk = np.zeros(len(lists), np.int32)
        for j, (s, l, g) in enumerate(lists):
            row_off[j + 1] = row_off[j] + len(l)
            start_off[j + 1] = start_off[j] + s.size
            kk[j] = s.shape[1]
        start = np.concatenate([s.reshape(-1) for s, _, _ in lists])
        length = np.concatenate([l for _, l, _ in lists])
        gain = np.concatenate([g for _, _, g in lists])
        link = np.full(row_off[-1], -7, np.int64)
        score = np.full(row_off[-1], -7, np.int64)
        _native.check(L, L.rv_chain_batch(h, len(lists), row_off.ctypes.data, start_off.ctypes.data, kk.ctypes.data, start.ctypes.data,
                                          length.ctypes.data, gain.ctypes.data, wpen, model, link.ctypes.data, score.ctypes.data))
        for j, (s, l, g) in enumerate(lists):
            wl = np.zeros(len(l), np.int64)
            ws = np.zeros(len(l), np.int64)
            reveallib.chain_dp(s, l, g, wpen, model, wl, ws)
            a, b = int(row_off[j]), int(row_off[j + 1])
            assert link[a:b].tolist() == wl.tolist(), "links of list %d" % j
            assert score[a:b].tolist() == ws.tolist(), "scores of list %d" % j
    finally:
        L.rv_index_free(h)


@pytest.mark.parametrize("model", [0, 1, 2])
def test_chain_batch_matches_host_recurrence_emulated(emu_lib, model):
    rng = np.random.default_rng(model + 5)
    check(emu_lib, random_lists(rng, 12, 60) + random_lists(rng, 6, 40, ties=True), 2, model)


def test_chain_batch_refuses_inconsistent_lists(emu_lib):
    h = ctypes.c_void_p()
    _native.check(emu_lib, emu_lib.rv_index_create(ctypes.byref(h), None))
    z = np.zeros(4, np.int64)
    row_off = np.array([0, 2], np.int64)
    start_off = np.array([0, 3], np.int64)   # 2 rows x k=2 would be 4 cells
    kk = np.array([2], np.int32)
    assert emu_lib.rv_chain_batch(h, 1, row_off.ctypes.data, start_off.ctypes.data, kk.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data, 1, 0,
                                  z.ctypes.data, z.ctypes.data) != 0
    emu_lib.rv_index_free(h)


@pytest.mark.gpu
@pytest.mark.parametrize("model", [0, 1, 2])
def test_chain_batch_matches_host_recurrence_cuda(cuda_lib, model):
    rng = np.random.default_rng(model + 50)
    check(cuda_lib, random_lists(rng, 40, 400) + random_lists(rng, 10, 1001) + random_lists(rng, 20, 120, ties=True), 1, model)
    check(cuda_lib, random_lists(rng, 300, 30), 3, model)
