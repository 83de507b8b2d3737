"""The REM recursion driver -- `index.align()` of the reference, with every array operation on the GPU.

Mirrors `align` (reveallib/interface.c:293-415) and the single-threaded `aligner` loop
(reveallib/reveal.c:731-1338): a LIFO queue of (sub)indexes (reveal.c:18-53); per step

  1. MUM extraction on the popped sub-index -- getmultimums when the MAIN index has more than two
     samples, else getmums_rem (reveal.c:802-829) -- unless the index carries precomputed
     `skipmums`;
  2. callback  mumpicker(mums, idx, precomputed=bool, minlength=int) -> () | (mum, skipleft, skipright)
     (reveal.c:839-895);
  3. callback  graphalign(idx, mum) -> None | (leading, trailing, matching, rest, merged, newleft, newright)
     (reveal.c:937-987);
  4. label scatter + split + T lower-casing + bubble_sort (reveal.c:1005-1252) -> rv_sub_split;
  5. push parallel, leading, trailing children (reveal.c:1296-1324).

Steps 1 and 4 are CUDA (rv_sweep.cu / rv_split.cu through the C-ABI); steps 2, 3, 5 are the
reference's host control flow.  `threads` is accepted for signature compatibility: the callback
sections of the reference are serialised by its `python` mutex anyway (reveal.c:779-780), and the
device work of a step is already parallel, so the steps run in the calling thread in the
reference's LIFO order (= the order of `threads=0`, interface.c:387-399).
"""
import ctypes

import numpy as np


class SubIndex(object):
    """A child index handed to the callbacks: the attribute surface of the reference's `index`
    type that schemes.graphmumpicker / rem.graphalign read."""

    def __init__(self, main, handle, n, depth, nsamples, nodes, leftnode, rightnode, skipmums):
        self.main = main
        self._sub = handle
        self._n = n
        self._depth = depth
        self._nsamples = nsamples
        self.nodes = nodes
        self.leftnode = leftnode
        self.rightnode = rightnode
        self.skipmums = skipmums

    n = property(lambda self: self._n)
    depth = property(lambda self: self._depth)
    nsamples = property(lambda self: self._nsamples)
    samples = property(lambda self: self.main.samples)
    nsep = property(lambda self: self.main.nsep)
    T = property(lambda self: self.main.T)
    SAi = property(lambda self: self.main.SAi)
    SO = property(lambda self: self.main.SO)

    def _arr(self, which):
        L = self.main._lib()
        a = np.empty(self._n, dtype=np.int32)
        self.main._call(L.rv_sub_get(self._sub, which, a.ctypes.data))
        return a

    SA = property(lambda self: self._arr(0).tolist())
    LCP = property(lambda self: self._arr(1).tolist())

    def _free(self):
        if self._sub is not None:
            self.main._lib().rv_sub_free(self._sub)
            self._sub = None


def _intervals(obj):
    """Python iterable of (begin, end) -> contiguous int64 [k,2] array, in iteration order."""
    rows = [(int(b), int(e)) for b, e in obj]
    a = np.asarray(rows, dtype=np.int64).reshape(-1, 2)
    return np.ascontiguousarray(a)


def _count_samples(main, intervals):
    """Number of distinct samples among the interval starts (reveal.c:1026-1041)."""
    if len(intervals) == 0:
        return 0
    begins = intervals[:, 0]
    if main._nsamples > 2:
        nsep = np.asarray(main._nsep, dtype=np.int64)
        return len(np.unique(np.searchsorted(nsep, begins, side="left")))  # SO[begin] = #nsep < begin
    nsep0 = main._nsep[0]
    return int((begins < nsep0).any()) + int((begins > nsep0).any())


def _extract(main, sub, minl, minn):
    """Step 1: MUMs of a sub-index in the shape the reference passes to mumpicker.  When the step that
    created the sub-index already swept it (single-launch path) the calls below answer from that result."""
    L = main._lib()
    if main._nsamples > 2:
        nr, nm = ctypes.c_int64(), ctypes.c_int64()
        main._call(L.rv_sub_mums_multi(sub, int(minl), int(minn), ctypes.byref(nr), ctypes.byref(nm)))
        hdr = np.empty((nr.value, 3), dtype=np.int64)
        mem = np.empty((nm.value, 2), dtype=np.int64)
        main._call(L.rv_sub_fetch(sub, hdr.ctypes.data, nr.value, mem.ctypes.data, nm.value))
        members = [tuple(x) for x in mem.tolist()]
        return [(l, n, tuple(members[first:first + n])) for l, n, first in hdr.tolist()]
    c = ctypes.c_int64()
    main._call(L.rv_sub_mums_pair(sub, int(minl), ctypes.byref(c)))
    rows = np.empty((c.value, 3), dtype=np.int64)
    main._call(L.rv_sub_fetch(sub, rows.ctypes.data, c.value, None, 0))
    return [(l, 2, ((0, a), (1, b))) for l, a, b in rows.tolist()]  # reveal.c:167-169


def align(main, mumpicker, graphalign, threads=0, wpen=0, wscore=0, minl=0, minn=0):
    from .reveallib import error
    L = main._lib()
    main._depth = 0
    main.main = main
    root = ctypes.c_void_p()
    main._call(L.rv_sub_root(main._handle(), ctypes.byref(root)))
    root_view = SubIndex(main, root, main._n, 0, main._nsamples, main.nodes, main.leftnode, main.rightnode, main.skipmums)
    queue = [(main, root_view)]  # (object handed to the callbacks, device view)
    nmums = 0
    try:
        while queue:
            idx, view = queue.pop()  # LIFO (reveal.c:21-26)
            try:
                if not callable(mumpicker):
                    raise TypeError("**** mumpicker isn't callable")
                if len(idx.skipmums) == 0:
                    multimums = _extract(main, view._sub, minl, minn)
                    precomputed = False
                else:
                    multimums = idx.skipmums
                    precomputed = True
                pick = mumpicker(multimums, idx, precomputed=precomputed, minlength=minl)
                if not isinstance(pick, tuple):
                    raise error("**** call to mumpicker failed")
                if len(pick) == 0:
                    continue  # no more MUMs in this sub-index
                mumobject, skipleft, skipright = pick
                mum_l, mum_n, spd = mumobject
                mum_sp = np.asarray([int(spd[i][1]) for i in range(mum_n)], dtype=np.int64)
                result = graphalign(idx, mumobject)
                if result is None:
                    continue
                if not isinstance(result, tuple):
                    raise error("**** call to graphalign failed")
                if len(result) != 7:
                    continue  # the reference silently drops an unparsable result (reveal.c:987-999)
                leading, trailing, matching, rest, merged, newleft, newright = result
                lead = _intervals(leading)
                trail = _intervals(trailing)
                par = _intervals(rest)
                match = _intervals(matching)
                kids = (ctypes.c_void_p * 3)()
                # children without precomputed skipmums will be swept first thing in their own step: let the
                # device do it in the same launch when the parent is small
                sweep = np.asarray([len(skipleft) == 0, len(skipright) == 0, 1], dtype=np.int32)
                main._call(L.rv_sub_step(view._sub, lead.ctypes.data, len(lead), trail.ctypes.data, len(trail), par.ctypes.data, len(par),
                                         mum_sp.ctypes.data, int(mum_n), int(mum_l), match.ctypes.data, len(match), sweep.ctypes.data,
                                         int(minl), int(minn), kids))
                main._Tdirty = True  # matched bases were lower-cased on the device
                depth = idx.depth + 1
                nmums += 1
                i_lead = i_trail = i_par = None
                if kids[0]:
                    i_lead = SubIndex(main, ctypes.c_void_p(kids[0]), int((lead[:, 1] - lead[:, 0]).sum()), depth, _count_samples(main, lead),
                                      leading, idx.leftnode, newright, skipleft)
                if kids[1]:
                    i_trail = SubIndex(main, ctypes.c_void_p(kids[1]), int((trail[:, 1] - trail[:, 0]).sum()), depth, _count_samples(main, trail),
                                       trailing, newleft, idx.rightnode, skipright)
                if kids[2]:
                    i_par = SubIndex(main, ctypes.c_void_p(kids[2]), int((par[:, 1] - par[:, 0]).sum()), depth, _count_samples(main, par),
                                     rest, idx.leftnode, idx.rightnode, [])
                for child in (i_par, i_lead, i_trail):  # push order of reveal.c:1296-1324
                    if child is not None:
                        queue.append((child, child))
            finally:
                view._free()
    finally:
        for _, v in queue:
            v._free()
    main._nmums = nmums
    return None
