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


_I64 = ctypes.c_int64


def _intervals(obj):
    """Python iterable of (begin, end) -> (ctypes int64 array [b0,e0,b1,e1,...], count, covered length, begins).
    Plain lists + ctypes: a recursion step handles a handful of intervals, numpy would cost more than it saves."""
    flat, begins, total = [], [], 0
    for b, e in obj:
        b = int(b)
        e = int(e)
        flat.append(b)
        flat.append(e)
        begins.append(b)
        total += e - b
    return (_I64 * (len(flat) or 1))(*flat), len(begins), total, begins


def _count_samples(main, begins):
    """Number of distinct samples among the interval starts (reveal.c:1026-1041)."""
    if not begins:
        return 0
    if main._nsamples > 2:
        nsep = main._nsep
        seen = set()
        for b in begins:  # SO[begin] = number of separators before begin
            lo, hi = 0, len(nsep)
            while lo < hi:
                mid = (lo + hi) >> 1
                if nsep[mid] < b:
                    lo = mid + 1
                else:
                    hi = mid
            seen.add(lo)
        return len(seen)
    nsep0 = main._nsep[0]
    left = right = 0
    for b in begins:
        if b < nsep0:
            left = 1
        elif b > nsep0:
            right = 1
    return left + right


def _extract(main, sub, minl, minn):
    """Step 1: MUMs of a sub-index in the shape the reference passes to mumpicker.  When the step that
    created the sub-index already swept it (single-launch path) the calls below answer from that result."""
    L = main._lib()
    if main._nsamples > 2:
        nr, nm = _I64(), _I64()
        main._call(L.rv_sub_mums_multi(sub, minl, minn, ctypes.byref(nr), ctypes.byref(nm)))
        r, m = nr.value, nm.value
        if r == 0:
            return []
        hdr = (_I64 * (3 * r))()
        mem = (_I64 * (2 * m))()
        main._call(L.rv_sub_fetch(sub, hdr, r, mem, m))
        mem = mem[:]
        hdr = hdr[:]
        members = list(zip(mem[0::2], mem[1::2]))
        return [(hdr[k], hdr[k + 1], tuple(members[hdr[k + 2]:hdr[k + 2] + hdr[k + 1]])) for k in range(0, 3 * r, 3)]
    c = _I64()
    main._call(L.rv_sub_mums_pair(sub, minl, ctypes.byref(c)))
    r = c.value
    if r == 0:
        return []
    rows = (_I64 * (3 * r))()
    main._call(L.rv_sub_fetch(sub, rows, r, None, 0))
    rows = rows[:]
    return [(rows[k], 2, ((0, rows[k + 1]), (1, rows[k + 2]))) for k in range(0, 3 * r, 3)]  # reveal.c:167-169


def align(main, mumpicker, graphalign, threads=0, wpen=0, wscore=0, minl=0, minn=0):
    from .reveallib_ctypes import error
    minl = int(minl)
    minn = int(minn)
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
                mum_sp = (_I64 * max(1, mum_n))(*[int(spd[i][1]) for i in range(mum_n)])
                result = graphalign(idx, mumobject)
                if result is None:
                    continue
                if not isinstance(result, tuple):
                    raise error("**** call to graphalign failed")
                if len(result) != 7:
                    continue  # the reference silently drops an unparsable result (reveal.c:987-999)
                leading, trailing, matching, rest, merged, newleft, newright = result
                lead, nlead, leadn, lead_b = _intervals(leading)
                trail, ntrail, trailn, trail_b = _intervals(trailing)
                par, npar, parn, par_b = _intervals(rest)
                match, nmatch, _, _ = _intervals(matching)
                kids = (ctypes.c_void_p * 3)()
                # children without precomputed skipmums will be swept first thing in their own step: let the
                # device do it in the same launch when the parent is small
                sweep = (ctypes.c_int32 * 3)(len(skipleft) == 0, len(skipright) == 0, 1)
                main._call(L.rv_sub_step(view._sub, lead, nlead, trail, ntrail, par, npar, mum_sp, int(mum_n), int(mum_l), match, nmatch,
                                         sweep, minl, minn, kids))
                main._Tdirty = True  # matched bases were lower-cased on the device
                depth = idx.depth + 1
                nmums += 1
                i_lead = i_trail = i_par = None
                if kids[0]:
                    i_lead = SubIndex(main, ctypes.c_void_p(kids[0]), leadn, depth, _count_samples(main, lead_b), leading, idx.leftnode,
                                      newright, skipleft)
                if kids[1]:
                    i_trail = SubIndex(main, ctypes.c_void_p(kids[1]), trailn, depth, _count_samples(main, trail_b), trailing, newleft,
                                       idx.rightnode, skipright)
                if kids[2]:
                    i_par = SubIndex(main, ctypes.c_void_p(kids[2]), parn, depth, _count_samples(main, par_b), rest, idx.leftnode,
                                     idx.rightnode, [])
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
