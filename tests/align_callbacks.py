"""Deterministic stand-ins for schemes.graphmumpicker / rem.graphalign, used to drive BOTH the
reference's `index.align` (unmodified C aligner in oracle/_ref) and reveal_b200's, so that the
sequence of sub-indexes, the MUM lists handed to the picker and the final text can be compared
step by step.  They exercise all three child classes (leading, trailing, parallel)."""


def make_callbacks(log, minlen=1, maxsteps=None):
    state = {"steps": 0}

    def mumpicker(mums, idx, precomputed=False, minlength=0):
        log.append(("pick", idx.depth, idx.n, idx.nsamples, tuple(sorted(idx.nodes)), precomputed, [tuple(m) for m in mums]))
        if maxsteps is not None and state["steps"] >= maxsteps:
            return ()
        cands = [m for m in mums if m[0] >= minlen]
        if not cands:
            return ()
        best = max(cands, key=lambda m: (m[0] * m[1]))  # first maximum
        state["steps"] += 1
        return (best, [], [])

    def graphalign(idx, mum):
        l, n, spd = mum
        nodes = set(idx.nodes)
        leading, trailing, matching, rest = set(), set(), set(), set(nodes)
        for _, p in spd:
            node = None
            for (b, e) in nodes:
                if b <= p and p + l <= e:
                    node = (b, e)
                    break
            if node is None:
                return None
            rest.discard(node)
            b, e = node
            if p > b:
                leading.add((b, p))
            if p + l < e:
                trailing.add((p + l, e))
            matching.add((p, p + l))
        log.append(("align", l, n, tuple(sorted(matching))))
        return (leading, trailing, matching, rest, None, ("newleft", l), ("newright", l))

    return mumpicker, graphalign
