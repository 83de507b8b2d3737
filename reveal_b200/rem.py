"""REM -- recursive exact matching: the driver that turns the GPU index into an alignment graph.

Host-side mirror of the reference's `reveal rem` (FASTA and GFA input), Python 3, on top of the drop-in `reveallib`:
the reference entry points keep their names and meaning --

  align_genomes(args) -> (G, idx)      reveal/rem.py:511-611   (args: the namespace of `reveal rem`, see rem_args)
  align(aobjs, ...)   -> (G, idx)      reveal/rem.py:616-712   (library call on (name, sequence) tuples)
  graphalign(idx, mum)                 reveal/rem.py:317-382   callback 2 of index.align
  graphmumpicker(mums, idx, ...)       reveal/schemes.py:197-361  callback 1 of index.align
  chain / trim_overlap / segment       reveal/schemes.py:20-191
  gapcost                              reveal/utils.py:162-183
  prune_nodes                          reveal/rem.py:385-446
  fasta_reader / read_fasta / write_gfa  reveal/utils.py:79-144, 304-375, 710-839

-- and G is the same networkx (Multi)DiGraph: nodes are `Interval(begin, end)` in index coordinates with
attributes `offsets` {path id: offset in that path} and `aligned`, edges carry `paths` (set of path ids),
`ofrom`, `oto`; G.graph holds paths / path2id / id2path / id2end / startnodes / endnodes.

What is different from the reference is the plumbing: no module-level globals (one `Rem` object per alignment owns
the graph and the position -> node map, so several alignments can live in one process), the position lookup is a
sorted array instead of an interval tree (nodes never overlap), and chaining is evaluated with numpy over all
candidate predecessors at once instead of a Python loop per pair.  Results are checked against graphs minted
from the reference's own driver (tests/golden/make_rem_golden.py, tests/test_rem.py).
"""
import argparse
import bisect
import collections
import gzip
import logging
import math
import os
import sys
import uuid

import networkx as nx
import numpy as np

log = logging.getLogger("reveal_b200.rem")


class Interval(collections.namedtuple("Interval", ["begin", "end"])):
    """A graph node: the half-open range [begin, end) of the index text it spells."""
    __slots__ = ()


def _real(G, sid):
    return not G.graph["id2path"][sid].startswith("*")


# ------------------------------------------------------------------------------------------------
# scoring (reveal/utils.py:162-183, reveal/schemes.py:20-191)
def gapcost(pointa, pointb, model="sumofpairs", convex=False, lambda_=1, epsilon_=0):
    """Penalty for the gap between two anchors given as per-sample coordinates (utils.py:162-183)."""
    assert len(pointa) == len(pointb)
    k = len(pointa)
    if model == "star-avg":
        return abs(sum(pointa[i] - pointb[i] for i in range(k))) // k
    if model == "star-med":
        return sorted(abs(pointa[i] - pointb[i]) for i in range(k))[k // 2]
    if model != "sumofpairs":
        log.warning("Unknown penalty model: %s.", model)
        return 0
    dist = [abs(pointa[i] - pointb[i]) for i in range(k)]
    p = min(dist) * epsilon_ if epsilon_ > 0 else 0
    for i in range(k):
        for j in range(i + 1, k):
            d = abs(dist[i] - dist[j])
            p += (math.log(d + 1) if convex else d) * lambda_
    return p


def _gap_matrix(ends, starts, model):
    """gapcost of every candidate predecessor (rows of `ends`, [m, k]) against one anchor (`starts`, [k])."""
    dist = np.abs(ends - starts[None, :])
    k = dist.shape[1]
    if model == "star-avg":
        return np.abs((ends - starts[None, :]).sum(axis=1)) // k
    if model == "star-med":
        return np.sort(dist, axis=1)[:, k // 2]
    pen = np.zeros(dist.shape[0], dtype=np.int64)
    for i in range(k):
        for j in range(i + 1, k):
            pen += np.abs(dist[:, i] - dist[:, j])
    return pen


def chain(mums, left, right, gcmodel="sumofpairs", wscore=1, wpen=1):
    """Best-scoring colinear chain of anchors between the bounds `left` and `right` (schemes.py:20-104).

    mums, left, right: (l, n, {sample: offset}).  Returns [(mum, score), ...] from the LAST anchor of the chain
    to the first, like the reference (its caller reverses it).

    An anchor may follow every earlier anchor (in the order of the first sample's coordinate) that ends before it
    in every sample; its score is the best predecessor score + wscore * l * n(n-1)/2 - wpen * gapcost.  Equal
    totals are resolved as the reference's sorted `active` list does: higher predecessor score first, then the
    predecessor that became available earlier, then the one processed earlier.
    """
    if len(mums) == 0:
        return []
    ref = sorted(mums[0][2].keys())[0]
    order = sorted(list(mums) + [right], key=lambda m: m[2][ref])
    keys = list(right[2].keys())
    m = len(order)
    # row 0 is `left`, rows 1..m the anchors in processing order
    start = np.empty((m + 1, len(keys)), dtype=np.int64)
    length = np.zeros(m + 1, dtype=np.int64)
    gain = np.zeros(m + 1, dtype=np.int64)
    start[0] = [left[2][k] for k in keys]
    for r, mum in enumerate(order, 1):
        start[r] = [mum[2][k] for k in keys]
        length[r] = mum[0]
        gain[r] = wscore * (mum[0] * ((mum[1] * (mum[1] - 1)) // 2))
    score = np.zeros(m + 1, dtype=np.int64)
    link = np.zeros(m + 1, dtype=np.int64)
    if _native_chain is not None and gcmodel in _MODELS:
        _native_chain(start, length, gain, int(wpen), _MODELS[gcmodel], link, score)
    else:
        _chain_numpy(start, length, gain, wpen, gcmodel, link, score)
    log.debug("Best score is: %d", score[m])
    path = []
    r = link[m]                                    # order[m-1] is `right` itself: the chain starts at its predecessor
    while r != 0:
        path.append((order[r - 1], int(score[r])))
        r = link[r]
    return path


_MODELS = {"sumofpairs": 0, "star-avg": 1, "star-med": 2}
try:  # the recurrence in C++ (csrc/ext/reveallib_module.cpp:mod_chain_dp); the numpy form below is its portable twin
    from .reveallib import chain_dp as _native_chain
except ImportError:  # extension not built (e.g. a source checkout used with the ctypes twin)
    _native_chain = None


try:  # the alignment graph in C++ for the duration of the recursion (csrc/ext/remcore_module.cpp); Rem's own methods are its twin
    from . import remcore as _remcore
except ImportError:
    _remcore = None


def _chain_numpy(start, length, gain, wpen, gcmodel, link, score):
    m = len(length) - 1
    end = start + length[:, None]
    joined = np.full(m + 1, -1, dtype=np.int64)     # iteration at which a processed anchor became a candidate
    joined[0] = 0
    for r in range(1, m + 1):
        ok = (end[:r] <= start[r][None, :]).all(axis=1)
        fresh = ok & (joined[:r] < 0)
        joined[:r][fresh] = r
        cand = np.nonzero(ok)[0]
        s = score[cand] + gain[r]
        total = s - wpen * _gap_matrix(end[cand], start[r], gcmodel)
        best = cand[total == total.max()]
        if len(best) > 1:
            best = best[np.lexsort((best, joined[best], -score[best]))]
        link[r] = best[0]
        score[r] = total.max()


def segment(mums):
    """The group of anchors over one and the same set of samples with the largest (total length x #samples)
    (schemes.py:107-125)."""
    groups = collections.OrderedDict()
    for mum in mums:
        groups.setdefault(tuple(sorted(gid for gid, _ in mum[2])), []).append(mum)
    best, pick = 0, None
    for part, members in groups.items():
        z = sum(m[0] for m in members) * len(part)
        if z > best:
            best, pick = z, part
    return groups[pick]


def trim_overlap(mums):
    """Removes overlap between anchors, one coordinate (= member slot of the anchor tuples) at a time
    (schemes.py:161-191): anchors contained in their neighbour go, overlapping ends are cut back."""
    for coord in range(len(mums[0][2])):
        if len(mums) <= 1:
            break
        mums.sort(key=lambda m: (m[2][coord][1], -m[0]))
        ends = [m[2][coord][1] + m[0] for m in mums]
        # the reference's filter indexes mums[i-1] with i == 0 too, i.e. compares the first anchor with the LAST one
        mums = [mum for i, mum in enumerate(mums) if (i == 0 and ends[1] > ends[0]) or ends[i - 1] < ends[i]]
        if len(mums) <= 1:
            break
        trimmed = [mums[0]]
        for mum in mums[1:]:
            prev = trimmed[-1]
            overlap = prev[2][coord][1] + prev[0] - mum[2][coord][1]
            if overlap > 0:
                if prev[0] - overlap > 0:
                    trimmed[-1] = (prev[0] - overlap, prev[1], prev[2])
                else:
                    del trimmed[-1]
                if mum[0] - overlap > 0:
                    trimmed.append((mum[0] - overlap, mum[1], tuple((k, v + overlap) for k, v in mum[2])))
            else:
                trimmed.append(mum)
        mums = trimmed
    return mums


# ------------------------------------------------------------------------------------------------
def rem_args(inputfiles=(), **overrides):
    """The option namespace of `reveal rem` with its defaults (reveal/reveal.py:74-99)."""
    a = dict(inputfiles=list(inputfiles), output=None, threads=0, minlength=20, pcutoff=1e-8, minn=2, gcmodel="sumofpairs",
             wpen=1, wscore=1, seedsize=10000, maxmums=1000, sa="", lcp="", cache=False, sa64=False, toupper=True, maxsize=None,
             contigs=True, trim=True, splitchain="largest", maxdepth=None)
    unknown = set(overrides) - set(a)
    if unknown:
        raise TypeError("unknown rem option(s): %s" % ", ".join(sorted(unknown)))
    a.update(overrides)
    return argparse.Namespace(**a)


class Rem(object):
    """One alignment in progress: the graph, the position -> node map and the two callbacks of index.align."""

    def __init__(self, args, G=None):
        self.args = args
        self.G = nx.MultiDiGraph() if G is None else G
        self.multi = isinstance(self.G, nx.MultiDiGraph)
        self.begins = []                   # sorted begin of every Interval node (bisect: C speed on the hot lookup)
        self.end_of = {}                   # begin -> end
        self._all_real = None
        self._coords = {}                  # index position -> ((path id, coordinate in that path), ...), see _lookup
        self.core = None                   # remcore.Graph while the recursion runs (see recursion_graph)
        self.batch_picker = None           # remcore.Graph.mumpicker_batch when the native callbacks are in use (see callbacks)
        self.shard = None                  # (rank, world, process group, index) of a sharded recursion, see align_genomes(shard=...)
        self.shard_stats = None
        for key, empty in (("paths", list), ("id2path", dict), ("path2id", dict), ("id2end", dict)):
            self.G.graph.setdefault(key, empty())

    # ---- position -> node ------------------------------------------------------------------
    def node_at(self, pos):
        i = bisect.bisect_right(self.begins, pos) - 1
        if i >= 0:
            begin = self.begins[i]
            end = self.end_of[begin]
            if pos < end:
                return Interval(begin, end)
        raise KeyError("no node covers index position %d" % pos)

    def _track(self, begin, end):
        bisect.insort(self.begins, begin)
        self.end_of[begin] = end

    def _untrack(self, begin):
        if begin in self.end_of:
            del self.end_of[begin]
            del self.begins[bisect.bisect_left(self.begins, begin)]

    # ---- input -----------------------------------------------------------------------------
    def add_sequence(self, index, name, seq):
        """One path of the graph = one sequence of the index, between its own start and end marker nodes
        (utils.py:326-347)."""
        g = self.G.graph
        g.setdefault("startnodes", [])
        g.setdefault("endnodes", [])
        name = name.replace(":", "").replace(";", "")
        if name in g["paths"]:
            raise ValueError("Fasta with this name: \"%s\" is already contained in the graph." % name)
        sid = len(g["paths"])
        g["paths"].append(name)
        g["path2id"][name] = sid
        g["id2path"][sid] = name
        g["id2end"][sid] = len(seq)
        begin, end = index.addsequence(seq)
        node = Interval(begin, end)
        self._track(begin, end)
        first, last = uuid.uuid4().hex, uuid.uuid4().hex
        self.G.add_node(first, offsets={sid: 0}, endpoint=True)
        g["startnodes"].append(first)
        self.G.add_node(node, offsets={sid: 0}, aligned=0)
        self.G.add_node(last, offsets={sid: len(seq)}, endpoint=True)
        g["endnodes"].append(last)
        self.G.add_edge(first, node, paths={sid}, ofrom="+", oto="+")
        self.G.add_edge(node, last, paths={sid}, ofrom="+", oto="+")

    def read_fasta(self, fasta, index, contigs=True, toupper=True):
        """contigs=True: the file is ONE sample whose sequences are its contigs; else every sequence is a sample
        (utils.py:304-375)."""
        if contigs:
            index.addsample(os.path.basename(fasta))
        for name, seq in fasta_reader(fasta, toupper=toupper):
            if not contigs:
                index.addsample(name)
            self.add_sequence(index, name, seq)

    def read_gfa(self, gfafile, index):
        """A GFA 1 graph as ONE sample of the index: every segment becomes its own `$`-terminated sequence and a
        node, links become edges, P lines give every node its per-path offsets and every traversed edge its paths;
        untraversed edges and nodes are dropped and each connected component gets one shared start and one shared
        end marker (utils.py:377-657, the `index is not None` form used by rem)."""
        G = self.G
        g = G.graph
        g.setdefault("startnodes", [])
        g.setdefault("endnodes", [])
        index.addsample(os.path.basename(gfafile))
        node_of = {}
        links, walks = [], []
        with (gzip.open(gfafile, "rt") if gfafile.endswith(".gz") else open(gfafile, "r")) as f:
            for line in f:
                if line.startswith("S"):
                    cols = line.strip().split("\t")
                    seq = cols[2] if len(cols) > 2 else ""
                    begin, end = index.addsequence(seq.upper())
                    node = Interval(begin, end)
                    self._track(begin, end)
                    G.add_node(node, aligned=0, offsets={})
                    node_of[int(cols[1])] = node
                elif line.startswith("L"):
                    links.append(line)
                elif line.startswith("P"):
                    walks.append(line)
        for line in links:
            e = line.strip().split("\t")
            if not self.multi and (e[2] != "+" or e[4] != "+"):
                continue  # a plain DiGraph only holds the forward-forward links
            tags = {"ofrom": e[2], "oto": e[4]}
            if len(e) > 5:
                tags["cigar"] = e[5]
            if "*" in e:
                for tag in e[7:]:
                    key, _, value = tag.split(":")
                    tags[key.lower()] = value
            tags["paths"] = set()
            G.add_edge(node_of[int(e[1])], node_of[int(e[3])], **tags)
        if not walks:
            raise ValueError("No paths defined in GFA: %s" % gfafile)
        firsts, lasts = set(), set()
        for line in walks:
            cols = line.rstrip().split("\t")
            name = cols[1]
            if not self.multi and name.startswith("*"):
                continue
            if name in g["paths"] or name in g["path2id"]:
                raise ValueError("Graph already contains path for: %s" % name)
            sid = len(g["path2id"])
            g["paths"].append(name)
            g["path2id"][name] = sid
            g["id2path"][sid] = name
            steps = [(int(x[:-1]), x[-1:]) for x in cols[2].split(",")]
            offset = 0
            prev = prev_orient = None
            for nid, orient in steps:
                node = node_of[nid]
                G.nodes[node]["offsets"][sid] = offset
                offset += node.end - node.begin
                if prev is not None:
                    if node not in G[prev]:
                        raise ValueError("Path %s steps to segment %d over a link the graph does not define" % (name, nid))
                    if self.multi:
                        for d in G[prev][node].values():
                            if d["oto"] == orient and d["ofrom"] == prev_orient:
                                d["paths"].add(sid)
                                break
                        else:
                            raise ValueError("Path %s: no link with orientation %s%s into segment %d" % (name, prev_orient, orient, nid))
                    else:
                        G[prev][node]["paths"].add(sid)
                prev, prev_orient = node, orient
            first, last = uuid.uuid4().hex, uuid.uuid4().hex
            G.add_node(first, offsets={sid: 0}, endpoint=True)
            G.add_edge(first, node_of[steps[0][0]], paths={sid}, ofrom="+", oto=steps[0][1])
            firsts.add(first)
            G.add_node(last, offsets={sid: offset}, endpoint=True)
            G.add_edge(node_of[steps[-1][0]], last, paths={sid}, ofrom=steps[-1][1], oto="+")
            lasts.add(last)
            g["id2end"][sid] = offset
        G.remove_edges_from([(u, v) for u, v, d in G.edges(data=True) if not d["paths"]])
        for node in [n for n, d in G.nodes(data=True) if not d["offsets"]]:
            if isinstance(node, Interval):
                self._untrack(node.begin)
            G.remove_node(node)
        # one start and one end marker per connected component, in place of one pair per path
        for comp in [c for c in nx.weakly_connected_components(G)]:
            for markers, bucket, forward in ((lasts, "endnodes", False), (firsts, "startnodes", True)):
                mine = [n for n in comp if n in markers]
                if not mine:
                    continue
                shared = uuid.uuid4().hex
                G.add_node(shared, offsets={}, seq="", endpoint=True)
                g[bucket].append(shared)
                for marker in mine:
                    G.nodes[shared]["offsets"].update(G.nodes[marker]["offsets"])
                    for other in list(G.successors(marker) if forward else G.predecessors(marker)):
                        d = (G[marker][other] if forward else G[other][marker])
                        d = d[0] if self.multi else d
                        u, v = (shared, other) if forward else (other, shared)
                        if not self.multi and G.has_edge(u, v):
                            G[u][v]["paths"].update(d["paths"])
                        else:
                            G.add_edge(u, v, paths=d["paths"], ofrom=d["ofrom"], oto=d["oto"])
            G.remove_nodes_from([n for n in comp if n in firsts or n in lasts])

    # ---- the graph during the recursion ----------------------------------------------------------
    def recursion_graph(self):
        """Context manager around index.align(): moves the graph into remcore.Graph (flat C++ vectors; graphalign and
        the coordinate look-ups of the mumpicker then run there) and rebuilds the networkx graph from it afterwards.
        Without the compiled module, or with RV_REM_PYTHON_GRAPH=1, the recursion works on the networkx graph itself."""
        return _RecursionGraph(self)

    def _load_core(self):
        if _remcore is None or os.environ.get("RV_REM_PYTHON_GRAPH", "0") not in ("", "0"):
            return
        G = self.G
        ids = G.graph["id2path"]
        ends = G.graph["id2end"]
        core = _remcore.Graph(self.multi, Interval, [not ids[sid].startswith("*") for sid in range(len(ids))],
                              [ends[sid] for sid in range(len(ids))])
        for node, d in G.nodes(data=True):
            extra = {k: v for k, v in d.items() if k not in ("offsets", "aligned")}
            core.add_node(node, d.get("aligned"), d.get("offsets") or {}, extra or None)
        for u, v, d in G.edges(data=True):
            extra = {k: x for k, x in d.items() if k not in ("paths", "ofrom", "oto")}
            core.add_edge(u, v, d["ofrom"], d["oto"], d["paths"], extra or None)
        self.core = core

    def _unload_core(self):
        core, self.core = self.core, None
        if core is None:
            return
        G = self.G
        import gc
        was_on = gc.isenabled()
        gc.disable()  # a burst of long-lived containers: generational collections in the middle of it only re-scan them
        try:
            self._rebuild_from(core)
        finally:
            if was_on:
                gc.enable()

    # ---- a rank's part of the graph as flat arrays: what travels between the ranks of a sharded recursion ----------------------
    @staticmethod
    def _pack_part(nodes, edges, texts, markers):
        """Node rows, edge rows and marked text stretches as one bytes object of flat numpy arrays (pickling hundreds of
        thousands of small tuples, dicts and sets costs several times more than filling arrays)."""
        import io
        import pickle

        import numpy as np
        mi = {m: i for i, m in enumerate(markers)}
        nb = np.fromiter((k.begin for k, _ in nodes), np.int64, len(nodes))
        ne = np.fromiter((k.end for k, _ in nodes), np.int64, len(nodes))
        al = np.fromiter((-1 if a.get("aligned") is None else a["aligned"] for _, a in nodes), np.int8, len(nodes))
        noff = np.fromiter((len(a.get("offsets") or ()) for _, a in nodes), np.int32, len(nodes))
        osid = np.fromiter((sid for _, a in nodes for sid in (a.get("offsets") or ())), np.int32, int(noff.sum()))
        oval = np.fromiter((v for _, a in nodes for v in (a.get("offsets") or {}).values()), np.int64, int(noff.sum()))
        nextra = {i: {k: v for k, v in a.items() if k not in ("offsets", "aligned")} for i, (_, a) in enumerate(nodes) if len(a) > 2}

        def ends(which):
            b = np.empty(len(edges), np.int64)
            e = np.empty(len(edges), np.int64)
            for i, row in enumerate(edges):
                k = row[which]
                if isinstance(k, str):
                    b[i], e[i] = -1 - mi[k], 0
                else:
                    b[i], e[i] = k.begin, k.end
            return b, e
        ub, ue = ends(0)
        vb, ve = ends(1)
        orient = np.fromiter(((a["ofrom"] == "-") | ((a["oto"] == "-") << 1) for _, _, a in edges), np.int8, len(edges))
        npath = np.fromiter((len(a["paths"]) for _, _, a in edges), np.int32, len(edges))
        paths = np.fromiter((x for _, _, a in edges for x in a["paths"]), np.int32, int(npath.sum()))
        eextra = {i: {k: v for k, v in a.items() if k not in ("paths", "ofrom", "oto")} for i, (_, _, a) in enumerate(edges) if len(a) > 3}
        tb = np.fromiter((b for b, _ in texts), np.int64, len(texts))
        tl = np.fromiter((len(t) for _, t in texts), np.int64, len(texts))
        tt = np.frombuffer("".join(t for _, t in texts).encode("latin-1"), np.uint8)
        buf = io.BytesIO()
        np.savez(buf, nb=nb, ne=ne, al=al, noff=noff, osid=osid, oval=oval, ub=ub, ue=ue, vb=vb, ve=ve, orient=orient, npath=npath, paths=paths,
                 tb=tb, tl=tl, tt=tt, extra=np.frombuffer(pickle.dumps((nextra, eextra, len(markers))), np.uint8))
        return buf.getvalue()

    @staticmethod
    def _unpack_part(blob, markers):
        """Inverse of _pack_part; marker ends are spelled with THIS rank's marker names (same creation order on every rank)."""
        import io
        import pickle

        import numpy as np
        z = np.load(io.BytesIO(blob))
        nextra, eextra, nmark = pickle.loads(z["extra"].tobytes())
        if nmark != len(markers):
            raise RuntimeError("sharded recursion: the ranks read different inputs")
        nb, ne, al, noff = z["nb"].tolist(), z["ne"].tolist(), z["al"].tolist(), z["noff"].tolist()
        osid, oval = z["osid"].tolist(), z["oval"].tolist()
        nodes, at = [], 0
        for i in range(len(nb)):
            attrs = {"offsets": dict(zip(osid[at:at + noff[i]], oval[at:at + noff[i]]))}
            if al[i] >= 0:
                attrs["aligned"] = al[i]
            at += noff[i]
            if i in nextra:
                attrs.update(nextra[i])
            nodes.append((Interval(nb[i], ne[i]), attrs))
        ub, ue, vb, ve = z["ub"].tolist(), z["ue"].tolist(), z["vb"].tolist(), z["ve"].tolist()
        orient, npath, paths = z["orient"].tolist(), z["npath"].tolist(), z["paths"].tolist()
        edges, at = [], 0
        for i in range(len(ub)):
            u = markers[-1 - ub[i]] if ub[i] < 0 else Interval(ub[i], ue[i])
            v = markers[-1 - vb[i]] if vb[i] < 0 else Interval(vb[i], ve[i])
            attrs = {"paths": set(paths[at:at + npath[i]]), "ofrom": "-" if orient[i] & 1 else "+", "oto": "-" if orient[i] & 2 else "+"}
            at += npath[i]
            if i in eextra:
                attrs.update(eextra[i])
            edges.append((u, v, attrs))
        tb, tl = z["tb"].tolist(), z["tl"].tolist()
        tt = z["tt"].tobytes().decode("latin-1")
        texts, at = [], 0
        for b, ln in zip(tb, tl):
            texts.append((b, tt[at:at + ln]))
            at += ln
        return nodes, edges, texts

    @staticmethod
    def _gather_blobs(blob, rank, world, group):
        """bytes of every rank -> list on rank 0 (None elsewhere).  Two collectives on the job's existing communicator (the lengths,
        then the padded bytes: all_gather); a `gather` would be point-to-point sends, whose NCCL connections are only set up at
        first use -- that set-up, not the megabytes, was most of the collection time."""
        import numpy as np
        import torch
        import torch.distributed as dist
        backend = dist.get_backend(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
        size = torch.tensor([len(blob) if blob is not None else 0], dtype=torch.int64, device=dev)
        sizes = [torch.zeros_like(size) for _ in range(world)]
        dist.all_gather(sizes, size, group=group)
        sizes = [int(x.item()) for x in sizes]
        width = max(max(sizes), 1)
        mine = torch.zeros(width, dtype=torch.uint8, device=dev)
        if blob:
            mine[:len(blob)] = torch.from_numpy(np.frombuffer(blob, dtype=np.uint8).copy()).to(dev)
        rows = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(rows, mine, group=group)
        if rank != 0:
            return None
        return [rows[r][:sizes[r]].cpu().numpy().tobytes() if sizes[r] else None for r in range(world)]

    def _collect_shards(self, nodes, edges):
        """Sharded recursion (index.align(shard_rank=, shard_world=)): every rank processed the tree above the cut and its own
        units below it, so its graph is final for the nodes of its units and stale for the units of the others.  The owners
        send their parts -- node rows, the edge rows touching them, the marked stretches of text -- to rank 0 (one gather of
        pickled rows: NCCL moves the bytes between GPUs, gloo in the CPU tests), which swaps them in.  Returns the merged rows
        on rank 0 and the rank's own (partly stale) rows elsewhere."""
        import bisect
        import time

        import torch.distributed as dist
        rank, world, group, idx = self.shard
        t0 = time.perf_counter()
        spans = sorted((b, e, owner) for owner, unit in idx.shard_units for b, e in unit)
        begins = [b for b, _, _ in spans]

        def owner_of(key):
            if isinstance(key, str):
                return None
            i = bisect.bisect_right(begins, key.begin) - 1
            if i >= 0 and key.begin < spans[i][1]:
                return spans[i][2]
            return None

        who = {key: owner_of(key) for key, _ in nodes}
        T = idx.T
        markers = [k for k, _ in nodes if isinstance(k, str)]   # start / end markers: random names, same creation order on every rank
        own_nodes = sum(1 for k, _ in nodes if who[k] == rank)
        mine = None
        if rank != 0:
            mine = self._pack_part([(k, a) for k, a in nodes if who[k] == rank],
                                   [(u, v, a) for u, v, a in edges if who.get(u) == rank or who.get(v) == rank],
                                   [(b, T[b:e]) for b, e, owner in spans if owner == rank], markers)
        t1 = time.perf_counter()
        parts = self._gather_blobs(mine, rank, world, group)
        t2 = time.perf_counter()
        self.shard_stats = {"units": len(idx.shard_units), "own_units": sum(1 for o, _ in idx.shard_units if o == rank),
                            "own_nodes": own_nodes, "pack_s": t1 - t0, "gather_s": t2 - t1,
                            "part_bytes": len(mine) if mine is not None else 0}
        if rank != 0:
            return nodes, edges
        foreign = {k for k, o in who.items() if o is not None and o != 0}
        out_nodes = [(k, a) for k, a in nodes if k not in foreign]
        out_edges = [(u, v, a) for u, v, a in edges if u not in foreign and v not in foreign]
        for r in range(1, world):
            pn, pe, pt = self._unpack_part(parts[r], markers)
            out_nodes.extend(pn)
            out_edges.extend(pe)
            for b, text in pt:
                idx.puttext(b, text)
        known = {k for k, _ in out_nodes}
        for u, v, _ in out_edges:
            if u not in known or v not in known:
                raise RuntimeError("sharded recursion: an edge joins the units of two ranks (%r -> %r); align this input unsharded" % (u, v))
        self.shard_stats["merge_s"] = time.perf_counter() - t2
        return out_nodes, out_edges

    def _rebuild_from(self, core):
        G = self.G
        nodes, edges = core.export()   # [(node, attributes)], [(u, v, attributes)], attribute dicts freshly made
        if self.shard is not None:
            nodes, edges = self._collect_shards(nodes, edges)
        keep = dict(G.graph)
        G.clear()
        G.graph.update(keep)
        self.end_of = {}
        if type(G) in (nx.MultiDiGraph, nx.DiGraph):
            # hundreds of thousands of nodes and edges: placed straight into the adjacency dicts, the way add_node / add_edge do
            node_of, succ, pred = G._node, G._succ, G._pred
            for key, attrs in nodes:
                node_of[key] = attrs
                succ[key] = {}
                pred[key] = {}
            if self.multi:
                for u, v, attrs in edges:
                    keyed = succ[u].get(v)
                    if keyed is None:
                        keyed = succ[u][v] = pred[v][u] = {}
                    keyed[len(keyed)] = attrs
            else:
                for u, v, attrs in edges:
                    succ[u][v] = pred[v][u] = attrs
        else:
            for key, attrs in nodes:
                G.add_node(key, **attrs)
            for u, v, attrs in edges:
                G.add_edge(u, v, **attrs)
        for key, attrs in nodes:
            if attrs.get("aligned") == 0:
                self.end_of[key.begin] = key.end
        self.begins = sorted(self.end_of)

    def _offsets_of(self, node):
        return self.core.node_offsets(node) if self.core is not None else self.G.nodes[node]["offsets"]

    # ---- graph surgery -----------------------------------------------------------------------
    # edge lists straight from the adjacency dicts, in the order networkx' in_edges / out_edges views would give them
    def _edges_in(self, node):
        pred = self.G._pred[node]
        if self.multi:
            return [(u, d) for u, keyed in pred.items() for d in keyed.values()]
        return list(pred.items())

    def _edges_out(self, node):
        succ = self.G._succ[node]
        if self.multi:
            return [(v, d) for v, keyed in succ.items() for d in keyed.values()]
        return list(succ.items())

    def _only_real_paths(self):
        """True when no path is a '*' (hidden) path -- always the case for FASTA input; lets the walks skip the
        per-edge path filter."""
        if self._all_real is None or self._all_real[0] != len(self.G.graph["id2path"]):
            self._all_real = (len(self.G.graph["id2path"]), all(not p.startswith("*") for p in self.G.graph["id2path"].values()))
        return self._all_real[1]

    def breaknode(self, node, pos, l):
        """Cuts [pos, pos+l) out of `node`; returns the matching piece and the set of left-over pieces
        (rem.py:14-131)."""
        G = self.G
        piece = Interval(pos, pos + l)
        if piece == node:
            self._untrack(node.begin)
            return node, set()
        att = G.nodes[node]
        ins, outs = self._edges_in(node), self._edges_out(node)
        shift = pos - node.begin
        mid_off = {s: o + shift for s, o in att["offsets"].items()}
        suf_off = {s: o + shift + l for s, o in att["offsets"].items()}
        # paths that run through the node on the other strand get mirrored edges between the pieces
        pos_paths, neg_paths = set(), set()
        if not ins and not outs:
            pos_paths = set(att["offsets"])
        for _, d in ins:
            (neg_paths if d["oto"] == "-" else pos_paths).update(d["paths"])
        for _, d in outs:
            (neg_paths if d["ofrom"] == "-" else pos_paths).update(d["paths"])
        assert not (pos_paths & neg_paths), "paths traverse node %s on both strands" % (node,)
        self._untrack(node.begin)
        G.add_node(piece, offsets=mid_off, aligned=0)
        others = set()
        head = tail = piece
        if node.begin != pos:
            head = Interval(node.begin, pos)
            G.add_node(head, offsets=att["offsets"], aligned=0)
            G.add_edge(head, piece, paths=set(pos_paths), ofrom="+", oto="+")
            if neg_paths:
                G.add_edge(piece, head, paths=set(neg_paths), ofrom="-", oto="-")
            self._track(head.begin, head.end)
            others.add(head)
        if node.end != pos + l:
            tail = Interval(pos + l, node.end)
            G.add_node(tail, offsets=suf_off, aligned=0)
            G.add_edge(piece, tail, paths=set(pos_paths), ofrom="+", oto="+")
            if neg_paths:
                G.add_edge(tail, piece, paths=set(neg_paths), ofrom="-", oto="-")
            self._track(tail.begin, tail.end)
            others.add(tail)
        G.remove_node(node)
        for u, d in ins:
            G.add_edge(u, head if d["oto"] == "+" else tail, **d)
        for v, d in outs:
            G.add_edge(tail if d["ofrom"] == "+" else head, v, **d)
        return piece, others

    def mergenodes(self, group):
        """Folds the nodes of `group` into its first member: offsets united, edges re-attached, parallel edges with
        the same orientation united by their paths (rem.py:133-205)."""
        G = self.G
        keep = group[0]
        offsets = {}
        for node in group:
            offsets.update(G.nodes[node]["offsets"])
        G.nodes[keep]["offsets"] = offsets
        G.nodes[keep]["aligned"] = 1
        for node in group[1:]:
            if self.multi:
                for u, d in self._edges_in(node):
                    for d2 in G._pred[keep].get(u, {}).values():  # an edge u -> keep with the same orientation: unite the paths
                        if d2["oto"] == d["oto"] and d2["ofrom"] == d["ofrom"]:
                            d2["paths"].update(d["paths"])
                            break
                    else:
                        G.add_edge(u, keep, **d)
                for v, d in self._edges_out(node):
                    for d2 in G._succ[keep].get(v, {}).values():
                        if d2["oto"] == d["oto"] and d2["ofrom"] == d["ofrom"]:
                            d2["paths"].update(d["paths"])
                            break
                    else:
                        G.add_edge(keep, v, **d)
            else:
                for u, _, d in list(G.in_edges(node, data=True)):
                    if G.has_edge(u, keep):
                        G[u][keep]["paths"].update(d["paths"])
                    else:
                        G.add_edge(u, keep, **d)
                for _, v, d in list(G.out_edges(node, data=True)):
                    if G.has_edge(keep, v):
                        G[keep][v]["paths"].update(d["paths"])
                    else:
                        G.add_edge(keep, v, **d)
            G.remove_node(node)
        return keep

    def _neighbours(self, node, backwards):
        """Neighbours over edges that carry at least one real (non '*') path (rem.py:207-233)."""
        G = self.G
        adj = G._pred[node] if backwards else G._succ[node]
        if self._only_real_paths():
            return adj  # every edge carries at least one path
        out = []
        for other, edges in adj.items():
            bundle = edges.values() if self.multi else (edges,)
            if any(_real(G, p) for e in bundle for p in e["paths"]):
                out.append(other)
        return out

    def _reach(self, source, backwards=False, through=()):
        """Breadth-first walk that stops at aligned nodes (class 1) and at the start/end markers (class 2) and runs
        on through unaligned ones (class 0), and through the aligned nodes listed in `through` (rem.py:235-262)."""
        G = self.G
        seen = {source}
        queue = collections.deque([source])
        while queue:
            for child in self._neighbours(queue.popleft(), backwards):
                if child in seen:
                    continue
                seen.add(child)
                data = G.nodes[child]
                if "aligned" not in data:
                    yield child, 2
                elif data["aligned"] == 0 or child in through:
                    queue.append(child)
                    yield child, 0
                else:
                    yield child, 1

    def segmentgraph(self, node, nodes):
        """Splits the intervals of a (sub)index around the freshly merged `node` into those that lie strictly
        before it, strictly after it, and the rest (rem.py:264-315)."""
        sides = []
        for backwards in (False, True):
            side, stops = set(), set()
            for c, kind in self._reach(node, backwards):
                (side if kind == 0 else stops).add(c)
            if len(stops) > 1:  # keep only what every stop reaches when walking back towards the node
                back = set()
                for stop in stops:
                    back.update(c for c, kind in self._reach(stop, not backwards, through=stops) if kind == 0)
                side &= back
            sides.append({(c.begin, c.end) for c in side if isinstance(c, Interval)} & nodes)
        trailing, leading = sides
        return leading, trailing, nodes - (leading | trailing)

    # ---- callback 2: graphalign(index, mum) (rem.py:317-382) --------------------------------
    def callbacks(self, minlength):
        """(mumpicker, graphalign) for index.align().  With the graph in C++ and the default options (splitchain="largest", a length
        threshold, no maxsize) both are methods of remcore.Graph -- index.align() then runs a whole recursion without entering
        a Python frame; otherwise the methods of this object, which are their readable twins."""
        args = self.args
        core = self.core
        native = (core is not None and hasattr(core, "mumpicker") and args.splitchain == "largest" and minlength != 0
                  and args.gcmodel in _MODELS and args.maxsize is None and os.environ.get("RV_REM_PYTHON_PICK", "0") in ("", "0"))
        self.batch_picker = None
        if not native:
            return self.graphmumpicker, self.graphalign
        self.batch_picker = getattr(core, "mumpicker_batch", None)   # the picks of a whole frontier batch in one call
        core.set_picker(bool(args.trim), int(args.maxmums), _MODELS[args.gcmodel], int(args.wscore), int(args.wpen), int(args.seedsize),
                        -1 if args.maxdepth is None else int(args.maxdepth))
        return core.mumpicker, core.graphalign_cb

    def graphalign(self, index, mum):
        try:
            l, n, spd = mum
            if self.core is not None:
                return self.core.graphalign(index.nodes, index.leftnode, index.rightnode, l, [pos for _, pos in spd])
            nodes = index.nodes
            G = self.G
            pieces = []
            matching = set()
            for _, pos in spd:
                matching.add((pos, pos + l))
                old = self.node_at(pos)
                assert old.end - old.begin >= l
                piece, others = self.breaknode(old, pos, l)
                pieces.append(piece)
                nodes.remove((old.begin, old.end))
                for o in others:
                    nodes.add((o.begin, o.end))
            merged = self.mergenodes(pieces)
            msamples = set(G.nodes[merged]["offsets"])
            leading, trailing, rest = self.segmentgraph(merged, set(nodes))
            newleft = newright = merged
            # a side whose intervals belong to samples outside the match is not cleanly cut by it: keep the old bound
            if any(not set(G.nodes[Interval(*iv)]["offsets"]) <= msamples for iv in leading):
                newright = index.rightnode
            if any(not set(G.nodes[Interval(*iv)]["offsets"]) <= msamples for iv in trailing):
                newleft = index.leftnode
            return leading, trailing, matching, rest, merged, newleft, newright
        except Exception:
            log.exception("graphalign failed")
            raise

    # ---- callback 1: graphmumpicker (schemes.py:197-361) -----------------------------------
    def _lookup(self, mum):
        """Anchor in index coordinates -> (l, n, {path id: offset in the path}) (schemes.py:127-150).

        The per-path coordinates of an index position never change while the position is still unaligned (breaking
        a node shifts the offsets of its pieces by exactly the cut, and only matched -- then aligned -- pieces are
        ever merged), and anchors only ever lie in unaligned text, so they are looked up once per position."""
        known = self._coords
        l, _, spd = mum
        n = 0
        point = {}
        for _, pos in spd:
            coords = known.get(pos)
            if coords is None and self.core is not None:
                coords = known[pos] = self.core.coords(pos)
            if coords is None:
                G = self.G
                begins = self.begins
                begin = begins[bisect.bisect_right(begins, pos) - 1]
                rel = pos - begin
                all_real = self._only_real_paths()
                coords = known[pos] = tuple((k, off + rel) for k, off in G._node[(begin, self.end_of[begin])]["offsets"].items()
                                            if all_real or _real(G, k))
            n += len(coords)
            point.update(coords)
        return (l, n, point)

    def _bounds(self, idx, keys):
        G = self.G
        if idx.leftnode is not None:
            off = self._offsets_of(idx.leftnode)
            left = {k: off[k] + (idx.leftnode[1] - idx.leftnode[0]) - 1 for k in keys}
        else:
            left = {k: -1 for k in keys}
        if idx.rightnode is not None:
            off = self._offsets_of(idx.rightnode)
            right = {k: off[k] for k in keys}
        else:
            right = {k: G.graph["id2end"][k] for k in keys}
        return (0, 0, left), (0, 0, right)

    def _small_enough(self, idx):
        """--maxbubblesize: true when every path between the bounds of idx is at most args.maxsize long."""
        G = self.G
        real = [G.graph["path2id"][p] for p in G.graph["paths"] if not p.startswith("*")]
        if idx.leftnode is None:
            lo = {k: 0 for k in real}
        else:
            off = self._offsets_of(idx.leftnode)
            lo = {k: off[k] + (idx.leftnode[1] - idx.leftnode[0]) for k in off}
        ro = {k: G.graph["id2end"][k] for k in real} if idx.rightnode is None else self._offsets_of(idx.rightnode)
        return all(ro[k] - lo[k] <= self.args.maxsize for k in set(lo) & set(ro))

    def graphmumpicker(self, mums, idx, precomputed=False, minlength=0):
        try:
            args = self.args
            if len(mums) == 0:
                return ()
            if precomputed:  # a chain handed down by the parent: split at its middle (schemes.py:346-351)
                half = len(mums) // 2
                return mums[half][0], mums[:half], mums[half + 1:]
            if args.maxdepth is not None and idx.depth > args.maxdepth:
                return ()
            if args.maxsize is not None and self._small_enough(idx):
                return ()
            if (self.core is not None and args.splitchain == "largest" and minlength != 0 and args.gcmodel in _MODELS
                    and os.environ.get("RV_REM_PYTHON_PICK", "0") in ("", "0")):
                # the default flow below, in C++ on the graph that lives there during the recursion (remcore.Graph.pick)
                return self.core.pick(mums, idx.nsamples, idx.leftnode, idx.rightnode, bool(args.trim), int(args.maxmums),
                                      _MODELS[args.gcmodel], int(args.wscore), int(args.wpen), int(args.seedsize))
            picked = [m for m in mums if m[1] == idx.nsamples]
            if not picked and idx.nsamples > 2:
                picked = segment(mums)
            if not picked:
                return ()
            if args.trim:
                picked = trim_overlap(picked)
                if not picked:
                    return ()
            picked.sort(key=lambda m: m[0], reverse=True)
            origin = {}
            rel = []
            for m in picked:
                r = self._lookup(m)
                rel.append(r)
                origin[tuple(r[2].values())] = m
            rel.sort(key=lambda m: (m[1], m[0]))
            rel = [m for m in rel if m[2].keys() == rel[-1][2].keys()]
            left, right = self._bounds(idx, list(rel[-1][2].keys()))
            skipleft, skipright = [], []
            if len(rel) == 1:
                split = rel[0]
            else:
                if len(rel) > args.maxmums:
                    rel = rel[-args.maxmums:]
                chained = chain(rel, left, right, gcmodel=args.gcmodel, wscore=args.wscore, wpen=args.wpen)[::-1]
                if not chained:
                    return ()
                if args.splitchain == "balanced":
                    split, opt = None, None
                    for m, _ in chained:
                        lseq = rseq = 0
                        for crd in m[2]:
                            lseq = m[2][crd]
                            rseq = right[2][crd] - m[2][crd] + m[0]
                        if opt is None or abs(lseq - rseq) < opt:
                            opt, split = abs(lseq - rseq), m
                else:
                    split = sorted(chained, key=lambda c: c[0][0])[-1][0]
                if args.seedsize > 0:
                    side, at_split = skipleft, 0
                    for m, score in chained:
                        if m == split:
                            at_split, side = score, skipright
                            continue
                        side.append((origin[tuple(m[2].values())], score - at_split))
                    skipleft = [(m, s) for m, s in skipleft if m[0] >= args.seedsize]
                    skipright = [(m, s) for m, s in skipright if m[0] >= args.seedsize]
            split = origin[tuple(split[2].values())]
            if minlength == 0:  # no length threshold: keep the anchor only if it is unlikely to occur by chance
                tests = 1
                for k in left[2]:
                    tests *= right[2][k] - left[2][k]
                p = (.25 ** (split[1] - 1)) ** split[0]
                if p > 0:
                    p = 1 - math.exp(math.log(1 - p) * tests)
                if p > args.pcutoff:
                    return ()
            return split, skipleft, skipright
        except Exception:
            log.exception("graphmumpicker failed")
            raise

    # ---- after the recursion ---------------------------------------------------------------------
    def prune_nodes(self, T=""):
        """Merges sibling nodes that spell the same sequence and have no other '+' neighbour on the shared side,
        until nothing changes (rem.py:385-446)."""
        G = self.G
        succ, pred, attrs = G._succ, G._pred, G._node   # the adjacency dicts themselves: this loop visits every node repeatedly
        multi = self.multi

        def plus(adj):
            """Neighbours over forward-forward edges, one entry per such edge."""
            if multi:
                return [v for v, keyed in adj.items() for d in keyed.values() if d["ofrom"] == "+" and d["oto"] == "+"]
            return [v for v, d in adj.items() if d["ofrom"] == "+" and d["oto"] == "+"]

        def spelled(node):
            data = attrs[node]
            if "seq" in data:
                return data["seq"]
            return T[node.begin:node.end] if isinstance(node, Interval) else None

        changed = True
        while changed:
            changed = False
            for node in list(attrs):
                if node not in attrs:
                    continue
                for forward in (True, False):
                    adj = succ[node] if forward else pred[node]
                    if len(adj) < 2:
                        continue  # fewer than two neighbours: nothing to merge on this side
                    by_seq = collections.OrderedDict()
                    for nei in plus(adj):
                        seq = spelled(nei)
                        if seq is not None:
                            by_seq.setdefault(seq, []).append(nei)
                    for group in by_seq.values():
                        if len(group) < 2:
                            continue
                        back = pred if forward else succ
                        if all(len(plus(back[v])) <= 1 for v in group):
                            for v in group[1:]:
                                if isinstance(v, Interval):
                                    self._untrack(v.begin)
                            self.mergenodes(group)
                            changed = True


class _RecursionGraph(object):
    def __init__(self, rem):
        self.rem = rem

    def __enter__(self):
        import gc
        self.gc_was_on = gc.isenabled()
        self.rem._load_core()
        # The recursion allocates millions of short-lived tuples (MUM lists) next to a large, long-lived heap; none of it is
        # cyclic garbage, so generational collections in the middle of it are pure overhead: switched off until the end.
        gc.disable()
        return self.rem

    def __exit__(self, *exc):
        import gc
        try:
            self.rem._unload_core()
        finally:
            if self.gc_was_on:
                gc.enable()
        return False


# ------------------------------------------------------------------------------------------------
def fasta_reader(fn, toupper=True, keepdash=False):
    """(name, sequence) of every record of a (gzipped) FASTA file (utils.py:79-144, default options)."""
    name, chunks = None, []
    with (gzip.open(fn, "rt") if fn.endswith(".gz") else open(fn, "r")) as f:
        for line in f:
            line = line.rstrip()
            if line.startswith(">"):
                if chunks and any(chunks):
                    yield name, "".join(chunks)
                name, chunks = line.replace(">", "").replace("\t", ""), []
            else:
                if toupper:
                    line = line.upper()
                if not keepdash:
                    line = line.replace("-", "")
                chunks.append(line)
    if chunks and any(chunks):
        yield name, "".join(chunks)


def _index_module(sa64=False):
    from . import reveallib, reveallib64
    return reveallib64 if sa64 else reveallib


def align_genomes(args, index_module=None, shard=None):
    """FASTA and GFA files -> (alignment graph, index) (rem.py:511-611).  `args`: see rem_args; `index_module` lets tests
    run the driver on another build of the drop-in extension.

    shard = (rank, world[, process group]): ONE alignment over the `world` processes of a torch.distributed job (one per GPU).
    Every rank reads the same inputs and builds the same index; the recursion is cut into units that the ranks share out
    (index.align(shard_rank=, shard_world=)); rank 0 returns the complete graph, the other ranks a graph that is only final
    for their own units.  Needs the graph in C++ (remcore) and this repository's extension."""
    mod = index_module if index_module is not None else _index_module(args.sa64)
    idx = mod.index(sa=args.sa, lcp=args.lcp, cache=args.cache)
    rem = Rem(args)
    for fn in args.inputfiles:
        if fn.endswith(".gfa") or fn.endswith(".gfa.gz"):
            rem.read_gfa(fn, idx)
        else:
            rem.read_fasta(fn, idx, contigs=args.contigs, toupper=args.toupper)
    if len(idx.samples) <= 1:
        raise ValueError("Specify at least 2 targets to construct alignment. In case of multi-fasta, consider contigs=False.")
    idx.construct()
    with rem.recursion_graph():
        mumpicker, graphalign = rem.callbacks(args.minlength)
        extra = {}
        if rem.batch_picker is not None and hasattr(mod, "align_stats"):   # this repository's extension: frontier batches
            extra["mumpicker_batch"] = rem.batch_picker
            extra["mums_as_rows"] = True    # pair MUM lists stay int64 rows between the device sweep and the native picker
        if shard is not None and shard[1] > 1:
            if rem.core is None:
                raise RuntimeError("a sharded recursion needs the compiled graph (reveal_b200.remcore)")
            rem.shard = (int(shard[0]), int(shard[1]), shard[2] if len(shard) > 2 else None, idx)
            extra.update(shard_rank=rem.shard[0], shard_world=rem.shard[1])
            if not os.environ.get("RV_REM_THREADS") and hasattr(_remcore, "set_threads"):
                # the ranks of one box share its cores: the pick threads of a rank (remcore's pool) are cut accordingly
                _remcore.set_threads(max(1, min(4, (os.cpu_count() or 2) // 2 // rem.shard[1])))
        idx.align(mumpicker, graphalign, threads=args.threads, wpen=args.wpen, wscore=args.wscore, minl=args.minlength, minn=args.minn, **extra)
    align_genomes.last_shard_stats = rem.shard_stats
    return rem.G, idx


def align(aobjs, ref=None, minlength=20, minn=2, seedsize=None, threads=0, targetsample=None, maxsamples=None, maxmums=10000, wpen=1,
          wscore=1, sa64=False, pcutoff=1e-8, gcmodel="sumofpairs", maxsize=None, trim=True, index_module=None):
    """Library entry: aligns (name, sequence) tuples, one sample each, between one shared start and end marker;
    returns (DiGraph, index) with the markers removed and equal siblings merged (rem.py:616-712)."""
    args = rem_args(minlength=minlength, minn=minn, seedsize=seedsize if seedsize is not None else 0, threads=threads, maxmums=maxmums,
                    wpen=wpen, wscore=wscore, sa64=sa64, pcutoff=pcutoff, gcmodel=gcmodel, maxsize=maxsize, trim=trim)
    mod = index_module if index_module is not None else _index_module(sa64)
    idx = mod.index()
    rem = Rem(args, nx.DiGraph())
    G = rem.G
    first, last = uuid.uuid4().hex, uuid.uuid4().hex
    G.add_node(first)
    G.add_node(last)
    for name, seq in aobjs:
        idx.addsample(name)
        begin, end = idx.addsequence(seq.upper())
        if end - begin > 0:
            node = Interval(begin, end)
            rem._track(begin, end)
            sid = len(G.graph["paths"])
            G.graph["path2id"][name] = sid
            G.graph["id2path"][sid] = name
            G.graph["id2end"][sid] = len(seq)
            G.graph["paths"].append(name)
            G.add_node(node, offsets={sid: 0}, aligned=0)
            G.add_edge(first, node, paths={sid}, ofrom="+", oto="+")
            G.add_edge(node, last, paths={sid}, ofrom="+", oto="+")
    idx.construct()
    with rem.recursion_graph():
        mumpicker, graphalign = rem.callbacks(minlength)
        extra = {"mumpicker_batch": rem.batch_picker, "mums_as_rows": True} if rem.batch_picker is not None and hasattr(mod, "align_stats") else {}
        idx.align(mumpicker, graphalign, threads=threads, wpen=wpen, wscore=wscore, minl=minlength, minn=minn, **extra)
    rem.prune_nodes(T=idx.T)
    G.remove_node(first)
    G.remove_node(last)
    return G, idx


def prune_nodes(G, T=""):
    """Module-level form of Rem.prune_nodes for graphs made by align_genomes (rem.py:385)."""
    Rem(rem_args(), G).prune_nodes(T=T)


def aligned_bases(G, idx):
    """(aligned bases, total bases, aligned nodes) as `reveal rem` reports them (rem.py:470-490)."""
    T = idx.T
    total = idx.n - T.count("$") - T.count("N")
    bases = nodes = 0
    for node, data in G.nodes(data=True):
        if isinstance(node, str) or not data.get("aligned"):
            continue
        nodes += 1
        length = node.end - node.begin
        if idx.nsamples > 2:
            bases += length * sum(1 for k in data["offsets"] if _real(G, k))
        else:
            bases += 2 * length
    return bases, total, nodes


def write_gfa(G, T, outputfile="reference.gfa", toupper=True):
    """GFA 1 with S, L and P lines; segment ids number the nodes in graph order (utils.py:710-839).  Sequence of
    aligned nodes is written upper-case (align_cmd runs seq2node(toupper=True) first, utils.py:1036-1045)."""
    if not outputfile.endswith(".gfa") and not outputfile.endswith(".gfa.gz"):
        outputfile += ".gfa.gz"
    order = [n for n in (nx.topological_sort(G) if not isinstance(G, nx.MultiDiGraph) else G.nodes()) if not isinstance(n, str)]
    ids = {n: i + 1 for i, n in enumerate(order)}
    with (gzip.open(outputfile, "wt") if outputfile.endswith(".gz") else open(outputfile, "w")) as f:
        f.write("H\tVN:Z:1.0\tCL:Z:%s\n" % " ".join(sys.argv))
        for node in order:
            data = G.nodes[node]
            seq = data["seq"] if "seq" in data else T[node.begin:node.end]
            if toupper and data.get("aligned"):
                seq = seq.upper()
            f.write("S\t%d\t%s\n" % (ids[node], seq))
            for _, to, d in G.out_edges(node, data=True):
                if not isinstance(to, str):
                    f.write("L\t%d\t%s\t%d\t%s\t%s\n" % (ids[node], d.get("ofrom", "+"), ids[to], d.get("oto", "+"), d.get("cigar", "0M")))
        for name, sid in G.graph["path2id"].items():
            walk, overlaps = [], []   # one overlap per link between two segments of the walk
            if "startnodes" not in G.graph:  # a graph of align(): no markers, the path is its nodes in offset order
                on_path = [n for n in order if sid in G.nodes[n]["offsets"]]
                walk = ["%d+" % ids[n] for n in sorted(on_path, key=lambda n: G.nodes[n]["offsets"][sid])]
                overlaps = ["0M"] * max(0, len(walk) - 1)
            for start in G.graph.get("startnodes", ()):
                if start in G and sid in G.nodes[start]["offsets"]:
                    node = start
                    while True:
                        nxt = [(v, d) for _, v, d in G.out_edges(node, data=True) if sid in d["paths"]]
                        if len(nxt) != 1:
                            log.warning("path %s stops or forks at %s", name, node)
                            break
                        prev = node
                        node, d = nxt[0]
                        if node in G.graph["endnodes"]:
                            break
                        if not isinstance(node, str):
                            walk.append("%d%s" % (ids[node], d.get("oto", "+")))
                            if not isinstance(prev, str):
                                overlaps.append(d.get("cigar", "0M"))
                    break
            f.write("P\t%s\t%s\t%s\n" % (name, ",".join(walk), ",".join(overlaps)))
    return outputfile


def align_cmd(args):
    """`reveal rem` for FASTA input: align, merge equal siblings (more than two paths), report, write the GFA
    (rem.py:448-509).  Returns (G, idx, output file)."""
    G, idx = align_genomes(args)
    if args.output is None:
        args.output = "_".join(os.path.basename(f).split(".")[0] for f in args.inputfiles) + ".gfa.gz"
    T = idx.T
    if len(G.graph["paths"]) > 2:
        prune_nodes(G, T=T)
    bases, total, nodes = aligned_bases(G, idx)
    log.info("%s (%.2f%% identity, %d bases out of %d aligned, %d nodes out of %d aligned).",
             "-".join(os.path.basename(f) for f in args.inputfiles), 100.0 * bases / max(total, 1), bases, total, nodes, G.number_of_nodes())
    out = write_gfa(G, T, outputfile=args.output)
    return G, idx, out


def main(argv=None):
    p = argparse.ArgumentParser(prog="python -m reveal_b200.rem", description="Recursive exact matching of FASTA / GFA files into a GFA graph.")
    p.add_argument("inputfiles", nargs="+")
    p.add_argument("-o", "--output", dest="output")
    p.add_argument("-t", "--threads", dest="threads", type=int, default=0)
    p.add_argument("-m", dest="minlength", type=int, default=20)
    p.add_argument("-p", dest="pcutoff", type=float, default=1e-8)
    p.add_argument("-n", dest="minn", type=int, default=2)
    p.add_argument("--gcmodel", dest="gcmodel", choices=["sumofpairs", "star-avg", "star-med"], default="sumofpairs")
    p.add_argument("--wp", dest="wpen", type=int, default=1)
    p.add_argument("--ws", dest="wscore", type=int, default=1)
    p.add_argument("--seedsize", dest="seedsize", type=int, default=10000)
    p.add_argument("--maxmums", dest="maxmums", type=int, default=1000)
    p.add_argument("--sa", dest="sa", default="")
    p.add_argument("--lcp", dest="lcp", default="")
    p.add_argument("--cache", dest="cache", default=False, action="store_true")
    p.add_argument("--64", dest="sa64", default=False, action="store_true")
    p.add_argument("--noupper", dest="toupper", action="store_false", default=True)
    p.add_argument("--maxbubblesize", dest="maxsize", type=int, default=None)
    p.add_argument("--nocontigs", dest="contigs", default=True, action="store_false")
    p.add_argument("--notrim", dest="trim", default=True, action="store_false")
    ns = p.parse_args(argv)
    logging.basicConfig(level=logging.INFO, format="%(asctime)s %(message)s")
    args = rem_args(**vars(ns))
    _, _, out = align_cmd(args)
    log.info("Graph written to: %s", out)


if __name__ == "__main__":
    main()
