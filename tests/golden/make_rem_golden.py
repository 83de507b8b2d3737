#!/usr/bin/env python
"""Generator of tests/golden/rem/*.json -- golden alignment graphs for the REM driver.
TEST INFRASTRUCTURE; runs only in the build container (it needs /root/reference).

The reference's driver (reveal/rem.py, schemes.py, utils.py) is Python 2 and there is no Python 2 here, so this
script makes a THROW-AWAY Python-3 rendering of those three files in a temporary directory (never in this
repository) with the mechanical substitutions listed in PATCHES below, gives it the reference's OWN compiled
extension (oracle/_ref, built unmodified from reveallib/*.c) as `reveallib`, a minimal `intervaltree` stand-in
(the package is absent; the driver uses add / remove / tree[pos] only), runs `align_genomes` + `prune_nodes`
exactly as `align_cmd` does (rem.py:448-462) and stores the resulting graph in a canonical, order-free form:

  nodes: [[ [[path id, offset], ...] sorted, length, aligned flag, sha1[:8] of the sequence ], ...] sorted
  edges: [[ from, to (ranks in `nodes`), ofrom+oto, [path ids] sorted ], ...] sorted
  walks: {path name: [node rank, ...]} following each path from its start node
  T_sha1, T_lower: the index text after alignment (matched bases lower-cased, reveal.c:1230-1234): digest and
         the [start, end) runs of lower-case characters

The substitutions keep Python-2 semantics where Python 3 would differ: integer `/`, `x > None` being true,
dicts with small int keys iterating in ascending key order.
"""
import argparse
import gzip
import hashlib
import json
import logging
import os
import re
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("REVEAL_REFERENCE_ROOT", "/root/reference")

PATCHES = [
    (r"print traceback\.format_exc\(\)", "print(traceback.format_exc())"),
    (r"\.node\[", ".nodes["),                                   # networkx >= 2
    (r"mums\[0\]\[2\]\.keys\(\)\[0\]", "sorted(mums[0][2].keys())[0]"),   # py2: small-int dict keys come out ascending
    (r"iter\(t\)\.next\(\)", "next(iter(t))"),
    (r"\)/2\)", ")//2)"),                                       # py2 integer division (schemes.py:71)
    (r"len\(chainedmums\)/2", "len(chainedmums)//2"),           # schemes.py:349-351
    (r"len\(pointa\)/2", "len(pointa)//2"),                     # utils.py:169
    (r"\)\)/len\(pointa\)", "))//len(pointa)"),                  # utils.py:166 (star-avg): integer division of integers
    (r"if tmpw>w or w==None:", "if w==None or tmpw>w:"),        # py2: int > None is True
    (r"xrange", "range"),
    (r"^class IntervalPatched\(intervaltree\.Interval\):", "class IntervalPatched(intervaltree.Interval):\n    __hash__=intervaltree.Interval.__hash__"),  # py3 drops __hash__ when __eq__ is defined
    (r"for node,data in G\.nodes\(data=True\):", "for node,data in list(G.nodes(data=True)):"),  # networkx 1 returned a list (rem.py:389)
    (r"f=fopen\(outputfile,'wb'\)", "f=fopen(outputfile,'w')"),   # write_gfa writes str (utils.py:720); plain .gfa only here
    (r" or type\(G\)==nx\.classes\.graphviews\.Sub(Multi)?DiGraph", ""),   # class names of networkx 2.0 (utils.py:730-733)
    (r"^import bubbles$", "bubbles=None"),                      # not used on this path
    (r"^import intervaltree$", "import rv_intervaltree as intervaltree"),
    (r"^from intervaltree import", "from rv_intervaltree import"),
]

INTERVALTREE = '''
import bisect
import collections


_Base = collections.namedtuple("Interval", ["begin", "end", "data"])


class Interval(_Base):
    __slots__ = ()

    def __new__(cls, begin, end, data=None):
        return _Base.__new__(cls, begin, end, data)


class IntervalTree(object):
    """add / remove / tree[pos] over intervals that are disjoint whenever tree[pos] is asked (graph nodes never
    overlap in index coordinates; breaknode adds the pieces of a node before it removes the node)."""

    def __init__(self):
        self.keys = []

    def add(self, iv):
        bisect.insort(self.keys, (iv.begin, iv.end))

    def remove(self, iv):
        i = bisect.bisect_left(self.keys, (iv.begin, iv.end))
        assert self.keys[i] == (iv.begin, iv.end)
        del self.keys[i]

    def __getitem__(self, pos):
        i = bisect.bisect_right(self.keys, (pos, float("inf"))) - 1
        if i >= 0 and self.keys[i][0] <= pos < self.keys[i][1]:
            return {Interval(self.keys[i][0], self.keys[i][1])}
        return set()
'''

SHIM = '''
import oracle.ref as _R
_m = _R.module(%d)
index = _m.index
error = _m.error
'''


def render(tmp):
    for name in ("rem", "schemes", "utils"):
        src = open(os.path.join(REF, "reveal", name + ".py")).read()
        for pat, rep in PATCHES:
            src = re.sub(pat, rep, src, flags=re.M)
        open(os.path.join(tmp, name + ".py"), "w").write(src)
    open(os.path.join(tmp, "rv_intervaltree.py"), "w").write(INTERVALTREE)
    open(os.path.join(tmp, "reveallib.py"), "w").write(SHIM % 32)
    open(os.path.join(tmp, "reveallib64.py"), "w").write(SHIM % 64)


def default_args(inputfiles, **kw):
    """The defaults of `reveal rem` (reveal/reveal.py:74-99)."""
    a = dict(inputfiles=list(inputfiles), output=None, threads=0, minlength=20, pcutoff=1e-8, minn=2, gcmodel="sumofpairs", wpen=1,
             wscore=1, seedsize=10000, maxmums=1000, mumplot=False, interactive=False, sa="", lcp="", cache=False, minsamples=1,
             maxsamples=None, reference=None, targetsample=None, gml=False, hwm=4000, toupper=True, maxsize=None, contigs=True,
             trim=True, sa64=False)
    a.update(kw)
    return argparse.Namespace(**a)


def canonical(G, T):
    """Order-free description of the alignment graph (see the module docstring)."""
    def key(node):
        return tuple(sorted((int(k), int(v)) for k, v in G.nodes[node]["offsets"].items()))
    nodes = {}
    for node, data in G.nodes(data=True):
        if isinstance(node, str):
            continue
        seq = data["seq"] if "seq" in data else T[node.begin:node.end]
        nodes[key(node)] = (len(seq), int(data.get("aligned", 0)), hashlib.sha1(seq.encode()).hexdigest()[:8])
    order = sorted(nodes)
    rank = {k: i for i, k in enumerate(order)}
    edges = []
    for u, v, d in G.edges(data=True):
        if isinstance(u, str) or isinstance(v, str):
            continue
        edges.append([rank[key(u)], rank[key(v)], d["ofrom"] + d["oto"], sorted(int(p) for p in d["paths"])])
    walks = {}
    for name, sid in G.graph["path2id"].items():
        walk = []
        if "startnodes" not in G.graph:  # rem.align(): markers already removed -- the path in offset order
            on_path = [n for n, d in G.nodes(data=True) if not isinstance(n, str) and sid in d["offsets"]]
            walks[name] = [rank[key(n)] for n in sorted(on_path, key=lambda n: G.nodes[n]["offsets"][sid])]
            continue
        for start in G.graph["startnodes"]:
            if sid in G.nodes[start]["offsets"]:
                node = start
                while True:
                    out = [(v, d) for _, v, d in G.out_edges(node, data=True) if sid in d["paths"]]
                    assert len(out) == 1, (name, node, out)
                    node = out[0][0]
                    if node in G.graph["endnodes"]:
                        break
                    if not isinstance(node, str):
                        walk.append(rank[key(node)])
                break
        walks[name] = walk
    lower = [[m.start(), m.end()] for m in re.finditer(r"[a-z]+", T)]
    return {"nodes": [[[list(kv) for kv in k], nodes[k][0], nodes[k][1], nodes[k][2]] for k in order], "edges": sorted(edges),
            "walks": walks, "T_sha1": hashlib.sha1(T.encode()).hexdigest(), "T_lower": lower, "n": len(T)}


CASES = [  # name, inputs (reference test files, or ("synth", n_genomes, length, seed)), overrides of the `rem` defaults
    ("1a_1b", ["1a.fa", "1b.fa"], {}),                                    # BASELINE configs[0]
    ("1a_1b_m15", ["1a.fa", "1b.fa"], {"minlength": 15}),
    ("1a_1b_notrim_noseed", ["1a.fa", "1b.fa"], {"trim": False, "seedsize": 0}),
    ("1a_1b_seed50", ["1a.fa", "1b.fa"], {"seedsize": 50}),               # precomputed chains handed to the children
    ("1a_1b_maxmums5", ["1a.fa", "1b.fa"], {"maxmums": 5}),
    ("1a_1b_1c", ["1a.fa", "1b.fa", "1c.fa"], {}),
    ("1a_1b_1c_n3", ["1a.fa", "1b.fa", "1c.fa"], {"minn": 3}),
    ("1a_1c_1d_1e", ["1a.fa", "1c.fa", "1d.fa", "1e.fa"], {"minlength": 15}),
    ("t1_t2", ["t1.fa", "t2.fa"], {"minlength": 5}),
    ("d1_d2", ["d1.fa", "d2.fa"], {"minlength": 10}),
    ("1e_1f_nocontigs", ["1e.fa", "1f.fa"], {"contigs": False, "minlength": 12}),
    ("synth2_4k", ("synth", 2, 4000, 21), {"minlength": 12}),           # small enough for the emulated kernels (CPU tests)
    ("synth3_3k", ("synth", 3, 3000, 22), {"minlength": 10}),
    ("synth4_2k_seed", ("synth", 4, 2000, 23), {"minlength": 8, "seedsize": 30, "minn": 3}),
    # option branches of graphmumpicker / chain on small inputs
    ("synth2_4k_m0_pvalue", ("synth", 2, 4000, 24), {"minlength": 0}),            # significance cut-off instead of a length
    ("synth3_3k_maxsize", ("synth", 3, 3000, 25), {"minlength": 10, "maxsize": 200}),
    ("synth3_3k_star_avg", ("synth", 3, 3000, 26), {"minlength": 10, "gcmodel": "star-avg"}),
    ("synth3_3k_star_med", ("synth", 3, 3000, 27), {"minlength": 10, "gcmodel": "star-med", "wpen": 3, "wscore": 2}),
    ("1c_1d_noupper", ["1c.fa", "1d.fa"], {"toupper": False}),                  # 1d.fa is lower-case: no matches without upper-casing
    # the library entry rem.align() on (name, sequence) tuples: plain DiGraph, shared start / end marker, prune_nodes
    ("align_api_3x3k", ("align", 3, 3000, 41), {"minlength": 10, "seedsize": 0}),
    ("align_api_2x4k", ("align", 2, 4000, 42), {"minlength": 12, "seedsize": 0, "trim": False}),
    # graph input (utils.read_gfa): the inputs are graphs the reference driver itself wrote from synthetic genomes
    ("gfa_x_gfa_4x20k", ("graphs", 4, 20000, 31, [[0, 1], [2, 3]]), {}),
    ("gfa_x_fasta_3x10k", ("graphs", 3, 10000, 32, [[0, 1], 2]), {"minlength": 15}),
    ("gfa3_x_gfa2_5x3k", ("graphs", 5, 3000, 33, [[0, 1, 2], [3, 4]]), {"minlength": 10}),   # small: emulated kernels
    # a hand-made graph in which one path runs through two segments on the reverse strand ('-' links in read_gfa, paths that
    # are not colinear).  The reference aligns the colinear flanks only: anchors in the inverted middle cost more gap penalty
    # than they score.  (Forcing them in -- wpen=0, or a single reversed segment with parallel +/- links -- crashes the
    # reference itself: `best` unbound in schemes.chain, a segmentation fault in the C aligner.)
    ("gfa_reverse_strand_x_fasta", ("revgraph", 6000, 51), {"minlength": 12}),
    ("synth2_200k", ("synth", 2, 200000, 11), {}),
    ("synth3_60k", ("synth", 3, 60000, 12), {}),
    ("synth5_30k_n3", ("synth", 5, 30000, 13), {"minn": 3, "minlength": 15}),
    ("synth2_1m", ("synth", 2, 1000000, 14), {}),
]
INPUTS = {}  # reference test file -> [[contig name, sequence], ...], written once to rem/inputs.json.gz
DIGEST_ABOVE = 4000  # graphs with more nodes are stored as digests of the canonical lists


def digest(obj):
    return hashlib.sha1(json.dumps(obj, separators=(",", ":"), sort_keys=True).encode()).hexdigest()


def write_fasta(path, name, seq):
    with open(path, "w") as f:
        f.write(">%s\n" % name)
        for i in range(0, len(seq), 100):
            f.write(seq[i:i + 100] + "\n")


def run_case(rem, tmp, inputs, overrides):
    out = {}
    if isinstance(inputs, tuple) and inputs[0] == "revgraph":
        from reveal_b200 import synth
        _, length, seed = inputs
        g0, g1 = [g.tobytes().decode() for g in synth.genomes(2, length, seed=seed)]
        cuts = [0, length // 4, length // 2, 3 * length // 4, length]
        segs = [g0[cuts[i]:cuts[i + 1]] for i in range(4)]
        gfa = "H\tVN:Z:1.0\n" + "".join("S\t%d\t%s\n" % (i + 1, sq) for i, sq in enumerate(segs))
        gfa += "L\t1\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nL\t3\t+\t4\t+\t0M\n"      # path fwd: 1+ 2+ 3+ 4+
        gfa += "L\t1\t+\t3\t-\t0M\nL\t3\t-\t2\t-\t0M\nL\t2\t-\t4\t+\t0M\n"      # path inv: 1+ 3- 2- 4+ (middle inverted)
        gfa += "P\tfwd\t1+,2+,3+,4+\t0M,0M,0M\nP\tinv\t1+,3-,2-,4+\t0M,0M,0M\n"
        gpath, fpath = os.path.join(tmp, "%s_rev.gfa" % seed), os.path.join(tmp, "%s_q.fa" % seed)
        open(gpath, "w").write(gfa)
        write_fasta(fpath, "query", g1)
        out["files"] = [["rev.gfa", gfa], ["q.fa", open(fpath).read()]]
        files = [gpath, fpath]
    elif isinstance(inputs, tuple) and inputs[0] == "align":
        from reveal_b200 import synth
        _, ng, length, seed = inputs
        aobjs = [("g%d" % k, g.tobytes().decode()) for k, g in enumerate(synth.genomes(ng, length, seed=seed))]
        G, idx = rem.align(aobjs, **overrides)
        out["align"] = [ng, length, seed]
        out.update(canonical(G, idx.T))
        out["counts"] = [len(out["nodes"]), len(out["edges"]), sum(n[2] != 0 for n in out["nodes"])]
        out["aligned_bases"] = sum(n[1] * len(n[0]) for n in out["nodes"] if n[2] != 0)
        out["args"] = dict(overrides)
        return out
    elif isinstance(inputs, tuple) and inputs[0] == "graphs":
        from reveal_b200 import synth
        _, ng, length, seed, groups = inputs
        fastas = []
        for k, g in enumerate(synth.genomes(ng, length, seed=seed)):
            fastas.append(os.path.join(tmp, "%s_g%d.fa" % (seed, k)))
            write_fasta(fastas[-1], "g%d" % k, g.tobytes().decode())
        files, texts = [], []
        for gi, group in enumerate(groups):
            if isinstance(group, list):  # align the group first and write its graph as the reference does (align_cmd)
                G, idx = rem.align_genomes(default_args([fastas[k] for k in group]))
                T = idx.T
                if len(G.graph["paths"]) > 2:
                    rem.prune_nodes(G, T=T)
                rem.seq2node(G, T, remap=False)
                path = os.path.join(tmp, "%s_graph%d.gfa" % (seed, gi))
                rem.write_gfa(G, T, outputfile=path)
                files.append(path)
                text = re.sub(r"^H\t[^\n]*\n", "H\tVN:Z:1.0\n", open(path).read())   # drop CL:Z:<command line> from the header
                open(path, "w").write(text)
                texts.append(["graph%d.gfa" % gi, text])
            else:
                files.append(fastas[group])
                texts.append(["g%d.fa" % group, open(fastas[group]).read()])
        out["files"] = texts  # the input files themselves (graphs as the reference wrote them)
    elif isinstance(inputs, tuple):
        from reveal_b200 import synth
        _, ng, length, seed = inputs
        files = []
        for k, g in enumerate(synth.genomes(ng, length, seed=seed)):
            files.append(os.path.join(tmp, "g%d.fa" % k))
            write_fasta(files[-1], "g%d" % k, g.tobytes().decode())
        out["synth"] = [ng, length, seed]
    else:
        files = [os.path.join(REF, "tests", f) for f in inputs]
        # the inputs travel with the golden: /root/reference does not exist where the GPU tests run
        out["inputs"] = list(inputs)
        for f in inputs:
            if f not in INPUTS:
                INPUTS[f] = [[name, seq] for name, seq in rem.fasta_reader(os.path.join(REF, "tests", f), toupper=False)]
    args = default_args(files, **overrides)
    G, idx = rem.align_genomes(args)
    T = idx.T
    if len(G.graph["paths"]) > 2:
        rem.prune_nodes(G, T=T)  # align_cmd, rem.py:460-461
    out.update(canonical(G, T))
    out["counts"] = [len(out["nodes"]), len(out["edges"]), sum(n[2] != 0 for n in out["nodes"])]
    out["aligned_bases"] = sum(n[1] * len(n[0]) for n in out["nodes"] if n[2] != 0)
    if len(out["nodes"]) > DIGEST_ABOVE:
        for k in ("nodes", "edges", "walks", "T_lower"):
            out[k + "_sha1"] = digest(out.pop(k))
    out["args"] = dict(overrides)
    return out


def main():
    sys.path.insert(0, ROOT)
    logging.TRACE = 1
    logging.trace = lambda msg, *a, **k: logging.log(1, msg, *a, **k)   # reveal/reveal.py:34-39
    logging.basicConfig(level=logging.WARNING)
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        render(tmp)
        sys.path.insert(0, tmp)
        import rem
        outdir = os.path.join(HERE, "rem")
        os.makedirs(outdir, exist_ok=True)
        for name, files, overrides in CASES:
            if only and name not in only:
                continue
            res = run_case(rem, tmp, files, overrides)
            dump(os.path.join(outdir, name + ".json.gz"), res)
            print("%-24s %6d nodes %6d edges, %6d aligned nodes, %d aligned bases" % ((name,) + tuple(res["counts"]) + (res["aligned_bases"],)))
        ipath = os.path.join(outdir, "inputs.json.gz")
        if os.path.exists(ipath):
            have = json.loads(gzip.open(ipath).read())
            have.update(INPUTS)
            INPUTS.update(have)
        dump(ipath, INPUTS)


def dump(path, obj):
    with gzip.GzipFile(path, "wb", mtime=0) as f:
        f.write(json.dumps(obj, separators=(",", ":"), sort_keys=True).encode())


if __name__ == "__main__":
    main()
