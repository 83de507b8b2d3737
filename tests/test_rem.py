"""The REM driver (reveal_b200/rem.py) against alignment graphs minted from the reference's own driver running on the
reference's own compiled extension (tests/golden/make_rem_golden.py): same nodes (per-path offsets, length, sequence,
aligned flag), same edges with orientations and paths, same walk of every path, same lower-cased text."""
import gzip
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_rem_golden as M  # noqa: E402  (canonical form + case table; it touches /root/reference only when run as a script)

from reveal_b200 import rem, synth  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "rem")


def load(name):
    return json.loads(gzip.open(os.path.join(GOLDEN, name + ".json.gz")).read())


def case_files(gold, tmp_path):
    files = []
    if "files" in gold:  # graph input: the files themselves travel with the golden
        for fn, text in gold["files"]:
            files.append(str(tmp_path / fn))
            with open(files[-1], "w") as f:
                f.write(text)
        return files
    if "synth" in gold:
        ng, length, seed = gold["synth"]
        for k, g in enumerate(synth.genomes(ng, length, seed=seed)):
            files.append(str(tmp_path / ("g%d.fa" % k)))
            M.write_fasta(files[-1], "g%d" % k, g.tobytes().decode())
        return files
    inputs = load("inputs")
    for fn in gold["inputs"]:
        files.append(str(tmp_path / fn))
        with open(files[-1], "w") as f:
            for name, seq in inputs[fn]:
                f.write(">%s\n%s\n" % (name, seq))
    return files


def run_case(name, tmp_path, index_module, shard=None):
    gold = load(name)
    if "align" in gold:  # the library entry on (name, sequence) tuples
        ng, length, seed = gold["align"]
        aobjs = [("g%d" % k, g.tobytes().decode()) for k, g in enumerate(synth.genomes(ng, length, seed=seed))]
        G, idx = rem.align(aobjs, index_module=index_module, **gold["args"])
        T = idx.T
    else:
        args = rem.rem_args(case_files(gold, tmp_path), **gold["args"])
        G, idx = rem.align_genomes(args, index_module=index_module, shard=shard)
        if shard is not None and shard[0] != 0:
            return G, idx   # only rank 0 holds the complete graph of a sharded recursion
        T = idx.T
        if len(G.graph["paths"]) > 2:
            rem.prune_nodes(G, T=T)
    got = M.canonical(G, T)
    assert [len(got["nodes"]), len(got["edges"]), sum(n[2] != 0 for n in got["nodes"])] == gold["counts"]
    assert sum(n[1] * len(n[0]) for n in got["nodes"] if n[2] != 0) == gold["aligned_bases"]
    # the reference's own report (rem.py:470-490): length x paths for more than two index samples, 2 x length else
    want = gold["aligned_bases"] if idx.nsamples > 2 else 2 * sum(n[1] for n in got["nodes"] if n[2] != 0)
    assert rem.aligned_bases(G, idx)[0] == want
    assert got["T_sha1"] == gold["T_sha1"]
    for k in ("nodes", "edges", "walks", "T_lower"):
        if k in gold:
            assert got[k] == gold[k], k
        else:
            assert M.digest(got[k]) == gold[k + "_sha1"], k
    return G, idx


def test_chain_matches_pairwise_definition():
    """The vectorised chain against a direct O(m^2) evaluation of the same recurrence on random anchors."""
    rng = np.random.default_rng(3)
    for trial in range(30):
        k = int(rng.integers(2, 5))
        m = int(rng.integers(1, 40))
        mums = []
        for _ in range(m):
            base = int(rng.integers(0, 2000))
            mums.append((int(rng.integers(5, 60)), k, {s: base + int(rng.integers(-40, 40)) + 100 * s for s in range(k)}))
        if len({mm[2][0] for mm in mums}) < m:
            continue
        left = (0, 0, {s: -1000 + 100 * s for s in range(k)})
        right = (0, 0, {s: 4000 + 100 * s for s in range(k)})
        got = rem.chain(list(mums), left, right)
        # direct evaluation
        order = sorted(mums + [right], key=lambda x: x[2][0])
        score = {id(left): 0}
        link = {}
        done = [left]
        for mm in order:
            best = None
            for a in done:
                if all(a[2][c] + a[0] <= mm[2][c] for c in mm[2]):
                    w = score[id(a)] + mm[0] * (mm[1] * (mm[1] - 1) // 2) - rem.gapcost([a[2][c] + a[0] for c in mm[2]], [mm[2][c] for c in mm[2]])
                    if best is None or w > best[0]:
                        best = (w, a)
            score[id(mm)], link[id(mm)] = best
            done.append(mm)
        assert (got[0][1] if got else None) == (score[id(link[id(right)])] if link[id(right)] is not left else None)
        total = score[id(right)]
        # the returned chain is colinear and adds up to the best total
        chain = [c for c, _ in got][::-1]
        for a, b in zip(chain, chain[1:]):
            assert all(a[2][c] + a[0] <= b[2][c] for c in a[2])
        acc, prev = 0, left
        for c in chain + [right]:
            acc += c[0] * (c[1] * (c[1] - 1) // 2) - rem.gapcost([prev[2][x] + prev[0] for x in c[2]], [c[2][x] for x in c[2]])
            prev = c
        assert acc == total


@pytest.mark.parametrize("model", ["sumofpairs", "star-avg", "star-med"])
def test_native_chain_equals_numpy_chain(model):
    """chain_dp of the compiled extension against its numpy twin: same links and scores, ties included."""
    if rem._native_chain is None:
        pytest.skip("extension module not built")
    rng = np.random.default_rng(9)
    for trial in range(60):
        k = int(rng.integers(1, 6))
        m = int(rng.integers(1, 120))
        start = np.sort(rng.integers(0, 300 if trial % 2 else 5000, size=(m + 1, 1)), axis=0) + rng.integers(-30, 30, size=(m + 1, k))
        start[0] = -100
        start[m] = 10000
        length = rng.integers(1, 40, size=m + 1)
        length[0] = length[m] = 0
        gain = rng.integers(0, 50, size=m + 1) * (1 if trial % 3 else 0)   # every third trial: all gains 0 -> many ties
        out = []
        for fn in (lambda *a: rem._native_chain(a[0], a[1], a[2], 2, rem._MODELS[model], a[3], a[4]),
                   lambda *a: rem._chain_numpy(a[0], a[1], a[2], 2, model, a[3], a[4])):
            link = np.zeros(m + 1, dtype=np.int64)
            score = np.zeros(m + 1, dtype=np.int64)
            fn(np.ascontiguousarray(start, dtype=np.int64), np.ascontiguousarray(length, dtype=np.int64),
               np.ascontiguousarray(gain, dtype=np.int64), link, score)
            out.append((link.tolist(), score.tolist()))
        assert out[0] == out[1]


def test_trim_overlap_and_gapcost_small_cases():
    a = (10, 2, ((0, 0), (1, 100)))
    b = (10, 2, ((0, 5), (1, 105)))      # overlaps a by 5 in both samples
    c = (4, 2, ((0, 2), (1, 300)))       # contained in a in sample 0
    out = rem.trim_overlap([a, b])
    assert sorted(out) == sorted([(5, 2, ((0, 0), (1, 100))), (5, 2, ((0, 10), (1, 110)))])
    # the reference's containment filter (schemes.py:171) judges the FIRST anchor by its successor: with c right
    # behind it, a itself is dropped together with c
    assert rem.trim_overlap([a, b, c]) == [b]
    assert rem.gapcost([0, 0, 0], [3, 5, 10]) == 2 + 7 + 5
    assert rem.gapcost([0, 0], [4, 9], model="star-med") == 9
    assert rem.gapcost([1, 2], [4, 9], model="star-avg") == 5


@pytest.mark.parametrize("name", ["t1_t2", "synth2_4k", "synth3_3k", "synth4_2k_seed", "gfa3_x_gfa2_5x3k", "synth2_4k_m0_pvalue"])
def test_rem_emulated_small(emu_reveallib, tmp_path, name):
    run_case(name, tmp_path, emu_reveallib.mod32)


@pytest.mark.parametrize("graph", ["native", "python"])
@pytest.mark.parametrize("name", [c[0] for c in M.CASES])
def test_rem_driver_on_reference_extension(tmp_path, monkeypatch, name, graph):
    """The driver alone: run on the reference's own compiled extension (oracle/_ref) it must rebuild the golden
    graphs exactly, whatever the size -- separates driver parity from index parity.  Once with the graph of the
    recursion in C++ (remcore.Graph) and once on the networkx graph itself (the Python twin of the same methods)."""
    import oracle.ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    if graph == "python":
        monkeypatch.setenv("RV_REM_PYTHON_GRAPH", "1")
        if name == "synth2_1m":
            pytest.skip("largest case: native graph only (suite time)")
    else:
        assert rem._remcore is not None, "reveal_b200/remcore extension module not built"
    run_case(name, tmp_path, R.module(32))


def test_rem_emulated_wide_index_module(emu_reveallib, tmp_path):
    """The same driver on reveallib64 (wider integers on the way out of the extension)."""
    run_case("synth3_3k", tmp_path, emu_reveallib.mod64)


@pytest.mark.gpu
@pytest.mark.parametrize("name", [c[0] for c in M.CASES])
def test_rem_matches_reference_graph_on_gpu(tmp_path, name):
    from reveal_b200 import reveallib
    run_case(name, tmp_path, reveallib)


@pytest.mark.gpu
def test_rem_gfa_spells_the_inputs(tmp_path):
    """write_gfa: walking every P line over the S lines gives back the input sequence."""
    from reveal_b200 import reveallib
    gold = load("1a_1b_1c")
    files = case_files(gold, tmp_path)
    args = rem.rem_args(files, output=str(tmp_path / "out.gfa"))
    G, idx, out = rem.align_cmd(args)
    seg, paths = {}, {}
    for line in open(out):
        f = line.rstrip("\n").split("\t")
        if f[0] == "S":
            seg[f[1]] = f[2]
        elif f[0] == "P":
            paths[f[1]] = [x[:-1] for x in f[2].split(",")]
    want = {}
    for fn in files:
        for name, seq in rem.fasta_reader(fn):
            want[name] = seq
    assert set(paths) == set(want)
    for name in want:
        assert "".join(seg[s] for s in paths[name]).upper() == want[name]


def test_rem_inconsistent_intervals_are_refused(emu_reveallib, tmp_path):
    """A graph whose paths run through one segment on both strands over parallel links: graphalign hands back
    leading intervals that are not part of the sub-index.  The reference's C aligner writes out of bounds on this
    input (segmentation fault); here the step is refused with reveallib.error."""
    length = 3000
    g0, g1 = [g.tobytes().decode() for g in synth.genomes(2, length, seed=51)]
    cuts = [0, length // 3, 2 * length // 3, length]
    gfa = "H\tVN:Z:1.0\n" + "".join("S\t%d\t%s\n" % (i + 1, g0[cuts[i]:cuts[i + 1]]) for i in range(3))
    gfa += "L\t1\t+\t2\t+\t0M\nL\t2\t+\t3\t+\t0M\nL\t1\t+\t2\t-\t0M\nL\t2\t-\t3\t+\t0M\n"
    gfa += "P\tfwd\t1+,2+,3+\t0M,0M\nP\tinv\t1+,2-,3+\t0M,0M\n"
    (tmp_path / "rev.gfa").write_text(gfa)
    (tmp_path / "q.fa").write_text(">q\n%s\n" % g1)
    args = rem.rem_args([str(tmp_path / "rev.gfa"), str(tmp_path / "q.fa")], minlength=12)
    with pytest.raises(emu_reveallib.error, match="intervals cover"):
        rem.align_genomes(args, index_module=emu_reveallib.mod32)


def test_rem_native_graph_equals_python_graph_fuzz(tmp_path, monkeypatch):
    """Random inputs and option mixes: the recursion on remcore.Graph (C++ graphalign + pick) and on the networkx
    graph (the Python twin) must give the same canonical graph.  Runs on the reference extension (fast, CPU)."""
    import oracle.ref as R
    if not R.available() or rem._remcore is None:
        pytest.skip("needs oracle/_ref and the remcore module")
    rng = np.random.default_rng(2024)
    for trial in range(30):
        ng = int(rng.integers(2, 5))
        length = int(rng.integers(1500, 6000))
        files = []
        for k, g in enumerate(synth.genomes(ng, length, seed=1000 + trial, snp=float(rng.choice([0.01, 0.03])), indel=0.002)):
            files.append(str(tmp_path / ("t%d_g%d.fa" % (trial, k))))
            M.write_fasta(files[-1], "g%d" % k, g.tobytes().decode())
        opts = dict(minlength=int(rng.integers(8, 20)), minn=int(rng.integers(2, ng + 1)), trim=bool(rng.integers(0, 2)),
                    seedsize=int(rng.choice([0, 15, 10000])), maxmums=int(rng.choice([2, 5, 1000])), wpen=int(rng.integers(0, 4)),
                    wscore=int(rng.integers(1, 4)), gcmodel=str(rng.choice(["sumofpairs", "star-avg", "star-med"])))
        got = []
        for mode in ("0", "1"):
            monkeypatch.setenv("RV_REM_PYTHON_GRAPH", mode)
            G, idx = rem.align_genomes(rem.rem_args(files, **opts), index_module=R.module(32))
            T = idx.T
            if ng > 2:
                rem.prune_nodes(G, T=T)
            got.append(M.canonical(G, T))
        assert got[0] == got[1], (trial, opts)
        assert sum(n[2] != 0 for n in got[0]["nodes"]) > 0 or opts["minlength"] > 15, (trial, opts)


def test_rem_driver_equals_reference_driver_fuzz(tmp_path):
    """Build container only (needs /root/reference): the reference's own driver, rendered for Python 3 into a
    temporary directory exactly as tests/golden/make_rem_golden.py does, against this driver on random genomes and
    option mixes -- both on the reference's compiled extension."""
    import importlib
    import logging
    import oracle.ref as R
    if not os.path.isdir(os.path.join(M.REF, "reveal")) or not R.available():
        pytest.skip("reference tree not present")
    rendered = tmp_path / "rendered"
    rendered.mkdir()
    M.render(str(rendered))
    logging.TRACE = 1
    logging.trace = lambda msg, *a, **k: logging.log(1, msg, *a, **k)
    sys.path.insert(0, str(rendered))
    sys.path.insert(0, os.path.dirname(HERE))
    try:
        refrem = importlib.import_module("rem")
        rng = np.random.default_rng(77)
        for trial in range(25):
            ng = int(rng.integers(2, 5))
            length = int(rng.integers(1500, 8000))
            files = []
            for k, g in enumerate(synth.genomes(ng, length, seed=2000 + trial, snp=float(rng.choice([0.01, 0.04])), indel=0.002)):
                files.append(str(tmp_path / ("r%d_g%d.fa" % (trial, k))))
                M.write_fasta(files[-1], "g%d" % k, g.tobytes().decode())
            opts = dict(minlength=int(rng.integers(8, 20)), minn=int(rng.integers(2, ng + 1)), trim=bool(rng.integers(0, 2)),
                        seedsize=int(rng.choice([0, 15, 10000])), maxmums=int(rng.choice([2, 5, 1000])), wpen=int(rng.integers(1, 4)),
                        wscore=int(rng.integers(1, 4)), gcmodel=str(rng.choice(["sumofpairs", "star-avg", "star-med"])))
            G1, i1 = refrem.align_genomes(M.default_args(files, **opts))
            T1 = i1.T
            G2, i2 = rem.align_genomes(rem.rem_args(files, **opts), index_module=R.module(32))
            T2 = i2.T
            if ng > 2:
                refrem.prune_nodes(G1, T=T1)
                rem.prune_nodes(G2, T=T2)
            assert M.canonical(G1, T1) == M.canonical(G2, T2), (trial, opts)
            if trial % 3 == 0 and ng > 2:
                # graph input: the graph the reference just made (all genomes but the last) against the last genome
                Gg, ig = refrem.align_genomes(M.default_args(files[:-1], **opts))
                Tg = ig.T
                if ng - 1 > 2:
                    refrem.prune_nodes(Gg, T=Tg)
                refrem.seq2node(Gg, Tg, remap=False)
                gfa = str(tmp_path / ("r%d.gfa" % trial))
                refrem.write_gfa(Gg, Tg, outputfile=gfa)
                pair = [gfa, files[-1]]
                o2 = dict(opts, minn=2)
                G1, i1 = refrem.align_genomes(M.default_args(pair, **o2))
                G2, i2 = rem.align_genomes(rem.rem_args(pair, **o2), index_module=R.module(32))
                assert M.canonical(G1, i1.T) == M.canonical(G2, i2.T), (trial, "graph input", o2)
    finally:
        sys.path.remove(str(rendered))
        for name in ("rem", "schemes", "utils", "reveallib", "reveallib64", "rv_intervaltree"):
            sys.modules.pop(name, None)


def test_remcore_refuses_bad_input_without_crashing():
    """The C++ graph reports misuse as Python exceptions."""
    if rem._remcore is None:
        pytest.skip("remcore module not built")
    g = rem._remcore.Graph(True, rem.Interval, [True, True], [100, 100])
    g.add_node(rem.Interval(0, 100), 0, {0: 0}, None)
    g.add_node(rem.Interval(101, 201), 0, {1: 0}, None)
    g.add_node("start", None, {0: 0, 1: 0}, {"endpoint": True})
    g.add_edge("start", rem.Interval(0, 100), "+", "+", {0}, None)
    g.add_edge("start", (101, 201), "+", "+", {1}, {"cigar": "0M"})
    assert g.stats()[:2] == (3, 2)
    assert g.coords(150) == ((1, 49),) and g.node_offsets((0, 100)) == {0: 0}
    with pytest.raises(KeyError):
        g.coords(100)                       # the separator between the two sequences belongs to no node
    with pytest.raises(KeyError):
        g.add_edge("nowhere", (0, 100), "+", "+", {0}, None)
    with pytest.raises(KeyError):
        g.node_offsets((5, 100))
    nodes = {(0, 100), (101, 201)}
    with pytest.raises(KeyError):
        g.graphalign(nodes, None, None, 20, [90, 150])      # [90, 110) does not fit into its node
    with pytest.raises(TypeError):
        g.graphalign([(0, 100)], None, None, 20, [10, 150])  # nodes must be the set of the sub-index
    with pytest.raises(TypeError):
        g.pick([(10, 2)], 2, None, None, True, 1000, 0, 1, 1, 0)
    assert g.pick([], 2, None, None, True, 1000, 0, 1, 1, 0) == ()
    # a proper step still works afterwards
    lead, trail, matching, rest, merged, newleft, newright = g.graphalign(nodes, None, None, 20, [10, 111])
    assert merged == rem.Interval(10, 30) and matching == {(10, 30), (111, 131)}
    assert lead == {(0, 10), (101, 111)} and trail == {(30, 100), (131, 201)} and rest == set()
    assert nodes == {(0, 10), (30, 100), (101, 111), (131, 201)} and newleft == merged and newright == merged
    n, e = g.export()
    assert len(n) == 6 and {k for k, _ in n} >= {rem.Interval(10, 30), "start"}
    assert [a for k, a in n if k == rem.Interval(10, 30)][0] == {"offsets": {0: 10, 1: 10}, "aligned": 1}


def test_write_gfa_equals_reference_writer(tmp_path):
    """Build container only: on the SAME graph object the reference's seq2node + write_gfa (rendered for Python 3) and
    this repository's write_gfa must produce the same file (header line aside: it carries the command line)."""
    import importlib
    import logging
    import oracle.ref as R
    if not os.path.isdir(os.path.join(M.REF, "reveal")) or not R.available():
        pytest.skip("reference tree not present")
    rendered = tmp_path / "rendered"
    rendered.mkdir()
    M.render(str(rendered))
    logging.TRACE = 1
    logging.trace = lambda msg, *a, **k: logging.log(1, msg, *a, **k)
    sys.path.insert(0, str(rendered))
    try:
        refutils = importlib.import_module("utils")
        for name in ("1a_1b", "1a_1b_1c", "gfa_x_fasta_3x10k"):
            gold = load(name)
            args = rem.rem_args(case_files(gold, tmp_path), **gold["args"])
            G, idx = rem.align_genomes(args, index_module=R.module(32))
            T = idx.T
            if len(G.graph["paths"]) > 2:
                rem.prune_nodes(G, T=T)
            mine = rem.write_gfa(G, T, outputfile=str(tmp_path / (name + "_mine.gfa")))
            # the reference looks nodes up by its own Interval type: same (begin, end), so relabel onto it
            import networkx as nx
            H = nx.relabel_nodes(G, {n: refutils.Interval(n.begin, n.end) for n in G if not isinstance(n, str)}, copy=True)
            refutils.seq2node(H, T, remap=False)
            theirs = str(tmp_path / (name + "_ref.gfa"))
            refutils.write_gfa(H, T, outputfile=theirs)
            a = [x for x in open(mine).read().splitlines() if not x.startswith("H")]
            b = [x for x in open(theirs).read().splitlines() if not x.startswith("H")]
            assert a == b, name
    finally:
        sys.path.remove(str(rendered))
        for mod in ("rem", "schemes", "utils", "reveallib", "reveallib64", "rv_intervaltree"):
            sys.modules.pop(mod, None)


def test_rem_gzipped_input_and_output(tmp_path):
    """FASTA.gz in, GFA.gz out, GFA.gz back in: same graph as with plain files (reference extension as the index)."""
    import oracle.ref as R
    if not R.available():
        pytest.skip("oracle/_ref not built")
    gold = load("synth3_3k")
    plain = case_files(gold, tmp_path)
    zipped = []
    for fn in plain:
        zipped.append(fn + ".gz")
        with gzip.open(zipped[-1], "wt") as f:
            f.write(open(fn).read())
    results = []
    for files in (plain, zipped):
        G, idx = rem.align_genomes(rem.rem_args(files, **gold["args"]), index_module=R.module(32))
        rem.prune_nodes(G, T=idx.T)
        results.append((G, idx))
    a, b = (M.canonical(G, idx.T) for G, idx in results)
    a["walks"] = sorted(a["walks"].values())
    b["walks"] = sorted(b["walks"].values())
    assert a == b
    G, idx = results[1]
    out = rem.write_gfa(G, idx.T, outputfile=str(tmp_path / "graph"))          # no extension: .gfa.gz is added
    assert out.endswith(".gfa.gz") and gzip.open(out, "rt").readline().startswith("H\tVN:Z:1.0")
    # the written graph as input again, against one more sequence
    extra = str(tmp_path / "extra.fa.gz")
    with gzip.open(extra, "wt") as f:
        f.write(">extra\n%s\n" % synth.genomes(1, 3000, seed=22)[0].tobytes().decode())
    G2, idx2 = rem.align_genomes(rem.rem_args([out, extra], minlength=10), index_module=R.module(32))
    assert len(G2.graph["paths"]) == 4 and rem.aligned_bases(G2, idx2)[0] > 0


# ---- sharded recursion: ONE alignment over the ranks of a torch.distributed job ----------------------------
def _shard_worker(rank, world, port, q, names, emu_path, tmp):
    """One rank: the same inputs, the same index, its own share of the recursion's units; rank 0 checks the collected graph
    against the golden one (the very graph the unsharded reference driver produced)."""
    import pathlib
    import traceback

    import torch.distributed as dist
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from reveal_b200 import reveallib
        if emu_path:
            reveallib._load(emu_path)   # the emulated kernels stand in for the GPUs (REVEAL_B200_TEST_HOOKS=1 from conftest)
        out = []
        for name in names:
            d = pathlib.Path(tmp) / ("%s_r%d" % (name, rank))
            d.mkdir()
            G, idx = run_case(name, d, reveallib, shard=(rank, world))
            units = idx.shard_units
            out.append((name, len(units), sorted({o for o, _ in units}), rem.align_genomes.last_shard_stats["own_nodes"]))
        dist.barrier()
        q.put((rank, "ok", out))
        dist.destroy_process_group()
    except Exception:
        q.put((rank, "error", traceback.format_exc()))


def _run_sharded(names, emu_path, tmp_path, world=2):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 400
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, q, names, emu_path, str(tmp_path))) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
    for rank, status, payload in res:
        assert status == "ok", "rank %d:\n%s" % (rank, payload)
    return res


def test_rem_sharded_recursion_world2_gloo(emu_lib, tmp_path):
    """Two ranks (gloo, emulated kernels) share ONE alignment: every rank runs the tree above the cut, the units below it are dealt
    out, rank 0 collects the refined parts -- and holds exactly the golden graph of the unsharded reference driver."""
    emu_path = os.path.join(HERE, "emu", "_build", "libreveal_emu.so")
    res = _run_sharded(["synth2_4k", "synth3_3k", "synth4_2k_seed"], emu_path, tmp_path)
    for rank, _, cases in res:
        for name, nunits, owners, own_nodes in cases:
            assert nunits >= 2 and owners == [0, 1], (name, nunits, owners)   # the cut produced units for both ranks
            assert own_nodes > 0


@pytest.mark.gpu
def test_rem_sharded_recursion_two_ranks_one_gpu(tmp_path):
    """The same on the CUDA library: two processes share GPU 0 (the test box has one GPU); graphs of up to 5 x 30 kbp."""
    res = _run_sharded(["synth2_200k", "synth3_60k", "synth5_30k_n3", "1a_1b"], None, tmp_path)
    for rank, _, cases in res:
        for name, nunits, owners, own_nodes in cases:
            assert owners == [0, 1], (name, nunits, owners)


def test_rem_batched_picks_with_device_chaining_emulated(emu_reveallib, tmp_path, monkeypatch):
    """Frontier batches hand all their MUM lists to remcore.Graph.mumpicker_batch; RV_REM_CHAIN=device sends every chaining recurrence
    of a batch through rv_chain_batch (here: the emulated kernel) -- the graphs stay the golden ones."""
    from reveal_b200 import remcore
    monkeypatch.setenv("RV_REM_CHAIN", "device")
    remcore._set_chain_library(os.path.join(HERE, "emu", "_build", "libreveal_emu.so"))
    try:
        before = remcore.chain_stats()
        for name in ("synth2_4k", "synth3_3k", "synth4_2k_seed"):
            run_case(name, tmp_path / name if (tmp_path / name).mkdir() is None else tmp_path, emu_reveallib.mod32)
        after = remcore.chain_stats()
        assert after["device"] and after["device_lists"] > before["device_lists"] and after["launches"] > before["launches"]
    finally:
        remcore._set_chain_library("")   # back to the library next to the module


def test_rem_pick_pool_emulated(emu_reveallib, tmp_path):
    """The C++ half of the picks of a frontier batch on remcore's thread pool (set_threads): the golden graph with three threads
    (the other emulated cases pin it with the default), the pool really used, and a forked child (which has none of the parent's
    threads) making its own."""
    from reveal_b200 import remcore
    prev = remcore.set_threads(3)
    try:
        before = remcore.chain_stats()
        d = tmp_path / "t3"
        d.mkdir()
        run_case("synth2_4k", d, emu_reveallib.mod32)
        after = remcore.chain_stats()
        assert after["threads"] == 3 and after["pooled_calls"] > before["pooled_calls"]
        remcore.set_threads(1)
        before = remcore.chain_stats()
        d = tmp_path / "t1"
        d.mkdir()
        run_case("t1_t2", d, emu_reveallib.mod32)
        assert remcore.chain_stats()["pooled_calls"] == before["pooled_calls"]
        pid = os.fork()
        if pid == 0:   # the child: same module state, no worker threads
            code = 1
            try:
                remcore.set_threads(2)
                d = tmp_path / "child"
                d.mkdir()
                run_case("synth2_4k", d, emu_reveallib.mod32)
                code = 0 if remcore.chain_stats()["pooled_calls"] > after["pooled_calls"] else 2
            finally:
                os._exit(code)
        assert os.waitpid(pid, 0)[1] == 0
    finally:
        remcore.set_threads(prev)
