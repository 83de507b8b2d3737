// remcore -- the alignment graph of the REM driver while the recursion runs, in C++ (host code).
//
// reveal_b200/rem.py keeps the graph as a networkx (Multi)DiGraph, like the reference (reveal/rem.py).  One recursion step
// cuts the matched piece out of one node per sample, folds the pieces into one node and walks the neighbourhood to sort the
// intervals of the sub-index into "before", "after" and "beside" the new node (graphalign -> breaknode, mergenodes,
// segmentgraph; reference: reveal/rem.py:14-382).  On networkx dictionaries that is ~0.2 ms of interpreter time per aligned
// MUM and the largest share of an end-to-end `rem`.  This module holds the same graph in flat vectors for the duration of the
// recursion: rem.py loads it from the networkx graph after the input is read (Graph.add_node / add_edge), calls
// Graph.graphalign from the callback, asks it for path coordinates (coords / node_offsets) in the mumpicker, and rebuilds the
// networkx graph from Graph.export() afterwards.  Same semantics as the Python methods of rem.Rem, which stay as the readable
// twin; tests/test_rem.py runs the golden graphs through both.
//
// Nodes are intervals [begin, end) of the index text (unique begin) or marker nodes (start / end of paths, any hashable
// Python object).  Edges carry ofrom / oto ('+' / '-'), a set of path ids and, rarely, extra GFA tags (kept as a dict).
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdlib.h>
#include <string>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <pthread.h>
#include <deque>
#include <map>
#include <unordered_map>
#include <vector>

#include "../../../include/reveal_b200.h"
#include "chain_dp.h"

namespace {

struct Edge {
    int u, v;
    char ofrom, oto;
    std::vector<int32_t> paths;  // sorted, unique
    PyObject *extra;             // dict of further attributes (cigar ...) or nullptr
    bool alive;
};

struct Node {
    int64_t begin, end;          // interval nodes
    PyObject *key;               // marker nodes: the Python object that names them (owned); nullptr for intervals
    int aligned;                 // -1: attribute absent (markers)
    PyObject *extra;             // dict of further node attributes (endpoint, seq ...) or nullptr
    std::vector<std::pair<int32_t, int64_t>> offsets;  // insertion-ordered, like the dict it mirrors
    std::vector<int> out, in;    // edge ids, insertion order (dead ones are skipped and compacted lazily)
    bool alive;
};

static void unite(std::vector<int32_t> &into, const std::vector<int32_t> &from) {
    std::vector<int32_t> r;
    r.reserve(into.size() + from.size());
    std::set_union(into.begin(), into.end(), from.begin(), from.end(), std::back_inserter(r));
    into.swap(r);
}

struct Graph {
    PyObject_HEAD
    bool multi;
    PyObject *interval_cls;
    std::vector<Node> *nodes;
    std::vector<Edge> *edges;
    std::unordered_map<int64_t, int> *by_begin;   // every alive interval node
    std::map<int64_t, int> *tracked;              // unaligned interval nodes only: position -> node lookups
    std::vector<char> *real;                      // per path id: not a '*' path
    std::vector<int64_t> *id2end;                 // per path id: length of the path
    PyObject *markers;                            // dict: marker object -> node id
    std::vector<uint32_t> *seen, *mark;           // epoch-stamped visit marks of the walks (no clearing per walk)
    uint32_t epoch;
    // options of mumpicker(): the default flow of Rem.graphmumpicker as a callable of this object
    int pk_trim, pk_model;
    long pk_maxmums, pk_maxdepth;   // maxdepth < 0: none
    long long pk_wscore, pk_wpen, pk_seedsize;
};

static int new_node(Graph *g) {
    g->nodes->emplace_back();
    Node &n = g->nodes->back();
    n.begin = n.end = 0;
    n.key = nullptr;
    n.extra = nullptr;
    n.aligned = -1;
    n.alive = true;
    return (int)g->nodes->size() - 1;
}

static int add_interval(Graph *g, int64_t b, int64_t e, int aligned, bool track) {
    int id = new_node(g);
    Node &n = (*g->nodes)[id];
    n.begin = b;
    n.end = e;
    n.aligned = aligned;
    (*g->by_begin)[b] = id;
    if (track) (*g->tracked)[b] = id;
    return id;
}

static bool is_real(const Graph *g, int32_t sid) { return sid < 0 || (size_t)sid >= g->real->size() || (*g->real)[sid]; }

// networkx semantics: a MultiDiGraph always gets a new parallel edge; a DiGraph updates the attributes of an existing one
static int add_edge(Graph *g, int u, int v, char ofrom, char oto, const std::vector<int32_t> &paths, PyObject *extra) {
    if (!g->multi) {
        for (int eid : (*g->nodes)[u].out) {
            Edge &e = (*g->edges)[eid];
            if (e.alive && e.v == v) {
                e.ofrom = ofrom;
                e.oto = oto;
                e.paths = paths;
                if (extra) {
                    Py_INCREF(extra);
                    Py_XSETREF(e.extra, extra);
                }
                return eid;
            }
        }
    }
    Edge e;
    e.u = u;
    e.v = v;
    e.ofrom = ofrom;
    e.oto = oto;
    e.paths = paths;
    e.extra = extra;
    Py_XINCREF(extra);
    e.alive = true;
    g->edges->push_back(e);
    int eid = (int)g->edges->size() - 1;
    (*g->nodes)[u].out.push_back(eid);
    (*g->nodes)[v].in.push_back(eid);
    return eid;
}

static void compact(Graph *g, std::vector<int> &list) {
    size_t w = 0;
    for (size_t r = 0; r < list.size(); r++)
        if ((*g->edges)[list[r]].alive) list[w++] = list[r];
    list.resize(w);
}

static void remove_node(Graph *g, int id) {
    Node &n = (*g->nodes)[id];
    for (int eid : n.out) {
        Edge &e = (*g->edges)[eid];
        if (!e.alive) continue;
        e.alive = false;
        Py_CLEAR(e.extra);
        if (e.v != id) compact(g, (*g->nodes)[e.v].in);
    }
    for (int eid : n.in) {
        Edge &e = (*g->edges)[eid];
        if (!e.alive) continue;
        e.alive = false;
        Py_CLEAR(e.extra);
        if (e.u != id) compact(g, (*g->nodes)[e.u].out);
    }
    n.out.clear();
    n.in.clear();
    if (!n.key) {
        auto it = g->by_begin->find(n.begin);
        if (it != g->by_begin->end() && it->second == id) g->by_begin->erase(it);
        auto jt = g->tracked->find(n.begin);
        if (jt != g->tracked->end() && jt->second == id) g->tracked->erase(jt);
    }
    n.alive = false;
    n.offsets.clear();
    Py_CLEAR(n.extra);
}

static int node_at(const Graph *g, int64_t pos) {
    auto it = g->tracked->upper_bound(pos);
    if (it == g->tracked->begin()) return -1;
    --it;
    const Node &n = (*g->nodes)[it->second];
    return pos < n.end ? it->second : -1;
}

struct EdgeCopy {
    int other;
    char ofrom, oto;
    std::vector<int32_t> paths;
    PyObject *extra;  // borrowed while the original edge is alive; INCREF'd by add_edge when re-attached
};

// rem.Rem.breaknode (reference: rem.py:14-131): cut [pos, pos+l) out of node `id`; returns the matching piece, the left-over
// pieces go to `others`
static int breaknode(Graph *g, int id, int64_t pos, int64_t l, std::vector<int> &others) {
    std::vector<Node> &N = *g->nodes;
    if (N[id].begin == pos && N[id].end == pos + l) {
        g->tracked->erase(N[id].begin);
        return id;
    }
    const int64_t nb = N[id].begin, ne = N[id].end;
    std::vector<EdgeCopy> ins, outs;
    std::vector<PyObject *> held;
    for (int eid : N[id].in) {
        Edge &e = (*g->edges)[eid];
        if (e.alive) { ins.push_back(EdgeCopy{e.u, e.ofrom, e.oto, e.paths, e.extra}); if (e.extra) { Py_INCREF(e.extra); held.push_back(e.extra); } }
    }
    for (int eid : N[id].out) {
        Edge &e = (*g->edges)[eid];
        if (e.alive) { outs.push_back(EdgeCopy{e.v, e.ofrom, e.oto, e.paths, e.extra}); if (e.extra) { Py_INCREF(e.extra); held.push_back(e.extra); } }
    }
    std::vector<int32_t> pos_paths, neg_paths;
    if (ins.empty() && outs.empty())
        for (auto &kv : N[id].offsets) pos_paths.push_back(kv.first);
    for (auto &c : ins)
        for (int32_t p : c.paths) (c.oto == '-' ? neg_paths : pos_paths).push_back(p);
    for (auto &c : outs)
        for (int32_t p : c.paths) (c.ofrom == '-' ? neg_paths : pos_paths).push_back(p);
    for (auto *v : {&pos_paths, &neg_paths}) {
        std::sort(v->begin(), v->end());
        v->erase(std::unique(v->begin(), v->end()), v->end());
    }
    const std::vector<std::pair<int32_t, int64_t>> base = N[id].offsets;
    const int64_t shift = pos - nb;
    g->tracked->erase(nb);
    g->by_begin->erase(nb);  // the head piece (if any) takes over this begin
    int piece = add_interval(g, pos, pos + l, 0, false);
    for (auto &kv : base) (*g->nodes)[piece].offsets.emplace_back(kv.first, kv.second + shift);
    int head = piece, tail = piece;
    if (nb != pos) {
        head = add_interval(g, nb, pos, 0, true);
        (*g->nodes)[head].offsets = base;
        add_edge(g, head, piece, '+', '+', pos_paths, nullptr);
        if (!neg_paths.empty()) add_edge(g, piece, head, '-', '-', neg_paths, nullptr);
        others.push_back(head);
    }
    if (ne != pos + l) {
        tail = add_interval(g, pos + l, ne, 0, true);
        for (auto &kv : base) (*g->nodes)[tail].offsets.emplace_back(kv.first, kv.second + shift + l);
        add_edge(g, piece, tail, '+', '+', pos_paths, nullptr);
        if (!neg_paths.empty()) add_edge(g, tail, piece, '-', '-', neg_paths, nullptr);
        others.push_back(tail);
    }
    remove_node(g, id);  // its begin may belong to the head piece by now: remove_node only erases map entries that still name `id`
    for (auto &c : ins) add_edge(g, c.other, c.oto == '+' ? head : tail, c.ofrom, c.oto, c.paths, c.extra);
    for (auto &c : outs) add_edge(g, c.ofrom == '+' ? tail : head, c.other, c.ofrom, c.oto, c.paths, c.extra);
    for (PyObject *o : held) Py_DECREF(o);
    return piece;
}

// rem.Rem.mergenodes (reference: rem.py:133-205)
static int mergenodes(Graph *g, const std::vector<int> &group) {
    std::vector<Node> &N = *g->nodes;
    const int keep = group[0];
    std::vector<std::pair<int32_t, int64_t>> offsets;
    for (int id : group)
        for (auto &kv : N[id].offsets) {
            bool found = false;
            for (auto &have : offsets)
                if (have.first == kv.first) { have.second = kv.second; found = true; break; }
            if (!found) offsets.push_back(kv);
        }
    N[keep].offsets = offsets;
    N[keep].aligned = 1;
    for (size_t k = 1; k < group.size(); k++) {
        const int id = group[k];
        std::vector<int> ins, outs;
        for (int eid : N[id].in) if ((*g->edges)[eid].alive) ins.push_back(eid);
        for (int eid : N[id].out) if ((*g->edges)[eid].alive) outs.push_back(eid);
        for (int eid : ins) {
            const Edge e = (*g->edges)[eid];
            bool merged = false;
            for (int kid : (*g->nodes)[keep].in) {
                Edge &k2 = (*g->edges)[kid];
                if (!k2.alive || k2.u != e.u) continue;
                if (!g->multi || (k2.oto == e.oto && k2.ofrom == e.ofrom)) { unite(k2.paths, e.paths); merged = true; break; }
            }
            if (!merged) add_edge(g, e.u, keep, e.ofrom, e.oto, e.paths, e.extra);
        }
        for (int eid : outs) {
            const Edge e = (*g->edges)[eid];
            bool merged = false;
            for (int kid : (*g->nodes)[keep].out) {
                Edge &k2 = (*g->edges)[kid];
                if (!k2.alive || k2.v != e.v) continue;
                if (!g->multi || (k2.oto == e.oto && k2.ofrom == e.ofrom)) { unite(k2.paths, e.paths); merged = true; break; }
            }
            if (!merged) add_edge(g, keep, e.v, e.ofrom, e.oto, e.paths, e.extra);
        }
        remove_node(g, id);
    }
    return keep;
}

static bool edge_counts(const Graph *g, const Edge &e) {
    for (int32_t p : e.paths)
        if (is_real(g, p)) return true;
    return false;
}

// rem.Rem._reach (reference bfs, rem.py:235-262): class 0 = unaligned (walked through), 1 = aligned (stop), 2 = marker (stop)
static void reach(Graph *g, int source, bool backwards, const std::vector<int> *through, std::vector<std::pair<int, int>> &out) {
    std::vector<Node> &N = *g->nodes;
    std::vector<uint32_t> &seen = *g->seen;
    if (seen.size() < N.size()) seen.resize(N.size() + N.size() / 2 + 16, 0);
    const uint32_t tick = ++g->epoch;
    std::deque<int> queue;
    seen[source] = tick;
    queue.push_back(source);
    while (!queue.empty()) {
        int cur = queue.front();
        queue.pop_front();
        const std::vector<int> &adj = backwards ? N[cur].in : N[cur].out;
        for (int eid : adj) {
            const Edge &e = (*g->edges)[eid];
            if (!e.alive || !edge_counts(g, e)) continue;
            int child = backwards ? e.u : e.v;
            if (seen[child] == tick) continue;
            seen[child] = tick;
            const Node &c = N[child];
            if (c.aligned < 0) {
                out.emplace_back(child, 2);
            } else if (c.aligned == 0 || (through && std::find(through->begin(), through->end(), child) != through->end())) {
                queue.push_back(child);
                out.emplace_back(child, 0);
            } else {
                out.emplace_back(child, 1);
            }
        }
    }
}

static PyObject *interval_tuple(int64_t b, int64_t e) { return Py_BuildValue("(LL)", (long long)b, (long long)e); }

// one side of rem.Rem.segmentgraph: the unaligned interval nodes strictly behind (or before) `node`, as a Python set of
// (begin, end) tuples restricted to `members`
static PyObject *side_of(Graph *g, int node, bool backwards, PyObject *members) {
    std::vector<std::pair<int, int>> found;
    reach(g, node, backwards, nullptr, found);
    std::vector<int> side, stops;
    for (auto &f : found) (f.second == 0 ? side : stops).push_back(f.first);
    if (stops.size() > 1) {
        std::vector<uint32_t> &back = *g->mark;
        if (back.size() < g->nodes->size()) back.resize(g->nodes->size() + g->nodes->size() / 2 + 16, 0);
        const uint32_t tag = ++g->epoch;
        for (int stop : stops) {
            std::vector<std::pair<int, int>> r;
            reach(g, stop, !backwards, &stops, r);
            for (auto &f : r)
                if (f.second == 0) back[f.first] = tag;
        }
        std::vector<int> kept;
        for (int s : side)
            if (back[s] == tag) kept.push_back(s);
        side.swap(kept);
    }
    PyObject *res = PySet_New(nullptr);
    if (!res) return nullptr;
    for (int s : side) {
        const Node &n = (*g->nodes)[s];
        if (n.key) continue;
        PyObject *t = interval_tuple(n.begin, n.end);
        if (!t) { Py_DECREF(res); return nullptr; }
        int has = PySet_Contains(members, t);
        if (has < 0 || (has && PySet_Add(res, t) < 0)) { Py_DECREF(t); Py_DECREF(res); return nullptr; }
        Py_DECREF(t);
    }
    return res;
}

static int find_node(Graph *g, PyObject *key) {
    if (PyTuple_Check(key) && PyTuple_GET_SIZE(key) >= 2 && PyLong_Check(PyTuple_GET_ITEM(key, 0))) {
        long long b = PyLong_AsLongLong(PyTuple_GET_ITEM(key, 0));
        auto it = g->by_begin->find(b);
        if (it == g->by_begin->end()) { PyErr_Format(PyExc_KeyError, "no interval node starts at %lld", b); return -1; }
        return it->second;
    }
    PyObject *v = PyDict_GetItemWithError(g->markers, key);
    if (!v) { if (!PyErr_Occurred()) PyErr_SetObject(PyExc_KeyError, key); return -1; }
    return (int)PyLong_AsLong(v);
}

static bool subset_of(const std::vector<std::pair<int32_t, int64_t>> &offs, const std::vector<std::pair<int32_t, int64_t>> &msamples) {
    for (auto &kv : offs) {
        bool in = false;
        for (auto &m : msamples)
            if (m.first == kv.first) { in = true; break; }
        if (!in) return false;
    }
    return true;
}

// ---- Python methods -----------------------------------------------------------------------------------------------------------
static PyObject *Graph_new(PyTypeObject *type, PyObject *, PyObject *) {
    Graph *g = (Graph *)type->tp_alloc(type, 0);
    if (!g) return nullptr;
    g->multi = true;
    g->interval_cls = nullptr;
    g->nodes = new std::vector<Node>();
    g->edges = new std::vector<Edge>();
    g->by_begin = new std::unordered_map<int64_t, int>();
    g->tracked = new std::map<int64_t, int>();
    g->real = new std::vector<char>();
    g->id2end = new std::vector<int64_t>();
    g->seen = new std::vector<uint32_t>();
    g->mark = new std::vector<uint32_t>();
    g->epoch = 0;
    g->markers = PyDict_New();
    return (PyObject *)g;
}

static int Graph_init(Graph *g, PyObject *args, PyObject *) {
    int multi = 1;
    PyObject *cls = nullptr, *real = nullptr, *ends = nullptr;
    if (!PyArg_ParseTuple(args, "pOOO", &multi, &cls, &real, &ends)) return -1;
    {
        PyObject *it = PyObject_GetIter(ends);
        if (!it) return -1;
        g->id2end->clear();
        while (PyObject *x = PyIter_Next(it)) {
            g->id2end->push_back((int64_t)PyLong_AsLongLong(x));
            Py_DECREF(x);
        }
        Py_DECREF(it);
        if (PyErr_Occurred()) return -1;
    }
    g->multi = multi != 0;
    Py_INCREF(cls);
    Py_XSETREF(g->interval_cls, cls);
    PyObject *it = PyObject_GetIter(real);
    if (!it) return -1;
    g->real->clear();
    while (PyObject *x = PyIter_Next(it)) {
        g->real->push_back(PyObject_IsTrue(x) ? 1 : 0);
        Py_DECREF(x);
    }
    Py_DECREF(it);
    return PyErr_Occurred() ? -1 : 0;
}

static void Graph_dealloc(Graph *g) {
    if (g->nodes)
        for (Node &n : *g->nodes) { Py_XDECREF(n.key); Py_XDECREF(n.extra); }
    if (g->edges)
        for (Edge &e : *g->edges) Py_XDECREF(e.extra);
    delete g->nodes;
    delete g->edges;
    delete g->by_begin;
    delete g->tracked;
    delete g->real;
    delete g->id2end;
    delete g->seen;
    delete g->mark;
    Py_XDECREF(g->markers);
    Py_XDECREF(g->interval_cls);
    Py_TYPE(g)->tp_free((PyObject *)g);
}

static bool read_offsets(PyObject *d, std::vector<std::pair<int32_t, int64_t>> &out) {
    PyObject *k, *v;
    Py_ssize_t p = 0;
    while (PyDict_Next(d, &p, &k, &v)) out.emplace_back((int32_t)PyLong_AsLong(k), (int64_t)PyLong_AsLongLong(v));
    return !PyErr_Occurred();
}

static bool read_paths(PyObject *s, std::vector<int32_t> &out) {
    PyObject *it = PyObject_GetIter(s);
    if (!it) return false;
    while (PyObject *x = PyIter_Next(it)) {
        out.push_back((int32_t)PyLong_AsLong(x));
        Py_DECREF(x);
    }
    Py_DECREF(it);
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return !PyErr_Occurred();
}

// add_node(key, aligned or None, offsets dict, extra dict or None)
static PyObject *Graph_add_node(Graph *g, PyObject *args) {
    PyObject *key, *aligned, *offsets, *extra;
    if (!PyArg_ParseTuple(args, "OOOO", &key, &aligned, &offsets, &extra)) return nullptr;
    int al = aligned == Py_None ? -1 : (int)PyLong_AsLong(aligned);
    int id;
    if (PyTuple_Check(key) && PyTuple_GET_SIZE(key) >= 2 && PyLong_Check(PyTuple_GET_ITEM(key, 0))) {
        long long b = PyLong_AsLongLong(PyTuple_GET_ITEM(key, 0)), e = PyLong_AsLongLong(PyTuple_GET_ITEM(key, 1));
        id = add_interval(g, b, e, al, al == 0);
    } else {
        id = new_node(g);
        Node &n = (*g->nodes)[id];
        n.key = key;
        Py_INCREF(key);
        n.aligned = al;
        PyObject *idobj = PyLong_FromLong(id);
        PyDict_SetItem(g->markers, key, idobj);
        Py_DECREF(idobj);
    }
    Node &n = (*g->nodes)[id];
    if (offsets != Py_None && !read_offsets(offsets, n.offsets)) return nullptr;
    if (extra != Py_None && PyDict_Size(extra) > 0) { n.extra = extra; Py_INCREF(extra); }
    if (PyErr_Occurred()) return nullptr;
    Py_RETURN_NONE;
}

// add_edge(u, v, ofrom, oto, paths, extra dict or None)
static PyObject *Graph_add_edge(Graph *g, PyObject *args) {
    PyObject *u, *v, *paths, *extra;
    const char *ofrom, *oto;
    if (!PyArg_ParseTuple(args, "OOssOO", &u, &v, &ofrom, &oto, &paths, &extra)) return nullptr;
    int iu = find_node(g, u);
    if (iu < 0) return nullptr;
    int iv = find_node(g, v);
    if (iv < 0) return nullptr;
    std::vector<int32_t> p;
    if (!read_paths(paths, p)) return nullptr;
    add_edge(g, iu, iv, ofrom[0], oto[0], p, (extra != Py_None && PyDict_Size(extra) > 0) ? extra : nullptr);
    Py_RETURN_NONE;
}

static PyObject *offsets_dict(const std::vector<std::pair<int32_t, int64_t>> &offs) {
    PyObject *d = PyDict_New();
    if (!d) return nullptr;
    for (auto &kv : offs) {
        PyObject *k = PyLong_FromLong(kv.first), *v = PyLong_FromLongLong(kv.second);
        PyDict_SetItem(d, k, v);
        Py_DECREF(k);
        Py_DECREF(v);
    }
    return d;
}

// coords(pos) -> ((path id, coordinate), ...) over the real paths of the unaligned node that covers index position pos
static PyObject *Graph_coords(Graph *g, PyObject *arg) {
    long long pos = PyLong_AsLongLong(arg);
    if (pos == -1 && PyErr_Occurred()) return nullptr;
    int id = node_at(g, pos);
    if (id < 0) { PyErr_Format(PyExc_KeyError, "no node covers index position %lld", pos); return nullptr; }
    const Node &n = (*g->nodes)[id];
    const int64_t rel = pos - n.begin;
    Py_ssize_t cnt = 0;
    for (auto &kv : n.offsets) cnt += is_real(g, kv.first);
    PyObject *t = PyTuple_New(cnt);
    Py_ssize_t w = 0;
    for (auto &kv : n.offsets)
        if (is_real(g, kv.first)) PyTuple_SET_ITEM(t, w++, Py_BuildValue("(iL)", (int)kv.first, (long long)(kv.second + rel)));
    return t;
}

// node_offsets(node) -> {path id: offset}
static PyObject *Graph_node_offsets(Graph *g, PyObject *key) {
    int id = find_node(g, key);
    if (id < 0) return nullptr;
    return offsets_dict((*g->nodes)[id].offsets);
}

// graphalign(nodes, leftnode, rightnode, l, positions) -> (leading, trailing, matching, rest, merged, newleft, newright)
// rem.Rem.graphalign (reference: rem.py:317-382); `nodes` (the set of the sub-index) is updated in place
static PyObject *Graph_graphalign(Graph *g, PyObject *args) {
    PyObject *nodes, *leftnode, *rightnode, *positions;
    long long l;
    if (!PyArg_ParseTuple(args, "OOOLO", &nodes, &leftnode, &rightnode, &l, &positions)) return nullptr;
    if (!PySet_Check(nodes)) { PyErr_SetString(PyExc_TypeError, "nodes must be a set"); return nullptr; }
    PyObject *seq = PySequence_Fast(positions, "positions must be a sequence");
    if (!seq) return nullptr;
    std::vector<int> pieces;
    PyObject *matching = PySet_New(nullptr);
    const Py_ssize_t np = PySequence_Fast_GET_SIZE(seq);
    for (Py_ssize_t k = 0; k < np; k++) {
        long long pos = PyLong_AsLongLong(PySequence_Fast_GET_ITEM(seq, k));
        if (pos == -1 && PyErr_Occurred()) { Py_DECREF(seq); Py_DECREF(matching); return nullptr; }
        PyObject *m = interval_tuple(pos, pos + l);
        PySet_Add(matching, m);
        Py_DECREF(m);
        int old = node_at(g, pos);
        if (old < 0 || (*g->nodes)[old].end - pos < l) {
            PyErr_Format(PyExc_KeyError, "no unaligned node holds [%lld, %lld)", pos, pos + l);
            Py_DECREF(seq); Py_DECREF(matching);
            return nullptr;
        }
        PyObject *oldt = interval_tuple((*g->nodes)[old].begin, (*g->nodes)[old].end);
        std::vector<int> others;
        int piece = breaknode(g, old, pos, l, others);
        pieces.push_back(piece);
        if (PySet_Discard(nodes, oldt) < 0) { Py_DECREF(oldt); Py_DECREF(seq); Py_DECREF(matching); return nullptr; }
        Py_DECREF(oldt);
        for (int o : others) {
            PyObject *t = interval_tuple((*g->nodes)[o].begin, (*g->nodes)[o].end);
            PySet_Add(nodes, t);
            Py_DECREF(t);
        }
    }
    Py_DECREF(seq);
    if (pieces.empty()) { Py_DECREF(matching); PyErr_SetString(PyExc_ValueError, "a match needs at least one position"); return nullptr; }
    int merged = mergenodes(g, pieces);
    const std::vector<std::pair<int32_t, int64_t>> msamples = (*g->nodes)[merged].offsets;
    PyObject *members = PySet_New(nodes);
    PyObject *trailing = members ? side_of(g, merged, false, members) : nullptr;
    PyObject *leading = trailing ? side_of(g, merged, true, members) : nullptr;
    PyObject *rest = nullptr, *merged_obj = nullptr, *ret = nullptr;
    if (leading) {
        rest = PySet_New(members);
        PyObject *it;
        for (PyObject *side : {leading, trailing}) {
            it = PyObject_GetIter(side);
            while (PyObject *x = PyIter_Next(it)) { PySet_Discard(rest, x); Py_DECREF(x); }
            Py_DECREF(it);
        }
        const Node &mn = (*g->nodes)[merged];
        merged_obj = PyObject_CallFunction(g->interval_cls, "LL", (long long)mn.begin, (long long)mn.end);
    }
    if (merged_obj) {
        PyObject *newleft = merged_obj, *newright = merged_obj;
        // a side with intervals of samples outside the match is not cleanly cut by it: keep the old bound
        for (int which = 0; which < 2; which++) {
            PyObject *side = which == 0 ? leading : trailing;
            PyObject *it = PyObject_GetIter(side);
            while (PyObject *x = PyIter_Next(it)) {
                int id = find_node(g, x);
                Py_DECREF(x);
                if (id < 0) { PyErr_Clear(); continue; }
                if (!subset_of((*g->nodes)[id].offsets, msamples)) {
                    if (which == 0) newright = rightnode; else newleft = leftnode;
                    break;
                }
            }
            Py_DECREF(it);
        }
        ret = PyTuple_Pack(7, leading, trailing, matching, rest, merged_obj, newleft, newright);
    }
    Py_XDECREF(members); Py_XDECREF(trailing); Py_XDECREF(leading); Py_XDECREF(rest); Py_XDECREF(merged_obj); Py_DECREF(matching);
    return ret;
}

// ---- the mumpicker (rem.Rem.graphmumpicker, default flow; reference: schemes.py:197-361) ---------------------------------------
// A few elements inline, the heap only beyond that: an anchor has one position per sample (two, mostly), and the mumpicker
// sorts, copies and filters lists of tens of thousands of anchors at the top of the recursion -- with std::vector members every
// one of those moves is an allocation.
template <class T, int N> struct SmallVec {
    T inl[N];
    std::vector<T> more;
    uint32_t count = 0;
    SmallVec() {}
    // copies and moves touch the elements in use only (two of the eight inline slots for a pair of genomes)
    SmallVec(const SmallVec &o) : more(o.more), count(o.count) { take(o); }
    SmallVec(SmallVec &&o) noexcept : more(std::move(o.more)), count(o.count) { take(o); }
    SmallVec &operator=(const SmallVec &o) {
        if (this != &o) { more = o.more; count = o.count; take(o); }
        return *this;
    }
    SmallVec &operator=(SmallVec &&o) noexcept {
        if (this != &o) { more = std::move(o.more); count = o.count; take(o); }
        return *this;
    }
    void take(const SmallVec &o) {
        const uint32_t k = count < (uint32_t)N ? count : (uint32_t)N;
        for (uint32_t i = 0; i < k; i++) inl[i] = o.inl[i];
    }
    size_t size() const { return count; }
    bool empty() const { return count == 0; }
    T *data() { return count <= (uint32_t)N ? inl : more.data(); }
    const T *data() const { return count <= (uint32_t)N ? inl : more.data(); }
    T *begin() { return data(); }
    T *end() { return data() + count; }
    const T *begin() const { return data(); }
    const T *end() const { return data() + count; }
    T &operator[](size_t i) { return data()[i]; }
    const T &operator[](size_t i) const { return data()[i]; }
    void push_back(const T &v) {
        if (count < (uint32_t)N) inl[count] = v;
        else {
            if (count == (uint32_t)N) more.assign(inl, inl + N);
            more.push_back(v);
        }
        count++;
    }
    template <class A, class B> void emplace_back(A a, B b) { push_back(T(a, b)); }
};

struct Mum {
    int64_t l;
    long n;
    SmallVec<std::pair<long, int64_t>, 8> sp;  // (sample of the index, position), in the order of the tuple
    PyObject *orig;                            // the caller's tuple while the anchor is untouched (borrowed)
    PyObject *spd;                             // the caller's position tuple while the positions are untouched (borrowed)
};

struct Rel {
    int64_t l;
    long n;
    SmallVec<std::pair<int32_t, int64_t>, 8> point;  // (path id, coordinate), insertion-ordered like the dict it mirrors
    int src;                                         // index into the picked anchors
    uint64_t value_hash() const {                    // of the coordinate values in order (the reference keys a dict by that tuple)
        uint64_t h = 0x9e3779b97f4a7c15ull ^ point.size();
        for (auto &kv : point) h = (h ^ (uint64_t)kv.second) * 0xff51afd7ed558ccdull, h ^= h >> 29;
        return h;
    }
    bool same_values(const Rel &o) const {
        if (point.size() != o.point.size()) return false;
        for (size_t i = 0; i < point.size(); i++)
            if (point[i].second != o.point[i].second) return false;
        return true;
    }
};

// origin: coordinate-value tuple of an anchor -> index of the LAST picked anchor with that tuple (the reference's
// `origin[tuple(r[2].values())] = m`, schemes.py:236-240).  A hash on the values, verified on the stored copy; the rare tuple
// whose hash slot was taken over by another tuple is found by a scan from the end.
struct Origin {
    std::unordered_map<uint64_t, int> slot;
    std::vector<Rel> by_src;   // rel of picked anchor i (before the list is sorted and cut)
    void build(const std::vector<Rel> &rel) {
        by_src = rel;
        slot.reserve(rel.size() * 2);
        for (size_t i = 0; i < rel.size(); i++) slot[rel[i].value_hash()] = (int)i;
    }
    int find(const Rel &r) const {
        auto it = slot.find(r.value_hash());
        if (it != slot.end() && by_src[(size_t)it->second].same_values(r)) return it->second;
        for (size_t i = by_src.size(); i-- > 0;)
            if (by_src[i].same_values(r)) return (int)i;
        return -1;
    }
};

static bool parse_mums(PyObject *list, std::vector<Mum> &out) {
    if (PyObject_CheckBuffer(list) && !PyList_Check(list) && !PyTuple_Check(list)) {
        // pair MUM rows straight from the device sweep (reveallib.mumrows): int64 triples (l, a, b) = (l, 2, ((0, a), (1, b)))
        Py_buffer view;
        if (PyObject_GetBuffer(list, &view, PyBUF_SIMPLE) != 0) return false;
        const int64_t *r = (const int64_t *)view.buf;
        const Py_ssize_t n = view.len / 24;
        out.reserve((size_t)n);
        for (Py_ssize_t i = 0; i < n; i++) {
            Mum m;
            m.l = r[3 * i];
            m.n = 2;
            m.orig = nullptr;
            m.spd = nullptr;
            m.sp.emplace_back(0L, r[3 * i + 1]);
            m.sp.emplace_back(1L, r[3 * i + 2]);
            out.push_back(m);
        }
        PyBuffer_Release(&view);
        return true;
    }
    if (!PyList_Check(list) && !PyTuple_Check(list) && PyObject_HasAttrString(list, "_rv_multi")) {
        // multi-MUM rows straight from the device sweep (reveallib.multimumrows): record rows (l, n, first) + member rows
        struct RvMultiView { const int64_t *hdr; int64_t nrec; const int64_t *mem; int64_t nmem; };   // layout of reveallib_module.cpp
        PyObject *cap = PyObject_GetAttrString(list, "_rv_multi");
        if (!cap) return false;
        const RvMultiView *v = (const RvMultiView *)PyCapsule_GetPointer(cap, "reveal_b200.RvMultiView");
        if (!v) { Py_DECREF(cap); return false; }
        out.reserve((size_t)v->nrec);
        for (int64_t i = 0; i < v->nrec; i++) {
            Mum m;
            m.l = v->hdr[3 * i];
            m.n = (long)v->hdr[3 * i + 1];
            m.orig = nullptr;
            m.spd = nullptr;
            const int64_t first = v->hdr[3 * i + 2];
            for (int64_t x = 0; x < m.n; x++) m.sp.emplace_back((long)v->mem[2 * (first + x)], v->mem[2 * (first + x) + 1]);
            out.push_back(std::move(m));
        }
        Py_DECREF(cap);
        return true;
    }
    PyObject *seq = PySequence_Fast(list, "mums must be a sequence");
    if (!seq) return false;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    out.reserve(n);
    for (Py_ssize_t i = 0; i < n; i++) {
        PyObject *t = PySequence_Fast_GET_ITEM(seq, i);
        if (!PyTuple_Check(t) || PyTuple_GET_SIZE(t) < 3) { PyErr_SetString(PyExc_TypeError, "anchor: (l, n, positions) expected"); Py_DECREF(seq); return false; }
        Mum m;
        m.l = PyLong_AsLongLong(PyTuple_GET_ITEM(t, 0));
        m.n = PyLong_AsLong(PyTuple_GET_ITEM(t, 1));
        m.orig = t;
        m.spd = PyTuple_GET_ITEM(t, 2);
        PyObject *sp = PySequence_Fast(m.spd, "anchor positions must be a sequence");
        if (!sp) { Py_DECREF(seq); return false; }
        for (Py_ssize_t k = 0; k < PySequence_Fast_GET_SIZE(sp); k++) {
            PyObject *pr = PySequence_Fast_GET_ITEM(sp, k);
            if (!PyTuple_Check(pr) || PyTuple_GET_SIZE(pr) < 2) { PyErr_SetString(PyExc_TypeError, "anchor position: (sample, pos) expected"); Py_DECREF(sp); Py_DECREF(seq); return false; }
            m.sp.emplace_back(PyLong_AsLong(PyTuple_GET_ITEM(pr, 0)), (int64_t)PyLong_AsLongLong(PyTuple_GET_ITEM(pr, 1)));
        }
        Py_DECREF(sp);
        out.push_back(std::move(m));
    }
    Py_DECREF(seq);
    return !PyErr_Occurred();
}

static PyObject *mum_object(const Mum &m) {
    if (m.orig) { Py_INCREF(m.orig); return m.orig; }
    PyObject *spd;
    if (m.spd) { spd = m.spd; Py_INCREF(spd); }
    else {
        spd = PyTuple_New((Py_ssize_t)m.sp.size());
        for (size_t k = 0; k < m.sp.size(); k++) PyTuple_SET_ITEM(spd, k, Py_BuildValue("(lL)", m.sp[k].first, (long long)m.sp[k].second));
    }
    PyObject *r = Py_BuildValue("(LlO)", (long long)m.l, m.n, spd);
    Py_DECREF(spd);
    return r;
}

// schemes.trim_overlap (schemes.py:161-191), one coordinate (= member slot) at a time
static void trim_overlap(std::vector<Mum> &mums) {
    if (mums.empty()) return;
    const size_t ncoord = mums[0].sp.size();
    for (size_t coord = 0; coord < ncoord; coord++) {
        if (mums.size() <= 1) break;
        // order by (position in this coordinate, longer first), stable: sorted as small (key, index) records; an anchor itself
        // is moved ONCE per coordinate, into the trimmed list
        const size_t n = mums.size();
        struct Key { int64_t pos, l; uint32_t at; };
        std::vector<Key> keyv(n);
        for (size_t i = 0; i < n; i++) keyv[i] = Key{mums[i].sp[coord].second, mums[i].l, (uint32_t)i};
        bool sorted = true;   // (lists come in suffix-array order or, further down, already in this order: skip the sort when they do)
        for (size_t i = 1; i < n && sorted; i++)
            sorted = keyv[i - 1].pos < keyv[i].pos || (keyv[i - 1].pos == keyv[i].pos && keyv[i - 1].l >= keyv[i].l);
        // (the index as the last key makes the order total: plain sort = stable sort)
        if (!sorted) std::sort(keyv.begin(), keyv.end(), [](const Key &a, const Key &b) { return a.pos != b.pos ? a.pos < b.pos : (a.l != b.l ? a.l > b.l : a.at < b.at); });
        std::vector<uint32_t> kept;
        kept.reserve(n);
        for (size_t i = 0; i < n; i++) {  // the reference compares the FIRST anchor with its successor, or (i-1 = -1) with the last one
            const int64_t end_i = keyv[i].pos + keyv[i].l, end_1 = keyv[1].pos + keyv[1].l, end_0 = keyv[0].pos + keyv[0].l;
            const Key &pk = keyv[(i + n - 1) % n];
            if ((i == 0 && end_1 > end_0) || pk.pos + pk.l < end_i) kept.push_back(keyv[i].at);
        }
        std::vector<Mum> trimmed;
        trimmed.reserve(kept.size());
        if (kept.size() <= 1) {
            for (uint32_t at : kept) trimmed.push_back(std::move(mums[at]));
            mums.swap(trimmed);
            break;
        }
        trimmed.push_back(std::move(mums[kept[0]]));
        for (size_t i = 1; i < kept.size(); i++) {
            Mum &mum = mums[kept[i]];   // (every index occurs once: not moved from yet)
            if (trimmed.empty()) {      // (the anchor before swallowed its predecessor and was dropped itself)
                trimmed.push_back(std::move(mum));
                continue;
            }
            const Mum &prev = trimmed.back();
            const int64_t overlap = prev.sp[coord].second + prev.l - mum.sp[coord].second;
            if (overlap > 0) {
                if (prev.l - overlap > 0) {
                    trimmed.back().l -= overlap;
                    trimmed.back().orig = nullptr;
                } else {
                    trimmed.pop_back();
                }
                if (mum.l - overlap > 0) {
                    mum.l -= overlap;
                    for (auto &p : mum.sp) p.second += overlap;
                    mum.orig = nullptr;
                    mum.spd = nullptr;
                    trimmed.push_back(std::move(mum));
                }
            } else {
                trimmed.push_back(std::move(mum));
            }
        }
        mums.swap(trimmed);
    }
}

// One call of the mumpicker's default flow, cut in two at its chaining recurrence so that the recurrences of MANY calls can run
// together (device: rv_chain_batch, one thread block per list; or the host loop rv_chain_dp):
//   pick_prepare  anchors of all samples, trim_overlap, path coordinates, bounds, the rows of the recurrence
//   pick_finish   back-tracking, split anchor, seeds for the children
struct PickJob {
    long nsamples = 0, maxmums = 0;
    int trim = 0, model = 0;
    long long wscore = 0, wpen = 0, seedsize = 0;
    PyObject *leftnode = nullptr, *rightnode = nullptr;   // borrowed
    std::vector<Mum> all, picked;
    std::vector<Rel> rel;
    Origin origin;
    std::vector<int32_t> keys;
    size_t k = 0, m = 0;
    std::vector<int64_t> left, right;
    std::vector<int> order;
    std::vector<int64_t> start, length, gain, link, score;
    bool need_dp = false;   // start / length / gain hold a recurrence of m + 1 rows; link / score are to be filled
    bool dp_done = false;
    int split = -1;
    // pick_parse -> pick_build: the bounding nodes as node ids (-1: None; -3: not found, the pending KeyError is kept in bound_err)
    int bound_id[2] = {-1, -1};
    PyObject *bound_err[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};
    // pick_build runs without the Python API (possibly on a worker thread): what it has to say waits here
    int err = 0;            // 1 KeyError, 2 ValueError
    int err_bound = -1;     // side whose kept KeyError is to be raised
    std::string err_msg;
    double pk[6] = {0, 0, 0, 0, 0, 0};
    ~PickJob() {
        for (auto &e : bound_err)
            for (PyObject *o : e) Py_XDECREF(o);
    }
};

static double pk_now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static double g_pk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
// The part of the preparation that reads Python objects: the anchors of the list and the node ids of the two bounds.
static int pick_parse(Graph *g, PyObject *list, PickJob &J) {
    const double t0 = pk_now();
    if (!parse_mums(list, J.all)) return -1;
    for (int side = 0; side < 2; side++) {
        PyObject *bound = side == 0 ? J.leftnode : J.rightnode;
        if (bound == Py_None) { J.bound_id[side] = -1; continue; }
        const int id = find_node(g, bound);
        if (id < 0) {  // raised only if the preparation gets as far as the bounds (like the one-piece flow did)
            PyErr_Fetch(&J.bound_err[side][0], &J.bound_err[side][1], &J.bound_err[side][2]);
            J.bound_id[side] = -3;
        } else {
            J.bound_id[side] = id;
        }
    }
    J.pk[0] += pk_now() - t0;
    return 1;
}
static int pick_fail(PickJob &J, int kind, const char *fmt, long long a = 0) {
    char buf[160];
    snprintf(buf, sizeof buf, fmt, a);
    J.err = kind;
    J.err_msg = buf;
    return -1;
}
// The rest of it: no Python API, nothing of the graph is written -- the jobs of a frontier batch run side by side on the
// threads of the pool.  -1: J.err / J.err_bound say what to raise (pick_raise).
static int pick_build(const Graph *g, PickJob &J) {
    double pk_t = pk_now(), pk_u;
#define PK_MARK(i) pk_u = pk_now(); J.pk[i] += pk_u - pk_t; pk_t = pk_u;
    const long nsamples = J.nsamples, maxmums = J.maxmums;
    const int trim = J.trim;
    const long long wscore = J.wscore;
    std::vector<Mum> &all = J.all, &picked = J.picked;
    {
        size_t full = 0;
        for (auto &m : all)
            if (m.n == nsamples) full++;
        if (full == all.size() && full > 0) picked = std::move(all);   // every anchor spans all samples (always, for two): no copy
        else
            for (auto &m : all)
                if (m.n == nsamples) picked.push_back(m);
    }
    if (picked.empty() && nsamples > 2) {  // schemes.segment: the sample group with the largest total length x group size
        std::vector<std::vector<long>> parts;
        std::vector<std::vector<int>> members;
        for (size_t i = 0; i < all.size(); i++) {
            std::vector<long> part;
            for (auto &p : all[i].sp) part.push_back(p.first);
            std::sort(part.begin(), part.end());
            size_t at = 0;
            for (; at < parts.size(); at++)
                if (parts[at] == part) break;
            if (at == parts.size()) { parts.push_back(part); members.emplace_back(); }
            members[at].push_back((int)i);
        }
        int64_t best = 0;
        int pick = -1;
        for (size_t at = 0; at < parts.size(); at++) {
            int64_t z = 0;
            for (int i : members[at]) z += all[i].l;
            z *= (int64_t)parts[at].size();
            if (z > best) { best = z; pick = (int)at; }
        }
        if (pick >= 0)
            for (int i : members[pick]) picked.push_back(all[i]);
    }
    if (picked.empty()) return 0;
    if (trim) {
        trim_overlap(picked);
        if (picked.empty()) return 0;
    }
    PK_MARK(1)
    {   // longest first, stable: sort an index permutation, move the anchors once
        std::vector<uint32_t> perm(picked.size());
        for (size_t i = 0; i < perm.size(); i++) perm[i] = (uint32_t)i;
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return picked[a].l > picked[b].l; });
        std::vector<Mum> sorted;
        sorted.reserve(picked.size());
        for (uint32_t i : perm) sorted.push_back(picked[i]);
        picked.swap(sorted);
    }
    PK_MARK(2)
    // anchors in path coordinates (schemes.lookup / maptooffsets)
    std::vector<Rel> &rel = J.rel;
    rel.assign(picked.size(), Rel());
    for (size_t i = 0; i < picked.size(); i++) {
        Rel &r = rel[i];
        r.l = picked[i].l;
        r.n = 0;
        r.src = (int)i;
        for (auto &p : picked[i].sp) {
            int id = node_at(g, p.second);
            if (id < 0) return pick_fail(J, 1, "no node covers index position %lld", (long long)p.second);
            const Node &nd = (*g->nodes)[id];
            const int64_t shift = p.second - nd.begin;
            for (auto &kv : nd.offsets) {
                if (!is_real(g, kv.first)) continue;
                r.n++;
                bool found = false;
                for (auto &have : r.point)
                    if (have.first == kv.first) { have.second = kv.second + shift; found = true; break; }
                if (!found) r.point.emplace_back(kv.first, kv.second + shift);
            }
        }
    }
    PK_MARK(3)
    J.origin.build(rel);
    PK_MARK(4)
    {
        std::vector<uint32_t> perm(rel.size());
        for (size_t i = 0; i < perm.size(); i++) perm[i] = (uint32_t)i;
        std::stable_sort(perm.begin(), perm.end(), [&](uint32_t a, uint32_t b) { return rel[a].n != rel[b].n ? rel[a].n < rel[b].n : rel[a].l < rel[b].l; });
        std::vector<Rel> sorted;
        sorted.reserve(rel.size());
        for (uint32_t i : perm) sorted.push_back(rel[i]);
        rel.swap(sorted);
    }
    auto same_keyset = [](const Rel &a, const Rel &b) {  // the same set of path ids (a path id occurs once per anchor)
        if (a.point.size() != b.point.size()) return false;
        for (auto &x : a.point) {
            bool found = false;
            for (auto &y : b.point)
                if (y.first == x.first) { found = true; break; }
            if (!found) return false;
        }
        return true;
    };
    {
        const Rel want = rel.back();
        bool all = true;
        for (auto &r : rel)
            if (!same_keyset(r, want)) { all = false; break; }
        if (!all) {
            std::vector<Rel> same;
            for (auto &r : rel)
                if (same_keyset(r, want)) same.push_back(r);
            rel.swap(same);
        }
    }
    std::vector<int32_t> &keys = J.keys;
    for (auto &kv : rel.back().point) keys.push_back(kv.first);
    const size_t k = J.k = keys.size();
    if (k == 0 || k > 64) return pick_fail(J, 2, "anchor over no path or over more than 64 paths");
    // bounds of the sub-index in path coordinates
    std::vector<int64_t> &left = J.left, &right = J.right;
    left.assign(k, 0);
    right.assign(k, 0);
    for (int side = 0; side < 2; side++) {
        const int id = J.bound_id[side];
        if (id == -1) {
            for (size_t c = 0; c < k; c++) {
                if (side == 0) left[c] = -1;
                else {
                    if (keys[c] < 0 || (size_t)keys[c] >= g->id2end->size()) return pick_fail(J, 1, "path without a length");
                    right[c] = (*g->id2end)[keys[c]];
                }
            }
            continue;
        }
        if (id < 0) {
            J.err = 1;
            J.err_bound = side;
            return -1;
        }
        const Node &nd = (*g->nodes)[id];
        for (size_t c = 0; c < k; c++) {
            bool found = false;
            for (auto &kv : nd.offsets)
                if (kv.first == keys[c]) {
                    if (side == 0) left[c] = kv.second + (nd.end - nd.begin) - 1; else right[c] = kv.second;
                    found = true;
                    break;
                }
            if (!found) return pick_fail(J, 1, "path %lld does not run through the bounding node", (long long)keys[c]);
        }
    }
    auto coord_of = [](const Rel &r, int32_t key) -> int64_t {
        for (auto &kv : r.point)
            if (kv.first == key) return kv.second;
        return 0;
    };
    if (rel.size() == 1) {
        J.split = 0;
    } else {
        if (maxmums > 0 && (long)rel.size() > maxmums) rel.erase(rel.begin(), rel.end() - maxmums);
        else if (maxmums <= 0 && !rel.empty()) { /* rel[-0:] is the whole list in Python as well */ }
        // schemes.chain: anchors + the right bound in the order of the smallest path id's coordinate
        int32_t ref = rel[0].point[0].first;
        for (auto &kv : rel[0].point) ref = std::min(ref, kv.first);
        const size_t m = J.m = rel.size() + 1;
        std::vector<int> &order = J.order;
        order.assign(m, 0);
        for (size_t i = 0; i < m; i++) order[i] = (int)i;  // index rel.size() stands for the right bound
        size_t refcol = 0;
        for (size_t c = 0; c < k; c++)
            if (keys[c] == ref) refcol = c;
        std::vector<int64_t> refcoord(m);
        for (size_t i = 0; i < m; i++) refcoord[i] = i == rel.size() ? right[refcol] : coord_of(rel[i], ref);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return refcoord[(size_t)a] < refcoord[(size_t)b]; });
        std::vector<int64_t> &start = J.start, &length = J.length, &gain = J.gain;
        start.assign((m + 1) * k, 0);
        length.assign(m + 1, 0);
        gain.assign(m + 1, 0);
        J.link.assign(m + 1, 0);
        J.score.assign(m + 1, 0);
        for (size_t c = 0; c < k; c++) start[c] = left[c];
        for (size_t r = 1; r <= m; r++) {
            const int i = order[r - 1];
            if ((size_t)i == rel.size()) {
                for (size_t c = 0; c < k; c++) start[r * k + c] = right[c];
            } else {
                for (size_t c = 0; c < k; c++) start[r * k + c] = coord_of(rel[i], keys[c]);
                length[r] = rel[i].l;
                gain[r] = wscore * (rel[i].l * ((rel[i].n * (rel[i].n - 1)) / 2));
            }
        }
        J.need_dp = true;
    }
    PK_MARK(5)
    return 1;
}
// the error pick_build left in the job, as a Python exception (always returns -1)
static int pick_raise(PickJob &J) {
    if (J.err_bound >= 0 && J.bound_err[J.err_bound][0]) {
        PyObject **e = J.bound_err[J.err_bound];
        PyErr_Restore(e[0], e[1], e[2]);
        e[0] = e[1] = e[2] = nullptr;
    } else {
        PyErr_SetString(J.err == 2 ? PyExc_ValueError : PyExc_KeyError, J.err_msg.c_str());
    }
    return -1;
}
static void pick_account(const PickJob &J) {
    for (int i = 0; i < 6; i++) g_pk[i] += J.pk[i];
}
// 1: go on (run the recurrence if need_dp, then pick_finish); 0: the answer is the empty tuple; -1: error set
static int pick_prepare(Graph *g, PyObject *list, PickJob &J) {
    if (pick_parse(g, list, J) < 0) return -1;
    const int st = pick_build(g, J);
    pick_account(J);
    return st < 0 ? pick_raise(J) : st;
}

static PyObject *pick_finish(Graph *g, PickJob &J) {
    std::vector<Mum> &picked = J.picked;
    std::vector<Rel> &rel = J.rel;
    const Origin &origin = J.origin;
    const long long seedsize = J.seedsize;
    std::vector<std::pair<int, int64_t>> skipleft, skipright;  // (picked index, score relative to the split)
    int split = J.split;                                        // index into rel
    if (J.need_dp || J.dp_done) {
        const size_t m = J.m;
        const std::vector<int> &order = J.order;
        const std::vector<int64_t> &link = J.link, &score = J.score;

        // the right bound sorts last (anchors lie inside the bounds; it was appended last and the sort is stable): back-track
        // from the last row, like rem.chain
        std::vector<std::pair<int, int64_t>> chained;  // (rel index, score), first anchor of the chain first
        for (int64_t r = link[m]; r != 0; r = link[r]) {
            if ((size_t)order[r - 1] == rel.size()) continue;  // never the bound itself
            chained.emplace_back(order[r - 1], score[r]);
        }
        std::reverse(chained.begin(), chained.end());
        if (chained.empty()) return PyTuple_New(0);
        for (auto &c : chained)  // "largest": the last of the longest anchors of the chain
            if (split < 0 || rel[c.first].l >= rel[split].l) split = c.first;
        if (seedsize > 0) {
            bool after = false;
            int64_t at_split = 0;
            for (auto &c : chained) {
                if (c.first == split) { at_split = c.second; after = true; continue; }
                const int src = origin.find(rel[c.first]);
                if (src < 0) continue;
                if (picked[(size_t)src].l < seedsize) continue;
                (after ? skipright : skipleft).emplace_back(src, c.second - at_split);
            }
        }
    }
    const int split_src = origin.find(rel[split]);
    if (split_src < 0) { PyErr_SetString(PyExc_RuntimeError, "picked anchor lost its origin"); return nullptr; }
    PyObject *anchor = mum_object(picked[(size_t)split_src]);
    PyObject *lists[2];
    for (int side = 0; side < 2; side++) {
        auto &src = side == 0 ? skipleft : skipright;
        lists[side] = PyList_New((Py_ssize_t)src.size());
        for (size_t i = 0; i < src.size(); i++) {
            PyObject *mo = mum_object(picked[src[i].first]);
            PyList_SET_ITEM(lists[side], i, Py_BuildValue("(NL)", mo, (long long)src[i].second));
        }
    }
    return Py_BuildValue("(NNN)", anchor, lists[0], lists[1]);
}


static inline void pick_dp_host(PickJob &J) {
    rv_chain_dp(J.start.data(), J.length.data(), J.gain.data(), (long)(J.m + 1), (long)J.k, (int64_t)J.wpen, J.model, J.link.data(), J.score.data());
}

// pick(mums, nsamples, leftnode, rightnode, trim, maxmums, model, wscore, wpen, seedsize) -> () | (anchor, skipleft, skipright)
static PyObject *Graph_pick(Graph *g, PyObject *args) {
    PyObject *list;
    PickJob J;
    if (!PyArg_ParseTuple(args, "OlOOpliLLL", &list, &J.nsamples, &J.leftnode, &J.rightnode, &J.trim, &J.maxmums, &J.model, &J.wscore, &J.wpen, &J.seedsize)) return nullptr;
    const int st = pick_prepare(g, list, J);
    if (st < 0) return nullptr;
    if (st == 0) return PyTuple_New(0);
    if (J.need_dp) {
        pick_dp_host(J);
        J.need_dp = false;
        J.dp_done = true;
    }
    return pick_finish(g, J);
}

// export() -> (nodes, edges): nodes = [(key, attribute dict)], edges = [(u, v, attribute dict)] -- the attribute dicts are
// complete (offsets / aligned / extras; paths / ofrom / oto / extras) and freshly made, ready to be placed in a networkx graph
static PyObject *Graph_export(Graph *g, PyObject *) {
    PyObject *nodes = PyList_New(0), *edges = PyList_New(0);
    PyObject *s_offsets = PyUnicode_InternFromString("offsets"), *s_aligned = PyUnicode_InternFromString("aligned");
    PyObject *s_paths = PyUnicode_InternFromString("paths"), *s_ofrom = PyUnicode_InternFromString("ofrom"), *s_oto = PyUnicode_InternFromString("oto");
    PyObject *plus = PyUnicode_InternFromString("+"), *minus = PyUnicode_InternFromString("-");
    std::vector<PyObject *> keys(g->nodes->size(), nullptr);
    bool ok = true;
    for (size_t i = 0; ok && i < g->nodes->size(); i++) {
        const Node &n = (*g->nodes)[i];
        if (!n.alive) continue;
        PyObject *key;
        if (n.key) { key = n.key; Py_INCREF(key); }
        else key = PyObject_CallFunction(g->interval_cls, "LL", (long long)n.begin, (long long)n.end);
        if (!key) { ok = false; break; }
        keys[i] = key;
        PyObject *attrs = n.extra ? PyDict_Copy(n.extra) : PyDict_New();
        PyObject *offs = offsets_dict(n.offsets);
        PyDict_SetItem(attrs, s_offsets, offs);
        Py_DECREF(offs);
        if (n.aligned >= 0) {
            PyObject *al = PyLong_FromLong(n.aligned);
            PyDict_SetItem(attrs, s_aligned, al);
            Py_DECREF(al);
        }
        PyObject *row = PyTuple_Pack(2, key, attrs);
        PyList_Append(nodes, row);
        Py_DECREF(row);
        Py_DECREF(attrs);
    }
    // edges in the order networkx would hold them: by source node, then by insertion
    for (size_t i = 0; ok && i < g->nodes->size(); i++) {
        const Node &n = (*g->nodes)[i];
        if (!n.alive) continue;
        for (int eid : n.out) {
            const Edge &e = (*g->edges)[eid];
            if (!e.alive) continue;
            PyObject *attrs = e.extra ? PyDict_Copy(e.extra) : PyDict_New();
            PyObject *paths = PySet_New(nullptr);
            for (int32_t p : e.paths) { PyObject *x = PyLong_FromLong(p); PySet_Add(paths, x); Py_DECREF(x); }
            PyDict_SetItem(attrs, s_paths, paths);
            Py_DECREF(paths);
            PyDict_SetItem(attrs, s_ofrom, e.ofrom == '-' ? minus : plus);
            PyDict_SetItem(attrs, s_oto, e.oto == '-' ? minus : plus);
            PyObject *row = PyTuple_Pack(3, keys[e.u], keys[e.v], attrs);
            PyList_Append(edges, row);
            Py_DECREF(row);
            Py_DECREF(attrs);
        }
    }
    for (PyObject *k : keys) Py_XDECREF(k);
    for (PyObject *o : {s_offsets, s_aligned, s_paths, s_ofrom, s_oto, plus, minus}) Py_DECREF(o);
    if (!ok) { Py_DECREF(nodes); Py_DECREF(edges); return nullptr; }
    PyObject *ret = PyTuple_Pack(2, nodes, edges);
    Py_DECREF(nodes); Py_DECREF(edges);
    return ret;
}

static PyObject *Graph_stats(Graph *g, PyObject *) {
    long alive_n = 0, alive_e = 0;
    for (const Node &n : *g->nodes) alive_n += n.alive;
    for (const Edge &e : *g->edges) alive_e += e.alive;
    return Py_BuildValue("(llll)", alive_n, alive_e, (long)g->nodes->size(), (long)g->edges->size());
}

// set_picker(trim, maxmums, model, wscore, wpen, seedsize, maxdepth or -1): options of mumpicker()
static PyObject *Graph_set_picker(Graph *g, PyObject *args) {
    if (!PyArg_ParseTuple(args, "pliLLLl", &g->pk_trim, &g->pk_maxmums, &g->pk_model, &g->pk_wscore, &g->pk_wpen, &g->pk_seedsize, &g->pk_maxdepth)) return nullptr;
    Py_RETURN_NONE;
}

// ---- the chaining recurrences of a batch on the device -----------------------------------------------------------------------
// remcore is host code (the graph of the driver); the one numeric kernel of the mumpicker -- the O(m^2) recurrence -- goes
// through the C-ABI of libreveal_b200.so (rv_chain_batch, csrc/rv_chain.cu) when many lists are waiting (frontier batching) or a
// list is long.  The library is the one next to this module; without it or without a GPU (the driver running on the
// reference's CPU extension, as in the CPU tests) the host loop rv_chain_dp does the same arithmetic.  RV_REM_CHAIN=host|device
// forces one or the other.
struct ChainApi {
    bool tried = false, ready = false;
    std::string path;   // override (tests: the emulated kernels)
    void *lib = nullptr;
    rv_index *h = nullptr;
    decltype(&::rv_index_create) create = nullptr;
    decltype(&::rv_index_free) destroy = nullptr;
    decltype(&::rv_chain_batch) chain = nullptr;
    decltype(&::rv_last_error) last_error = nullptr;
    long long device_lists = 0, host_lists = 0, launches = 0;
};
static ChainApi g_chain;

extern "C" PyObject *PyInit_remcore(void);
static bool chain_device_ready() {
    ChainApi &c = g_chain;
    if (c.tried) return c.ready;
    c.tried = true;
    const char *mode = getenv("RV_REM_CHAIN");
    if (mode && mode[0] == 'h') return false;
    std::string path = c.path;
    if (path.empty()) {
        Dl_info info;
        if (!dladdr((void *)&PyInit_remcore, &info) || !info.dli_fname) return false;
        path = info.dli_fname;
        size_t slash = path.rfind('/');
        path = (slash == std::string::npos ? std::string(".") : path.substr(0, slash)) + "/libreveal_b200.so";
    }
    c.lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!c.lib) return false;
    c.create = (decltype(c.create))dlsym(c.lib, "rv_index_create");
    c.destroy = (decltype(c.destroy))dlsym(c.lib, "rv_index_free");
    c.chain = (decltype(c.chain))dlsym(c.lib, "rv_chain_batch");
    c.last_error = (decltype(c.last_error))dlsym(c.lib, "rv_last_error");
    if (!c.create || !c.destroy || !c.chain || !c.last_error) return false;
    if (c.create(&c.h, nullptr) != 0) return false;   // no GPU: the host loop it is
    c.ready = true;
    return true;
}

// ---- worker threads for the preparation of a frontier batch ---------------------------------------------------------------------
// The picks of a batch are independent and their preparation (trim_overlap, path coordinates, the rows of the recurrence, the
// short recurrences themselves) is plain C++ on private data plus reads of a graph that nobody writes meanwhile: the jobs are
// dealt out to a small pool.  The calling thread keeps the GIL and works along; the workers never touch a Python object.
// Threads: RV_REM_THREADS, else remcore.set_threads(n), else min(4, cores / 2); 1 = no pool.  The pool is made on first use and
// never torn down (no joins at interpreter exit); a forked child starts without one.
struct Pool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv;
    std::function<void(size_t)> fn;
    size_t n = 0;
    std::atomic<size_t> generation{0}, next{0};
    std::atomic<int> inside{0};          // workers between "may I" and "done"
    std::atomic<bool> accepting{false};  // the job description (fn, n) is valid and there may be jobs left
    static void relax() {
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
    }
    void run() {
        for (;;) {
            const size_t i = next.fetch_add(1);
            if (i >= n) break;
            fn(i);
        }
    }
    // The calls of a recursion come a few milliseconds apart and in pairs (preparation, then the short recurrences): a worker
    // that has just finished keeps looking for the next call for a while (a futex wake-up costs more than the work of a small
    // batch) and only then goes to sleep.
    void worker() {
        size_t seen = 0;
        for (;;) {
            bool got = false;
            const auto t0 = std::chrono::steady_clock::now();
            for (int spin = 0;; spin++) {
                if (generation.load(std::memory_order_acquire) != seen) { got = true; break; }
                relax();
                if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(600)) break;
            }
            if (!got) {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return generation.load(std::memory_order_acquire) != seen; });
            }
            seen = generation.load(std::memory_order_acquire);
            // A worker that wakes up late must not hold the caller back, nor read a job description that is being replaced:
            // it announces itself, then asks; the caller closes the call first and waits for the announced ones only.
            inside.fetch_add(1);
            if (accepting.load()) run();
            inside.fetch_sub(1);
        }
    }
    explicit Pool(int workers) {
        for (int i = 0; i < workers; i++) th.emplace_back([this] { worker(); });
    }
    void parallel_for(size_t count, const std::function<void(size_t)> &f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            fn = f;
            n = count;
            next.store(0);
            accepting.store(true);
            generation.fetch_add(1);
        }
        cv.notify_all();
        run();                     // returns when every job has been claimed
        accepting.store(false);
        while (inside.load() != 0) relax();  // the claimed ones are being finished: short jobs, no sleeping here
    }
};
static Pool *g_pool = nullptr;
static long long g_pooled_calls = 0, g_pooled_jobs = 0;
static int g_threads = 0;  // 0: not decided yet
static void pool_after_fork() { g_pool = nullptr; }  // the child has none of the parent's threads (the old object is left alone)
static int pool_threads() {
    if (g_threads > 0) return g_threads;
    int t = 0;
    if (const char *e = getenv("RV_REM_THREADS")) t = atoi(e);
    if (t <= 0) {
        const unsigned hw = std::thread::hardware_concurrency();
        t = (int)(hw / 2);
        if (t > 4) t = 4;  // (C2 on a 16-core host: picks 0.58 s alone, 0.42 s with 4 threads, 0.43 s with 8)
    }
    if (t < 1) t = 1;
    g_threads = t;
    return t;
}
static void for_each_job(size_t count, const std::function<void(size_t)> &f) {
    const int t = pool_threads();
    if (t <= 1 || count < 4) {
        for (size_t i = 0; i < count; i++) f(i);
        return;
    }
    if (!g_pool) {
        static bool hooked = false;
        if (!hooked) { pthread_atfork(nullptr, nullptr, pool_after_fork); hooked = true; }
        g_pool = new Pool(t - 1);
    }
    g_pooled_calls++;
    g_pooled_jobs += (long long)count;
    g_pool->parallel_for(count, f);
}

// runs the recurrences of the prepared jobs: the long ones together on the device when that pays, the rest on the host
static bool run_recurrences(std::vector<PickJob *> &jobs) {
    static const size_t SMALL = 48;          // rows below which a list is not worth a thread block
    static const double WORTH = 150000.0;    // sum of rows^2 from which one launch + two copies beat the host loop
    std::vector<PickJob *> big;
    double work = 0;
    for (PickJob *J : jobs) {
        if (!J->need_dp) continue;
        if (J->m + 1 >= SMALL) {
            big.push_back(J);
            work += (double)(J->m + 1) * (double)(J->m + 1);
        }
    }
    const char *mode = getenv("RV_REM_CHAIN");
    const bool force = mode && mode[0] == 'd';
    bool on_device = !big.empty() && (force || work >= WORTH) && chain_device_ready();
    if (on_device) {
        // every list in one call; lists may differ in gap model / penalty only across calls of one batch: group by (wpen, model)
        const int64_t wpen = big[0]->wpen;
        const int model = big[0]->model;
        std::vector<int64_t> row_off(1, 0), start_off(1, 0), start, length, gain;
        std::vector<int32_t> kk;
        for (PickJob *J : big) {
            if (J->wpen != wpen || J->model != model) { on_device = false; break; }
            row_off.push_back(row_off.back() + (int64_t)J->m + 1);
            start_off.push_back(start_off.back() + (int64_t)J->start.size());
            kk.push_back((int32_t)J->k);
            start.insert(start.end(), J->start.begin(), J->start.end());
            length.insert(length.end(), J->length.begin(), J->length.end());
            gain.insert(gain.end(), J->gain.begin(), J->gain.end());
        }
        if (on_device) {
            std::vector<int64_t> link((size_t)row_off.back()), score((size_t)row_off.back());
            int status;
            Py_BEGIN_ALLOW_THREADS;
            status = g_chain.chain(g_chain.h, (int32_t)big.size(), row_off.data(), start_off.data(), kk.data(), start.data(), length.data(), gain.data(),
                                   wpen, model, link.data(), score.data());
            Py_END_ALLOW_THREADS;
            if (status != 0) {
                PyErr_Format(PyExc_RuntimeError, "rv_chain_batch failed: %s", g_chain.last_error());
                return false;
            }
            for (size_t j = 0; j < big.size(); j++) {
                std::copy(link.begin() + row_off[j], link.begin() + row_off[j + 1], big[j]->link.begin());
                std::copy(score.begin() + row_off[j], score.begin() + row_off[j + 1], big[j]->score.begin());
                big[j]->need_dp = false;
                big[j]->dp_done = true;
            }
            g_chain.device_lists += (long long)big.size();
            g_chain.launches++;
        }
    }
    std::vector<PickJob *> rest;
    for (PickJob *J : jobs)
        if (J->need_dp) rest.push_back(J);
    for_each_job(rest.size(), [&](size_t i) {
        pick_dp_host(*rest[i]);
        rest[i]->need_dp = false;
        rest[i]->dp_done = true;
    });
    g_chain.host_lists += (long long)rest.size();
    return true;
}

// front part of the mumpicker callback: 0 -> *result is the answer; 1 -> J is prepared (recurrence, then pick_finish); -1 error
// defer_build: only the Python half of the preparation (pick_parse) is done here; the caller runs pick_build for its whole batch
static int mumpicker_front(Graph *g, PyObject *mums, PyObject *idx, int precomputed, PyObject **result, PickJob &J, bool defer_build = false) {
    *result = nullptr;
    const Py_ssize_t n = PyObject_Length(mums);
    if (n < 0) return -1;
    if (n == 0) { *result = PyTuple_New(0); return *result ? 0 : -1; }
    if (precomputed) {  // a chain handed down by the parent: split at its middle (schemes.py:346-351)
        const Py_ssize_t half = n / 2;
        PyObject *item = PySequence_GetItem(mums, half);
        if (!item) return -1;
        PyObject *mum = PySequence_GetItem(item, 0);
        Py_DECREF(item);
        if (!mum) return -1;
        PyObject *left = PySequence_GetSlice(mums, 0, half), *right = PySequence_GetSlice(mums, half + 1, n);
        if (!left || !right) { Py_DECREF(mum); Py_XDECREF(left); Py_XDECREF(right); return -1; }
        *result = Py_BuildValue("(NNN)", mum, left, right);
        return *result ? 0 : -1;
    }
    if (g->pk_maxdepth >= 0) {
        PyObject *d = PyObject_GetAttrString(idx, "depth");
        if (!d) return -1;
        const long depth = PyLong_AsLong(d);
        Py_DECREF(d);
        if (depth > g->pk_maxdepth) { *result = PyTuple_New(0); return *result ? 0 : -1; }
    }
    PyObject *ns = PyObject_GetAttrString(idx, "nsamples"), *ln = PyObject_GetAttrString(idx, "leftnode"), *rn = PyObject_GetAttrString(idx, "rightnode");
    int st = -1;
    if (ns && ln && rn) {
        J.nsamples = PyLong_AsLong(ns);
        J.leftnode = ln;    // kept alive by the index object for the duration of the call / the batch
        J.rightnode = rn;
        J.trim = g->pk_trim;
        J.maxmums = g->pk_maxmums;
        J.model = g->pk_model;
        J.wscore = g->pk_wscore;
        J.wpen = g->pk_wpen;
        J.seedsize = g->pk_seedsize;
        st = defer_build ? pick_parse(g, mums, J) : pick_prepare(g, mums, J);
        if (st == 0) { *result = PyTuple_New(0); if (!*result) st = -1; }
    }
    Py_XDECREF(ns);
    Py_XDECREF(ln);
    Py_XDECREF(rn);
    return st;
}

// mumpicker(mums, idx, precomputed=False, minlength=0): the callback index.align() expects (schemes.py:197-361, default options:
// splitchain="largest", a length threshold, no maxsize) without a Python frame in between -- Rem.graphmumpicker is the readable twin.
static PyObject *Graph_mumpicker(Graph *g, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"mums", "idx", "precomputed", "minlength", nullptr};
    PyObject *mums, *idx, *minlength = nullptr;
    int precomputed = 0;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "OO|pO", (char **)kwlist, &mums, &idx, &precomputed, &minlength)) return nullptr;
    PickJob J;
    PyObject *result = nullptr;
    const int st = mumpicker_front(g, mums, idx, precomputed, &result, J);
    if (st < 0) return nullptr;
    if (st == 0) return result;
    std::vector<PickJob *> one(1, &J);
    if (!run_recurrences(one)) return nullptr;
    return pick_finish(g, J);
}

// mumpicker_batch(entries, minlength=0) -> [pick, ...]: the mumpicker for every (mums, idx, precomputed) of a frontier batch of
// index.align().  The picks of different sub-indexes do not depend on each other (an anchor's path coordinates do not change
// while its text is unaligned), so all lists are prepared first and their chaining recurrences run together -- on the device
// (rv_chain_batch) when the batch holds enough work.
static PyObject *Graph_mumpicker_batch(Graph *g, PyObject *args, PyObject *kwds) {
    static const char *kwlist[] = {"entries", "minlength", nullptr};
    PyObject *entries, *minlength = nullptr;
    if (!PyArg_ParseTupleAndKeywords(args, kwds, "O|O", (char **)kwlist, &entries, &minlength)) return nullptr;
    PyObject *seq = PySequence_Fast(entries, "entries must be a sequence of (mums, idx, precomputed)");
    if (!seq) return nullptr;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    std::vector<PickJob> jobs((size_t)n);
    std::vector<PyObject *> results((size_t)n, nullptr);
    std::vector<int> state((size_t)n, 0);
    std::vector<PickJob *> pending;
    bool ok = true;
    for (Py_ssize_t i = 0; ok && i < n; i++) {
        PyObject *e = PySequence_Fast_GET_ITEM(seq, i);
        PyObject *mums, *idx;
        int precomputed = 0;
        if (!PyArg_ParseTuple(e, "OO|p", &mums, &idx, &precomputed)) { ok = false; break; }
        state[(size_t)i] = mumpicker_front(g, mums, idx, precomputed, &results[(size_t)i], jobs[(size_t)i], true);
        if (state[(size_t)i] < 0) ok = false;
        else if (state[(size_t)i] == 1) pending.push_back(&jobs[(size_t)i]);
    }
    if (ok) {  // the C++ half of every preparation, side by side
        std::vector<int> built(pending.size(), 0);
        for_each_job(pending.size(), [&](size_t j) { built[j] = pick_build(g, *pending[j]); });
        std::vector<PickJob *> go;
        size_t j = 0;
        for (Py_ssize_t i = 0; i < n; i++) {
            if (state[(size_t)i] != 1) continue;
            PickJob &J = jobs[(size_t)i];
            pick_account(J);
            const int st = built[j++];
            if (st < 0) {
                if (ok) pick_raise(J);  // the first failure is the one reported
                ok = false;
            } else if (st == 0) {
                state[(size_t)i] = 0;
                results[(size_t)i] = PyTuple_New(0);
                if (!results[(size_t)i]) ok = false;
            } else {
                go.push_back(&J);
            }
        }
        pending.swap(go);
    }
    if (ok) ok = run_recurrences(pending);
    PyObject *out = ok ? PyList_New(n) : nullptr;
    for (Py_ssize_t i = 0; ok && i < n; i++) {
        if (state[(size_t)i] == 1) results[(size_t)i] = pick_finish(g, jobs[(size_t)i]);
        if (!results[(size_t)i]) { ok = false; break; }
        PyList_SET_ITEM(out, i, results[(size_t)i]);
        results[(size_t)i] = nullptr;
    }
    for (PyObject *r : results) Py_XDECREF(r);
    Py_DECREF(seq);
    if (!ok) { Py_XDECREF(out); return nullptr; }
    return out;
}

// graphalign_cb(idx, mum): the second callback of index.align() (rem.py:320-382) on this graph
static PyObject *Graph_graphalign(Graph *g, PyObject *args);
static PyObject *Graph_graphalign_cb(Graph *g, PyObject *args) {
    PyObject *idx, *mum;
    if (!PyArg_ParseTuple(args, "OO", &idx, &mum)) return nullptr;
    PyObject *l, *spd;
    if (!PyTuple_Check(mum) || PyTuple_GET_SIZE(mum) < 3) { PyErr_SetString(PyExc_TypeError, "anchor: (l, n, positions) expected"); return nullptr; }
    l = PyTuple_GET_ITEM(mum, 0);
    spd = PyTuple_GET_ITEM(mum, 2);
    const Py_ssize_t k = PyObject_Length(spd);
    if (k < 0) return nullptr;
    PyObject *positions = PyList_New(k);
    for (Py_ssize_t i = 0; i < k; i++) {
        PyObject *pr = PySequence_GetItem(spd, i);
        PyObject *pos = pr ? PySequence_GetItem(pr, 1) : nullptr;
        Py_XDECREF(pr);
        if (!pos) { Py_DECREF(positions); return nullptr; }
        PyList_SET_ITEM(positions, i, pos);
    }
    PyObject *nodes = PyObject_GetAttrString(idx, "nodes"), *ln = PyObject_GetAttrString(idx, "leftnode"), *rn = PyObject_GetAttrString(idx, "rightnode");
    PyObject *ret = nullptr;
    if (nodes && ln && rn) {
        PyObject *a = Py_BuildValue("(OOOOO)", nodes, ln, rn, l, positions);
        if (a) {
            ret = Graph_graphalign(g, a);
            Py_DECREF(a);
        }
    }
    Py_XDECREF(nodes);
    Py_XDECREF(ln);
    Py_XDECREF(rn);
    Py_DECREF(positions);
    return ret;
}

static PyMethodDef Graph_methods[] = {
    {"set_picker", (PyCFunction)Graph_set_picker, METH_VARARGS, "set_picker(trim, maxmums, model, wscore, wpen, seedsize, maxdepth or -1): options of mumpicker()"},
    {"mumpicker", (PyCFunction)(void (*)(void))Graph_mumpicker, METH_VARARGS | METH_KEYWORDS,
     "mumpicker(mums, idx, precomputed=False, minlength=0) -> () | (anchor, skipleft, skipright): the callback of index.align(), default flow"},
    {"mumpicker_batch", (PyCFunction)(void (*)(void))Graph_mumpicker_batch, METH_VARARGS | METH_KEYWORDS,
     "mumpicker_batch([(mums, idx, precomputed), ...], minlength=0) -> [pick, ...]: the mumpicker for a whole frontier batch of index.align(); "
     "the chaining recurrences of the batch run together (on the device when that pays)"},
    {"graphalign_cb", (PyCFunction)Graph_graphalign_cb, METH_VARARGS, "graphalign_cb(idx, mum): the graphalign callback of index.align() on this graph"},
    {"add_node", (PyCFunction)Graph_add_node, METH_VARARGS, "add_node(key, aligned or None, offsets, extra or None)"},
    {"add_edge", (PyCFunction)Graph_add_edge, METH_VARARGS, "add_edge(u, v, ofrom, oto, paths, extra or None)"},
    {"coords", (PyCFunction)Graph_coords, METH_O, "coords(pos) -> ((path id, coordinate), ...) of the real paths through index position pos"},
    {"node_offsets", (PyCFunction)Graph_node_offsets, METH_O, "node_offsets(node) -> {path id: offset}"},
    {"graphalign", (PyCFunction)Graph_graphalign, METH_VARARGS,
     "graphalign(nodes, leftnode, rightnode, l, positions) -> (leading, trailing, matching, rest, merged, newleft, newright)"},
    {"pick", (PyCFunction)Graph_pick, METH_VARARGS,
     "pick(mums, nsamples, leftnode, rightnode, trim, maxmums, model, wscore, wpen, seedsize) -> () | (anchor, skipleft, skipright)"},
    {"export", (PyCFunction)Graph_export, METH_NOARGS, "export() -> (node rows, edge rows) of the alive graph"},
    {"stats", (PyCFunction)Graph_stats, METH_NOARGS, "(alive nodes, alive edges, node slots, edge slots)"},
    {nullptr, nullptr, 0, nullptr}};

static PyTypeObject GraphType = {PyVarObject_HEAD_INIT(nullptr, 0)};

static PyObject *mod_pick_phases(PyObject *, PyObject *) {
    return Py_BuildValue("{s:d,s:d,s:d,s:d,s:d,s:d}", "parse", g_pk[0], "filter_trim", g_pk[1], "sort", g_pk[2], "lookup", g_pk[3], "origin", g_pk[4], "rows", g_pk[5]);
}
static PyObject *mod_chain_stats(PyObject *, PyObject *) {
    return Py_BuildValue("{s:O,s:L,s:L,s:L,s:i,s:L,s:L}", "device", g_chain.ready ? Py_True : Py_False, "device_lists", g_chain.device_lists, "host_lists",
                         g_chain.host_lists, "launches", g_chain.launches, "threads", pool_threads(), "pooled_calls", g_pooled_calls, "pooled_jobs", g_pooled_jobs);
}
// set_threads(n) -> previous setting: threads that prepare the picks of a frontier batch (1: the calling thread alone; 0: decide
// again from RV_REM_THREADS / the core count).  A pool that exists keeps its size.
static PyObject *mod_set_threads(PyObject *, PyObject *args) {
    int n = 0;
    if (!PyArg_ParseTuple(args, "i", &n)) return nullptr;
    const int before = g_threads;
    g_threads = n < 0 ? 0 : (n > 64 ? 64 : n);
    return PyLong_FromLong(before);
}

static PyObject *mod_set_chain_library(PyObject *, PyObject *args) {
    const char *path;
    if (!PyArg_ParseTuple(args, "s", &path)) return nullptr;
    const char *e = getenv("REVEAL_B200_TEST_HOOKS");
    if (!e || e[0] != '1') {
        PyErr_SetString(PyExc_RuntimeError, "remcore._set_chain_library is a test hook (REVEAL_B200_TEST_HOOKS=1)");
        return nullptr;
    }
    if (g_chain.h && g_chain.destroy) g_chain.destroy(g_chain.h);
    g_chain = ChainApi();
    g_chain.path = path;
    Py_RETURN_NONE;
}
static PyMethodDef module_methods[] = {
    {"pick_phases", mod_pick_phases, METH_NOARGS, "seconds this process spent in the phases of the mumpicker's preparation (diagnostics)"},
    {"chain_stats", mod_chain_stats, METH_NOARGS, "where the chaining recurrences of this process ran: lists on the device / on the host, device launches"},
    {"set_threads", mod_set_threads, METH_VARARGS, "set_threads(n) -> previous: threads preparing the picks of a frontier batch (1 = none besides the caller, 0 = default)"},
    {"_set_chain_library", mod_set_chain_library, METH_VARARGS, "test hook: the C-ABI library rv_chain_batch is taken from (the emulated kernels)"},
    {nullptr, nullptr, 0, nullptr}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "remcore",
                                       "Alignment graph of the REM driver during the recursion (see reveal_b200/rem.py).", -1, module_methods};

}  // namespace

extern "C" __attribute__((visibility("default"))) PyObject *PyInit_remcore(void) {
    GraphType.tp_name = "remcore.Graph";
    GraphType.tp_basicsize = sizeof(Graph);
    GraphType.tp_flags = Py_TPFLAGS_DEFAULT;
    GraphType.tp_doc = "Graph(multi, interval_class, real_path_flags, path_lengths)";
    GraphType.tp_new = Graph_new;
    GraphType.tp_init = (initproc)Graph_init;
    GraphType.tp_dealloc = (destructor)Graph_dealloc;
    GraphType.tp_methods = Graph_methods;
    if (PyType_Ready(&GraphType) < 0) return nullptr;
    PyObject *m = PyModule_Create(&moduledef);
    if (!m) return nullptr;
    Py_INCREF(&GraphType);
    PyModule_AddObject(m, "Graph", (PyObject *)&GraphType);
    return m;
}
